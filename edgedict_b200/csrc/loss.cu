// loss.cu -- RNN-Transducer loss for sm_100a.
//
// Replaces the reference's GPU path  warp-transducer/include/detail/gpu_rnnt.h:82-215
// (memset + reduce_max + reduce_exp + alphas + betas + grad + D2H, three host syncs) with
//
//   1. rnnt_denom_kernel   one warp per (b,t,u) row: one streaming 128-bit pass over the logits,
//                          writes -logsumexp and the two log-probs the lattice needs
//                          (blank, label[u]) as compact [B,T,U] arrays        (reads 1*s*N bytes)
//   2. rnnt_lattice_kernel alpha and beta wavefronts concurrently (grid = B x 2), each touching
//                          2 floats per cell instead of V-strided gathers      (latency bound)
//   3. rnnt_grad_kernel    one warp per row: re-reads the logits, writes d loss / d logits
//                          (fp32 or bf16, optionally in place), zero on padded cells, scaled
//                          by the upstream gradient on the device               (2*s*N bytes)
//
// No host synchronisation anywhere except in the warp-transducer compatible entry point
// compute_rnnt_loss(), whose contract returns costs in HOST memory (include/rnnt.h).
//
// Arithmetic follows include/detail/gpu_rnnt_kernel.h:5-179 and rnnt_helper.h:17-24.
#include "common.cuh"
#include "../../include/rnnt.h"
#include "../../include/edgedict_b200.h"

namespace {

template <typename T> struct M;
template <> struct M<float> {
    static __device__ __forceinline__ float exp_fast(float x) { return fast_exp(x); }
    static __device__ __forceinline__ float exp_acc(float x) { return expf(x); }
    static __device__ __forceinline__ float log_acc(float x) { return logf(x); }
    static __device__ __forceinline__ float log1p_acc(float x) { return log1pf(x); }
    static __device__ __forceinline__ float ninf() { return -INFINITY; }
};
template <> struct M<double> {
    static __device__ __forceinline__ double exp_fast(double x) { return exp(x); }
    static __device__ __forceinline__ double exp_acc(double x) { return exp(x); }
    static __device__ __forceinline__ double log_acc(double x) { return log(x); }
    static __device__ __forceinline__ double log1p_acc(double x) { return log1p(x); }
    static __device__ __forceinline__ double ninf() { return -(double)INFINITY; }
};

template <typename T>
__device__ __forceinline__ T lse2(T a, T b) {   // rnnt_helper.h:17-24
    if (a == M<T>::ninf()) return b;
    if (b == M<T>::ninf()) return a;
    return (a > b) ? M<T>::log1p_acc(M<T>::exp_acc(b - a)) + a : M<T>::log1p_acc(M<T>::exp_acc(a - b)) + b;
}

// 4-wide row access; VEC=true requires V % 4 == 0 (row starts are then 16-byte aligned for fp32)
template <typename T, bool VEC> struct Row4 {
    static __device__ __forceinline__ void load(const T* row, int v, int V, T (&x)[4]) {
        if (VEC) {
            if (sizeof(T) == 4) {
                float4 q;
                asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(row + v));
                x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
            } else {
                const double2* p = reinterpret_cast<const double2*>(row + v);
                double2 a = p[0], b = p[1];
                x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = (v + i < V) ? row[v + i] : M<T>::ninf();
        }
    }
};

// bf16 logits (written by the fused joint GEMM epilogue): 4 values per 8-byte load
template <bool VEC> struct Row4<__nv_bfloat16, VEC> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* row, int v, int V, float (&x)[4]) {
        if (VEC) {
            const uint2 q = *reinterpret_cast<const uint2*>(row + v);
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
            const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
            x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = (v + i < V) ? __bfloat162float(row[v + i]) : -INFINITY;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// 1. denominators + (blank, label) gather
// ---------------------------------------------------------------------------------------------
template <typename T, bool VEC, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
rnnt_denom_kernel(const T* __restrict__ logits, const int* __restrict__ labels,
                  const int* __restrict__ xlen, const int* __restrict__ ylen,
                  T* __restrict__ denom, T* __restrict__ lpb, T* __restrict__ lpl,
                  int B, int maxT, int maxU, int V, int blank) {
    const int lane = threadIdx.x & 31;
    const long ncells = (long)B * maxT * maxU;
    const long wstride = (long)gridDim.x * WARPS;
    for (long cell = (long)blockIdx.x * WARPS + (threadIdx.x >> 5); cell < ncells; cell += wstride) {
        const int u = (int)(cell % maxU);
        const long bt = cell / maxU;
        const int t = (int)(bt % maxT);
        const int b = (int)(bt / maxT);
        const int Tn = xlen[b], Un = ylen[b] + 1;
        if (t >= Tn || u >= Un) continue;                // padded cell: never read
        const T* row = logits + cell * (long)V;
        const int lab = (u < Un - 1) ? labels[b * (maxU - 1) + u] : -1;
        // online max / sum-exp over this lane's 4-wide chunks
        T m = M<T>::ninf(), s = 0, xb = 0, xl = 0;
        for (int v = lane * 4; v < V; v += 128) {
            T x[4];
            Row4<T, VEC>::load(row, v, V, x);
            T cm = fmax(fmax(x[0], x[1]), fmax(x[2], x[3]));
            T nm = fmax(m, cm);
            T acc = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc += M<T>::exp_fast(x[i] - nm);       // exp(-inf) = 0 for the masked tail
                if (v + i == blank) xb = x[i];
                if (v + i == lab) xl = x[i];
            }
            s = s * M<T>::exp_fast(m - nm) + acc;       // m = -inf on first chunk => s*0
            m = nm;
        }
        T gm = warp_max(m);
        s *= M<T>::exp_fast(m - gm);
        s = warp_sum(s);
        const T d = -gm - M<T>::log_acc(s);
        // the lane that saw the blank / label column owns the value: reduce by sum of one-hot
        xb = warp_sum(xb);
        xl = warp_sum(xl);
        if (lane == 0) {
            denom[cell] = d;
            lpb[cell] = d + xb;
            lpl[cell] = d + xl;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 2. alpha / beta wavefronts.  grid = (B, 2): y == 0 -> alphas, y == 1 -> betas.
//    blockDim.x = maxU rounded up to a warp.  Thread u walks its own column t = n - u.
// ---------------------------------------------------------------------------------------------
template <typename T, int PF>
__global__ void rnnt_lattice_kernel(const T* __restrict__ lpb, const T* __restrict__ lpl,
                                    const int* __restrict__ xlen, const int* __restrict__ ylen,
                                    T* __restrict__ alphas, T* __restrict__ betas,
                                    T* __restrict__ ll_fwd, T* __restrict__ ll_bwd,
                                    int maxT, int maxU, int do_beta) {
    extern __shared__ unsigned char sm_raw[];
    T* sh = reinterpret_cast<T*>(sm_raw);               // [2][blockDim.x]
    const int b = blockIdx.x, u = threadIdx.x;
    const int Tn = xlen[b], Un = ylen[b] + 1;
    const long base = (long)b * maxT * maxU;
    const T* pb = lpb + base;
    const T* pl = lpl + base;
    const int ND = Tn + Un - 1;                          // number of anti-diagonals
    const bool act = u < Un;
    const int W = blockDim.x;
    if (blockIdx.y == 0) {
        T* al = alphas + base;
        T self = M<T>::ninf();                           // alpha(t-1,u)
        // register prefetch ring: values for iterations n0 .. n0+PF-1
        T cb[PF], cl[PF], nb[PF], nl[PF];
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            int t = i - u;
            cb[i] = (act && t >= 1 && t < Tn) ? pb[(long)(t - 1) * maxU + u] : T(0);
            cl[i] = (act && u >= 1 && t >= 0 && t < Tn) ? pl[(long)t * maxU + u - 1] : T(0);
        }
        for (int n0 = 0; n0 < ND; n0 += PF) {
#pragma unroll
            for (int i = 0; i < PF; ++i) {               // issue next block's loads early
                int t = n0 + PF + i - u;
                nb[i] = (act && t >= 1 && t < Tn) ? pb[(long)(t - 1) * maxU + u] : T(0);
                nl[i] = (act && u >= 1 && t >= 0 && t < Tn) ? pl[(long)t * maxU + u - 1] : T(0);
            }
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                const int n = n0 + i;
                if (n < ND) {
                    const int t = n - u;
                    if (act && t >= 0 && t < Tn) {
                        T a;
                        if (n == 0) a = 0;
                        else {
                            T stay = (t > 0) ? self + cb[i] : M<T>::ninf();
                            T emit = (u > 0) ? sh[((n - 1) & 1) * W + u - 1] + cl[i] : M<T>::ninf();
                            a = lse2(emit, stay);
                        }
                        al[(long)t * maxU + u] = a;
                        self = a;
                        sh[(n & 1) * W + u] = a;
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int i = 0; i < PF; ++i) { cb[i] = nb[i]; cl[i] = nl[i]; }
        }
        if (u == Un - 1) ll_fwd[b] = self + pb[(long)(Tn - 1) * maxU + Un - 1];
    } else {
        if (!do_beta) return;
        T* be = betas + base;
        T self = M<T>::ninf();                           // beta(t+1,u)
        T cb[PF], cl[PF], nb[PF], nl[PF];
#pragma unroll
        for (int i = 0; i < PF; ++i) {
            int t = (ND - 1 - i) - u;
            bool ok = act && t >= 0 && t < Tn;
            cb[i] = ok ? pb[(long)t * maxU + u] : T(0);
            cl[i] = (ok && u < Un - 1) ? pl[(long)t * maxU + u] : T(0);
        }
        for (int n0 = 0; n0 < ND; n0 += PF) {
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                int t = (ND - 1 - (n0 + PF + i)) - u;
                bool ok = act && t >= 0 && t < Tn;
                nb[i] = ok ? pb[(long)t * maxU + u] : T(0);
                nl[i] = (ok && u < Un - 1) ? pl[(long)t * maxU + u] : T(0);
            }
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                const int k = n0 + i;                    // k-th step, diagonal n = ND-1-k
                if (k < ND) {
                    const int n = ND - 1 - k;
                    const int t = n - u;
                    if (act && t >= 0 && t < Tn) {
                        T v;
                        if (k == 0) v = cb[i];           // (T-1, U-1): log p(blank)
                        else {
                            T stay = (t < Tn - 1) ? self + cb[i] : M<T>::ninf();
                            T emit = (u < Un - 1) ? sh[((k - 1) & 1) * W + u + 1] + cl[i] : M<T>::ninf();
                            v = lse2(emit, stay);
                        }
                        be[(long)t * maxU + u] = v;
                        self = v;
                        sh[(k & 1) * W + u] = v;
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int i = 0; i < PF; ++i) { cb[i] = nb[i]; cl[i] = nl[i]; }
        }
        if (u == 0) ll_bwd[b] = self;
    }
}

// ---------------------------------------------------------------------------------------------
// 3. gradient wrt logits (gpu_rnnt_kernel.h:143-179), fused zero-fill of padded cells, upstream
//    gradient and 1/B scaling applied on the device.
// ---------------------------------------------------------------------------------------------
template <typename TO> struct Store4;
template <> struct Store4<float> {
    static __device__ __forceinline__ void st(float* p, const float (&g)[4]) {
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "f"(g[0]), "f"(g[1]), "f"(g[2]), "f"(g[3]) : "memory");
    }
};
template <> struct Store4<double> {
    static __device__ __forceinline__ void st(double* p, const double (&g)[4]) {
        reinterpret_cast<double2*>(p)[0] = make_double2(g[0], g[1]);
        reinterpret_cast<double2*>(p)[1] = make_double2(g[2], g[3]);
    }
};
template <> struct Store4<__nv_bfloat16> {
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&g)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(g[0], g[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(g[2], g[3]);
        uint2 q;
        q.x = *reinterpret_cast<uint32_t*>(&a);
        q.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = q;
    }
};

template <typename T, typename TO, bool VEC, int WARPS, typename TI = T>
__global__ void __launch_bounds__(WARPS * 32)
rnnt_grad_kernel(const TI* logits, TO* grads, const int* __restrict__ labels,
                 const int* __restrict__ xlen, const int* __restrict__ ylen,
                 const T* __restrict__ denom, const T* __restrict__ alphas,
                 const T* __restrict__ betas, const T* __restrict__ ll_fwd,
                 const T* __restrict__ gscale, int gscale_per_batch, T hscale,
                 int B, int maxT, int maxU, int V, int blank) {
    const int lane = threadIdx.x & 31;
    const long ncells = (long)B * maxT * maxU;
    const long wstride = (long)gridDim.x * WARPS;
    for (long cell = (long)blockIdx.x * WARPS + (threadIdx.x >> 5); cell < ncells; cell += wstride) {
        const int u = (int)(cell % maxU);
        const long bt = cell / maxU;
        const int t = (int)(bt % maxT);
        const int b = (int)(bt / maxT);
        const int Tn = xlen[b], Un = ylen[b] + 1;
        const TI* row = logits + cell * (long)V;
        TO* orow = grads + cell * (long)V;
        if (t >= Tn || u >= Un) {                         // padded: zero, logits never read
            for (int v = lane * 4; v < V; v += 128) {
                if (VEC) {
                    const T z[4] = {0, 0, 0, 0};
                    Store4<TO>::st(orow + v, z);
                } else {
                    for (int i = 0; i < 4 && v + i < V; ++i) orow[v + i] = TO(T(0));
                }
            }
            continue;
        }
        const T sc = hscale * (gscale ? gscale[gscale_per_batch ? b : 0] : T(1));
        const T a = alphas[cell], bt_ = betas[cell], ll = ll_fwd[b], d = denom[cell];
        const int lab = (u < Un - 1) ? labels[b * (maxU - 1) + u] : -1;
        // scalar pieces shared by the row
        const T c_all = a + bt_ - ll + d;                 // exp(c_all + x_v) = exp(a+b+logp-ll)
        T c_blank = M<T>::ninf();                         // log-factor subtracted at v == blank
        if (t < Tn - 1) c_blank = a - ll + d + betas[cell + maxU];
        else if (u == Un - 1) c_blank = a - ll + d;
        const T c_lab = (lab >= 0) ? a - ll + d + betas[cell + 1] : M<T>::ninf();
        for (int v = lane * 4; v < V; v += 128) {
            T x[4];
            Row4<TI, VEC>::load(row, v, V, x);
            T g[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                T gr = M<T>::exp_fast(c_all + x[i]);
                if (v + i == blank) gr -= M<T>::exp_acc(c_blank + x[i]);
                if (v + i == lab) gr -= M<T>::exp_acc(c_lab + x[i]);
                g[i] = gr * sc;
            }
            if (VEC) {
                Store4<TO>::st(orow + v, g);
            } else {
                for (int i = 0; i < 4 && v + i < V; ++i) orow[v + i] = TO(g[i]);
            }
        }
    }
}

// bf16 logits -> bf16 gradients (in place allowed), the fused-joint path: 8 values per lane and iteration
// (one 16-byte load, one 16-byte store); the rare blank / label corrections are applied outside the unrolled
// exponentials.  Same arithmetic as rnnt_grad_kernel (gpu_rnnt_kernel.h:143-179).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
rnnt_grad_bf16x8_kernel(const __nv_bfloat16* logits, __nv_bfloat16* grads, const int* __restrict__ labels,
                        const int* __restrict__ xlen, const int* __restrict__ ylen, const float* __restrict__ denom,
                        const float* __restrict__ alphas, const float* __restrict__ betas,
                        const float* __restrict__ ll_fwd, const float* __restrict__ gscale, int gscale_per_batch,
                        float hscale, int B, int maxT, int maxU, int V, int blank) {
    const int lane = threadIdx.x & 31;
    const long ncells = (long)B * maxT * maxU;
    const long wstride = (long)gridDim.x * WARPS;
    for (long cell = (long)blockIdx.x * WARPS + (threadIdx.x >> 5); cell < ncells; cell += wstride) {
        const int u = (int)(cell % maxU);
        const long bt = cell / maxU;
        const int t = (int)(bt % maxT);
        const int b = (int)(bt / maxT);
        const int Tn = xlen[b], Un = ylen[b] + 1;
        const __nv_bfloat16* row = logits + cell * (long)V;
        __nv_bfloat16* orow = grads + cell * (long)V;
        if (t >= Tn || u >= Un) {
            for (int v = lane * 8; v < V; v += 256) *reinterpret_cast<uint4*>(orow + v) = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        const float sc = hscale * (gscale ? gscale[gscale_per_batch ? b : 0] : 1.f);
        const float a = alphas[cell], bt_ = betas[cell], ll = ll_fwd[b], d = denom[cell];
        const int lab = (u < Un - 1) ? labels[b * (maxU - 1) + u] : -1;
        const float c_all = a + bt_ - ll + d;
        float c_blank = -INFINITY;
        if (t < Tn - 1) c_blank = a - ll + d + betas[cell + maxU];
        else if (u == Un - 1) c_blank = a - ll + d;
        const float c_lab = (lab >= 0) ? a - ll + d + betas[cell + 1] : -INFINITY;
        constexpr float LOG2E = 1.4426950408889634f;
        const float c2 = c_all * LOG2E;
        // logits and grads may be the same buffer: all four 16-byte loads of a group are issued before the first
        // store, otherwise the possible aliasing serialises load -> store -> load and one row costs four DRAM
        // round trips
        for (int v0 = lane * 8; v0 < V; v0 += 1024) {
            uint4 qq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (v0 + k * 256 < V) qq[k] = *reinterpret_cast<const uint4*>(row + v0 + k * 256);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = v0 + k * 256;
                if (v >= V) break;
                const uint32_t w[4] = {qq[k].x, qq[k].y, qq[k].z, qq[k].w};
                float x[8], g[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
                    x[2 * i] = f.x; x[2 * i + 1] = f.y;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = fast_ex2(fmaf(x[i], LOG2E, c2));
                const unsigned ib = (unsigned)(blank - v), il = (unsigned)(lab - v);
                if (ib < 8u || il < 8u) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if ((unsigned)i == ib) g[i] -= expf(c_blank + x[i]);
                        if ((unsigned)i == il) g[i] -= expf(c_lab + x[i]);
                    }
                }
                uint32_t ow[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    __nv_bfloat162 pk = __floats2bfloat162_rn(g[2 * i] * sc, g[2 * i + 1] * sc);
                    ow[i] = *reinterpret_cast<uint32_t*>(&pk);
                }
                *reinterpret_cast<uint4*>(orow + v) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
        }
    }
}

template <typename T>
struct Workspace {
    T *denom, *lpb, *lpl, *alphas, *betas, *ll_fwd, *ll_bwd;
    static size_t bytes(int B, int maxT, int maxU) {
        return sizeof(T) * ((size_t)B * maxT * maxU * 5 + 2 * (size_t)B);
    }
    Workspace(void* ws, int B, int maxT, int maxU) {
        const size_t n = (size_t)B * maxT * maxU;
        T* p = reinterpret_cast<T*>(ws);
        denom = p; lpb = p + n; lpl = p + 2 * n; alphas = p + 3 * n; betas = p + 4 * n;
        ll_fwd = p + 5 * n; ll_bwd = ll_fwd + B;
    }
};

inline int row_grid(long ncells, int warps) {
    long blocks = (ncells + warps - 1) / warps;
    long cap = (long)eb_num_sms() * 16;                  // grid-stride above 16 CTAs/SM
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

template <typename T>
int loss_fwd(const T* logits, const int* labels, const int* xlen, const int* ylen, int B, int maxT,
             int maxU, int V, int blank, void* ws, int need_beta, cudaStream_t st) {
    if (!logits || (!labels && maxU > 1) || !xlen || !ylen || !ws || B <= 0 || maxT <= 0 || maxU <= 0 || V <= 0 ||
        blank < 0 || blank >= V || maxU > 1024)
        return EB_ERR_INVALID;
    Workspace<T> w(ws, B, maxT, maxU);
    const long ncells = (long)B * maxT * maxU;
    constexpr int WARPS = 8;
    const bool vec = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
    if (vec)
        rnnt_denom_kernel<T, true, WARPS><<<row_grid(ncells, WARPS), WARPS * 32, 0, st>>>(
            logits, labels, xlen, ylen, w.denom, w.lpb, w.lpl, B, maxT, maxU, V, blank);
    else
        rnnt_denom_kernel<T, false, WARPS><<<row_grid(ncells, WARPS), WARPS * 32, 0, st>>>(
            logits, labels, xlen, ylen, w.denom, w.lpb, w.lpl, B, maxT, maxU, V, blank);
    EB_CHECK_LAUNCH();
    const int threads = ((maxU + 31) / 32) * 32;
    rnnt_lattice_kernel<T, 8><<<dim3(B, 2), threads, 2 * threads * sizeof(T), st>>>(
        w.lpb, w.lpl, xlen, ylen, w.alphas, w.betas, w.ll_fwd, w.ll_bwd, maxT, maxU, need_beta);
    EB_CHECK_LAUNCH();
    return EB_OK;
}

template <typename T, typename TO>
int loss_bwd(const T* logits, TO* grads, const int* labels, const int* xlen, const int* ylen, int B,
             int maxT, int maxU, int V, int blank, void* ws, const T* gscale, int per_batch,
             T hscale, cudaStream_t st) {
    if (!logits || !grads || !ws) return EB_ERR_INVALID;
    Workspace<T> w(ws, B, maxT, maxU);
    const long ncells = (long)B * maxT * maxU;
    constexpr int WARPS = 8;
    const bool vec = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(grads) & 15) == 0);
    if (vec)
        rnnt_grad_kernel<T, TO, true, WARPS><<<row_grid(ncells, WARPS), WARPS * 32, 0, st>>>(
            logits, grads, labels, xlen, ylen, w.denom, w.alphas, w.betas, w.ll_fwd, gscale,
            per_batch, hscale, B, maxT, maxU, V, blank);
    else
        rnnt_grad_kernel<T, TO, false, WARPS><<<row_grid(ncells, WARPS), WARPS * 32, 0, st>>>(
            logits, grads, labels, xlen, ylen, w.denom, w.alphas, w.betas, w.ll_fwd, gscale,
            per_batch, hscale, B, maxT, maxU, V, blank);
    EB_CHECK_LAUNCH();
    return EB_OK;
}

template <typename T>
__global__ void neg_copy_kernel(const T* ll, T* costs, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) costs[i] = -ll[i];
}

template <typename T>
rnntStatus_t compat_entry(const T* acts, T* grads, const int* labels, const int* label_lengths,
                          const int* input_lengths, int V, int B, T* costs, void* workspace,
                          rnntOptions o) {
    // argument validation mirrors src/rnnt_entrypoint.cpp:49-59
    if (!acts || (!labels && o.maxU > 1) || !label_lengths || !input_lengths || !costs || !workspace || V <= 0 ||
        B <= 0 || o.maxT <= 0 || o.maxU <= 0)
        return RNNT_STATUS_INVALID_VALUE;
    if (o.loc != RNNT_GPU) {
        // The reference prints a diagnostic when the requested location is not compiled in
        // (rnnt_entrypoint.cpp:86-88).  This library is GPU-only by design: no CPU fallback.
        fprintf(stderr, "CPU execution requested, but edgedict_b200 is a GPU-only (sm_100a) build\n");
        return RNNT_STATUS_EXECUTION_FAILED;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(o.stream);
    int rc = loss_fwd<T>(acts, labels, input_lengths, label_lengths, B, o.maxT, o.maxU, V,
                         o.blank_label, workspace, grads != nullptr, st);
    if (rc == EB_ERR_INVALID) return RNNT_STATUS_INVALID_VALUE;
    if (rc != EB_OK) return RNNT_STATUS_EXECUTION_FAILED;
    if (grads) {
        rc = loss_bwd<T, T>(acts, grads, labels, input_lengths, label_lengths, B, o.maxT, o.maxU, V,
                            o.blank_label, workspace, nullptr, 0, T(1), st);
        if (rc != EB_OK) return RNNT_STATUS_EXECUTION_FAILED;
    }
    Workspace<T> w(workspace, B, o.maxT, o.maxU);
    if (cudaMemcpyAsync(costs, w.ll_fwd, sizeof(T) * B, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return RNNT_STATUS_MEMOPS_FAILED;
    if (cudaStreamSynchronize(st) != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
    for (int i = 0; i < B; ++i) costs[i] = -costs[i];   // gpu_rnnt.h:209-213
    return RNNT_STATUS_SUCCESS;
}

}  // namespace

// ------------------------------ warp-transducer compatible C ABI ------------------------------
extern "C" {

__attribute__((visibility("default"))) int get_warprnnt_version() { return 1; }

__attribute__((visibility("default"))) const char* rnntGetStatusString(rnntStatus_t status) {
    switch (status) {
        case RNNT_STATUS_SUCCESS: return "no error";
        case RNNT_STATUS_MEMOPS_FAILED: return "cuda memcpy or memset failed";
        case RNNT_STATUS_INVALID_VALUE: return "invalid value";
        case RNNT_STATUS_EXECUTION_FAILED: return "execution failed";
        default: return "unknown error";
    }
}

__attribute__((visibility("default"))) rnntStatus_t
compute_rnnt_loss(const float* const activations, float* gradients, const int* const flat_labels,
                  const int* const label_lengths, const int* const input_lengths, int alphabet_size,
                  int minibatch, float* costs, void* workspace, rnntOptions options) {
    return compat_entry<float>(activations, gradients, flat_labels, label_lengths, input_lengths,
                               alphabet_size, minibatch, costs, workspace, options);
}

__attribute__((visibility("default"))) rnntStatus_t
compute_rnnt_loss_fp64(const double* const activations, double* gradients,
                       const int* const flat_labels, const int* const label_lengths,
                       const int* const input_lengths, int alphabet_size, int minibatch,
                       double* costs, void* workspace, rnntOptions options) {
    return compat_entry<double>(activations, gradients, flat_labels, label_lengths, input_lengths,
                                alphabet_size, minibatch, costs, workspace, options);
}

__attribute__((visibility("default"))) rnntStatus_t
get_workspace_size(int maxT, int maxU, int minibatch, bool gpu, size_t* size_bytes,
                   size_t dtype_size) {
    if (minibatch <= 0 || maxT <= 0 || maxU <= 0 || !size_bytes) return RNNT_STATUS_INVALID_VALUE;
    (void)gpu;  // only the GPU location exists in this build; same size either way
    *size_bytes = dtype_size * ((size_t)minibatch * maxT * maxU * 5 + 2 * (size_t)minibatch);
    return RNNT_STATUS_SUCCESS;
}

}  // extern "C"

// ------------------------------ device-resident (no host sync) ABI ----------------------------
EB_API size_t eb_rnnt_workspace_bytes(int B, int maxT, int maxU, int dtype_size) {
    return (size_t)dtype_size * ((size_t)B * maxT * maxU * 5 + 2 * (size_t)B);
}

EB_API int eb_rnnt_loss_fwd(const void* logits, const int* labels, const int* xlen, const int* ylen,
                            int B, int maxT, int maxU, int V, int blank, int dtype_size,
                            void* workspace, void* costs_dev, int need_beta, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc;
    if (dtype_size == 4) {
        rc = loss_fwd<float>((const float*)logits, labels, xlen, ylen, B, maxT, maxU, V, blank,
                             workspace, need_beta, st);
        if (rc) return rc;
        if (costs_dev) {
            Workspace<float> w(workspace, B, maxT, maxU);
            neg_copy_kernel<float><<<(B + 127) / 128, 128, 0, st>>>(w.ll_fwd, (float*)costs_dev, B);
        }
    } else if (dtype_size == 8) {
        rc = loss_fwd<double>((const double*)logits, labels, xlen, ylen, B, maxT, maxU, V, blank,
                              workspace, need_beta, st);
        if (rc) return rc;
        if (costs_dev) {
            Workspace<double> w(workspace, B, maxT, maxU);
            neg_copy_kernel<double><<<(B + 127) / 128, 128, 0, st>>>(w.ll_fwd, (double*)costs_dev, B);
        }
    } else {
        return EB_ERR_INVALID;
    }
    EB_CHECK_LAUNCH();
    return EB_OK;
}

EB_API int eb_rnnt_loss_bwd(const void* logits, void* grads, int grads_bf16, const int* labels,
                            const int* xlen, const int* ylen, int B, int maxT, int maxU, int V,
                            int blank, int dtype_size, void* workspace, const void* gscale_dev,
                            int gscale_per_batch, double host_scale, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype_size == 4) {
        if (grads_bf16)
            return loss_bwd<float, __nv_bfloat16>((const float*)logits, (__nv_bfloat16*)grads, labels,
                                                  xlen, ylen, B, maxT, maxU, V, blank, workspace,
                                                  (const float*)gscale_dev, gscale_per_batch,
                                                  (float)host_scale, st);
        return loss_bwd<float, float>((const float*)logits, (float*)grads, labels, xlen, ylen, B, maxT,
                                      maxU, V, blank, workspace, (const float*)gscale_dev,
                                      gscale_per_batch, (float)host_scale, st);
    }
    if (dtype_size == 8 && !grads_bf16)
        return loss_bwd<double, double>((const double*)logits, (double*)grads, labels, xlen, ylen, B,
                                        maxT, maxU, V, blank, workspace, (const double*)gscale_dev,
                                        gscale_per_batch, host_scale, st);
    return EB_ERR_INVALID;
}

// Lattice only: denom / lpb / lpl of the workspace were already produced by eb_joint_logits_lse.
EB_API int eb_rnnt_loss_lattice(const int* xlen, const int* ylen, int B, int maxT, int maxU, void* workspace,
                                float* costs_dev, int need_beta, void* stream) {
    if (!xlen || !ylen || !workspace || B <= 0 || maxT <= 0 || maxU <= 0 || maxU > 1024) return EB_ERR_INVALID;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Workspace<float> w(workspace, B, maxT, maxU);
    const int threads = ((maxU + 31) / 32) * 32;
    rnnt_lattice_kernel<float, 8><<<dim3(B, 2), threads, 2 * threads * sizeof(float), st>>>(
        w.lpb, w.lpl, xlen, ylen, w.alphas, w.betas, w.ll_fwd, w.ll_bwd, maxT, maxU, need_beta);
    EB_CHECK_LAUNCH();
    if (costs_dev) neg_copy_kernel<float><<<(B + 127) / 128, 128, 0, st>>>(w.ll_fwd, costs_dev, B);
    EB_CHECK_LAUNCH();
    return EB_OK;
}

// Gradient wrt bf16 logits, written as bf16 (grads16 may alias logits16: in place).
EB_API int eb_rnnt_loss_bwd_bf16(const void* logits16, void* grads16, const int* labels, const int* xlen,
                                 const int* ylen, int B, int maxT, int maxU, int V, int blank, void* workspace,
                                 const float* gscale_dev, int gscale_per_batch, double host_scale, void* stream) {
    if (!logits16 || !grads16 || !workspace) return EB_ERR_INVALID;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Workspace<float> w(workspace, B, maxT, maxU);
    const long ncells = (long)B * maxT * maxU;
    constexpr int WARPS = 8;
    const bool vec = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits16) & 7) == 0) &&
                     ((reinterpret_cast<uintptr_t>(grads16) & 7) == 0);
    const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(logits16);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(grads16);
    const bool vec8 = (V % 8 == 0) && ((reinterpret_cast<uintptr_t>(logits16) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(grads16) & 15) == 0);
    if (vec8)
        rnnt_grad_bf16x8_kernel<WARPS><<<row_grid(ncells, WARPS), WARPS * 32, 0, st>>>(
            in, out, labels, xlen, ylen, w.denom, w.alphas, w.betas, w.ll_fwd, gscale_dev, gscale_per_batch,
            (float)host_scale, B, maxT, maxU, V, blank);
    else if (vec)
        rnnt_grad_kernel<float, __nv_bfloat16, true, WARPS, __nv_bfloat16><<<row_grid(ncells, WARPS), WARPS * 32, 0, st>>>(
            in, out, labels, xlen, ylen, w.denom, w.alphas, w.betas, w.ll_fwd, gscale_dev, gscale_per_batch,
            (float)host_scale, B, maxT, maxU, V, blank);
    else
        rnnt_grad_kernel<float, __nv_bfloat16, false, WARPS, __nv_bfloat16><<<row_grid(ncells, WARPS), WARPS * 32, 0, st>>>(
            in, out, labels, xlen, ylen, w.denom, w.alphas, w.betas, w.ll_fwd, gscale_dev, gscale_per_batch,
            (float)host_scale, B, maxT, maxU, V, blank);
    EB_CHECK_LAUNCH();
    return EB_OK;
}

// debugging / test access to the lattice (device pointers into the workspace)
EB_API int eb_rnnt_workspace_views(void* workspace, int B, int maxT, int maxU, int dtype_size,
                                   void** denom, void** alphas, void** betas, void** ll_fwd,
                                   void** ll_bwd) {
    if (dtype_size == 4) {
        Workspace<float> w(workspace, B, maxT, maxU);
        *denom = w.denom; *alphas = w.alphas; *betas = w.betas; *ll_fwd = w.ll_fwd; *ll_bwd = w.ll_bwd;
    } else {
        Workspace<double> w(workspace, B, maxT, maxU);
        *denom = w.denom; *alphas = w.alphas; *betas = w.betas; *ll_fwd = w.ll_fwd; *ll_bwd = w.ll_bwd;
    }
    return EB_OK;
}
