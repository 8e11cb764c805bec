// lstm_tc.cu -- persistent-RNN LSTM layer on tensor cores (bf16 operands, fp32 accumulate / state),
// the bf16-mode replacement of lstm.cu for H % 64 == 0, H <= 1024 (all BASELINE configs).
//
// Forward (one launch for all T steps, H/8 CTAs x 256 threads):
//   * CTA k owns hidden units [8k, 8k+8): 32 gate rows of W_hh arranged as two m16 tiles
//     (i|f) and (g|o), so that after mma.sync.m16n8k16 one thread holds all four gates of a
//     (unit, batch) pair and the cell update is thread-local;
//   * W_hh is held in REGISTERS as mma A-fragments for the whole sequence: each of the 8 warps
//     owns a K-range of H/8 columns (<= 8 k-steps x 2 m-tiles x 4 regs = 64 registers);
//   * h_{t-1} (bf16, [batch][H]) is exchanged through an L2-resident double buffer; each warp
//     pulls only its own K-range with cp.async and feeds ldmatrix B-fragments; the 8 partial
//     accumulators are summed through shared memory; one grid barrier per timestep.
// Backward (BPTT): 2-D decomposition, cluster of 8 CTAs per 64-unit slice: CTA (js, rs) holds
//   W_hh[rs-th eighth of the 4H gate rows, 64 units of slice js] as A-fragments, multiplies by its
//   K-slice of dG_t (bf16 exchange buffer), reduces across warps in shared memory and across the
//   cluster through distributed shared memory, then the owning threads run the gate-gradient
//   math for step t-1 with dh/dc carried in registers.
//
// Semantics: nn.LSTM cell, gate order i|f|g|o (rnnt/models.py:45-46 -> torch.nn.LSTM).
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/edgedict_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int NW = 8;            // warps per CTA
constexpr int UPC = 8;           // hidden units per CTA (forward)
constexpr int PAD = 8;           // bf16 elements of row padding (16 B) -> conflict-free ldmatrix
constexpr int NB = 32;           // batch tile

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    unsigned a = (unsigned)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ldg_u32(const __nv_bfloat16* p) {
    return *reinterpret_cast<const uint32_t*>(p);
}
// every warp's lane 0 polls; no block-wide barrier on the wait side
__device__ __forceinline__ void warp_wait(const unsigned* ctr, unsigned target) {
    if ((threadIdx.x & 31) == 0) {
        spin_wait_ge(ctr, target);
    }
    __syncwarp();
}

struct FwdP {
    const float* xg;              // [B,T,4H] fp32
    const __nv_bfloat16* whh;     // [4H,H] bf16
    const float* h0; const float* c0;
    float* y; __nv_bfloat16* y16; float* hT; float* cT; float* gates; float* cseq;
    __nv_bfloat16* hx;            // [2][NB][H] exchange
    unsigned* bar;
    int B, T, H;
};

__global__ void __launch_bounds__(NW * 32, 1) lstm_tc_fwd_kernel(FwdP p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int H = p.H, B = p.B, T = p.T;
    const int HP = H + PAD;
    __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(smraw);                 // [NB][HP]
    float* red = reinterpret_cast<float*>(smraw + (size_t)NB * HP * 2);          // [NW][32][32]
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int j0 = blockIdx.x * UPC;
    const unsigned ncta = gridDim.x;
    const int nks = H / 16;                                  // k-steps in total
    const int ksper = (nks + NW - 1) / NW;                   // per warp (<= 8)
    const int ks0 = w * ksper;
    const int myks = max(0, min(ksper, nks - ks0));

    // ---- resident A fragments: afr[mt][ks][4]; mt 0 = rows (i: 0-7, f: 8-15), mt 1 = (g, o)
    uint32_t afr[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const int k = (ks0 + ks) * 16 + (l & 3) * 2;
            const int u = j0 + (l >> 2);
            const __nv_bfloat16* rlo = p.whh + ((long)(mt * 2 + 0) * H + u) * H + k;   // gate i / g
            const __nv_bfloat16* rhi = p.whh + ((long)(mt * 2 + 1) * H + u) * H + k;   // gate f / o
            const bool ok = ks < myks;
            afr[mt][ks][0] = ok ? ldg_u32(rlo) : 0u;
            afr[mt][ks][1] = ok ? ldg_u32(rhi) : 0u;
            afr[mt][ks][2] = ok ? ldg_u32(rlo + 8) : 0u;
            afr[mt][ks][3] = ok ? ldg_u32(rhi + 8) : 0u;
        }

    // ---- the (unit, batch) pair this thread finalises every step
    const int ju = l >> 2;
    const int bb = (w >> 1) * 8 + (l & 3) * 2 + (w & 1);
    const int j = j0 + ju;
    const bool own = bb < B;
    float c_state = (own && p.c0) ? p.c0[(long)bb * H + j] : 0.f;
    __nv_bfloat16* hx[2] = {p.hx, p.hx + (size_t)NB * H};
    if (own) hx[1][(long)bb * H + j] = __float2bfloat16(p.h0 ? p.h0[(long)bb * H + j] : 0.f);
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(p.bar, 1u);
    unsigned epoch = 1;

    float px[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) px[g] = own ? __ldg(p.xg + ((long)bb * T + 0) * 4 * H + (long)g * H + j) : 0.f;

    for (int t = 0; t < T; ++t) {
        const __nv_bfloat16* hprev = hx[(t + 1) & 1];
        __nv_bfloat16* hnext = hx[t & 1];
        warp_wait(p.bar, epoch * ncta);
        // pull this warp's K-range of h_{t-1}: NB rows x (myks*16) bf16
        {
            const int chunks_per_row = myks * 2;                 // 16-byte chunks
            const int kbase = ks0 * 16;
            for (int i = l; i < NB * chunks_per_row; i += 32) {
                const int r = i / chunks_per_row, q = i % chunks_per_row;
                cp_async16(hs + (size_t)r * HP + kbase + q * 8, hprev + (size_t)r * H + kbase + q * 8);
            }
            cp_async_wait_all();
            __syncwarp();
        }
        float acc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            if (ks < myks) {
                uint32_t b01[4], b23[4];
                // lanes 0-7 / 8-15 / 16-23 / 24-31 address matrices (nt, k-lo), (nt, k-hi), (nt+1, k-lo), (nt+1, k-hi)
                const int mrow = (l & 7) + ((l >> 4) & 1) * 8;
                const int mk = (ks0 + ks) * 16 + ((l >> 3) & 1) * 8;
                ldmatrix_x4(b01, hs + (size_t)mrow * HP + mk);
                ldmatrix_x4(b23, hs + (size_t)(16 + mrow) * HP + mk);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_bf16(acc[mt][0], afr[mt][ks], b01[0], b01[1]);
                    mma_bf16(acc[mt][1], afr[mt][ks], b01[2], b01[3]);
                    mma_bf16(acc[mt][2], afr[mt][ks], b23[0], b23[1]);
                    mma_bf16(acc[mt][3], afr[mt][ks], b23[2], b23[3]);
                }
            }
        }
        // partials -> shared: red[w][slot][lane], slot = nt*8 + mt*4 + i
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) red[(w * 32 + nt * 8 + mt * 4 + i) * 32 + l] = acc[mt][nt][i];
        __syncthreads();
        {
            // this thread: nt = w>>1, batch offset = w&1 -> slots (i: mt0,c=off) (f: mt0,c=2+off) (g: mt1,off) (o: mt1,2+off)
            const int nt = w >> 1, off = w & 1;
            float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int sw = 0; sw < NW; ++sw) {
                const float* r = red + (size_t)(sw * 32 + nt * 8) * 32 + l;
                s[0] += r[(0 + off) * 32];
                s[1] += r[(2 + off) * 32];
                s[2] += r[(4 + off) * 32];
                s[3] += r[(6 + off) * 32];
            }
            if (own) {
                const float ig = sigmoidf_(s[0] + px[0]);
                const float fg = sigmoidf_(s[1] + px[1]);
                const float gg = tanhf(s[2] + px[2]);
                const float og = sigmoidf_(s[3] + px[3]);
                c_state = fg * c_state + ig * gg;
                const float hn = og * tanhf(c_state);
                const long bt = (long)bb * T + t;
                hnext[(long)bb * H + j] = __float2bfloat16(hn);
                p.y[bt * H + j] = hn;
                if (p.y16) p.y16[bt * H + j] = __float2bfloat16(hn);
                if (p.gates) {
                    float* gp = p.gates + bt * 4 * H + j;
                    gp[0] = ig; gp[H] = fg; gp[2 * (long)H] = gg; gp[3 * (long)H] = og;
                }
                if (p.cseq) p.cseq[bt * H + j] = c_state;
                if (t == T - 1) { p.hT[(long)bb * H + j] = hn; p.cT[(long)bb * H + j] = c_state; }
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(p.bar, 1u);
        ++epoch;
        if (t + 1 < T) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
                px[g] = own ? __ldg(p.xg + ((long)bb * T + t + 1) * 4 * H + (long)g * H + j) : 0.f;
        }
    }
}

// ------------------------------------------------------------------------------------------------
struct BwdP {
    const float* dy; const float* gates; const float* cseq; const float* c0;
    const __nv_bfloat16* whhT;    // [H,4H] bf16 = W_hh^T
    const float* dhT; const float* dcT;
    __nv_bfloat16* dg16;          // [B,T,4H] bf16 gate-preactivation gradients (output)
    float* dh0; float* dc0;
    __nv_bfloat16* gx;            // [2][NB][4H] exchange
    unsigned* bar;
    unsigned* gbar;               // per 64-unit slice counters (non-cluster variant)
    float* pglob;                 // [H/64][8][JS][NB] partial tiles in L2 (non-cluster variant)
    int B, T, H;
};

constexpr int JS = 64;           // units per cluster (j-slice); 8 CTAs of a cluster split the 4H rows

// CLUSTER = true : the 8 CTAs of a 64-unit slice form a thread-block cluster; partial tiles are
//                  exchanged through distributed shared memory behind one hardware cluster barrier.
// CLUSTER = false: same decomposition on a plain cooperative grid; partial tiles go through L2 and
//                  a per-slice software barrier (used when 16 clusters of 8 cannot be co-resident).
template <bool CLUSTER>
__global__ void __launch_bounds__(NW * 32, 1) lstm_tc_bwd_kernel(BwdP p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int H = p.H, B = p.B, T = p.T, H4 = 4 * H;
    const int KR = H4 / 8;                                   // gate rows (contraction) per CTA = H/2
    const int KP = KR + PAD;
    __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(smraw);                 // [NB][KP]
    float* red = reinterpret_cast<float*>(smraw + (size_t)NB * KP * 2);          // [NW][64][32]
    float* part = red + NW * 64 * 32;                                            // [JS][NB]
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int rs = blockIdx.x & 7;                           // which eighth of the 4H rows (= cluster rank)
    const int js = blockIdx.x >> 3;                          // which 64-unit slice
    const int r0 = rs * KR;
    const unsigned ncta = gridDim.x;
    const int nks = KR / 16;
    const int ksper = (nks + NW - 1) / NW;                   // <= 4
    const int ks0 = w * ksper;
    const int myks = max(0, min(ksper, nks - ks0));

    // A(m = unit, k = gate row) = W_hh[r, j] = whhT[j][r]; afr[mt][ks][4], 4 m-tiles of 16 units
    uint32_t afr[4][4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int k = r0 + (ks0 + ks) * 16 + (l & 3) * 2;
            const int u = js * JS + mt * 16 + (l >> 2);
            const __nv_bfloat16* rlo = p.whhT + (long)u * H4 + k;
            const __nv_bfloat16* rhi = p.whhT + (long)(u + 8) * H4 + k;
            const bool ok = ks < myks;
            afr[mt][ks][0] = ok ? ldg_u32(rlo) : 0u;
            afr[mt][ks][1] = ok ? ldg_u32(rhi) : 0u;
            afr[mt][ks][2] = ok ? ldg_u32(rlo + 8) : 0u;
            afr[mt][ks][3] = ok ? ldg_u32(rhi + 8) : 0u;
        }

    // phase-A ownership: unit = js*64 + rs*8 + w, batch = lane
    const int j = js * JS + rs * 8 + w;
    const int bb = l;
    const bool own = bb < B;
    float dh = (own && p.dhT) ? p.dhT[(long)bb * H + j] : 0.f;
    float dc = (own && p.dcT) ? p.dcT[(long)bb * H + j] : 0.f;
    __nv_bfloat16* gx[2] = {p.gx, p.gx + (size_t)NB * H4};
    unsigned epoch = 0;

    for (int t = T - 1; t >= 0; --t) {
        __nv_bfloat16* gcur = gx[t & 1];
        // ---- phase A: gate gradients of step t for the owned (unit, batch)
        if (own) {
            const long bt = (long)bb * T + t;
            const float* gp = p.gates + bt * H4 + j;
            const float ig = gp[0], fg = gp[H], gg = gp[2 * (long)H], og = gp[3 * (long)H];
            const float ct = p.cseq[bt * H + j];
            const float cprev = (t > 0) ? p.cseq[(bt - 1) * H + j] : (p.c0 ? p.c0[(long)bb * H + j] : 0.f);
            const float tc = tanhf(ct);
            const float dht = p.dy[bt * H + j] + dh;
            const float dct = dc + dht * og * (1.f - tc * tc);
            const float da[4] = {dct * gg * ig * (1.f - ig), dct * cprev * fg * (1.f - fg),
                                 dct * ig * (1.f - gg * gg), dht * tc * og * (1.f - og)};
            dc = dct * fg;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const __nv_bfloat16 v = __float2bfloat16(da[g]);
                p.dg16[bt * H4 + (long)g * H + j] = v;
                gcur[(long)bb * H4 + (long)g * H + j] = v;
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(p.bar, 1u);
        ++epoch;
        warp_wait(p.bar, epoch * ncta);
        // ---- phase B: partial dh_rec[unit (64), batch] over this CTA's K-slice of dG_t
        {
            const int chunks_per_row = myks * 2;
            const int kbase = ks0 * 16;
            for (int i = l; i < NB * chunks_per_row; i += 32) {
                const int r = i / chunks_per_row, q = i % chunks_per_row;
                cp_async16(gs + (size_t)r * KP + kbase + q * 8, gcur + (size_t)r * H4 + r0 + kbase + q * 8);
            }
            cp_async_wait_all();
            __syncwarp();
        }
        float acc[4][4][4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            if (ks < myks) {
                uint32_t b01[4], b23[4];
                const int mrow = (l & 7) + ((l >> 4) & 1) * 8;
                const int mk = (ks0 + ks) * 16 + ((l >> 3) & 1) * 8;
                ldmatrix_x4(b01, gs + (size_t)mrow * KP + mk);
                ldmatrix_x4(b23, gs + (size_t)(16 + mrow) * KP + mk);
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    mma_bf16(acc[mt][0], afr[mt][ks], b01[0], b01[1]);
                    mma_bf16(acc[mt][1], afr[mt][ks], b01[2], b01[3]);
                    mma_bf16(acc[mt][2], afr[mt][ks], b23[0], b23[1]);
                    mma_bf16(acc[mt][3], afr[mt][ks], b23[2], b23[3]);
                }
            }
        }
        // cross-warp reduction: red[w][slot][lane], slot = mt*16 + nt*4 + i
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) red[(w * 64 + mt * 16 + nt * 4 + i) * 32 + l] = acc[mt][nt][i];
        __syncthreads();
        // thread (w,l) sums slots [8w, 8w+8): slot -> (mt = slot/16, nt = (slot/4)%4, i = slot%4)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int slot = w * 8 + q;
            float s = 0.f;
#pragma unroll
            for (int sw = 0; sw < NW; ++sw) s += red[(sw * 64 + slot) * 32 + l];
            const int mt = slot >> 4, nt = (slot >> 2) & 3, i = slot & 3;
            const int unit = mt * 16 + (l >> 2) + (i >> 1) * 8;
            const int bcol = nt * 8 + (l & 3) * 2 + (i & 1);
            if (CLUSTER) part[unit * NB + bcol] = s;
            else p.pglob[(((size_t)js * 8 + rs) * JS + unit) * NB + bcol] = s;
        }
        if (CLUSTER) {
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();                                 // all 8 partial tiles of the slice are visible
            // reduce-scatter through distributed shared memory: this CTA finalises units [8*rs, 8*rs+8)
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float* rp = cluster.map_shared_rank(part, c);
                s += rp[(rs * 8 + w) * NB + l];
            }
            dh = s;                                          // dh_rec for (unit j, batch l) at step t-1
            // `part` / `red` are rewritten only after the next grid barrier, which every CTA of the
            // cluster reaches after finishing the remote reads above.
        } else {
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(p.gbar + js, 1u);
            warp_wait(p.gbar + js, 8u * (unsigned)(T - t));
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                s += __ldcg(p.pglob + (((size_t)js * 8 + c) * JS + rs * 8 + w) * NB + l);
            dh = s;
            // pglob is rewritten one full grid barrier later (same argument as above)
        }
    }
    if (own) {
        p.dh0[(long)bb * H + j] = dh;
        p.dc0[(long)bb * H + j] = dc;
    }
    if (CLUSTER) cg::this_cluster().sync();                  // no CTA exits while its smem may be read
}

inline bool tc_ok(int B, int H) { return H % JS == 0 && H <= 1024 && B >= 1; }

}  // namespace

EB_API int eb_lstm_tc_supported(int B, int H) { return tc_ok(B, H) ? 1 : 0; }

constexpr size_t TC_HDR = 1024;   // [0,256) grid barrier, [256,1024) per-slice barriers

EB_API size_t eb_lstm_tc_scratch_bytes(int B, int H) {
    if (!tc_ok(B, H)) return 0;
    return TC_HDR + sizeof(__nv_bfloat16) * (size_t)2 * NB * 4 * H + sizeof(float) * (size_t)(H / JS) * 8 * JS * NB;
}

// how many 8-CTA clusters of the BPTT kernel can be co-resident (diagnostic + path selection)
EB_API int eb_lstm_tc_max_clusters(int H) {
    if (H % JS || H > 1024) return -1;
    const int KR = 4 * H / 8;
    const size_t smem = (size_t)NB * (KR + PAD) * 2 + sizeof(float) * (NW * 64 * 32 + JS * NB);
    if (cudaFuncSetAttribute(lstm_tc_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((H / JS) * 8);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 8; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lstm_tc_bwd_kernel<true>, &cfg) != cudaSuccess) { (void)cudaGetLastError(); return -3; }
    return n;
}

// xg [B,T,4H] fp32; whh16 [4H,H] bf16.  B > 32 is processed in batch tiles of 32 (independent
// utterances), one launch per tile.
EB_API int eb_lstm_tc_fwd(const float* xg, const void* whh16, const float* h0, const float* c0, float* y,
                          void* y16, float* hT, float* cT, float* gates_save, float* cseq_save,
                          void* scratch, int B, int T, int H, void* stream) {
    if (!xg || !whh16 || !y || !hT || !cT || !scratch || T <= 0 || !tc_ok(B, H)) return EB_ERR_INVALID;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)NB * (H + PAD) * 2 + sizeof(float) * NW * 32 * 32;
    EB_CUDA(cudaFuncSetAttribute(lstm_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int b0 = 0; b0 < B; b0 += NB) {
        const int nb = (B - b0 < NB) ? (B - b0) : NB;
        FwdP p;
        p.xg = xg + (size_t)b0 * T * 4 * H;
        p.whh = reinterpret_cast<const __nv_bfloat16*>(whh16);
        p.h0 = h0 ? h0 + (size_t)b0 * H : nullptr;
        p.c0 = c0 ? c0 + (size_t)b0 * H : nullptr;
        p.y = y + (size_t)b0 * T * H;
        p.y16 = y16 ? reinterpret_cast<__nv_bfloat16*>(y16) + (size_t)b0 * T * H : nullptr;
        p.hT = hT + (size_t)b0 * H;
        p.cT = cT + (size_t)b0 * H;
        p.gates = gates_save ? gates_save + (size_t)b0 * T * 4 * H : nullptr;
        p.cseq = cseq_save ? cseq_save + (size_t)b0 * T * H : nullptr;
        p.bar = reinterpret_cast<unsigned*>(scratch);
        p.hx = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(scratch) + TC_HDR);
        p.B = nb; p.T = T; p.H = H;
        EB_CUDA(cudaMemsetAsync(scratch, 0, TC_HDR + sizeof(__nv_bfloat16) * (size_t)2 * NB * H, st));
        void* args[] = {&p};
        EB_CUDA(cudaLaunchCooperativeKernel((void*)lstm_tc_fwd_kernel, dim3(H / UPC), dim3(NW * 32), args, smem, st));
    }
    return EB_OK;
}

// whhT16 [H,4H] bf16 (W_hh transposed).  dg16 [B,T,4H] bf16 out; dh0/dc0 [B,H] fp32 out.
EB_API int eb_lstm_tc_bwd(const float* dy, const float* gates, const float* cseq, const float* c0,
                          const void* whhT16, const float* dhT, const float* dcT, void* dg16, float* dh0,
                          float* dc0, void* scratch, int B, int T, int H, void* stream) {
    if (!dy || !gates || !cseq || !whhT16 || !dg16 || !dh0 || !dc0 || !scratch || T <= 0 || !tc_ok(B, H))
        return EB_ERR_INVALID;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int KR = 4 * H / 8;
    const size_t smem = (size_t)NB * (KR + PAD) * 2 + sizeof(float) * (NW * 64 * 32 + JS * NB);
    const int nclusters = H / JS;
    static int use_cluster = -1;            // decided once: env override, else occupancy query
    if (use_cluster < 0) {
        const char* e = getenv("EDGEDICT_LSTM_CLUSTER");
        if (e) use_cluster = atoi(e) ? 1 : 0;
        else use_cluster = (eb_lstm_tc_max_clusters(1024) >= 16) ? 1 : 0;
    }
    bool cluster = use_cluster == 1 && eb_lstm_tc_max_clusters(H) >= nclusters;
    EB_CUDA(cudaFuncSetAttribute(lstm_tc_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EB_CUDA(cudaFuncSetAttribute(lstm_tc_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    char* base = reinterpret_cast<char*>(scratch);
    for (int b0 = 0; b0 < B; b0 += NB) {
        const int nb = (B - b0 < NB) ? (B - b0) : NB;
        BwdP p;
        p.dy = dy + (size_t)b0 * T * H;
        p.gates = gates + (size_t)b0 * T * 4 * H;
        p.cseq = cseq + (size_t)b0 * T * H;
        p.c0 = c0 ? c0 + (size_t)b0 * H : nullptr;
        p.whhT = reinterpret_cast<const __nv_bfloat16*>(whhT16);
        p.dhT = dhT ? dhT + (size_t)b0 * H : nullptr;
        p.dcT = dcT ? dcT + (size_t)b0 * H : nullptr;
        p.dg16 = reinterpret_cast<__nv_bfloat16*>(dg16) + (size_t)b0 * T * 4 * H;
        p.dh0 = dh0 + (size_t)b0 * H;
        p.dc0 = dc0 + (size_t)b0 * H;
        p.bar = reinterpret_cast<unsigned*>(base);
        p.gbar = reinterpret_cast<unsigned*>(base + 256);
        p.gx = reinterpret_cast<__nv_bfloat16*>(base + TC_HDR);
        p.pglob = reinterpret_cast<float*>(base + TC_HDR + sizeof(__nv_bfloat16) * (size_t)2 * NB * 4 * H);
        p.B = nb; p.T = T; p.H = H;
        EB_CUDA(cudaMemsetAsync(scratch, 0, TC_HDR + sizeof(__nv_bfloat16) * (size_t)2 * NB * 4 * H, st));
        bool launched = false;
        if (cluster) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nclusters * 8);
            cfg.blockDim = dim3(NW * 32);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attrs[2];
            attrs[0].id = cudaLaunchAttributeClusterDimension;
            attrs[0].val.clusterDim.x = 8; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
            attrs[1].id = cudaLaunchAttributeCooperative;
            attrs[1].val.cooperative = 1;
            cfg.attrs = attrs;
            cfg.numAttrs = 2;
            cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_tc_bwd_kernel<true>, p);
            if (e == cudaSuccess) launched = true;
            else { (void)cudaGetLastError(); cluster = false; use_cluster = 0; }   // never launch un-guaranteed
        }
        if (!launched) {
            void* args[] = {&p};
            EB_CUDA(cudaLaunchCooperativeKernel((void*)lstm_tc_bwd_kernel<false>, dim3(nclusters * 8), dim3(NW * 32),
                                                args, smem, st));
        }
    }
    return EB_OK;
}
