// decode.cu -- streaming greedy decode as ONE persistent kernel per audio chunk (sm_100a).
//
// Replaces the Python loop of PytorchStreamDecoder.decode (rnnt/stream.py:93-120): per chunk the
// reference runs the stateful encoder, then for every encoder frame joint -> argmax(.item(): a
// host sync) -> optional predictor step.  Here the whole chunk for S concurrent streams is a
// "phase program" (built once by the host, edgedict_b200/stream_engine.py) that one cooperative
// kernel walks with a grid barrier between dependent phases -- no launch gaps, no host syncs:
//
//   LN      y = LayerNorm(x1 (+ x2))                 nn.LayerNorm + residual (models.py:47,66-70,124)
//   PAIR    y[s,t/2] = mean(x[s,t], x[s,t+1])        TimeReduction (models.py:21-29)
//   LSTM    one cell step for S streams              nn.LSTM step (models.py:45-46,145-147); the
//           x rows may come from an embedding table indexed by the last token, and the update can
//           be masked per stream (predictor advances only on non-blank, stream.py:111-116)
//   LINEAR  y = act(x1 W1^T (+ x2 W2^T) + b)         Linear / joint (models.py:129,148,163-167)
//   ARGMAX  token = argmax(logits) with the <unk> rule (stream.py:105-108: logit := 0, re-argmax)
//   COPY    y = x
//
// fp32 CUDA-core arithmetic throughout: the north star asks for token-for-token identical greedy
// output, which bf16 near-ties would break; a chunk of 64 streams is ~0.3 ms against 120 ms of
// audio, so the tensor pipe is not needed to be three orders of magnitude faster than real time.
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/edgedict_b200.h"

namespace {

constexpr int TR = 64, TC = 32, KC = 32;     // tile rows (streams), tile cols, k chunk

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned target) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(ctr, 1u);
        spin_wait_ge(ctr, target);
    }
    __syncthreads();
}

// acc[r][q] += sum_k A(row 2*rp + r, k) * Bw(col cg*4 + q, k) for one K segment
template <typename AF, typename BF>
__device__ __forceinline__ void tile_segment(float (&acc)[2][4], AF arow, BF brow, int K, int nrows, int ncols,
                                             float* As, float* Bs) {
    const int tid = threadIdx.x, rp = tid >> 3, cg = tid & 7;
    for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();
        for (int i = tid; i < TR * KC; i += 256) {
            const int r = i / KC, k = i % KC;
            float v = 0.f;
            if (r < nrows && k0 + k < K) v = __ldcg(arow(r) + k0 + k);   // activations: written in-kernel
            As[k * (TR + 4) + r] = v;
        }
        for (int i = tid; i < TC * KC; i += 256) {
            const int n = i / KC, k = i % KC;
            float v = 0.f;
            if (n < ncols && k0 + k < K) v = __ldg(brow(n) + k0 + k);    // weights: read-only
            Bs[k * (TC + 4) + n] = v;
        }
        __syncthreads();
        const int kmax = min(KC, K - k0);
#pragma unroll 8
        for (int k = 0; k < kmax; ++k) {
            const float2 a = *reinterpret_cast<const float2*>(As + k * (TR + 4) + rp * 2);
            const float4 b = *reinterpret_cast<const float4*>(Bs + k * (TC + 4) + cg * 4);
            acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
            acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
            acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
            acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
        }
    }
}

__device__ void phase_lstm(const EbPhase& p, float* As, float* Bs) {
    const int S = p.S, H = p.N;
    const int ctiles = (H + 7) / 8, rtiles = (S + TR - 1) / TR;
    const int tid = threadIdx.x, rp = tid >> 3, cg = tid & 7;
    for (int tile = blockIdx.x; tile < ctiles * rtiles; tile += gridDim.x) {
        const int s0 = (tile / ctiles) * TR, j0 = (tile % ctiles) * 8;
        const int nrows = min(TR, S - s0), nunits = min(8, H - j0);
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        auto x1row = [&](int r) -> const float* {
            if (p.flags & 2) return p.x1 + (long)__ldcg(p.tok_in + s0 + r) * p.ldx1;     // embedding row of the last token
            return p.x1 + (long)(s0 + r) * p.ldx1;
        };
        auto hrow = [&](int r) -> const float* { return p.x2 + (long)(s0 + r) * p.ldx2; };
        auto w1row = [&](int n) -> const float* { return p.w1 + ((long)(n & 3) * H + j0 + (n >> 2)) * p.ldw1; };
        auto w2row = [&](int n) -> const float* { return p.w2 + ((long)(n & 3) * H + j0 + (n >> 2)) * p.ldw2; };
        tile_segment(acc, x1row, w1row, p.K1, nrows, nunits * 4, As, Bs);
        tile_segment(acc, hrow, w2row, p.K2, nrows, nunits * 4, As, Bs);
        if (cg < nunits) {
            const int j = j0 + cg;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int s = s0 + rp * 2 + r;
                if (s >= S) continue;
                const bool active = !(p.flags & 4) || (__ldcg(p.tok_in + s) != p.aux);
                float hn;
                if (active) {
                    float g4[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) g4[g] = acc[r][g] + p.b1[(long)g * H + j] + p.b2[(long)g * H + j];
                    const float ig = sigmoidf_(g4[0]), fg = sigmoidf_(g4[1]), gg = tanhf(g4[2]), og = sigmoidf_(g4[3]);
                    const float cn = fg * __ldcg(p.c + (long)s * H + j) + ig * gg;
                    p.c[(long)s * H + j] = cn;
                    hn = og * tanhf(cn);
                } else {
                    hn = __ldcg(p.x2 + (long)s * p.ldx2 + j);
                }
                p.y[(long)s * p.ldy + j] = hn;
                if (p.y2) p.y2[(long)s * H + j] = hn;
            }
        }
    }
}

__device__ void phase_linear(const EbPhase& p, float* As, float* Bs) {
    const int S = p.S, N = p.N;
    const int ctiles = (N + TC - 1) / TC, rtiles = (S + TR - 1) / TR;
    const int tid = threadIdx.x, rp = tid >> 3, cg = tid & 7;
    for (int tile = blockIdx.x; tile < ctiles * rtiles; tile += gridDim.x) {
        const int s0 = (tile / ctiles) * TR, n0 = (tile % ctiles) * TC;
        const int nrows = min(TR, S - s0), ncols = min(TC, N - n0);
        float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        auto x1row = [&](int r) -> const float* { return p.x1 + (long)(s0 + r) * p.ldx1; };
        auto w1row = [&](int n) -> const float* { return p.w1 + (long)(n0 + n) * p.ldw1; };
        tile_segment(acc, x1row, w1row, p.K1, nrows, ncols, As, Bs);
        if (p.K2 > 0) {
            auto x2row = [&](int r) -> const float* { return p.x2 + (long)(s0 + r) * p.ldx2; };
            auto w2row = [&](int n) -> const float* { return p.w2 + (long)(n0 + n) * p.ldw2; };
            tile_segment(acc, x2row, w2row, p.K2, nrows, ncols, As, Bs);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int s = s0 + rp * 2 + r;
            if (s >= S) continue;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int n = n0 + cg * 4 + q;
                if (n >= N) continue;
                float v = acc[r][q] + (p.b1 ? p.b1[n] : 0.f);
                if (p.flags & 1) v = tanhf(v);
                p.y[(long)s * p.ldy + n] = v;
            }
        }
    }
}

__device__ void phase_ln(const EbPhase& p) {
    const int lane = threadIdx.x & 31, H = p.N;
    for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < p.S; r += gridDim.x * 8) {
        const float* x = p.x1 + (long)r * p.ldx1;
        const float* x2 = p.x2 ? p.x2 + (long)r * p.ldx2 : nullptr;
        float s = 0.f;
        for (int c = lane; c < H; c += 32) s += __ldcg(x + c) + (x2 ? __ldcg(x2 + c) : 0.f);
        const float mu = warp_sum(s) / H;
        float q = 0.f;
        for (int c = lane; c < H; c += 32) {
            float d = __ldcg(x + c) + (x2 ? __ldcg(x2 + c) : 0.f) - mu;
            q += d * d;
        }
        const float rs = rsqrtf(warp_sum(q) / H + 1e-5f);
        for (int c = lane; c < H; c += 32) {
            float z = __ldcg(x + c) + (x2 ? __ldcg(x2 + c) : 0.f);
            p.y[(long)r * p.ldy + c] = (z - mu) * rs * p.w1[c] + p.b1[c];
        }
    }
}

__device__ void phase_argmax(const EbPhase& p) {
    const int lane = threadIdx.x & 31, V = p.N, blank = p.aux, unk = p.aux2;
    (void)blank;
    for (int s = blockIdx.x * 8 + (threadIdx.x >> 5); s < p.S; s += gridDim.x * 8) {
        const float* x = p.x1 + (long)s * p.ldx1;
        int pred = -1;
        for (int pass = 0; pass < 2; ++pass) {
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int v = lane; v < V; v += 32) {
                float val = __ldcg(x + v);
                if (pass == 1 && v == pred) val = 0.f;              // stream.py:107 `prob[:, pred] = 0`
                if (val > best) { best = val; bi = v; }             // strict >: first index wins ties
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            pred = bi;
            if (pred != unk) break;
        }
        if (p.flags & 8) {                                     // batched greedy (models.py:253-255): accumulate
            const float best = __ldcg(x + pred);               // log_softmax(x)[pred] = -log sum exp(x - max)
            float sum = 0.f;
            for (int v = lane; v < V; v += 32) sum += expf(__ldcg(x + v) - best);
            sum = warp_sum(sum);
            if (lane == 0) p.y[s] += -logf(sum);
        }
        if (lane == 0) {
            p.tok_out[s] = pred;
            if (p.hist) p.hist[(long)s * p.hist_ld + p.hist_col] = pred;
        }
    }
}

__global__ void __launch_bounds__(256) decode_program_kernel(const EbPhase* __restrict__ prog, int nphase, unsigned* bar) {
    __shared__ __align__(16) float As[KC * (TR + 4)];
    __shared__ __align__(16) float Bs[KC * (TC + 4)];
    __shared__ EbPhase ph;
    unsigned epoch = 0;
    for (int i = 0; i < nphase; ++i) {
        __syncthreads();
        if (threadIdx.x < sizeof(EbPhase) / 4)
            reinterpret_cast<int*>(&ph)[threadIdx.x] = reinterpret_cast<const int*>(prog + i)[threadIdx.x];
        __syncthreads();
        const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gn = (long)gridDim.x * blockDim.x;
        switch (ph.type) {
            case EB_PH_LN: phase_ln(ph); break;
            case EB_PH_PAIR: {
                const int H = ph.N, n2 = ph.aux / 2;                // x1 [S, n, H] -> y [S, n/2, H]
                for (long k = gtid; k < (long)ph.S * n2 * H; k += gn) {
                    const int c = (int)(k % H);
                    const long st = k / H;
                    const int t2 = (int)(st % n2), s = (int)(st / n2);
                    const float* x = ph.x1 + ((long)s * ph.aux + 2 * t2) * H + c;
                    ph.y[k] = 0.5f * (__ldcg(x) + __ldcg(x + H));
                }
            } break;
            case EB_PH_LSTM: phase_lstm(ph, As, Bs); break;
            case EB_PH_LINEAR: phase_linear(ph, As, Bs); break;
            case EB_PH_ARGMAX: phase_argmax(ph); break;
            case EB_PH_COPY:
                for (long k = gtid; k < (long)ph.S * ph.N; k += gn) ph.y[k] = __ldcg(ph.x1 + k);
                break;
            default: break;
        }
        ++epoch;
        if (i + 1 < nphase) grid_sync(bar, epoch * gridDim.x);
    }
}

}  // namespace

EB_API int eb_decode_run(const void* phases_dev, int nphase, void* barrier_dev, int max_ctas, void* stream) {
    if (!phases_dev || nphase <= 0 || !barrier_dev) return EB_ERR_INVALID;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    EB_CUDA(cudaMemsetAsync(barrier_dev, 0, 4, st));
    int grid = eb_num_sms();
    if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
    const EbPhase* prog = reinterpret_cast<const EbPhase*>(phases_dev);
    unsigned* bar = reinterpret_cast<unsigned*>(barrier_dev);
    void* args[] = {(void*)&prog, (void*)&nphase, (void*)&bar};
    EB_CUDA(cudaLaunchCooperativeKernel((void*)decode_program_kernel, dim3(grid), dim3(256), args, 0, st));
    return EB_OK;
}

EB_API int eb_decode_phase_size(void) { return (int)sizeof(EbPhase); }
