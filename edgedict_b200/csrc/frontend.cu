// frontend.cu -- log-mel front end (the step before the hot path), fp32.
//
// Reference: rnnt/features.py:126-164 (FilterbankFeatures.forward: pre-emphasis -> torch.stft(center, reflect,
// hann of win_length centred in n_fft) -> power -> mel matmul -> log(x + 1e-20) -> zero frames >= ceil(L/hop)) and
// rnnt/transforms.py:37-51 (Downsample: stack n_frame consecutive frames, zero-pad to a multiple).
//
// Decomposition (all on the caller's stream):
//   K1 fe_preemph_pad   : x[B,L] -> xp[B,Lp]  pre-emphasised, reflect-padded by n_fft/2, rows padded with zeros
//                         to Lp = multiple of hop, so that frame g = b*(Lp/hop) + f of the FLAT buffer starts at
//                         g*hop: the framing is a strided view (row stride hop), never materialised;
//   G1 eb_gemm_f32      : spec[g, 0:NB | NB:2NB] = frames[g, :] @ (window * cos | -window * sin)   (direct DFT as an
//                         exact-fp32 GEMM: n_fft = 512 is a 512-deep contraction, 0.4 GFLOP per second of audio);
//   K2 fe_power         : P[g,k] = re^2 + im^2;
//   G2 eb_gemm_f32      : mel[g,m] = P[g,:] @ fb[m,:]^T;
//   K3 fe_log_stack     : out[b,t,s*n_mels+m] = log(mel[g(b, t*n_frame+s), m] + 1e-20), zero for masked / padded
//                         frames -- directly in the [B, T, n_mels*n_frame] layout Encoder.forward consumes.
#include "common.cuh"
#include "../../include/edgedict_b200.h"

namespace {

__global__ void fe_preemph_pad_kernel(const float* __restrict__ x, float* __restrict__ xp, int L, long Lp, int pad,
                                      float preemph, int use_preemph) {
    const int b = blockIdx.y;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Lp) return;
    float v = 0.f;
    if (i < (long)L + 2 * pad) {
        long r = i - pad;                               // reflect (no edge repeat): -1 -> 1, L -> L-2
        if (r < 0) r = -r;
        if (r >= L) r = 2L * (L - 1) - r;
        const float* row = x + (long)b * L;
        v = row[r];
        if (use_preemph && r > 0) v -= preemph * row[r - 1];
    }
    xp[(long)b * Lp + i] = v;
}

__global__ void fe_power_kernel(const float* __restrict__ spec, float* __restrict__ pw, long rows, int NB) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * NB) return;
    const long g = i / NB;
    const int k = (int)(i % NB);
    const float re = spec[g * 2 * NB + k], im = spec[g * 2 * NB + NB + k];
    pw[i] = re * re + im * im;
}

__global__ void fe_log_stack_kernel(const float* __restrict__ mel, float* __restrict__ out, int R, int F, int seq_len,
                                    int n_mels, int n_frame, int Tout, int take_log) {
    const int b = blockIdx.y;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int W = n_mels * n_frame;
    if (i >= (long)Tout * W) return;
    const int t = (int)(i / W), c = (int)(i % W);
    const int s = c / n_mels, m = c % n_mels;
    const int f = t * n_frame + s;
    float v = 0.f;
    if (f < F && f < seq_len) {
        v = mel[((long)b * R + f) * n_mels + m];
        if (take_log) v = logf(v + 1e-20f);
    }
    out[(long)b * Tout * W + i] = v;
}

// SpecAugment masking (rnnt/transforms.py:53-147): x [B, D1, D2]; spans [B, nmask, 2] = [start, end) along dim `axis`
// (1: frequency rows, 2: time columns); masked elements := fill.
__global__ void fe_mask_kernel(float* __restrict__ x, const int* __restrict__ spans, int nmask, int D1, int D2, int axis,
                               float fill) {
    const int b = blockIdx.y;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)D1 * D2) return;
    const int pos = axis == 1 ? (int)(i / D2) : (int)(i % D2);
    const int* sp = spans + (long)b * nmask * 2;
    bool hit = false;
    for (int m = 0; m < nmask; ++m) hit = hit || (pos >= sp[2 * m] && pos < sp[2 * m + 1]);
    if (hit) x[(long)b * D1 * D2 + i] = fill;
}

}  // namespace

EB_API int eb_fe_mask(float* x, const int* spans, int B, int D1, int D2, int nmask, int axis, float fill, void* stream) {
    if (!x || !spans || B <= 0 || D1 <= 0 || D2 <= 0 || nmask <= 0 || (axis != 1 && axis != 2)) return EB_ERR_INVALID;
    dim3 grid((unsigned)(((long)D1 * D2 + 255) / 256), B);
    fe_mask_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, spans, nmask, D1, D2, axis, fill);
    EB_CHECK_LAUNCH();
    return EB_OK;
}

EB_API int eb_fe_preemph_pad(const float* x, float* xp, int B, int L, long Lp, int pad, float preemph,
                             int use_preemph, void* stream) {
    if (!x || !xp || B <= 0 || L <= 1 || pad < 0 || pad >= L || Lp < (long)L + 2 * pad) return EB_ERR_INVALID;
    dim3 grid((unsigned)((Lp + 255) / 256), B);
    fe_preemph_pad_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, xp, L, Lp, pad, preemph, use_preemph);
    EB_CHECK_LAUNCH();
    return EB_OK;
}

EB_API int eb_fe_power(const float* spec, float* power, long rows, int nbins, void* stream) {
    if (!spec || !power || rows <= 0 || nbins <= 0) return EB_ERR_INVALID;
    const long n = rows * nbins;
    fe_power_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(spec, power, rows, nbins);
    EB_CHECK_LAUNCH();
    return EB_OK;
}

EB_API int eb_fe_log_stack(const float* mel, float* out, int B, int rows_per_utt, int n_frames, int seq_len,
                           int n_mels, int n_stack, int t_out, int take_log, void* stream) {
    if (!mel || !out || B <= 0 || rows_per_utt < n_frames || n_frames <= 0 || n_mels <= 0 || n_stack <= 0 ||
        t_out <= 0)
        return EB_ERR_INVALID;
    const long per = (long)t_out * n_mels * n_stack;
    dim3 grid((unsigned)((per + 255) / 256), B);
    fe_log_stack_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mel, out, rows_per_utt, n_frames, seq_len,
                                                                                  n_mels, n_stack, t_out, take_log);
    EB_CHECK_LAUNCH();
    return EB_OK;
}
