/* include/edgedict_b200.h -- C ABI of libedgedict_b200.so (sm_100a).
 *
 * Plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in
 * _host; every call is asynchronous on `stream` (a cudaStream_t passed as void*) and returns
 * 0 on success, 2 for invalid arguments, 3 for a CUDA error.  No call allocates memory:
 * callers own all buffers (same ownership rule as warp-transducer, README.md:36-37).
 *
 * Each entry point names the piece of the reference it replaces (paths relative to
 * /root/reference).  The reference has no FFI for the model path (it calls torch.nn modules),
 * so those entry points mirror the module boundaries of rnnt/models.py.
 */
#pragma once
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNN-T loss, device-resident costs, no host sync ------------------------------------
 * replaces warp-transducer/include/detail/gpu_rnnt.h:82-215 (GpuRNNT::compute_cost_and_score)
 * and warprnnt_pytorch/__init__.py:10-50 (_RNNT.forward/backward).
 *   logits [B,maxT,maxU,V] (dtype_size 4|8), labels [B,maxU-1] int32, xlen/ylen [B] int32. */
size_t eb_rnnt_workspace_bytes(int B, int maxT, int maxU, int dtype_size);
int eb_rnnt_loss_fwd(const void* logits, const int* labels, const int* xlen, const int* ylen,
                     int B, int maxT, int maxU, int V, int blank, int dtype_size,
                     void* workspace, void* costs_dev /* [B], may be NULL */, int need_beta,
                     void* stream);
/* grads = d(sum_b gscale[b]*cost_b)/d logits * host_scale; grads may alias logits (in place);
 * grads_bf16: write bf16 instead of fp32 (dtype_size 4 only). */
int eb_rnnt_loss_bwd(const void* logits, void* grads, int grads_bf16, const int* labels,
                     const int* xlen, const int* ylen, int B, int maxT, int maxU, int V, int blank,
                     int dtype_size, void* workspace, const void* gscale_dev /* [1]|[B]|NULL */,
                     int gscale_per_batch, double host_scale, void* stream);
/* bf16-mode fused path: the joint's output GEMM writes bf16 logits AND the softmax statistics
 * (eb_joint_logits_lse below), then only the lattice runs, and the gradient is taken on bf16 logits. */
int eb_rnnt_loss_lattice(const int* xlen, const int* ylen, int B, int maxT, int maxU, void* workspace,
                         float* costs_dev, int need_beta, void* stream);
int eb_rnnt_loss_bwd_bf16(const void* logits16, void* grads16, const int* labels, const int* xlen, const int* ylen,
                          int B, int maxT, int maxU, int V, int blank, void* workspace, const float* gscale_dev,
                          int gscale_per_batch, double host_scale, void* stream);
int eb_rnnt_workspace_views(void* workspace, int B, int maxT, int maxU, int dtype_size,
                            void** denom, void** alphas, void** betas, void** ll_fwd, void** ll_bwd);

/* ---- fp32 GEMM (parity mode of every Linear / LSTM input projection) ----------------------
 * replaces the cuBLAS/MKL calls behind nn.Linear and nn.LSTM's input GEMM (rnnt/models.py:45-46,
 * 129,148,163-167).  C = alpha*A'B' + beta*C + bias[n], A'(m,k)=A[m*sam+k*sak],
 * B'(k,n)=B[k*sbk+n*sbn]. */
int eb_gemm_f32(const float* A, long sam, long sak, const float* B, long sbk, long sbn, float* C,
                long ldc, const float* bias, int M, int N, int K, float alpha, float beta,
                void* stream);

/* ---- bf16 tensor-core GEMM (tcgen05 + TMA + TMEM), fp32 accumulate ------------------------
 * same call sites in bf16 mode.  A: [M,K] (a_mn_major=0, K contiguous) or [K,M] (a_mn_major=1);
 * B: [N,K] (b_mn_major=0) or [K,N] (b_mn_major=1); C row-major [M,N] fp32 or bf16;
 * C = A*B (+ bias[n]) (+ C when accumulate).  Pointers 16-byte aligned, contiguous dim % 8 == 0. */
int eb_gemm_bf16(const void* A, int a_mn_major, const void* B, int b_mn_major, void* C, int c_bf16,
                 const float* bias, int accumulate, long M, int N, long K, void* stream);

/* eb_gemm_bf16 with launch flags.  EB_GEMM_CORESIDENT (A and B K-major only): a 115 KB / 192-thread
 * configuration that shares an SM with one CTA of a persistent recurrent kernel (eb_lstm_tc_fwd), used by the
 * layer-wavefront schedule of the encoder stack (functional.LSTMStack). */
#define EB_GEMM_CORESIDENT 1
int eb_gemm_bf16_ex(const void* A, int a_mn_major, const void* B, int b_mn_major, void* C, int c_bf16,
                    const float* bias, int accumulate, long M, int N, long K, int flags, void* stream);

/* cta_group::2 tiles (two CTAs of a cluster share one 256 x 256 tile, each staging half of the B operand) for the
 * bf16-output GEMMs with K-major A -- the joint's logits (+LSE) and d-hidden products -- and, on request only, for
 * split-K weight gradients with both operands MN-major.  mode -1 = automatic (bf16-output products with enough 256-row
 * blocks for every CTA pair), 0 = never, 1 = whenever legal; returns the previous mode.  Default: EDGEDICT_GEMM_PAIR. */
int eb_gemm_pair_mode(int mode);

/* C16[M,N] = bf16((A B) * (1 - hid16[M,N]^2)): the joint's d-hidden GEMM with the derivative of Joint.forward's Tanh
 * (rnnt/models.py:164) applied in the epilogue, so that d(pre-activation) leaves the GEMM directly. */
int eb_gemm_bf16_dtanh(const void* A, int a_mn_major, const void* B, int b_mn_major, void* C16, const void* hid16,
                       long M, int N, long K, void* stream);

/* joint output layer + softmax statistics in one GEMM (bf16 mode): replaces the second Linear of Joint
 * (rnnt/models.py:165) together with reduce_max/reduce_exp (warp-transducer reduce.h:45-104) and the
 * blank/label gathers of the lattice kernels.  denom/lpb/lpl: the first three arrays of the loss workspace. */
int eb_joint_logits_lse(const void* hidden16, const void* w2_16, const float* b2, void* logits16, const int* labels,
                        const int* xlen, const int* ylen, float* denom, float* lpb, float* lpl, int B, int maxT,
                        int maxU, int V, int J, int blank, void* stream);

/* ---- LSTM layer, recurrent part (persistent kernel) ---------------------------------------
 * replaces the time loop of nn.LSTM (rnnt/models.py:45-46,64-65,145-147,154-155).
 * xg [B,T,4H] = W_ih x + b_ih + b_hh (gate order i|f|g|o); whh [4H,H]; h0/c0 may be NULL (zeros).
 * gates_save [B,T,4H] / cseq_save [B,T,H] may be NULL for inference.  scratch: zero-size-checked
 * by eb_lstm_scratch_bytes(B,H). */
size_t eb_lstm_scratch_bytes(int B, int H);
int eb_lstm_seq_fwd(const float* xg, const float* whh, const float* h0, const float* c0, float* y,
                    float* hT, float* cT, float* gates_save, float* cseq_save, void* scratch, int B,
                    int T, int H, void* stream);
/* BPTT: dgates [B,T,4H] (may alias gates) = d loss / d gate pre-activations; dh0/dc0 [B,H] out. */
int eb_lstm_seq_bwd(const float* dy, const float* gates, const float* cseq, const float* c0,
                    const float* whh, const float* dhT, const float* dcT, float* dgates, float* dh0,
                    float* dc0, void* scratch, int B, int T, int H, void* stream);

/* ---- LSTM layer on tensor cores (bf16 mode; H % 64 == 0, H <= 1024) ---------------------------
 * same contract as eb_lstm_seq_fwd/bwd with bf16 recurrent operands (fp32 accumulation, fp32 cell
 * state): whh16 [4H,H] bf16, whhT16 [H,4H] bf16 (= W_hh^T), y16 optional bf16 copy of y, dg16
 * [B,T,4H] bf16 gate-preactivation gradients. */
int eb_lstm_tc_supported(int B, int H);
size_t eb_lstm_tc_scratch_bytes(int B, int H);
int eb_lstm_tc_set_trace(void* dev_buf, int steps);     /* debug: per-stage clock64 stamps of CTA 0 of the BPTT kernel */
int eb_lstm_tc_max_clusters(int H, int cluster_size);   /* co-resident clusters of the BPTT kernel (diagnostic) */
int eb_lstm_tc_fwd(const float* xg, const void* whh16, const float* h0, const float* c0, float* y, void* y16,
                   float* hT, float* cT, float* gates_save, float* cseq_save, void* scratch, int B, int T,
                   int H, void* stream);
int eb_lstm_tc_bwd(const float* dy, const float* gates, const float* cseq, const float* c0,
                   const void* whhT16, const float* dhT, const float* dcT, void* dg16, float* dh0, float* dc0,
                   void* scratch, int B, int T, int H, void* stream);
/* eb_lstm_tc_bwd over a time axis stored chunk-major (functional.LSTMStack's wavefront buffers): chunk c is a contiguous
 * [B, chunk_lens[c], D] block, the blocks follow each other; one launch walks all chunks (T = sum of chunk_lens,
 * nchunks <= 8, chunk_lens a HOST array) instead of one launch per chunk with the (dh, dc) carry through memory. */
int eb_lstm_tc_bwd_chunks(const float* dy, const float* gates, const float* cseq, const float* c0,
                          const void* whhT16, const float* dhT, const float* dcT, void* dg16, float* dh0, float* dc0,
                          void* scratch, int B, const int* chunk_lens, int nchunks, int H, void* stream);

/* ---- LSTM layer on tcgen05 tensor cores inside thread-block clusters (bf16 mode; H % 256 == 0, H <= 1024) ----
 * csrc/lstm_c4.cu: W_hh slices resident in shared memory, h / dG exchanged through L2 with TMA pulls, partial gate
 * sums reduced across a cluster through distributed shared memory.  Same cell semantics as eb_lstm_tc_* with a
 * CTA-private layout of the forward saves:
 *   gsave: post-activation gates in bf16, csave: cell states in fp32, sizes from eb_lstm_c4_{g,c}save_bytes;
 *   hprev16 [B,T,H] bf16 = h_{t-1} for every step (frame 0 = h0): the operand of the dW_hh GEMM (optional).
 * eb_lstm_c4_supported also checks that all clusters of a launch are co-resident (cooperative cluster launch). */
int eb_lstm_c4_supported(int B, int H);
int eb_lstm_c4_max_clusters(int H, int which);        /* diagnostic: co-resident clusters (0 fwd, 4 / 8 bwd) */
int eb_lstm_c4_bwd_cluster(int H);                      /* cluster size the BPTT kernel uses (8 / 4), 0 = cannot run */
int eb_lstm_c4_set_trace(void* dev_buf, int steps);     /* debug: per-stage clock64 stamps of CTA 0 ([steps][16] int64) */
size_t eb_lstm_c4_scratch_bytes(int B, int H);
size_t eb_lstm_c4_gsave_bytes(int B, int T, int H);
size_t eb_lstm_c4_csave_bytes(int B, int T, int H);
int eb_lstm_c4_fwd(const float* xg, const void* whh16, const float* h0, const float* c0, float* y, void* hprev16,
                   float* hT, float* cT, void* gsave, void* csave, float* gates_std, float* cseq_std, void* scratch,
                   int B, int T, int H, void* stream);   /* gates_std / cseq_std: optional saves in eb_lstm_tc_bwd's layout */
int eb_lstm_c4_bwd(const float* dy, const void* gsave, const void* csave, const float* c0, const void* whhT16,
                   const float* dhT, const float* dcT, void* dg16, float* dh0, float* dc0, void* scratch, int B,
                   int T, int H, void* stream);

/* ---- LayerNorm(x + res) fwd/bwd, TimeReduction, Embedding -------------------------------
 * rnnt/models.py:47,66-69,124 ; :21-29 ; :150-153.  *_bf16 outputs are optional side copies. */
int eb_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                     void* y_bf16, float* mean, float* rstd, long rows, int H, float eps, void* stream);
int eb_layernorm_bwd(const float* dy, const float* x, const float* res, const float* gamma,
                     const float* mean, const float* rstd, float* dz, float* dgamma_accum,
                     float* dbeta_accum, long rows, int H, void* stream);
int eb_time_reduce_fwd(const float* x, float* y, void* y_bf16, int B, int T, int H, void* stream);
int eb_time_reduce_bwd(const float* dy, float* dx, int B, int T, int H, void* stream);
int eb_embedding_fwd(const void* ids, int ids_are_int64, const float* W, float* out, void* out_bf16,
                     int B, int U, int E, int prepend_bos, int bos, void* stream);
int eb_embedding_bwd(const void* ids, int ids_are_int64, const float* dout, float* dW_accum, int B,
                     int U, int E, int prepend_bos, int bos, int pad, void* stream);

/* ---- Joint network pieces (rnnt/models.py:169-179) --------------------------------------
 * hidden[b,t,u,:] = tanh(ep[b,t,:] + dp[b,u,:]) with ep = W1e*h_enc + b1, dp = W1d*h_dec. */
int eb_joint_hidden_fwd(const float* ep, const float* dp, void* hidden, int hidden_bf16, int B, int T,
                        int U, int J, void* stream);
int eb_joint_hidden_bwd(void* dhidden_inout, const void* hidden, int is_bf16, float* dep, float* ddp,
                        int B, int T, int U, int J, void* stream);
/* the same two reductions when d(pre-activation) [B,T,U,J] bf16 is already available (eb_gemm_bf16_dtanh) */
int eb_joint_dpre_reduce(const void* dpre16, float* dep, float* ddp, int B, int T, int U, int J, void* stream);

/* ---- streaming greedy decode: one persistent kernel per audio chunk -----------------------------
 * replaces PytorchStreamDecoder.decode's Python loop (rnnt/stream.py:93-120).  The host builds a
 * phase program once (edgedict_b200/stream_engine.py) and launches it per chunk; see decode.cu. */
enum { EB_PH_LN = 0, EB_PH_PAIR = 1, EB_PH_LSTM = 2, EB_PH_LINEAR = 3, EB_PH_ARGMAX = 4, EB_PH_COPY = 5 };
typedef struct EbPhase {
    int32_t type, S, K1, K2, N, flags, ldx1, ldx2, ldw1, ldw2, ldy, aux, aux2, hist_ld, hist_col, pad_;
    const float *x1, *x2, *w1, *w2, *b1, *b2;
    float *y, *y2, *c;
    const int32_t* tok_in;
    int32_t* tok_out;
    int32_t* hist;
} EbPhase;
/* flags: 1 = tanh epilogue (LINEAR); 2 = x1 rows are embedding rows indexed by tok_in (LSTM);
 *        4 = masked update: streams whose tok_in equals aux (blank) keep their state (LSTM);
 *        8 = ARGMAX also accumulates log_softmax(x)[argmax] into y[s] (batched greedy decode). */
int eb_decode_phase_size(void);
int eb_decode_run(const void* phases_dev, int nphase, void* barrier_dev, int max_ctas, void* stream);

/* ---- reductions, casts, optimizer -------------------------------------------------------- */
int eb_colsum(const void* x, int x_bf16, float* out_accum, long rows, int N, void* stream);
int eb_cast_bf16(const float* x, void* y, long n, void* stream);
int eb_transpose_to_bf16(const void* x, int x_bf16, void* y, long rows, long cols, void* stream);
int eb_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
                 float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);
/* Adam / AdamW with gradient clip and loss-scale handling on the device (no host read of the norm):
 * sumsq = device scalar sum(g^2) of the unscaled bucket (eb_sumsq) or NULL; max_norm > 0 clips like
 * torch.nn.utils.clip_grad_norm_ (cli/baseline.py:239-245); a non-finite norm skips the step (loss-scale overflow);
 * adamw = 1 selects the reference's AdamW update (modules/optimizer.py:283-290). */
int eb_adam_step_ex(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, float grad_scale, const float* sumsq, float max_norm,
                    int adamw, void* stream);
int eb_sumsq(const float* x, long n, float* out_accum, void* stream);

/* ---- log-mel front end (the step before the path) ------------------------------
 * replaces FilterbankFeatures.forward (rnnt/features.py:126-164) + Downsample (rnnt/transforms.py:37-51).
 * eb_fe_preemph_pad : x[B,L] -> xp[B,Lp]: pre-emphasis (features.py:137-141) and the reflect padding of
 *                     torch.stft(center=True) by `pad` = n_fft/2 samples; positions >= L+2*pad are zero.  With Lp a
 *                     multiple of hop, frame g = b*(Lp/hop)+f starts at flat offset g*hop, so the STFT is
 *                     eb_gemm_f32 on a strided view (sam = hop) against the windowed DFT basis [n_fft, 2*nbins].
 * eb_fe_power       : spec[rows, re(nbins) | im(nbins)] -> power[rows, nbins]  (features.py:149)
 * eb_fe_log_stack   : mel[B*rows_per_utt, n_mels] -> out[B, t_out, n_mels*n_stack]: log(x + 1e-20)
 *                     (features.py:155-156), zero for frames >= seq_len (features.py:160-164) and for the
 *                     stacking pad (transforms.py:41-45); out is the [B,T,F] layout Encoder.forward takes. */
int eb_fe_preemph_pad(const float* x, float* xp, int B, int L, long Lp, int pad, float preemph,
                      int use_preemph, void* stream);
int eb_fe_power(const float* spec, float* power, long rows, int nbins, void* stream);
int eb_fe_log_stack(const float* mel, float* out, int B, int rows_per_utt, int n_frames, int seq_len,
                    int n_mels, int n_stack, int t_out, int take_log, void* stream);
/* SpecAugment masks (rnnt/transforms.py:53-147) in place on x [B, D1, D2]: spans int32 [B, nmask, 2] = [start, end)
 * along axis 1 (frequency) or 2 (time); masked elements := fill. */
int eb_fe_mask(float* x, const int* spans, int B, int D1, int D2, int nmask, int axis, float fill, void* stream);

#ifdef __cplusplus
}
#endif
