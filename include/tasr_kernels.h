/*
 * libtasr_kernels.so — C-ABI of the B200 (sm_100a) hot path for Eminkorkut/Turkish-ASR-Model:
 *   raw 16 kHz waveform -> log-mel/CMVN -> Conv2d subsampler -> Conformer encoder (SwiGLU FFN, RoPE
 *   multi-query attention, GroupNorm, depthwise conv + BatchNorm) -> log-softmax + CTC loss and
 *   gradient -> clip + AdamW, plus greedy (argmax/collapse) decoding.
 *
 * The reference has no FFI of its own (it is pure PyTorch); the boundary it offers is its nn.Module /
 * callable API.  Every entry point below therefore cites the reference call site whose
 * library kernel(s) it replaces (paths relative to the reference checkout).
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless named h_*;
 *   - the library never allocates or frees device memory, never synchronises the device, launches
 *     only on the passed stream (CUDA-graph capturable);
 *   - returns 0 (TASR_OK) or a negative tasr_status; tasr_last_error() gives a message;
 *   - activations are token-major: (B, T', d) row-major == (M = B*T', d).
 */
#ifndef TASR_KERNELS_H_
#define TASR_KERNELS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tasr_stream_t; /* cudaStream_t */

enum tasr_status {
  TASR_OK = 0,
  TASR_ERR_SHAPE = -1,   /* unsupported / inconsistent shape */
  TASR_ERR_ALIGN = -2,   /* misaligned pointer or leading dimension */
  TASR_ERR_ARCH = -3,    /* device is not sm_100 */
  TASR_ERR_CUDA = -4,    /* launch or driver error; see tasr_last_error() */
  TASR_ERR_WORKSPACE = -5
};

const char* tasr_status_string(int status);
const char* tasr_last_error(void);
int tasr_version(void);
/* 0 when the current device is compute capability 10.x, TASR_ERR_ARCH otherwise. */
int tasr_check_device(void);
/* number of CUDA kernels this library has launched in this process (monotonic). */
uint64_t tasr_launch_count(void);
/* Optional device-resident counter added to every dropout seed at kernel run time (NULL to disable), so that a
 * captured CUDA graph draws fresh dropout masks on every replay.  Process-global configuration. */
int tasr_set_dropout_seed_ptr(const uint64_t* dev_ptr);
/* The pointer set by the last tasr_set_dropout_seed_ptr call (the owner clears it before freeing the counter). */
const uint64_t* tasr_get_dropout_seed_ptr(void);

/* ------------------------------------------------------------------------------------------------
 * Dense contractions on tcgen05 / TMEM (bf16 in, fp32 accumulate), operands staged by TMA.
 * Replaces: nn.Linear / 1x1 Conv1d / Conv2d-as-GEMM call sites —
 *   model/conformer.py:20,24 (FF), :81,86 (pointwise convs), :153 (conv2 after im2col), :185
 *   (input_proj), :209 (fc); model/attention.py:218-222,249 (q/k/v/out projections);
 *   and their autograd backward (dgrad / wgrad).
 *
 * C[m, n] = sum_k A(m, k) * B(n, k)
 *   a_mn_major == 0: A(m,k) = A[m*lda + k]   (K contiguous)      == 1: A(m,k) = A[k*lda + m]
 *   b_mn_major == 0: B(n,k) = B[n*ldb + k]                       == 1: B(n,k) = B[k*ldb + n]
 * so   forward  y = x W^T      : A = x (K-major),  B = W (out,in) (K-major)
 *      dgrad    dx = dy W      : A = dy (K-major), B = W (out,in) used MN-major (k = out, n = in)
 *      wgrad    dW = dy^T x    : A = dy MN-major (m = out, k = token), B = x MN-major (n = in)
 * Alignment: base pointers 16 B, lda/ldb multiples of 8 elements.
 * ---------------------------------------------------------------------------------------------- */
enum tasr_epilogue {
  TASR_EPI_STORE = 0,      /* out = alpha*(acc + bias[n])                      (bf16 or f32, out_f32)   */
  TASR_EPI_RESID = 1,      /* out_f32 = aux_f32 + alpha*dropout(acc + bias[n])                           */
  TASR_EPI_SWIGLU = 2,     /* dual-B: g|v -> out2 (bf16, N=2*n_half), out = dropout(silu(g)*v) (bf16)    */
  TASR_EPI_GLU = 3,        /* dual-B: a|b -> out2, out = a*sigmoid(b)                                    */
  TASR_EPI_SILU = 4,       /* z = acc+bias -> out2 (bf16), out = silu(z) (bf16)                          */
  TASR_EPI_SWIGLU_BWD = 5, /* acc = dh; aux = g|v; out(bf16, 2*n_half wide) = dg|dv                      */
  TASR_EPI_GLU_BWD = 6,    /* acc = du; aux = a|b; out = da|db                                           */
  TASR_EPI_SILU_BWD = 7,   /* out(bf16) = acc * silu'(aux)                                               */
  TASR_EPI_ATOMIC = 8,     /* out_f32[m*ldo + remap(n)] += alpha*acc  (split-K wgrad)                    */
  TASR_EPI_ROPE = 9        /* out(bf16) = rope(acc + bias): rotary embedding (model/attention.py:62-70) on the 64-wide
                              heads in columns [0, remap_p0) with position m % n_half, aux = cos|sin table (>= n_half,
                              32, 2) fp32; columns >= remap_p0 are stored unrotated (the fused q|k|v projection)  */
};

typedef struct tasr_gemm_args {
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  const void* A; /* bf16 */
  int64_t lda;
  const void* B; /* bf16 */
  int64_t ldb;
  int32_t epilogue; /* enum tasr_epilogue */
  int32_t out_f32;  /* STORE only: 1 -> out is float32, 0 -> bf16 */
  void* out;
  int64_t ldo;
  void* out2;
  int64_t ldo2;
  const float* bias; /* fp32, length N (2*n_half for dual-B); may be NULL */
  const void* aux;
  int64_t ldaux;
  float alpha;
  int32_t n_half;   /* dual-B modes: N passed is n_half; B rows [0,n_half) and [n_half, 2*n_half) */
  float drop_p;     /* 0 -> no dropout */
  uint64_t seed;
  int32_t split_k;  /* ATOMIC only: number of splits of the reduction (>=1) */
  int32_t remap_p0; /* ATOMIC only: if >0, column n is written at (n % p0) * p1 + n / p0 */
  int32_t remap_p1;
  float* colsum;    /* ATOMIC only, may be NULL: colsum[m] += alpha * sum_k A(m, k) -- for a wgrad (A = dy) this is the
                       bias gradient, produced by one extra N = 16 UMMA per k-step against a tile of ones (replaces a
                       separate column-sum pass over dy) */
} tasr_gemm_args;

int tasr_gemm_bf16(const tasr_gemm_args* args, tasr_stream_t stream);

/* One-time initialisation of immutable per-device constant tables (FFT twiddles).  Synchronous;
 * call once per process and device before the first kernel call (the Python binding does). */
int tasr_init(void);

/* ------------------------------------------------------------------------------------------------
 * Log-mel front-end, batched and fused.
 * Replaces: data/preprocessing.py:81-110 AudioPreprocessor.extract_features (torchaudio
 *   MelSpectrogram -> torch.stft/cuFFT + MelScale matmul, :98; AmplitudeToDB top_db=80, :101;
 *   CMVN _normalize :112-116) and the zero padding of data/dataset.py:309 (collate_fn).
 *   wave      (B, wave_ld) fp32, utterance b uses the first n_samples[b] samples (each needs > 200)
 *   window    (n_fft) fp32 periodic Hann; fb (n_fft/2+1, n_mels) fp32 mel filterbank (torchaudio's)
 *   ranges    (2*n_mels) int32 from tasr_mel_filter_ranges(fb)
 *   feats     (B, Tmax, n_mels) fp32 out; frames t >= 1 + n_samples[b]/hop are zero-filled
 * Only n_fft = 400, hop = 160 (the reference's configuration), n_mels <= 128.
 * ---------------------------------------------------------------------------------------------- */
int tasr_mel_filter_ranges(const float* fb, int n_mels, int32_t* ranges, tasr_stream_t stream);
size_t tasr_mel_workspace_bytes(int B, int n_mels);
int tasr_mel_forward(const float* wave, int64_t wave_ld, const int32_t* n_samples, int B, int Tmax,
                     const float* window, const float* fb, const int32_t* ranges, int n_mels, int n_fft,
                     int hop, int normalize, float* feats, void* workspace, size_t workspace_bytes,
                     tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm over (channels-in-group x all padded time) per sample, token-major (B, T, d) fp32 input.
 * Replaces: model/conformer.py:45-49 TransposeGroupNorm.forward (transpose + native_group_norm +
 *   transpose), 5 uses per block (:121,124,78,133,135), and its backward.
 *   stats (B, G, 2) fp32 = (mean, rstd) saved for backward.  out is bf16 (GEMM operand) or fp32.
 *   bwd: dres (B,T,d) fp32 = (accumulate ? dres : 0) + dx;  dgamma/dbeta (d) are accumulated (+=);
 *        cast_out (bf16, may be NULL) = bf16(cast_alpha * dropout_mask(cast_seed) * dres): the operand of the
 *        next backward GEMMs, written in the same pass (same mask function as tasr_cast_f32_bf16).
 * Requires d/G % 4 == 0 and d/4 a divisor of 256.
 * ---------------------------------------------------------------------------------------------- */
size_t tasr_groupnorm_workspace_bytes(int B, int T, int d);
int tasr_groupnorm_fwd(const float* x, int B, int T, int d, int G, float eps, const float* gamma, const float* beta,
                       void* out, int out_bf16, float* stats, void* workspace, size_t workspace_bytes,
                       tasr_stream_t stream);
int tasr_groupnorm_bwd(const void* dy, int dy_bf16, const float* x, int B, int T, int d, int G, const float* stats,
                       const float* gamma, float* dres, int accumulate, float* dgamma, float* dbeta, void* cast_out,
                       float cast_alpha, float cast_drop_p, uint64_t cast_seed, void* workspace,
                       size_t workspace_bytes, tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Depthwise Conv1d k=31 pad=15 along time (token-major bf16) with fused BatchNorm partial statistics;
 * BatchNorm1d (batch statistics over B*T incl. padding, running-stat update) + SiLU.
 * Replaces: model/conformer.py:83 depthwise_conv, :84 batch_norm, :85 swish (+ :82 GLU backward).
 *   weight (d, 31) fp32 == the reference (d,1,31) tensor; bn_partial (tasr_dwconv_bn_parts, d, 2) fp32.
 *   bn stats (d, 2) fp32 = (mean, rstd).
 *   dwconv bwd: dw (M,d) bf16 in; ab (M,2d) saved GLU input (may be NULL -> du written to `du`);
 *   dab (M,2d) bf16 out; dweight (d,31), dbias (d) accumulated (+=).  The two halves are independent kernels:
 *   dab == du == NULL runs only the weight/bias-gradient kernel, dweight == NULL only the data kernel (so that a
 *   caller can put the weight half, a leaf of the backward graph, on another stream).
 * ---------------------------------------------------------------------------------------------- */
int tasr_dwconv_bn_parts(int B, int T);
int tasr_dwconv31_fwd(const void* u, int B, int T, int d, const float* weight, const float* bias, void* out,
                      float* bn_partial, tasr_stream_t stream);
int tasr_dwconv31_bwd(const void* dw, const void* u, const void* ab, int B, int T, int d, const float* weight,
                      void* dab, void* du, float* dweight, float* dbias, tasr_stream_t stream);
int tasr_bn_finalize(const float* partial, int npart, int d, int64_t count, float eps, float momentum, int training,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked, float* stats,
                     tasr_stream_t stream);
int tasr_bn_silu_fwd(const void* w, int64_t M, int d, const float* stats, const float* gamma, const float* beta,
                     void* out, tasr_stream_t stream);
size_t tasr_bn_bwd_workspace_bytes(int64_t M, int d);
int tasr_bn_silu_bwd(const void* ds, const void* w, int64_t M, int d, const float* stats, const float* gamma,
                     const float* beta, void* dw, float* dgamma, float* dbeta, void* workspace,
                     size_t workspace_bytes, tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise helpers.
 *   cast:   out_bf16[i] = bf16(alpha * in[i] * dropout_mask(seed, i))   (dropout backward of
 *           model/conformer.py:25 + autocast casts)
 *   colsum: out[c] += sum_r in[r][c]                                    (bias gradients)
 *   rope:   in-place rotary embedding on the first rot_cols columns (64-wide heads) of (M, ld) bf16;
 *           cos_sin (T, 32, 2) fp32; row r has position r % T.  model/attention.py:62-70,228-230.
 * ---------------------------------------------------------------------------------------------- */
int tasr_cast_f32_bf16(const float* in, void* out, int64_t n, float alpha, float drop_p, uint64_t seed,
                       tasr_stream_t stream);
int tasr_colsum_bf16(const void* in, int64_t M, int N, int64_t ld, float* out, tasr_stream_t stream);
int tasr_rope_inplace(void* qkv, int64_t M, int T, int ld, int rot_cols, const float* cos_sin, int inverse,
                      tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-query flash attention (tcgen05), forward / backward.
 * Replaces: model/attention.py:233-245 (expand + scaled_dot_product_attention / _standard_attention).
 *   qkv (B*T, d+128) bf16 = q (H heads x 64) | k (64) | v (64), RoPE already applied to q and k.
 *   key_lengths (B) int64 device or NULL: keys t >= key_lengths[b] are masked (reference mask
 *   model/conformer.py:187-202).  A fully masked row yields 0.
 *   ctx (B*T, d) bf16; lse2 (B, H, T) fp32 (log2 domain) saved for backward.
 *   bwd: dqkv (B*T, d+128) bf16; with cos_sin != NULL the inverse RoPE is fused so that dqkv is the
 *   gradient of the un-rotated projections.
 * ---------------------------------------------------------------------------------------------- */
int tasr_mqa_attention_fwd(const void* qkv, int B, int T, int H, int d, const int64_t* key_lengths, float drop_p,
                           uint64_t seed, void* ctx, float* lse2, tasr_stream_t stream);
size_t tasr_mqa_attention_bwd_workspace_bytes(int B, int T, int H, int d);
int tasr_mqa_attention_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse2, int B, int T, int H,
                           int d, const int64_t* key_lengths, float drop_p, uint64_t seed, const float* cos_sin,
                           void* dqkv, void* workspace, size_t workspace_bytes, tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Weight re-packing for the subsampler GEMMs.
 * Replaces: the permute/view of model/conformer.py:183 (folded into the weight layout instead of the activations).
 *   pack_weight_remap: out_bf16[n][(k % q)*(K/q) + k/q] = in_f32[n][k]  (conv2: q=9; input_proj: q=F2)
 * ---------------------------------------------------------------------------------------------- */
int tasr_pack_weight_remap(const float* in, int64_t N, int K, int q, void* out, tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Conv2d subsampler as implicit GEMM (no im2col buffer; 4-D TMA gathers, tcgen05 main loop).
 * Replaces: model/conformer.py:150-155,177-183 and the autograd backward of both convolutions.
 *   conv1_fwd : x (B,T,F) fp32 -> y1 (B,T1,F1,d) bf16 NHWC = silu(conv1(x))
 *   conv2_fwd : y1, w2p (d, 9d) bf16 packed (co,kh,kw,ci), bias (d) -> z2 (pre-activation) and
 *               y2 = silu(z2), both (B,T2,F2,d) bf16 == the (B*T2, F2*d) operand of input_proj
 *   conv2_dgrad: dz2 -> dy1 (B,T1,F1,d) bf16 (one GEMM per output-parity class)
 *   conv2_wgrad: dW2 (d,d,3,3) fp32 += dz2^T (*) y1
 *   conv1_bwd : dy1, x -> dW1 (d,1,3,3), db1 (d) fp32 (+=); conv1's pre-activation is recomputed
 * T, F are the mel-feature dimensions; d % 128 == 0; F must reduce 80 -> 40 -> 20 style (F1 = 2*F2, F2 % 4 == 0).
 * ---------------------------------------------------------------------------------------------- */
int tasr_conv1_fwd(const float* x, int B, int T, int F, int d, const float* w1, const float* b1, void* y1,
                   tasr_stream_t stream);
int tasr_conv1_bwd(const void* dy1, const float* x, int B, int T, int F, int d, const float* w1, const float* b1,
                   float* dw1, float* db1, tasr_stream_t stream);
int tasr_conv2_fwd(const void* y1, int B, int T, int F, int d, const void* w2p, const float* bias, void* z2, void* y2,
                   tasr_stream_t stream);
int tasr_conv2_dgrad(const void* dz2, int B, int T, int F, int d, const void* w2p, void* dy1, tasr_stream_t stream);
int tasr_conv2_wgrad(const void* dz2, const void* y1, int B, int T, int F, int d, float* dw2, tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Augmentation (BASELINE configs[4]).
 *   resample_sinc: torchaudio.functional.resample(lowpass_filter_width=6, rolloff=0.99, sinc_interp_hann)
 *     as used by data/preprocessing.py:191-228 SpeedPerturbation; per utterance orig/new frequency ALREADY
 *     divided by their gcd; out length ceil(new*N/orig); y (B, y_ld) is zero beyond it up to max_out.
 *   specaugment: data/preprocessing.py:132-188; params (B, nmask, 3) int32 = (axis 0 freq | 1 time, start, end)
 *     drawn on the host like torchaudio's mask_along_axis; positions [start, end) are set to 0.0 in place.
 * ---------------------------------------------------------------------------------------------- */
int tasr_resample_sinc(const float* x, int64_t x_ld, const int32_t* n_in, const int32_t* orig_freq,
                       const int32_t* new_freq, int B, float* y, int64_t y_ld, int max_out, tasr_stream_t stream);
int tasr_specaugment(float* feats, int B, int T, int F, const int32_t* params, int nmask, const int64_t* frames,
                     tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused log-softmax + CTC loss (mean reduction, zero_infinity) + gradient w.r.t. the logits.
 * Replaces: trainer/trainer.py:167-173 (log_softmax + nn.CTCLoss(blank=0, zero_infinity=True)) and
 *   their backward.  logits (B,T,V) bf16 or fp32 with row pitch ld >= V elements (dlogits: same pitch); targets (B,Smax) int64 padded; lengths int64 (B),
 *   all on the device.  loss (1) fp32 = mean_b(nll_b / max(S_b,1)), infeasible samples contribute 0;
 *   nll (B) fp32 or NULL; dlogits (or NULL), same shape and row pitch as the logits =
 *   grad_scale * (softmax - occupancy) / (B * max(S_b,1)) for t < input_lengths[b], else 0.
 *   logits_bf16: 0 = fp32 logits and gradient, 1 = bf16 logits and gradient (the training path: the gradient is the
 *   bf16 operand of the classifier's backward GEMMs), 2 = bf16 logits, fp32 gradient (the arithmetic is fp32 in all
 *   three; 2 exposes it unrounded).
 * ---------------------------------------------------------------------------------------------- */
size_t tasr_ctc_workspace_bytes(int B, int T, int V, int Smax);
int tasr_ctc_loss_fwd_bwd(const void* logits, int logits_bf16, int64_t ld, int B, int T, int V, const int64_t* targets, int Smax,
                          const int64_t* input_lengths, const int64_t* target_lengths, int blank, float grad_scale,
                          float* loss, float* nll, void* dlogits, void* workspace, size_t workspace_bytes,
                          tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Global-norm clip + AdamW on flat fp32 buffers (also refreshes the bf16 shadow weights).
 * Replaces: trainer/trainer.py:189-195, main.py:106-110.
 *   hyper (9 floats, device): lr, beta1, beta2, eps, weight_decay, 1-beta1^t, 1-beta2^t, max_norm, grad_div
 *   sumsq (device double): sum of squared gradients (tasr_grad_sumsq, after all-reduce under DP).
 * ---------------------------------------------------------------------------------------------- */
int tasr_grad_sumsq(const float* g, int64_t n, double* out, tasr_stream_t stream);
int tasr_clip_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n, const float* hyper,
                    const double* sumsq, float* norm_out, tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Greedy CTC decode to token ids.  Replaces: utils/decoding.py:149,163; data/tokenizer.py:44-54.
 *   ids (B,T) int64 argmax; tokens (B,T) int64 collapsed (repeats merged, blanks dropped, -1 padded);
 *   out_len (B) int32.  lengths (B) int64 or NULL = frames to decode per utterance.
 * ---------------------------------------------------------------------------------------------- */
int tasr_argmax_collapse(const void* logits, int logits_bf16, int64_t ld, int B, int T, int V, const int64_t* lengths, int blank,
                         int64_t* ids, int64_t* tokens, int32_t* out_len, tasr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 operand mode of the encoder forward (csrc/fp32_mode.cu; north_star "1e-4 in fp32 mode").
 * Replaces, for callers that ask for fp32 results: model/conformer.py:172-211 and model/attention.py:195-251 as they
 * run WITHOUT autocast (trainer/trainer.py:227-282 validate on CPU, inference.py:101-128, BASELINE configs[0]).
 * The contractions still run on tasr_gemm_bf16: an fp32 operand is expanded into bf16 pieces x = h0 + h1 + h2 and the
 * product terms are concatenated along K (A' = [a0|a0|a1..], B' = [b0|b1|b0..]), K' = nterms * K.
 *   terms: 4 bits per term, term t uses piece (terms >> 4t) & 15; nterms in 1..6.
 *   split_terms: out (M, nterms*K) bf16 from in (M, K [or 2K for act 2/3]) fp32 with row pitch ld_in;
 *                act 0 none | 1 silu | 2 silu(in[:, :K]) * in[:, K:] | 3 in[:, :K] * sigmoid(in[:, K:]);
 *                remap_q > 0: destination column (c % q) * (K / q) + c / q (same packing as pack_weight_remap).
 *   conv1      : x (B,T,F) -> y1 (B,T1,F1,d) fp32 channels-last = silu(conv1(x)).
 *   im2col_split: y1 -> (B*T2*F2, nterms*9d) bf16, column t*9d + (kh*3+kw)*d + c (conv2's operand).
 *   rope       : in place on the first rot_cols (multiple of 64) fp32 columns of qkv (M, ld); cos_sin (>=T, 32, 2).
 *   mqa_fwd    : qkv (B*T, d+128) fp32 = [q heads | k | v] -> ctx (B*T, d) fp32; key_lengths (B) int64 or NULL.
 *   dwconv31   : u (B,T,d) fp32 -> out fp32 (+bias); bn_partial (tasr_f32_dwconv_parts(B,T), d, 2) sums or NULL.
 *   bn_silu    : out = silu((w - mean) * rstd * gamma + beta) with stats (d,2) from tasr_bn_finalize.
 * ---------------------------------------------------------------------------------------------- */
int tasr_f32_split_terms(const float* in, int64_t M, int K, int64_t ld_in, int act, int remap_q, uint32_t terms, int nterms,
                         void* out, tasr_stream_t stream);
int tasr_f32_conv1(const float* x, int B, int T, int F, int d, const float* w1, const float* b1, float* y1,
                   tasr_stream_t stream);
int tasr_f32_im2col_split(const float* y1, int B, int T, int F, int d, uint32_t terms, int nterms, void* out,
                          tasr_stream_t stream);
int tasr_f32_rope(float* qkv, int64_t M, int T, int ld, int rot_cols, const float* cos_sin, tasr_stream_t stream);
int tasr_f32_mqa_fwd(const float* qkv, int B, int T, int H, int d, const int64_t* key_lengths, float* ctx,
                     tasr_stream_t stream);
int tasr_f32_dwconv_parts(int B, int T);
int tasr_f32_dwconv31(const float* u, int B, int T, int d, const float* weight, const float* bias, float* out,
                      float* bn_partial, tasr_stream_t stream);
int tasr_f32_bn_silu(const float* w, int64_t M, int d, const float* stats, const float* gamma, const float* beta,
                     float* out, tasr_stream_t stream);
/* u (M, d) = ab[:, :d] * sigmoid(ab[:, d:]) on a dense (M, 2d) fp32 matrix (nn.GLU, model/conformer.py:60,82). */
int tasr_f32_glu(const float* ab, int64_t M, int d, float* u, tasr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TASR_KERNELS_H_ */
