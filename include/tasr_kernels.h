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
  TASR_EPI_ATOMIC = 8      /* out_f32[m*ldo + remap(n)] += alpha*acc  (split-K wgrad)                    */
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
} tasr_gemm_args;

int tasr_gemm_bf16(const tasr_gemm_args* args, tasr_stream_t stream);
/* Same contract, plain CUDA-core kernel.  Debug/triage aid for the tests only; never used by the
 * product path. */
int tasr_gemm_bf16_debug(const tasr_gemm_args* args, tasr_stream_t stream);

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

#ifdef __cplusplus
}
#endif
#endif /* TASR_KERNELS_H_ */
