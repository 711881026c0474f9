// fp32 operand mode of the encoder forward (north_star: logits within 1e-4 of the reference in fp32 mode;
// reference: trainer/trainer.py:227-282 validate(), inference.py:101-128, and the CPU run of BASELINE configs[0]).
//
// The dense contractions stay on the tcgen05 bf16 main loop of gemm.cu: an fp32 operand x is expanded into bf16
// pieces  x = h0 + h1 (+ h2),  h0 = bf16(x), h1 = bf16(x - h0), ...  (each subtraction is exact in fp32), and the
// product A.B^T is evaluated as the sum of the piece products that matter,
//     3 terms:  a0 b0 + a0 b1 + a1 b0                      (relative error ~2^-17 per product)
//     6 terms:  ... + a0 b2 + a1 b1 + a2 b0                (~2^-24)
// by CONCATENATING the pieces along K:  A' = [a0 | a0 | a1 ...] (M, nterms*K),  B' = [b0 | b1 | b0 ...] (N, nterms*K),
// so that one ordinary bf16 GEMM with K' = nterms*K and fp32 accumulation in TMEM produces the fp32 result.  The
// kernels here build those operands (fusing the activation that precedes the contraction) and do the remaining
// element-wise / small-reduction pieces of the forward in IEEE fp32 (this file is compiled WITHOUT --use_fast_math).
// Throughput is not the point of this mode: it is the parity instrument and the validation / inference path when
// fp32 results are asked for.
#include "common.cuh"
#include <math.h>

namespace {
constexpr int NT = 256;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }

__device__ __forceinline__ void pieces3(float x, float* h) {
  const bf16 a = __float2bfloat16_rn(x);
  h[0] = __bfloat162float(a);
  float r = x - h[0];
  const bf16 b = __float2bfloat16_rn(r);
  h[1] = __bfloat162float(b);
  r -= h[1];
  h[2] = __bfloat162float(__float2bfloat16_rn(r));
}

// out[r][t*K + dst(c)] = piece[(terms >> 4t) & 15] of act(in[r][...c...])
//   act 0: x          1: silu(x)        2: silu(in[r][c]) * in[r][K + c]        3: in[r][c] * sigmoid(in[r][K + c])
//   remap_q > 0 (weights): dst(c) = (c % q) * (K / q) + c / q, else dst(c) = c
__global__ void __launch_bounds__(NT) split_terms_kernel(const float* __restrict__ in, long long M, int K, long long ld_in,
                                                         int act, int remap_q, uint32_t terms, int nterms,
                                                         bf16* __restrict__ out) {
  const long long total = M * K;
  const long long ldo = (long long)nterms * K;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const long long r = i / K;
    const int c = (int)(i - r * K);
    const float* row = in + r * ld_in;
    float x = row[c];
    if (act == 1) x = silu_f(x);
    else if (act == 2) x = silu_f(x) * row[K + c];
    else if (act == 3) x = x * sigmoid_f(row[K + c]);
    float h[3];
    pieces3(x, h);
    const int dc = remap_q > 0 ? (c % remap_q) * (K / remap_q) + c / remap_q : c;
    bf16* o = out + r * ldo + dc;
    for (int t = 0; t < nterms; ++t) o[(long long)t * K] = __float2bfloat16_rn(h[(terms >> (4 * t)) & 15u]);
  }
}

// conv1 (Conv2d(1, d, 3, stride 2, pad 1) + SiLU) -> y1 (B, T1, F1, d) fp32, channels last
__global__ void __launch_bounds__(NT) conv1_f32_kernel(const float* __restrict__ x, int B, int T, int F, int d, int T1, int F1,
                                                       const float* __restrict__ w1, const float* __restrict__ b1,
                                                       float* __restrict__ y1) {
  const long long total = (long long)B * T1 * F1 * d;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const int c = (int)(i % d);
    long long p = i / d;
    const int f1 = (int)(p % F1);
    p /= F1;
    const int t1 = (int)(p % T1);
    const int b = (int)(p / T1);
    const float* xb = x + (long long)b * T * F;
    float acc = b1[c];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int tt = 2 * t1 - 1 + kh;
      if (tt < 0 || tt >= T) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ff = 2 * f1 - 1 + kw;
        if (ff < 0 || ff >= F) continue;
        acc = fmaf(w1[c * 9 + kh * 3 + kw], xb[(long long)tt * F + ff], acc);
      }
    }
    y1[i] = silu_f(acc);
  }
}

// im2col of y1 for conv2 (3x3, stride 2, pad 1), expanded into bf16 piece terms:
//   out[(b,t2,f2)][t*9d + (kh*3+kw)*d + c] = piece of y1[b, 2 t2 - 1 + kh, 2 f2 - 1 + kw, c]   (0 outside)
__global__ void __launch_bounds__(NT) im2col_split_kernel(const float* __restrict__ y1, int B, int T1, int F1, int d, int T2,
                                                          int F2, uint32_t terms, int nterms, bf16* __restrict__ out) {
  const int K = 9 * d;
  const long long rows = (long long)B * T2 * F2;
  const long long total = rows * K;
  const long long ldo = (long long)nterms * K;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const long long r = i / K;
    const int k = (int)(i - r * K);
    const int tap = k / d, c = k - tap * d;
    const int kh = tap / 3, kw = tap - kh * 3;
    long long p = r;
    const int f2 = (int)(p % F2);
    p /= F2;
    const int t2 = (int)(p % T2);
    const int b = (int)(p / T2);
    const int t1 = 2 * t2 - 1 + kh, f1 = 2 * f2 - 1 + kw;
    float x = 0.f;
    if (t1 >= 0 && t1 < T1 && f1 >= 0 && f1 < F1) x = y1[(((long long)b * T1 + t1) * F1 + f1) * d + c];
    float h[3];
    pieces3(x, h);
    bf16* o = out + r * ldo + k;
    for (int t = 0; t < nterms; ++t) o[(long long)t * K] = __float2bfloat16_rn(h[(terms >> (4 * t)) & 15u]);
  }
}

// rotary position embedding on fp32 q|k columns, in place; cos_sin (>= T, 32, 2)
__global__ void __launch_bounds__(NT) rope_f32_kernel(float* __restrict__ qkv, long long M, int T, int ld, int rot_cols,
                                                      const float* __restrict__ cos_sin) {
  const int pairs_per_row = rot_cols / 2;
  const long long total = M * pairs_per_row;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const long long r = i / pairs_per_row;
    const int pi = (int)(i - r * pairs_per_row);
    const int head = pi / 32, j = pi - head * 32;
    const int t = (int)(r % T);
    const float c = cos_sin[((long long)t * 32 + j) * 2], s = cos_sin[((long long)t * 32 + j) * 2 + 1];
    float* p = qkv + r * ld + head * 64 + j;
    const float x1 = p[0], x2 = p[32];
    p[0] = x1 * c - x2 * s;
    p[32] = x2 * c + x1 * s;
  }
}

// multi-query attention core in fp32: softmax(q k^T / 8 over keys < L_b) v.  CTA = 16 queries of one (b, h).
constexpr int AQ = 16;
__global__ void __launch_bounds__(NT) mqa_f32_kernel(const float* __restrict__ qkv, int T, int H, int d,
                                                     const long long* __restrict__ key_len, float* __restrict__ ctx, int Tpad) {
  extern __shared__ float sm[];
  float* sq = sm;                 // [AQ][64]
  float* sinv = sq + AQ * 64;     // [AQ]
  float* S = sinv + AQ;           // [AQ][Tpad]
  const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
  const int ld = d + 128;
  const int L = key_len ? (int)min((long long)T, key_len[b]) : T;
  const float* base = qkv + (long long)b * T * ld;
  for (int i = threadIdx.x; i < AQ * 64; i += NT) {
    const int q = i >> 6, c = i & 63;
    sq[i] = (q0 + q < T) ? base[(long long)(q0 + q) * ld + h * 64 + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < L; j += NT) {
    float kr[64];
    const float4* kp = reinterpret_cast<const float4*>(base + (long long)j * ld + d);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float4 v = kp[c];
      kr[4 * c] = v.x; kr[4 * c + 1] = v.y; kr[4 * c + 2] = v.z; kr[4 * c + 3] = v.w;
    }
    for (int q = 0; q < AQ; ++q) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) acc = fmaf(sq[q * 64 + c], kr[c], acc);
      S[q * Tpad + j] = acc * 0.125f;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = warp; q < AQ; q += NT / 32) {
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) m = fmaxf(m, S[q * Tpad + j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float e = expf(S[q * Tpad + j] - m);
      S[q * Tpad + j] = e;
      s += e;
    }
    s = warp_sum(s);
    if (lane == 0) sinv[q] = (L > 0 && s > 0.f) ? 1.f / s : 0.f;
  }
  __syncthreads();
  const int q = threadIdx.x >> 4, c4 = (threadIdx.x & 15) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < L; ++j) {
    const float p = S[q * Tpad + j];
    const float4 v = *reinterpret_cast<const float4*>(base + (long long)j * ld + d + 64 + c4);
    acc.x = fmaf(p, v.x, acc.x); acc.y = fmaf(p, v.y, acc.y); acc.z = fmaf(p, v.z, acc.z); acc.w = fmaf(p, v.w, acc.w);
  }
  if (q0 + q < T) {
    const float inv = sinv[q];
    *reinterpret_cast<float4*>(ctx + ((long long)b * T + q0 + q) * d + h * 64 + c4) =
        make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

// depthwise Conv1d(k = 31, pad 15) over time + bias, fp32; optional BatchNorm partial sums per (b, 64-row tile)
constexpr int DW_ROWS = 64;
__global__ void dwconv31_f32_kernel(const float* __restrict__ u, int T, int d, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ part) {
  const int c = threadIdx.x;  // blockDim.x == d
  const int tile = blockIdx.x, b = blockIdx.y;
  float wk[31];
#pragma unroll
  for (int k = 0; k < 31; ++k) wk[k] = w[c * 31 + k];
  const float bc = bias[c];
  const float* ub = u + (long long)b * T * d;
  float s = 0.f, ss = 0.f;
  const int t0 = tile * DW_ROWS, t1 = min(T, t0 + DW_ROWS);
  for (int t = t0; t < t1; ++t) {
    float acc = bc;
#pragma unroll
    for (int k = 0; k < 31; ++k) {
      const int tt = t + k - 15;
      if (tt >= 0 && tt < T) acc = fmaf(wk[k], ub[(long long)tt * d + c], acc);
    }
    out[((long long)b * T + t) * d + c] = acc;
    s += acc;
    ss = fmaf(acc, acc, ss);
  }
  if (part != nullptr) {
    float* p = part + (((long long)b * gridDim.x + tile) * d + c) * 2;
    p[0] = s;
    p[1] = ss;
  }
}

// out = silu((w - mean) * rstd * gamma + beta), stats (d, 2) = (mean, rstd) from tasr_bn_finalize
__global__ void __launch_bounds__(NT) bn_silu_f32_kernel(const float* __restrict__ w, long long M, int d,
                                                         const float* __restrict__ stats, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ out) {
  const long long total = M * d;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const int c = (int)(i % d);
    const float y = (w[i] - stats[2 * c]) * stats[2 * c + 1] * gamma[c] + beta[c];
    out[i] = silu_f(y);
  }
}

// u[r][c] = ab[r][c] * sigmoid(ab[r][d + c])   (nn.GLU over channels, model/conformer.py:60,82)
__global__ void __launch_bounds__(NT) glu_f32_kernel(const float* __restrict__ ab, long long M, int d, float* __restrict__ u) {
  const long long total = M * d;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const long long r = i / d;
    const int c = (int)(i - r * d);
    u[i] = ab[r * 2 * d + c] * sigmoid_f(ab[r * 2 * d + d + c]);
  }
}

inline int grid_for(long long total) { return (int)imin64((long long)148 * 16, (total + NT - 1) / NT); }
}  // namespace

extern "C" int tasr_f32_split_terms(const float* in, int64_t M, int K, int64_t ld_in, int act, int remap_q, uint32_t terms,
                                    int nterms, void* out, tasr_stream_t stream) {
  if (M <= 0 || K <= 0 || nterms < 1 || nterms > 6 || act < 0 || act > 3 || (remap_q > 0 && K % remap_q)) return TASR_ERR_SHAPE;
  split_terms_kernel<<<grid_for(M * K), NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, M, K, ld_in, act, remap_q, terms, nterms, reinterpret_cast<bf16*>(out));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_f32_conv1(const float* x, int B, int T, int F, int d, const float* w1, const float* b1, float* y1,
                              tasr_stream_t stream) {
  if (B <= 0 || T <= 0 || F <= 0 || d <= 0) return TASR_ERR_SHAPE;
  const int T1 = (T - 1) / 2 + 1, F1 = (F - 1) / 2 + 1;
  conv1_f32_kernel<<<grid_for((long long)B * T1 * F1 * d), NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, B, T, F, d, T1, F1,
                                                                                                            w1, b1, y1);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_f32_im2col_split(const float* y1, int B, int T, int F, int d, uint32_t terms, int nterms, void* out,
                                     tasr_stream_t stream) {
  if (B <= 0 || T <= 0 || F <= 0 || d <= 0 || nterms < 1 || nterms > 6) return TASR_ERR_SHAPE;
  const int T1 = (T - 1) / 2 + 1, F1 = (F - 1) / 2 + 1, T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  im2col_split_kernel<<<grid_for((long long)B * T2 * F2 * 9 * d), NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      y1, B, T1, F1, d, T2, F2, terms, nterms, reinterpret_cast<bf16*>(out));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_f32_rope(float* qkv, int64_t M, int T, int ld, int rot_cols, const float* cos_sin, tasr_stream_t stream) {
  if (M <= 0 || T <= 0 || rot_cols <= 0 || rot_cols % 64 || rot_cols > ld) return TASR_ERR_SHAPE;
  rope_f32_kernel<<<grid_for(M * (rot_cols / 2)), NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(qkv, M, T, ld, rot_cols, cos_sin);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_f32_mqa_fwd(const float* qkv, int B, int T, int H, int d, const int64_t* key_lengths, float* ctx,
                                tasr_stream_t stream) {
  if (B <= 0 || T <= 0 || H <= 0 || d != H * 64) return TASR_ERR_SHAPE;
  const int Tpad = (T + 31) / 32 * 32 + 1;
  const size_t smem = (size_t)(AQ * 64 + AQ + (size_t)AQ * Tpad) * sizeof(float);
  if (smem > 227 * 1024) return TASR_ERR_SHAPE;  // T' <= ~3500 (140 s of audio)
  cudaError_t e = cudaFuncSetAttribute(mqa_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return tasr_set_cuda_error(e);
  dim3 grid(cdiv(T, AQ), H, B);
  mqa_f32_kernel<<<grid, NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      qkv, T, H, d, reinterpret_cast<const long long*>(key_lengths), ctx, Tpad);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_f32_dwconv_parts(int B, int T) { return B * cdiv(T, DW_ROWS); }

extern "C" int tasr_f32_dwconv31(const float* u, int B, int T, int d, const float* weight, const float* bias, float* out,
                                 float* bn_partial, tasr_stream_t stream) {
  if (B <= 0 || T <= 0 || d <= 0 || d > 1024) return TASR_ERR_SHAPE;
  dim3 grid(cdiv(T, DW_ROWS), B);
  dwconv31_f32_kernel<<<grid, d, 0, reinterpret_cast<cudaStream_t>(stream)>>>(u, T, d, weight, bias, out, bn_partial);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_f32_bn_silu(const float* w, int64_t M, int d, const float* stats, const float* gamma, const float* beta,
                                float* out, tasr_stream_t stream) {
  if (M <= 0 || d <= 0) return TASR_ERR_SHAPE;
  bn_silu_f32_kernel<<<grid_for(M * d), NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, M, d, stats, gamma, beta, out);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_f32_glu(const float* ab, int64_t M, int d, float* u, tasr_stream_t stream) {
  if (M <= 0 || d <= 0) return TASR_ERR_SHAPE;
  glu_f32_kernel<<<grid_for(M * d), NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ab, M, d, u);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
