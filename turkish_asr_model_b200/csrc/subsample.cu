// Conv2d subsampler (two 3x3 / stride 2 / pad 1 convolutions + SiLU).
// Replaces (reference): model/conformer.py:150-155,177-183 (Conv2d(1,d)+SiLU, Conv2d(d,d)+SiLU,
//   permute/view) and their backward.
//
//   conv1 (K = 9, bandwidth bound) is never materialised: it is recomputed inside the producer of conv2's
//   im2col operand  col[(b,t2,f2)][(kh,kw,c)] = silu(conv1(x))[b, c, 2*t2-1+kh, 2*f2-1+kw]   (0 outside),
//   and conv2 itself runs on the tcgen05 GEMM (tasr_gemm_bf16, SiLU epilogue) with the weight packed to
//   (co, kh, kw, ci).  Backward: dCol = dZ2 * W2 (GEMM), then col2im + SiLU' + the conv1 weight/bias
//   gradient in one pass (dZ1 is never materialised either).
#include "common.cuh"

namespace {

constexpr int NT = 256;

// conv1 pre-activation for 8 consecutive channels at conv1-output position (h, w)
__device__ __forceinline__ void conv1_point(const float* __restrict__ xb, int T, int F, int h, int w,
                                            const float* __restrict__ w1s, const float* __restrict__ b1s, int d, int c0,
                                            float* z, float* xin) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int tt = 2 * h - 1 + i, ff = 2 * w - 1 + j;
      xin[i * 3 + j] = (tt >= 0 && tt < T && ff >= 0 && ff < F) ? xb[(long long)tt * F + ff] : 0.f;
    }
#pragma unroll
  for (int q = 0; q < 8; ++q) z[q] = b1s[c0 + q];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 wa = *reinterpret_cast<const float4*>(w1s + k * d + c0);
    const float4 wb = *reinterpret_cast<const float4*>(w1s + k * d + c0 + 4);
    const float xv = xin[k];
    z[0] = fmaf(xv, wa.x, z[0]); z[1] = fmaf(xv, wa.y, z[1]); z[2] = fmaf(xv, wa.z, z[2]); z[3] = fmaf(xv, wa.w, z[3]);
    z[4] = fmaf(xv, wb.x, z[4]); z[5] = fmaf(xv, wb.y, z[5]); z[6] = fmaf(xv, wb.z, z[6]); z[7] = fmaf(xv, wb.w, z[7]);
  }
}

// x (B, T, F) fp32; w1 (d, 1, 3, 3) fp32; col (B*T2*F2, 9*d) bf16
__global__ void __launch_bounds__(NT) conv1_im2col_kernel(const float* __restrict__ x, int B, int T, int F, int d,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,
                                                          int T1, int F1, int T2, int F2, bf16* __restrict__ col) {
  extern __shared__ float sh_w[];  // [9][d] weights, [d] bias
  float* w1s = sh_w;
  float* b1s = sh_w + 9 * d;
  for (int i = threadIdx.x; i < 9 * d; i += NT) {
    const int k = i / d, c = i - k * d;
    w1s[i] = w1[c * 9 + k];
  }
  for (int i = threadIdx.x; i < d; i += NT) b1s[i] = b1[i];
  __syncthreads();
  const int lanes_per_item = d >> 3;                 // threads per (pixel, tap)
  const int items_per_iter = NT / lanes_per_item;
  const int c0 = (threadIdx.x % lanes_per_item) << 3;
  const int sub = threadIdx.x / lanes_per_item;
  const long long total = (long long)B * T2 * F2 * 9;
  for (long long item = (long long)blockIdx.x * items_per_iter + sub; item < total;
       item += (long long)gridDim.x * items_per_iter) {
    const long long pix = item / 9;
    const int tap = (int)(item - pix * 9);
    const int f2 = (int)(pix % F2);
    const long long bt = pix / F2;
    const int t2 = (int)(bt % T2), b = (int)(bt / T2);
    const int h = 2 * t2 - 1 + tap / 3, w = 2 * f2 - 1 + tap % 3;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (h >= 0 && h < T1 && w >= 0 && w < F1) {
      float z[8], xin[9];
      conv1_point(x + (long long)b * T * F, T, F, h, w, w1s, b1s, d, c0, z, xin);
      o.x = pack_bf16x2(siluf_(z[0]), siluf_(z[1]));
      o.y = pack_bf16x2(siluf_(z[2]), siluf_(z[3]));
      o.z = pack_bf16x2(siluf_(z[4]), siluf_(z[5]));
      o.w = pack_bf16x2(siluf_(z[6]), siluf_(z[7]));
    }
    *reinterpret_cast<uint4*>(col + (pix * 9 + tap) * d + c0) = o;
  }
}

// dcol (B*T2*F2, 9*d) bf16 -> dW1 (d,1,3,3), db1 (d) (atomic accumulate)
__global__ void __launch_bounds__(NT) col2im_conv1_bwd_kernel(const bf16* __restrict__ dcol, const float* __restrict__ x,
                                                              int B, int T, int F, int d, const float* __restrict__ w1,
                                                              const float* __restrict__ b1, int T1, int F1, int T2, int F2,
                                                              float* __restrict__ dw1, float* __restrict__ db1) {
  extern __shared__ float sh_w[];  // [9][d] weights, [d] bias, then reduction buffer [10][d]
  float* w1s = sh_w;
  float* b1s = sh_w + 9 * d;
  float* red = b1s + d;
  for (int i = threadIdx.x; i < 9 * d; i += NT) {
    const int k = i / d, c = i - k * d;
    w1s[i] = w1[c * 9 + k];
  }
  for (int i = threadIdx.x; i < d; i += NT) b1s[i] = b1[i];
  for (int i = threadIdx.x; i < 10 * d; i += NT) red[i] = 0.f;
  __syncthreads();
  const int lanes_per_item = d >> 3;
  const int items_per_iter = NT / lanes_per_item;
  const int c0 = (threadIdx.x % lanes_per_item) << 3;
  const int sub = threadIdx.x / lanes_per_item;
  float gw[9][8];
  float gb[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    gb[q] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) gw[k][q] = 0.f;
  }
  const long long total = (long long)B * T1 * F1;
  for (long long pixel = (long long)blockIdx.x * items_per_iter + sub; pixel < total;
       pixel += (long long)gridDim.x * items_per_iter) {
    const int w = (int)(pixel % F1);
    const long long bh = pixel / F1;
    const int h = (int)(bh % T1), b = (int)(bh / T1);
    float g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = 0.f;
    // taps (kh, kw) with 2*t2 - 1 + kh == h and 2*f2 - 1 + kw == w
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int th = h + 1 - kh;
      if (th < 0 || (th & 1)) continue;
      const int t2 = th >> 1;
      if (t2 >= T2) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tw = w + 1 - kw;
        if (tw < 0 || (tw & 1)) continue;
        const int f2 = tw >> 1;
        if (f2 >= F2) continue;
        const long long pix2 = ((long long)b * T2 + t2) * F2 + f2;
        const uint4 u = *reinterpret_cast<const uint4*>(dcol + (pix2 * 9 + kh * 3 + kw) * d + c0);
        float2 p;
        p = unpack_bf16x2(u.x); g[0] += p.x; g[1] += p.y;
        p = unpack_bf16x2(u.y); g[2] += p.x; g[3] += p.y;
        p = unpack_bf16x2(u.z); g[4] += p.x; g[5] += p.y;
        p = unpack_bf16x2(u.w); g[6] += p.x; g[7] += p.y;
      }
    }
    float z[8], xin[9];
    conv1_point(x + (long long)b * T * F, T, F, h, w, w1s, b1s, d, c0, z, xin);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float dz = g[q] * silu_gradf_(z[q]);
      gb[q] += dz;
#pragma unroll
      for (int k = 0; k < 9; ++k) gw[k][q] = fmaf(dz, xin[k], gw[k][q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(&red[k * d + c0 + q], gw[k][q]);
    atomicAdd(&red[9 * d + c0 + q], gb[q]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * d; i += NT) {
    const int k = i / d, c = i - k * d;
    if (k < 9) atomicAdd(dw1 + c * 9 + k, red[i]);
    else atomicAdd(db1 + c, red[i]);
  }
}

// out[n][(k % q) * (K / q) + k / q] = bf16(in[n][k])      (conv2: q = 9, input_proj: q = F2)
__global__ void __launch_bounds__(NT) pack_weight_remap_kernel(const float* __restrict__ in, long long N, int K, int q,
                                                               bf16* __restrict__ out) {
  const int inner = K / q;
  const long long total = N * K;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const long long n = i / K;
    const int k = (int)(i - n * K);
    out[n * K + (long long)(k % q) * inner + k / q] = __float2bfloat16(in[i]);
  }
}

}  // namespace

extern "C" int tasr_conv1_im2col(const float* x, int B, int T, int F, int d, const float* w1, const float* b1, void* col,
                                 tasr_stream_t stream) {
  if (d % 8 || d > 2048 || (NT % (d / 8)) || B <= 0 || T <= 0 || F <= 0) return TASR_ERR_SHAPE;
  const int T1 = (T - 1) / 2 + 1, F1 = (F - 1) / 2 + 1, T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  const size_t sm = (size_t)10 * d * sizeof(float);
  const long long total = (long long)B * T2 * F2 * 9;
  const int per = NT / (d / 8);
  const int grid = (int)imin64((long long)148 * 16, (total + per - 1) / per);
  conv1_im2col_kernel<<<grid, NT, sm, reinterpret_cast<cudaStream_t>(stream)>>>(x, B, T, F, d, w1, b1, T1, F1, T2, F2,
                                                                                reinterpret_cast<bf16*>(col));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_col2im_conv1_bwd(const void* dcol, const float* x, int B, int T, int F, int d, const float* w1,
                                     const float* b1, float* dw1, float* db1, tasr_stream_t stream) {
  if (d % 8 || d > 1024 || (NT % (d / 8)) || B <= 0 || T <= 0 || F <= 0) return TASR_ERR_SHAPE;
  const int T1 = (T - 1) / 2 + 1, F1 = (F - 1) / 2 + 1, T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  const size_t sm = (size_t)20 * d * sizeof(float);
  const long long total = (long long)B * T1 * F1;
  const int per = NT / (d / 8);
  const int grid = (int)imin64((long long)148 * 4, (total + per - 1) / per);
  col2im_conv1_bwd_kernel<<<grid, NT, sm, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(dcol), x, B, T, F, d, w1, b1, T1, F1, T2, F2, dw1, db1);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_pack_weight_remap(const float* in, int64_t N, int K, int q, void* out, tasr_stream_t stream) {
  if (N <= 0 || K <= 0 || q <= 0 || K % q) return TASR_ERR_SHAPE;
  const int grid = (int)imin64((long long)148 * 8, (N * K + NT - 1) / NT);
  pack_weight_remap_kernel<<<grid, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, N, K, q, reinterpret_cast<bf16*>(out));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
