// Depthwise Conv1d (kernel 31, zero pad 15) along time on token-major (B, T, d) bf16 tensors, with the
// BatchNorm partial statistics fused into the forward, and (backward) the weight / bias gradient and the
// GLU backward fused into one kernel.  Halo tile staged in shared memory, sliding register window.
// Replaces (reference): model/conformer.py:63-67,83 depthwise_conv (cuDNN depthwise + 2 transposes),
//   the GLU backward of :82, and the statistics pass of BatchNorm1d :84.
#include "common.cuh"

namespace {

constexpr int KW = 31, HALO = 15;
constexpr int TT = 64;        // time steps per CTA
constexpr int CC = 64;        // channels per CTA (32 bf16x2 lanes)
constexpr int OPT = 8;        // outputs per thread
constexpr int DW_THREADS = 256;  // 32 channel-pair lanes x 8 strips
constexpr int ROWS = TT + 2 * HALO;

__device__ __forceinline__ void load_tile(bf162 (*tile)[CC / 2], const bf16* __restrict__ src, int b, int T, int d, int t0,
                                          int c0) {
  // ROWS x 128 B, 8 threads (16 B each) per row
  for (int i = threadIdx.x; i < ROWS * 8; i += DW_THREADS) {
    const int r = i >> 3, v = i & 7;
    const int t = t0 - HALO + r;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (t >= 0 && t < T) val = *reinterpret_cast<const uint4*>(src + ((long long)b * T + t) * d + c0 + v * 8);
    *reinterpret_cast<uint4*>(&tile[r][v * 4]) = val;
  }
}

// out[t] = bias + sum_k in[t + k - 15] * w[k]      (FLIP: w index 30 - k, used for the input gradient)
template <bool FLIP>
__device__ __forceinline__ void conv_strip(const bf162 (*tile)[CC / 2], int lane, int strip, const float2* wreg, float2* acc) {
#pragma unroll
  for (int j = 0; j < OPT + KW - 1; ++j) {
    const float2 v = __bfloat1622float2(tile[strip * OPT + j][lane]);
#pragma unroll
    for (int o = 0; o < OPT; ++o) {
      const int k = j - o;
      if (k >= 0 && k < KW) {
        const float2 w = wreg[FLIP ? (KW - 1 - k) : k];
        acc[o].x = fmaf(v.x, w.x, acc[o].x);
        acc[o].y = fmaf(v.y, w.y, acc[o].y);
      }
    }
  }
}

// weight (d, 31) fp32 [reference layout (d,1,31)], bias (d)
__global__ void __launch_bounds__(DW_THREADS) dwconv_fwd_kernel(const bf16* __restrict__ u, int T, int d,
                                                                const float* __restrict__ weight,
                                                                const float* __restrict__ bias, bf16* __restrict__ out,
                                                                float* __restrict__ bn_partial) {
  __shared__ __align__(16) bf162 tile[ROWS][CC / 2];
  __shared__ float red[8][CC][2];
  const int c0 = blockIdx.x * CC, t0 = blockIdx.y * TT, b = blockIdx.z;
  const int lane = threadIdx.x & 31, strip = threadIdx.x >> 5;
  load_tile(tile, u, b, T, d, t0, c0);
  float2 wreg[KW];
  const int ch = c0 + 2 * lane;
#pragma unroll
  for (int k = 0; k < KW; ++k) wreg[k] = make_float2(weight[ch * KW + k], weight[(ch + 1) * KW + k]);
  const float2 bv = make_float2(bias[ch], bias[ch + 1]);
  __syncthreads();
  float2 acc[OPT];
#pragma unroll
  for (int o = 0; o < OPT; ++o) acc[o] = bv;
  conv_strip<false>(tile, lane, strip, wreg, acc);
  float2 s = make_float2(0, 0), ss = make_float2(0, 0);
#pragma unroll
  for (int o = 0; o < OPT; ++o) {
    const int t = t0 + strip * OPT + o;
    if (t < T) {
      bf162 q = __floats2bfloat162_rn(acc[o].x, acc[o].y);
      *reinterpret_cast<bf162*>(out + ((long long)b * T + t) * d + ch) = q;
      const float2 r = __bfloat1622float2(q);  // statistics of what BatchNorm will actually read
      s.x += r.x; s.y += r.y;
      ss.x += r.x * r.x; ss.y += r.y * r.y;
    }
  }
  if (bn_partial != nullptr) {
    red[strip][2 * lane][0] = s.x; red[strip][2 * lane][1] = ss.x;
    red[strip][2 * lane + 1][0] = s.y; red[strip][2 * lane + 1][1] = ss.y;
    __syncthreads();
    if (threadIdx.x < CC * 2) {
      const int c = threadIdx.x >> 1, which = threadIdx.x & 1;
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) a += red[r][c][which];
      const long long part = (long long)b * gridDim.y + blockIdx.y;
      bn_partial[(part * d + c0 + c) * 2 + which] = a;
    }
  }
}

// Backward, kernel A: du = corr(dw, flipped weight) with the GLU backward fused -> dab (M, 2d) (or du).
__global__ void __launch_bounds__(DW_THREADS) dwconv_bwd_data_kernel(const bf16* __restrict__ dwv, const bf16* __restrict__ ab,
                                                                     int T, int d, const float* __restrict__ weight,
                                                                     bf16* __restrict__ dab, bf16* __restrict__ du_out) {
  __shared__ __align__(16) bf162 tile[ROWS][CC / 2];
  const int c0 = blockIdx.x * CC, t0 = blockIdx.y * TT, b = blockIdx.z;
  const int lane = threadIdx.x & 31, strip = threadIdx.x >> 5;
  const int ch = c0 + 2 * lane;
  load_tile(tile, dwv, b, T, d, t0, c0);
  float2 wreg[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) wreg[k] = make_float2(weight[ch * KW + k], weight[(ch + 1) * KW + k]);
  __syncthreads();
  float2 acc[OPT];
#pragma unroll
  for (int o = 0; o < OPT; ++o) acc[o] = make_float2(0.f, 0.f);
  conv_strip<true>(tile, lane, strip, wreg, acc);
#pragma unroll
  for (int o = 0; o < OPT; ++o) {
    const int t = t0 + strip * OPT + o;
    if (t < T) {
      const long long row = (long long)b * T + t;
      if (ab != nullptr) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const bf162*>(ab + row * 2 * d + ch));
        const float2 gt = __bfloat1622float2(*reinterpret_cast<const bf162*>(ab + row * 2 * d + d + ch));
        const float s0 = sigmoidf_(gt.x), s1 = sigmoidf_(gt.y);
        *reinterpret_cast<bf162*>(dab + row * 2 * d + ch) = __floats2bfloat162_rn(acc[o].x * s0, acc[o].y * s1);
        *reinterpret_cast<bf162*>(dab + row * 2 * d + d + ch) =
            __floats2bfloat162_rn(acc[o].x * a.x * s0 * (1.f - s0), acc[o].y * a.y * s1 * (1.f - s1));
      }
      if (du_out != nullptr)
        *reinterpret_cast<bf162*>(du_out + row * d + ch) = __floats2bfloat162_rn(acc[o].x, acc[o].y);
    }
  }
}

// Backward, kernel B: dweight[c][k] += sum_t dw[t,c] * u[t+k-15,c], dbias[c] += sum_t dw[t,c].
// lane = channel pair, warp w = taps 4w..4w+3 (sliding 4-value window over u); a CTA walks a segment of time tiles
// with the accumulators in registers and issues its atomics once.
__global__ void __launch_bounds__(DW_THREADS) dwconv_bwd_weight_kernel(const bf16* __restrict__ dwv, const bf16* __restrict__ u,
                                                                       int T, int d, float* __restrict__ dweight,
                                                                       float* __restrict__ dbias, int tiles_per_cta) {
  __shared__ __align__(16) bf162 utile[ROWS][CC / 2];
  __shared__ __align__(16) bf162 gtile[TT][CC / 2];
  const int c0 = blockIdx.x * CC, b = blockIdx.z;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ch = c0 + 2 * lane;
  const int tap0 = 4 * w;
  float2 acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
  float2 gsum = make_float2(0.f, 0.f);
  const int ntiles = (T + TT - 1) / TT;
  const int tile_begin = blockIdx.y * tiles_per_cta, tile_end = min(ntiles, tile_begin + tiles_per_cta);
  for (int ti = tile_begin; ti < tile_end; ++ti) {
    const int t0 = ti * TT;
    __syncthreads();
    load_tile(utile, u, b, T, d, t0, c0);
    for (int i = threadIdx.x; i < TT * 8; i += DW_THREADS) {
      const int r = i >> 3, v = i & 7;
      const int t = t0 + r;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (t < T) val = *reinterpret_cast<const uint4*>(dwv + ((long long)b * T + t) * d + c0 + v * 8);
      *reinterpret_cast<uint4*>(&gtile[r][v * 4]) = val;
    }
    __syncthreads();
    // u row for (t, tap) is t + tap (the tile starts at t0 - 15)
    float2 win[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) win[k + 1] = __bfloat1622float2(utile[tap0 + k][lane]);
#pragma unroll 8
    for (int t = 0; t < TT; ++t) {
      win[0] = win[1]; win[1] = win[2]; win[2] = win[3];
      const int r = min(t + tap0 + 3, ROWS - 1);
      win[3] = __bfloat1622float2(utile[r][lane]);
      const float2 g = __bfloat1622float2(gtile[t][lane]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[k].x = fmaf(g.x, win[k].x, acc[k].x);
        acc[k].y = fmaf(g.y, win[k].y, acc[k].y);
      }
      if (w == 7) { gsum.x += g.x; gsum.y += g.y; }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int tap = tap0 + k;
    if (tap < KW) {
      atomicAdd(dweight + (long long)ch * KW + tap, acc[k].x);
      atomicAdd(dweight + (long long)(ch + 1) * KW + tap, acc[k].y);
    }
  }
  if (w == 7 && dbias != nullptr) {
    atomicAdd(dbias + ch, gsum.x);
    atomicAdd(dbias + ch + 1, gsum.y);
  }
}

}  // namespace

extern "C" int tasr_dwconv_bn_parts(int B, int T) { return B * cdiv(T, TT); }

extern "C" int tasr_dwconv31_fwd(const void* u, int B, int T, int d, const float* weight, const float* bias,
                                 void* out, float* bn_partial, tasr_stream_t stream) {
  if (d % CC || B <= 0 || T <= 0) return TASR_ERR_SHAPE;
  dim3 grid(d / CC, cdiv(T, TT), B);
  dwconv_fwd_kernel<<<grid, DW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(u), T, d, weight, bias, reinterpret_cast<bf16*>(out), bn_partial);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_dwconv31_bwd(const void* dw, const void* u, const void* ab, int B, int T, int d,
                                 const float* weight, void* dab, void* du, float* dweight, float* dbias,
                                 tasr_stream_t stream) {
  if (d % CC || B <= 0 || T <= 0) return TASR_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ntiles = cdiv(T, TT);
  dim3 grid_a(d / CC, ntiles, B);
  dwconv_bwd_data_kernel<<<grid_a, DW_THREADS, 0, st>>>(reinterpret_cast<const bf16*>(dw), reinterpret_cast<const bf16*>(ab), T, d,
                                                        weight, reinterpret_cast<bf16*>(dab), reinterpret_cast<bf16*>(du));
  TASR_CHECK_LAUNCH();
  int nseg = 592 / max(1, B * (d / CC));  // ~4 CTAs per SM, otherwise as few segments (= atomics) as possible
  nseg = max(1, min(nseg, ntiles));
  const int tiles_per_cta = cdiv(ntiles, nseg);
  dim3 grid_b(d / CC, cdiv(ntiles, tiles_per_cta), B);
  dwconv_bwd_weight_kernel<<<grid_b, DW_THREADS, 0, st>>>(reinterpret_cast<const bf16*>(dw), reinterpret_cast<const bf16*>(u), T, d,
                                                          dweight, dbias, tiles_per_cta);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
