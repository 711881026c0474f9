// Depthwise Conv1d (kernel 31, zero pad 15) along time on token-major (B, T, d) bf16 tensors, with the
// BatchNorm partial statistics fused into the forward, and (backward) the weight / bias gradient and the
// GLU backward fused into one kernel.  Halo tile staged in shared memory, sliding register window, packed fp32x2 FMAs.
// Replaces (reference): model/conformer.py:63-67,83 depthwise_conv (cuDNN depthwise + 2 transposes),
//   the GLU backward of :82, and the statistics pass of BatchNorm1d :84.
#include "common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

namespace {

constexpr int KW = 31, HALO = 15;
constexpr int TT = 128;       // time steps per tile
constexpr int CC = 64;        // channels per CTA (32 channel-pair lanes)
constexpr int OPT = 16;       // outputs per thread
constexpr int DW_THREADS = 256;  // 32 channel-pair lanes x 8 strips
constexpr int ROWS = TT + 2 * HALO;

// Stages the halo tile (ROWS x 128 B, 8 threads of 16 B per row) and the CTA's weights.  All global loads of both are
// issued before the first shared-memory store (fixed trip counts, values held in registers): a load -> store loop with a
// run-time bound is compiled into one exposed DRAM round trip per iteration, and staging the weights and the tile one
// after the other costs two round trips instead of one.
// The (d, 31) weight rows of this CTA's 64 channels are one contiguous run: read it coalesced and keep it as
// [tap][channel pair] so that a thread's 31 taps are conflict-free 8-byte reads.
__device__ __forceinline__ void load_tile_and_weights(bf162 (*tile)[CC / 2], float2 (*wsm)[CC / 2], const bf16* __restrict__ src,
                                                      const float* __restrict__ weight, int b, int T, int d, int t0, int c0) {
  constexpr int N = ROWS * 8, IT = (N + DW_THREADS - 1) / DW_THREADS;
  constexpr int NW = CC * KW, ITW = (NW + DW_THREADS - 1) / DW_THREADS;
  uint4 val[IT];
  float wv[ITW];
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int i = threadIdx.x + it * DW_THREADS;
    const int r = i >> 3, v = i & 7;
    const int t = t0 - HALO + r;
    val[it] = make_uint4(0, 0, 0, 0);
    if (i < N && t >= 0 && t < T) val[it] = *reinterpret_cast<const uint4*>(src + ((long long)b * T + t) * d + c0 + v * 8);
  }
#pragma unroll
  for (int it = 0; it < ITW; ++it) {
    const int i = threadIdx.x + it * DW_THREADS;
    wv[it] = i < NW ? weight[(long long)c0 * KW + i] : 0.f;
  }
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int i = threadIdx.x + it * DW_THREADS;
    if (i < N) *reinterpret_cast<uint4*>(&tile[i >> 3][(i & 7) * 4]) = val[it];
  }
  float* flat = reinterpret_cast<float*>(wsm);
#pragma unroll
  for (int it = 0; it < ITW; ++it) {
    const int i = threadIdx.x + it * DW_THREADS;
    const int c = i / KW, k = i - c * KW;
    if (i < NW) flat[(k * (CC / 2) + (c >> 1)) * 2 + (c & 1)] = wv[it];
  }
}

// out[t] = bias + sum_k in[t + k - 15] * w[k]      (FLIP: w index 30 - k, used for the input gradient)
// packed fp32x2 FMAs on the channel pair
template <bool FLIP>
__device__ __forceinline__ void conv_strip(const bf162 (*tile)[CC / 2], int lane, int strip, const float2* wreg, float2* acc) {
#pragma unroll
  for (int j = 0; j < OPT + KW - 1; ++j) {
    const float2 v = __bfloat1622float2(tile[strip * OPT + j][lane]);
#pragma unroll
    for (int o = 0; o < OPT; ++o) {
      const int k = j - o;
      if (k >= 0 && k < KW) acc[o] = __ffma2_rn(v, wreg[FLIP ? (KW - 1 - k) : k], acc[o]);
    }
  }
}

// weight (d, 31) fp32 [reference layout (d,1,31)], bias (d)
__global__ void __launch_bounds__(DW_THREADS, 3) dwconv_fwd_kernel(const bf16* __restrict__ u, int T, int d,
                                                                   const float* __restrict__ weight,
                                                                   const float* __restrict__ bias, bf16* __restrict__ out,
                                                                   float* __restrict__ bn_partial) {
  __shared__ __align__(16) bf162 tile[ROWS][CC / 2];
  __shared__ __align__(16) float2 wsm[KW][CC / 2];
  __shared__ float red[8][CC][2];
  const int c0 = blockIdx.x * CC, t0 = blockIdx.y * TT, b = blockIdx.z;
  const int lane = threadIdx.x & 31, strip = threadIdx.x >> 5;
  load_tile_and_weights(tile, wsm, u, weight, b, T, d, t0, c0);
  const int ch = c0 + 2 * lane;
  const float2 bv = make_float2(bias[ch], bias[ch + 1]);
  __syncthreads();
  float2 wreg[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) wreg[k] = wsm[k][lane];
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
// STAGE: the (a | gate) rows the GLU backward needs are fetched with cp.async into shared memory while the halo tile is
// staged and the taps run, so the epilogue does not wait on global loads issued after the arithmetic.
template <bool STAGE>
__global__ void __launch_bounds__(DW_THREADS, 3) dwconv_bwd_data_kernel(const bf16* __restrict__ dwv, const bf16* __restrict__ ab,
                                                                        int T, int d, const float* __restrict__ weight,
                                                                        bf16* __restrict__ dab, bf16* __restrict__ du_out) {
  __shared__ __align__(16) bf162 tile[ROWS][CC / 2];
  __shared__ __align__(16) float2 wsm[KW][CC / 2];
  extern __shared__ __align__(16) uint8_t dw_dyn[];
  bf162 (*abt)[2][CC / 2] = reinterpret_cast<bf162 (*)[2][CC / 2]>(dw_dyn);  // [TT][a | gate][channel pair]
  const int c0 = blockIdx.x * CC, t0 = blockIdx.y * TT, b = blockIdx.z;
  const int lane = threadIdx.x & 31, strip = threadIdx.x >> 5;
  const int ch = c0 + 2 * lane;
  if (STAGE && ab != nullptr) {
    // TT rows x 2 halves x 8 chunks of 16 B
#pragma unroll
    for (int it = 0; it < TT * 16 / DW_THREADS; ++it) {
      const int i = threadIdx.x + it * DW_THREADS;
      const int r = i >> 4, hv = (i >> 3) & 1, v = i & 7;
      const int t = t0 + r;
      if (t < T) {
        const bf16* src = ab + ((long long)b * T + t) * 2 * d + hv * d + c0 + v * 8;
        const uint32_t dst = smem_u32(&abt[r][hv][v * 4]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  load_tile_and_weights(tile, wsm, dwv, weight, b, T, d, t0, c0);
  __syncthreads();
  float2 wreg[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) wreg[k] = wsm[k][lane];
  float2 acc[OPT];
#pragma unroll
  for (int o = 0; o < OPT; ++o) acc[o] = make_float2(0.f, 0.f);
  conv_strip<true>(tile, lane, strip, wreg, acc);
  if (STAGE && ab != nullptr) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
#pragma unroll
  for (int o = 0; o < OPT; ++o) {
    const int tl = strip * OPT + o, t = t0 + tl;
    if (t < T) {
      const long long row = (long long)b * T + t;
      if (ab != nullptr) {
        float2 a, gt;
        if (STAGE) {
          a = __bfloat1622float2(abt[tl][0][lane]);
          gt = __bfloat1622float2(abt[tl][1][lane]);
        } else {
          a = __bfloat1622float2(*reinterpret_cast<const bf162*>(ab + row * 2 * d + ch));
          gt = __bfloat1622float2(*reinterpret_cast<const bf162*>(ab + row * 2 * d + d + ch));
        }
        const float s0 = fmaf(0.5f, tanh_approx(0.5f * gt.x), 0.5f), s1 = fmaf(0.5f, tanh_approx(0.5f * gt.y), 0.5f);
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
// lane = channel pair; warp w: tap group w & 3 (taps 8g..8g+7, a sliding 8-value window over u) and time half w >> 2 of
// each tile.  The tiles are converted to fp32 pairs once when they are staged, so the inner loop is 2 LDS.64 + 8 FFMA2
// per time step.  A CTA walks a segment of time tiles with the accumulators in registers and issues its atomics once.
constexpr int WG_TAPS = 8;
__global__ void __launch_bounds__(DW_THREADS) dwconv_bwd_weight_kernel(const bf16* __restrict__ dwv, const bf16* __restrict__ u,
                                                                       int T, int d, float* __restrict__ dweight,
                                                                       float* __restrict__ dbias, int tiles_per_cta) {
  extern __shared__ __align__(16) float2 dsm[];
  float2 (*utile)[CC / 2] = reinterpret_cast<float2 (*)[CC / 2]>(dsm);                       // [ROWS + 1]
  float2 (*gtile)[CC / 2] = reinterpret_cast<float2 (*)[CC / 2]>(dsm + (ROWS + 1) * (CC / 2));  // [TT]
  const int c0 = blockIdx.x * CC, b = blockIdx.z;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int tg = w & 3, th = w >> 2;
  const int tap0 = WG_TAPS * tg;
  float2 acc[WG_TAPS];
#pragma unroll
  for (int k = 0; k < WG_TAPS; ++k) acc[k] = make_float2(0.f, 0.f);
  float2 gsum = make_float2(0.f, 0.f);
  const int ntiles = (T + TT - 1) / TT;
  const int tile_begin = blockIdx.y * tiles_per_cta, tile_end = min(ntiles, tile_begin + tiles_per_cta);
  for (int ti = tile_begin; ti < tile_end; ++ti) {
    const int t0 = ti * TT;
    __syncthreads();
    // stage u rows t0-15 .. t0+TT+15 (+1 zero row: tap 31 does not exist) and dw rows t0 .. t0+TT-1 as fp32 pairs;
    // all loads first, then the conversions and stores
    constexpr int NU = (ROWS + 1) * 8, ITU = (NU + DW_THREADS - 1) / DW_THREADS, ITG = TT * 8 / DW_THREADS;
    uint4 uv[ITU], gv[ITG];
#pragma unroll
    for (int it = 0; it < ITU; ++it) {
      const int i = threadIdx.x + it * DW_THREADS;
      const int r = i >> 3, v = i & 7;
      const int t = t0 - HALO + r;
      uv[it] = make_uint4(0, 0, 0, 0);
      if (r < ROWS && t >= 0 && t < T) uv[it] = *reinterpret_cast<const uint4*>(u + ((long long)b * T + t) * d + c0 + v * 8);
    }
#pragma unroll
    for (int it = 0; it < ITG; ++it) {
      const int i = threadIdx.x + it * DW_THREADS;
      const int r = i >> 3, v = i & 7;
      const int t = t0 + r;
      gv[it] = make_uint4(0, 0, 0, 0);
      if (t < T) gv[it] = *reinterpret_cast<const uint4*>(dwv + ((long long)b * T + t) * d + c0 + v * 8);
    }
#pragma unroll
    for (int it = 0; it < ITU; ++it) {
      const int i = threadIdx.x + it * DW_THREADS;
      if (i < NU) {
        float4* dst = reinterpret_cast<float4*>(&utile[i >> 3][(i & 7) * 4]);
        const float2 p0 = unpack_bf16x2(uv[it].x), p1 = unpack_bf16x2(uv[it].y), p2 = unpack_bf16x2(uv[it].z), p3 = unpack_bf16x2(uv[it].w);
        dst[0] = make_float4(p0.x, p0.y, p1.x, p1.y);
        dst[1] = make_float4(p2.x, p2.y, p3.x, p3.y);
      }
    }
#pragma unroll
    for (int it = 0; it < ITG; ++it) {
      const int i = threadIdx.x + it * DW_THREADS;
      float4* dst = reinterpret_cast<float4*>(&gtile[i >> 3][(i & 7) * 4]);
      const float2 p0 = unpack_bf16x2(gv[it].x), p1 = unpack_bf16x2(gv[it].y), p2 = unpack_bf16x2(gv[it].z), p3 = unpack_bf16x2(gv[it].w);
      dst[0] = make_float4(p0.x, p0.y, p1.x, p1.y);
      dst[1] = make_float4(p2.x, p2.y, p3.x, p3.y);
    }
    __syncthreads();
    // u row for (t, tap) is t + tap (the tile starts at t0 - 15)
    const int tb = th * (TT / 2);
    float2 win[WG_TAPS];
#pragma unroll
    for (int k = 0; k < WG_TAPS - 1; ++k) win[k + 1] = utile[tb + tap0 + k][lane];
#pragma unroll 8
    for (int t = tb; t < tb + TT / 2; ++t) {
#pragma unroll
      for (int k = 0; k < WG_TAPS - 1; ++k) win[k] = win[k + 1];
      win[WG_TAPS - 1] = utile[t + tap0 + WG_TAPS - 1][lane];  // <= row ROWS (the zero row) for the last tap group
      const float2 g = gtile[t][lane];
#pragma unroll
      for (int k = 0; k < WG_TAPS; ++k) acc[k] = __ffma2_rn(g, win[k], acc[k]);
      if (tg == 0) gsum = __fadd2_rn(gsum, g);
    }
  }
  // reduce: the two time halves through shared memory, then the CTAs of the cluster (neighbouring utterances, same
  // channels) through distributed shared memory; each CTA of the cluster adds its share of the 64 x 32 sums (31 taps +
  // bias) to global memory, so the number of same-address atomics drops by the cluster size.
  __syncthreads();
  float* red = reinterpret_cast<float*>(dsm);  // [2][CC][32]
#pragma unroll
  for (int k = 0; k < WG_TAPS; ++k) {
    const int tap = tap0 + k;
    if (tap < KW) {
      red[(th * CC + 2 * lane) * 32 + tap] = acc[k].x;
      red[(th * CC + 2 * lane + 1) * 32 + tap] = acc[k].y;
    }
  }
  if (tg == 0) {
    red[(th * CC + 2 * lane) * 32 + 31] = gsum.x;
    red[(th * CC + 2 * lane + 1) * 32 + 31] = gsum.y;
  }
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  const int csz = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int per = (CC * 32) / csz;
  for (int i = threadIdx.x; i < per; i += DW_THREADS) {
    const int idx = crank * per + i;
    float a = 0.f;
    for (int r = 0; r < csz; ++r) {
      const float* rp = cluster.map_shared_rank(red, r);
      a += rp[idx] + rp[CC * 32 + idx];
    }
    const int c = idx >> 5, tap = idx & 31;
    if (tap < KW) atomicAdd(dweight + (long long)(c0 + c) * KW + tap, a);
    else if (dbias != nullptr) atomicAdd(dbias + c0 + c, a);
  }
  cluster.sync();  // keep every CTA's sums alive until its peers have read them
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
  if (dab != nullptr || du != nullptr) {  // data half (main chain)
    // TASR_DWCONV_STAGE=0: GLU operands read in the epilogue (A/B switch)
    static const bool stage = [] { const char* e = getenv("TASR_DWCONV_STAGE"); return !(e && e[0] == '0'); }();
    constexpr int ABSMEM = TT * 2 * (CC / 2) * (int)sizeof(bf162);
    if (stage && ab != nullptr) {
      static TasrPerDevice attr_done_a;
      if (!attr_done_a.get()) {
        cudaError_t e = cudaFuncSetAttribute(dwconv_bwd_data_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ABSMEM);
        if (e != cudaSuccess) return tasr_set_cuda_error(e);
        attr_done_a.set();
      }
      dwconv_bwd_data_kernel<true><<<grid_a, DW_THREADS, ABSMEM, st>>>(
          reinterpret_cast<const bf16*>(dw), reinterpret_cast<const bf16*>(ab), T, d, weight, reinterpret_cast<bf16*>(dab),
          reinterpret_cast<bf16*>(du));
    } else {
      dwconv_bwd_data_kernel<false><<<grid_a, DW_THREADS, 0, st>>>(
          reinterpret_cast<const bf16*>(dw), reinterpret_cast<const bf16*>(ab), T, d, weight, reinterpret_cast<bf16*>(dab),
          reinterpret_cast<bf16*>(du));
    }
    TASR_CHECK_LAUNCH();
  }
  if (dweight == nullptr) return TASR_OK;  // weight half is a leaf of the backward graph: callers may run it elsewhere
  int nseg = 1184 / max(1, B * (d / CC));  // up to ~8 CTAs per SM of work, otherwise as few segments (= atomics) as possible
  nseg = max(1, min(nseg, ntiles));
  const int tiles_per_cta = cdiv(ntiles, nseg);
  dim3 grid_b(d / CC, cdiv(ntiles, tiles_per_cta), B);
  constexpr int WSMEM = ((ROWS + 1) + TT) * (CC / 2) * (int)sizeof(float2);
  static TasrPerDevice attr_done;
  if (!attr_done.get()) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WSMEM);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done.set();
  }
  {
    int cz = 8;
    while (B % cz) cz >>= 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid_b;
    cfg.blockDim = dim3(DW_THREADS);
    cfg.dynamicSmemBytes = WSMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = cz;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, dwconv_bwd_weight_kernel, reinterpret_cast<const bf16*>(dw),
                                       reinterpret_cast<const bf16*>(u), T, d, dweight, dbias, tiles_per_cta);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
  }
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
