// conv1 (1 -> d channels, 3x3, stride 2) on the tensor cores, forward and weight / bias gradient.
//
// The CUDA-core kernels (conv_gemm.cu) are issue bound: 9 FMA + SiLU per output (forward) and 18 FMA + SiLU' per output
// (backward: recompute of the pre-activation + the dW accumulation) for 435 M outputs per step.  Here the FMAs move to
// tcgen05 and the CUDA cores keep only the element-wise part:
//
//   P   (128 pixels x 64) bf16  im2col patch tile of one 128-pixel tile, built in shared memory by the CTA
//                               cols 0..8 hi(x), 9 = 1, 16..24 lo(x) = bf16(x - hi), 32..40 hi(x) again, 41 = 1
//   Wt  (d x 64) bf16           cols 0..8 hi(w), 9 = hi(b), 16..24 hi(w), 32..40 lo(w), 41 = lo(b)
//   Z   = P Wt^T  (K = 48)      = (x_hi + x_lo) w_hi + x_hi w_lo + b : the pre-activation to ~2^-16 (fp32 accumulate)
//   fwd: y1 = silu(Z) -> bf16, stored by the thread that owns the pixel row
//   bwd: dZ = dy1 * silu'(Z) -> bf16 tile in shared memory;  D (d x 32) += dZ^T P[:, 0:32]  (MN-major operands, the
//        same P tile): cols 0..8 + 16..24 = dW1, col 9 = db1 (the ones column); D stays in TMEM for the whole kernel.
//
// One persistent CTA per SM walks 128-pixel tiles; 512 threads: warp w owns TMEM lanes (= pixel rows) 32 (w % 4) .. +31
// and the channel quarter w / 4.  Replaces (reference): model/conformer.py:150-151 (Conv2d(1, d, 3, 2, 1) + SiLU) and
// its autograd backward for the weight and bias (the input needs no gradient).
#include "common.cuh"

namespace {

constexpr int TC_THREADS = 512;
constexpr int TPIX = 128;       // pixels per tile
constexpr int CD = 256;         // channels handled by the tensor-core path
constexpr int P_BYTES = TPIX * 128;   // 16 KB, 128-byte swizzled rows
constexpr int W_BYTES = CD * 128;     // 32 KB
constexpr int DZ_BYTES = TPIX * CD * 2;  // 64 KB: four [128 x 64] bf16 chunks

struct Conv1TcParams {
  const float* x;    // (B, T, F)
  int B, T, F, T1, F1;
  long long npix;    // B * T1 * F1
  int ntiles;
  const float* w1;   // (d, 9)
  const float* b1;   // (d)
  bf16* y1;          // fwd out (npix, d)
  const bf16* dy1;   // bwd in (npix, d)
  float* dw1;        // (d, 9)  +=
  float* db1;        // (d)     +=
  int ch0;           // first of the 256 output channels this launch computes (d > 256: one launch per 256-channel chunk)
};

// byte offset of bf16 element (row r, col c) in a [rows x 64] tile with 128-byte rows and the TMA/UMMA 128 B swizzle
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)(r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1)));
}
__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16(v);
  lo = __float2bfloat16(v - __bfloat162float(hi));
}

// Wt tile: row = channel.  Written once per CTA.
__device__ __forceinline__ void build_weight_tile(uint8_t* sW, const float* __restrict__ w1, const float* __restrict__ b1) {
  for (int i = threadIdx.x; i < CD * 8; i += TC_THREADS)  // zero (16-byte pieces)
    reinterpret_cast<uint4*>(sW)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < CD * 10; i += TC_THREADS) {
    const int c = i / 10, k = i - c * 10;
    const float v = k < 9 ? w1[c * 9 + k] : b1[c];
    bf16 hi, lo;
    split_bf16(v, hi, lo);
    *reinterpret_cast<bf16*>(sW + sw128_off(c, k)) = hi;
    *reinterpret_cast<bf16*>(sW + sw128_off(c, 32 + k)) = lo;
    if (k < 9) *reinterpret_cast<bf16*>(sW + sw128_off(c, 16 + k)) = hi;
  }
}

// P tile of tile index `tile`: thread (pixel r = tid & 127, part = tid >> 7) writes taps 3 part .. 3 part + 2 (part < 3)
// or the ones columns (part == 3).  Columns that are never written stay zero (the buffers are zeroed once).
// Split in two so that the global loads of a tile are in flight while the previous tile is processed.
struct PatchRegs {
  float v[3];
  bool valid;
};
__device__ __forceinline__ PatchRegs fetch_patch(const Conv1TcParams& p, int tile) {
  PatchRegs pr;
  pr.v[0] = pr.v[1] = pr.v[2] = 0.f;
  const int r = threadIdx.x & (TPIX - 1), part = threadIdx.x >> 7;
  const long long pix = (long long)tile * TPIX + r;
  pr.valid = tile < p.ntiles && pix < p.npix;
  if (part == 3 || !pr.valid) return pr;
  const long long bh = pix / p.F1;
  const int w = (int)(pix - bh * p.F1);
  const int b = (int)(bh / p.T1);
  const int h = (int)(bh - (long long)b * p.T1);
  const int tt = 2 * h - 1 + part;
  if (tt < 0 || tt >= p.T) return pr;
  const float* xr = p.x + ((long long)b * p.T + tt) * p.F;
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
    const int ff = 2 * w - 1 + kw;
    if (ff >= 0 && ff < p.F) pr.v[kw] = xr[ff];
  }
  return pr;
}
__device__ __forceinline__ void store_patch(uint8_t* sP, const PatchRegs& pr) {
  const int r = threadIdx.x & (TPIX - 1), part = threadIdx.x >> 7;
  if (part == 3) {
    const bf16 one = __float2bfloat16(pr.valid ? 1.f : 0.f);
    *reinterpret_cast<bf16*>(sP + sw128_off(r, 9)) = one;
    *reinterpret_cast<bf16*>(sP + sw128_off(r, 41)) = one;
    return;
  }
#pragma unroll
  for (int kw = 0; kw < 3; ++kw) {
    bf16 hi, lo;
    split_bf16(pr.v[kw], hi, lo);
    const int k = part * 3 + kw;
    *reinterpret_cast<bf16*>(sP + sw128_off(r, k)) = hi;
    *reinterpret_cast<bf16*>(sP + sw128_off(r, 16 + k)) = lo;
    *reinterpret_cast<bf16*>(sP + sw128_off(r, 32 + k)) = hi;
  }
}
__device__ __forceinline__ void build_patch_tile(uint8_t* sP, const Conv1TcParams& p, int tile) {
  store_patch(sP, fetch_patch(p, tile));
}

// Z (128 x 256) = P Wt^T over K = 48 (three 16-wide steps), issued by one thread
__device__ __forceinline__ void issue_gemm1(uint32_t tZ, const uint8_t* sP, const uint8_t* sW) {
  constexpr uint32_t idesc = umma_idesc_bf16(128, CD, 0, 0);
  const uint32_t pa = smem_u32(sP), wa = smem_u32(sW);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    umma_bf16(tZ, umma_desc_sw128(pa + k * 32, 16, 1024), umma_desc_sw128(wa + k * 32, 16, 1024), idesc, k > 0);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) conv1_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmY, const Conv1TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                       // 32 KB
  uint8_t* sP = smem + W_BYTES;             // 2 x 16 KB
  uint8_t* sY = smem + W_BYTES + 2 * P_BYTES;  // 2 x 64 KB output staging: four [128 x 64] bf16 chunks each (TMA store)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sY + 2 * DZ_BYTES);  // z0, z1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, cq = warp >> 2;
  const int rloc = quarter * 32 + lane;
  if (tid == 0) {
    tma_prefetch_desc(&tmY);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 2 * P_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
  build_weight_tile(sW, p.w1 + p.ch0 * 9, p.b1 + p.ch0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;

  const int n = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA
  if (n > 0) {
    build_patch_tile(sP, p, blockIdx.x);
    if (n > 1) build_patch_tile(sP + P_BYTES, p, blockIdx.x + gridDim.x);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_gemm1(tmem, sP, sW);
      umma_commit(&bars[0]);
    }
    for (int i = 0; i < n; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int zb = i & 1;
      // input patch of tile i+2: its loads are in flight while tile i is processed (stored after the column loop)
      const PatchRegs nxt = fetch_patch(p, (i + 2 < n) ? tile + 2 * (int)gridDim.x : p.ntiles);
      if (tid == 0) bulk_wait_read<1>();  // the stores of tile i-2 have read their staging tile (two tiles alternate)
      tc_fence_before();
      __syncthreads();  // everyone has drained Z buffer (i+1)&1 (tile i-1); patch i+1 visible; staging free
      if (tid == 0 && i + 1 < n) {
        tc_fence_after();
        issue_gemm1(tmem + ((i + 1) & 1) * CD, sP + ((i + 1) & 1) * P_BYTES, sW);
        umma_commit(&bars[(i + 1) & 1]);
      }
      mbar_wait(&bars[zb], (uint32_t)(i >> 1) & 1u);
      __syncwarp();
      tc_fence_after();
      uint8_t* sYt = sY + zb * DZ_BYTES;
      uint8_t* row = sYt + cq * (TPIX * 128) + rloc * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t u[16];
        tmem_ld16(tmem + zb * CD + lane_addr + cq * 64 + c * 16, u);
        tmem_ld_wait();
        const float2 half2 = make_float2(0.5f, 0.5f);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 z = make_float2(__uint_as_float(u[2 * j]), __uint_as_float(u[2 * j + 1]));
          const float2 hh = __fmul2_rn(z, half2);
          const float2 t = make_float2(tanh_approx(hh.x), tanh_approx(hh.y));
          const float2 y = __ffma2_rn(hh, t, hh);
          pk[j] = pack_bf16x2(y.x, y.y);
        }
        *reinterpret_cast<uint4*>(row + (((2 * c) ^ (rloc & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(row + (((2 * c + 1) ^ (rloc & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      if (i + 2 < n) store_patch(sP + zb * P_BYTES, nxt);  // buffer i&1: its product (tile i) is complete
      fence_proxy_async_smem();
      __syncthreads();  // output tile staged, patch i+2 written
      if (tid == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) tma_store_2d(&tmY, sYt + q * (TPIX * 128), p.ch0 + q * 64, tile * TPIX);  // rows >= npix clipped
        bulk_commit();
      }
    }
    if (tid == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// backward (weight and bias gradient)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) conv1_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmG, const Conv1TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                   // 32 KB
  uint8_t* sP = smem + W_BYTES;                         // 2 x 16 KB
  uint8_t* sDZ = smem + W_BYTES + 2 * P_BYTES;          // 64 KB
  uint8_t* sG = sDZ + DZ_BYTES;                         // 64 KB: dy1 tile (TMA, four [128 x 64] chunks)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + DZ_BYTES);  // z, d, g, (unused)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, cq = warp >> 2;
  const int rloc = quarter * 32 + lane;
  if (tid == 0) {
    tma_prefetch_desc(&tmG);
    for (int q = 0; q < 4; ++q) mbar_init(&bars[q], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 2 * P_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
  build_weight_tile(sW, p.w1 + p.ch0 * 9, p.b1 + p.ch0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  auto load_dy = [&](int tile) {  // thread 0: the tile's 128 x 256 gradient block (rows >= npix zero-filled)
    mbar_expect_tx(&bars[2], DZ_BYTES);
#pragma unroll
    for (int q = 0; q < 4; ++q) tma_load_2d(sG + q * (TPIX * 128), &tmG, &bars[2], p.ch0 + q * 64, tile * TPIX);
  };
  const uint32_t tZ = tmem, tD = tmem + CD;  // D: two accumulators (channel halves) of 32 columns
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
  constexpr uint32_t idesc2 = umma_idesc_bf16(128, 32, 1, 1);

  const int n = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (n > 0) {
    build_patch_tile(sP, p, blockIdx.x);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      load_dy(blockIdx.x);
      tc_fence_after();
      issue_gemm1(tZ, sP, sW);
      umma_commit(&bars[0]);
    }
    const float2 half2 = make_float2(0.5f, 0.5f), one2 = make_float2(1.f, 1.f), mone2 = make_float2(-1.f, -1.f);
    for (int i = 0; i < n; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const uint32_t ph = (uint32_t)i & 1u;
      const long long pix = (long long)tile * TPIX + rloc;
      const bool pvalid = pix < p.npix;
      // this thread's 64 dy values of the tile: shared memory -> registers, then the buffer is refilled for tile i+1
      (void)pvalid;
      uint4 g[8];
      {
        mbar_wait(&bars[2], ph);
        const uint8_t* grow = sG + cq * (TPIX * 128) + rloc * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] = *reinterpret_cast<const uint4*>(grow + ((q ^ (rloc & 7)) << 4));
        __syncthreads();
        if (tid == 0 && i + 1 < n) {
          fence_proxy_async_smem();
          load_dy(tile + (int)gridDim.x);
        }
      }
      // input patch of tile i+1: loads in flight while this tile is processed (stored after the column loop)
      const PatchRegs nxt = fetch_patch(p, (i + 1 < n) ? tile + (int)gridDim.x : p.ntiles);
      // the dW / db product of tile i-1 has finished reading the dZ tile and patch buffer (i+1)&1
      if (i > 0) mbar_wait(&bars[1], ph ^ 1u);
      mbar_wait(&bars[0], ph);
      __syncwarp();
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t u[16];
        tmem_ld16(tZ + lane_addr + cq * 64 + c * 16, u);
        tmem_ld_wait();

        const uint32_t gw[8] = {g[2 * c].x, g[2 * c].y, g[2 * c].z, g[2 * c].w, g[2 * c + 1].x, g[2 * c + 1].y, g[2 * c + 1].z, g[2 * c + 1].w};
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // silu'(z) = (1 + t)(1 + h (1 - t)) / 2,  h = z / 2, t = tanh(h)
          const float2 z = make_float2(__uint_as_float(u[2 * j]), __uint_as_float(u[2 * j + 1]));
          const float2 dy = unpack_bf16x2(gw[j]);
          const float2 hh = __fmul2_rn(z, half2);
          const float2 t = make_float2(tanh_approx(hh.x), tanh_approx(hh.y));
          const float2 a1 = __fadd2_rn(t, one2);
          const float2 b1m = __ffma2_rn(t, mone2, one2);
          const float2 cc = __ffma2_rn(hh, b1m, one2);
          const float2 dz = __fmul2_rn(__fmul2_rn(dy, half2), __fmul2_rn(a1, cc));
          pk[j] = pack_bf16x2(dz.x, dz.y);
        }
        // channel block cq (64 channels) is chunk cq of the dZ tile: [128 pixel rows x 128 B], columns c*16 .. c*16+15
        uint8_t* row = sDZ + cq * (TPIX * 128) + rloc * 128;
        *reinterpret_cast<uint4*>(row + (((2 * c) ^ (rloc & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(row + (((2 * c + 1) ^ (rloc & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      if (i + 1 < n) store_patch(sP + ((i + 1) & 1) * P_BYTES, nxt);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncthreads();  // Z drained, dZ tile and patch i+1 complete
      if (tid == 0) {
        tc_fence_after();
        const uint32_t za = smem_u32(sDZ), pa = smem_u32(sP + (i & 1) * P_BYTES);
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)  // D[mh] (128 channels x 32) += dZ^T P   (reduction over the 128 pixels)
#pragma unroll
          for (int k = 0; k < TPIX / 16; ++k)
            umma_bf16(tD + mh * 32, umma_desc_sw128(za + mh * 2 * (TPIX * 128) + k * 2048, TPIX * 128, 1024),
                      umma_desc_sw128(pa + k * 2048, 8192, 1024), idesc2, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&bars[1]);
        if (i + 1 < n) {
          issue_gemm1(tZ, sP + ((i + 1) & 1) * P_BYTES, sW);
          umma_commit(&bars[0]);
        }
      }
    }
    mbar_wait(&bars[1], (uint32_t)(n - 1) & 1u);
    __syncwarp();
    tc_fence_after();
    // D: lane = channel (two halves), columns 0..8 hi-patch part, 9 = bias, 16..24 lo-patch part
    if (cq < 2) {
      uint32_t u[32];
      tmem_ld32(tD + cq * 32 + lane_addr, u);
      tmem_ld_wait();
      const int ch = cq * 128 + rloc;
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(p.dw1 + (p.ch0 + ch) * 9 + k, __uint_as_float(u[k]) + __uint_as_float(u[16 + k]));
      atomicAdd(p.db1 + p.ch0 + ch, __uint_as_float(u[9]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

typedef CUresult (*PFN_encodeTiledTc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// (npix, d) bf16 row-major, box 64 columns x 128 rows, 128 B swizzle
int make_out_map(CUtensorMap* m, void* base, long long npix, int d) {
  static PFN_encodeTiledTc enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return TASR_ERR_CUDA;
    enc = reinterpret_cast<PFN_encodeTiledTc>(fn);
  }
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)npix};
  cuuint64_t strides[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {64, TPIX};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TASR_OK : TASR_ERR_CUDA;
}

int g_tc_sms = 0;
int tc_num_sms() {
  if (g_tc_sms <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_tc_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_tc_sms <= 0) g_tc_sms = 148;
  }
  return g_tc_sms;
}

bool fill(Conv1TcParams* p, const float* x, int B, int T, int F, int d, const float* w1, const float* b1) {
  if (d <= 0 || d % CD || B <= 0 || T <= 0 || F <= 0) return false;  // d = 256 (default), 512 (Conformer-M), ...
  p->x = x; p->B = B; p->T = T; p->F = F;
  p->T1 = (T - 1) / 2 + 1;
  p->F1 = (F - 1) / 2 + 1;
  p->npix = (long long)B * p->T1 * p->F1;
  const long long nt = (p->npix + TPIX - 1) / TPIX;
  if (p->npix > 0x7fffff00LL || nt > 0x3fffffffLL) return false;  // pixel rows are 32-bit TMA coordinates
  p->ntiles = (int)nt;
  p->w1 = w1; p->b1 = b1;
  p->y1 = nullptr; p->dy1 = nullptr; p->dw1 = nullptr; p->db1 = nullptr;
  p->ch0 = 0;
  return true;
}

}  // namespace

// returns TASR_OK when the tensor-core path ran, TASR_ERR_SHAPE when the shape is not covered (caller falls back to the
// CUDA-core kernel)
int tasr_conv1_tc_fwd(const float* x, int B, int T, int F, int d, const float* w1, const float* b1, void* y1,
                      cudaStream_t st) {
  Conv1TcParams p;
  if (!fill(&p, x, B, T, F, d, w1, b1)) return TASR_ERR_SHAPE;
  p.y1 = reinterpret_cast<bf16*>(y1);
  if (reinterpret_cast<uintptr_t>(y1) & 15) return TASR_ERR_SHAPE;
  CUtensorMap tmY;
  if (make_out_map(&tmY, y1, p.npix, d) != TASR_OK) return TASR_ERR_CUDA;
  constexpr int SMEM = W_BYTES + 2 * P_BYTES + 2 * DZ_BYTES + 64 + 1024;
  static TasrPerDevice attr_done;
  if (!attr_done.get()) {
    cudaError_t e = cudaFuncSetAttribute(conv1_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done.set();
  }
  const int grid = p.ntiles < tc_num_sms() ? p.ntiles : tc_num_sms();
  for (p.ch0 = 0; p.ch0 < d; p.ch0 += CD) {  // output channels are independent: one launch per 256-channel chunk
    conv1_tc_fwd_kernel<<<grid, TC_THREADS, SMEM, st>>>(tmY, p);
    TASR_CHECK_LAUNCH();
  }
  return TASR_OK;
}

int tasr_conv1_tc_bwd(const void* dy1, const float* x, int B, int T, int F, int d, const float* w1, const float* b1,
                      float* dw1, float* db1, cudaStream_t st) {
  Conv1TcParams p;
  if (!fill(&p, x, B, T, F, d, w1, b1)) return TASR_ERR_SHAPE;
  p.dy1 = reinterpret_cast<const bf16*>(dy1);
  p.dw1 = dw1;
  p.db1 = db1;
  if (reinterpret_cast<uintptr_t>(dy1) & 15) return TASR_ERR_SHAPE;
  CUtensorMap tmG;
  if (make_out_map(&tmG, const_cast<void*>(dy1), p.npix, d) != TASR_OK) return TASR_ERR_CUDA;
  constexpr int SMEM = W_BYTES + 2 * P_BYTES + 2 * DZ_BYTES + 64 + 1024;
  static_assert(SMEM <= 232448, "shared memory budget");
  static TasrPerDevice attr_done;
  if (!attr_done.get()) {
    cudaError_t e = cudaFuncSetAttribute(conv1_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done.set();
  }
  const int grid = p.ntiles < tc_num_sms() ? p.ntiles : tc_num_sms();
  for (p.ch0 = 0; p.ch0 < d; p.ch0 += CD) {
    conv1_tc_bwd_kernel<<<grid, TC_THREADS, SMEM, st>>>(tmG, p);
    TASR_CHECK_LAUNCH();
  }
  return TASR_OK;
}
