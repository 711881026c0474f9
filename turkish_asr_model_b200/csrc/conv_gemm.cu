// Conv2d subsampler as IMPLICIT GEMM on tcgen05: no im2col buffer, operands gathered by 4-D TMA.
// Replaces (reference): model/conformer.py:150-155,177-183  Conv2d(1,d,3,s2,p1)+SiLU, Conv2d(d,d,3,s2,p1)+SiLU
//   (cuDNN implicit GEMM forward, dgrad, wgrad).
//
// Layouts: y1 = silu(conv1(x)) is NHWC bf16 (B, T1, F1, d); conv2 output z2 / y2 is (B, T2, F2, d) which IS
// the (B*T2, F2*d) matrix input_proj consumes (weight packed to (n, f*d + c)).
//
// Stride-2 gather without TMA element strides: a 3x3/stride-2/pad-1 tap (kh, kw) reads rows h = 2t'-1+kh, i.e.
// only rows of ONE parity (kh = 1 -> even rows, index t'; kh = 0 -> odd rows, index t'-1; kh = 2 -> odd rows,
// index t').  So y1 is described by four "parity class" tensor maps (doubled strides, base offset (ph, pw)),
// and every tap is a dense 4-D box {64 ch, 4 f, 32 t, 1 b} of one class map at an offset of 0 or -1; rows that
// fall into the zero padding are out of bounds of the map and arrive as zeros.  The same class maps describe
// the gradient dy1 for the backward data pass, which runs one GEMM per parity class (1, 2, 2 and 4 taps).
//
//   FWD   : z2[pix, co]  = sum_{tap, ci} y1[pix@tap, ci] * W2p[co, tap*d + ci]      (SiLU epilogue -> z2, y2)
//   DGRAD : dy1[class pix, ci] = sum_{tap in class, co} dz2[pix@tap, co] * W2p[co, tap*d + ci]
//   WGRAD : dW2[co, ci, tap] += sum_pix dz2[pix, co] * y1[pix@tap, ci]               (split over pixel blocks)
// The main loop / barrier / TMEM structure is the one of gemm.cu (persistent, warp-specialised, 2 accumulators).
#include "gemm_common.cuh"
#include <mutex>

namespace {

enum { CONV_FWD = 0, CONV_DGRAD = 1, CONV_WGRAD = 2 };

struct ConvDev {
  int B, d;
  int T2, F2;        // conv2 output map
  int nt_blk;        // time blocks per utterance for the M (FWD/DGRAD) or reduction (WGRAD) tiling
  int nf_blk;        // frequency blocks (of 4)
  int ph, pw;        // DGRAD: parity class of the output rows / columns
  int ntaps;         // DGRAD: taps in this class
  int taps[4];       // DGRAD: tap ids (kh*3+kw)
  int kc;            // d / 64
  int tiles_m, tiles_n, splits, kb_per_split, total_kb;
  const float* bias;
  float* dw;         // WGRAD: fp32 gradient (co, ci, 3, 3), accumulated
};

constexpr int CV_STAGES = 3;
constexpr int CV_RINGG = 2;

template <int MODE, int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap cm0, const __grid_constant__ CUtensorMap cm1,
                 const __grid_constant__ CUtensorMap cm2, const __grid_constant__ CUtensorMap cm3,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmZ,
                 const __grid_constant__ CUtensorMap tmY, const ConvDev p) {
  // cm0..3 : parity-class maps of y1 (FWD, WGRAD) or dy1 (DGRAD, output)      index = ph*2 + pw
  // tmW    : W2p (d, 9d) 2-D           tmZ : z2 (FWD out2) | dz2 (DGRAD/WGRAD in), 4-D     tmY : y2 (FWD out), 4-D
  constexpr int STAGES = CV_STAGES;
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr bool A_MN = (MODE == CONV_WGRAD);
  constexpr bool B_MN = (MODE != CONV_FWD);
  constexpr int TROWS = (MODE == CONV_WGRAD) ? 16 : 32;  // time rows per TMA box (x4 frequency bins)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sC = sB + STAGES * B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sC + 2 * CV_RINGG * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n * p.splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&cm0); tma_prefetch_desc(&cm1); tma_prefetch_desc(&cm2); tma_prefetch_desc(&cm3);
    tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmZ); tma_prefetch_desc(&tmY);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dependency_wait();  // prologue above overlaps the previous kernel's tail (PDL)

  auto class_map = [&](int idx) -> const CUtensorMap* {
    return idx == 0 ? &cm0 : (idx == 1 ? &cm1 : (idx == 2 ? &cm2 : &cm3));
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.tiles_n;
        const int m_tile = (tile / p.tiles_n) % p.tiles_m;
        const int split = tile / (p.tiles_n * p.tiles_m);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
          uint8_t* a_dst = sA + s * A_BYTES;
          uint8_t* b_dst = sB + s * B_BYTES;
          if (MODE == CONV_FWD) {
            // M tile -> (b, tblk, fblk);  k-block -> (tap, ci chunk)
            const int fblk = m_tile % p.nf_blk;
            const int tblk = (m_tile / p.nf_blk) % p.nt_blk;
            const int b = m_tile / (p.nf_blk * p.nt_blk);
            const int tap = kb / p.kc, c0 = (kb - tap * p.kc) * 64;
            const int kh = tap / 3, kw = tap - kh * 3;
            const CUtensorMap* cm = class_map(((kh + 1) & 1) * 2 + ((kw + 1) & 1));
            tma_load_4d(a_dst, cm, &full_bar[s], c0, 4 * fblk - (kw == 0), 32 * tblk - (kh == 0), b);
            const int n0 = n_tile * BN;
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmW, &full_bar[s], tap * p.d + c0, n0 + 64 * j);
          } else if (MODE == CONV_DGRAD) {
            // M tile -> (b, iblk, jblk) of the parity class;  k-block -> (tap of the class, co chunk)
            const int jblk = m_tile % p.nf_blk;
            const int iblk = (m_tile / p.nf_blk) % p.nt_blk;
            const int b = m_tile / (p.nf_blk * p.nt_blk);
            const int ti = kb / p.kc, co0 = (kb - ti * p.kc) * 64;
            const int tap = p.taps[ti];
            const int kh = tap / 3, kw = tap - kh * 3;
            // t' = i + (ph + 1 - kh) / 2  ->  +1 only for (ph = 1, kh = 0)
            tma_load_4d(a_dst, &tmZ, &full_bar[s], co0, 4 * jblk + (p.pw == 1 && kw == 0), 32 * iblk + (p.ph == 1 && kh == 0), b);
            const int n0 = n_tile * BN;
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmW, &full_bar[s], tap * p.d + n0 + 64 * j, co0);
          } else {
            // WGRAD: M tile -> co block, N tile -> (tap, ci block);  k-block -> pixel block (b, tblk16, fblk)
            const int fblk = kb % p.nf_blk;
            const int tblk = (kb / p.nf_blk) % p.nt_blk;
            const int b = kb / (p.nf_blk * p.nt_blk);
            const int nper = p.d / BN;
            const int tap = n_tile / nper, ci0 = (n_tile - tap * nper) * BN;
            const int kh = tap / 3, kw = tap - kh * 3;
            const int m0 = m_tile * BM;
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_4d(a_dst + j * 8192, &tmZ, &full_bar[s], m0 + 64 * j, 4 * fblk, 16 * tblk, b);
            const CUtensorMap* cm = class_map(((kh + 1) & 1) * 2 + ((kw + 1) & 1));
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(b_dst + j * 8192, cm, &full_bar[s], ci0 + 64 * j, 4 * fblk - (kw == 0), 16 * tblk - (kh == 0), b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int split = tile / (p.tiles_n * p.tiles_m);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
        const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], aph ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * A_BYTES);
          const uint32_t b_base = smem_u32(sB + s * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? umma_desc_sw128(a_base + k * 2048, 8192, 1024)
                                        : umma_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_desc_sw128(b_base + k * 2048, 8192, 1024)
                                        : umma_desc_sw128(b_base + k * 32, 16, 1024);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue: 2 groups x 4 warps =====================
    const int ew = warp - 2;
    const int grp = ew >> 2;
    const int q = warp & 3;
    const int rloc = q * 32 + lane;
    const bool leader = (q == 2) && lane == 0;
    const int bar_id = 1 + grp;
    uint8_t* ring_base = sC + grp * CV_RINGG * STAGE_BYTES;
    uint32_t tcount = 0, ring = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int n_tile = tile % p.tiles_n;
      const int m_tile = (tile / p.tiles_n) % p.tiles_m;
      const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
      mbar_wait(&tfull_bar[acc], aph);
      __syncwarp();
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      if (MODE == CONV_WGRAD) {
        // rows = co, columns = (tap, ci): scatter-add into the reference (co, ci, kh, kw) layout
        const int nper = p.d / BN;
        const int tap = n_tile / nper, ci0 = (n_tile - tap * nper) * BN;
        const int co = m_tile * BM + rloc;
#pragma unroll 1
        for (int g = grp; g < BN / 64; g += 2) {
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t u[32];
            tmem_ld32(tbase + g * 64 + h * 32, u);
            tmem_ld_wait();
            if (co < p.d) {
              float* o = p.dw + ((long long)co * p.d + ci0 + g * 64 + h * 32) * 9 + tap;
#pragma unroll
              for (int i = 0; i < 32; ++i) atomicAdd(o + i * 9, __uint_as_float(u[i]));
            }
          }
        }
      } else {
        const int fblk = m_tile % p.nf_blk;
        const int tblk = (m_tile / p.nf_blk) % p.nt_blk;
        const int b = m_tile / (p.nf_blk * p.nt_blk);
        const int n0 = n_tile * BN;
        constexpr int NBUF = (MODE == CONV_FWD) ? 2 : 1;
#pragma unroll 1
        for (int g = grp; g < BN / 64; g += 2) {
          if (leader) bulk_wait_read<CV_RINGG - NBUF>();
          named_bar_sync(bar_id, EPI_GROUP_THREADS);
          uint8_t* buf0 = ring_base + (ring % CV_RINGG) * STAGE_BYTES;
          uint8_t* buf1 = ring_base + ((ring + 1) % CV_RINGG) * STAGE_BYTES;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const int col0 = n0 + g * 64 + h * 32;
            uint32_t u[32];
            tmem_ld32(tbase + g * 64 + h * 32, u);
            tmem_ld_wait();
            float* v = reinterpret_cast<float*>(u);
            if (MODE == CONV_FWD) {
              float y[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                v[i] = bf16_round(v[i] + p.bias[col0 + i]);
                y[i] = siluf_(v[i]);
              }
              stage_bf16_half(buf0, rloc, h, y);
              stage_bf16_half(buf1, rloc, h, v);
            } else {
              stage_bf16_half(buf0, rloc, h, v);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, EPI_GROUP_THREADS);
          if (leader) {
            const int c = n0 + g * 64;
            if (MODE == CONV_FWD) {
              tma_store_4d(&tmY, buf0, c, 4 * fblk, 32 * tblk, b); bulk_commit();
              tma_store_4d(&tmZ, buf1, c, 4 * fblk, 32 * tblk, b); bulk_commit();
            } else {
              tma_store_4d(class_map(p.ph * 2 + p.pw), buf0, c, 4 * fblk, 32 * tblk, b); bulk_commit();
            }
          }
          ring += NBUF;
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
    if (leader) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// conv1 (1 -> d channels, K = 9): direct, bandwidth bound.  One 8-channel slice per thread.
// ------------------------------------------------------------------------------------------------
constexpr int NT = 256;

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

// Each thread: 8 channels x PX = 4 neighbouring output columns of one row (the 3 x 9 input patch and the 72
// weights are loaded once for 32 outputs).
constexpr int PX = 4;

__device__ __forceinline__ void load_patch(const float* __restrict__ xb, int T, int F, int h, int w0, float (*xp)[2 * PX + 1]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int tt = 2 * h - 1 + i;
    const bool rok = tt >= 0 && tt < T;
#pragma unroll
    for (int j = 0; j < 2 * PX + 1; ++j) {
      const int ff = 2 * w0 - 1 + j;
      xp[i][j] = (rok && ff >= 0 && ff < F) ? xb[(long long)tt * F + ff] : 0.f;
    }
  }
}
__device__ __forceinline__ void conv1_rows(const float (*xp)[2 * PX + 1], const float* __restrict__ w1s,
                                           const float* __restrict__ b1s, int d, int c0, float (*z)[8]) {
#pragma unroll
  for (int px = 0; px < PX; ++px)
#pragma unroll
    for (int q = 0; q < 8; ++q) z[px][q] = b1s[c0 + q];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 wa = *reinterpret_cast<const float4*>(w1s + k * d + c0);
    const float4 wb = *reinterpret_cast<const float4*>(w1s + k * d + c0 + 4);
    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      const float xv = xp[k / 3][2 * px + k % 3];
#pragma unroll
      for (int q = 0; q < 8; ++q) z[px][q] = fmaf(xv, wv[q], z[px][q]);
    }
  }
}

// x (B, T, F) fp32 -> y1 (B, T1, F1, d) bf16 = silu(conv1(x))
__global__ void __launch_bounds__(NT) conv1_fwd_kernel(const float* __restrict__ x, int B, int T, int F, int d,
                                                       const float* __restrict__ w1, const float* __restrict__ b1, int T1,
                                                       int F1, bf16* __restrict__ y1) {
  extern __shared__ float sh_w[];
  float* w1s = sh_w;
  float* b1s = sh_w + 9 * d;
  for (int i = threadIdx.x; i < 9 * d; i += NT) {
    const int k = i / d, c = i - k * d;
    w1s[i] = w1[c * 9 + k];
  }
  for (int i = threadIdx.x; i < d; i += NT) b1s[i] = b1[i];
  __syncthreads();
  const int lanes_per_item = d >> 3;
  const int items_per_iter = NT / lanes_per_item;
  const int c0 = (threadIdx.x % lanes_per_item) << 3;
  const int sub = threadIdx.x / lanes_per_item;
  const int wgroups = (F1 + PX - 1) / PX;
  const long long total = (long long)B * T1 * wgroups;
  for (long long item = (long long)blockIdx.x * items_per_iter + sub; item < total; item += (long long)gridDim.x * items_per_iter) {
    const int wg = (int)(item % wgroups);
    const long long bh = item / wgroups;
    const int h = (int)(bh % T1), b = (int)(bh / T1);
    const int w0 = wg * PX;
    float xp[3][2 * PX + 1], z[PX][8];
    load_patch(x + (long long)b * T * F, T, F, h, w0, xp);
    conv1_rows(xp, w1s, b1s, d, c0, z);
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      if (w0 + px < F1) {
        uint4 o;
        o.x = pack_bf16x2(siluf_(z[px][0]), siluf_(z[px][1]));
        o.y = pack_bf16x2(siluf_(z[px][2]), siluf_(z[px][3]));
        o.z = pack_bf16x2(siluf_(z[px][4]), siluf_(z[px][5]));
        o.w = pack_bf16x2(siluf_(z[px][6]), siluf_(z[px][7]));
        *reinterpret_cast<uint4*>(y1 + (((long long)b * T1 + h) * F1 + w0 + px) * d + c0) = o;
      }
    }
  }
}

// dy1 (B, T1, F1, d) bf16 -> dW1 (d,1,3,3), db1 (d)  (+=);  z1 is recomputed from x
__global__ void __launch_bounds__(NT) conv1_bwd_kernel(const bf16* __restrict__ dy1, const float* __restrict__ x, int B,
                                                       int T, int F, int d, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, int T1, int F1, float* __restrict__ dw1,
                                                       float* __restrict__ db1) {
  extern __shared__ float sh_w[];
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
  const int wgroups = (F1 + PX - 1) / PX;
  const long long total = (long long)B * T1 * wgroups;
  for (long long item = (long long)blockIdx.x * items_per_iter + sub; item < total; item += (long long)gridDim.x * items_per_iter) {
    const int wg = (int)(item % wgroups);
    const long long bh = item / wgroups;
    const int h = (int)(bh % T1), b = (int)(bh / T1);
    const int w0 = wg * PX;
    uint4 u[PX];
#pragma unroll
    for (int px = 0; px < PX; ++px)
      u[px] = (w0 + px < F1) ? *reinterpret_cast<const uint4*>(dy1 + (((long long)b * T1 + h) * F1 + w0 + px) * d + c0)
                             : make_uint4(0, 0, 0, 0);
    float xp[3][2 * PX + 1], z[PX][8];
    load_patch(x + (long long)b * T * F, T, F, h, w0, xp);
    conv1_rows(xp, w1s, b1s, d, c0, z);
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      float g[8];
      float2 t;
      t = unpack_bf16x2(u[px].x); g[0] = t.x; g[1] = t.y;
      t = unpack_bf16x2(u[px].y); g[2] = t.x; g[3] = t.y;
      t = unpack_bf16x2(u[px].z); g[4] = t.x; g[5] = t.y;
      t = unpack_bf16x2(u[px].w); g[6] = t.x; g[7] = t.y;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float dz = g[q] * silu_gradf_(z[px][q]);
        gb[q] += dz;
#pragma unroll
        for (int k = 0; k < 9; ++k) gw[k][q] = fmaf(dz, xp[k / 3][2 * px + k % 3], gw[k][q]);
      }
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_enc = nullptr;
std::once_flag g_enc_once;
int g_sms = 148;

void init_enc() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    g_enc = reinterpret_cast<PFN_encodeTiled>(fn);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
  if (g_sms <= 0) g_sms = 148;
}

// 4-D bf16 map {c, f, t, b} with explicit byte strides for f, t, b; box {64, 4, trows, 1}; 128 B swizzle
int make_map4(CUtensorMap* m, const void* base, uint64_t c, uint64_t f, uint64_t t, uint64_t b, uint64_t sf, uint64_t st_,
              uint64_t sb, uint32_t trows) {
  std::call_once(g_enc_once, init_enc);
  if (!g_enc) return TASR_ERR_CUDA;
  cuuint64_t dims[4] = {c, f, t, b};
  cuuint64_t strides[3] = {sf, st_, sb};
  cuuint32_t box[4] = {64, 4, trows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TASR_OK : TASR_ERR_CUDA;
}
int make_map2(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld) {
  std::call_once(g_enc_once, init_enc);
  if (!g_enc) return TASR_ERR_CUDA;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TASR_OK : TASR_ERR_CUDA;
}

struct Geom {
  int T1, F1, T2, F2;
};
Geom geom(int T, int F) {
  Geom g;
  g.T1 = (T - 1) / 2 + 1; g.F1 = (F - 1) / 2 + 1;
  g.T2 = (g.T1 - 1) / 2 + 1; g.F2 = (g.F1 - 1) / 2 + 1;
  return g;
}

// the four parity-class views of an NHWC (B, T1, F1, d) tensor
int class_maps(CUtensorMap* cm, const void* base, int B, int T1, int F1, int d, uint32_t trows) {
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const uint64_t Ti = (uint64_t)(T1 - ph + 1) / 2, Fj = (uint64_t)(F1 - pw + 1) / 2;
      const uint8_t* b0 = reinterpret_cast<const uint8_t*>(base) + ((size_t)ph * F1 + pw) * d * 2;
      if (Ti == 0 || Fj == 0) {  // degenerate (T1 == 1): point at a 1-row view; coordinates will be out of bounds anyway
        int rc = make_map4(&cm[ph * 2 + pw], base, d, 1, 1, B, (uint64_t)2 * d * 2, (uint64_t)2 * F1 * d * 2,
                           (uint64_t)T1 * F1 * d * 2, trows);
        if (rc) return rc;
        continue;
      }
      int rc = make_map4(&cm[ph * 2 + pw], b0, d, Fj, Ti, B, (uint64_t)2 * d * 2, (uint64_t)2 * F1 * d * 2,
                         (uint64_t)T1 * F1 * d * 2, trows);
      if (rc) return rc;
    }
  return TASR_OK;
}

template <int MODE, int BN>
int launch_conv(const CUtensorMap* cm, const CUtensorMap& tmW, const CUtensorMap& tmZ, const CUtensorMap& tmY, ConvDev& p,
                cudaStream_t st) {
  constexpr int SMEM = CV_STAGES * (BM * BK * 2 + BN * BK * 2) + 2 * CV_RINGG * STAGE_BYTES + (2 * CV_STAGES + 4) * 8 + 16 + 1024;
  static_assert(SMEM <= 232448, "shared memory budget");
  auto kern = conv_gemm_kernel<MODE, BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (err != cudaSuccess) return tasr_set_cuda_error(err);
    attr_done = true;
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.splits;
  const int grid = (int)(total < g_sms ? total : g_sms);
  cudaError_t lerr = launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), (size_t)SMEM, st, cm[0], cm[1], cm[2], cm[3], tmW, tmZ, tmY, p);
  if (lerr != cudaSuccess) return tasr_set_cuda_error(lerr);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

bool conv_shape_ok(int B, int T, int F, int d) {
  if (B <= 0 || T <= 0 || F <= 0 || d % 128 || d > 1024) return false;
  Geom g = geom(T, F);
  return g.F2 % 4 == 0 && g.F1 == 2 * g.F2;  // 80 mel bins -> 40 -> 20 (model/conformer.py:157)
}

}  // namespace

extern "C" int tasr_conv1_fwd(const float* x, int B, int T, int F, int d, const float* w1, const float* b1, void* y1,
                              tasr_stream_t stream) {
  if (d % 8 || d > 2048 || (NT % (d / 8)) || B <= 0 || T <= 0 || F <= 0) return TASR_ERR_SHAPE;
  Geom g = geom(T, F);
  const long long total = (long long)B * g.T1 * ((g.F1 + PX - 1) / PX);
  const int per = NT / (d / 8);
  const int grid = (int)imin64((long long)148 * 8, (total + per - 1) / per);
  conv1_fwd_kernel<<<grid, NT, (size_t)10 * d * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      x, B, T, F, d, w1, b1, g.T1, g.F1, reinterpret_cast<bf16*>(y1));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_conv1_bwd(const void* dy1, const float* x, int B, int T, int F, int d, const float* w1, const float* b1,
                              float* dw1, float* db1, tasr_stream_t stream) {
  if (d % 8 || d > 1024 || (NT % (d / 8)) || B <= 0 || T <= 0 || F <= 0) return TASR_ERR_SHAPE;
  Geom g = geom(T, F);
  const long long total = (long long)B * g.T1 * ((g.F1 + PX - 1) / PX);
  const int per = NT / (d / 8);
  const int grid = (int)imin64((long long)148 * 4, (total + per - 1) / per);
  conv1_bwd_kernel<<<grid, NT, (size_t)20 * d * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(dy1), x, B, T, F, d, w1, b1, g.T1, g.F1, dw1, db1);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

// y1 (B,T1,F1,d) bf16; w2p (d, 9d) bf16 packed (co, kh, kw, ci); bias (d) fp32 -> z2 (pre-activation, may be NULL ->
// still written to y2's buffer is NOT done; pass a buffer), y2 = silu(z2); both (B, T2, F2, d) bf16
extern "C" int tasr_conv2_fwd(const void* y1, int B, int T, int F, int d, const void* w2p, const float* bias, void* z2,
                              void* y2, tasr_stream_t stream) {
  if (!conv_shape_ok(B, T, F, d) || z2 == nullptr || y2 == nullptr) return TASR_ERR_SHAPE;
  Geom g = geom(T, F);
  CUtensorMap cm[4], tmW, tmZ, tmY;
  int rc = class_maps(cm, y1, B, g.T1, g.F1, d, 32);
  if (rc) return rc;
  if ((rc = make_map2(&tmW, w2p, (uint64_t)9 * d, d, (uint64_t)9 * d))) return rc;
  const uint64_t sf = (uint64_t)d * 2, st_ = (uint64_t)g.F2 * d * 2, sb = (uint64_t)g.T2 * g.F2 * d * 2;
  if ((rc = make_map4(&tmZ, z2, d, g.F2, g.T2, B, sf, st_, sb, 32))) return rc;
  if ((rc = make_map4(&tmY, y2, d, g.F2, g.T2, B, sf, st_, sb, 32))) return rc;
  ConvDev p = {};
  p.B = B; p.d = d; p.T2 = g.T2; p.F2 = g.F2;
  p.nt_blk = cdiv(g.T2, 32); p.nf_blk = g.F2 / 4;
  p.kc = d / 64;
  p.total_kb = 9 * p.kc; p.kb_per_split = p.total_kb; p.splits = 1;
  p.tiles_m = B * p.nt_blk * p.nf_blk;
  p.bias = bias;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d % 256 == 0) { p.tiles_n = d / 256; return launch_conv<CONV_FWD, 256>(cm, tmW, tmZ, tmY, p, st); }
  p.tiles_n = d / 128;
  return launch_conv<CONV_FWD, 128>(cm, tmW, tmZ, tmY, p, st);
}

// dz2 (B,T2,F2,d) bf16, w2p (d, 9d) -> dy1 (B,T1,F1,d) bf16 (every element written exactly once)
extern "C" int tasr_conv2_dgrad(const void* dz2, int B, int T, int F, int d, const void* w2p, void* dy1,
                                tasr_stream_t stream) {
  if (!conv_shape_ok(B, T, F, d)) return TASR_ERR_SHAPE;
  Geom g = geom(T, F);
  CUtensorMap cm[4], tmW, tmZ;
  int rc = class_maps(cm, dy1, B, g.T1, g.F1, d, 32);
  if (rc) return rc;
  if ((rc = make_map2(&tmW, w2p, (uint64_t)9 * d, d, (uint64_t)9 * d))) return rc;
  const uint64_t sf = (uint64_t)d * 2, st_ = (uint64_t)g.F2 * d * 2, sb = (uint64_t)g.T2 * g.F2 * d * 2;
  if ((rc = make_map4(&tmZ, dz2, d, g.F2, g.T2, B, sf, st_, sb, 32))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const int Ti = (g.T1 - ph + 1) / 2, Fj = (g.F1 - pw + 1) / 2;
      if (Ti <= 0 || Fj <= 0) continue;
      ConvDev p = {};
      p.B = B; p.d = d; p.T2 = g.T2; p.F2 = g.F2;
      p.ph = ph; p.pw = pw;
      p.nt_blk = cdiv(Ti, 32); p.nf_blk = cdiv(Fj, 4);
      p.kc = d / 64;
      p.ntaps = 0;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
          if (((kh + 1) & 1) == ph && ((kw + 1) & 1) == pw) p.taps[p.ntaps++] = kh * 3 + kw;
      p.total_kb = p.ntaps * p.kc; p.kb_per_split = p.total_kb; p.splits = 1;
      p.tiles_m = B * p.nt_blk * p.nf_blk;
      if (d % 256 == 0) { p.tiles_n = d / 256; rc = launch_conv<CONV_DGRAD, 256>(cm, tmW, tmZ, tmZ, p, st); }
      else { p.tiles_n = d / 128; rc = launch_conv<CONV_DGRAD, 128>(cm, tmW, tmZ, tmZ, p, st); }
      if (rc) return rc;
    }
  return TASR_OK;
}

// dW2 (d, d, 3, 3) fp32 += sum_pix dz2[pix, co] * y1[pix@tap, ci]
extern "C" int tasr_conv2_wgrad(const void* dz2, const void* y1, int B, int T, int F, int d, float* dw2,
                                tasr_stream_t stream) {
  if (!conv_shape_ok(B, T, F, d)) return TASR_ERR_SHAPE;
  Geom g = geom(T, F);
  CUtensorMap cm[4], tmZ;
  int rc = class_maps(cm, y1, B, g.T1, g.F1, d, 16);
  if (rc) return rc;
  const uint64_t sf = (uint64_t)d * 2, st_ = (uint64_t)g.F2 * d * 2, sb = (uint64_t)g.T2 * g.F2 * d * 2;
  if ((rc = make_map4(&tmZ, dz2, d, g.F2, g.T2, B, sf, st_, sb, 16))) return rc;
  ConvDev p = {};
  p.B = B; p.d = d; p.T2 = g.T2; p.F2 = g.F2;
  p.nt_blk = cdiv(g.T2, 16); p.nf_blk = g.F2 / 4;
  p.kc = d / 64;
  p.total_kb = B * p.nt_blk * p.nf_blk;
  p.tiles_m = d / 128;
  p.dw = dw2;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int bn = (d % 256 == 0) ? 256 : 128;
  p.tiles_n = 9 * (d / bn);
  int splits = (2 * g_sms + p.tiles_m * p.tiles_n - 1) / (p.tiles_m * p.tiles_n);
  if (splits < 1) splits = 1;
  if (splits > p.total_kb) splits = p.total_kb;
  p.kb_per_split = cdiv(p.total_kb, splits);
  p.splits = cdiv(p.total_kb, p.kb_per_split);
  if (bn == 256) return launch_conv<CONV_WGRAD, 256>(cm, cm[0], tmZ, tmZ, p, st);
  return launch_conv<CONV_WGRAD, 128>(cm, cm[0], tmZ, tmZ, p, st);
}
