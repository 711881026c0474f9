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
#include <stdlib.h>
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

// forward stages two output tensors (z2, y2) per column group: 3 operand stages + 2 staging buffers per epilogue group;
// dgrad / wgrad stage one: the second buffer is traded for a fourth operand stage (deeper TMA look-ahead)
__host__ __device__ constexpr int cv_stages(int mode, int bn) { return (mode == 0 || bn < 256) ? 3 : 4; }
__host__ __device__ constexpr int cv_ringg(int mode, int bn) { return (mode == 0 || bn < 256) ? 2 : 1; }

// MC: the CTA pair (2k, 2k+1) is a cluster that walks neighbouring M tiles in lockstep (same N tile, same k-blocks);
// each CTA fetches half of the weight (B) tile of a k-block and TMA-multicasts it to both, so the L2 -> shared-memory
// traffic per k-block drops from 48 KB to 32 KB per SM (the main loop is operand-feed bound).  A stage is released to the
// producers when BOTH CTAs' UMMAs have consumed it (multicast tcgen05.commit on the empty barriers).
template <int MODE, int BN, bool MC>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap cm0, const __grid_constant__ CUtensorMap cm1,
                 const __grid_constant__ CUtensorMap cm2, const __grid_constant__ CUtensorMap cm3,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmZ,
                 const __grid_constant__ CUtensorMap tmY, const ConvDev p) {
  // cm0..3 : parity-class maps of y1 (FWD, WGRAD) or dy1 (DGRAD, output)      index = ph*2 + pw
  // tmW    : W2p (d, 9d) 2-D           tmZ : z2 (FWD out2) | dz2 (DGRAD/WGRAD in), 4-D     tmY : y2 (FWD out), 4-D
  constexpr int STAGES = cv_stages(MODE, BN);
  constexpr int CV_RINGG = cv_ringg(MODE, BN);
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr bool A_MN = (MODE == CONV_WGRAD);
  constexpr bool B_MN = (MODE != CONV_FWD);

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
      mbar_init(&empty_bar[s], MC ? 2 : 1);
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
  if (MC) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast write
  grid_dependency_wait();  // prologue above overlaps the previous kernel's tail (PDL)
  const int crank = MC ? (int)cluster_ctarank() : 0;
  // MC: both CTAs of a pair run the same number of tiles; a pair whose second tile does not exist repeats the last
  // tile (identical values are written twice)
  const int tile_end = MC ? total_tiles + crank : total_tiles;

  auto class_map = [&](int idx) -> const CUtensorMap* {
    return idx == 0 ? &cm0 : (idx == 1 ? &cm1 : (idx == 2 ? &cm2 : &cm3));
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile_raw = blockIdx.x; tile_raw < tile_end; tile_raw += gridDim.x) {
        const int tile = min(tile_raw, total_tiles - 1);
        const int n_tile = tile % p.tiles_n;
        const int m_tile = (tile / p.tiles_n) % p.tiles_m;
        const int split = tile / (p.tiles_n * p.tiles_m);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);  // MC: half of B comes from the peer CTA
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
            if (MC) {
#pragma unroll
              for (int jj = 0; jj < BN / 128; ++jj) {
                const int j = crank * (BN / 128) + jj;
                tma_load_2d_mc(b_dst + j * 8192, &tmW, &full_bar[s], tap * p.d + c0, n0 + 64 * j, (uint16_t)3);
              }
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmW, &full_bar[s], tap * p.d + c0, n0 + 64 * j);
            }
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
            if (MC) {
#pragma unroll
              for (int jj = 0; jj < BN / 128; ++jj) {
                const int j = crank * (BN / 128) + jj;
                tma_load_2d_mc(b_dst + j * 8192, &tmW, &full_bar[s], tap * p.d + n0 + 64 * j, co0, (uint16_t)3);
              }
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmW, &full_bar[s], tap * p.d + n0 + 64 * j, co0);
            }
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
      for (int tile_raw = blockIdx.x; tile_raw < tile_end; tile_raw += gridDim.x, ++tcount) {
        const int tile = min(tile_raw, total_tiles - 1);
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
          if (MC) umma_commit_mc(&empty_bar[s], (uint16_t)3);
          else umma_commit(&empty_bar[s]);
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
    for (int tile_raw = blockIdx.x; tile_raw < tile_end; tile_raw += gridDim.x, ++tcount) {
      const int tile = min(tile_raw, total_tiles - 1);
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
                v[i] += p.bias[col0 + i];
                y[i] = silu_tanh(v[i]);
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
  if (MC) cluster_sync_all();  // the peer may still multicast into this CTA's shared memory / arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// conv1 (1 -> d channels, K = 9): direct on the CUDA cores (issue bound, not HBM bound: 9 FMA + SiLU per output).
// Work item = one output row segment of PX columns x (32 lanes x CH channels); a warp walks items on its own (no CTA
// barrier in the loop, so the FMA / MUFU / store phases of different warps overlap).  The 3 x (2 PX + 1) input patch of
// an item is staged by the warp in its private shared-memory slot, every value stored twice (x, x) so one LDS.64 feeds
// the packed fp32x2 FMAs (FFMA2 on channel pairs); the next item's patch is in flight while this one is computed.
// SiLU uses one MUFU.TANH: silu(z) = h + h tanh(h), h = z / 2.
// ------------------------------------------------------------------------------------------------
constexpr int NT = 256;
constexpr int NWARP = NT / 32;
constexpr int PX = 5;               // output columns per item (F1 = 40 -> 8 items per row)
constexpr int PCOLS = 2 * PX + 1;   // 11 patch columns; 3 x 11 = 33 patch values: lane l stages value l, lane 0 also value 32
constexpr int PSLOT = 3 * (PCOLS + 1);

// A warp keeps one (column group, channel slice) for the whole kernel and walks output rows with a fixed stride, so the
// per-item index arithmetic is a handful of adds (no divisions), and the backward's accumulators stay on one slice.
struct WarpWalk {
  int w0, c0, npx;         // first output column, first channel of this lane, valid columns (<= PX)
  int row, rstep, nrows;   // current output row (b * T1 + h), stride, B * T1
  bool active;
  // patch staging (lane l stages patch value l = (r0, j0); lane 0 also value 32 = (2, PCOLS - 1))
  int r0, j0, ff0, ff1, pb, ph;
  bool colok0, colok1;
  float v0, v1;
  int T, F, T1;
  const float* x;

  __device__ __forceinline__ void init(const float* x_, int B, int T_, int F_, int T1_, int F1, int d, int chw) {
    const int lane = threadIdx.x & 31;
    x = x_; T = T_; F = F_; T1 = T1_;
    nrows = B * T1;
    const int wgroups = (F1 + PX - 1) / PX, nch = (d + chw - 1) / chw;
    const int slices = wgroups * nch;
    const int nwarps = gridDim.x * NWARP;
    const int used = (nwarps / slices) * slices;
    const int gw = blockIdx.x * NWARP + (threadIdx.x >> 5);
    active = gw < used;
    const int slice = gw % slices;
    w0 = (slice / nch) * PX;
    c0 = (slice % nch) * chw + lane * (chw / 32);
    npx = min(PX, F1 - w0);
    row = gw / slices;
    rstep = used / slices;
    r0 = lane / PCOLS;
    j0 = lane - r0 * PCOLS;
    ff0 = 2 * w0 - 1 + j0;
    ff1 = 2 * w0 - 1 + PCOLS - 1;
    colok0 = ff0 >= 0 && ff0 < F;
    colok1 = lane == 0 && ff1 < F;
    pb = row / T1;
    ph = row - pb * T1;
  }
  // loads the patch of the row (pb, ph) and advances (pb, ph) by rstep rows
  __device__ __forceinline__ void prefetch(bool valid) {
    v0 = v1 = 0.f;
    if (valid) {
      const int tt0 = 2 * ph - 1 + r0, tt1 = 2 * ph + 1;
      const float* xb = x + (long long)pb * T * F;
      if (colok0 && tt0 >= 0 && tt0 < T) v0 = xb[tt0 * F + ff0];
      if (colok1 && tt1 < T) v1 = xb[tt1 * F + ff1];
    }
    ph += rstep;
    while (ph >= T1) {
      ph -= T1;
      ++pb;
    }
  }
  __device__ __forceinline__ void commit(float2* __restrict__ slot) const {
    slot[r0 * (PCOLS + 1) + j0] = make_float2(v0, v0);
    if ((threadIdx.x & 31) == 0) slot[2 * (PCOLS + 1) + PCOLS - 1] = make_float2(v1, v1);
    __syncwarp();
  }
};

// z[px][q2] (channel pairs) = b1 + sum_k patch[k of px] * w1[k]
template <int CH>
__device__ __forceinline__ void conv1_cols(const float2* __restrict__ patch, const float* __restrict__ w1s,
                                           const float* __restrict__ b1s, int d, int c0, float2 (*z)[CH / 2]) {
#pragma unroll
  for (int q = 0; q < CH / 2; ++q) {
    const float2 bv = *reinterpret_cast<const float2*>(b1s + c0 + 2 * q);
#pragma unroll
    for (int px = 0; px < PX; ++px) z[px][q] = bv;
  }
#pragma unroll 1
  for (int kh = 0; kh < 3; ++kh) {  // not unrolled: keeps only one kernel row of weights (3 x CH) live
    const float2* xrow = patch + kh * (PCOLS + 1);
    const float* wrow = w1s + kh * 3 * d + c0;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      float2 wv[CH / 2];
#pragma unroll
      for (int q = 0; q < CH / 4; ++q) {
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + kw * d + 4 * q);
        wv[2 * q] = make_float2(w4.x, w4.y);
        wv[2 * q + 1] = make_float2(w4.z, w4.w);
      }
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        const float2 xx = xrow[2 * px + kw];
#pragma unroll
        for (int q = 0; q < CH / 2; ++q) z[px][q] = __ffma2_rn(xx, wv[q], z[px][q]);
      }
    }
  }
}

__device__ __forceinline__ void stage_conv1_weights(float* w1s, float* b1s, const float* __restrict__ w1,
                                                    const float* __restrict__ b1, int d, int dpad) {
  for (int i = threadIdx.x; i < 9 * dpad; i += NT) {
    const int k = i / dpad, c = i - k * dpad;
    w1s[i] = c < d ? w1[c * 9 + k] : 0.f;
  }
  for (int i = threadIdx.x; i < dpad; i += NT) b1s[i] = i < d ? b1[i] : 0.f;
}

// silu on a channel pair: h + h tanh(h), h = z / 2  (packed FMUL2 / FFMA2 around two MUFU.TANH)
__device__ __forceinline__ uint32_t silu_pack(float2 z) {
  const float2 h = __fmul2_rn(z, make_float2(0.5f, 0.5f));
  const float2 t = make_float2(tanh_approx(h.x), tanh_approx(h.y));
  const float2 o = __ffma2_rn(h, t, h);
  return pack_bf16x2(o.x, o.y);
}

// x (B, T, F) fp32 -> y1 (B, T1, F1, d) bf16 = silu(conv1(x));  8 channels x PX columns per thread
__global__ void __launch_bounds__(NT, 2) conv1_fwd_kernel(const float* __restrict__ x, int B, int T, int F, int d, int dpad,
                                                          const float* __restrict__ w1, const float* __restrict__ b1, int T1,
                                                          int F1, bf16* __restrict__ y1) {
  extern __shared__ __align__(16) float sh_w[];
  float* w1s = sh_w;                 // [9][dpad]
  float* b1s = sh_w + 9 * dpad;      // [dpad]
  float2* slots = reinterpret_cast<float2*>(b1s + dpad) + (threadIdx.x >> 5) * 2 * PSLOT;  // per warp [2][PSLOT]
  stage_conv1_weights(w1s, b1s, w1, b1, d, dpad);
  __syncthreads();
  WarpWalk wk;
  wk.init(x, B, T, F, T1, F1, d, 256);
  if (!wk.active) return;
  wk.prefetch(wk.row < wk.nrows);
  wk.commit(slots);
  bf16* yp = y1 + ((long long)wk.row * F1 + wk.w0) * d + wk.c0;
  const long long ystep = (long long)wk.rstep * F1 * d;
  const bool cok = wk.c0 < d;
  int buf = 0;
  for (int row = wk.row; row < wk.nrows; row += wk.rstep, buf ^= 1, yp += ystep) {
    wk.prefetch(row + wk.rstep < wk.nrows);
    float2 z[PX][4];
    conv1_cols<8>(slots + buf * PSLOT, w1s, b1s, dpad, wk.c0, z);
    if (cok) {
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        if (px < wk.npx) {
          uint4 o;
          o.x = silu_pack(z[px][0]);
          o.y = silu_pack(z[px][1]);
          o.z = silu_pack(z[px][2]);
          o.w = silu_pack(z[px][3]);
          *reinterpret_cast<uint4*>(yp + (long long)px * d) = o;
        }
      }
    }
    wk.commit(slots + (buf ^ 1) * PSLOT);
  }
}

// dy1 (B, T1, F1, d) bf16 -> dW1 (d,1,3,3), db1 (d)  (+=);  z1 is recomputed from x.  4 channels x PX columns per thread
// (36 + 4 gradient accumulators in registers for the whole kernel) so that two CTAs fit an SM.
__global__ void __launch_bounds__(NT, 2) conv1_bwd_kernel(const bf16* __restrict__ dy1, const float* __restrict__ x, int B,
                                                          int T, int F, int d, int dpad, const float* __restrict__ w1,
                                                          const float* __restrict__ b1, int T1, int F1, float* __restrict__ dw1,
                                                          float* __restrict__ db1) {
  extern __shared__ __align__(16) float sh_w[];
  float* w1s = sh_w;
  float* b1s = sh_w + 9 * dpad;
  float* red = b1s + dpad;  // [10][dpad]
  float2* slots = reinterpret_cast<float2*>(red + 10 * dpad) + (threadIdx.x >> 5) * 2 * PSLOT;
  stage_conv1_weights(w1s, b1s, w1, b1, d, dpad);
  for (int i = threadIdx.x; i < 10 * dpad; i += NT) red[i] = 0.f;
  __syncthreads();
  WarpWalk wk;
  wk.init(x, B, T, F, T1, F1, d, 128);
  const int c0 = wk.c0;
  const bool cok = c0 < d;
  if (wk.active) {
    float2 gw[9][2], gb[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      gb[q] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 9; ++k) gw[k][q] = make_float2(0.f, 0.f);
    }
    wk.prefetch(wk.row < wk.nrows);
    wk.commit(slots);
    const bf16* gp = dy1 + ((long long)wk.row * F1 + wk.w0) * d + c0;
    const long long gstep = (long long)wk.rstep * F1 * d;
    const float2 half2 = make_float2(0.5f, 0.5f), one2 = make_float2(1.f, 1.f), mone2 = make_float2(-1.f, -1.f);
    int buf = 0;
    for (int row = wk.row; row < wk.nrows; row += wk.rstep, buf ^= 1, gp += gstep) {
      wk.prefetch(row + wk.rstep < wk.nrows);
      const float2* patch = slots + buf * PSLOT;
      uint2 u[PX];
#pragma unroll
      for (int px = 0; px < PX; ++px)
        u[px] = (cok && px < wk.npx) ? *reinterpret_cast<const uint2*>(gp + (long long)px * d) : make_uint2(0, 0);
      float2 z[PX][2];
      conv1_cols<4>(patch, w1s, b1s, dpad, c0, z);
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        const uint32_t uu[2] = {u[px].x, u[px].y};
        float2 dz[2];  // 2 x dz: silu'(z) = (1 + t)(1 + h (1 - t)) / 2 with h = z / 2, t = tanh(h); the 1/2 is applied at the end
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float2 g = unpack_bf16x2(uu[q]);
          const float2 h = __fmul2_rn(z[px][q], half2);
          const float2 t = make_float2(tanh_approx(h.x), tanh_approx(h.y));
          const float2 a1 = __fadd2_rn(t, one2);
          const float2 b1m = __ffma2_rn(t, mone2, one2);
          const float2 c = __ffma2_rn(h, b1m, one2);
          dz[q] = __fmul2_rn(g, __fmul2_rn(a1, c));
          gb[q] = __fadd2_rn(gb[q], dz[q]);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float2 xx = patch[(k / 3) * (PCOLS + 1) + 2 * px + (k % 3)];
          gw[k][0] = __ffma2_rn(dz[0], xx, gw[k][0]);
          gw[k][1] = __ffma2_rn(dz[1], xx, gw[k][1]);
        }
      }
      wk.commit(slots + (buf ^ 1) * PSLOT);
    }
    if (cok) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          atomicAdd(&red[k * dpad + c0 + 2 * q], 0.5f * gw[k][q].x);
          atomicAdd(&red[k * dpad + c0 + 2 * q + 1], 0.5f * gw[k][q].y);
        }
        atomicAdd(&red[9 * dpad + c0 + 2 * q], 0.5f * gb[q].x);
        atomicAdd(&red[9 * dpad + c0 + 2 * q + 1], 0.5f * gb[q].y);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * dpad; i += NT) {
    const int k = i / dpad, c = i - k * dpad;
    if (c < d) {
      if (k < 9) atomicAdd(dw1 + c * 9 + k, red[i]);
      else atomicAdd(db1 + c, red[i]);
    }
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
  constexpr int CV_STAGES = cv_stages(MODE, BN), CV_RINGG = cv_ringg(MODE, BN);
  constexpr int SMEM = CV_STAGES * (BM * BK * 2 + BN * BK * 2) + 2 * CV_RINGG * STAGE_BYTES + (2 * CV_STAGES + 4) * 8 + 16 + 1024;
  static_assert(SMEM <= 232448, "shared memory budget");
  const long long total = (long long)p.tiles_m * p.tiles_n * p.splits;
  cudaError_t lerr;
  // forward / dgrad at full width: CTA pairs share the weight tile by TMA multicast (TASR_CONV_MC=0 disables)
  static const bool mc_enabled = [] { const char* e = getenv("TASR_CONV_MC"); return !(e && e[0] == '0'); }();
  if (MODE != CONV_WGRAD && BN == 256 && p.tiles_n == 1 && p.splits == 1 && total >= 2 && g_sms >= 2 && mc_enabled) {
    auto kern = conv_gemm_kernel<MODE, BN, (MODE != CONV_WGRAD && BN == 256)>;
    static TasrPerDevice attr_done;
    if (!attr_done.get()) {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (err != cudaSuccess) return tasr_set_cuda_error(err);
      attr_done.set();
    }
    const long long pairs = (total + 1) / 2;
    const int grid = 2 * (int)(pairs < g_sms / 2 ? pairs : g_sms / 2);
    lerr = launch_pdl_cluster(kern, dim3(grid), dim3(GEMM_THREADS), (size_t)SMEM, st, 2, cm[0], cm[1], cm[2], cm[3], tmW, tmZ,
                              tmY, p);
  } else {
    auto kern = conv_gemm_kernel<MODE, BN, false>;
    static TasrPerDevice attr_done;
    if (!attr_done.get()) {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (err != cudaSuccess) return tasr_set_cuda_error(err);
      attr_done.set();
    }
    const int grid = (int)(total < g_sms ? total : g_sms);
    lerr = launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), (size_t)SMEM, st, cm[0], cm[1], cm[2], cm[3], tmW, tmZ, tmY, p);
  }
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

// tensor-core path (conv1_tc.cu): TASR_ERR_SHAPE when the shape is not covered
int tasr_conv1_tc_fwd(const float* x, int B, int T, int F, int d, const float* w1, const float* b1, void* y1, cudaStream_t st);
int tasr_conv1_tc_bwd(const void* dy1, const float* x, int B, int T, int F, int d, const float* w1, const float* b1,
                      float* dw1, float* db1, cudaStream_t st);

extern "C" int tasr_conv1_fwd(const float* x, int B, int T, int F, int d, const float* w1, const float* b1, void* y1,
                              tasr_stream_t stream) {
  if (d % 8 || d > 1024 || B <= 0 || T <= 0 || F <= 0) return TASR_ERR_SHAPE;
  {
    const int rc = tasr_conv1_tc_fwd(x, B, T, F, d, w1, b1, y1, reinterpret_cast<cudaStream_t>(stream));
    if (rc != TASR_ERR_SHAPE) return rc;
  }
  Geom g = geom(T, F);
  const int dpad = cdiv(d, 256) * 256;
  const int slices = cdiv(g.F1, PX) * (dpad / 256);
  const long long total = (long long)B * g.T1 * slices;
  if (total > 0x3fffffffLL || (long long)T * F > 0x3fffffffLL) return TASR_ERR_SHAPE;
  int grid = (int)imin64(cdiv(total, NWARP), 2 * g_sms);
  if (grid * NWARP < slices) grid = cdiv(slices, NWARP);
  const size_t smem = ((size_t)10 * dpad + NWARP * 2 * PSLOT * 2) * sizeof(float);
  static TasrPerDevice attr_done;
  if (!attr_done.get()) {
    cudaError_t e = cudaFuncSetAttribute(conv1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done.set();
  }
  conv1_fwd_kernel<<<grid, NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, B, T, F, d, dpad, w1, b1, g.T1, g.F1,
                                                                              reinterpret_cast<bf16*>(y1));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_conv1_bwd(const void* dy1, const float* x, int B, int T, int F, int d, const float* w1, const float* b1,
                              float* dw1, float* db1, tasr_stream_t stream) {
  if (d % 8 || d > 1024 || B <= 0 || T <= 0 || F <= 0) return TASR_ERR_SHAPE;
  {
    const int rc = tasr_conv1_tc_bwd(dy1, x, B, T, F, d, w1, b1, dw1, db1, reinterpret_cast<cudaStream_t>(stream));
    if (rc != TASR_ERR_SHAPE) return rc;
  }
  Geom g = geom(T, F);
  const int dpad = cdiv(d, 128) * 128;
  const int slices = cdiv(g.F1, PX) * (dpad / 128);
  const long long total = (long long)B * g.T1 * slices;
  if (total > 0x3fffffffLL || (long long)T * F > 0x3fffffffLL) return TASR_ERR_SHAPE;
  int grid = (int)imin64(cdiv(total, NWARP), 2 * g_sms);
  if (grid * NWARP < slices) grid = cdiv(slices, NWARP);
  const size_t smem = ((size_t)20 * dpad + NWARP * 2 * PSLOT * 2) * sizeof(float);
  static TasrPerDevice attr_done;
  if (!attr_done.get()) {
    cudaError_t e = cudaFuncSetAttribute(conv1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done.set();
  }
  conv1_bwd_kernel<<<grid, NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(dy1), x, B, T, F, d,
                                                                              dpad, w1, b1, g.T1, g.F1, dw1, db1);
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
