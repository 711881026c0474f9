// bf16 GEMM on tcgen05 / TMEM: persistent, warp-specialised, TMA in and TMA out, fused epilogues.
//
//   grid  = min(#tiles, #SMs) persistent CTAs, static round-robin tile schedule (n fastest, so CTAs that
//           run together share A rows in L2)
//   warp 0     : TMA producer       (smem ring of STAGES x {A 128x64, B BNx64} bf16, 128 B swizzle)
//   warp 1     : TMEM allocator + UMMA issuer (cta_group::1, 128 x BN x 16 per instruction)
//   warps 2..5 : epilogue: TMEM -> registers -> fused math -> swizzled smem staging -> TMA store
//   TMEM holds TWO accumulators (2 x BN fp32 columns): the epilogue of tile i overlaps the main loop of
//   tile i+1, which matters because most GEMMs of this model have K = 256 (4 k-blocks per tile).
//
// Replaces (reference): every nn.Linear / 1x1 Conv1d / Conv2d-as-GEMM and their autograd backward, see
// include/tasr_kernels.h.
#include "gemm_common.cuh"
#include <stdio.h>
#include <string.h>
#include <mutex>

namespace {

// ------------------------------------------------------------------------------------------------
// the kernel
//   EPI    epilogue mode (compile time)          BN     accumulator width (dual-B modes: BN/2 output cols)
//   STAGES smem pipeline depth                   RINGG  staging buffers per epilogue group
//   two epilogue groups of 4 warps each take alternate 64-column groups of a tile
// ------------------------------------------------------------------------------------------------
template <int EPI, int BN, int STAGES, int RINGG, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2, const GemmDev p) {
  constexpr bool DUAL = (EPI == TASR_EPI_SWIGLU || EPI == TASR_EPI_GLU);
  constexpr int A_BYTES = BM * BK * 2;  // 16 KB
  constexpr int B_BYTES = BN * BK * 2;
  // Weight gradients (split-K, EPI_ATOMIC): a CTA gets one tile (the grid is one wave), so a single accumulator is
  // enough; the 16 columns behind it hold the bias-gradient accumulator (A x ones).
  constexpr bool WG = (EPI == TASR_EPI_ATOMIC);
  constexpr int NACC = WG ? 1 : 2;
  constexpr int ONES_BYTES = WG ? 512 : 0;
  constexpr uint32_t TMEM_COLS = WG ? (BN + 16 <= 256 ? 256u : 512u) : 2u * BN;  // power of two
  constexpr int TILE_N = DUAL ? BN / 2 : BN;
  constexpr int NBUF = DUAL ? 3 : ((EPI == TASR_EPI_SILU || EPI == TASR_EPI_SWIGLU_BWD || EPI == TASR_EPI_GLU_BWD) ? 2 : 1);
  static_assert(RINGG >= NBUF, "staging ring too small");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sC = sB + STAGES * B_BYTES;  // 2 groups x RINGG staging buffers
  uint8_t* sOnes = sC + 2 * RINGG * STAGE_BYTES;  // WG: 16 (N) x 16 (K) bf16 ones as 4 un-swizzled core matrices
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sOnes + ONES_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int total_tiles = p.tiles_m * p.tiles_n * p.splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmO2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], G_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (WG) {
    if (threadIdx.x < ONES_BYTES / 4) reinterpret_cast<uint32_t*>(sOnes)[threadIdx.x] = 0x3F803F80u;  // bf16 1.0 x 2
    fence_proxy_async_smem();  // the tensor core reads shared memory through the async proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dependency_wait();  // prologue above overlaps the previous kernel's tail (PDL)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.tiles_n;
        const int m_tile = (tile / p.tiles_n) % p.tiles_m;
        const int split = tile / (p.tiles_n * p.tiles_m);
        const int m0 = m_tile * BM, n0 = n_tile * TILE_N;
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait_backoff(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
          const int k0 = kb * BK;
          uint8_t* a_dst = sA + s * A_BYTES;
          uint8_t* b_dst = sB + s * B_BYTES;
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) {
            if (A_MN) tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, k0);
            else      tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[s], k0, m0 + 64 * j);
          }
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            const int nrow = DUAL ? (j < BN / 128 ? n0 + 64 * j : p.n_half + n0 + 64 * (j - BN / 128)) : n0 + 64 * j;
            if (B_MN) tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[s], nrow, k0);
            else      tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[s], k0, nrow);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t idesc_ones = umma_idesc_bf16(BM, 16, A_MN ? 1 : 0, 0);
      const uint64_t ones_desc = umma_desc_noswizzle(smem_u32(sOnes), 128, 256);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int split = tile / (p.tiles_n * p.tiles_m);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
        const uint32_t acc = NACC == 2 ? (tcount & 1u) : 0u, aph = NACC == 2 ? ((tcount >> 1) & 1u) : (tcount & 1u);
        const bool do_colsum = WG && p.colsum != nullptr && (tile % p.tiles_n) == 0;
        mbar_wait_backoff(&tempty_bar[acc], aph ^ 1u);  // epilogue has drained this accumulator
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
            // K-major: +32 B per 16-wide K step inside the 128 B swizzle row.
            // MN-major: +2 swizzle atoms (16 K-rows x 128 B) per step; 64-wide M|N chunks are 8 KB apart.
            const uint64_t adesc = A_MN ? umma_desc_sw128(a_base + k * 2048, 8192, 1024)
                                        : umma_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_desc_sw128(b_base + k * 2048, 8192, 1024)
                                        : umma_desc_sw128(b_base + k * 32, 16, 1024);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            // bias gradient: the same A k-slice against ones -> sum over k of A(m, k) in 16 identical columns
            if (do_colsum) umma_bf16(tmem_base + BN, adesc, ones_desc, idesc_ones, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage when these MMAs retire
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue: 2 super-groups x 8 warps =====================
    // A super-group owns alternate 64-column groups of the tile; inside it, warp set `hsel` takes the 32-column
    // half hsel of every group, 16 columns at a time (small register footprint -> 16 warps hide the latencies).
    const int ew = warp - 2;          // 0..15
    const int sg = ew >> 3;           // super-group
    const int hsel = (ew >> 2) & 1;   // 32-column half inside a 64-column group
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int rloc = q * 32 + lane;
    const bool leader = ((ew & 7) == 0) && lane == 0;
    const int bar_id = 1 + sg;
    uint8_t* ring_base = sC + sg * RINGG * STAGE_BYTES;
    constexpr bool F32_MODE = (EPI == TASR_EPI_RESID || EPI == TASR_EPI_ATOMIC);
    constexpr bool HAS_AUX_BF16 = (EPI == TASR_EPI_SWIGLU_BWD || EPI == TASR_EPI_GLU_BWD || EPI == TASR_EPI_SILU_BWD);
    const bool f32_out = F32_MODE || (EPI == TASR_EPI_STORE && p.out_f32);
    const bool direct_atomic = (EPI == TASR_EPI_ATOMIC) && (p.remap_p0 > 0);
    uint32_t tcount = 0, ring = 0;
    const uint32_t s32 = gemm_drop_seed32(p);  // dropout seed digest, once per thread
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int n_tile = tile % p.tiles_n;
      const int m_tile = (tile / p.tiles_n) % p.tiles_m;
      const int m0 = m_tile * BM, n0 = n_tile * TILE_N;
      const int row = m0 + rloc;
      const uint32_t acc = NACC == 2 ? (tcount & 1u) : 0u, aph = NACC == 2 ? ((tcount >> 1) & 1u) : (tcount & 1u);
      // saved-tensor operands of this thread's first column group: in flight while we wait for the accumulator
      AuxBf16 ax_first[2];
      AuxF32 res_first[2];
      if (HAS_AUX_BF16) {
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) preload_aux_bf16<EPI>(p, row, n0 + sg * 64 + hsel * 32 + sub * 16, ax_first[sub]);
      }
      if (EPI == TASR_EPI_RESID) {
#pragma unroll
        for (int u = 0; u < 2; ++u) preload_aux_f32(p, row, n0 + sg * 64 + u * 32 + hsel * 16, res_first[u]);
      }
      // The saved tensor of the NEXT tile of this CTA goes to L2 now: its loads (issued one tile later, when the
      // accumulator is usually already waiting, so nothing hides them) then cost an L2 hit instead of an HBM round trip.
      if ((HAS_AUX_BF16 || EPI == TASR_EPI_RESID) && !(p.flags & 1)) {
        const int nt_ = tile + gridDim.x;
        if (nt_ < total_tiles) {
          const int nn0 = (nt_ % p.tiles_n) * TILE_N;
          const int nrow = ((nt_ / p.tiles_n) % p.tiles_m) * BM + rloc;
          if (nrow < p.M) {
#pragma unroll 1
            for (int g = sg; g < TILE_N / 64; g += 2) {
              if (nn0 + g * 64 >= p.N) break;
              if (EPI == TASR_EPI_RESID) {  // this thread: 2 x 16 fp32 (64 B each) at +0 and +32 columns
                const float* a0 = reinterpret_cast<const float*>(p.aux) + (long long)nrow * p.ldaux + nn0 + g * 64 + hsel * 16;
                prefetch_l2(a0);
                prefetch_l2(a0 + 32);
              } else {                      // 32 bf16 (64 B) of each half
                const bf16* a0 = reinterpret_cast<const bf16*>(p.aux) + (long long)nrow * p.ldaux + nn0 + g * 64 + hsel * 32;
                prefetch_l2(a0);
                if (EPI != TASR_EPI_SILU_BWD) prefetch_l2(a0 + p.n_half);
              }
            }
          }
        }
      }
      mbar_wait(&tfull_bar[acc], aph);
      __syncwarp();
      tc_fence_after();
      const uint32_t tbase = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int g = sg; g < TILE_N / 64; g += 2) {  // this super-group's 64-column groups
        if (n0 + g * 64 >= p.N) break;              // fully out of range (uniform across the super-group)
        if (f32_out) {
          if (EPI == TASR_EPI_STORE || F32_MODE) {
            AuxF32 res[2];
            if (EPI == TASR_EPI_RESID) {
              if (g == sg) { res[0] = res_first[0]; res[1] = res_first[1]; }
              else {
#pragma unroll
                for (int u = 0; u < 2; ++u) preload_aux_f32(p, row, n0 + g * 64 + u * 32 + hsel * 16, res[u]);
              }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {  // two 32-column fp32 staging buffers per group; this warp set: 16 of the 32
              const int col0 = n0 + g * 64 + u * 32 + hsel * 16;
              uint32_t lo_u[16];
              tmem_ld16(tbase + g * 64 + u * 32 + hsel * 16, lo_u);
              tmem_ld_wait();
              float* lo = reinterpret_cast<float*>(lo_u);
              if (EPI == TASR_EPI_RESID) {
                epilogue_resid16(p, s32, row, col0, lo, res[u]);
              } else {
                float t3[16];
                epilogue_math<EPI, 16>(p, s32, row, col0, lo, lo, t3);
              }
              if (direct_atomic) {
                if (row < p.M) {
                  float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo;
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    int c = col0 + i;
                    if (c < p.N) {
                      c = (c % p.remap_p0) * p.remap_p1 + c / p.remap_p0;
                      atomicAdd(o + c, lo[i]);
                    }
                  }
                }
                continue;
              }
              if (leader) bulk_wait_read<RINGG - 1>();
              named_bar_sync(bar_id, G_SG_THREADS);
              uint8_t* buf = ring_base + (ring % RINGG) * STAGE_BYTES;
              stage_f32_16(buf, rloc, hsel * 4, lo);
              fence_proxy_async_smem();
              named_bar_sync(bar_id, G_SG_THREADS);
              if (leader) {
                const int c = n0 + g * 64 + u * 32;
                if (EPI == TASR_EPI_ATOMIC) tma_reduce_add_2d(&tmO, buf, c, m0);
                else tma_store_2d(&tmO, buf, c, m0);
                bulk_commit();
              }
              ++ring;
            }
          }
        } else if (!F32_MODE) {
          // bf16 outputs: NBUF 64-column staging buffers per group
          AuxBf16 ax[2];
          if (HAS_AUX_BF16) {
            if (g == sg) { ax[0] = ax_first[0]; ax[1] = ax_first[1]; }
            else {
#pragma unroll
              for (int sub = 0; sub < 2; ++sub) preload_aux_bf16<EPI>(p, row, n0 + g * 64 + hsel * 32 + sub * 16, ax[sub]);
            }
          }
          uint8_t* buf0 = ring_base + (ring % RINGG) * STAGE_BYTES;
          uint8_t* buf1 = ring_base + ((ring + 1) % RINGG) * STAGE_BYTES;
          uint8_t* buf2 = ring_base + ((ring + 2) % RINGG) * STAGE_BYTES;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const int coff = g * 64 + hsel * 32 + sub * 16;
            const int col0 = n0 + coff;
            const int chunk0 = hsel * 4 + sub * 2;
            uint32_t lo_u[16], hi_u[16];
            tmem_ld16(tbase + coff, lo_u);
            if (DUAL) tmem_ld16(tbase + BN / 2 + coff, hi_u);
            // rotary embedding: the partner columns (same head, other 32-column half) sit in this thread's TMEM lane
            const bool rot = (EPI == TASR_EPI_ROPE) && (n0 + g * 64 < p.remap_p0);
            if (EPI == TASR_EPI_ROPE && rot) tmem_ld16(tbase + g * 64 + (hsel ^ 1) * 32 + sub * 16, hi_u);
            tmem_ld_wait();
            float* lo = reinterpret_cast<float*>(lo_u);
            float* hi = reinterpret_cast<float*>(hi_u);
            float t3[16];
            if (HAS_AUX_BF16) epilogue_bwd16<EPI>(p, s32, row, col0, lo, hi, ax[sub]);
            else if (EPI == TASR_EPI_ROPE) epilogue_rope16(p, row, col0, hsel, sub, rot, lo, hi);
            else epilogue_math<EPI, 16>(p, s32, row, col0, lo, hi, t3);
            if (sub == 0) {
              // the staging buffers are needed only now: the math above overlapped the TMA stores (their reads of
              // these buffers) of the previous column group
              if (leader) bulk_wait_read<RINGG - NBUF>();
              named_bar_sync(bar_id, G_SG_THREADS);
            }
            if (DUAL) {
              stage_bf16_16(buf0, rloc, chunk0, t3);
              stage_bf16_16(buf1, rloc, chunk0, lo);
              stage_bf16_16(buf2, rloc, chunk0, hi);
            } else if (EPI == TASR_EPI_SILU) {
              stage_bf16_16(buf0, rloc, chunk0, t3);
              stage_bf16_16(buf1, rloc, chunk0, lo);
            } else if (EPI == TASR_EPI_SWIGLU_BWD || EPI == TASR_EPI_GLU_BWD) {
              stage_bf16_16(buf0, rloc, chunk0, lo);
              stage_bf16_16(buf1, rloc, chunk0, hi);
            } else {
              stage_bf16_16(buf0, rloc, chunk0, lo);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, G_SG_THREADS);
          if (leader) {
            const int c = n0 + g * 64;
            if (DUAL) {
              tma_store_2d(&tmO, buf0, c, m0); bulk_commit();
              tma_store_2d(&tmO2, buf1, c, m0); bulk_commit();
              tma_store_2d(&tmO2, buf2, p.n_half + c, m0); bulk_commit();
            } else if (EPI == TASR_EPI_SILU) {
              tma_store_2d(&tmO, buf0, c, m0); bulk_commit();
              if (p.out2 != nullptr) tma_store_2d(&tmO2, buf1, c, m0);
              bulk_commit();
            } else if (EPI == TASR_EPI_SWIGLU_BWD || EPI == TASR_EPI_GLU_BWD) {
              tma_store_2d(&tmO, buf0, c, m0); bulk_commit();
              tma_store_2d(&tmO, buf1, p.n_half + c, m0); bulk_commit();
            } else {
              tma_store_2d(&tmO, buf0, c, m0); bulk_commit();
            }
          }
          ring += NBUF;
        }
      }
      if (WG && p.colsum != nullptr && n_tile == 0 && sg == 0 && hsel == 0) {  // one warp per TMEM lane quarter
        uint32_t cs[16];
        tmem_ld16(tmem_base + BN + ((uint32_t)(q * 32) << 16), cs);
        tmem_ld_wait();
        if (row < p.M) atomicAdd(p.colsum + row, __uint_as_float(cs[0]) * p.alpha);
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);  // all 512 epilogue threads: this accumulator may be overwritten
    }
    if (leader) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;
std::once_flag g_encode_once;

void init_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
}

int make_tmap_2d(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* base, uint64_t inner, uint64_t outer,
                 uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  std::call_once(g_encode_once, init_encode);
  if (!g_encode) return TASR_ERR_CUDA;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * esize};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TASR_OK : TASR_ERR_CUDA;
}

}  // namespace

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows of pitch ld elements,
// 128 B swizzle, zero fill out of bounds.
int tasr_make_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                           uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_2d(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, inner, outer, ld_elems, box_inner, box_outer);
}

namespace {

int fill_dev(const tasr_gemm_args* a, GemmDev* p, bool* dual) {
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return TASR_ERR_SHAPE;
  *dual = (a->epilogue == TASR_EPI_SWIGLU || a->epilogue == TASR_EPI_GLU);
  if ((*dual || a->epilogue == TASR_EPI_SWIGLU_BWD || a->epilogue == TASR_EPI_GLU_BWD) && a->n_half != a->N)
    return TASR_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15) || (a->lda & 7) || (a->ldb & 7))
    return TASR_ERR_ALIGN;
  p->M = a->M; p->N = a->N; p->K = a->K;
  p->epi = a->epilogue; p->out_f32 = a->out_f32;
  p->out = a->out; p->ldo = a->ldo; p->out2 = a->out2; p->ldo2 = a->ldo2;
  p->bias = a->bias; p->aux = a->aux; p->ldaux = a->ldaux;
  p->alpha = a->alpha; p->n_half = a->n_half;
  p->drop_thresh = tasr_drop_thresh16(a->drop_p);
  p->drop_inv_keep = tasr_drop_inv_keep(p->drop_thresh);
  if ((a->epilogue == TASR_EPI_SWIGLU_BWD || a->epilogue == TASR_EPI_GLU_BWD || a->epilogue == TASR_EPI_SILU_BWD) &&
      ((a->N & 15) || (a->ldaux & 7) || (reinterpret_cast<uintptr_t>(a->aux) & 15) || (a->n_half & 7)))
    return TASR_ERR_ALIGN;
  if (a->epilogue == TASR_EPI_RESID && ((a->N & 15) || (a->ldaux & 3) || (reinterpret_cast<uintptr_t>(a->aux) & 15)))
    return TASR_ERR_ALIGN;
  if (p->drop_thresh && (a->N & 31)) return TASR_ERR_SHAPE;  // the pair-hash dropout mask is generated in runs of 32 columns
  p->seed = a->seed;
  p->seed_ptr = g_tasr_seed_ptr;
  p->remap_p0 = a->remap_p0; p->remap_p1 = a->remap_p1;
  p->colsum = (a->epilogue == TASR_EPI_ATOMIC) ? a->colsum : nullptr;
  static int env_flags = -1;
  if (env_flags < 0) {
    const char* e = getenv("TASR_GEMM_FLAGS");
    env_flags = e ? atoi(e) : 0;
  }
  p->flags = env_flags;
  if (a->colsum != nullptr && a->epilogue != TASR_EPI_ATOMIC) return TASR_ERR_SHAPE;
  const int num_kb = (a->K + BK - 1) / BK;
  int splits = (a->epilogue == TASR_EPI_ATOMIC && a->split_k > 1) ? a->split_k : 1;
  if (splits > num_kb) splits = num_kb;
  p->kb_per_split = (num_kb + splits - 1) / splits;
  p->splits = (num_kb + p->kb_per_split - 1) / p->kb_per_split;
  p->tiles_m = p->tiles_n = 1;
  return TASR_OK;
}

template <int EPI, int BN, int STAGES, int RINGG, bool A_MN, bool B_MN>
int launch_tc(const tasr_gemm_args* a, GemmDev& p, cudaStream_t st) {
  constexpr bool DUAL = (EPI == TASR_EPI_SWIGLU || EPI == TASR_EPI_GLU);
  constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + 2 * RINGG * STAGE_BYTES + (EPI == TASR_EPI_ATOMIC ? 512 : 0) +
                       (2 * STAGES + 4) * 8 + 16 + 1024;
  static_assert(SMEM <= 232448, "shared memory budget");
  constexpr int TILE_N = DUAL ? BN / 2 : BN;
  CUtensorMap tmA, tmB, tmO, tmO2;
  int rc;
  if (A_MN) rc = tasr_make_tmap_2d_bf16(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, 64);
  else      rc = tasr_make_tmap_2d_bf16(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, 64, 64);
  if (rc) return rc;
  const uint64_t nrows = DUAL ? (uint64_t)2 * a->n_half : (uint64_t)a->N;
  if (B_MN) rc = tasr_make_tmap_2d_bf16(&tmB, a->B, nrows, (uint64_t)a->K, (uint64_t)a->ldb, 64, 64);
  else      rc = tasr_make_tmap_2d_bf16(&tmB, a->B, (uint64_t)a->K, nrows, (uint64_t)a->ldb, 64, 64);
  if (rc) return rc;
  // output maps (TMA store clips rows >= M and columns >= the map width)
  const bool f32_out = EPI == TASR_EPI_RESID || EPI == TASR_EPI_ATOMIC || (EPI == TASR_EPI_STORE && a->out_f32);
  const bool direct_atomic = EPI == TASR_EPI_ATOMIC && a->remap_p0 > 0;
  const uint64_t out_cols = (EPI == TASR_EPI_SWIGLU_BWD || EPI == TASR_EPI_GLU_BWD) ? (uint64_t)2 * a->n_half : (uint64_t)a->N;
  if (!direct_atomic) {
    if (f32_out) {
      if ((reinterpret_cast<uintptr_t>(a->out) & 15) || (a->ldo & 3)) return TASR_ERR_ALIGN;
      rc = make_tmap_2d(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->out, out_cols, (uint64_t)a->M, (uint64_t)a->ldo, 32, 128);
    } else {
      if ((reinterpret_cast<uintptr_t>(a->out) & 15) || (a->ldo & 7)) return TASR_ERR_ALIGN;
      rc = make_tmap_2d(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->out, out_cols, (uint64_t)a->M, (uint64_t)a->ldo, 64, 128);
    }
    if (rc) return rc;
  } else {
    tmO = tmA;
  }
  if (a->out2 != nullptr && (DUAL || EPI == TASR_EPI_SILU)) {
    if ((reinterpret_cast<uintptr_t>(a->out2) & 15) || (a->ldo2 & 7)) return TASR_ERR_ALIGN;
    const uint64_t cols2 = DUAL ? (uint64_t)2 * a->n_half : (uint64_t)a->N;
    rc = make_tmap_2d(&tmO2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->out2, cols2, (uint64_t)a->M, (uint64_t)a->ldo2, 64, 128);
    if (rc) return rc;
  } else {
    if (DUAL) return TASR_ERR_SHAPE;  // dual-B modes need out2
    tmO2 = tmO;
  }
  auto kern = gemm_tc_kernel<EPI, BN, STAGES, RINGG, A_MN, B_MN>;
  static TasrPerDevice attr_done;
  if (!attr_done.get()) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (err != cudaSuccess) return tasr_set_cuda_error(err);
    attr_done.set();
  }
  p.tiles_m = cdiv(a->M, BM);
  p.tiles_n = cdiv(a->N, TILE_N);
  const long long total = (long long)p.tiles_m * p.tiles_n * p.splits;
  // as few CTAs as finish in the same number of rounds: e.g. 332 tiles -> 3 rounds -> 111 CTAs with 3 tiles each instead
  // of 148 with 2-3; the makespan is the same and the other SMs stay free for the kernels of the second stream
  const int num_sms = tasr_num_sms();
  const long long rounds = (total + num_sms - 1) / num_sms;
  int grid = (int)((total + rounds - 1) / rounds);
  if (p.flags & 2) grid = (int)(total < num_sms ? total : num_sms);  // experiment: every SM takes tiles
  cudaError_t lerr = launch_pdl(kern, dim3(grid), dim3(G_THREADS), (size_t)SMEM, st, tmA, tmB, tmO, tmO2, p);
  if (lerr != cudaSuccess) return tasr_set_cuda_error(lerr);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

// Tile width.  Wide (256) tiles halve the shared-memory traffic per flop, but with ~148 persistent CTAs the number
// of scheduling rounds is what counts for the small GEMMs of this model: pick the width with fewer
// (rounds x tile width), ties to the wide tile.
inline bool use_wide(int M, int N, int splits) {
  if (!((N % 256 == 0) || (N > 512 && (N % 256) > 128))) return false;
  const long long tm = (M + BM - 1) / BM;
  const long long wide_tiles = tm * ((N + 255) / 256) * splits, narrow_tiles = tm * ((N + 127) / 128) * splits;
  const long long sms = tasr_num_sms();
  const long long cost_wide = ((wide_tiles + sms - 1) / sms) * 256, cost_narrow = ((narrow_tiles + sms - 1) / sms) * 128;
  return cost_wide <= cost_narrow;
}

template <int EPI, bool A_MN, bool B_MN>
int launch_single(const tasr_gemm_args* a, GemmDev& p, cudaStream_t st) {
  // split-K weight gradients stream two long operands and write one small tile at the end: a single staging buffer
  // per epilogue group buys one more operand stage (deeper TMA look-ahead)
  constexpr bool LONGK = (EPI == TASR_EPI_ATOMIC);
  constexpr int R = LONGK ? 1 : 2;
  // Weight gradients run on the side stream next to the main chain: a smaller shared-memory footprint (3 x 48 KB or
  // 4 x 32 KB of operand stages instead of 4 / 6) lets their CTAs share an SM with the GroupNorm / depthwise / BatchNorm
  // kernels of the main chain.  Stand-alone the kernel is 2 % slower, the training step 1.8 % faster (measured A/B,
  // TASR_GEMM_FLAGS=4 restores the deep pipeline, 32 forces narrow tiles with 3 stages).
  if (LONGK && (p.flags & 4)) {
    if (use_wide(a->M, a->N, p.splits)) return launch_tc<EPI, 256, LONGK ? 4 : 3, R, A_MN, B_MN>(a, p, st);
    return launch_tc<EPI, 128, LONGK ? 6 : 4, R, A_MN, B_MN>(a, p, st);
  }
  if (LONGK && (p.flags & 32)) return launch_tc<EPI, 128, 3, R, A_MN, B_MN>(a, p, st);
  if (use_wide(a->M, a->N, p.splits)) return launch_tc<EPI, 256, 3, R, A_MN, B_MN>(a, p, st);
  return launch_tc<EPI, 128, 4, R, A_MN, B_MN>(a, p, st);
}

}  // namespace

extern "C" int tasr_gemm_bf16(const tasr_gemm_args* a, tasr_stream_t stream) {
  GemmDev p;
  bool dual;
  int rc = fill_dev(a, &p, &dual);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::call_once(g_encode_once, init_encode);
  const int am = a->a_mn_major ? 1 : 0, bm = a->b_mn_major ? 1 : 0;
  switch (a->epilogue) {
    case TASR_EPI_STORE:
      if (!am && !bm) return launch_single<TASR_EPI_STORE, false, false>(a, p, st);
      if (!am && bm) {
        // long reductions with a small output (FFN up-projection dgrad, K = 2 dff): MMA-bound, so trade the second
        // staging buffer for two more operand stages
        // long reductions with a small output (FFN up-projection dgrad, K = 2 dff): the round-count model of use_wide()
        // would pick 128-wide tiles here, but then the long A operand is fetched once per N tile and the main loop
        // starves on it (measured: 35.8 -> 29.2 us at d = 256, 106 -> 76 us at d = 512 with one 256-wide tile per row
        // block); a single staging buffer pays for a fourth operand stage.  TASR_GEMM_FLAGS=64 restores the narrow tiles.
        if (a->K >= 2048 && !a->out_f32 && a->N % 256 == 0 && !(p.flags & 64))
          return launch_tc<TASR_EPI_STORE, 256, 4, 1, false, true>(a, p, st);
        // mid-K data gradients (K = 512 .. 1024): six operand stages and one staging buffer beat four and two by
        // 4-8 % stand-alone (d = 512: 39.4 -> 36.9 us at K = 1024, 26.7 -> 24.6 us at K = 512); flag 128 = old rule.
        if (a->K >= ((p.flags & 128) ? 2048 : 512) && !a->out_f32 && !use_wide(a->M, a->N, p.splits))
          return launch_tc<TASR_EPI_STORE, 128, 6, 1, false, true>(a, p, st);
        return launch_single<TASR_EPI_STORE, false, true>(a, p, st);
      }
      if (am && bm) return launch_single<TASR_EPI_STORE, true, true>(a, p, st);
      return launch_tc<TASR_EPI_STORE, 128, 4, 2, true, false>(a, p, st);
    case TASR_EPI_RESID:
      if (!am && !bm) {
        // FFN down-projection (K = dff): 16-32 k-blocks per tile, so operand look-ahead is worth more than a second
        // staging buffer (34.0 -> 31.4 us at d = 256, 73.0 -> 68.8 us at d = 512; 256-wide tiles were slower)
        if (a->K >= 1024 && !use_wide(a->M, a->N, p.splits)) return launch_tc<TASR_EPI_RESID, 128, 6, 1, false, false>(a, p, st);
        return launch_single<TASR_EPI_RESID, false, false>(a, p, st);
      }
      break;
    case TASR_EPI_ROPE:
      if (am || bm || a->out_f32 || a->aux == nullptr || a->n_half <= 0 || a->remap_p0 <= 0 || (a->remap_p0 % 64) ||
          a->remap_p0 > a->N || (reinterpret_cast<uintptr_t>(a->aux) & 15))
        return TASR_ERR_SHAPE;
      return launch_single<TASR_EPI_ROPE, false, false>(a, p, st);
    case TASR_EPI_SWIGLU:
      if (am || bm || (a->n_half % 64)) break;
      if (a->n_half % 128 == 0) return launch_tc<TASR_EPI_SWIGLU, 256, 2, 3, false, false>(a, p, st);
      return launch_tc<TASR_EPI_SWIGLU, 128, 3, 3, false, false>(a, p, st);
    case TASR_EPI_GLU:
      if (am || bm || (a->n_half % 64)) break;
      if (a->n_half % 128 == 0) return launch_tc<TASR_EPI_GLU, 256, 2, 3, false, false>(a, p, st);
      return launch_tc<TASR_EPI_GLU, 128, 3, 3, false, false>(a, p, st);
    case TASR_EPI_SILU:
      if (!am && !bm) return launch_single<TASR_EPI_SILU, false, false>(a, p, st);
      break;
    case TASR_EPI_SWIGLU_BWD:
      // always 128-wide tiles: with 256-wide ones a super-group walks two column groups per tile and the saved
      // gate|up tile of the second one is fetched with nothing to hide it (measured: 50.5 vs 84.4 us at d = 256,
      // 101 vs 145 us at d = 512, where the round-count model of use_wide() used to pick the wide tile)
      if (!am && bm) {
        if (a->K >= 512) return launch_tc<TASR_EPI_SWIGLU_BWD, 128, 5, 2, false, true>(a, p, st);
        return launch_tc<TASR_EPI_SWIGLU_BWD, 128, 4, 2, false, true>(a, p, st);
      }
      break;
    case TASR_EPI_GLU_BWD:
      if (!am && bm) return launch_tc<TASR_EPI_GLU_BWD, 128, 4, 2, false, true>(a, p, st);  // see SWIGLU_BWD
      break;
    case TASR_EPI_SILU_BWD:
      if (!am && bm) return launch_single<TASR_EPI_SILU_BWD, false, true>(a, p, st);  // wide tiles measured faster here
      break;
    case TASR_EPI_ATOMIC:
      if (am && bm) return launch_single<TASR_EPI_ATOMIC, true, true>(a, p, st);
      break;
    default: break;
  }
  return TASR_ERR_SHAPE;  // epilogue / operand-major combination not instantiated (see include/tasr_kernels.h)
}
