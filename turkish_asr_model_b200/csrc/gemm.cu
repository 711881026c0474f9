// bf16 GEMM on tcgen05 / TMEM with TMA-fed shared-memory pipeline and fused epilogues.
// One CTA = one 128 x BN output tile (cta_group::1), 6 warps:
//   warp 0     : TMA producer (one elected lane)
//   warp 1     : TMEM allocator + UMMA issuer (one elected lane)
//   warps 2..5 : epilogue, one TMEM lane (= tile row) per thread
// Two CTAs fit per SM (3 stages x 32 KB, 128 TMEM columns each), so one CTA's epilogue overlaps the
// other's main loop.
#include "common.cuh"
#include <stdio.h>
#include <string.h>
#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 192;

struct GemmDev {
  int M, N, K;
  int epi;
  int out_f32;
  void* out;
  long long ldo;
  void* out2;
  long long ldo2;
  const float* bias;
  const void* aux;
  long long ldaux;
  float alpha;
  int n_half;
  uint32_t drop_thresh;
  float drop_inv_keep;
  unsigned long long seed;
  int kb_per_split;
  int remap_p0, remap_p1;
};

// ------------------------------------------------------------------------------------------------
// Epilogue on one row chunk of 32 columns (shared by the tcgen05 kernel and the debug kernel).
//   lo[]: accumulator columns [col0, col0+32);   hi[]: dual-B modes only, the paired half.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_bf16_chunk(bf16* dst, const float* v, int nvalid) {
  if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      d4[i] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) dst[i] = __float2bfloat16(v[i]);
  }
}
__device__ __forceinline__ void load_bf16_chunk(const bf16* src, float* v, int nvalid) {
  if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = s4[i];
      float2 f;
      f = unpack_bf16x2(u.x); v[8 * i + 0] = f.x; v[8 * i + 1] = f.y;
      f = unpack_bf16x2(u.y); v[8 * i + 2] = f.x; v[8 * i + 3] = f.y;
      f = unpack_bf16x2(u.z); v[8 * i + 4] = f.x; v[8 * i + 5] = f.y;
      f = unpack_bf16x2(u.w); v[8 * i + 6] = f.x; v[8 * i + 7] = f.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (i < nvalid) ? __bfloat162float(src[i]) : 0.f;
  }
}
__device__ __forceinline__ void store_f32_chunk(float* dst, const float* v, int nvalid) {
  if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nvalid) dst[i] = v[i];
  }
}
__device__ __forceinline__ void load_f32_chunk(const float* src, float* v, int nvalid) {
  if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 f = s4[i];
      v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (i < nvalid) ? src[i] : 0.f;
  }
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

__device__ __forceinline__ void epilogue_chunk(const GemmDev& p, int row, int col0, float* lo, float* hi) {
  const int nvalid = min(32, p.N - col0);
  if (nvalid <= 0) return;
  const long long r = row;
  switch (p.epi) {
    case TASR_EPI_STORE: {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float b = (p.bias != nullptr && i < nvalid) ? p.bias[col0 + i] : 0.f;
        lo[i] = p.alpha * (lo[i] + b);
      }
      if (p.out_f32)
        store_f32_chunk(reinterpret_cast<float*>(p.out) + r * p.ldo + col0, lo, nvalid);
      else
        store_bf16_chunk(reinterpret_cast<bf16*>(p.out) + r * p.ldo + col0, lo, nvalid);
    } break;
    case TASR_EPI_RESID: {
      float res[32];
      load_f32_chunk(reinterpret_cast<const float*>(p.aux) + r * p.ldaux + col0, res, nvalid);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float b = (p.bias != nullptr && i < nvalid) ? p.bias[col0 + i] : 0.f;
        float v = lo[i] + b;
        if (p.drop_thresh) v *= dropout_scale(p.seed, (unsigned long long)(r * p.N + col0 + i), p.drop_thresh, p.drop_inv_keep);
        lo[i] = res[i] + p.alpha * v;
      }
      store_f32_chunk(reinterpret_cast<float*>(p.out) + r * p.ldo + col0, lo, nvalid);
    } break;
    case TASR_EPI_SWIGLU:
    case TASR_EPI_GLU: {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float b0 = (p.bias != nullptr && i < nvalid) ? p.bias[col0 + i] : 0.f;
        float b1 = (p.bias != nullptr && i < nvalid) ? p.bias[p.n_half + col0 + i] : 0.f;
        lo[i] = bf16_round(lo[i] + b0);
        hi[i] = bf16_round(hi[i] + b1);
      }
      bf16* o2 = reinterpret_cast<bf16*>(p.out2) + r * p.ldo2;
      store_bf16_chunk(o2 + col0, lo, nvalid);
      store_bf16_chunk(o2 + p.n_half + col0, hi, nvalid);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = (p.epi == TASR_EPI_SWIGLU) ? siluf_(lo[i]) * hi[i] : lo[i] * sigmoidf_(hi[i]);
        if (p.drop_thresh) v *= dropout_scale(p.seed, (unsigned long long)(r * p.N + col0 + i), p.drop_thresh, p.drop_inv_keep);
        lo[i] = v;
      }
      store_bf16_chunk(reinterpret_cast<bf16*>(p.out) + r * p.ldo + col0, lo, nvalid);
    } break;
    case TASR_EPI_SILU: {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float b = (p.bias != nullptr && i < nvalid) ? p.bias[col0 + i] : 0.f;
        lo[i] = bf16_round(lo[i] + b);
      }
      if (p.out2 != nullptr) store_bf16_chunk(reinterpret_cast<bf16*>(p.out2) + r * p.ldo2 + col0, lo, nvalid);
#pragma unroll
      for (int i = 0; i < 32; ++i) lo[i] = siluf_(lo[i]);
      store_bf16_chunk(reinterpret_cast<bf16*>(p.out) + r * p.ldo + col0, lo, nvalid);
    } break;
    case TASR_EPI_SWIGLU_BWD:
    case TASR_EPI_GLU_BWD: {
      float g[32], v[32];
      const bf16* ax = reinterpret_cast<const bf16*>(p.aux) + r * p.ldaux;
      load_bf16_chunk(ax + col0, g, nvalid);
      load_bf16_chunk(ax + p.n_half + col0, v, nvalid);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float d = lo[i];
        if (p.drop_thresh) d *= dropout_scale(p.seed, (unsigned long long)(r * p.N + col0 + i), p.drop_thresh, p.drop_inv_keep);
        if (p.epi == TASR_EPI_SWIGLU_BWD) {
          lo[i] = d * v[i] * silu_gradf_(g[i]);  // d/dg
          hi[i] = d * siluf_(g[i]);              // d/dv
        } else {
          float s = sigmoidf_(v[i]);
          lo[i] = d * s;                         // d/da
          hi[i] = d * g[i] * s * (1.f - s);      // d/db
        }
      }
      bf16* o = reinterpret_cast<bf16*>(p.out) + r * p.ldo;
      store_bf16_chunk(o + col0, lo, nvalid);
      store_bf16_chunk(o + p.n_half + col0, hi, nvalid);
    } break;
    case TASR_EPI_SILU_BWD: {
      float z[32];
      load_bf16_chunk(reinterpret_cast<const bf16*>(p.aux) + r * p.ldaux + col0, z, nvalid);
#pragma unroll
      for (int i = 0; i < 32; ++i) lo[i] = lo[i] * silu_gradf_(z[i]);
      store_bf16_chunk(reinterpret_cast<bf16*>(p.out) + r * p.ldo + col0, lo, nvalid);
    } break;
    case TASR_EPI_ATOMIC: {
      float* o = reinterpret_cast<float*>(p.out) + r * p.ldo;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i >= nvalid) break;
        int c = col0 + i;
        if (p.remap_p0 > 0) c = (c % p.remap_p0) * p.remap_p1 + c / p.remap_p0;
        atomicAdd(o + c, p.alpha * lo[i]);
      }
    } break;
    default: break;
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool A_MN, bool B_MN, bool DUAL>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  constexpr int A_BYTES = BM * BK * 2;  // 16 KB
  constexpr int B_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = BN;  // power of two >= 32

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int m0 = m_tile * BM;
  const int n0 = DUAL ? n_tile * (BN / 2) : n_tile * BN;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(num_kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (num_kb > 0) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        for (int i = 0; i < num_kb; ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
          const int k0 = (kb_begin + i) * BK;
          uint8_t* a_dst = sA + s * A_BYTES;
          uint8_t* b_dst = sB + s * B_BYTES;
#pragma unroll
          for (int j = 0; j < BM / 64; ++j) {
            if (A_MN) tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[s], m0 + 64 * j, k0);
            else      tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[s], k0, m0 + 64 * j);
          }
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            int nrow = DUAL ? (j < BN / 128 ? n0 + 64 * j : p.n_half + n0 + 64 * (j - BN / 128)) : n0 + 64 * j;
            if (B_MN) tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[s], nrow, k0);
            else      tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[s], k0, nrow);
          }
        }
      }
    } else if (warp == 1) {
      // ===================== UMMA issuer =====================
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
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
            umma_bf16(tmem_base, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);                    // frees the smem stage when the MMAs retire
          if (i == num_kb - 1) umma_commit(accum_bar);   // accumulator complete
        }
        __syncwarp();
      }
    } else {
      // ===================== epilogue =====================
      const int q = warp & 3;  // TMEM lane quarter this warp may access
      const int row = m0 + q * 32 + lane;
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      constexpr int NCH = DUAL ? BN / 64 : BN / 32;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        uint32_t lo_u[32], hi_u[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
        tmem_ld32(taddr, lo_u);
        if (DUAL) tmem_ld32(taddr + BN / 2, hi_u);
        tmem_ld_wait();
        if (row < p.M) {
          float* lo = reinterpret_cast<float*>(lo_u);
          float* hi = reinterpret_cast<float*>(hi_u);
          epilogue_chunk(p, row, n0 + c * 32, lo, hi);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// Debug CUDA-core kernel (tests / triage only): one thread per (row, 32-col chunk)
// ------------------------------------------------------------------------------------------------
__global__ void gemm_debug_kernel(const bf16* A, long long lda, int a_mn, const bf16* B, long long ldb, int b_mn,
                                  GemmDev p, int dual) {
  const int chunks = (p.N + 31) / 32;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.M * chunks) return;
  const int row = (int)(gid / chunks);
  const int col0 = (int)(gid % chunks) * 32;
  float lo[32], hi[32];
  for (int i = 0; i < 32; ++i) {
    float s0 = 0.f, s1 = 0.f;
    const int n = col0 + i;
    if (n < p.N) {
      for (int k = 0; k < p.K; ++k) {
        float a = __bfloat162float(a_mn ? A[(long long)k * lda + row] : A[(long long)row * lda + k]);
        float b = __bfloat162float(b_mn ? B[(long long)k * ldb + n] : B[(long long)n * ldb + k]);
        s0 += a * b;
        if (dual) {
          float b1 = __bfloat162float(b_mn ? B[(long long)k * ldb + p.n_half + n] : B[(long long)(p.n_half + n) * ldb + k]);
          s1 += a * b1;
        }
      }
    }
    lo[i] = s0;
    hi[i] = s1;
  }
  epilogue_chunk(p, row, col0, lo, hi);
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

}  // namespace

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows of pitch ld elements,
// box = 64 (inner) x box_outer, 128 B swizzle, zero fill out of bounds.
int tasr_make_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                           uint32_t box_inner, uint32_t box_outer) {
  std::call_once(g_encode_once, init_encode);
  if (!g_encode) return TASR_ERR_CUDA;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TASR_OK : TASR_ERR_CUDA;
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
  if (a->drop_p > 0.f) {
    double t = (double)a->drop_p * 4294967296.0;
    p->drop_thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    if (p->drop_thresh == 0) p->drop_thresh = 1;
    p->drop_inv_keep = 1.f / (1.f - a->drop_p);
  } else {
    p->drop_thresh = 0; p->drop_inv_keep = 1.f;
  }
  p->seed = a->seed;
  p->remap_p0 = a->remap_p0; p->remap_p1 = a->remap_p1;
  const int num_kb = (a->K + BK - 1) / BK;
  int splits = (a->epilogue == TASR_EPI_ATOMIC && a->split_k > 1) ? a->split_k : 1;
  if (splits > num_kb) splits = num_kb;
  p->kb_per_split = (num_kb + splits - 1) / splits;
  return TASR_OK;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, bool DUAL>
int launch_tc(const tasr_gemm_args* a, const GemmDev& p, cudaStream_t st) {
  constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 1) * 8 + 16 + 1024;
  CUtensorMap tmA, tmB;
  int rc;
  // A: K-major -> dims {K, M}; MN-major -> dims {M, K}
  if (A_MN) rc = tasr_make_tmap_2d_bf16(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda, 64, 64);
  else      rc = tasr_make_tmap_2d_bf16(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, 64, 64);
  if (rc) return rc;
  const uint64_t nrows = DUAL ? (uint64_t)2 * a->n_half : (uint64_t)a->N;
  if (B_MN) rc = tasr_make_tmap_2d_bf16(&tmB, a->B, nrows, (uint64_t)a->K, (uint64_t)a->ldb, 64, 64);
  else      rc = tasr_make_tmap_2d_bf16(&tmB, a->B, (uint64_t)a->K, nrows, (uint64_t)a->ldb, 64, 64);
  if (rc) return rc;
  auto kern = gemm_tc_kernel<BN, STAGES, A_MN, B_MN, DUAL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done = true;
  }
  const int tile_n = DUAL ? BN / 2 : BN;
  const int num_kb = (a->K + BK - 1) / BK;
  const int splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  dim3 grid(cdiv(a->N, tile_n), cdiv(a->M, BM), splits);
  kern<<<grid, GEMM_THREADS, SMEM, st>>>(tmA, tmB, p);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

}  // namespace

extern "C" int tasr_gemm_bf16(const tasr_gemm_args* a, tasr_stream_t stream) {
  GemmDev p;
  bool dual;
  int rc = fill_dev(a, &p, &dual);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dual) {
    if (a->a_mn_major || a->b_mn_major) return TASR_ERR_SHAPE;
    if (a->n_half % 64) return TASR_ERR_SHAPE;
    return launch_tc<128, 3, false, false, true>(a, p, st);
  }
  if (!a->a_mn_major && !a->b_mn_major) return launch_tc<128, 3, false, false, false>(a, p, st);
  if (!a->a_mn_major && a->b_mn_major) return launch_tc<128, 3, false, true, false>(a, p, st);
  if (a->a_mn_major && a->b_mn_major) return launch_tc<128, 3, true, true, false>(a, p, st);
  return launch_tc<128, 3, true, false, false>(a, p, st);
}

extern "C" int tasr_gemm_bf16_debug(const tasr_gemm_args* a, tasr_stream_t stream) {
  GemmDev p;
  bool dual;
  int rc = fill_dev(a, &p, &dual);
  if (rc) return rc;
  p.kb_per_split = (a->K + BK - 1) / BK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)a->M * ((a->N + 31) / 32);
  gemm_debug_kernel<<<cdiv(total, 128), 128, 0, st>>>(reinterpret_cast<const bf16*>(a->A), a->lda, a->a_mn_major,
                                                      reinterpret_cast<const bf16*>(a->B), a->ldb, a->b_mn_major, p,
                                                      dual ? 1 : 0);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
