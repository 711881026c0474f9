// Shared device code of the tcgen05 GEMM family (gemm.cu: plain 2-D operands; conv_gemm.cu: implicit-GEMM
// Conv2d with N-D TMA operands): parameter block, fused epilogue math, swizzled staging writes.
#pragma once
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 320;       // producer warp + MMA warp + 8 epilogue warps
constexpr int EPI_THREADS = 256;
constexpr int EPI_GROUP_THREADS = 128;  // one epilogue group = 4 warps = the 4 TMEM lane quarters
// gemm.cu: 16 epilogue warps = 2 super-groups of 8 warps (each TMEM lane quarter twice: one warp per 32-column half)
constexpr int G_THREADS = 576;
constexpr int G_EPI_THREADS = 512;
constexpr int G_SG_THREADS = 256;
constexpr int STAGE_BYTES = 16384;  // 128 rows x 128 B

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
  const unsigned long long* seed_ptr;
  int kb_per_split;
  int splits;
  int remap_p0, remap_p1;
  int tiles_m, tiles_n;
  float* colsum;
  int flags;  // experiment switches (TASR_GEMM_FLAGS in the environment); 0 in production
};

// W consecutive columns of one row (W = 16 or 32), vectorised when the chunk is full and 16-byte aligned
template <int W>
__device__ __forceinline__ void load_bf16_w(const bf16* src, float* v, int nvalid) {
  if (nvalid == W && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < W / 8; ++i) {
      uint4 u = s4[i];
      float2 f;
      f = unpack_bf16x2(u.x); v[8 * i + 0] = f.x; v[8 * i + 1] = f.y;
      f = unpack_bf16x2(u.y); v[8 * i + 2] = f.x; v[8 * i + 3] = f.y;
      f = unpack_bf16x2(u.z); v[8 * i + 4] = f.x; v[8 * i + 5] = f.y;
      f = unpack_bf16x2(u.w); v[8 * i + 6] = f.x; v[8 * i + 7] = f.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = (i < nvalid) ? __bfloat162float(src[i]) : 0.f;
  }
}
template <int W>
__device__ __forceinline__ void load_f32_w(const float* src, float* v, int nvalid) {
  if (nvalid == W && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
      float4 f = s4[i];
      v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = (i < nvalid) ? src[i] : 0.f;
  }
}
template <int W>
__device__ __forceinline__ void load_bias_w(const float* __restrict__ bias, int col0, int nvalid, float* b) {
  if (bias == nullptr) {
#pragma unroll
    for (int i = 0; i < W; ++i) b[i] = 0.f;
  } else if (nvalid == W && ((reinterpret_cast<uintptr_t>(bias + col0) & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
      const float4 f = __ldg(s4 + i);
      b[4 * i] = f.x; b[4 * i + 1] = f.y; b[4 * i + 2] = f.z; b[4 * i + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < W; ++i) b[i] = (i < nvalid) ? bias[col0 + i] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// Operand pre-loading for the epilogues that read a saved tensor (16-column pieces).  The loads are issued before
// the thread waits for the accumulator / the staging barrier so that their (HBM) latency overlaps with those waits.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gemm_drop_seed32(const GemmDev& p) {
  if (!p.drop_thresh) return 0u;
  return tasr_seed_mix(p.seed_ptr ? p.seed + *p.seed_ptr : p.seed);
}
struct AuxBf16 {  // 16 bf16 of the first half (g | a | z) and 16 of the second half (v | b)
  uint4 a[2], b[2];
};
template <int EPI>
__device__ __forceinline__ void preload_aux_bf16(const GemmDev& p, int row, int col0, AuxBf16& x) {
  x.a[0] = x.a[1] = x.b[0] = x.b[1] = make_uint4(0, 0, 0, 0);
  if (row < p.M && col0 + 16 <= p.N) {
    const bf16* ax = reinterpret_cast<const bf16*>(p.aux) + (long long)row * p.ldaux + col0;
    const uint4* s0 = reinterpret_cast<const uint4*>(ax);
    x.a[0] = s0[0]; x.a[1] = s0[1];
    if (EPI != TASR_EPI_SILU_BWD) {
      const uint4* s1 = reinterpret_cast<const uint4*>(ax + p.n_half);
      x.b[0] = s1[0]; x.b[1] = s1[1];
    }
  }
}
__device__ __forceinline__ float bf16_at(const uint4* q, int k) {  // element k (0..15) of 2 packed uint4
  const uint32_t w = reinterpret_cast<const uint32_t*>(q)[k >> 1];
  return __uint_as_float((k & 1) ? (w & 0xFFFF0000u) : (w << 16));
}
// math of the *_BWD epilogues on a pre-loaded piece: lo = accumulator in, lo/hi = outputs
template <int EPI>
__device__ __forceinline__ void epilogue_bwd16(const GemmDev& p, uint32_t s32, int row, int col0, float* lo, float* hi, const AuxBf16& x) {
  uint32_t dbase = 0;
  const uint32_t dseed_hi = 0;
  if (p.drop_thresh) dbase = tasr_hash_pair_base_s32(s32, (unsigned long long)((long long)row * p.N + col0) >> 1);
  if (EPI == TASR_EPI_SILU_BWD) {
#pragma unroll
    for (int i = 0; i < 16; ++i) lo[i] = lo[i] * silu_grad_tanh(bf16_at(x.a, i));
  }
#pragma unroll
  for (int i = 0; i < (EPI == TASR_EPI_SILU_BWD ? 0 : 16); i += 2) {
    float s0 = 1.f, s1 = 1.f;
    if (p.drop_thresh) dropout_scale2_fast(dbase, dseed_hi, i >> 1, p.drop_thresh, p.drop_inv_keep, s0, s1);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = i + j;
      const float d = lo[k] * (j == 0 ? s0 : s1);
      const float g = bf16_at(x.a, k), v = bf16_at(x.b, k);
      if (EPI == TASR_EPI_SWIGLU_BWD) {
        const float sg = sigmoid_tanh(g);
        const float dsg = d * sg;
        lo[k] = dsg * v * fmaf(g, 1.f - sg, 1.f);     // d/dg
        hi[k] = dsg * g;                               // d/dv
      } else {
        const float sv = sigmoid_tanh(v);
        const float dsv = d * sv;
        lo[k] = dsv;                                   // d/da
        hi[k] = dsv * g * (1.f - sv);                  // d/db
      }
    }
  }
}
struct AuxF32 {  // 16 fp32 residual values
  float4 v[4];
};
__device__ __forceinline__ void preload_aux_f32(const GemmDev& p, int row, int col0, AuxF32& x) {
  x.v[0] = x.v[1] = x.v[2] = x.v[3] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < p.M && col0 + 16 <= p.N) {
    const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + (long long)row * p.ldaux + col0);
    x.v[0] = s[0]; x.v[1] = s[1]; x.v[2] = s[2]; x.v[3] = s[3];
  }
}
__device__ __forceinline__ void epilogue_resid16(const GemmDev& p, uint32_t s32, int row, int col0, float* lo, const AuxF32& x) {
  const int nvalid = (row < p.M) ? max(0, min(16, p.N - col0)) : 0;
  float bb[16];
  load_bias_w<16>(p.bias, col0, nvalid, bb);
  uint32_t dbase = 0;
  const uint32_t dseed_hi = 0;
  if (p.drop_thresh) dbase = tasr_hash_pair_base_s32(s32, (unsigned long long)((long long)row * p.N + col0) >> 1);
  const float* res = reinterpret_cast<const float*>(x.v);
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    float v0 = lo[i] + bb[i], v1 = lo[i + 1] + bb[i + 1];
    if (p.drop_thresh) {
      float s0, s1;
      dropout_scale2_fast(dbase, dseed_hi, i >> 1, p.drop_thresh, p.drop_inv_keep, s0, s1);
      v0 *= s0; v1 *= s1;
    }
    lo[i] = res[i] + p.alpha * v0;
    lo[i + 1] = res[i + 1] + p.alpha * v1;
  }
}

// ------------------------------------------------------------------------------------------------
// Fused epilogue math on one row chunk of W columns [col0, col0+W) of output row `row`.
//   in : lo = accumulator;  hi = paired accumulator (dual-B modes)
//   out: lo = primary output, hi = second output, t3 = third output (see table below)
//     STORE / RESID / SILU_BWD / ATOMIC : lo -> out
//     SWIGLU / GLU                      : t3 -> out (h|u), lo -> out2[:, col] (g|a), hi -> out2[:, n_half+col]
//     SILU                              : t3 -> out (silu(z)), lo -> out2 (z)
//     SWIGLU_BWD / GLU_BWD              : lo -> out[:, col], hi -> out[:, n_half+col]
// Rows >= M or columns >= N produce don't-care values (clipped by the TMA store / masked by the caller).
// The mode is a template parameter: only that mode's code is generated (the step is epilogue-bound for the
// K = 256 GEMMs); W = 16 keeps the register footprint small enough for 16 epilogue warps per CTA.
// ------------------------------------------------------------------------------------------------
template <int EPI, int W>
__device__ __forceinline__ void epilogue_math(const GemmDev& p, uint32_t s32, int row, int col0, float* lo, float* hi, float* t3) {
  const int nvalid = (row < p.M) ? max(0, min(W, p.N - col0)) : 0;
  const long long r = row;
  // dropout: pair-hash over the run of W/2 pairs of this chunk (N % 32 == 0 is checked on the host, so a run never
  // crosses a 2^32 boundary of the pair index); s32 = tasr_seed_mix(seed), computed once per thread by the caller
  uint32_t dbase = 0;
  const uint32_t dseed_hi = 0;
  if (p.drop_thresh) dbase = tasr_hash_pair_base_s32(s32, (unsigned long long)(r * p.N + col0) >> 1);
  if (EPI == TASR_EPI_STORE) {
    load_bias_w<W>(p.bias, col0, nvalid, t3);
#pragma unroll
    for (int i = 0; i < W; ++i) lo[i] = p.alpha * (lo[i] + t3[i]);
  } else if (EPI == TASR_EPI_RESID) {
    load_bias_w<W>(p.bias, col0, nvalid, t3);
#pragma unroll
    for (int i = 0; i < W; ++i) lo[i] += t3[i];
    load_f32_w<W>(reinterpret_cast<const float*>(p.aux) + r * p.ldaux + col0, t3, nvalid);
#pragma unroll
    for (int i = 0; i < W; i += 2) {
      float v0 = lo[i], v1 = lo[i + 1];
      if (p.drop_thresh) {
        float s0, s1;
        dropout_scale2_fast(dbase, dseed_hi, i >> 1, p.drop_thresh, p.drop_inv_keep, s0, s1);
        v0 *= s0; v1 *= s1;
      }
      lo[i] = t3[i] + p.alpha * v0;
      lo[i + 1] = t3[i + 1] + p.alpha * v1;
    }
  } else if (EPI == TASR_EPI_SWIGLU || EPI == TASR_EPI_GLU) {
    load_bias_w<W>(p.bias, col0, nvalid, t3);
#pragma unroll
    for (int i = 0; i < W; ++i) lo[i] += t3[i];
    load_bias_w<W>(p.bias ? p.bias + p.n_half : nullptr, col0, nvalid, t3);
#pragma unroll
    for (int i = 0; i < W; ++i) hi[i] += t3[i];
#pragma unroll
    for (int i = 0; i < W; i += 2) {
      float s0 = 1.f, s1 = 1.f;
      if (p.drop_thresh) dropout_scale2_fast(dbase, dseed_hi, i >> 1, p.drop_thresh, p.drop_inv_keep, s0, s1);
      const float v0 = (EPI == TASR_EPI_SWIGLU) ? silu_tanh(lo[i]) * hi[i] : lo[i] * sigmoid_tanh(hi[i]);
      const float v1 = (EPI == TASR_EPI_SWIGLU) ? silu_tanh(lo[i + 1]) * hi[i + 1] : lo[i + 1] * sigmoid_tanh(hi[i + 1]);
      t3[i] = v0 * s0;
      t3[i + 1] = v1 * s1;
    }
  } else if (EPI == TASR_EPI_SILU) {
    load_bias_w<W>(p.bias, col0, nvalid, t3);
#pragma unroll
    for (int i = 0; i < W; ++i) {
      lo[i] += t3[i];
      t3[i] = silu_tanh(lo[i]);
    }
  } else if (EPI == TASR_EPI_SWIGLU_BWD || EPI == TASR_EPI_GLU_BWD) {
    const bf16* ax = reinterpret_cast<const bf16*>(p.aux) + r * p.ldaux;
    load_bf16_w<W>(ax + col0, hi, nvalid);             // g | a
    load_bf16_w<W>(ax + p.n_half + col0, t3, nvalid);  // v | b
#pragma unroll
    for (int i = 0; i < W; i += 2) {
      float s0 = 1.f, s1 = 1.f;
      if (p.drop_thresh) dropout_scale2_fast(dbase, dseed_hi, i >> 1, p.drop_thresh, p.drop_inv_keep, s0, s1);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = i + j;
        const float d = lo[k] * (j == 0 ? s0 : s1);
        const float g = hi[k], v = t3[k];
        if (EPI == TASR_EPI_SWIGLU_BWD) {
          const float sg = sigmoid_tanh(g);
          const float dsg = d * sg;
          lo[k] = dsg * v * fmaf(g, 1.f - sg, 1.f);     // d/dg
          hi[k] = dsg * g;                               // d/dv
        } else {
          const float sv = sigmoid_tanh(v);
          const float dsv = d * sv;
          lo[k] = dsv;                                   // d/da
          hi[k] = dsv * g * (1.f - sv);                  // d/db
        }
      }
    }
  } else if (EPI == TASR_EPI_SILU_BWD) {
    load_bf16_w<W>(reinterpret_cast<const bf16*>(p.aux) + r * p.ldaux + col0, t3, nvalid);
#pragma unroll
    for (int i = 0; i < W; ++i) lo[i] = lo[i] * silu_grad_tanh(t3[i]);
  } else if (EPI == TASR_EPI_ATOMIC) {
#pragma unroll
    for (int i = 0; i < W; ++i) lo[i] *= p.alpha;
  }
}

// Rotary position embedding on a 16-column piece of a 64-wide head (model/attention.py:62-70): with x1 = first half
// and x2 = second half of the head, out1 = x1 cos - x2 sin, out2 = x2 cos + x1 sin.  `lo` = this thread's 16 columns
// (accumulator), `pr` = the 16 partner columns 32 away in the same head (own TMEM lane, read by the caller);
// hsel = which half `lo` is in; sub = which 16 of the 32 rotary indices.  Position = row % n_half (absolute frame index).
__device__ __forceinline__ void epilogue_rope16(const GemmDev& p, int row, int col0, int hsel, int sub, bool rot, float* lo,
                                                const float* pr) {
  const int nvalid = (row < p.M) ? max(0, min(16, p.N - col0)) : 0;
  float bb[16];
  load_bias_w<16>(p.bias, col0, nvalid, bb);
#pragma unroll
  for (int i = 0; i < 16; ++i) lo[i] += bb[i];
  if (!rot) return;
  load_bias_w<16>(p.bias, col0 + (hsel ? -32 : 32), 16, bb);  // rotated heads are always complete (remap_p0 % 64 == 0)
  const int t = row % p.n_half;
  const float4* cs = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + ((long long)t * 32 + sub * 16) * 2);
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float4 v = __ldg(cs + (i >> 1));  // (cos_i, sin_i, cos_i+1, sin_i+1)
    const float p0 = pr[i] + bb[i], p1 = pr[i + 1] + bb[i + 1];
    lo[i] = hsel ? fmaf(p0, v.y, lo[i] * v.x) : fmaf(-p0, v.y, lo[i] * v.x);
    lo[i + 1] = hsel ? fmaf(p1, v.w, lo[i + 1] * v.z) : fmaf(-p1, v.w, lo[i + 1] * v.z);
  }
}

// 16 columns of row r into a [128 rows x 128 B] swizzled staging buffer, starting at 16-byte chunk `chunk0`
__device__ __forceinline__ void stage_bf16_16(uint8_t* buf, int r, int chunk0, const float* v) {
  uint8_t* base = buf + r * 128;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(base + (((chunk0 + i) ^ (r & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void stage_f32_16(uint8_t* buf, int r, int chunk0, const float* v) {
  uint8_t* base = buf + r * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(base + (((chunk0 + i) ^ (r & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// staging writes: row r of a [128 rows x 128 B] buffer in the TMA 128-byte swizzle
__device__ __forceinline__ void stage_bf16_half(uint8_t* buf, int r, int half, const float* v) {
  uint8_t* base = buf + r * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(base + (((half * 4 + i) ^ (r & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void stage_f32(uint8_t* buf, int r, const float* v) {
  uint8_t* base = buf + r * 128;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(base + ((i ^ (r & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}


}  // namespace
