// Multi-query flash attention on tcgen05 / TMEM, forward and backward.
// One shared K/V head (64 dims) for H query heads, key-padding mask by utterance length, online
// softmax, optional dropout on the probabilities (counter-based mask, recomputed in backward).
// Replaces (reference): model/attention.py:233-245 (expand K/V to H heads + F.scaled_dot_product_attention
//   with an additive -inf key mask in training, or the materialised (B,H,T,T) _standard_attention in
//   eval) and the SDPA backward.
//
// Tensors (bf16, token-major, RoPE already applied):  qkv (B*T, d + 128): q heads | k | v;
// ctx (B*T, d).  S = Q K^T and O_blk = P V run as UMMA (A/B from shared memory, D in TMEM); one thread
// owns one query row (TMEM lane) for the softmax; P is written back to shared memory in the 128-byte
// swizzled K-major layout so that it can be the A operand of the next UMMA (and, in backward, be
// re-read MN-major as P^T without a transpose).
#include "common.cuh"

namespace {

constexpr int DH = 64;
constexpr int BQ = 128;   // query rows per tile
constexpr int BKV = 128;  // keys per block
constexpr int ATT_THREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;

struct AttnParams {
  int B, T, H, d;
  const long long* key_len;  // (B) valid keys per utterance, or nullptr (all T keys valid)
  float scale;               // 1/sqrt(dh)
  uint32_t drop_thresh;
  float drop_inv_keep;
  unsigned long long seed;
  const unsigned long long* seed_ptr;
  bf16* ctx;                 // fwd out (B*T, d)
  float* lse2;               // (B, H, T) log2-domain logsumexp
  // backward
  const float* delta;        // (B, H, T) rowsum(dO * O)
  float* dq_acc;             // (B*T, d) fp32, zero-initialised
  bf16* dqkv;                // (B*T, d + 128): dk | dv written at cols d.., dq left to the finalize kernel
};

// write 32 consecutive bf16 (row r, columns [c0, c0+32)) of a [128 x 128] tile stored as two K-major
// 128 B-swizzled sub-tiles of 64 columns (16 KB each)
__device__ __forceinline__ void store_tile_chunk(uint8_t* tile, int r, int c0, const float* v) {
  uint8_t* base = tile + (c0 >> 6) * 16384 + r * 128;
  const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(base + (((chunk0 + i) ^ (r & 7)) << 4)) = u;
  }
}

// dropout index space: element (b, h, q, key) -> ((b H + h) T + q) * even(T) + key, so (key, key + 1) pairs share a hash
// Dropout mask of one (key, key + 1) pair inside a 32-key chunk: the chunk's affine hash part is computed once
// (tasr_hash_pair_base_s32 of the chunk's first pair), pair j of the chunk costs one add + the mixer.  Forward and
// backward walk the same 32-key chunks, so they see the same mask.  thresh_hi = thresh16 << 16.
__device__ __forceinline__ void attn_drop_pair(uint32_t base32, uint32_t j, uint32_t thresh_hi, bool& keep0, bool& keep1) {
  dropout_keep2_fast(base32, j, thresh_hi, keep0, keep1);
}

// ------------------------------------------------------------------------------------------------
// forward: grid (ceil(T/128), H, B), 256 threads.  Two threads share a query row (= TMEM lane): warps 0-3 own key
// columns [0,64) of every 128-key S tile and columns [0,32) of O, warps 4-7 the other halves; the block row-max and
// the final row sum are exchanged through shared memory.  K(j+1) is fetched while block j is in the softmax, V(j+1)
// while S(j+1) is computed.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 2) mqa_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;            // 16 KB
  uint8_t* sK = smem + 16384;    // 16 KB
  uint8_t* sV = smem + 32768;    // 16 KB
  uint8_t* sP = smem + 49152;    // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 81920);  // q, k, v, s, o
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* xch = reinterpret_cast<float*>(smem + 81920 + 64);    // [2][128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int rloc = quarter * 32 + lane;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int Lk = p.key_len ? (int)max(0LL, min((long long)p.T, p.key_len[b])) : p.T;
  const int nkv = (Lk + BKV - 1) / BKV;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 128;

  auto load_k = [&](int j) {
    mbar_expect_tx(&bars[1], 16384);
    tma_load_3d(sK, &tm_qkv, &bars[1], p.d, j * BKV, b);
    tma_load_3d(sK + 8192, &tm_qkv, &bars[1], p.d, j * BKV + 64, b);
  };
  auto load_v = [&](int j) {
    mbar_expect_tx(&bars[2], 16384);
    tma_load_3d(sV, &tm_qkv, &bars[2], p.d + DH, j * BKV, b);
    tma_load_3d(sV + 8192, &tm_qkv, &bars[2], p.d + DH, j * BKV + 64, b);
  };
  if (tid == 0 && nkv > 0) {
    mbar_expect_tx(&bars[0], 16384);
    tma_load_3d(sQ, &tm_qkv, &bars[0], h * DH, q0, b);
    tma_load_3d(sQ + 8192, &tm_qkv, &bars[0], h * DH, q0 + 64, b);
    load_k(0);
    load_v(0);
  }
  const unsigned long long dseed = (p.drop_thresh && p.seed_ptr) ? p.seed + *p.seed_ptr : p.seed;
  const uint32_t dseed32 = tasr_seed_mix(dseed), thresh_hi = p.drop_thresh << 16;
  const float scale2 = p.scale * LOG2E;
  float m_run = -INFINITY, l_part = 0.f;
  float o[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) o[i] = 0.f;
  const int qrow = q0 + rloc;
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, BKV, 0, 0);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);
  const int Tp = (p.T + 1) & ~1;  // even row pitch of the dropout index space
  const unsigned long long drow = ((unsigned long long)(b * p.H + h) * p.T + qrow) * (unsigned long long)Tp;

  for (int j = 0; j < nkv; ++j) {
    const uint32_t ph = (uint32_t)j & 1u;
    if (tid == 0) {
      if (j == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1], ph);
      tc_fence_after();
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        umma_bf16(tS, umma_desc_sw128(qa + k * 32, 16, 1024), umma_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0);
      umma_commit(&bars[3]);
    }
    mbar_wait(&bars[3], ph);
    __syncwarp();
    tc_fence_after();
    if (tid == 0 && j + 1 < nkv) load_k(j + 1);  // S(j) is complete: the K tile is free
    const int nvalid = Lk - j * BKV;             // keys of this block that exist (>= 1)
    const bool tail = nvalid < BKV;
    // pass 1: row max over this thread's 64 columns
    float mx = -INFINITY;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;
      uint32_t u[32];
      tmem_ld32(tS + lane_addr + c * 32, u);
      tmem_ld_wait();
      if (!tail) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(u[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(u[i]));
      }
    }
    xch[half * 128 + rloc] = mx;
    __syncthreads();
    mx = fmaxf(mx, xch[(half ^ 1) * 128 + rloc]) * scale2;  // scale2 > 0; at least one valid key per block
    const float m_new = fmaxf(m_run, mx);
    const float corr = fast_exp2(m_run - m_new);  // 0 on the first block (m_run = -inf)
    float lsum = 0.f;
    // pass 2: probabilities -> shared memory (bf16, A operand of the PV product)
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;
      uint32_t u[32];
      tmem_ld32(tS + lane_addr + c * 32, u);
      tmem_ld_wait();
      float pv[32];
      const uint32_t dbase = tasr_hash_pair_base_s32(dseed32, (drow + (unsigned long long)(j * BKV + c * 32)) >> 1);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float e0 = fast_exp2(fmaf(__uint_as_float(u[i]), scale2, -m_new));
        float e1 = fast_exp2(fmaf(__uint_as_float(u[i + 1]), scale2, -m_new));
        if (tail) {
          if (c * 32 + i >= nvalid) e0 = 0.f;
          if (c * 32 + i + 1 >= nvalid) e1 = 0.f;
        }
        lsum += e0 + e1;
        if (p.drop_thresh) {
          bool k0, k1;
          attn_drop_pair(dbase, (uint32_t)(i >> 1), thresh_hi, k0, k1);
          e0 = k0 ? e0 * p.drop_inv_keep : 0.f;
          e1 = k1 ? e1 * p.drop_inv_keep : 0.f;
        }
        pv[i] = e0;
        pv[i + 1] = e1;
      }
      store_tile_chunk(sP, rloc, c * 32, pv);
    }
    l_part = l_part * corr + lsum;
    m_run = m_new;
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] *= corr;
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&bars[2], ph);
      tc_fence_after();
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k)
        umma_bf16(tO, umma_desc_sw128(pa + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                  umma_desc_sw128(va + k * 2048, 8192, 1024), idesc_o, k > 0);
      umma_commit(&bars[4]);
    }
    mbar_wait(&bars[4], ph);
    __syncwarp();
    tc_fence_after();
    if (tid == 0 && j + 1 < nkv) load_v(j + 1);  // PV(j) is complete: the V tile (and P) are free
    {
      uint32_t u[32];
      tmem_ld32(tO + lane_addr + half * 32, u);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(u[i]);
    }
    tc_fence_before();  // the next S / PV products are issued only after a later __syncthreads
  }
  xch[half * 128 + rloc] = l_part;
  __syncthreads();
  const float l_run = l_part + xch[(half ^ 1) * 128 + rloc];
  if (qrow < p.T) {
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    bf16* dst = p.ctx + ((long long)b * p.T + qrow) * p.d + h * DH + half * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
      u.y = pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
      u.z = pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
      u.w = pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
      reinterpret_cast<uint4*>(dst)[i] = u;
    }
    if (p.lse2 && half == 0) p.lse2[((long long)b * p.H + h) * p.T + qrow] = l_run > 0.f ? m_run + log2f(l_run) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// backward prep: delta[b,h,t] = sum_j dO[row, 64h+j] * O[row, 64h+j]   (one warp per (row, head))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ dctx, const bf16* __restrict__ ctx, int B, int T,
                                                         int H, int d, float* __restrict__ delta) {
  // 8 threads per (row, head), 8 elements (16 B) each
  const long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long item = gidx >> 3;
  const int sub = (int)(gidx & 7);
  const bool ok = item < (long long)B * T * H;
  float s = 0.f;
  long long row = 0;
  int h = 0;
  if (ok) {
    row = item / H;
    h = (int)(item - row * H);
    const long long off = row * d + h * DH + sub * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(dctx + off);
    const uint4 c = *reinterpret_cast<const uint4*>(ctx + off);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16x2(av[i]), y = unpack_bf16x2(cv[i]);
      s = fmaf(x.x, y.x, fmaf(x.y, y.y, s));
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && sub == 0) {
    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
    delta[((long long)b * H + h) * T + t] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// backward: persistent, one CTA per SM.  Work unit = (utterance b, 128-key block, head, 128-query tile); the units of
// the key blocks that exist (key-length aware) are dealt to the CTAs as equal contiguous ranges, so a CTA keeps K/V
// and the dK/dV accumulators (TMEM) while it stays on one key block and adds them to the fp32 gradient buffer when it
// leaves it.  512 threads: warp w owns TMEM lanes (= query rows) 32 (w % 4) .. + 31 and the 32-column chunk w / 4 of the
// S / dP tiles (four threads per row).  Q / dO tiles are double-buffered, and the S / dP products of unit i+1 are issued right
// behind the dV / dK / dQ products of unit i, so they run while unit i's dQ is drained.
// ------------------------------------------------------------------------------------------------
constexpr int ATT_BWD_THREADS = 512;
constexpr int ATT_BWD_MAX_B = 4096;

struct BwdUnit {
  int b, kblk, h, qi;
  long long g;  // key-block index over the whole batch (b, kblk)
};

__device__ __forceinline__ BwdUnit decode_unit(long long u, int per, int nq, const int* __restrict__ pre, int& bhint) {
  BwdUnit r;
  r.g = u / per;
  const int rem = (int)(u - r.g * per);
  r.h = rem / nq;
  r.qi = rem - r.h * nq;
  while (pre[bhint + 1] <= r.g) ++bhint;
  r.b = bhint;
  r.kblk = (int)(r.g - pre[bhint]);
  return r;
}

__global__ void __launch_bounds__(ATT_BWD_THREADS) mqa_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                  const __grid_constant__ CUtensorMap tm_do, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;             // 16 KB
  uint8_t* sV = smem + 16384;     // 16 KB
  uint8_t* sQ = smem + 32768;     // 2 x 16 KB
  uint8_t* sDO = smem + 65536;    // 2 x 16 KB
  uint8_t* sP = smem + 98304;     // 32 KB
  uint8_t* sDS = smem + 131072;   // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 163840);  // kv, q0, q1, mma1, mma2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  int* pre = reinterpret_cast<int*>(smem + 163840 + 64);        // [B + 1] prefix of key blocks per utterance

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, cgrp = warp >> 2;
  const int rloc = quarter * 32 + lane;  // TMEM lane = row of the tile
  const int ld = p.d + 2 * DH;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < p.B; i += ATT_BWD_THREADS) {
    const int Lk = p.key_len ? (int)max(0LL, min((long long)p.T, p.key_len[i])) : p.T;
    pre[i + 1] = (Lk + BKV - 1) / BKV;
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    pre[0] = 0;
    for (int i = 1; i <= p.B; ++i) pre[i] += pre[i - 1];
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + 128, tDK = tmem + 256, tDV = tmem + 320, tDQ = tmem + 384;
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;

  const int nq = (p.T + BQ - 1) / BQ;
  const int per = p.H * nq;
  const long long U = (long long)pre[p.B] * per;
  const long long u_begin = U * blockIdx.x / gridDim.x, u_end = U * (blockIdx.x + 1) / gridDim.x;
  const int n = (int)(u_end - u_begin);

  const unsigned long long dseed = (p.drop_thresh && p.seed_ptr) ? p.seed + *p.seed_ptr : p.seed;
  const uint32_t dseed32 = tasr_seed_mix(dseed), thresh_hi = p.drop_thresh << 16;
  const float scale2 = p.scale * LOG2E;
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, BKV, 0, 0);     // Q K^T, dO V^T
  constexpr uint32_t idesc_t = umma_idesc_bf16(128, DH, 1, 1);      // P^T dO, dS^T Q
  constexpr uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);      // dS K
  const int Tp = (p.T + 1) & ~1;  // even row pitch of the dropout index space

  // ---- issuer-side helpers (thread 0 only) ----
  uint32_t kv_phase = 0;
  auto issue_qdo = [&](const BwdUnit& un, int buf) {
    uint64_t* bar = &bars[1 + buf];
    const int q0 = un.qi * BQ;
    mbar_expect_tx(bar, 32768);
    tma_load_3d(sQ + buf * 16384, &tm_qkv, bar, un.h * DH, q0, un.b);
    tma_load_3d(sQ + buf * 16384 + 8192, &tm_qkv, bar, un.h * DH, q0 + 64, un.b);
    tma_load_3d(sDO + buf * 16384, &tm_do, bar, un.h * DH, q0, un.b);
    tma_load_3d(sDO + buf * 16384 + 8192, &tm_do, bar, un.h * DH, q0 + 64, un.b);
  };
  auto load_kv = [&](const BwdUnit& un) {
    const int k0 = un.kblk * BKV;
    mbar_expect_tx(&bars[0], 32768);
    tma_load_3d(sK, &tm_qkv, &bars[0], p.d, k0, un.b);
    tma_load_3d(sK + 8192, &tm_qkv, &bars[0], p.d, k0 + 64, un.b);
    tma_load_3d(sV, &tm_qkv, &bars[0], p.d + DH, k0, un.b);
    tma_load_3d(sV + 8192, &tm_qkv, &bars[0], p.d + DH, k0 + 64, un.b);
    mbar_wait(&bars[0], kv_phase);
    kv_phase ^= 1u;
  };
  auto issue_s_dp = [&](int i) {  // S = Q K^T and dP = dO V^T of local unit i
    const int buf = i & 1;
    mbar_wait(&bars[1 + buf], (uint32_t)(i >> 1) & 1u);
    tc_fence_after();
    const uint32_t qa = smem_u32(sQ + buf * 16384), ka = smem_u32(sK), da = smem_u32(sDO + buf * 16384), va = smem_u32(sV);
#pragma unroll
    for (int k = 0; k < DH / 16; ++k)
      umma_bf16(tS, umma_desc_sw128(qa + k * 32, 16, 1024), umma_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
    for (int k = 0; k < DH / 16; ++k)
      umma_bf16(tDP, umma_desc_sw128(da + k * 32, 16, 1024), umma_desc_sw128(va + k * 32, 16, 1024), idesc_s, k > 0);
    umma_commit(&bars[3]);
  };

  if (n > 0) {
    int bhint = 0;
    BwdUnit cur = decode_unit(u_begin, per, nq, pre, bhint);
    if (tid == 0) {
      issue_qdo(cur, 0);
      if (n > 1) {
        int bh2 = bhint;
        issue_qdo(decode_unit(u_begin + 1, per, nq, pre, bh2), 1);
      }
      load_kv(cur);
      issue_s_dp(0);
    }
    bool first_in_group = true;
    for (int i = 0; i < n; ++i) {
      const int buf = i & 1;
      const uint32_t ph = (uint32_t)i & 1u;
      int bh1 = bhint;
      BwdUnit nxt = cur;
      const bool has_next = i + 1 < n;
      if (has_next) nxt = decode_unit(u_begin + i + 1, per, nq, pre, bh1);
      const bool next_same = has_next && nxt.g == cur.g;
      const int b = cur.b, h = cur.h, k0 = cur.kblk * BKV, q0 = cur.qi * BQ;
      const int Lk = p.key_len ? (int)max(0LL, min((long long)p.T, p.key_len[b])) : p.T;
      uint8_t* cQ = sQ + buf * 16384;
      uint8_t* cDO = sDO + buf * 16384;
      const int qrow = q0 + rloc;
      const bool qvalid = qrow < p.T;
      float lse2 = 0.f, delta = 0.f;
      if (qvalid) {
        lse2 = p.lse2[((long long)b * p.H + h) * p.T + qrow];
        delta = p.delta[((long long)b * p.H + h) * p.T + qrow];
      }
      // rows that do not exist (or saw no key): lse2 = +inf makes every probability exp2(-inf) = 0
      if (!qvalid || lse2 == -INFINITY) lse2 = INFINITY;
      const unsigned long long drow = ((unsigned long long)(b * p.H + h) * p.T + qrow) * (unsigned long long)Tp;
      mbar_wait(&bars[3], ph);
      __syncwarp();
      tc_fence_after();
      {
        const int c = cgrp;  // 32-column chunk of the 128-key tile
        uint32_t us[32], ud[32];
        tmem_ld32(tS + lane_addr + c * 32, us);
        tmem_ld32(tDP + lane_addr + c * 32, ud);
        tmem_ld_wait();
        float pv[32], dsv[32];
        const uint32_t dbase = tasr_hash_pair_base_s32(dseed32, (drow + (unsigned long long)(k0 + c * 32)) >> 1);
        const int nvalid = Lk - k0 - c * 32;  // keys of this chunk that exist
        const float inv_keep = p.drop_thresh ? p.drop_inv_keep : 1.f;
#pragma unroll
        for (int i2 = 0; i2 < 32; i2 += 2) {
          bool keep0 = true, keep1 = true;
          if (p.drop_thresh) attn_drop_pair(dbase, (uint32_t)(i2 >> 1), thresh_hi, keep0, keep1);
          float pr0 = fast_exp2(fmaf(__uint_as_float(us[i2]), scale2, -lse2));
          float pr1 = fast_exp2(fmaf(__uint_as_float(us[i2 + 1]), scale2, -lse2));
          if (nvalid < 32) {  // only in the last key block of an utterance
            if (i2 >= nvalid) pr0 = 0.f;
            if (i2 + 1 >= nvalid) pr1 = 0.f;
          }
          const float m0 = keep0 ? inv_keep : 0.f, m1 = keep1 ? inv_keep : 0.f;
          dsv[i2] = (pr0 * p.scale) * fmaf(__uint_as_float(ud[i2]), m0, -delta);
          dsv[i2 + 1] = (pr1 * p.scale) * fmaf(__uint_as_float(ud[i2 + 1]), m1, -delta);
          pv[i2] = pr0 * m0;
          pv[i2 + 1] = pr1 * m1;
        }
        store_tile_chunk(sP, rloc, c * 32, pv);
        store_tile_chunk(sDS, rloc, c * 32, dsv);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t pa = smem_u32(sP), sa = smem_u32(sDS), qa = smem_u32(cQ), da = smem_u32(cDO), ka = smem_u32(sK);
        const uint32_t acc0 = first_in_group ? 0u : 1u;
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k)  // dV += P^T dO   (reduction over query rows)
          umma_bf16(tDV, umma_desc_sw128(pa + k * 2048, 16384, 1024), umma_desc_sw128(da + k * 2048, 8192, 1024), idesc_t,
                    k > 0 ? 1u : acc0);
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k)  // dK += dS^T Q
          umma_bf16(tDK, umma_desc_sw128(sa + k * 2048, 16384, 1024), umma_desc_sw128(qa + k * 2048, 8192, 1024), idesc_t,
                    k > 0 ? 1u : acc0);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)  // dQ = dS K    (reduction over keys)
          umma_bf16(tDQ, umma_desc_sw128(sa + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    umma_desc_sw128(ka + k * 2048, 8192, 1024), idesc_q, k > 0);
        umma_commit(&bars[4]);
        if (next_same) issue_s_dp(i + 1);  // same K/V: S / dP of the next unit run while this unit's dQ is drained
      }
      mbar_wait(&bars[4], ph);
      __syncwarp();
      tc_fence_after();
      if (tid == 0 && i + 2 < n) {  // this unit's Q / dO buffer is free again
        int bh2 = bh1;
        issue_qdo(decode_unit(u_begin + i + 2, per, nq, pre, bh2), buf);
      }
      {
        uint32_t u[16];
        tmem_ld16(tDQ + lane_addr + cgrp * 16, u);
        tmem_ld_wait();
        if (qvalid) {
          float4* dst = reinterpret_cast<float4*>(p.dq_acc + ((long long)b * p.T + qrow) * ld + h * DH + cgrp * 16);
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2)
            atomicAdd(dst + i2, make_float4(__uint_as_float(u[4 * i2]), __uint_as_float(u[4 * i2 + 1]),
                                            __uint_as_float(u[4 * i2 + 2]), __uint_as_float(u[4 * i2 + 3])));
        }
      }
      first_in_group = false;
      if (!next_same) {
        // leaving this key block: add dK (half 0) / dV (half 1), summed over the heads and query tiles seen, to the buffer
        const int krow = k0 + rloc;
        {
          uint32_t u[32];
          tmem_ld32(tDK + cgrp * 32 + lane_addr, u);  // tDK | tDV are adjacent: chunks 0,1 = dK, 2,3 = dV
          tmem_ld_wait();
          if (krow < Lk) {
            float4* dst = reinterpret_cast<float4*>(p.dq_acc + ((long long)b * p.T + krow) * ld + p.d + cgrp * 32);
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2)
              atomicAdd(dst + i2, make_float4(__uint_as_float(u[4 * i2]), __uint_as_float(u[4 * i2 + 1]),
                                              __uint_as_float(u[4 * i2 + 2]), __uint_as_float(u[4 * i2 + 3])));
          }
        }
        tc_fence_before();
        __syncthreads();  // everyone has drained dK / dV before the next key block overwrites them (and K / V)
        if (has_next && tid == 0) {
          tc_fence_after();
          load_kv(nxt);
          issue_s_dp(i + 1);
        }
        first_in_group = true;
      }
      tc_fence_before();
      cur = nxt;
      bhint = bh1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// gradient buffer (M, d + 128) fp32 -> dqkv bf16: inverse RoPE on the q heads and on k, plain conversion of v.
// A thread converts elements i..i+3 and i+32..i+35 of one head (i multiple of 4).
__global__ void __launch_bounds__(256) attn_dq_finalize_kernel(const float* __restrict__ acc, bf16* __restrict__ dqkv,
                                                               long long M, int T, int d, const float* __restrict__ cs) {
  const int ld = d + 2 * DH;
  const int quads = ld >> 3;  // (head, i4) with i4 < 8
  const long long total = M * quads;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const long long row = idx / quads;
    const int qidx = (int)(idx - row * quads);
    const int head = qidx >> 3, i = (qidx & 7) << 2;
    const float* src = acc + row * ld + head * DH;
    const float4 x1 = *reinterpret_cast<const float4*>(src + i), x2 = *reinterpret_cast<const float4*>(src + i + 32);
    float4 c0 = make_float4(1.f, 0.f, 1.f, 0.f), c1 = c0;  // (cos, sin) pairs
    if (cs != nullptr && head * DH < d + DH) {  // rotated heads: q heads and k
      const int t = (int)(row % T);
      c0 = *reinterpret_cast<const float4*>(cs + (t * 32 + i) * 2);
      c1 = *reinterpret_cast<const float4*>(cs + (t * 32 + i) * 2 + 4);
    }
    bf16* o = dqkv + row * ld + head * DH;
    uint2 lo, hi;
    lo.x = pack_bf16x2(x1.x * c0.x + x2.x * c0.y, x1.y * c0.z + x2.y * c0.w);
    lo.y = pack_bf16x2(x1.z * c1.x + x2.z * c1.y, x1.w * c1.z + x2.w * c1.w);
    hi.x = pack_bf16x2(x2.x * c0.x - x1.x * c0.y, x2.y * c0.z - x1.y * c0.w);
    hi.y = pack_bf16x2(x2.z * c1.x - x1.z * c1.y, x2.w * c1.z - x1.w * c1.w);
    *reinterpret_cast<uint2*>(o + i) = lo;
    *reinterpret_cast<uint2*>(o + i + 32) = hi;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_3d(CUtensorMap* m, const void* base, int cols, int T, int B, long long ld) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  static PFN_encodeTiled enc = nullptr;
  if (!enc) {
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return TASR_ERR_CUDA;
    enc = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)T * ld * 2};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? TASR_OK : TASR_ERR_CUDA;
}

void fill_params(AttnParams* p, int B, int T, int H, int d, const int64_t* key_len, float drop_p, uint64_t seed) {
  p->B = B; p->T = T; p->H = H; p->d = d;
  p->key_len = reinterpret_cast<const long long*>(key_len);
  p->scale = 0.125f;
  p->drop_thresh = tasr_drop_thresh16(drop_p);
  p->drop_inv_keep = tasr_drop_inv_keep(p->drop_thresh);
  p->seed = seed;
  p->seed_ptr = g_tasr_seed_ptr;
  p->ctx = nullptr; p->lse2 = nullptr; p->delta = nullptr; p->dq_acc = nullptr; p->dqkv = nullptr;
}

}  // namespace

extern "C" int tasr_mqa_attention_fwd(const void* qkv, int B, int T, int H, int d, const int64_t* key_lengths,
                                      float drop_p, uint64_t seed, void* ctx, float* lse2, tasr_stream_t stream) {
  if (B <= 0 || T <= 0 || H <= 0 || d != H * DH) return TASR_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(qkv) & 15) return TASR_ERR_ALIGN;
  CUtensorMap tm;
  int rc = make_tmap_3d(&tm, qkv, d + 2 * DH, T, B, d + 2 * DH);
  if (rc) return rc;
  AttnParams p;
  fill_params(&p, B, T, H, d, key_lengths, drop_p, seed);
  p.ctx = reinterpret_cast<bf16*>(ctx);
  p.lse2 = lse2;
  constexpr int SMEM = 81920 + 64 + 1024 + 1024;  // tiles, barriers, row exchange, alignment slack
  static TasrPerDevice attr_done;
  if (!attr_done.get()) {
    cudaError_t e = cudaFuncSetAttribute(mqa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done.set();
  }
  dim3 grid(cdiv(T, BQ), H, B);
  mqa_fwd_kernel<<<grid, ATT_THREADS, SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(tm, p);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" size_t tasr_mqa_attention_bwd_workspace_bytes(int B, int T, int H, int d) {
  return (size_t)B * H * T * sizeof(float) + (size_t)B * T * (d + 2 * DH) * sizeof(float) + 256;
}

// dqkv (B*T, d+128) bf16 out: gradients w.r.t. the PRE-RoPE q | k | v when cos_sin != NULL (the inverse
// rotation is fused), w.r.t. the rotated tensors otherwise.
extern "C" int tasr_mqa_attention_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse2, int B, int T,
                                      int H, int d, const int64_t* key_lengths, float drop_p, uint64_t seed,
                                      const float* cos_sin, void* dqkv, void* workspace, size_t workspace_bytes,
                                      tasr_stream_t stream) {
  if (B <= 0 || B > ATT_BWD_MAX_B || T <= 0 || H <= 0 || d != H * DH) return TASR_ERR_SHAPE;
  if (workspace_bytes < tasr_mqa_attention_bwd_workspace_bytes(B, T, H, d)) return TASR_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* delta = reinterpret_cast<float*>(workspace);
  size_t off = ((size_t)B * H * T * sizeof(float) + 255) & ~(size_t)255;
  float* dq_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + off);
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, (size_t)B * T * (d + 2 * DH) * sizeof(float), st);
  if (e != cudaSuccess) return tasr_set_cuda_error(e);
  CUtensorMap tm_qkv, tm_do;
  int rc = make_tmap_3d(&tm_qkv, qkv, d + 2 * DH, T, B, d + 2 * DH);
  if (rc) return rc;
  rc = make_tmap_3d(&tm_do, dctx, d, T, B, d);
  if (rc) return rc;
  const long long nw = (long long)B * T * H;
  attn_delta_kernel<<<cdiv(nw * 8, 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(dctx), reinterpret_cast<const bf16*>(ctx),
                                                        B, T, H, d, delta);
  TASR_CHECK_LAUNCH();
  AttnParams p;
  fill_params(&p, B, T, H, d, key_lengths, drop_p, seed);
  p.lse2 = const_cast<float*>(lse2);
  p.delta = delta;
  p.dq_acc = dq_acc;
  p.dqkv = reinterpret_cast<bf16*>(dqkv);
  constexpr int SMEM_MAX = 163840 + 64 + (ATT_BWD_MAX_B + 1) * 4 + 1024;
  const int smem_bytes = 163840 + 64 + (B + 1) * 4 + 1024;
  static TasrPerDevice attr_done;
  const int n_sms = tasr_num_sms();
  if (!attr_done.get()) {
    e = cudaFuncSetAttribute(mqa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done.set();
  }
  const long long max_units = (long long)B * cdiv(T, BKV) * H * cdiv(T, BQ);
  const int grid = (int)imin64(n_sms, max_units);
  mqa_bwd_kernel<<<grid, ATT_BWD_THREADS, smem_bytes, st>>>(tm_qkv, tm_do, p);
  TASR_CHECK_LAUNCH();
  const long long total = (long long)B * T * ((d + 2 * DH) / 8);
  attn_dq_finalize_kernel<<<(int)imin64((long long)148 * 8, (total + 255) / 256), 256, 0, st>>>(
      dq_acc, reinterpret_cast<bf16*>(dqkv), (long long)B * T, T, d, cos_sin);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
