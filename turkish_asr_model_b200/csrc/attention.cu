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

// dropout index space: rows of even pitch so that (key, key + 1) pairs share one hash in forward and backward
__device__ __forceinline__ unsigned long long drop_index(int b, int h, int H, int T, int q, int k) {
  return (((unsigned long long)(b * H + h) * T + q) * (unsigned long long)((T + 1) & ~1)) + k;
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
          float s0, s1;
          dropout_scale2(dseed, drow + (unsigned long long)(j * BKV + c * 32 + i), p.drop_thresh, p.drop_inv_keep, s0, s1);
          e0 *= s0;
          e1 *= s1;
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
__global__ void attn_delta_kernel(const bf16* __restrict__ dctx, const bf16* __restrict__ ctx, int B, int T, int H, int d,
                                  float* __restrict__ delta) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)B * T * H) return;
  const long long row = warp / H;
  const int h = (int)(warp - row * H);
  const long long off = row * d + h * DH + lane * 2;
  const float2 a = __bfloat1622float2(*reinterpret_cast<const bf162*>(dctx + off));
  const float2 c = __bfloat1622float2(*reinterpret_cast<const bf162*>(ctx + off));
  const float s = warp_sum(a.x * c.x + a.y * c.y);
  if (lane == 0) {
    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
    delta[((long long)b * H + h) * T + t] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// backward: grid (ceil(T/128) kv blocks, B); loops over heads and query tiles.
// 256 threads: warps 0-3 own columns [0,64) of the S / dP tiles, warps 4-7 columns [64,128) (one TMEM lane =
// one query row per thread pair).  Q / dO tiles are double-buffered: the TMA for iteration it+1 is in flight
// while iteration it runs.
// ------------------------------------------------------------------------------------------------
constexpr int ATT_BWD_THREADS = 256;

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

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int rloc = quarter * 32 + lane;  // TMEM lane = row of the tile
  const int j = blockIdx.x, b = blockIdx.y;
  const int k0 = j * BKV;
  const int Lk = p.key_len ? (int)max(0LL, min((long long)p.T, p.key_len[b])) : p.T;
  const int ld = p.d + 2 * DH;
  const int krow = k0 + rloc;

  if (k0 >= Lk) {  // fully masked key block: zero gradients
    if (krow < p.T) {
      uint4* dst = reinterpret_cast<uint4*>(p.dqkv + ((long long)b * p.T + krow) * ld + p.d) + half * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tDP = tmem + 128, tDV = tmem + 256, tDK = tmem + 320, tDQ = tmem + 384;
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;

  const int nq = (p.T + BQ - 1) / BQ;
  const int niter = p.H * nq;
  auto issue_qdo = [&](int it) {  // thread 0 only
    const int h = it / nq, q0 = (it - h * nq) * BQ, buf = it & 1;
    uint64_t* bar = &bars[1 + buf];
    mbar_expect_tx(bar, 32768);
    tma_load_3d(sQ + buf * 16384, &tm_qkv, bar, h * DH, q0, b);
    tma_load_3d(sQ + buf * 16384 + 8192, &tm_qkv, bar, h * DH, q0 + 64, b);
    tma_load_3d(sDO + buf * 16384, &tm_do, bar, h * DH, q0, b);
    tma_load_3d(sDO + buf * 16384 + 8192, &tm_do, bar, h * DH, q0 + 64, b);
  };
  if (tid == 0) {
    mbar_expect_tx(&bars[0], 32768);
    tma_load_3d(sK, &tm_qkv, &bars[0], p.d, k0, b);
    tma_load_3d(sK + 8192, &tm_qkv, &bars[0], p.d, k0 + 64, b);
    tma_load_3d(sV, &tm_qkv, &bars[0], p.d + DH, k0, b);
    tma_load_3d(sV + 8192, &tm_qkv, &bars[0], p.d + DH, k0 + 64, b);
    issue_qdo(0);
  }
  const unsigned long long dseed = (p.drop_thresh && p.seed_ptr) ? p.seed + *p.seed_ptr : p.seed;
  const float scale2 = p.scale * LOG2E;
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, BKV, 0, 0);     // Q K^T, dO V^T
  constexpr uint32_t idesc_t = umma_idesc_bf16(128, DH, 1, 1);      // P^T dO, dS^T Q
  constexpr uint32_t idesc_q = umma_idesc_bf16(128, DH, 0, 1);      // dS K
  const int Tp = (p.T + 1) & ~1;  // even row pitch of the dropout index space

  for (int it = 0; it < niter; ++it) {
    const int h = it / nq, qi = it - h * nq;
    const int buf = it & 1;
    const uint32_t ph = (uint32_t)it & 1u;          // mma barriers complete once per iteration
    const uint32_t qph = (uint32_t)(it >> 1) & 1u;  // each q/dO buffer completes every other iteration
    const int q0 = qi * BQ;
    uint8_t* cQ = sQ + buf * 16384;
    uint8_t* cDO = sDO + buf * 16384;
    if (tid == 0) {
      if (it + 1 < niter) issue_qdo(it + 1);  // the other buffer was released by the end of iteration it-1
      if (it == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1 + buf], qph);
      tc_fence_after();
      const uint32_t qa = smem_u32(cQ), ka = smem_u32(sK), da = smem_u32(cDO), va = smem_u32(sV);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        umma_bf16(tS, umma_desc_sw128(qa + k * 32, 16, 1024), umma_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        umma_bf16(tDP, umma_desc_sw128(da + k * 32, 16, 1024), umma_desc_sw128(va + k * 32, 16, 1024), idesc_s, k > 0);
      umma_commit(&bars[3]);
    }
    const int qrow = q0 + rloc;
    const bool qvalid = qrow < p.T;
    float lse2 = 0.f, delta = 0.f;
    if (qvalid) {
      lse2 = p.lse2[((long long)b * p.H + h) * p.T + qrow];
      delta = p.delta[((long long)b * p.H + h) * p.T + qrow];
    }
    const bool row_ok = qvalid && lse2 != -INFINITY;
    const unsigned long long drow = ((unsigned long long)(b * p.H + h) * p.T + qrow) * (unsigned long long)Tp;
    mbar_wait(&bars[3], ph);
    __syncwarp();
    tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;  // 32-column chunk of the 128-key tile
      uint32_t us[32], ud[32];
      tmem_ld32(tS + lane_addr + c * 32, us);
      tmem_ld32(tDP + lane_addr + c * 32, ud);
      tmem_ld_wait();
      float pv[32], dsv[32];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int key = k0 + c * 32 + i;
        float s0 = 1.f, s1 = 1.f;
        if (p.drop_thresh) dropout_scale2(dseed, drow + key, p.drop_thresh, p.drop_inv_keep, s0, s1);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          float pr = 0.f, ds = 0.f;
          if (row_ok && key + jj < Lk) {
            const float ms = jj == 0 ? s0 : s1;
            pr = fast_exp2(__uint_as_float(us[i + jj]) * scale2 - lse2);
            ds = pr * (__uint_as_float(ud[i + jj]) * ms - delta) * p.scale;
            pr *= ms;
          }
          pv[i + jj] = pr;
          dsv[i + jj] = ds;
        }
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
#pragma unroll
      for (int k = 0; k < BQ / 16; ++k)  // dV += P^T dO   (reduction over query rows)
        umma_bf16(tDV, umma_desc_sw128(pa + k * 2048, 16384, 1024), umma_desc_sw128(da + k * 2048, 8192, 1024), idesc_t,
                  (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < BQ / 16; ++k)  // dK += dS^T Q
        umma_bf16(tDK, umma_desc_sw128(sa + k * 2048, 16384, 1024), umma_desc_sw128(qa + k * 2048, 8192, 1024), idesc_t,
                  (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k)  // dQ = dS K    (reduction over keys)
        umma_bf16(tDQ, umma_desc_sw128(sa + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                  umma_desc_sw128(ka + k * 2048, 8192, 1024), idesc_q, k > 0);
      umma_commit(&bars[4]);
    }
    mbar_wait(&bars[4], ph);
    __syncwarp();
    tc_fence_after();
    {
      uint32_t u[32];
      tmem_ld32(tDQ + lane_addr + half * 32, u);
      tmem_ld_wait();
      if (qvalid) {
        float4* dst = reinterpret_cast<float4*>(p.dq_acc + ((long long)b * p.T + qrow) * p.d + h * DH + half * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          atomicAdd(dst + i, make_float4(__uint_as_float(u[4 * i]), __uint_as_float(u[4 * i + 1]),
                                         __uint_as_float(u[4 * i + 2]), __uint_as_float(u[4 * i + 3])));
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  // dK, dV of this key block (summed over heads and query tiles); half 0 writes dK, half 1 writes dV
  {
    bf16* dst = p.dqkv + ((long long)b * p.T + krow) * ld + p.d + half * DH;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t u[32];
      tmem_ld32((half == 0 ? tDK : tDV) + c * 32 + lane_addr, u);
      tmem_ld_wait();
      if (krow < p.T) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(u[8 * i + 0]), __uint_as_float(u[8 * i + 1]));
          v.y = pack_bf16x2(__uint_as_float(u[8 * i + 2]), __uint_as_float(u[8 * i + 3]));
          v.z = pack_bf16x2(__uint_as_float(u[8 * i + 4]), __uint_as_float(u[8 * i + 5]));
          v.w = pack_bf16x2(__uint_as_float(u[8 * i + 6]), __uint_as_float(u[8 * i + 7]));
          reinterpret_cast<uint4*>(dst + c * 32)[i] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// dq (fp32 accumulate) -> inverse RoPE -> bf16 into dqkv[:, 0:d];  inverse RoPE in place on dk
__global__ void __launch_bounds__(256) attn_dq_finalize_kernel(const float* __restrict__ dq_acc, bf16* __restrict__ dqkv,
                                                               long long M, int T, int d, const float* __restrict__ cs) {
  const int ld = d + 2 * DH;
  const int pairs = (d + DH) >> 1;
  const long long total = M * pairs;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const long long row = idx / pairs;
    const int pidx = (int)(idx - row * pairs);
    const int head = pidx >> 5, i = pidx & 31;
    const int t = (int)(row % T);
    float c = 1.f, s = 0.f;
    if (cs != nullptr) { c = cs[(t * 32 + i) * 2]; s = -cs[(t * 32 + i) * 2 + 1]; }
    bf16* o = dqkv + row * ld + head * DH;
    float x1, x2;
    if (head * DH < d) {
      const float* src = dq_acc + row * d + head * DH;
      x1 = src[i]; x2 = src[i + 32];
    } else {
      x1 = __bfloat162float(o[i]); x2 = __bfloat162float(o[i + 32]);
    }
    o[i] = __float2bfloat16(x1 * c - x2 * s);
    o[i + 32] = __float2bfloat16(x2 * c + x1 * s);
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
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(mqa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done = true;
  }
  dim3 grid(cdiv(T, BQ), H, B);
  mqa_fwd_kernel<<<grid, ATT_THREADS, SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(tm, p);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" size_t tasr_mqa_attention_bwd_workspace_bytes(int B, int T, int H, int d) {
  return (size_t)B * H * T * sizeof(float) + (size_t)B * T * d * sizeof(float) + 256;
}

// dqkv (B*T, d+128) bf16 out: gradients w.r.t. the PRE-RoPE q | k | v when cos_sin != NULL (the inverse
// rotation is fused), w.r.t. the rotated tensors otherwise.
extern "C" int tasr_mqa_attention_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse2, int B, int T,
                                      int H, int d, const int64_t* key_lengths, float drop_p, uint64_t seed,
                                      const float* cos_sin, void* dqkv, void* workspace, size_t workspace_bytes,
                                      tasr_stream_t stream) {
  if (B <= 0 || T <= 0 || H <= 0 || d != H * DH) return TASR_ERR_SHAPE;
  if (workspace_bytes < tasr_mqa_attention_bwd_workspace_bytes(B, T, H, d)) return TASR_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* delta = reinterpret_cast<float*>(workspace);
  size_t off = ((size_t)B * H * T * sizeof(float) + 255) & ~(size_t)255;
  float* dq_acc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + off);
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, (size_t)B * T * d * sizeof(float), st);
  if (e != cudaSuccess) return tasr_set_cuda_error(e);
  CUtensorMap tm_qkv, tm_do;
  int rc = make_tmap_3d(&tm_qkv, qkv, d + 2 * DH, T, B, d + 2 * DH);
  if (rc) return rc;
  rc = make_tmap_3d(&tm_do, dctx, d, T, B, d);
  if (rc) return rc;
  const long long nw = (long long)B * T * H;
  attn_delta_kernel<<<cdiv(nw * 32, 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(dctx), reinterpret_cast<const bf16*>(ctx),
                                                        B, T, H, d, delta);
  TASR_CHECK_LAUNCH();
  AttnParams p;
  fill_params(&p, B, T, H, d, key_lengths, drop_p, seed);
  p.lse2 = const_cast<float*>(lse2);
  p.delta = delta;
  p.dq_acc = dq_acc;
  p.dqkv = reinterpret_cast<bf16*>(dqkv);
  constexpr int SMEM = 163840 + 64 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    e = cudaFuncSetAttribute(mqa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done = true;
  }
  dim3 grid(cdiv(T, BKV), B);
  mqa_bwd_kernel<<<grid, ATT_BWD_THREADS, SMEM, st>>>(tm_qkv, tm_do, p);
  TASR_CHECK_LAUNCH();
  const long long total = (long long)B * T * ((d + DH) / 2);
  attn_dq_finalize_kernel<<<(int)imin64((long long)148 * 8, (total + 255) / 256), 256, 0, st>>>(
      dq_acc, reinterpret_cast<bf16*>(dqkv), (long long)B * T, T, d, cos_sin);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
