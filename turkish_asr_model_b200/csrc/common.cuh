// Shared device/host helpers for the sm_100a kernels of libtasr_kernels.so.
// Everything here is written for Blackwell (B200, sm_100a) only: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (UMMA + TMEM) wrappers as inline PTX.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/tasr_kernels.h"

extern unsigned long long g_tasr_launches;
// optional device counter added to every dropout seed (lets a captured CUDA graph draw fresh masks per replay)
extern const unsigned long long* g_tasr_seed_ptr;  // kernels launched by this library (bench.py reports it)

#define TASR_CHECK_LAUNCH()                                     \
  do {                                                          \
    ++g_tasr_launches;                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return tasr_set_cuda_error(e__);    \
  } while (0)

int tasr_set_cuda_error(cudaError_t e);  // records the message, returns TASR_ERR_CUDA

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

#ifdef __CUDACC__
// launch with the programmatic-dependent-launch attribute (the kernel must call grid_dependency_wait())
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                             int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = cluster_x;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
#endif

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// One-time-per-DEVICE flag (cudaFuncSetAttribute and friends are per device, a process may drive several GPUs).
struct TasrPerDevice {
  unsigned long long mask = 0ull;  // bit = device ordinal (ordinals >= 64 are simply redone on every call)
  static int current() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
  }
  bool get() const {
    const int dev = current();
    return dev < 64 && ((__atomic_load_n(&mask, __ATOMIC_ACQUIRE) >> dev) & 1ull);
  }
  void set() {
    const int dev = current();
    if (dev < 64) __atomic_fetch_or(&mask, 1ull << dev, __ATOMIC_RELEASE);
  }
};
// SM count of the current device (cached per device)
static inline int tasr_num_sms() {
  static int cache[64] = {0};
  const int dev = TasrPerDevice::current();
  if (dev < 64 && cache[dev] > 0) return cache[dev];
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev < 64) cache[dev] = n;
  return n;
}
static inline long long imin64(long long a, long long b) { return a < b ? a : b; }

// ----------------------------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// d/dx silu(x) = s + x*s*(1-s)
__device__ __forceinline__ float silu_gradf_(float x) {
  float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}
// MUFU.TANH forms: sigmoid(x) = 1/2 + tanh(x/2)/2 (one SFU op instead of ex2 + rcp)
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_tanh(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx(h), h);
}
// d/dx silu(x) = (1 + t)(1 + h (1 - t)) / 2,  h = x/2, t = tanh(h)
__device__ __forceinline__ float silu_grad_tanh(float x) {
  const float h = 0.5f * x, t = tanh_approx(h);
  return 0.5f * (1.f + t) * fmaf(h, 1.f - t, 1.f);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  bf162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}

// Counter-based dropout mask, evaluated identically in forward and backward (no mask is stored).
// One hash serves TWO neighbouring elements: with x0 = lo32(pair) + hi32(pair) * GOLD and s32 = a 32-bit digest of the
// 64-bit seed,
//   t = x0 * C1 + s32;  t ^= t >> 15;  t *= C2;   lane1 = t >> 16,  lane0 = (t * GOLD) >> 16
// and element idx is kept iff lane(idx & 1) >= thresh16 = round(p * 65536).  For a run of pairs the affine part is
// computed once (base = x0 * C1 + s32) and pair j costs: one add (base + j * C1), the xorshift, two multiplies and
// two compares (lane >= thresh16  <=>  word >= thresh16 << 16).
#define TASR_HASH_C1 0x7FEB352Du
#define TASR_HASH_C2 0x846CA68Bu
#define TASR_HASH_GOLD 0x9E3779B1u
__host__ __device__ __forceinline__ uint32_t tasr_seed_mix(unsigned long long seed) {
  uint32_t x = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x85EBCA6Bu);
  x ^= x >> 16; x *= TASR_HASH_C1;
  x ^= x >> 15; x *= TASR_HASH_C2;
  x ^= x >> 16;
  return x;
}
// affine part of the hash of pair `pair0` (pair0 + j follows by adding j * C1 as long as lo32 does not wrap)
__host__ __device__ __forceinline__ uint32_t tasr_hash_pair_base_s32(uint32_t s32, unsigned long long pair0) {
  return ((uint32_t)pair0 + (uint32_t)(pair0 >> 32) * TASR_HASH_GOLD) * TASR_HASH_C1 + s32;
}
__host__ __device__ __forceinline__ uint32_t tasr_hash_pair_base(unsigned long long seed, unsigned long long pair0) {
  return tasr_hash_pair_base_s32(tasr_seed_mix(seed), pair0);
}
__host__ __device__ __forceinline__ uint32_t tasr_hash_finish(uint32_t t) {
  t ^= t >> 15;
  t *= TASR_HASH_C2;
  return t;
}
// both 16-bit lanes packed: lane0 in the low half, lane1 in the high half
__host__ __device__ __forceinline__ uint32_t tasr_hash_pair(unsigned long long seed, unsigned long long pair_idx) {
  const uint32_t t = tasr_hash_finish(tasr_hash_pair_base(seed, pair_idx));
  return (t & 0xFFFF0000u) | ((t * TASR_HASH_GOLD) >> 16);
}
__host__ __device__ __forceinline__ uint32_t tasr_drop_thresh16(float p) {
  if (!(p > 0.f)) return 0u;
  double t = (double)p * 65536.0 + 0.5;
  uint32_t v = t >= 65535.0 ? 65535u : (uint32_t)t;
  return v == 0 ? 1u : v;
}
__host__ __device__ __forceinline__ float tasr_drop_inv_keep(uint32_t thresh16) {
  return thresh16 ? 65536.f / (65536.f - (float)thresh16) : 1.f;
}
__device__ __forceinline__ float dropout_scale(unsigned long long seed, unsigned long long idx,
                                               uint32_t thresh16, float inv_keep) {
  const uint32_t h = tasr_hash_pair(seed, idx >> 1);
  const uint32_t lane = (idx & 1) ? (h >> 16) : (h & 0xFFFFu);
  return lane >= thresh16 ? inv_keep : 0.f;
}
// pair j of a run whose affine part is `base32` (tasr_hash_pair_base of the run's first pair)
__device__ __forceinline__ void dropout_keep2_fast(uint32_t base32, uint32_t j, uint32_t thresh_hi, bool& keep0, bool& keep1) {
  const uint32_t t = tasr_hash_finish(base32 + j * TASR_HASH_C1);
  keep1 = t >= thresh_hi;                    // thresh_hi = thresh16 << 16
  keep0 = t * TASR_HASH_GOLD >= thresh_hi;
}
__device__ __forceinline__ void dropout_scale2_fast(uint32_t base32, uint32_t /*unused*/, uint32_t j, uint32_t thresh16,
                                                    float inv_keep, float& s0, float& s1) {
  bool k0, k1;
  dropout_keep2_fast(base32, j, thresh16 << 16, k0, k1);
  s0 = k0 ? inv_keep : 0.f;
  s1 = k1 ? inv_keep : 0.f;
}
// both elements of the pair (idx even, idx + 1) with one hash
__device__ __forceinline__ void dropout_scale2(unsigned long long seed, unsigned long long idx_even, uint32_t thresh16,
                                               float inv_keep, float& s0, float& s1) {
  const uint32_t h = tasr_hash_pair(seed, idx_even >> 1);
  s0 = (h & 0xFFFFu) >= thresh16 ? inv_keep : 0.f;
  s1 = (h >> 16) >= thresh16 ? inv_keep : 0.f;
}

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// PTX: mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// same, yielding the issue slots between polls (producer / MMA warps share their SM sub-partitions with epilogue warps)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(40);
}
// generic-proxy smem writes -> visible to the async proxy (TMA / UMMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// PTX: TMA loads (tile mode), completing on an mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// multicast: the box lands at the same shared-memory offset (and signals the mbarrier at the same offset) in every CTA
// of the cluster whose bit is set in cta_mask
__device__ __forceinline__ void tma_load_2d_mc(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// PTX: tcgen05 (TMEM alloc, UMMA, commit, ld)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued UMMAs of this thread arrive on `bar` when complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at this offset in every CTA of the cluster named by cta_mask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets lane (quarter*32+i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)m),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// PTX: TMA stores / reductions (shared -> global), bulk async groups, named barriers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   (uint64_t)m),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor drains; everything before this wait (barrier init, TMEM allocation, descriptor prefetch)
// overlaps with the predecessor's tail, everything after it sees the predecessor's memory.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// UMMA shared-memory matrix descriptor, 128-byte swizzle (layout type 2), descriptor version 1.
//   start address / LBO / SBO are encoded in 16-byte units.
//   K-major tile  (rows = M|N index, 128 B = 64 bf16 of K per row): SBO = 1024 (8 rows), LBO unused (=16 B)
//   MN-major tile (rows = K index, 128 B = 64 bf16 of M|N per row): SBO = 1024 (8 K rows),
//                                                                   LBO = bytes between 64-wide M|N chunks
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // version = 1 (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// Same, no swizzle (layout type 0), K-major: core matrices of 8 rows x 16 B stored contiguously (128 B each);
// LBO = bytes between core matrices adjacent along K, SBO = bytes between 8-row groups along M|N.
__device__ __forceinline__ uint64_t umma_desc_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // version = 1 (Blackwell)
  return d;
}
// UMMA instruction descriptor: bf16 x bf16 -> fp32, M x N tile, per-operand major-ness (0 = K, 1 = MN)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4)                       // D format: fp32
         | (1u << 7)                     // A format: bf16
         | (1u << 10)                    // B format: bf16
         | ((uint32_t)a_mn_major << 15)  //
         | ((uint32_t)b_mn_major << 16)  //
         | ((uint32_t)(n >> 3) << 17)    //
         | ((uint32_t)(m >> 4) << 24);
}
#endif  // __CUDACC__
