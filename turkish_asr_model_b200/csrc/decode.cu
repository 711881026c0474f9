// Greedy CTC decoding up to token IDs: argmax over the vocabulary (first maximum on ties, like
// torch.argmax), collapse repeats, drop blanks.
// Replaces (reference): utils/decoding.py:149,163 (torch.argmax), inference.py:125, utils/metrics.py:24
//   and the Python loop of data/tokenizer.py:44-54 (ctc_decode) up to the filtered id list.
#include "common.cuh"

namespace {

template <typename TL>
__device__ __forceinline__ float ldv(const TL* p, long long i);
template <>
__device__ __forceinline__ float ldv<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float ldv<bf16>(const bf16* p, long long i) { return __bfloat162float(p[i]); }

template <typename TL>
__global__ void __launch_bounds__(256) argmax_kernel(const TL* __restrict__ logits, long long ld, long long rows, int V,
                                                     long long* __restrict__ ids) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const TL* row = logits + warp * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < V; c += 32) {
    const float v = ldv(row, c);
    if (v > best || (v != v && best == best)) { best = v; bi = c; }  // NaN counts as maximal (torch semantics)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const bool take = (ob > best) || (ob != ob && best == best) || (ob == best && oi < bi) || (ob != ob && best != best && oi < bi);
    if (take) { best = ob; bi = oi; }
  }
  if (lane == 0) ids[warp] = bi == 0x7fffffff ? 0 : bi;
}

// one CTA per utterance: keep[t] = id[t] != id[t-1] && id[t] != blank, compact in order
__global__ void __launch_bounds__(256) collapse_kernel(const long long* __restrict__ ids, int T,
                                                       const long long* __restrict__ lengths, int blank,
                                                       long long* __restrict__ tokens, int* __restrict__ out_len) {
  const int b = blockIdx.x;
  const int L = lengths ? (int)min((long long)T, lengths[b]) : T;
  const long long* row = ids + (long long)b * T;
  long long* dst = tokens + (long long)b * T;
  __shared__ int warp_counts[8];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < L; t0 += 256) {
    const int t = t0 + threadIdx.x;
    bool keep = false;
    long long id = 0;
    if (t < L) {
      id = row[t];
      keep = (id != blank) && (t == 0 || id != row[t - 1]);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_counts[w] = __popc(mask);
    __syncthreads();
    int off = base;
    for (int i = 0; i < w; ++i) off += warp_counts[i];
    if (keep) dst[off + __popc(mask & ((1u << lane) - 1))] = id;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int i = 0; i < 8; ++i) tot += warp_counts[i];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_len[b] = base;
  for (int t = base + threadIdx.x; t < T; t += 256) dst[t] = -1;
}
}  // namespace

// ids (B, T) int64: per-frame argmax; tokens (B, T) int64: collapsed ids, padded with -1; out_len (B) int32
extern "C" int tasr_argmax_collapse(const void* logits, int logits_bf16, int64_t ld, int B, int T, int V, const int64_t* lengths,
                                    int blank, int64_t* ids, int64_t* tokens, int32_t* out_len, tasr_stream_t stream) {
  if (B <= 0 || T <= 0 || V <= 0 || ld < V) return TASR_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)B * T;
  if (logits_bf16)
    argmax_kernel<bf16><<<cdiv(rows * 32, 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(logits), ld, rows, V,
                                                              reinterpret_cast<long long*>(ids));
  else
    argmax_kernel<float><<<cdiv(rows * 32, 256), 256, 0, st>>>(reinterpret_cast<const float*>(logits), ld, rows, V,
                                                               reinterpret_cast<long long*>(ids));
  TASR_CHECK_LAUNCH();
  if (tokens != nullptr) {
    collapse_kernel<<<B, 256, 0, st>>>(reinterpret_cast<const long long*>(ids), T, reinterpret_cast<const long long*>(lengths),
                                       blank, reinterpret_cast<long long*>(tokens), out_len);
    TASR_CHECK_LAUNCH();
  }
  return TASR_OK;
}
