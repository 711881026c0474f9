// Fused log-softmax + CTC loss + gradient w.r.t. the logits.
// Replaces (reference): trainer/trainer.py:167-168 (permute + F.log_softmax), :76,:173
//   nn.CTCLoss(blank=0, zero_infinity=True) forward (ctc_loss_gpu alpha kernel) and its backward
//   (beta kernel + gradient-collect kernel + log_softmax backward): >= 4 passes over (T', B, V) in the
//   reference, here the logits are read once for the row statistics and once for the gradient.
//
//   1. ctc_rowstats: one warp per (b, t < L'_b): lse = logsumexp(logits[b,t,:]); gathers
//      lp[b,t,0] = blank and lp[b,t,1+s] = label s log-probabilities (the only entries the lattice needs).
//   2. ctc_alpha_beta: one CTA per utterance, one thread per lattice state (2S+1) for alpha and one for
//      beta, running concurrently; alpha/beta rows go to an L2-resident workspace.
//   3. ctc_grad: one warp per (b, t): occupancies per label, then writes
//      dlogits = (softmax - occupancy) * grad_scale / (B * max(S_b, 1)), zeros for t >= L'_b.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr float NEG_INF = -INFINITY;

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == NEG_INF) return NEG_INF;
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  if (m == NEG_INF) return NEG_INF;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

template <typename TL>
__device__ __forceinline__ float ldlogit(const TL* p, long long i);
template <>
__device__ __forceinline__ float ldlogit<float>(const float* p, long long i) { return p[i]; }
template <>
__device__ __forceinline__ float ldlogit<bf16>(const bf16* p, long long i) { return __bfloat162float(p[i]); }

template <typename TL>
__global__ void __launch_bounds__(256) ctc_rowstats_kernel(const TL* __restrict__ logits, long long ld, int B, int T, int V,
                                                           const long long* __restrict__ targets, int Smax,
                                                           const long long* __restrict__ in_len,
                                                           const long long* __restrict__ tgt_len, int blank,
                                                           float* __restrict__ lse_out, float* __restrict__ lp) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * T) return;
  const int b = warp / T, t = warp - b * T;
  const int L = (int)min((long long)T, in_len[b]);
  if (t >= L) return;
  const TL* row = logits + ((long long)b * T + t) * ld;
  float lse;
  if (sizeof(TL) == 2 && ((V | (int)ld) & 7) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {
    // 16-byte loads, one pass: per-lane running maximum and rescaled sum, combined across the warp at the end
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
    const int nch = V >> 3;
    float m = NEG_INF, s = 0.f;
    for (int c = lane; c < nch; c += 32) {
      const uint4 q = r4[c];
      const float2 p0 = unpack_bf16x2(q.x), p1 = unpack_bf16x2(q.y), p2 = unpack_bf16x2(q.z), p3 = unpack_bf16x2(q.w);
      const float v[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
      float cm = v[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) cm = fmaxf(cm, v[i]);
      if (cm > m) {
        s *= expf(m - cm);  // m = -inf on the first chunk: s is 0
        m = cm;
      }
      if (m != NEG_INF) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s += expf(v[i] - m);  // a NaN logit makes the row's lse NaN
      }
    }
    const float mw = warp_max(m);
    s = warp_sum(m == NEG_INF ? 0.f : s * expf(m - mw));
    lse = mw + logf(s);
  } else {
    float m = NEG_INF;
    for (int c = lane; c < V; c += 32) m = fmaxf(m, ldlogit(row, c));
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(ldlogit(row, c) - m);
    s = warp_sum(s);
    lse = m + logf(s);
  }
  if (lane == 0) lse_out[(long long)b * T + t] = lse;
  const int S = (int)tgt_len[b];
  float* lpr = lp + ((long long)b * T + t) * (Smax + 1);
  for (int j = lane; j <= S; j += 32) {
    const int c = (j == 0) ? blank : (int)targets[(long long)b * Smax + j - 1];
    lpr[j] = ldlogit(row, c) - lse;
  }
}

// blockDim.x = 2 * NSP (NSP = states rounded up to 32): first half alpha, second half beta.
__global__ void ctc_alpha_beta_kernel(int T, int Smax, const long long* __restrict__ targets,
                                      const long long* __restrict__ in_len, const long long* __restrict__ tgt_len,
                                      int blank, int V, const float* __restrict__ lp, float* __restrict__ alpha,
                                      float* __restrict__ beta, float* __restrict__ nll_out, float* __restrict__ loss,
                                      int B, unsigned short* __restrict__ slot_of_class, int* __restrict__ first_occ) {
  extern __shared__ float sh_ab[];  // [2 dirs][2 buffers][NSP + 2]
  const int b = blockIdx.x;
  const int NSP = blockDim.x >> 1;
  const int dir = threadIdx.x >= NSP ? 1 : 0;
  const int s = threadIdx.x - dir * NSP;
  const int S = (int)tgt_len[b];
  const int L = (int)min((long long)T, in_len[b]);
  const int NS = 2 * S + 1;
  const int NSmax = 2 * Smax + 1;
  const long long* tg = targets + (long long)b * Smax;

  // first occurrence of every label + class -> slot table (used by the gradient kernel)
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const long long tok = tg[j];
    int f = j;
    for (int i = 0; i < j; ++i)
      if (tg[i] == tok) { f = i; break; }
    first_occ[(long long)b * Smax + j] = (tok == blank) ? -1 : f;   // label == blank folds into slot 0
    if (f == j && tok != blank && tok >= 0 && tok < V) slot_of_class[(long long)b * V + tok] = (unsigned short)(j + 1);
  }
  if (threadIdx.x == 0 && blank >= 0 && blank < V) slot_of_class[(long long)b * V + blank] = 0;

  float* bufbase = sh_ab + dir * 2 * (NSP + 2);
  // buffers are addressed with a +2 / +0 guard so that s-1, s-2 (alpha) and s+1, s+2 (beta) are in range
  float* buf0 = bufbase;
  float* buf1 = bufbase + (NSP + 2);
  for (int i = threadIdx.x; i < 4 * (NSP + 2); i += blockDim.x) sh_ab[i] = NEG_INF;
  __syncthreads();

  const bool active = s < NS;
  const int my_src = (s & 1) ? (s >> 1) + 1 : 0;  // column of lp: 0 = blank, 1 + label index
  // can this state take the skip transition (from s-2 for alpha, to s+2 for beta)?
  bool skip = false;
  if (active && (s & 1)) {
    if (dir == 0) skip = (s >= 3) && (tg[s >> 1] != tg[(s >> 1) - 1]);
    else skip = (s + 2 < NS) && (tg[s >> 1] != tg[(s >> 1) + 1]);
  }
  const long long lp_stride = Smax + 1;
  const float* lpb = lp + (long long)b * T * lp_stride;
  float* outp = (dir == 0 ? alpha : beta) + (long long)b * T * NSmax;

  if (L > 0) {
    // initial step
    const int tinit = dir == 0 ? 0 : L - 1;
    float v = NEG_INF;
    if (active) {
      const float e = lpb[(long long)tinit * lp_stride + my_src];
      if (dir == 0) { if (s <= 1) v = e; }
      else { if (s >= NS - 2) v = e; }
      outp[(long long)tinit * NSmax + s] = v;
    }
    float* cur = buf0;
    float* nxt = buf1;
    if (dir == 0) cur[2 + s] = active ? v : NEG_INF; else cur[s] = active ? v : NEG_INF;
    __syncthreads();
    // The emission of step t comes from global memory (L2): fetched PF steps ahead into a register ring, otherwise every
    // step of the serial recursion waits a full L2 round trip (measured: ~800 cycles per step with a one-step prefetch).
    // The loads are unconditional (clamped row, column 0 for idle threads; the surplus values are never consumed): a
    // predicated load is compiled into load + select, and the select waits for the load inside the same step.
    constexpr int PF = 8;
    float e_ring[PF];
    const float* lpe = lpb + (active ? my_src : 0);
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int st = min(1 + u, L - 1);
      e_ring[u] = lpe[(long long)(dir == 0 ? st : L - 1 - st) * lp_stride];
    }
    auto lattice_step = [&](int step, float e) {
      const int t = dir == 0 ? step : L - 1 - step;
      float nv = NEG_INF;
      if (active) {
        float a0, a1, a2;
        if (dir == 0) { a0 = cur[2 + s]; a1 = cur[2 + s - 1]; a2 = skip ? cur[2 + s - 2] : NEG_INF; }
        else { a0 = cur[s]; a1 = cur[s + 1]; a2 = skip ? cur[s + 2] : NEG_INF; }
        // out-of-lattice neighbours: alpha s-1 < 0 reads the guard (-inf); beta s+1 >= NS reads -inf too
        if (dir == 1 && s + 1 >= NS) a1 = NEG_INF;
        // NaN logits stay NaN (the loss and the step are then skipped by the trainer).  ex2.approx / lg2.approx in place
        // of expf / logf were measured: no change (118.3 -> 117.4 us), the step is not bound by this arithmetic.
        nv = lse3(a0, a1, a2) + e;
        outp[(long long)t * NSmax + s] = nv;
      }
      if (dir == 0) nxt[2 + s] = nv; else nxt[s] = nv;
    };
    // full groups of PF steps: no guard inside, and a ring slot is refilled only after its value has been consumed, so
    // the load lands in the slot's own register (a guard or an early refill costs a register move that waits for the load)
    int base = 1;
    for (; base + PF <= L; base += PF) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int step = base + u;
        lattice_step(step, e_ring[u]);
        const int sn = min(step + PF, L - 1);
        e_ring[u] = lpe[(long long)(dir == 0 ? sn : L - 1 - sn) * lp_stride];
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
      }
    }
    // the last L - base < PF steps: their emissions are already in the ring
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      if (base + u < L) {  // L is uniform over the CTA
        lattice_step(base + u, e_ring[u]);
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
      }
    }
    if (dir == 0 && s == 0) {
      const float aL = cur[2 + NS - 1];
      const float aL1 = NS >= 2 ? cur[2 + NS - 2] : NEG_INF;
      float nll = -lse2(aL, aL1);
      nll_out[b] = nll;
      // zero_infinity zeroes infeasible (+inf) samples only: a NaN sample makes the mean NaN, like nn.CTCLoss
      if (nll != INFINITY) atomicAdd(loss, nll / (float)max(S, 1) / (float)B);
    }
  } else if (threadIdx.x == 0) {
    nll_out[b] = (S == 0) ? 0.f : INFINITY;
  }
}

// Same recursion for targets longer than 255 labels (more than 1024 lattice states for both directions): 512 threads
// per direction, each walking its states with a stride.  Long-form transcripts only; not tuned.
__global__ void __launch_bounds__(1024) ctc_alpha_beta_long_kernel(int T, int Smax, const long long* __restrict__ targets,
                                                                   const long long* __restrict__ in_len,
                                                                   const long long* __restrict__ tgt_len, int blank, int V,
                                                                   const float* __restrict__ lp, float* __restrict__ alpha,
                                                                   float* __restrict__ beta, float* __restrict__ nll_out,
                                                                   float* __restrict__ loss, int B,
                                                                   unsigned short* __restrict__ slot_of_class,
                                                                   int* __restrict__ first_occ, int NSP) {
  extern __shared__ float sh_ab[];  // [2 dirs][2 buffers][NSP + 2]
  const int b = blockIdx.x;
  const int NTD = blockDim.x >> 1;
  const int dir = threadIdx.x >= NTD ? 1 : 0;
  const int tid = threadIdx.x - dir * NTD;
  const int S = (int)tgt_len[b];
  const int L = (int)min((long long)T, in_len[b]);
  const int NS = 2 * S + 1, NSmax = 2 * Smax + 1;
  const long long* tg = targets + (long long)b * Smax;
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const long long tok = tg[j];
    int f = j;
    for (int i = 0; i < j; ++i)
      if (tg[i] == tok) { f = i; break; }
    first_occ[(long long)b * Smax + j] = (tok == blank) ? -1 : f;
    if (f == j && tok != blank && tok >= 0 && tok < V) slot_of_class[(long long)b * V + tok] = (unsigned short)(j + 1);
  }
  if (threadIdx.x == 0 && blank >= 0 && blank < V) slot_of_class[(long long)b * V + blank] = 0;
  float* cur = sh_ab + dir * 2 * (NSP + 2);
  float* nxt = cur + (NSP + 2);
  for (int i = threadIdx.x; i < 4 * (NSP + 2); i += blockDim.x) sh_ab[i] = NEG_INF;
  __syncthreads();
  const long long lp_stride = Smax + 1;
  const float* lpb = lp + (long long)b * T * lp_stride;
  float* outp = (dir == 0 ? alpha : beta) + (long long)b * T * NSmax;
  const int off = dir == 0 ? 2 : 0;  // alpha reads s-1, s-2 through a 2-element guard; beta reads s+1, s+2 past the end
  if (L > 0) {
    const int tinit = dir == 0 ? 0 : L - 1;
    for (int s = tid; s < NS; s += NTD) {
      const int src = (s & 1) ? (s >> 1) + 1 : 0;
      float v = NEG_INF;
      if (dir == 0 ? (s <= 1) : (s >= NS - 2)) v = lpb[(long long)tinit * lp_stride + src];
      outp[(long long)tinit * NSmax + s] = v;
      cur[off + s] = v;
    }
    __syncthreads();
    for (int step = 1; step < L; ++step) {
      const int t = dir == 0 ? step : L - 1 - step;
      for (int s = tid; s < NS; s += NTD) {
        const int src = (s & 1) ? (s >> 1) + 1 : 0;
        bool skip = false;
        if (s & 1) {
          if (dir == 0) skip = (s >= 3) && (tg[s >> 1] != tg[(s >> 1) - 1]);
          else skip = (s + 2 < NS) && (tg[s >> 1] != tg[(s >> 1) + 1]);
        }
        float a0, a1, a2;
        if (dir == 0) { a0 = cur[2 + s]; a1 = cur[2 + s - 1]; a2 = skip ? cur[2 + s - 2] : NEG_INF; }
        else { a0 = cur[s]; a1 = (s + 1 < NS) ? cur[s + 1] : NEG_INF; a2 = skip ? cur[s + 2] : NEG_INF; }
        const float nv = lse3(a0, a1, a2) + lpb[(long long)t * lp_stride + src];
        outp[(long long)t * NSmax + s] = nv;
        nxt[off + s] = nv;
      }
      __syncthreads();
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (dir == 0 && tid == 0) {
      const float aL = cur[2 + NS - 1];
      const float aL1 = NS >= 2 ? cur[2 + NS - 2] : NEG_INF;
      const float nll = -lse2(aL, aL1);
      nll_out[b] = nll;
      if (nll != INFINITY) atomicAdd(loss, nll / (float)max(S, 1) / (float)B);
    }
  } else if (threadIdx.x == 0) {
    nll_out[b] = (S == 0) ? 0.f : INFINITY;
  }
}

__device__ __forceinline__ void store_grad(float* p, long long i, float v) { p[i] = v; }
__device__ __forceinline__ void store_grad(bf16* p, long long i, float v) { p[i] = __float2bfloat16(v); }

// TL = logits dtype, TG = gradient dtype (bf16 logits may ask for an fp32 gradient: the arithmetic is fp32 either way,
// bf16 is only the storage format the classifier's backward GEMMs consume)
template <typename TL, typename TG>
__global__ void __launch_bounds__(256) ctc_grad_kernel(const TL* __restrict__ logits, long long ld, int B, int T, int V, int Smax,
                                                       const long long* __restrict__ in_len,
                                                       const long long* __restrict__ tgt_len,
                                                       const float* __restrict__ lse_in, const float* __restrict__ lp,
                                                       const float* __restrict__ alpha, const float* __restrict__ beta,
                                                       const float* __restrict__ nll_in,
                                                       const unsigned short* __restrict__ slot_of_class,
                                                       const int* __restrict__ first_occ, float grad_scale,
                                                       TG* __restrict__ dlogits, long long ldg) {
  extern __shared__ float sh_gam[];  // per warp: Smax + 1 occupancies
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + wib;
  if (warp >= B * T) return;
  const int b = warp / T, t = warp - b * T;
  const int L = (int)min((long long)T, in_len[b]);
  const int S = (int)tgt_len[b];
  const float nll = nll_in[b];
  TG* drow = dlogits + ((long long)b * T + t) * ldg;
  // bf16 in / bf16 out with 16-byte aligned rows: eight classes per load / store
  const bool vec = sizeof(TL) == 2 && sizeof(TG) == 2 && ((V | (int)ld | (int)ldg) & 7) == 0 &&
                   ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits) |
                     reinterpret_cast<uintptr_t>(slot_of_class)) & 15) == 0;
  if (t >= L || nll == INFINITY || nll != nll) {
    // padding frames and infeasible samples: exactly 0 (zero_infinity); a NaN sample poisons its rows so that the
    // non-finite gradient norm makes the optimizer skip the step (reference trainer/trainer.py:178-181)
    const float fill = (nll != nll && t < L) ? nll : 0.f;
    if (vec) {
      const uint32_t f2 = pack_bf16x2(fill, fill);
      for (int c = lane; c < (V >> 3); c += 32) reinterpret_cast<uint4*>(drow)[c] = make_uint4(f2, f2, f2, f2);
    } else {
      for (int c = lane; c < V; c += 32) store_grad(drow, c, fill);
    }
    return;
  }
  float* gam = sh_gam + wib * (Smax + 1);
  for (int j = lane; j <= Smax; j += 32) gam[j] = 0.f;
  __syncwarp();
  const int NS = 2 * S + 1, NSmax = 2 * Smax + 1;
  const float* ar = alpha + ((long long)b * T + t) * NSmax;
  const float* br = beta + ((long long)b * T + t) * NSmax;
  const float* lpr = lp + ((long long)b * T + t) * (Smax + 1);
  for (int s = lane; s < NS; s += 32) {
    const int src = (s & 1) ? (s >> 1) + 1 : 0;
    const float ab = ar[s] + br[s];
    if (ab == NEG_INF) continue;
    const float g = expf(ab + nll - lpr[src]);
    int slot = 0;
    if (s & 1) {
      const int f = first_occ[(long long)b * Smax + (s >> 1)];
      slot = f < 0 ? 0 : f + 1;
    }
    atomicAdd(&gam[slot], g);
  }
  __syncwarp();
  const float scale = grad_scale / ((float)B * (float)max(S, 1));
  const float lse = lse_in[(long long)b * T + t];
  const TL* row = logits + ((long long)b * T + t) * ld;
  const unsigned short* slots = slot_of_class + (long long)b * V;
  if (vec) {
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
    const uint4* s4 = reinterpret_cast<const uint4*>(slots);
    for (int c = lane; c < (V >> 3); c += 32) {
      const uint4 q = r4[c], sq = s4[c];
      const uint32_t qw[4] = {q.x, q.y, q.z, q.w}, sw[4] = {sq.x, sq.y, sq.z, sq.w};
      uint32_t ow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = unpack_bf16x2(qw[i]);
        float v0 = expf(x.x - lse), v1 = expf(x.y - lse);
        const uint32_t s0 = sw[i] & 0xFFFFu, s1 = sw[i] >> 16;
        if (s0 != 0xFFFFu) v0 -= gam[s0];
        if (s1 != 0xFFFFu) v1 -= gam[s1];
        ow[i] = pack_bf16x2(v0 * scale, v1 * scale);
      }
      reinterpret_cast<uint4*>(drow)[c] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    return;
  }
  for (int c = lane; c < V; c += 32) {
    float v = expf(ldlogit(row, c) - lse);
    const unsigned short sl = slots[c];
    if (sl != 0xFFFF) v -= gam[sl];
    store_grad(drow, c, v * scale);
  }
}

inline int nsp_for(int Smax) { return ((2 * Smax + 1 + 31) / 32) * 32; }

}  // namespace

// workspace layout: lse (B*T) | lp (B*T*(Smax+1)) | alpha, beta (B*T*(2Smax+1) each) | nll (B) |
//                   first_occ (B*Smax int) | slot_of_class (B*V u16)
extern "C" size_t tasr_ctc_workspace_bytes(int B, int T, int V, int Smax) {
  size_t n = 0;
  n += (size_t)B * T * 4;
  n += (size_t)B * T * (Smax + 1) * 4;
  n += (size_t)2 * B * T * (2 * Smax + 1) * 4;
  n += (size_t)B * 4;
  n += (size_t)B * (Smax > 0 ? Smax : 1) * 4;
  n += (size_t)B * V * 2;
  return n + 256;
}

extern "C" int tasr_ctc_loss_fwd_bwd(const void* logits, int logits_bf16, int64_t ld, int B, int T, int V, const int64_t* targets,
                                     int Smax, const int64_t* input_lengths, const int64_t* target_lengths, int blank,
                                     float grad_scale, float* loss, float* nll, void* dlogits, void* workspace,
                                     size_t workspace_bytes, tasr_stream_t stream) {
  if (B <= 0 || T <= 0 || V <= 0 || Smax < 0 || V > 65535 || ld < V) return TASR_ERR_SHAPE;
  const int NSP = nsp_for(Smax);
  const size_t ab_smem = (size_t)4 * (NSP + 2) * sizeof(float);
  if (ab_smem > 200 * 1024) return TASR_ERR_SHAPE;  // targets beyond ~6000 labels per utterance
  if (workspace_bytes < tasr_ctc_workspace_bytes(B, T, V, Smax)) return TASR_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  float* lse = reinterpret_cast<float*>(w); w += (size_t)B * T * 4;
  float* lp = reinterpret_cast<float*>(w); w += (size_t)B * T * (Smax + 1) * 4;
  float* alpha = reinterpret_cast<float*>(w); w += (size_t)B * T * (2 * Smax + 1) * 4;
  float* beta = reinterpret_cast<float*>(w); w += (size_t)B * T * (2 * Smax + 1) * 4;
  float* nll_ws = reinterpret_cast<float*>(w); w += (size_t)B * 4;
  int* first_occ = reinterpret_cast<int*>(w); w += (size_t)B * (Smax > 0 ? Smax : 1) * 4;
  w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(w) + 15) & ~(uintptr_t)15);  // 16-byte loads in ctc_grad (slack: +256)
  unsigned short* slots = reinterpret_cast<unsigned short*>(w);
  cudaError_t e = cudaMemsetAsync(slots, 0xFF, (size_t)B * V * 2, st);
  if (e != cudaSuccess) return tasr_set_cuda_error(e);
  e = cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (e != cudaSuccess) return tasr_set_cuda_error(e);
  const long long* tg = reinterpret_cast<const long long*>(targets);
  const long long* il = reinterpret_cast<const long long*>(input_lengths);
  const long long* tl = reinterpret_cast<const long long*>(target_lengths);
  const int rows = B * T;
  const int grid_rows = cdiv((long long)rows * 32, 256);
  if (logits_bf16)
    ctc_rowstats_kernel<bf16><<<grid_rows, 256, 0, st>>>(reinterpret_cast<const bf16*>(logits), ld, B, T, V, tg, Smax, il, tl,
                                                         blank, lse, lp);
  else
    ctc_rowstats_kernel<float><<<grid_rows, 256, 0, st>>>(reinterpret_cast<const float*>(logits), ld, B, T, V, tg, Smax, il, tl,
                                                          blank, lse, lp);
  TASR_CHECK_LAUNCH();
  float* nll_dst = nll != nullptr ? nll : nll_ws;
  if (2 * NSP <= 1024) {  // one thread per lattice state (targets up to 255 labels)
    ctc_alpha_beta_kernel<<<B, 2 * NSP, ab_smem, st>>>(T, Smax, tg, il, tl, blank, V, lp, alpha, beta, nll_dst, loss, B, slots,
                                                       first_occ);
  } else {
    if (ab_smem > 48 * 1024) {
      cudaError_t ea = cudaFuncSetAttribute(ctc_alpha_beta_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ab_smem);
      if (ea != cudaSuccess) return tasr_set_cuda_error(ea);
    }
    ctc_alpha_beta_long_kernel<<<B, 1024, ab_smem, st>>>(T, Smax, tg, il, tl, blank, V, lp, alpha, beta, nll_dst, loss, B, slots,
                                                         first_occ, NSP);
  }
  TASR_CHECK_LAUNCH();
  if (dlogits != nullptr) {
    const size_t sm = (size_t)8 * (Smax + 1) * sizeof(float);
    auto launch_grad = [&](auto kern, auto* lg, auto* dg) -> cudaError_t {
      if (sm > 48 * 1024) {
        cudaError_t eg = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (eg != cudaSuccess) return eg;
      }
      kern<<<grid_rows, 256, sm, st>>>(lg, ld, B, T, V, Smax, il, tl, lse, lp, alpha, beta, nll_dst, slots, first_occ, grad_scale,
                                       dg, ld);
      return cudaSuccess;
    };
    cudaError_t eg;
    if (logits_bf16 == 1)
      eg = launch_grad(ctc_grad_kernel<bf16, bf16>, reinterpret_cast<const bf16*>(logits), reinterpret_cast<bf16*>(dlogits));
    else if (logits_bf16 == 2)
      eg = launch_grad(ctc_grad_kernel<bf16, float>, reinterpret_cast<const bf16*>(logits), reinterpret_cast<float*>(dlogits));
    else
      eg = launch_grad(ctc_grad_kernel<float, float>, reinterpret_cast<const float*>(logits), reinterpret_cast<float*>(dlogits));
    if (eg != cudaSuccess) return tasr_set_cuda_error(eg);
    TASR_CHECK_LAUNCH();
  }
  return TASR_OK;
}
