// GroupNorm over (channels-in-group x ALL padded time) per sample on token-major (B, T, d) tensors,
// and BatchNorm1d(+SiLU) over (B*T) per channel, forward and backward.  HBM-bound kernels:
// vectorised 16-byte accesses, warp-shuffle / shared-memory reductions, partials reduced in double.
// Replaces (reference): model/conformer.py:45-49 TransposeGroupNorm.forward (2 transposes +
//   native_group_norm) and its backward; :84-85 BatchNorm1d + SiLU in ConformerConvModule.
#include "common.cuh"
#include <stdlib.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_bf16(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4_bf16(bf16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---------------------------------------------------------------- GroupNorm forward
// partial[b][chunk][G][2] = (sum, sumsq) over the chunk's rows
__global__ void __launch_bounds__(NT) gn_stats_kernel(const float* __restrict__ x, int T, int d, int G, int rows_per_cta,
                                                      float* __restrict__ partial) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tpr = d >> 2;             // threads per row
  const int rlanes = NT / tpr;        // rows in flight
  const int col = (threadIdx.x % tpr) << 2;
  const int rl = threadIdx.x / tpr;
  const int t0 = chunk * rows_per_cta, t1 = min(T, t0 + rows_per_cta);
  float s = 0.f, ss = 0.f;
#pragma unroll 4
  for (int t = t0 + rl; t < t1; t += rlanes) {
    float4 v = ld4(x + ((long long)b * T + t) * d + col);
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  __shared__ float sh_s[NT], sh_ss[NT];
  sh_s[threadIdx.x] = s;
  sh_ss[threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < G) {
    const int cpg = d / G, tpg = cpg >> 2;  // threads (float4 columns) per group
    float a = 0.f, c = 0.f;
    for (int r = 0; r < rlanes; ++r)
      for (int j = 0; j < tpg; ++j) {
        int idx = r * tpr + threadIdx.x * tpg + j;
        a += sh_s[idx];
        c += sh_ss[idx];
      }
    float* o = partial + (((long long)b * gridDim.x + chunk) * G + threadIdx.x) * 2;
    o[0] = a;
    o[1] = c;
  }
}

template <bool OUT_BF16>
__global__ void __launch_bounds__(NT) gn_apply_kernel(const float* __restrict__ x, int T, int d, int G, int rows_per_cta,
                                                      const float* __restrict__ partial, int nchunk_stats, float eps,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      void* __restrict__ out, float* __restrict__ stats_out) {
  const int b = blockIdx.y;
  __shared__ float sh_mean[64], sh_rstd[64];
  if (threadIdx.x < G) {
    double s = 0.0, ss = 0.0;
    for (int c = 0; c < nchunk_stats; ++c) {
      const float* p = partial + (((long long)b * nchunk_stats + c) * G + threadIdx.x) * 2;
      s += p[0];
      ss += p[1];
    }
    const double n = (double)T * (d / G);
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    sh_mean[threadIdx.x] = (float)mean;
    sh_rstd[threadIdx.x] = rstd;
    if (blockIdx.x == 0 && stats_out != nullptr) {
      stats_out[((long long)b * G + threadIdx.x) * 2] = (float)mean;
      stats_out[((long long)b * G + threadIdx.x) * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  const int tpr = d >> 2, rlanes = NT / tpr;
  const int col = (threadIdx.x % tpr) << 2, rl = threadIdx.x / tpr;
  const int g = col / (d / G);
  const float mean = sh_mean[g], rstd = sh_rstd[g];
  const float4 ga = ld4(gamma + col), be = ld4(beta + col);
  const int t0 = blockIdx.x * rows_per_cta, t1 = min(T, t0 + rows_per_cta);
#pragma unroll 4
  for (int t = t0 + rl; t < t1; t += rlanes) {
    const long long off = ((long long)b * T + t) * d + col;
    float4 v = ld4(x + off);
    v.x = (v.x - mean) * rstd * ga.x + be.x;
    v.y = (v.y - mean) * rstd * ga.y + be.y;
    v.z = (v.z - mean) * rstd * ga.z + be.z;
    v.w = (v.w - mean) * rstd * ga.w + be.w;
    if (OUT_BF16) st4_bf16(reinterpret_cast<bf16*>(out) + off, v);
    else *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + off) = v;
  }
}

// ---------------------------------------------------------------- GroupNorm backward
// partial[b][chunk][d][2] = per-channel (sum dy*xhat, sum dy) over the chunk's rows
template <bool DY_BF16>
__global__ void __launch_bounds__(NT) gn_bwd_stats_kernel(const void* __restrict__ dy, const float* __restrict__ x, int T,
                                                          int d, int G, int rows_per_cta, const float* __restrict__ stats,
                                                          float* __restrict__ partial) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tpr = d >> 2, rlanes = NT / tpr;
  const int col = (threadIdx.x % tpr) << 2, rl = threadIdx.x / tpr;
  const int g = col / (d / G);
  const float mean = stats[((long long)b * G + g) * 2], rstd = stats[((long long)b * G + g) * 2 + 1];
  const int t0 = chunk * rows_per_cta, t1 = min(T, t0 + rows_per_cta);
  float4 a = make_float4(0, 0, 0, 0), c = make_float4(0, 0, 0, 0);
#pragma unroll 4
  for (int t = t0 + rl; t < t1; t += rlanes) {
    const long long off = ((long long)b * T + t) * d + col;
    float4 xv = ld4(x + off);
    float4 g4 = DY_BF16 ? ld4_bf16(reinterpret_cast<const bf16*>(dy) + off) : ld4(reinterpret_cast<const float*>(dy) + off);
    a.x += g4.x * (xv.x - mean) * rstd; a.y += g4.y * (xv.y - mean) * rstd;
    a.z += g4.z * (xv.z - mean) * rstd; a.w += g4.w * (xv.w - mean) * rstd;
    c.x += g4.x; c.y += g4.y; c.z += g4.z; c.w += g4.w;
  }
  __shared__ float4 sh_a[NT], sh_c[NT];
  sh_a[threadIdx.x] = a;
  sh_c[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x < tpr) {
    for (int r = 1; r < rlanes; ++r) {
      float4 a2 = sh_a[r * tpr + threadIdx.x], c2 = sh_c[r * tpr + threadIdx.x];
      a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w;
      c.x += c2.x; c.y += c2.y; c.z += c2.z; c.w += c2.w;
    }
    float* o = partial + (((long long)b * gridDim.x + chunk) * d + col) * 2;
    o[0] = a.x; o[1] = c.x; o[2] = a.y; o[3] = c.y; o[4] = a.z; o[5] = c.z; o[6] = a.w; o[7] = c.w;
  }
}

// dx = rstd * (dy*gamma - S1/n - xhat*S2/n);  dres_out = (accumulate ? dres_in : 0) + dx
template <bool DY_BF16>
__global__ void __launch_bounds__(NT) gn_bwd_apply_kernel(const void* __restrict__ dy, const float* __restrict__ x, int T,
                                                          int d, int G, int rows_per_cta, const float* __restrict__ stats,
                                                          const float* __restrict__ partial, int nchunk_stats,
                                                          const float* __restrict__ gamma, float* __restrict__ dres,
                                                          int accumulate, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, bf16* __restrict__ cast_out,
                                                          float cast_alpha, uint32_t cast_thresh, float cast_inv_keep,
                                                          unsigned long long cast_seed,
                                                          const unsigned long long* __restrict__ seed_ptr) {
  const int b = blockIdx.y;
  if (cast_thresh && seed_ptr) cast_seed += *seed_ptr;
  extern __shared__ float sh_dyn[];  // d floats: a_c ; d floats: c_c ; G: S1 ; G: S2
  float* sh_a = sh_dyn;
  float* sh_c = sh_dyn + d;
  float* sh_s1 = sh_c + d;
  float* sh_s2 = sh_s1 + G;
  for (int ch = threadIdx.x; ch < d; ch += NT) {
    double a = 0.0, c = 0.0;
    for (int k = 0; k < nchunk_stats; ++k) {
      const float* p = partial + (((long long)b * nchunk_stats + k) * d + ch) * 2;
      a += p[0];
      c += p[1];
    }
    sh_a[ch] = (float)a;
    sh_c[ch] = (float)c;
    if (blockIdx.x == 0) {
      if (dgamma) atomicAdd(dgamma + ch, (float)a);
      if (dbeta) atomicAdd(dbeta + ch, (float)c);
    }
  }
  __syncthreads();
  const int cpg = d / G;
  if (threadIdx.x < G) {
    float s1 = 0.f, s2 = 0.f;
    for (int j = 0; j < cpg; ++j) {
      const int ch = threadIdx.x * cpg + j;
      s1 += gamma[ch] * sh_c[ch];
      s2 += gamma[ch] * sh_a[ch];
    }
    const float inv_n = 1.f / ((float)T * cpg);
    sh_s1[threadIdx.x] = s1 * inv_n;
    sh_s2[threadIdx.x] = s2 * inv_n;
  }
  __syncthreads();
  const int tpr = d >> 2, rlanes = NT / tpr;
  const int col = (threadIdx.x % tpr) << 2, rl = threadIdx.x / tpr;
  const int g = col / cpg;
  const float mean = stats[((long long)b * G + g) * 2], rstd = stats[((long long)b * G + g) * 2 + 1];
  const float s1 = sh_s1[g], s2 = sh_s2[g];
  const float4 ga = ld4(gamma + col);
  const int t0 = blockIdx.x * rows_per_cta, t1 = min(T, t0 + rows_per_cta);
#pragma unroll 4
  for (int t = t0 + rl; t < t1; t += rlanes) {
    const long long off = ((long long)b * T + t) * d + col;
    float4 xv = ld4(x + off);
    float4 g4 = DY_BF16 ? ld4_bf16(reinterpret_cast<const bf16*>(dy) + off) : ld4(reinterpret_cast<const float*>(dy) + off);
    float4 r;
    r.x = rstd * (g4.x * ga.x - s1 - (xv.x - mean) * rstd * s2);
    r.y = rstd * (g4.y * ga.y - s1 - (xv.y - mean) * rstd * s2);
    r.z = rstd * (g4.z * ga.z - s1 - (xv.z - mean) * rstd * s2);
    r.w = rstd * (g4.w * ga.w - s1 - (xv.w - mean) * rstd * s2);
    if (accumulate) {
      float4 o = ld4(dres + off);
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    *reinterpret_cast<float4*>(dres + off) = r;
    if (cast_out != nullptr) {  // fused "dy = bf16(alpha * dropout_mask * dres)" for the next backward GEMMs
      float4 c4 = make_float4(r.x * cast_alpha, r.y * cast_alpha, r.z * cast_alpha, r.w * cast_alpha);
      if (cast_thresh) {
        float s0, s1, s2, s3;
        dropout_scale2(cast_seed, (unsigned long long)off, cast_thresh, cast_inv_keep, s0, s1);
        dropout_scale2(cast_seed, (unsigned long long)off + 2, cast_thresh, cast_inv_keep, s2, s3);
        c4.x *= s0; c4.y *= s1; c4.z *= s2; c4.w *= s3;
      }
      st4_bf16(cast_out + off, c4);
    }
  }
}

// ---------------------------------------------------------------- BatchNorm (+SiLU)
// reduce partial[npart][d][2] (sum, sumsq) -> mean / rstd; update running stats (training)
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int npart, int d, long long count, float eps,
                                   float momentum, int training, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
                                   float* __restrict__ stats) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= d) return;
  const int ch = warp;
  if (training) {
    double s = 0.0, ss = 0.0;
    for (int k = lane; k < npart; k += 32) {
      const float* p = partial + ((long long)k * d + ch) * 2;
      s += p[0];
      ss += p[1];
    }
    s = warp_sum_d(s);
    ss = warp_sum_d(ss);
    if (lane == 0) {
      const double mean = s / (double)count;
      double var = ss / (double)count - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[ch * 2] = (float)mean;
      stats[ch * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean) {
        const double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
        running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
      }
      if (ch == 0 && num_batches_tracked) *num_batches_tracked += 1;
    }
  } else if (lane == 0) {
    stats[ch * 2] = running_mean[ch];
    stats[ch * 2 + 1] = rsqrtf(running_var[ch] + eps);
  }
}

// s = silu((w - mean) * rstd * gamma + beta)
// A thread owns 4 fixed channels (its constants live in registers) and walks rows.
__global__ void __launch_bounds__(NT) bn_silu_apply_kernel(const bf16* __restrict__ w, long long M, int d,
                                                           const float* __restrict__ stats, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, bf16* __restrict__ out) {
  const int tpr = d >> 2, rlanes = NT / tpr;
  const int col = (threadIdx.x % tpr) << 2, rl = threadIdx.x / tpr;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float mean = stats[(col + j) * 2], rstd = stats[(col + j) * 2 + 1];
    sc[j] = rstd * gamma[col + j];
    sh[j] = beta[col + j] - mean * sc[j];
  }
  const long long stride = (long long)gridDim.x * rlanes;
#pragma unroll 4
  for (long long r = (long long)blockIdx.x * rlanes + rl; r < M; r += stride) {
    float4 v = ld4_bf16(w + r * d + col);
    v.x = silu_tanh(fmaf(v.x, sc[0], sh[0]));
    v.y = silu_tanh(fmaf(v.y, sc[1], sh[1]));
    v.z = silu_tanh(fmaf(v.z, sc[2], sh[2]));
    v.w = silu_tanh(fmaf(v.w, sc[3], sh[3]));
    st4_bf16(out + r * d + col, v);
  }
}
__global__ void __launch_bounds__(NT) bn_silu_bwd_stats_kernel(const bf16* __restrict__ ds, const bf16* __restrict__ w,
                                                               long long M, int d, int rows_per_cta,
                                                               const float* __restrict__ stats,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* __restrict__ partial) {
  const int tpr = d >> 2, rlanes = NT / tpr;
  const int col = (threadIdx.x % tpr) << 2, rl = threadIdx.x / tpr;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(M, r0 + (long long)rows_per_cta);
  float mean[4], rstd[4], ga[4], be[4];
  for (int j = 0; j < 4; ++j) {
    mean[j] = stats[(col + j) * 2]; rstd[j] = stats[(col + j) * 2 + 1];
    ga[j] = gamma[col + j]; be[j] = beta[col + j];
  }
  float a[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
#pragma unroll 4
  for (long long r = r0 + rl; r < r1; r += rlanes) {
    float4 wv = ld4_bf16(w + r * d + col), dv = ld4_bf16(ds + r * d + col);
    const float wq[4] = {wv.x, wv.y, wv.z, wv.w}, dq[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float wh = (wq[j] - mean[j]) * rstd[j];
      const float dz = dq[j] * silu_grad_tanh(wh * ga[j] + be[j]);
      a[j] += dz * wh;
      c[j] += dz;
    }
  }
  __shared__ float sh[NT][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[threadIdx.x][j] = a[j]; sh[threadIdx.x][4 + j] = c[j]; }
  __syncthreads();
  if (threadIdx.x < tpr) {
    for (int r = 1; r < rlanes; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j] += sh[r * tpr + threadIdx.x][j]; c[j] += sh[r * tpr + threadIdx.x][4 + j]; }
    float* o = partial + ((long long)blockIdx.x * d + col) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = a[j]; o[2 * j + 1] = c[j]; }
  }
}
// reduce partials -> sums[d][2]; accumulate dgamma / dbeta
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int npart, int d, float* __restrict__ sums,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= d) return;
  double a = 0.0, c = 0.0;
  for (int k = lane; k < npart; k += 32) {
    const float* p = partial + ((long long)k * d + warp) * 2;
    a += p[0];
    c += p[1];
  }
  a = warp_sum_d(a);
  c = warp_sum_d(c);
  if (lane == 0) {
    sums[warp * 2] = (float)a;
    sums[warp * 2 + 1] = (float)c;
    if (dgamma) atomicAdd(dgamma + warp, (float)a);
    if (dbeta) atomicAdd(dbeta + warp, (float)c);
  }
}
// dw = gamma*rstd*(dz - C/M - what*A/M)
__global__ void __launch_bounds__(NT) bn_silu_bwd_apply_kernel(const bf16* __restrict__ ds, const bf16* __restrict__ w,
                                                               long long M, int d, const float* __restrict__ stats,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const float* __restrict__ sums, bf16* __restrict__ dw) {
  const int tpr = d >> 2, rlanes = NT / tpr;
  const int col = (threadIdx.x % tpr) << 2, rl = threadIdx.x / tpr;
  const float invM = 1.f / (float)M;
  float mean[4], rstd[4], ga[4], be[4], c1[4], c2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mean[j] = stats[(col + j) * 2]; rstd[j] = stats[(col + j) * 2 + 1];
    ga[j] = gamma[col + j]; be[j] = beta[col + j];
    c2[j] = sums[(col + j) * 2] * invM;
    c1[j] = sums[(col + j) * 2 + 1] * invM;
  }
  const long long stride = (long long)gridDim.x * rlanes;
#pragma unroll 4
  for (long long r = (long long)blockIdx.x * rlanes + rl; r < M; r += stride) {
    const float4 wv = ld4_bf16(w + r * d + col), dv = ld4_bf16(ds + r * d + col);
    const float wq[4] = {wv.x, wv.y, wv.z, wv.w}, dq[4] = {dv.x, dv.y, dv.z, dv.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float wh = (wq[j] - mean[j]) * rstd[j];
      const float dz = dq[j] * silu_grad_tanh(fmaf(wh, ga[j], be[j]));
      o[j] = ga[j] * rstd[j] * (dz - c1[j] - wh * c2[j]);
    }
    st4_bf16(dw + r * d + col, make_float4(o[0], o[1], o[2], o[3]));
  }
}

// ---------------------------------------------------------------- single-pass GroupNorm on a thread-block cluster
// One cluster of GN_CL CTAs per utterance: every CTA stages its slice of rows in shared memory while accumulating the
// per-group sums, the partial sums are exchanged through distributed shared memory, and the rows are normalised
// straight from shared memory: x is read from global memory exactly once (the two-kernel path reads it twice).
constexpr int GN_CL = 8;
constexpr long long GN_FUSED_MAX_SLICE = 8ll << 20;  // bytes of one utterance (T x d fp32) kept hot in L2 between the passes

template <bool OUT_BF16>
__global__ void __launch_bounds__(NT)
gn_fused_fwd_kernel(const float* __restrict__ x, int T, int d, int G, int rows_per_cta, float eps,
                    const float* __restrict__ gamma, const float* __restrict__ beta, void* __restrict__ out,
                    float* __restrict__ stats_out) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float part[64][2];                          // this CTA's per-group (sum, sumsq)
  __shared__ float sh_s[NT], sh_ss[NT];
  __shared__ float sh_mean[64], sh_rstd[64];
  const int ncl = (int)cluster.num_blocks();
  const int b = blockIdx.x / ncl, rank = blockIdx.x % ncl;
  const int tpr = d >> 2, rlanes = NT / tpr;
  const int col = (threadIdx.x % tpr) << 2, rl = threadIdx.x / tpr;
  const int t0 = rank * rows_per_cta, t1 = min(T, t0 + rows_per_cta);
  float s = 0.f, ss = 0.f;
#pragma unroll 4
  for (int t = t0 + rl; t < t1; t += rlanes) {
    const float4 v = ld4(x + ((long long)b * T + t) * d + col);
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  sh_s[threadIdx.x] = s;
  sh_ss[threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < G) {
    const int tpg = (d / G) >> 2;
    float a = 0.f, c = 0.f;
    for (int r = 0; r < rlanes; ++r)
      for (int j = 0; j < tpg; ++j) {
        const int idx = r * tpr + threadIdx.x * tpg + j;
        a += sh_s[idx];
        c += sh_ss[idx];
      }
    part[threadIdx.x][0] = a;
    part[threadIdx.x][1] = c;
  }
  cluster.sync();
  if (threadIdx.x < G) {
    double a = 0.0, c = 0.0;
    for (int r = 0; r < ncl; ++r) {
      const float* rp = cluster.map_shared_rank(&part[0][0], r);
      a += rp[threadIdx.x * 2];
      c += rp[threadIdx.x * 2 + 1];
    }
    const double n = (double)T * (d / G);
    const double mean = a / n;
    double var = c / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    sh_mean[threadIdx.x] = (float)mean;
    sh_rstd[threadIdx.x] = rstd;
    if (rank == 0 && stats_out != nullptr) {
      stats_out[((long long)b * G + threadIdx.x) * 2] = (float)mean;
      stats_out[((long long)b * G + threadIdx.x) * 2 + 1] = rstd;
    }
  }
  cluster.sync();  // also keeps every CTA's `part` alive until all peers have read it
  const int g = col / (d / G);
  const float mean = sh_mean[g], rstd = sh_rstd[g];
  const float4 ga = ld4(gamma + col), be = ld4(beta + col);
#pragma unroll 4
  for (int t = t0 + rl; t < t1; t += rlanes) {
    const long long off = ((long long)b * T + t) * d + col;
    float4 v = ld4(x + off);  // second read of this CTA's slice: served by L2
    v.x = (v.x - mean) * rstd * ga.x + be.x;
    v.y = (v.y - mean) * rstd * ga.y + be.y;
    v.z = (v.z - mean) * rstd * ga.z + be.z;
    v.w = (v.w - mean) * rstd * ga.w + be.w;
    if (OUT_BF16) st4_bf16(reinterpret_cast<bf16*>(out) + off, v);
    else *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + off) = v;
  }
}

// Same kernel with the CTA's rows held on chip between the statistics and the normalisation (at most RREG + RSH rows
// per thread: the first RREG stay in registers, the others in shared memory): all loads of the slice are in flight at
// once and x is not read a second time, not even from L2.  Same summation order as gn_fused_fwd_kernel, so the results
// are bit-identical.
template <bool OUT_BF16, int RREG, int RSH>
__global__ void __launch_bounds__(NT, 4)
gn_fused_fwd_reg_kernel(const float* __restrict__ x, int T, int d, int G, int rows_per_cta, float eps,
                        const float* __restrict__ gamma, const float* __restrict__ beta, void* __restrict__ out,
                        float* __restrict__ stats_out) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float part[64][2];
  __shared__ float sh_s[NT], sh_ss[NT];
  __shared__ float sh_mean[64], sh_rstd[64];
  __shared__ float4 stash[RSH][NT];
  const int ncl = (int)cluster.num_blocks();
  const int b = blockIdx.x / ncl, rank = blockIdx.x % ncl;
  const int tpr = d >> 2, rlanes = NT / tpr;
  const int col = (threadIdx.x % tpr) << 2, rl = threadIdx.x / tpr;
  const int t0 = rank * rows_per_cta, t1 = min(T, t0 + rows_per_cta);
  const long long base = ((long long)b * T + t0 + rl) * d + col;
  const long long rstride = (long long)rlanes * d;
  // rows RREG.. go straight to shared memory (cp.async: no registers held while they are in flight)
#pragma unroll
  for (int i = 0; i < RSH; ++i)
    if (t0 + rl + (RREG + i) * rlanes < t1)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&stash[i][threadIdx.x])),
                   "l"(x + base + (RREG + i) * rstride)
                   : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  float4 v[RREG];
#pragma unroll
  for (int i = 0; i < RREG; ++i) {
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t0 + rl + i * rlanes < t1) v[i] = ld4(x + base + i * rstride);
  }
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < RREG; ++i) {
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // a thread only ever reads the stash entries it fetched itself
#pragma unroll
  for (int i = 0; i < RSH; ++i)
    if (t0 + rl + (RREG + i) * rlanes < t1) {
      const float4 h = stash[i][threadIdx.x];
      s += (h.x + h.y) + (h.z + h.w);
      ss += (h.x * h.x + h.y * h.y) + (h.z * h.z + h.w * h.w);
    }
  sh_s[threadIdx.x] = s;
  sh_ss[threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x < G) {
    const int tpg = (d / G) >> 2;
    float a = 0.f, c = 0.f;
    for (int r = 0; r < rlanes; ++r)
      for (int j = 0; j < tpg; ++j) {
        const int idx = r * tpr + threadIdx.x * tpg + j;
        a += sh_s[idx];
        c += sh_ss[idx];
      }
    part[threadIdx.x][0] = a;
    part[threadIdx.x][1] = c;
  }
  cluster.sync();
  if (threadIdx.x < G) {
    double a = 0.0, c = 0.0;
    for (int r = 0; r < ncl; ++r) {
      const float* rp = cluster.map_shared_rank(&part[0][0], r);
      a += rp[threadIdx.x * 2];
      c += rp[threadIdx.x * 2 + 1];
    }
    const double n = (double)T * (d / G);
    const double mean = a / n;
    double var = c / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    sh_mean[threadIdx.x] = (float)mean;
    sh_rstd[threadIdx.x] = rstd;
    if (rank == 0 && stats_out != nullptr) {
      stats_out[((long long)b * G + threadIdx.x) * 2] = (float)mean;
      stats_out[((long long)b * G + threadIdx.x) * 2 + 1] = rstd;
    }
  }
  cluster.sync();  // also keeps every CTA's `part` alive until all peers have read it
  const int g = col / (d / G);
  const float mean = sh_mean[g], rstd = sh_rstd[g];
  const float4 ga = ld4(gamma + col), be = ld4(beta + col);
#pragma unroll
  for (int i = 0; i < RREG + RSH; ++i) {
    if (t0 + rl + i * rlanes < t1) {
      const long long off = base + i * rstride;
      float4 w = i < RREG ? v[i < RREG ? i : 0] : stash[i < RREG ? 0 : i - RREG][threadIdx.x];  // own values: no barrier needed
      w.x = (w.x - mean) * rstd * ga.x + be.x;
      w.y = (w.y - mean) * rstd * ga.y + be.y;
      w.z = (w.z - mean) * rstd * ga.z + be.z;
      w.w = (w.w - mean) * rstd * ga.w + be.w;
      if (OUT_BF16) st4_bf16(reinterpret_cast<bf16*>(out) + off, w);
      else *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + off) = w;
    }
  }
}

// backward: dx = rstd * (dy*gamma - S1/n - xhat*S2/n); per-channel sums of dy*xhat and dy exchanged through DSMEM
template <bool DY_BF16>
__global__ void __launch_bounds__(NT, 4)
gn_fused_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x, int T, int d, int G, int rows_per_cta,
                    const float* __restrict__ stats, const float* __restrict__ gamma, float* __restrict__ dres,
                    int accumulate, float* __restrict__ dgamma, float* __restrict__ dbeta, bf16* __restrict__ cast_out,
                    float cast_alpha, uint32_t cast_thresh, float cast_inv_keep, unsigned long long cast_seed,
                    const unsigned long long* __restrict__ seed_ptr) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sh_dyn2[];
  // layout: part a (d) | part c (d) | tot a (d) | tot c (d) | S1 (G) | S2 (G)
  const int ncl = (int)cluster.num_blocks();
  const int b = blockIdx.x / ncl, rank = blockIdx.x % ncl;
  float* pa = sh_dyn2;
  float* pc = pa + d;
  float* ta = pc + d;
  float* tcx = ta + d;
  float* s1s = tcx + d;
  float* s2s = s1s + G;
  __shared__ float4 red_a[NT], red_c[NT];
  if (cast_thresh && seed_ptr) cast_seed += *seed_ptr;
  const int tpr = d >> 2, rlanes = NT / tpr;
  const int col = (threadIdx.x % tpr) << 2, rl = threadIdx.x / tpr;
  const int cpg = d / G;
  const int g = col / cpg;
  const float mean = stats[((long long)b * G + g) * 2], rstd = stats[((long long)b * G + g) * 2 + 1];
  const int t0 = rank * rows_per_cta, t1 = min(T, t0 + rows_per_cta);
  float4 a = make_float4(0, 0, 0, 0), c = make_float4(0, 0, 0, 0);
#pragma unroll 4
  for (int t = t0 + rl; t < t1; t += rlanes) {
    const long long off = ((long long)b * T + t) * d + col;
    float4 xv = ld4(x + off);
    const float4 g4 = DY_BF16 ? ld4_bf16(reinterpret_cast<const bf16*>(dy) + off) : ld4(reinterpret_cast<const float*>(dy) + off);
    xv.x = (xv.x - mean) * rstd; xv.y = (xv.y - mean) * rstd; xv.z = (xv.z - mean) * rstd; xv.w = (xv.w - mean) * rstd;
    a.x += g4.x * xv.x; a.y += g4.y * xv.y; a.z += g4.z * xv.z; a.w += g4.w * xv.w;
    c.x += g4.x; c.y += g4.y; c.z += g4.z; c.w += g4.w;
  }
  red_a[threadIdx.x] = a;
  red_c[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x < tpr) {
    for (int r = 1; r < rlanes; ++r) {
      const float4 a2 = red_a[r * tpr + threadIdx.x], c2 = red_c[r * tpr + threadIdx.x];
      a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w;
      c.x += c2.x; c.y += c2.y; c.z += c2.z; c.w += c2.w;
    }
    *reinterpret_cast<float4*>(pa + col) = a;
    *reinterpret_cast<float4*>(pc + col) = c;
  }
  cluster.sync();
  for (int ch = threadIdx.x; ch < d; ch += NT) {
    float av = 0.f, cv = 0.f;
    for (int r = 0; r < ncl; ++r) {
      av += cluster.map_shared_rank(pa, r)[ch];
      cv += cluster.map_shared_rank(pc, r)[ch];
    }
    ta[ch] = av;
    tcx[ch] = cv;
    if (rank == 0) {
      if (dgamma) atomicAdd(dgamma + ch, av);
      if (dbeta) atomicAdd(dbeta + ch, cv);
    }
  }
  __syncthreads();
  if (threadIdx.x < G) {
    float s1 = 0.f, s2 = 0.f;
    for (int j = 0; j < cpg; ++j) {
      const int ch = threadIdx.x * cpg + j;
      s1 += gamma[ch] * tcx[ch];
      s2 += gamma[ch] * ta[ch];
    }
    const float inv_n = 1.f / ((float)T * cpg);
    s1s[threadIdx.x] = s1 * inv_n;
    s2s[threadIdx.x] = s2 * inv_n;
  }
  cluster.sync();  // peers have finished reading this CTA's partial sums; S1/S2 visible to the whole CTA
  const float s1 = s1s[g], s2 = s2s[g];
  const float4 ga = ld4(gamma + col);
#pragma unroll 4
  for (int t = t0 + rl; t < t1; t += rlanes) {
    const long long off = ((long long)b * T + t) * d + col;
    float4 xh = ld4(x + off);  // second read of this CTA's slice of x and dy: served by L2
    const float4 g4 = DY_BF16 ? ld4_bf16(reinterpret_cast<const bf16*>(dy) + off) : ld4(reinterpret_cast<const float*>(dy) + off);
    xh.x = (xh.x - mean) * rstd; xh.y = (xh.y - mean) * rstd; xh.z = (xh.z - mean) * rstd; xh.w = (xh.w - mean) * rstd;
    float4 r;
    r.x = rstd * (g4.x * ga.x - s1 - xh.x * s2);
    r.y = rstd * (g4.y * ga.y - s1 - xh.y * s2);
    r.z = rstd * (g4.z * ga.z - s1 - xh.z * s2);
    r.w = rstd * (g4.w * ga.w - s1 - xh.w * s2);
    if (accumulate) {
      const float4 o = ld4(dres + off);
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    *reinterpret_cast<float4*>(dres + off) = r;
    if (cast_out != nullptr) {
      float4 c4 = make_float4(r.x * cast_alpha, r.y * cast_alpha, r.z * cast_alpha, r.w * cast_alpha);
      if (cast_thresh) {
        float q0, q1, q2, q3;
        dropout_scale2(cast_seed, (unsigned long long)off, cast_thresh, cast_inv_keep, q0, q1);
        dropout_scale2(cast_seed, (unsigned long long)off + 2, cast_thresh, cast_inv_keep, q2, q3);
        c4.x *= q0; c4.y *= q1; c4.z *= q2; c4.w *= q3;
      }
      st4_bf16(cast_out + off, c4);
    }
  }
}

// cluster size: as many CTAs per utterance as keep the whole grid resident (ctas_per_sm CTAs fit an SM)
int gn_cluster_size(int B, int T, int ctas_per_sm) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int cl = GN_CL;
  while (cl > 1 && ((long long)B * cl > (long long)ctas_per_sm * sms || cl > T)) cl >>= 1;
  return cl;
}
template <typename... KArgs, typename... Args>
cudaError_t gn_launch_cluster(void (*kern)(KArgs...), int B, int cl, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * cl);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

bool gn_shape_ok(int d, int G) {
  if (d <= 0 || G <= 0 || G > 64 || d % G) return false;
  const int cpg = d / G;
  if (cpg % 4) return false;
  const int tpr = d / 4;
  return tpr <= NT && (NT % tpr) == 0;
}
int gn_rows_per_cta(int B, int T) {
  int chunks = max(1, 1184 / max(B, 1));  // ~8 CTAs (of 8 warps) per SM: these kernels are latency bound otherwise
  int rows = cdiv(T, chunks);
  return max(rows, 8);
}

}  // namespace

extern "C" int tasr_groupnorm_chunks(int B, int T) { return cdiv(T, gn_rows_per_cta(B, T)); }

extern "C" size_t tasr_groupnorm_workspace_bytes(int B, int T, int d) {
  return (size_t)B * tasr_groupnorm_chunks(B, T) * d * 2 * sizeof(float);
}

extern "C" int tasr_groupnorm_fwd(const float* x, int B, int T, int d, int G, float eps, const float* gamma,
                                  const float* beta, void* out, int out_bf16, float* stats, void* workspace,
                                  size_t workspace_bytes, tasr_stream_t stream) {
  if (!gn_shape_ok(d, G) || B <= 0 || T <= 0) return TASR_ERR_SHAPE;
  if (workspace_bytes < tasr_groupnorm_workspace_bytes(B, T, d)) return TASR_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((long long)T * d * sizeof(float) <= GN_FUSED_MAX_SLICE) {  // one cluster per utterance; its slice is re-read through L2
    const int cl = gn_cluster_size(B, T, 6);
    const int frows = cdiv(T, cl);
    // rows of a CTA's slice per thread; up to GN_REG_ROWS of them stay on chip (TASR_GN_REG=0: always re-read from L2)
    constexpr int GN_RREG = 4, GN_RSH = 8, GN_REG_ROWS = GN_RREG + GN_RSH;
    static const bool reg_ok = [] { const char* v = getenv("TASR_GN_REG"); return !(v && v[0] == '0'); }();
    if (reg_ok && d <= 4 * NT && NT % (d / 4) == 0 && cdiv(frows, NT / (d / 4)) <= GN_REG_ROWS) {
      cudaError_t e = out_bf16 ? gn_launch_cluster(gn_fused_fwd_reg_kernel<true, GN_RREG, GN_RSH>, B, cl, 0, st, x, T, d, G, frows, eps,
                                                   gamma, beta, out, stats)
                               : gn_launch_cluster(gn_fused_fwd_reg_kernel<false, GN_RREG, GN_RSH>, B, cl, 0, st, x, T, d, G, frows, eps,
                                                   gamma, beta, out, stats);
      if (e != cudaSuccess) return tasr_set_cuda_error(e);
      TASR_CHECK_LAUNCH();
      return TASR_OK;
    }
    cudaError_t e = out_bf16 ? gn_launch_cluster(gn_fused_fwd_kernel<true>, B, cl, 0, st, x, T, d, G, frows, eps, gamma, beta, out, stats)
                             : gn_launch_cluster(gn_fused_fwd_kernel<false>, B, cl, 0, st, x, T, d, G, frows, eps, gamma, beta, out, stats);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    TASR_CHECK_LAUNCH();
    return TASR_OK;
  }
  const int rows = gn_rows_per_cta(B, T), nchunk = cdiv(T, rows);
  float* partial = reinterpret_cast<float*>(workspace);
  dim3 grid(nchunk, B);
  gn_stats_kernel<<<grid, NT, 0, st>>>(x, T, d, G, rows, partial);
  TASR_CHECK_LAUNCH();
  if (out_bf16) gn_apply_kernel<true><<<grid, NT, 0, st>>>(x, T, d, G, rows, partial, nchunk, eps, gamma, beta, out, stats);
  else gn_apply_kernel<false><<<grid, NT, 0, st>>>(x, T, d, G, rows, partial, nchunk, eps, gamma, beta, out, stats);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_groupnorm_bwd(const void* dy, int dy_bf16, const float* x, int B, int T, int d, int G,
                                  const float* stats, const float* gamma, float* dres, int accumulate, float* dgamma,
                                  float* dbeta, void* cast_out, float cast_alpha, float cast_drop_p, uint64_t cast_seed,
                                  void* workspace, size_t workspace_bytes, tasr_stream_t stream) {
  const uint32_t cthresh = tasr_drop_thresh16(cast_drop_p);
  const float cinv = tasr_drop_inv_keep(cthresh);
  bf16* cout_ = reinterpret_cast<bf16*>(cast_out);
  if (!gn_shape_ok(d, G) || B <= 0 || T <= 0) return TASR_ERR_SHAPE;
  if (workspace_bytes < tasr_groupnorm_workspace_bytes(B, T, d)) return TASR_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((long long)T * d * sizeof(float) <= GN_FUSED_MAX_SLICE) {
    // up to 4 CTAs per SM (64 registers): 24 % faster than 2 per SM stand-alone on cold operands (more loads in flight);
    // inside the training step, where the kernel shares the memory system with the weight-gradient stream, +0.3 %
    const int cl = gn_cluster_size(B, T, 4);
    const int frows = cdiv(T, cl);
    const size_t fsm = ((size_t)4 * d + 2 * G) * sizeof(float);
    const unsigned long long seed64 = cast_seed;
    cudaError_t e = dy_bf16 ? gn_launch_cluster(gn_fused_bwd_kernel<true>, B, cl, fsm, st, dy, x, T, d, G, frows, stats, gamma, dres,
                                                accumulate, dgamma, dbeta, cout_, cast_alpha, cthresh, cinv, seed64, g_tasr_seed_ptr)
                            : gn_launch_cluster(gn_fused_bwd_kernel<false>, B, cl, fsm, st, dy, x, T, d, G, frows, stats, gamma, dres,
                                                accumulate, dgamma, dbeta, cout_, cast_alpha, cthresh, cinv, seed64, g_tasr_seed_ptr);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    TASR_CHECK_LAUNCH();
    return TASR_OK;
  }
  const int rows = gn_rows_per_cta(B, T), nchunk = cdiv(T, rows);
  float* partial = reinterpret_cast<float*>(workspace);
  dim3 grid(nchunk, B);
  const size_t sm = (size_t)(2 * d + 2 * G) * sizeof(float);
  if (dy_bf16) {
    gn_bwd_stats_kernel<true><<<grid, NT, 0, st>>>(dy, x, T, d, G, rows, stats, partial);
    TASR_CHECK_LAUNCH();
    gn_bwd_apply_kernel<true><<<grid, NT, sm, st>>>(dy, x, T, d, G, rows, stats, partial, nchunk, gamma, dres, accumulate,
                                                    dgamma, dbeta, cout_, cast_alpha, cthresh, cinv, cast_seed, g_tasr_seed_ptr);
  } else {
    gn_bwd_stats_kernel<false><<<grid, NT, 0, st>>>(dy, x, T, d, G, rows, stats, partial);
    TASR_CHECK_LAUNCH();
    gn_bwd_apply_kernel<false><<<grid, NT, sm, st>>>(dy, x, T, d, G, rows, stats, partial, nchunk, gamma, dres, accumulate,
                                                     dgamma, dbeta, cout_, cast_alpha, cthresh, cinv, cast_seed, g_tasr_seed_ptr);
  }
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

// ---- BatchNorm + SiLU
extern "C" int tasr_bn_finalize(const float* partial, int npart, int d, int64_t count, float eps, float momentum,
                                int training, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                float* stats, tasr_stream_t stream) {
  if (d <= 0 || (training && npart <= 0)) return TASR_ERR_SHAPE;
  bn_finalize_kernel<<<cdiv((long long)d * 32, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, npart, d, count, eps, momentum, training, running_mean, running_var,
      reinterpret_cast<long long*>(num_batches_tracked), stats);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_bn_silu_fwd(const void* w, int64_t M, int d, const float* stats, const float* gamma,
                                const float* beta, void* out, tasr_stream_t stream) {
  const int tpr = d / 4;
  if (d % 4 || tpr > NT || (NT % tpr) || M <= 0) return TASR_ERR_SHAPE;
  const int grid = (int)imin64((long long)148 * 6, (M + (NT / tpr) - 1) / (NT / tpr));
  bn_silu_apply_kernel<<<grid, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(w), M, d, stats, gamma, beta, reinterpret_cast<bf16*>(out));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_bn_bwd_parts(int64_t M) { return (int)imin64((long long)296, (long long)((M + 31) / 32)); }

extern "C" size_t tasr_bn_bwd_workspace_bytes(int64_t M, int d) {
  return ((size_t)tasr_bn_bwd_parts(M) * d * 2 + (size_t)d * 2) * sizeof(float);
}

extern "C" int tasr_bn_silu_bwd(const void* ds, const void* w, int64_t M, int d, const float* stats,
                                const float* gamma, const float* beta, void* dw, float* dgamma, float* dbeta,
                                void* workspace, size_t workspace_bytes, tasr_stream_t stream) {
  const int tpr = d / 4;
  if (d % 4 || tpr > NT || (NT % tpr) || M <= 0) return TASR_ERR_SHAPE;
  if (workspace_bytes < tasr_bn_bwd_workspace_bytes(M, d)) return TASR_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nparts_max = tasr_bn_bwd_parts(M);
  const int rows = (int)((M + nparts_max - 1) / nparts_max);
  const int nparts = (int)((M + rows - 1) / rows);
  float* partial = reinterpret_cast<float*>(workspace);
  float* sums = partial + (size_t)nparts_max * d * 2;
  bn_silu_bwd_stats_kernel<<<nparts, NT, 0, st>>>(reinterpret_cast<const bf16*>(ds), reinterpret_cast<const bf16*>(w), M, d,
                                                  rows, stats, gamma, beta, partial);
  TASR_CHECK_LAUNCH();
  bn_bwd_finalize_kernel<<<cdiv((long long)d * 32, 256), 256, 0, st>>>(partial, nparts, d, sums, dgamma, dbeta);
  TASR_CHECK_LAUNCH();
  const int grid = (int)imin64((long long)148 * 6, (M + (NT / tpr) - 1) / (NT / tpr));
  bn_silu_bwd_apply_kernel<<<grid, NT, 0, st>>>(reinterpret_cast<const bf16*>(ds), reinterpret_cast<const bf16*>(w), M, d,
                                                stats, gamma, beta, sums, reinterpret_cast<bf16*>(dw));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
