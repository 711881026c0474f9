// Global-norm gradient clipping + AdamW on flat fp32 buffers, one fused multi-tensor pass that also
// refreshes the bf16 shadow copy of the weights used by the tcgen05 GEMMs.
// Replaces (reference): trainer/trainer.py:189-195 (clip_grad_norm_(params, 1.0) + AdamW step:
//   foreach norm / mul / addcdiv kernels), main.py:106-110 (AdamW lr 5e-4, wd 1e-6).
#include "common.cuh"

namespace {
constexpr int NT = 256;

__global__ void __launch_bounds__(NT) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  double acc = 0.0;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
    const float4 v = *reinterpret_cast<const float4*>(g + i * 4);
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    acc += (double)v * v;
  }
  acc = warp_sum_d(acc);
  __shared__ double sh[NT / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < NT / 32; ++i) t += sh[i];
    atomicAdd(out, t);
  }
}

// hyper: [lr, beta1, beta2, eps, weight_decay, bias_corr1, bias_corr2, max_norm (<=0: no clipping), grad_div]
__global__ void __launch_bounds__(NT) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, bf16* __restrict__ shadow, long long n,
                                                   const float* __restrict__ hyper, const double* __restrict__ sumsq,
                                                   float* __restrict__ norm_out) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float bc1 = hyper[5], bc2 = hyper[6], max_norm = hyper[7], gdiv = hyper[8];
  float coef = 1.f / gdiv;
  bool skip = false;
  if (sumsq != nullptr) {
    const float total = (float)sqrt(*sumsq) / gdiv;
    if (norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
    if (!(total == total) || total == INFINITY) skip = true;          // non-finite gradients: leave the weights alone
    if (max_norm > 0.f) coef *= fminf(1.f, max_norm / (total + 1e-6f));
  }
  if (skip) return;
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    pi -= step_size * mi / denom;
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (shadow != nullptr) shadow[i] = __float2bfloat16(pi);
  }
}
}  // namespace

// out (double, device) += sum(g^2); zero it first with tasr_zero or cudaMemsetAsync
extern "C" int tasr_grad_sumsq(const float* g, int64_t n, double* out, tasr_stream_t stream) {
  if (n <= 0) return TASR_OK;
  if (reinterpret_cast<uintptr_t>(g) & 15) return TASR_ERR_ALIGN;
  const int grid = (int)imin64((long long)148 * 4, ((n >> 2) + NT) / NT);
  sumsq_kernel<<<grid, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g, n, out);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_clip_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, int64_t n,
                               const float* hyper, const double* sumsq, float* norm_out, tasr_stream_t stream) {
  if (n <= 0) return TASR_OK;
  const int grid = (int)imin64((long long)148 * 8, (n + NT - 1) / NT);
  adamw_kernel<<<grid, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, reinterpret_cast<bf16*>(shadow_bf16), n,
                                                                        hyper, sumsq, norm_out);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
