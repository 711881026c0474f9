// Small HBM-bound helpers: fp32 -> bf16 casts (with scale / dropout mask), column sums (bias
// gradients), rotary position embedding on the fused QKV buffer.
// Replaces (reference): autocast weight/activation casts; dropout backward of model/conformer.py:23,25;
//   bias-gradient reductions of nn.Linear backward; model/attention.py:62-70,228-230 (rotate_half /
//   apply_rotary_pos_emb: 2 x (chunk, neg, cat, mul, mul, add)).
#include "common.cuh"

namespace {

constexpr int NT = 256;

__global__ void __launch_bounds__(NT) cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out,
                                                           long long n, float alpha, uint32_t thresh, float inv_keep,
                                                           unsigned long long seed, const unsigned long long* seed_ptr) {
  if (thresh && seed_ptr) seed += *seed_ptr;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
    float4 v = *reinterpret_cast<const float4*>(in + i * 4);
    v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
    if (thresh) {
      float s0, s1, s2, s3;
      dropout_scale2(seed, (unsigned long long)i * 4, thresh, inv_keep, s0, s1);
      dropout_scale2(seed, (unsigned long long)i * 4 + 2, thresh, inv_keep, s2, s3);
      v.x *= s0; v.y *= s1; v.z *= s2; v.w *= s3;
    }
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + i * 4) = u;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float v = in[i] * alpha;
    if (thresh) v *= dropout_scale(seed, i, thresh, inv_keep);
    out[i] = __float2bfloat16(v);
  }
}

// out[c] += sum_r in[r][c]   (in: (M, N) bf16, row pitch ld).  blockDim = (32, 8): a warp covers 256 columns
// (16 B per lane), 8 row lanes, 4 independent rows in flight per thread.
__global__ void __launch_bounds__(NT) colsum_bf16_kernel(const bf16* __restrict__ in, long long M, int N, long long ld,
                                                         int rows_per_cta, float* __restrict__ out) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 8;  // first of this thread's 8 columns
  const long long r0 = (long long)blockIdx.y * rows_per_cta, r1 = min(M, r0 + (long long)rows_per_cta);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (c + 8 <= N && (ld & 7) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    long long r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(in + (r + 8 * k) * ld + c);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f;
        f = unpack_bf16x2(u[k].x); acc[0] += f.x; acc[1] += f.y;
        f = unpack_bf16x2(u[k].y); acc[2] += f.x; acc[3] += f.y;
        f = unpack_bf16x2(u[k].z); acc[4] += f.x; acc[5] += f.y;
        f = unpack_bf16x2(u[k].w); acc[6] += f.x; acc[7] += f.y;
      }
    }
    for (; r < r1; r += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(in + r * ld + c);
      float2 f;
      f = unpack_bf16x2(u.x); acc[0] += f.x; acc[1] += f.y;
      f = unpack_bf16x2(u.y); acc[2] += f.x; acc[3] += f.y;
      f = unpack_bf16x2(u.z); acc[4] += f.x; acc[5] += f.y;
      f = unpack_bf16x2(u.w); acc[6] += f.x; acc[7] += f.y;
    }
  } else {
    for (long long r = r0 + threadIdx.y; r < r1; r += 8)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c + i < N) acc[i] += __bfloat162float(in[r * ld + c + i]);
  }
  __shared__ float sh[8][32][9];
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  // 256 threads: one column each
  const int t = threadIdx.y * 32 + threadIdx.x;
  const int lane_x = t >> 3, i = t & 7;
  float s = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) s += sh[y][lane_x][i];
  const int col = (blockIdx.x * 32 + lane_x) * 8 + i;
  if (col < N) atomicAdd(out + col, s);
}

// In-place NeoX-style RoPE on the first `rot_cols` columns of each row (blocks of 64: H query heads and
// the shared key head).  x' = x cos + rotate_half(x) sin; `sign` = -1 applies the inverse rotation
// (used for the gradient).  cs: (T, 32, 2) fp32 = (cos, sin) per position and frequency.
__global__ void __launch_bounds__(NT) rope_kernel(bf16* __restrict__ qkv, long long M, int T, int ld, int rot_cols,
                                                  const float* __restrict__ cs, float sign) {
  const int pairs_per_row = rot_cols >> 1;  // (i, i+32) pairs: 32 per head
  const long long total = M * pairs_per_row;
  for (long long idx = (long long)blockIdx.x * NT + threadIdx.x; idx < total; idx += (long long)gridDim.x * NT) {
    const long long row = idx / pairs_per_row;
    const int p = (int)(idx - row * pairs_per_row);
    const int head = p >> 5, i = p & 31;
    const int t = (int)(row % T);
    const float c = cs[(t * 32 + i) * 2], s = sign * cs[(t * 32 + i) * 2 + 1];
    bf16* base = qkv + row * ld + head * 64;
    const float x1 = __bfloat162float(base[i]), x2 = __bfloat162float(base[i + 32]);
    base[i] = __float2bfloat16(x1 * c - x2 * s);
    base[i + 32] = __float2bfloat16(x2 * c + x1 * s);
  }
}

// out[n][(k % q) * (K / q) + k / q] = bf16(in[n][k])      (conv2: q = 9, input_proj: q = F2)
__global__ void __launch_bounds__(NT) pack_weight_remap_kernel(const float* __restrict__ in, long long N, int K, int q,
                                                               bf16* __restrict__ out) {
  const int inner = K / q;
  const long long total = N * K;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
    const long long n = i / K;
    const int k = (int)(i - n * K);
    out[n * K + (long long)(k % q) * inner + k / q] = __float2bfloat16(in[i]);
  }
}

}  // namespace

extern "C" int tasr_cast_f32_bf16(const float* in, void* out, int64_t n, float alpha, float drop_p, uint64_t seed,
                                  tasr_stream_t stream) {
  if (n <= 0) return TASR_OK;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 7)) return TASR_ERR_ALIGN;
  const uint32_t thresh = tasr_drop_thresh16(drop_p);
  const float inv_keep = tasr_drop_inv_keep(thresh);
  const int grid = (int)imin64((long long)148 * 8, ((n >> 2) + NT) / NT);
  cast_f32_bf16_kernel<<<grid, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, reinterpret_cast<bf16*>(out), n, alpha,
                                                                                 thresh, inv_keep, seed, g_tasr_seed_ptr);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_colsum_bf16(const void* in, int64_t M, int N, int64_t ld, float* out, tasr_stream_t stream) {
  if (M <= 0 || N <= 0 || (ld & 1)) return TASR_ERR_SHAPE;
  const int col_blocks = cdiv(N, 256);
  int row_blocks = (int)((M + 127) / 128);
  if (row_blocks > 888 / col_blocks) row_blocks = 888 / col_blocks;
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_cta = (int)((M + row_blocks - 1) / row_blocks);
  row_blocks = (int)((M + rows_per_cta - 1) / rows_per_cta);
  dim3 grid(col_blocks, row_blocks), block(32, 8);
  colsum_bf16_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(in), M, N,
                                                                                 ld, rows_per_cta, out);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_rope_inplace(void* qkv, int64_t M, int T, int ld, int rot_cols, const float* cos_sin, int inverse,
                                 tasr_stream_t stream) {
  if (rot_cols % 64 || M <= 0 || T <= 0) return TASR_ERR_SHAPE;
  const long long total = M * (rot_cols / 2);
  const int grid = (int)imin64((long long)148 * 8, (total + NT - 1) / NT);
  rope_kernel<<<grid, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<bf16*>(qkv), M, T, ld, rot_cols,
                                                                       cos_sin, inverse ? -1.f : 1.f);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_pack_weight_remap(const float* in, int64_t N, int K, int q, void* out, tasr_stream_t stream) {
  if (N <= 0 || K <= 0 || q <= 0 || K % q) return TASR_ERR_SHAPE;
  const int grid = (int)imin64((long long)148 * 8, (N * K + NT - 1) / NT);
  pack_weight_remap_kernel<<<grid, NT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, N, K, q, reinterpret_cast<bf16*>(out));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
