// Fused log-mel front-end: reflect-pad framing + Hann window + 400-point real FFT + power + mel
// filterbank + 10*log10 in ONE kernel (frames staged through shared memory), followed by the
// per-utterance top_db clamp and CMVN.
// Replaces (reference): data/preprocessing.py:52-64 (MelSpectrogram/AmplitudeToDB ctor),
//   :98 (mel_transform -> torch.stft + MelScale), :101 (amplitude_to_db, top_db=80),
//   :112-116 (_normalize: per-bin mean / unbiased std over time, eps 1e-8).
//
// 400-point real FFT = 200-point complex FFT of z[n] = x[2n] + i x[2n+1] (200 = 8 x 5 x 5, one radix-8
// and two radix-5 passes in shared memory) + the real-FFT split.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int NFFT = 400;
constexpr int HOP = 160;
constexpr int NBIN = 201;
constexpr int FR = 32;         // frames per CTA
constexpr int MEL_NT = 512;           // two CTAs (104 KB each) per SM -> 1024 threads per SM
constexpr int RAW = (FR - 1) * HOP + NFFT;  // 5360 samples staged per CTA
constexpr int ZLD = 201;       // float2 per frame (200 + 1 pad)
constexpr int PLD = 203;       // floats per frame of the power spectrum

__device__ float2 g_tw400[NFFT];  // e^{-2 pi i j / 400}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_negi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

__device__ __forceinline__ void dft4(float2 b0, float2 b1, float2 b2, float2 b3, float2& y0, float2& y1, float2& y2,
                                     float2& y3) {
  float2 c0 = cadd(b0, b2), c1 = csub(b0, b2), c2 = cadd(b1, b3), c3 = mul_negi(csub(b1, b3));
  y0 = cadd(c0, c2); y2 = csub(c0, c2); y1 = cadd(c1, c3); y3 = csub(c1, c3);
}
// forward 8-point DFT, natural order in and out
__device__ __forceinline__ void dft8(float2* x) {
  const float h = 0.70710678118654752f;
  float2 a0 = cadd(x[0], x[4]), a4 = csub(x[0], x[4]);
  float2 a1 = cadd(x[1], x[5]), a5 = csub(x[1], x[5]);
  float2 a2 = cadd(x[2], x[6]), a6 = csub(x[2], x[6]);
  float2 a3 = cadd(x[3], x[7]), a7 = csub(x[3], x[7]);
  a5 = make_float2(h * (a5.x + a5.y), h * (a5.y - a5.x));    // * (h, -h)
  a6 = mul_negi(a6);                                         // * (-i)
  a7 = make_float2(h * (a7.y - a7.x), -h * (a7.x + a7.y));   // * (-h, -h)
  dft4(a0, a1, a2, a3, x[0], x[2], x[4], x[6]);
  dft4(a4, a5, a6, a7, x[1], x[3], x[5], x[7]);
}
// forward 5-point DFT
__device__ __forceinline__ void dft5(float2* x) {
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494745f;
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247314f;
  float2 t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
  float2 m1 = make_float2(x[0].x + c1 * t1.x + c2 * t2.x, x[0].y + c1 * t1.y + c2 * t2.y);
  float2 m2 = make_float2(x[0].x + c2 * t1.x + c1 * t2.x, x[0].y + c2 * t1.y + c1 * t2.y);
  float2 u1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  float2 u2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  x[0] = make_float2(x[0].x + t1.x + t2.x, x[0].y + t1.y + t2.y);
  // X1 = m1 - i u1, X4 = m1 + i u1, X2 = m2 - i u2, X3 = m2 + i u2
  x[1] = make_float2(m1.x + u1.y, m1.y - u1.x);
  x[4] = make_float2(m1.x - u1.y, m1.y + u1.x);
  x[2] = make_float2(m2.x + u2.y, m2.y - u2.x);
  x[3] = make_float2(m2.x - u2.y, m2.y + u2.x);
}

// nonzero bin range of every mel filter (called once per filterbank)
__global__ void mel_filter_ranges_kernel(const float* __restrict__ fb, int n_mels, int* __restrict__ ranges) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_mels) return;
  int lo = NBIN, hi = 0;
  for (int k = 0; k < NBIN; ++k)
    if (fb[k * n_mels + m] != 0.f) {
      lo = min(lo, k);
      hi = max(hi, k + 1);
    }
  if (hi == 0) lo = 0;
  ranges[m] = lo;
  ranges[n_mels + m] = hi;
}

template <int MEL_THREADS>
__global__ void __launch_bounds__(MEL_THREADS)
mel_logpower_kernel(const float* __restrict__ wave, long long wave_ld, const int* __restrict__ n_samples,
                    const float* __restrict__ window, const float* __restrict__ fb, const int* __restrict__ ranges,
                    int n_mels, float* __restrict__ feats, int Tmax, int* __restrict__ utt_max) {
  extern __shared__ __align__(16) uint8_t smem_mel[];
  float* raw = reinterpret_cast<float*>(smem_mel);                 // RAW
  float* win = raw + RAW;                                          // 400
  float2* tw = reinterpret_cast<float2*>(win + NFFT);              // 400
  float2* Z = tw + NFFT;                                           // FR * ZLD
  float* P = reinterpret_cast<float*>(Z + FR * ZLD);               // FR * PLD
  int* rng = reinterpret_cast<int*>(P + FR * PLD);                 // 2 * 128
  unsigned short* slot = reinterpret_cast<unsigned short*>(rng + 2 * 128);  // 200: where bin k of the 200-pt FFT was left

  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FR;
  const int N = n_samples[b];
  const int T = 1 + N / HOP;
  if (t0 >= T) return;
  const int tid = threadIdx.x;
  const float* w = wave + (long long)b * wave_ld;

  {
    // all global loads are issued before the first shared-memory store (a load -> store loop with a run-time trip
    // count exposes one DRAM round trip per iteration)
    constexpr int IT = (RAW + MEL_THREADS - 1) / MEL_THREADS;
    float v[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = tid + it * MEL_THREADS;
      const int gi = HOP * t0 + i;  // index in the reflect-padded signal
      v[it] = 0.f;
      if (i < RAW && gi < N + NFFT) {
        int j = gi - NFFT / 2;
        if (j < 0) j = -j;
        if (j >= N) j = 2 * (N - 1) - j;
        v[it] = w[j];
      }
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = tid + it * MEL_THREADS;
      if (i < RAW) raw[i] = v[it];
    }
  }
  for (int i = tid; i < NFFT; i += MEL_THREADS) {
    win[i] = window[i];
    tw[i] = g_tw400[i];
  }
  for (int i = tid; i < 2 * n_mels; i += MEL_THREADS) rng[i] = ranges[i];
  for (int k = tid; k < 200; k += MEL_THREADS) {
    const int k1 = k & 7, k2 = k >> 3;
    slot[k] = (unsigned short)(25 * k1 + 5 * (k2 % 5) + k2 / 5);
  }
  __syncthreads();

  // pass 1: radix-8 over n1 (n = 25 n1 + n2), twiddle W200^(n2 k1)
  for (int idx = tid; idx < FR * 25; idx += MEL_THREADS) {
    const int f = idx / 25, n2 = idx - f * 25;
    const float* r = raw + f * HOP;
    float2 x[8];
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) {
      // n is even and HOP is even: 8-byte loads, consecutive n2 -> consecutive banks
      const int n = 2 * (25 * n1 + n2);
      const float2 rv = *reinterpret_cast<const float2*>(r + n), wv = *reinterpret_cast<const float2*>(win + n);
      x[n1] = make_float2(rv.x * wv.x, rv.y * wv.y);
    }
    dft8(x);
    float2* z = Z + f * ZLD;
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
      float2 v = x[k1];
      if (k1 > 0) v = cmul(v, tw[2 * n2 * k1]);  // 2 n2 k1 <= 336 < 400
      z[25 * k1 + n2] = v;
    }
  }
  __syncthreads();
  // pass 2a: for each (k1, b5): radix-5 over a (n2 = 5a + b5), twiddle W25^(b5 c)
  for (int idx = tid; idx < FR * 40; idx += MEL_THREADS) {
    const int f = idx / 40, rem = idx - f * 40;
    const int k1 = rem / 5, b5 = rem - k1 * 5;
    float2* z = Z + f * ZLD + 25 * k1;
    float2 x[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) x[a] = z[5 * a + b5];
    dft5(x);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      float2 v = x[c];
      if (c > 0 && b5 > 0) v = cmul(v, tw[16 * b5 * c]);  // <= 256 < 400
      z[5 * c + b5] = v;
    }
  }
  __syncthreads();
  // pass 2b: for each (k1, c): radix-5 over b5 -> Z[k1 + 8 (c + 5 e)] kept at slot 25 k1 + 5 c + e
  for (int idx = tid; idx < FR * 40; idx += MEL_THREADS) {
    const int f = idx / 40, rem = idx - f * 40;
    const int k1 = rem / 5, c = rem - k1 * 5;
    float2* z = Z + f * ZLD + 25 * k1 + 5 * c;
    float2 x[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) x[q] = z[q];
    dft5(x);
#pragma unroll
    for (int e = 0; e < 5; ++e) z[e] = x[e];
  }
  __syncthreads();
  // real-FFT split + power
  for (int idx = tid; idx < FR * NBIN; idx += MEL_THREADS) {
    const int f = idx / NBIN, k = idx - f * NBIN;
    const float2* z = Z + f * ZLD;
    const int ka = (k == 200) ? 0 : k, kb = (k == 0 || k == 200) ? 0 : 200 - k;
    const float2 za = z[slot[ka]];
    const float2 zb = z[slot[kb]];
    // E = (Za + conj(Zb))/2 ; O = (Za - conj(Zb))/(2i)
    const float er = 0.5f * (za.x + zb.x), ei = 0.5f * (za.y - zb.y);
    const float orr = 0.5f * (za.y + zb.y), oi = -0.5f * (za.x - zb.x);
    const float2 wk = (k == 200) ? make_float2(-1.f, 0.f) : tw[k];
    const float xr = er + wk.x * orr - wk.y * oi;
    const float xi = ei + wk.x * oi + wk.y * orr;
    P[f * PLD + k] = xr * xr + xi * xi;
  }
  __syncthreads();
  // mel filterbank + dB, running max for top_db
  float vmax = 0.f;  // values are stored shifted by +200 so that they are positive
  for (int idx = tid; idx < FR * n_mels; idx += MEL_THREADS) {
    const int f = idx / n_mels, m = idx - f * n_mels;
    const int t = t0 + f;
    if (t >= T) continue;
    const int lo = rng[m], hi = rng[n_mels + m];
    const float* pf = P + f * PLD;
    float acc = 0.f;
    for (int k = lo; k < hi; ++k) acc = fmaf(pf[k], __ldg(fb + k * n_mels + m), acc);
    const float db = 10.f * log10f(fmaxf(acc, 1e-10f));
    feats[((long long)b * Tmax + t) * n_mels + m] = db;
    vmax = fmaxf(vmax, db + 200.f);
  }
  vmax = warp_max(vmax);
  if ((tid & 31) == 0 && vmax > 0.f) atomicMax(utt_max + b, __float_as_int(vmax));
}

// per-utterance, per-bin sum / sum of squares (double) of the clamped dB values
__global__ void mel_cmvn_stats_kernel(const float* __restrict__ feats, const int* __restrict__ n_samples, int Tmax,
                                      int n_mels, const int* __restrict__ utt_max, double* __restrict__ stats,
                                      int frames_per_cta) {
  const int b = blockIdx.y;
  const int T = 1 + n_samples[b] / HOP;
  const int t_begin = blockIdx.x * frames_per_cta;
  if (t_begin >= T) return;
  const int t_end = min(T, t_begin + frames_per_cta);
  const float cutoff = __int_as_float(utt_max[b]) - 200.f - 80.f;
  // blockDim = (n_mels_pad32, 8): x = bin, y = frame lane
  const int m = threadIdx.x;
  double s = 0.0, ss = 0.0;
  if (m < n_mels)
    for (int t = t_begin + threadIdx.y; t < t_end; t += blockDim.y) {
      float v = fmaxf(feats[((long long)b * Tmax + t) * n_mels + m], cutoff);
      s += v;
      ss += (double)v * v;
    }
  __shared__ double sh[2][8][128];
  sh[0][threadIdx.y][m] = s;
  sh[1][threadIdx.y][m] = ss;
  __syncthreads();
  if (threadIdx.y == 0 && m < n_mels) {
    for (int y = 1; y < blockDim.y; ++y) {
      s += sh[0][y][m];
      ss += sh[1][y][m];
    }
    atomicAdd(stats + ((long long)b * n_mels + m) * 2, s);
    atomicAdd(stats + ((long long)b * n_mels + m) * 2 + 1, ss);
  }
}

// blockDim = (n_mels padded to 32, 8): x = bin (its mean / deviation are computed once per thread), y = frame lane
__global__ void mel_cmvn_apply_kernel(float* __restrict__ feats, const int* __restrict__ n_samples, int Tmax,
                                      int n_mels, const int* __restrict__ utt_max, const double* __restrict__ stats,
                                      int normalize, int frames_per_cta) {
  const int b = blockIdx.y;
  const int T = 1 + n_samples[b] / HOP;
  const float cutoff = __int_as_float(utt_max[b]) - 200.f - 80.f;
  const int m = threadIdx.x;
  if (m >= n_mels) return;
  float mean_f = 0.f, den = 1.f;
  if (normalize) {
    const double s = stats[((long long)b * n_mels + m) * 2], ss = stats[((long long)b * n_mels + m) * 2 + 1];
    const double mean = s / T;
    const double var = (ss - s * mean) / (double)(T - 1);  // unbiased; T == 1 -> NaN like the reference
    const float sd = (float)sqrt(var > 0.0 ? var : (var == var ? 0.0 : var));
    mean_f = (float)mean;
    den = sd + 1e-8f;
  }
  const int t_begin = blockIdx.x * frames_per_cta, t_end = min(Tmax, t_begin + frames_per_cta);
  float* base = feats + (long long)b * Tmax * n_mels + m;
#pragma unroll 4
  for (int t = t_begin + threadIdx.y; t < t_end; t += blockDim.y) {
    float* p = base + (long long)t * n_mels;
    float v = 0.f;
    if (t < T) {
      v = fmaxf(*p, cutoff);
      if (normalize) v = (v - mean_f) / den;
    }
    *p = v;
  }
}

}  // namespace

extern "C" int tasr_mel_init(void) {
  static float2 h_tw[NFFT];
  for (int j = 0; j < NFFT; ++j) {
    double a = -2.0 * M_PI * (double)j / (double)NFFT;
    h_tw[j] = make_float2((float)cos(a), (float)sin(a));
  }
  cudaError_t e = cudaMemcpyToSymbol(g_tw400, h_tw, sizeof(h_tw));
  if (e != cudaSuccess) return tasr_set_cuda_error(e);
  return TASR_OK;
}

extern "C" int tasr_mel_filter_ranges(const float* fb, int n_mels, int* ranges, tasr_stream_t stream) {
  if (n_mels <= 0 || n_mels > 128) return TASR_ERR_SHAPE;
  mel_filter_ranges_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(fb, n_mels, ranges);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" size_t tasr_mel_workspace_bytes(int B, int n_mels) {
  return (size_t)B * 16 + (size_t)B * n_mels * 2 * sizeof(double);
}

extern "C" int tasr_mel_forward(const float* wave, int64_t wave_ld, const int32_t* n_samples, int B, int Tmax,
                                const float* window, const float* fb, const int32_t* ranges, int n_mels,
                                int n_fft, int hop, int normalize, float* feats, void* workspace,
                                size_t workspace_bytes, tasr_stream_t stream) {
  if (n_fft != NFFT || hop != HOP || n_mels <= 0 || n_mels > 128 || B <= 0 || Tmax <= 0) return TASR_ERR_SHAPE;
  if (workspace_bytes < tasr_mel_workspace_bytes(B, n_mels)) return TASR_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int* utt_max = reinterpret_cast<int*>(workspace);
  double* stats = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + (size_t)B * 16);
  cudaError_t e = cudaMemsetAsync(workspace, 0, tasr_mel_workspace_bytes(B, n_mels), st);
  if (e != cudaSuccess) return tasr_set_cuda_error(e);

  const size_t smem = (size_t)RAW * 4 + NFFT * 4 + NFFT * 8 + (size_t)FR * ZLD * 8 + (size_t)FR * PLD * 4 + 2 * 128 * 4 + 200 * 2 + 16;
  static TasrPerDevice attr_done;
  if (!attr_done.get()) {
    e = cudaFuncSetAttribute(mel_logpower_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mel_logpower_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mel_logpower_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tasr_set_cuda_error(e);
    attr_done.set();
  }
  dim3 g1(cdiv(Tmax, FR), B);
  // TASR_MEL_THREADS = 256 (the round-1 CTA size) / 512 / 1024: A/B switch
  static const int nt = [] { const char* v = getenv("TASR_MEL_THREADS"); return v ? atoi(v) : MEL_NT; }();
  if (nt == 256)
    mel_logpower_kernel<256><<<g1, 256, smem, st>>>(wave, wave_ld, n_samples, window, fb, ranges, n_mels, feats, Tmax, utt_max);
  else if (nt == 1024)
    mel_logpower_kernel<1024><<<g1, 1024, smem, st>>>(wave, wave_ld, n_samples, window, fb, ranges, n_mels, feats, Tmax, utt_max);
  else
    mel_logpower_kernel<512><<<g1, 512, smem, st>>>(wave, wave_ld, n_samples, window, fb, ranges, n_mels, feats, Tmax, utt_max);
  TASR_CHECK_LAUNCH();
  if (normalize) {
    const int fpc = 128;
    dim3 g2(cdiv(Tmax, fpc), B);
    dim3 b2(((n_mels + 31) / 32) * 32, 8);
    mel_cmvn_stats_kernel<<<g2, b2, 0, st>>>(feats, n_samples, Tmax, n_mels, utt_max, stats, fpc);
    TASR_CHECK_LAUNCH();
  }
  const int fpa = 64;
  dim3 g3(cdiv(Tmax, fpa), B);
  dim3 b3(((n_mels + 31) / 32) * 32, 8);
  mel_cmvn_apply_kernel<<<g3, b3, 0, st>>>(feats, n_samples, Tmax, n_mels, utt_max, stats, normalize, fpa);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
