// Waveform / feature augmentation on the GPU (BASELINE configs[4]).
// Replaces (reference): data/preprocessing.py:191-228 SpeedPerturbation -> torchaudio.functional.resample
//   (functional.py:1305-1432: hann-windowed sinc, lowpass_filter_width 6, rolloff 0.99) which materialises an
//   (new x 1 x (orig + 2*width)) filter bank (1.14 GB and 4 s per call on the CPU for 16000 -> 17777 Hz);
//   here the <= 2*width + 1 non-zero taps of every output sample are evaluated on the fly.
//   data/preprocessing.py:132-188 SpecAugment -> torchaudio mask_along_axis (value 0.0).
#include "common.cuh"
#include <math.h>
#include <stdio.h>

namespace {

// y[b][j], j = q*n + r:  sum_k x[q*o + k] * h_r[k],  t = (-r/n + k/o) * base clamped to +-6,
// h_r[k] = (base/o) * sinc(pi t) * cos^2(pi t / 12)           (orig_freq o, new_freq n already divided by gcd)
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long long x_ld,
                                                       const int* __restrict__ n_in, const int* __restrict__ orig,
                                                       const int* __restrict__ neww, float* __restrict__ y, long long y_ld,
                                                       int max_out) {
  const int b = blockIdx.y;
  const int o = orig[b], n = neww[b];
  const int N = n_in[b];
  const long long out_len = ((long long)n * N + o - 1) / o;
  const float* xb = x + (long long)b * x_ld;
  float* yb = y + (long long)b * y_ld;
  const double base = (double)min(o, n) * 0.99;
  const int width = (int)ceil(6.0 * (double)o / base);
  const double scale = base / (double)o;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < max_out; j += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (j < out_len) {
      if (o == n) {
        acc = xb[j];
      } else {
        const long long q = j / n;
        const int r = (int)(j - q * n);
        // torchaudio divides -r by new_freq in float32 before promoting to float64 (functional.py:1364); keep that rounding
        const double centre = -(double)__fdiv_rn((float)(-r), (float)n);  // IEEE-rounded, like the host
        // taps with |(-r/n + k/o) * base| < 6
        const int kc = (int)floor(centre * (double)o);
        for (int k = kc - width; k <= kc + width + 1; ++k) {
          if (k < -width || k >= width + o) continue;  // outside torchaudio's kernel support
          const long long xi = q * (long long)o + k;
          if (xi < 0 || xi >= N) continue;             // zero padding
          double t = (-centre + (double)k / (double)o) * base;
          t = fmin(fmax(t, -6.0), 6.0);
          const double win = cos(t * M_PI / 12.0);
          const double tp = t * M_PI;
          const double sinc = (tp == 0.0) ? 1.0 : sin(tp) / tp;
          const float h = (float)(sinc * win * win * scale);
          acc = fmaf(xb[xi], h, acc);
        }
      }
    }
    yb[j] = acc;
  }
}

// params: (B, nmask, 3) int32 = (axis: 0 freq | 1 time, start, end); feats (B, T, F) in place; frames (B) valid frames
__global__ void specaugment_kernel(float* __restrict__ feats, int T, int F, const int* __restrict__ params, int nmask,
                                   const long long* __restrict__ frames) {
  const int b = blockIdx.y;
  const int Tb = frames ? (int)min((long long)T, frames[b]) : T;
  float* fb = feats + (long long)b * T * F;
  for (int m = 0; m < nmask; ++m) {
    const int axis = params[(b * nmask + m) * 3], s = max(0, params[(b * nmask + m) * 3 + 1]);
    int e = params[(b * nmask + m) * 3 + 2];
    if (axis == 0) {
      e = min(e, F);
      const int w = e - s;
      if (w <= 0) continue;
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)Tb * w; i += (long long)gridDim.x * blockDim.x)
        fb[(i / w) * F + s + (i % w)] = 0.f;
    } else {
      e = min(e, Tb);
      const long long cnt = (long long)(e - s) * F;
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x)
        fb[(long long)s * F + i] = 0.f;
    }
  }
}

}  // namespace

extern "C" int tasr_resample_sinc(const float* x, int64_t x_ld, const int32_t* n_in, const int32_t* orig_freq,
                                  const int32_t* new_freq, int B, float* y, int64_t y_ld, int max_out, tasr_stream_t stream) {
  if (B <= 0 || max_out <= 0) return TASR_ERR_SHAPE;
  dim3 grid(min(cdiv(max_out, 256), 148 * 8), B);
  resample_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, x_ld, n_in, orig_freq, new_freq, y, y_ld, max_out);
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}

extern "C" int tasr_specaugment(float* feats, int B, int T, int F, const int32_t* params, int nmask, const int64_t* frames,
                                tasr_stream_t stream) {
  if (B <= 0 || T <= 0 || F <= 0 || nmask < 0) return TASR_ERR_SHAPE;
  if (nmask == 0) return TASR_OK;
  dim3 grid(32, B);
  specaugment_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(feats, T, F, params, nmask,
                                                                              reinterpret_cast<const long long*>(frames));
  TASR_CHECK_LAUNCH();
  return TASR_OK;
}
