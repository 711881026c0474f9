"""Augmentation oracle (numpy float64).  TEST INFRASTRUCTURE ONLY.

SpeedPerturbation (reference data/preprocessing.py:191-228) = torchaudio.functional.resample(waveform, sr,
int(sr / speed)) with the defaults lowpass_filter_width=6, rolloff=0.99, sinc_interp_hann (torchaudio 2.11
functional/functional.py:1305-1432), restated sample by sample.
SpecAugment (data/preprocessing.py:132-188, torchaudio mask_along_axis) = slice fill with 0.0.
"""
import math

import numpy as np


def speed_to_freqs(speed, sample_rate=16000):
    """(orig, new) divided by their gcd, as torchaudio does before building the kernel."""
    new_freq = int(sample_rate / speed)
    g = math.gcd(sample_rate, new_freq)
    return sample_rate // g, new_freq // g


def resample_sinc(x, orig, new, lowpass_filter_width=6, rolloff=0.99):
    """x (N,) -> (ceil(new*N/orig),) float64; orig/new already reduced by their gcd."""
    x = np.asarray(x, dtype=np.float64)
    n_in = x.shape[0]
    if orig == new:
        return x.copy()
    base = min(orig, new) * rolloff
    width = int(math.ceil(lowpass_filter_width * orig / base))
    out_len = int(math.ceil(new * n_in / orig))
    y = np.zeros(out_len)
    ks = np.arange(-width, width + orig)
    for j in range(out_len):
        q, r = divmod(j, new)
        # torchaudio forms -r/new from an int64 arange divided in the default dtype (float32) and only then promotes
        # to float64 (functional.py:1364); reproduce that rounding, it moves the taps by up to ~1e-4
        t = (float(np.float32(-r) / np.float32(new)) + ks / orig) * base
        t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
        nz = np.abs(t) < lowpass_filter_width
        k = ks[nz]
        tt = t[nz] * math.pi
        win = np.cos(tt / lowpass_filter_width / 2) ** 2
        sinc = np.where(tt == 0, 1.0, np.sin(tt) / np.where(tt == 0, 1.0, tt))
        h = sinc * win * (base / orig)
        xi = q * orig + k
        ok = (xi >= 0) & (xi < n_in)
        y[j] = np.dot(x[xi[ok]], h[ok])
    return y


def spec_augment(feats, params):
    """feats (T, F); params list of (axis 'f'|'t', start, end) -> masked copy."""
    out = np.array(feats, copy=True)
    for axis, s, e in params:
        if e <= s:
            continue
        if axis == "f":
            out[:, max(s, 0):e] = 0.0
        else:
            out[max(s, 0):e, :] = 0.0
    return out
