"""Log-mel front-end oracle (numpy float64).  TEST INFRASTRUCTURE ONLY.

Follows reference data/preprocessing.py:52-64 (MelSpectrogram(n_fft=400, win=400, hop=160, f 0-8000,
80 mels, htk) + AmplitudeToDB('power', top_db=80)), :98-104 (extract_features) and :112-116 (CMVN),
and the torchaudio 2.11 functions they call: functional.spectrogram (center=True, reflect pad,
periodic hann, power=2), functional.melscale_fbanks (norm=None, htk), functional.amplitude_to_DB
(amin=1e-10, ref=1, one cutoff per utterance).
"""
import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_MELS = 80


def hann_periodic(n=N_FFT):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + f / 700.0)


def mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (m / 2595.0) - 1.0)


def melscale_fbanks(n_freqs=N_FFT // 2 + 1, f_min=0.0, f_max=8000.0, n_mels=N_MELS, sample_rate=SAMPLE_RATE):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') -> (n_freqs, n_mels)."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_pts = np.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = mel_to_hz_htk(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


def num_frames(n_samples):
    """mel frame count T = 1 + N // hop (center=True STFT); integer-exact parity item."""
    return 1 + int(n_samples) // HOP


def log_mel(wave, normalize=True, fb=None, window=None):
    """wave: (N,) float -> (T, 80) float64 log-mel (+CMVN)."""
    w = np.asarray(wave, dtype=np.float64).reshape(-1)
    n = w.shape[0]
    if n <= N_FFT // 2:
        raise ValueError("reflect padding needs N > n_fft/2 (same as torch.stft)")
    fb = melscale_fbanks() if fb is None else np.asarray(fb, dtype=np.float64)
    window = hann_periodic() if window is None else np.asarray(window, dtype=np.float64)
    x = np.pad(w, (N_FFT // 2, N_FFT // 2), mode="reflect")
    t = num_frames(n)
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(t)[:, None]
    frames = x[idx] * window[None, :]
    spec = np.fft.rfft(frames, n=N_FFT, axis=1)
    power = spec.real ** 2 + spec.imag ** 2
    mel = power @ fb
    db = 10.0 * np.log10(np.maximum(mel, 1e-10))
    db = np.maximum(db, db.max() - 80.0)
    if normalize:
        mean = db.mean(axis=0, keepdims=True)
        std = db.std(axis=0, ddof=1, keepdims=True)
        db = (db - mean) / (std + 1e-8)
    return db


def log_mel_batch(waves, lengths, normalize=True, fb=None, window=None):
    """Batched form with collate_fn-style zero padding (reference data/dataset.py:283-312).
    waves: (B, Nmax); lengths: (B,) samples.  Returns feats (B, Tmax, 80) float64, frame counts (B,) int64."""
    lengths = [int(v) for v in lengths]
    ts = [num_frames(n) for n in lengths]
    tmax = max(ts)
    out = np.zeros((len(lengths), tmax, (N_MELS if fb is None else np.asarray(fb).shape[1])), dtype=np.float64)
    for b, n in enumerate(lengths):
        out[b, : ts[b]] = log_mel(np.asarray(waves[b])[:n], normalize=normalize, fb=fb, window=window)
    return out, np.asarray(ts, dtype=np.int64)
