"""Drop-in mirror of reference data/preprocessing.py:16-129 (AudioPreprocessor): same constructor and
methods, with the log-mel / dB / CMVN arithmetic done by the fused sm_100a kernel (tasr_mel_forward)
instead of torchaudio's MelSpectrogram + AmplitudeToDB on the CPU.  Adds a batched entry point
(extract_features_batch) that also produces the collate_fn-style zero padding."""
import math
from typing import Optional, Tuple

import torch

from .. import _lib as L

TARGET_SAMPLE_RATE = 16000


def _melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') (the table the reference's
    MelSpectrogram builds at construction, data/preprocessing.py:52-61)."""
    try:
        import torchaudio
        return torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm=None, mel_scale="htk")
    except Exception:  # torchaudio absent: same formula
        all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
        m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
        m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
        m_pts = torch.linspace(m_min, m_max, n_mels + 2)
        f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
        down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
        up = slopes[:, 2:] / f_diff[1:]
        return torch.max(torch.zeros(1), torch.min(down, up))


class AudioPreprocessor:
    def __init__(self, sample_rate: int = TARGET_SAMPLE_RATE, n_mels: int = 80, n_fft: int = 400, hop_length: int = 160,
                 win_length: int = 400, f_min: float = 0.0, f_max: Optional[float] = 8000.0, normalize: bool = True,
                 device: str = "cpu"):
        if n_fft != 400 or hop_length != 160 or win_length != 400:
            raise L.TasrError("the B200 mel kernel implements the reference configuration n_fft=win=400, hop=160")
        self.sample_rate = sample_rate
        self.n_mels = n_mels
        self.normalize = normalize
        self.device = device  # where results are returned (the arithmetic always runs on the CUDA device)
        self._fb_cpu = _melscale_fbanks(n_fft // 2 + 1, f_min, f_max if f_max is not None else sample_rate / 2.0, n_mels,
                                        sample_rate).float().contiguous()
        self._win_cpu = torch.hann_window(win_length)
        self._dev = None

    def _tables(self, dev):
        if self._dev is None or self._dev[0] != dev:
            fb = self._fb_cpu.to(dev)
            self._dev = (dev, fb, self._win_cpu.to(dev), L.mel_filter_ranges(fb))
        return self._dev[1:]

    def load_audio(self, path: str) -> Tuple[torch.Tensor, int]:
        """reference data/preprocessing.py:66-79 (file I/O stays on torchaudio)."""
        import torchaudio
        waveform, sr = torchaudio.load(path)
        if waveform.shape[0] > 1:
            waveform = torch.mean(waveform, dim=0, keepdim=True)
        if sr != self.sample_rate:
            waveform = torchaudio.transforms.Resample(orig_freq=sr, new_freq=self.sample_rate)(waveform)
        return waveform, self.sample_rate

    def extract_features_batch(self, waves: torch.Tensor, lengths: torch.Tensor, tmax: Optional[int] = None):
        """waves (B, Nmax) fp32 (CUDA), lengths (B,) samples -> feats (B, Tmax, n_mels) on the CUDA device
        (frames >= T_b zero-filled like collate_fn), frame counts (B,) int64 = 1 + N // 160."""
        if not waves.is_cuda:
            waves = waves.cuda(non_blocking=True)
        dev = waves.device
        fb, win, ranges = self._tables(dev)
        lengths_cpu = lengths if not lengths.is_cuda else None
        n32 = lengths.to(device=dev, dtype=torch.int32)
        if tmax is None:
            if lengths_cpu is None:
                lengths_cpu = lengths.cpu()
            tmax = 1 + int(lengths_cpu.max()) // 160
        feats = L.mel_forward(waves.contiguous().float(), n32, tmax, win, fb, ranges, normalize=self.normalize)
        frames = 1 + torch.div(lengths.to(torch.int64), 160, rounding_mode="floor")
        return feats, frames

    def extract_features(self, waveform: torch.Tensor) -> torch.Tensor:
        """reference data/preprocessing.py:81-110: (1, N) or (N,) -> (T, n_mels)."""
        if waveform.dim() == 2:
            waveform = waveform[0]
        n = waveform.shape[0]
        if n <= 200:
            raise RuntimeError("reflect padding needs more than n_fft // 2 = 200 samples (same as torch.stft)")
        feats, _ = self.extract_features_batch(waveform.reshape(1, -1), torch.tensor([n]))
        return feats[0].to(self.device)

    def __call__(self, path: str) -> torch.Tensor:
        waveform, _ = self.load_audio(path)
        return self.extract_features(waveform)


class SpecAugment:
    """reference data/preprocessing.py:132-188 (2 frequency masks of width <= 27, 2 time masks of width <= 100,
    value 0.0 after CMVN).  Mask parameters are drawn on the host exactly like torchaudio's mask_along_axis
    (two torch.rand(1) per mask: value, then min_value); the masking itself is a slice fill."""

    def __init__(self, freq_mask_param: int = 27, time_mask_param: int = 100, n_freq_masks: int = 2,
                 n_time_masks: int = 2):
        self.freq_mask_param = freq_mask_param
        self.time_mask_param = time_mask_param
        self.n_freq_masks = n_freq_masks
        self.n_time_masks = n_time_masks

    @staticmethod
    def draw(size, param):
        value = torch.rand(1) * param
        min_value = torch.rand(1) * (size - value)
        start = int(min_value.long())
        end = start + int(value.long())
        return start, end

    def mask_params(self, n_frames, n_mels):
        out = []
        for _ in range(self.n_freq_masks):
            out.append(("f",) + self.draw(n_mels, self.freq_mask_param))
        for _ in range(self.n_time_masks):
            out.append(("t",) + self.draw(n_frames, self.time_mask_param))
        return out

    def __call__(self, features: torch.Tensor, params=None) -> torch.Tensor:
        """features (T, n_mels) -> masked copy."""
        t, f = features.shape
        params = self.mask_params(t, f) if params is None else params
        out = features.clone()
        for axis, s, e in params:
            if e <= s:
                continue
            if axis == "f":
                out[:, max(s, 0):e] = 0.0
            else:
                out[max(s, 0):e, :] = 0.0
        return out


    def apply_batch(self, feats: torch.Tensor, frames: torch.Tensor, params=None) -> torch.Tensor:
        """Batched GPU form: feats (B, Tmax, n_mels) CUDA, masked in place; per-utterance parameters are drawn on the
        host exactly as the reference would draw them utterance by utterance (size = that utterance's frame count)."""
        B, T, F = feats.shape
        frames_cpu = frames.cpu().tolist()
        if params is None:
            params = [self.mask_params(int(n), F) for n in frames_cpu]
        flat = [[(0 if a == "f" else 1), int(s), int(e)] for per in params for (a, s, e) in per]
        nmask = len(params[0]) if params else 0
        p = torch.tensor(flat, dtype=torch.int32).view(B, nmask, 3).to(feats.device)
        L.specaugment_(feats, p, frames.to(device=feats.device, dtype=torch.int64))
        return feats


class SpeedPerturbation:
    """reference data/preprocessing.py:191-228: pick speed in `speeds` with torch.randint, resample with
    torchaudio.functional.resample(waveform, sr, int(sr / speed)).  The resampling runs on the GPU
    (tasr_resample_sinc evaluates the <= 15 non-zero windowed-sinc taps per output sample on the fly)."""

    def __init__(self, speeds=(0.9, 1.0, 1.1)):
        self.speeds = list(speeds)

    def pick(self):
        return self.speeds[int(torch.randint(len(self.speeds), (1,)).item())]

    @staticmethod
    def freqs(speed: float, sample_rate: int):
        import math
        new_freq = int(sample_rate / speed)
        g = math.gcd(sample_rate, new_freq)
        return sample_rate // g, new_freq // g

    def apply_batch(self, waves: torch.Tensor, n_samples: torch.Tensor, sample_rate: int = TARGET_SAMPLE_RATE, speeds=None):
        """waves (B, Nmax) CUDA, n_samples (B,) -> (resampled (B, Nmax'), new lengths (B,) int64 = ceil(new*N/orig))."""
        B = waves.shape[0]
        if speeds is None:
            speeds = [self.pick() for _ in range(B)]
        ns = n_samples.cpu().tolist()
        orig, new, out_len = [], [], []
        for n, sp in zip(ns, speeds):
            o, m = (1, 1) if sp == 1.0 else self.freqs(sp, sample_rate)
            orig.append(o)
            new.append(m)
            out_len.append(-(-m * int(n) // o))
        dev = waves.device if waves.is_cuda else torch.device("cuda")
        y = L.resample_sinc(waves.to(dev).contiguous().float(), torch.tensor(ns, dtype=torch.int32, device=dev),
                            torch.tensor(orig, dtype=torch.int32, device=dev), torch.tensor(new, dtype=torch.int32, device=dev),
                            max(out_len))
        return y, torch.tensor(out_len, dtype=torch.int64)

    def __call__(self, waveform: torch.Tensor, sample_rate: int):
        """Reference call signature: (1, N) or (N,) waveform -> (perturbed waveform, sample_rate)."""
        speed = self.pick()
        if speed == 1.0:
            return waveform, sample_rate
        w = waveform.reshape(1, -1)
        y, n = self.apply_batch(w, torch.tensor([w.shape[1]]), sample_rate, speeds=[speed])
        y = y[:, : int(n[0])].to(waveform.device)
        return (y if waveform.dim() == 2 else y[0]), sample_rate
