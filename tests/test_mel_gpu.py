"""Fused log-mel kernel (tasr_mel_forward) vs the numpy float64 oracle (oracle/mel.py)."""
import numpy as np
import pytest
import torch

from oracle import mel as om
from turkish_asr_model_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _run(cuda, lengths, seed, normalize=True, zero_head=()):
    g = torch.Generator().manual_seed(seed)
    nmax = max(lengths)
    wave = torch.zeros(len(lengths), nmax)
    for b, n in enumerate(lengths):
        wave[b, :n] = 0.1 * torch.randn(n, generator=g)
        if b in zero_head:
            wave[b, : min(8000, n // 2)] = 0.0
    fb = torch.tensor(om.melscale_fbanks(), dtype=torch.float32, device=cuda)
    win = torch.tensor(om.hann_periodic(), dtype=torch.float32, device=cuda)
    ranges = L.mel_filter_ranges(fb)
    ns = torch.tensor(lengths, dtype=torch.int32, device=cuda)
    tmax = max(om.num_frames(n) for n in lengths)
    feats = L.mel_forward(wave.to(cuda), ns, tmax, win, fb, ranges, normalize=normalize)
    torch.cuda.synchronize()
    ref, ts = om.log_mel_batch(wave.numpy(), lengths, normalize=normalize)
    return feats.cpu().numpy().astype(np.float64), ref, ts


def test_mel_c1_batch(cuda):
    got, ref, ts = _run(cuda, [160000] * 8, 1234, zero_head=(0, 4))
    assert got.shape == ref.shape == (8, 1001, 80)  # frame count parity: T = 1 + N // 160
    assert np.abs(got - ref).max() <= 1e-4  # north_star: log-mel 1e-4 absolute (fp32)


def test_mel_ragged_and_padding(cuda):
    lengths = [80000, 123457, 240000, 201, 16000, 159999, 160001]
    got, ref, ts = _run(cuda, lengths, 7, zero_head=(2,))
    assert got.shape == ref.shape
    for b, t in enumerate(ts):
        assert np.all(got[b, t:] == 0.0)  # collate-style zero padding
    # T = 2 for N = 201: std over two frames, still finite
    assert np.abs(got - ref).max() <= 1e-4


def test_mel_unnormalized_db(cuda):
    got, ref, ts = _run(cuda, [48000, 32000], 3, normalize=False, zero_head=(0,))
    assert np.abs(got - ref).max() <= 2e-4  # raw dB values (|x| up to 100): 2e-6 relative
