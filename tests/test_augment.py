"""Augmentation: oracle vs torchaudio (CPU) and CUDA kernels vs oracle (GPU)."""
import numpy as np
import pytest
import torch

from oracle import augment as oa


def test_resample_oracle_vs_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    from torchaudio.functional.functional import _apply_sinc_resample_kernel, _get_sinc_resample_kernel
    import math
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4000, generator=g)
    for speed in (1.1, 0.9):  # 0.9 builds torchaudio's full 17777 x 16014 kernel bank once (~1 GB, a few seconds)
        new_freq = int(16000 / speed)
        ref32 = torchaudio.functional.resample(x, 16000, new_freq)[0].numpy()  # what SpeedPerturbation calls
        gcd = math.gcd(16000, new_freq)
        k, width = _get_sinc_resample_kernel(16000, new_freq, gcd)
        ref64 = _apply_sinc_resample_kernel(x.double(), 16000, new_freq, gcd, k.double(), width)[0].numpy()
        o, n = oa.speed_to_freqs(speed)
        got = oa.resample_sinc(x[0].numpy(), o, n)
        assert got.shape == ref32.shape  # length parity: ceil(new * N / orig)
        # torchaudio's own fp32 kernel bank, convolved in float64: the oracle restates exactly this
        assert np.abs(got - ref64).max() < 1e-6
        # torch's fp32 conv1d over the 3 k+ tap (mostly zero) kernel carries ~4e-4 of its own rounding noise
        assert np.abs(got - ref32).max() < 2e-3
        del k


def test_specaugment_host_draws_match_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    from turkish_asr_model_b200.data.preprocessing import SpecAugment
    feats = torch.randn(600, 80)
    sa = SpecAugment()
    torch.manual_seed(3)
    params = sa.mask_params(600, 80)
    got = oa.spec_augment(feats.numpy(), params)
    torch.manual_seed(3)  # the reference: features.T.unsqueeze(0) -> 2 freq masks, 2 time masks (preprocessing.py:167-186)
    spec = feats.t().unsqueeze(0)
    for _ in range(2):
        spec = torchaudio.transforms.FrequencyMasking(27)(spec)
    for _ in range(2):
        spec = torchaudio.transforms.TimeMasking(100)(spec)
    assert np.array_equal(got, spec.squeeze(0).t().numpy())


@pytest.mark.gpu
def test_resample_kernel_vs_oracle(cuda):
    from turkish_asr_model_b200.data.preprocessing import SpeedPerturbation
    g = torch.Generator().manual_seed(1)
    lengths = [3000, 2500, 1800]
    waves = torch.zeros(3, 3000)
    for i, n in enumerate(lengths):
        waves[i, :n] = torch.randn(n, generator=g)
    speeds = [0.9, 1.1, 1.0]
    sp = SpeedPerturbation()
    y, new_len = sp.apply_batch(waves.to(cuda), torch.tensor(lengths), speeds=speeds)
    for i, (n, s) in enumerate(zip(lengths, speeds)):
        o, m = (1, 1) if s == 1.0 else oa.speed_to_freqs(s)
        ref = oa.resample_sinc(waves[i, :n].numpy(), o, m)
        assert int(new_len[i]) == ref.shape[0]  # integer parity (177770 / 145450 for 160000 samples)
        assert np.abs(y[i, : ref.shape[0]].cpu().numpy() - ref).max() < 1e-5
        assert torch.all(y[i, ref.shape[0]:] == 0)
    assert SpeedPerturbation.freqs(0.9, 16000) == (16000, 17777) and SpeedPerturbation.freqs(1.1, 16000) == (3200, 2909)
    assert -(-17777 * 160000 // 16000) == 177770 and -(-2909 * 160000 // 3200) == 145450


@pytest.mark.gpu
def test_specaugment_kernel_vs_oracle(cuda):
    from turkish_asr_model_b200.data.preprocessing import SpecAugment
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(3, 700, 80, generator=g)
    frames = torch.tensor([700, 612, 505])
    sa = SpecAugment()
    torch.manual_seed(9)
    params = [sa.mask_params(int(n), 80) for n in frames]
    out = sa.apply_batch(feats.clone().to(cuda), frames, params=params).cpu().numpy()
    for b in range(3):
        n = int(frames[b])
        ref = oa.spec_augment(feats[b, :n].numpy(), params[b])
        assert np.array_equal(out[b, :n], ref)
        assert np.array_equal(out[b, n:], feats[b, n:].numpy())  # padding frames untouched (they are zero in practice)


@pytest.mark.gpu
def test_batched_inference_ids(cuda):
    from oracle import conformer as oc
    from oracle import mel as om
    from turkish_asr_model_b200.inference import BatchedInference
    from turkish_asr_model_b200.model import TurkishASRModel
    torch.manual_seed(6)
    model = TurkishASRModel(80, 128, 2, 1, 50, dropout=0.1)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    inf = BatchedInference(model.to(cuda))
    g = torch.Generator().manual_seed(7)
    ns = torch.tensor([24000, 17000])
    waves = torch.zeros(2, 24000)
    for i in range(2):
        waves[i, : ns[i]] = 0.1 * torch.randn(int(ns[i]), generator=g)
    logits, lengths = inf.logits(waves.to(cuda), ns)
    feats, frames = om.log_mel_batch(waves.numpy(), ns.tolist())
    ref = oc.forward(torch.from_numpy(feats.astype(np.float32)), torch.from_numpy(frames), sd, 2, 1, training=False)
    assert lengths.tolist() == (frames // 4).tolist()
    assert ((logits.float().cpu() - ref).abs().max() / ref.abs().max()).item() < 2e-2
    ids = inf.transcribe_ids(waves.to(cuda), ns)
    _, ref_ids = oc.greedy_ids(logits.float().cpu(), lengths)
    assert ids == ref_ids
    # staged host batches (copy stream + reusable device buffers) give the same ids; handles can be prepared ahead
    hw = waves.pin_memory()
    h0, h1 = inf.stage(hw), inf.stage(hw)
    assert inf.transcribe_ids(h0, ns) == ref_ids and inf.transcribe_ids(h1, ns) == ref_ids
    h2 = inf.stage(hw)
    assert inf.transcribe_ids(h2, ns) == ref_ids and len(inf._staged) <= 3
    jobs = [inf.submit(inf.stage(hw), ns) for _ in range(2)]  # asynchronous form: results are read later
    assert [j.result() for j in jobs] == [ref_ids, ref_ids]
