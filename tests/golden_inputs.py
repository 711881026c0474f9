"""Seeded inputs shared by the golden-fixture tests: they rebuild exactly what tools/make_golden.py fed to the
reference (the fixtures store outputs and input checksums only)."""
import numpy as np
import torch

GCFG = dict(n_mels=80, d_model=128, n_heads=2, n_blocks=1, n_classes=32)


def big_case_inputs(g):
    """Rebuild the seeded inputs of tools/make_golden.py::big_case (the fixture stores only checksums of x)."""
    from turkish_asr_model_b200.model import TurkishASRModel
    d, H, nb, V, B, T, S, ts, vs = [int(v) for v in g["cfg"]]
    torch.manual_seed(0)
    m = TurkishASRModel(80, d, H, nb, V, dropout=0.0)
    gen = torch.Generator().manual_seed(4321)
    x = torch.randn(B, T, 80, generator=gen)
    il = torch.from_numpy(g["input_lengths"])
    for b in range(B):
        x[b, il[b]:] = 0.0
    assert abs(float(x.double().sum()) - g["x_checksum"][0]) < 1e-6 * max(1.0, g["x_checksum"][1])
    targets = torch.randint(1, V, (B, S), generator=gen)
    assert np.array_equal(targets.numpy(), g["targets"])
    return m, x, il, targets, torch.from_numpy(g["target_lengths"]), (d, H, nb, V, ts, vs)


def trainer_golden_batches():
    """Seeded batches of tools/make_golden.py::trainer_case."""
    gen = torch.Generator().manual_seed(77)
    batches = []
    for i in range(3):
        T = 67 + 8 * i
        xb = torch.randn(2, T, 80, generator=gen)
        ilb = torch.tensor([T, T - 20])
        xb[1, T - 20:] = 0.0
        batches.append((xb, torch.randint(1, GCFG["n_classes"], (2, 5), generator=gen), ilb, torch.tensor([5, 3])))
    return batches
