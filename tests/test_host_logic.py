"""CPU: host-side logic of the drop-in package (sampler, collate, flat parameter layout, ABI exports)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest
import torch

from oracle import sampler as osamp
from turkish_asr_model_b200 import _lib as L
from turkish_asr_model_b200.data.dataset import BucketingSampler, collate_fn
from turkish_asr_model_b200.data.preprocessing import SpecAugment
from turkish_asr_model_b200.engine import FlatParams
from turkish_asr_model_b200.model import TurkishASRModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tasr_kernels.h")).read()
    declared = set(re.findall(r"\b(tasr_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tasr_stream_t"}
    assert len(declared) >= 40
    lib = ctypes.CDLL(L.LIB_PATH)  # loads without a GPU; no compute calls here
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(L._SIGNATURES) <= declared
    assert lib.tasr_version() == 1


def test_no_cpu_fallback():
    x = torch.zeros(2, 8, 256)
    with pytest.raises(L.TasrError):
        L.groupnorm_fwd(x, 32, torch.ones(256), torch.zeros(256))
    m = TurkishASRModel(80, 128, 2, 1, 16)
    with pytest.raises(L.TasrError):
        m(torch.zeros(1, 40, 80))


def test_bucketing_sampler_bit_exact_vs_golden_and_oracle():
    g = np.load(os.path.join(GOLD, "sampler_golden.npz"))
    sizes = g["sizes"].tolist()
    for key in g.files:
        if not key.startswith("order_"):
            continue
        _, bs, seed, drop = key.split("_")
        bs, seed, drop = int(bs[2:]), int(seed[4:]), bool(int(drop[4:]))
        random.seed(seed)  # reference semantics: global random.shuffle
        got = list(iter(BucketingSampler(None, bs, shuffle=True, drop_last=drop, lengths=sizes)))
        assert got == g[key].tolist(), key
        # private stream with seed=... reproduces random.seed(seed + epoch); random.shuffle(...)
        s = BucketingSampler(None, bs, shuffle=True, drop_last=drop, lengths=sizes, seed=seed)
        assert list(iter(s)) == g[key].tolist()
        assert len(s) == len(got)


def test_rank_sharded_sampler_union_is_reference_at_global_batch():
    rng = np.random.RandomState(0)
    sizes = rng.randint(1000, 9000, size=517).tolist()
    B, W = 8, 4
    random.seed(11)
    # the reference algorithm at bucket size W*B; the (single) short bucket is dropped so all ranks step together
    ref_batches = [b for b in osamp.bucketing_buckets(sizes, B * W, shuffle=True, drop_last=False) if len(b) == B * W]
    per_rank = [list(iter(BucketingSampler(None, B, lengths=sizes, rank=r, world_size=W, seed=11))) for r in range(W)]
    assert len({len(p) for p in per_rank}) == 1  # same number of steps on every rank
    nsteps = len(per_rank[0]) // B
    assert nsteps == len(ref_batches)
    for k in range(nsteps):
        union = []
        for r in range(W):
            union += per_rank[r][k * B:(k + 1) * B]
        assert union == ref_batches[k]  # union over ranks == the reference's global batch, in order


def test_collate_fn():
    feats = [torch.randn(7, 80), torch.randn(11, 80), torch.randn(3, 80)]
    tg = [torch.tensor([1, 2, 3]), torch.tensor([4]), torch.tensor([5, 6])]
    f, t, il, tl = collate_fn(list(zip(feats, tg)))
    assert f.shape == (3, 11, 80) and t.shape == (3, 3)
    assert il.tolist() == [7, 11, 3] and tl.tolist() == [3, 1, 2]
    assert torch.all(f[0, 7:] == 0) and torch.all(t[1, 1:] == 0)
    assert collate_fn([None]) == (None, None, None, None)


def test_flat_params_layout():
    torch.manual_seed(0)
    m = TurkishASRModel(80, 128, 2, 2, 40)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    fp = FlatParams(m)
    after = m.state_dict()
    assert list(after.keys()) == list(before.keys())
    for k in before:
        assert torch.equal(after[k], before[k]), k
    d = 128
    # fused QKV view is contiguous: q (d,d) | k (64,d) | v (64,d)
    wqkv = fp.view(fp.params, "blocks.0.attn.linear_q.weight", (d + 128, d))
    assert torch.equal(wqkv[:d], m.blocks[0].attn.linear_q.weight)
    assert torch.equal(wqkv[d:d + 64], m.blocks[0].attn.linear_k.weight)
    assert torch.equal(wqkv[d + 64:], m.blocks[0].attn.linear_v.weight)
    bqkv = fp.view(fp.params, "blocks.0.attn.linear_q.bias", (d + 128,))
    assert torch.equal(bqkv[d:d + 64], m.blocks[0].attn.linear_k.bias)
    # dead parameters live in the tail the optimizer skips
    assert fp.offsets["blocks.0.norm_conv.norm.weight"] >= fp.live_numel
    assert fp.offsets["fc.bias"] < fp.live_numel
    assert all(o % 8 == 0 for o in fp.offsets.values())
    # parameters alias the flat buffer
    with torch.no_grad():
        m.fc.bias.add_(1.0)
    assert torch.equal(fp.view(fp.params, "fc.bias"), m.fc.bias)
    assert fp.still_valid()
    assert "blocks.1.norm_conv.norm.bias" not in fp.grad_views()


def test_specaugment_draws_like_torchaudio():
    torch.manual_seed(5)
    sa = SpecAugment()
    params = sa.mask_params(600, 80)
    assert [p[0] for p in params] == ["f", "f", "t", "t"]
    torch.manual_seed(5)
    v = torch.rand(1) * 27
    mn = torch.rand(1) * (80 - v)
    assert params[0][1:] == (int(mn.long()), int(mn.long()) + int(v.long()))
    x = torch.ones(600, 80)
    y = sa(x, params)
    for axis, s, e in params:
        if axis == "f":
            assert torch.all(y[:, s:e] == 0)
        else:
            assert torch.all(y[s:e, :] == 0)
