"""CPU: the oracle (oracle/) against the golden fixtures produced by running the reference
(tools/make_golden.py), and - when /root/reference is mounted - against the live reference."""
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

from oracle import conformer as oc
from oracle import ctc as octc
from oracle import mel as om
from oracle import sampler as osamp

from golden_inputs import GCFG, big_case_inputs, trainer_golden_batches

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"


def _golden_model_sd():
    """Same seeded construction as tools/make_golden.py; the checksums in the fixture prove the weights match."""
    from turkish_asr_model_b200.model import TurkishASRModel
    torch.manual_seed(0)
    m = TurkishASRModel(GCFG["n_mels"], GCFG["d_model"], GCFG["n_heads"], GCFG["n_blocks"], GCFG["n_classes"], dropout=0.0)
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def test_mel_oracle_vs_golden():
    g = np.load(os.path.join(GOLD, "mel_golden.npz"))
    for i, n in enumerate(g["lengths"]):
        ref = g["feat%d" % i]
        got = om.log_mel(g["wave%d" % i])
        assert got.shape == ref.shape == (om.num_frames(n), 80)  # frame-count parity (integer, exact)
        assert np.abs(got - ref).max() < 1e-4, i  # the reference is fp32; north_star tolerance 1e-4


def test_seeded_init_matches_reference_checksums():
    g = np.load(os.path.join(GOLD, "model_golden.npz"))
    _, sd = _golden_model_sd()
    assert list(g["checksum_names"]) == list(sd.keys())  # state_dict layout identical to the reference
    for name, (s, a) in zip(g["checksum_names"], g["checksums"]):
        v = sd[str(name)].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(a)), name
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, abs(a)), name


def test_model_oracle_vs_golden():
    g = np.load(os.path.join(GOLD, "model_golden.npz"))
    _, sd = _golden_model_sd()
    x = torch.from_numpy(g["x"])
    il = torch.from_numpy(g["input_lengths"])
    bn_state = {}
    logits = oc.forward(x, il, sd, GCFG["n_heads"], GCFG["n_blocks"], training=True, bn_state=bn_state)
    assert logits.shape == g["logits_train"].shape == (2, oc.encoder_frames(67), 32)
    assert np.abs(logits.numpy() - g["logits_train"]).max() < 1e-5
    loss = oc.ctc_loss_torch(logits, torch.from_numpy(g["targets"]), il, torch.from_numpy(g["target_lengths"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    l2, nll, grad = octc.ctc_loss_and_grad(logits.numpy(), g["targets"], (g["input_lengths"] // 4), g["target_lengths"])
    assert abs(l2 - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    for i, name in enumerate(g["bn_names"]):
        name = str(name)
        if "num_batches" in name:
            continue
        assert np.abs(bn_state[name].numpy() - g["bn%d" % i]).max() < 1e-6, name
    sd_eval = dict(sd)
    sd_eval.update(bn_state)
    le = oc.forward(x, il, sd_eval, GCFG["n_heads"], GCFG["n_blocks"], training=False)
    assert np.abs(le.numpy() - g["logits_eval"]).max() < 1e-5
    ids, toks = oc.greedy_ids(le)
    assert np.array_equal(ids.numpy(), g["greedy_ids"])  # integer, bit-exact


def test_model_oracle_grad_norms_vs_golden():
    g = np.load(os.path.join(GOLD, "model_golden.npz"))
    m, sd = _golden_model_sd()
    pnames = [n for n, _ in m.named_parameters()]
    sdr = {k: (v.clone().requires_grad_(True) if k in pnames else v) for k, v in sd.items()}
    x, il = torch.from_numpy(g["x"]), torch.from_numpy(g["input_lengths"])
    logits = oc.forward(x, il, sdr, GCFG["n_heads"], GCFG["n_blocks"], training=True)
    oc.ctc_loss_torch(logits, torch.from_numpy(g["targets"]), il, torch.from_numpy(g["target_lengths"])).backward()
    for name, ref in zip(g["grad_names"], g["grad_norms"]):
        gr = sdr[str(name)].grad
        if ref < 0:
            assert gr is None, name  # dead norm_conv parameters: grad None in the reference too
        else:
            assert abs(float(gr.norm()) - ref) <= 1e-4 * max(ref, 1e-6) + 1e-7, name


def test_sampler_oracle_vs_golden():
    g = np.load(os.path.join(GOLD, "sampler_golden.npz"))
    sizes = g["sizes"].tolist()
    for key in g.files:
        if not key.startswith("order_"):
            continue
        _, bs, seed, drop = key.split("_")
        bs, seed, drop = int(bs[2:]), int(seed[4:]), bool(int(drop[4:]))
        random.seed(seed)
        got = osamp.bucketing_order(sizes, bs, shuffle=True, drop_last=drop)
        assert got == g[key].tolist(), key  # bit-exact


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not mounted")
def test_oracle_vs_live_reference():
    sys.path.insert(0, REF)
    sys.modules.setdefault("jiwer", types.ModuleType("jiwer"))
    from data.preprocessing import AudioPreprocessor
    from model.conformer import TurkishASRModel as RefModel
    g = torch.Generator().manual_seed(99)
    w = 0.1 * torch.randn(20000, generator=g)
    assert np.abs(AudioPreprocessor().extract_features(w).numpy() - om.log_mel(w.numpy())).max() < 1e-4
    torch.manual_seed(3)
    ref = RefModel(80, 256, 4, 2, 50, dropout=0.0).train()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    x = torch.randn(3, 131, 80, generator=g)
    il = torch.tensor([131, 100, 64])
    assert (ref(x, il) - oc.forward(x, il, sd, 4, 2, training=True)).abs().max().item() < 1e-5
    assert (ref(x, None) - oc.forward(x, None, sd, 4, 2, training=True)).abs().max().item() < 1e-5
    ref.eval()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    with torch.no_grad():
        assert (ref(x, il) - oc.forward(x, il, sd, 4, 2, training=False)).abs().max().item() < 1e-5


def test_ctc_oracle_vs_torch_ctc_loss():
    """The fp64 CTC restatement (oracle/ctc.py) against torch's own CTCLoss + autograd on the CPU, i.e. exactly the
    reference's call (trainer/trainer.py:76,167-173: log_softmax -> CTCLoss(blank=0, reduction='mean',
    zero_infinity=True)), including repeated labels, a short input, an empty target and an infeasible sample."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    B, T, V, S = 5, 40, 17, 12
    logits = torch.randn(B, T, V, generator=g, dtype=torch.float64, requires_grad=True)
    targets = torch.randint(1, V, (B, S), generator=g)
    targets[1, 3] = targets[1, 2]                      # repeated label (needs a blank in between)
    tl = torch.tensor([12, 7, 0, 12, 3])
    il = torch.tensor([40, 25, 10, 11, 40])            # sample 3: 12 labels in 11 frames -> infeasible -> 0 loss, 0 grad
    lp = F.log_softmax(logits, dim=-1).transpose(0, 1)
    ref = F.ctc_loss(lp, targets, il, tl, blank=0, reduction="mean", zero_infinity=True)
    ref.backward()
    loss, nll, grad = octc.ctc_loss_and_grad(logits.detach().numpy(), targets.numpy(), il.numpy(), tl.numpy())
    assert abs(float(loss) - float(ref.detach())) < 1e-9
    assert np.abs(grad - logits.grad.numpy()).max() < 1e-9
    assert np.abs(grad[3]).max() == 0.0 and np.abs(grad[1, 25:]).max() == 0.0


# ------------------------------------------------------------------------------------------------
# full-depth fixtures (BASELINE configs[0]/[1] depth: 8 blocks d=256 V=1000 B=8x10 s; configs[2]: 16 blocks d=512)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_golden.npz", "cm_golden.npz"])
def test_full_depth_oracle_vs_golden(name):
    g = np.load(os.path.join(GOLD, name))
    m, x, il, targets, tl, (d, H, nb, V, ts, vs) = big_case_inputs(g)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    pnames = [n for n, _ in m.named_parameters()]
    sdr = {k: (v.clone().requires_grad_(True) if k in pnames else v) for k, v in sd.items()}
    logits = oc.forward(x, il, sdr, H, nb, training=True)
    ref = g["logits_sub"]
    assert np.abs(logits.detach()[:, ::ts, ::vs].numpy() - ref).max() < 1e-4 * float(g["logits_absmax"])
    loss = oc.ctc_loss_torch(logits, targets, il, tl)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    loss.backward()
    for pname, ref_norm in zip(g["grad_names"], g["grad_norms"]):
        gr = sdr[str(pname)].grad
        if ref_norm < 0:
            assert gr is None, pname
        else:
            assert abs(float(gr.norm()) - ref_norm) <= 2e-3 * max(ref_norm, 1e-6) + 1e-6, pname


@pytest.mark.parametrize("accum", [1, 2])
def test_trainer_oracle_vs_golden(accum):
    """oracle/trainer.py against the reference's own Trainer.train_epoch (3 batches, CPU): average loss, step
    counters, scheduler position and the parameter update itself (every 16th element of every parameter)."""
    from oracle.trainer import TrainerOracle
    g = np.load(os.path.join(GOLD, "trainer_golden.npz"))
    m, sd = _golden_model_sd()
    pnames = [n for n, _ in m.named_parameters()]
    tr = TrainerOracle(sd, pnames, GCFG["n_heads"], GCFG["n_blocks"], accumulation_steps=accum,
                       scheduler_fn=lambda o: torch.optim.lr_scheduler.OneCycleLR(o, max_lr=5e-4, total_steps=100,
                                                                                  pct_start=0.1, anneal_strategy="cos"))
    init = torch.cat([sd[n].reshape(-1) for n in pnames]).clone()
    avg, _ = tr.train_epoch(trainer_golden_batches())
    assert abs(avg - float(g["avg_loss_accum%d" % accum])) < 1e-5 * abs(avg)
    assert tr.global_step == int(g["global_step_accum%d" % accum])
    assert tr.sched.last_epoch == int(g["sched_last_epoch_accum%d" % accum])
    final = torch.cat([tr.sd[n].detach().reshape(-1) for n in pnames])
    delta = (final - init)[::16].numpy()
    ref = g["delta_sub_accum%d" % accum]
    # AdamW's first steps move every weight by ~lr * sign(g): entries whose gradient is ~0 may flip, the rest agree
    close = np.abs(delta - ref) <= 2e-5 + 1e-2 * np.abs(ref)
    assert close.mean() > 0.995, close.mean()
    cos = float((delta * ref).sum() / (np.linalg.norm(delta) * np.linalg.norm(ref)))
    assert cos > 0.999, cos
