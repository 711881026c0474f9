"""GPU: parity at the depth the benchmark runs, against fixtures written by running the reference
(tools/make_golden.py): BASELINE configs[0]/[1] = 8 blocks, d_model 256, V = 1000, batch 8 x 10 s; configs[2] =
Conformer-M, 16 blocks, d_model 512.  bf16 operand error accumulates through 40 / 80 GroupNorms; the north_star
tolerance (logits 2e-2 relative in bf16) is asserted here in two norms: max|d| / max|ref| and the Frobenius ratio."""
import os

import numpy as np
import pytest
import torch

from golden_inputs import big_case_inputs
from turkish_asr_model_b200 import _lib as L

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["c1_golden.npz", "cm_golden.npz"])
def test_full_depth_logits_loss_and_gradients(cuda, name):
    g = np.load(os.path.join(GOLD, name))
    model, x, il, targets, tl, (d, H, nb, V, ts, vs) = big_case_inputs(g)
    model = model.to(cuda).train()
    logits = model(x.to(cuda), il)
    torch.cuda.synchronize()
    got = logits.detach().float().cpu()[:, ::ts, ::vs].numpy()
    ref = g["logits_sub"]
    assert got.shape == ref.shape
    absmax = float(g["logits_absmax"])
    assert np.abs(got - ref).max() < 2e-2 * absmax, np.abs(got - ref).max() / absmax
    assert np.linalg.norm(got - ref) < 2e-2 * np.linalg.norm(ref), np.linalg.norm(got - ref) / np.linalg.norm(ref)
    loss, _, dlogits = L.ctc_loss_fwd_bwd(logits.detach(), targets.to(cuda), (il // 4).to(cuda), tl.to(cuda))
    assert abs(loss.item() - float(g["loss"])) < 2e-2 * abs(float(g["loss"]))
    logits.backward(dlogits)
    torch.cuda.synchronize()
    grads = dict(model.named_parameters())
    rels = {}
    for pname, ref_norm in zip(g["grad_names"], g["grad_norms"]):
        p = grads[str(pname)]
        if ref_norm < 0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, pname
        elif not str(pname).endswith("depthwise_conv.bias"):
            rels[str(pname)] = abs(float(p.grad.norm()) - ref_norm) / max(ref_norm, 1e-8)
    worst = max(rels, key=rels.get)
    assert rels[worst] < 8e-2, (worst, rels[worst])
    assert np.median(list(rels.values())) < 2e-2


# ------------------------------------------------------------------ fp32 operand mode (north_star: 1e-4)
def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def test_fp32_mode_c1_logits_and_loss_1e4(cuda):
    """BASELINE configs[0]: default model, 8 blocks, V=1000, batch 8 x 10 s, forward + CTC loss in fp32 mode against the
    fixture written by the reference's own fp32 CPU run: logits within 1e-4 (max norm, relative), loss within 1e-4."""
    g = np.load(os.path.join(GOLD, "c1_golden.npz"))
    model, x, il, targets, tl, (d, H, nb, V, ts, vs) = big_case_inputs(g)
    model = model.to(cuda).train().set_precision("fp32")
    with torch.no_grad():
        logits = model(x.to(cuda), il)
    assert logits.dtype == torch.float32
    got = logits.cpu()[:, ::ts, ::vs].numpy()
    ref = g["logits_sub"]
    err = np.abs(got - ref).max() / float(g["logits_absmax"])
    assert err < 1e-4, err
    loss, _, _ = L.ctc_loss_fwd_bwd(logits, targets.to(cuda), (il // 4).to(cuda), tl.to(cuda), want_grad=False)
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    with pytest.raises(L.TasrError):  # forward-only mode
        model(x.to(cuda), il)


@pytest.mark.parametrize("nterms", [3, 6])
def test_fp32_mode_small_model_train_eval_and_bn(cuda, nterms):
    """fp32 mode against the oracle on a 2-block model: train-mode forward (batch statistics, running-stat update),
    eval-mode forward, unmasked forward; 3-term and 6-term piece expansions."""
    from oracle import conformer as oc
    from turkish_asr_model_b200.model import TurkishASRModel
    torch.manual_seed(11)
    model = TurkishASRModel(80, 256, 4, 2, 100, dropout=0.0)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(12)
    x = torch.randn(3, 203, 80, generator=gen)
    il = torch.tensor([203, 150, 101])
    for b in range(3):
        x[b, il[b]:] = 0.0
    model = model.to(cuda).train().set_precision("fp32")
    model.engine().f32_terms = nterms
    bn_state = {}
    ref = oc.forward(x, il, sd, 4, 2, training=True, bn_state=bn_state)
    with torch.no_grad():
        out = model(x.to(cuda), il)
        out_nomask = model(x.to(cuda), None)
    assert _rel(out, ref) < 1e-4, _rel(out, ref)
    assert _rel(out_nomask, oc.forward(x, None, sd, 4, 2, training=True)) < 1e-4
    model.eval()
    sd_eval = dict(sd)
    # two train-mode forwards advanced the running statistics twice; restore the single-step state of the oracle
    model.load_state_dict({k: (bn_state[k] if k in bn_state else v) for k, v in model.state_dict().items()}, strict=False)
    sd_eval.update(bn_state)
    with torch.no_grad():
        oe = model(x.to(cuda), il)
    assert _rel(oe, oc.forward(x, il, sd_eval, 4, 2, training=False)) < 1e-4
