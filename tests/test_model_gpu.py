"""End-to-end parity of the drop-in TurkishASRModel (B200 kernels, bf16 operands) against the fp32 CPU
oracle (oracle/conformer.py) on identical synthetic inputs and weights.

Tolerances from BASELINE.json north_star: encoder logits 2e-2 relative (bf16 mode), CTC loss 1e-3
relative given the same logits, gradients compared at bf16 resolution."""
import numpy as np
import pytest
import torch

from oracle import conformer as oc
from turkish_asr_model_b200 import _lib as L
from turkish_asr_model_b200.model import TurkishASRModel

pytestmark = pytest.mark.gpu


def _make(cuda, d=256, H=4, nb=2, V=100, B=3, T=203, seed=0):
    torch.manual_seed(seed)
    model = TurkishASRModel(80, d, H, nb, V, dropout=0.0)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, T, 80, generator=g)
    il = torch.tensor([T, max(T - 53, 8), max(T // 2, 8)][:B])
    for b in range(B):
        x[b, il[b]:] = 0.0
    return model.to(cuda).train(), sd, x, il


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def test_state_dict_layout_and_flat_params(cuda):
    model, sd, x, il = _make(cuda)
    keys = list(model.state_dict().keys())
    assert keys == list(sd.keys())
    assert "blocks.0.norm_conv.norm.weight" in keys and "blocks.0.attn.rotary_emb.inv_freq" in keys
    assert "blocks.0.conv.batch_norm.num_batches_tracked" in keys
    model(x.to(cuda), il)  # flattens
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]) or k.endswith(("running_mean", "running_var", "num_batches_tracked")), k


@pytest.mark.parametrize("masked", [True, False])
def test_forward_logits_parity(cuda, masked):
    model, sd, x, il = _make(cuda)
    lengths = il if masked else None
    ref = oc.forward(x, lengths, sd, 4, 2, training=True)
    out = model(x.to(cuda), lengths)
    torch.cuda.synchronize()
    assert out.shape == ref.shape == (3, oc.encoder_frames(203), 100)
    assert _rel(out, ref) < 2e-2
    # BatchNorm running statistics were updated like the reference does in train mode
    bn_state = {}
    oc.forward(x, lengths, sd, 4, 2, training=True, bn_state=bn_state)
    for k, v in bn_state.items():
        assert _rel(model.state_dict()[k], v) < 2e-2, k


def test_eval_forward_parity(cuda):
    model, sd, x, il = _make(cuda)
    model.eval()
    with torch.no_grad():
        out = model(x.to(cuda), il)
    ref = oc.forward(x, il, sd, 4, 2, training=False)
    assert _rel(out, ref) < 2e-2


def test_backward_parity(cuda):
    model, sd, x, il = _make(cuda)
    B, V = 3, 100
    g = torch.Generator().manual_seed(5)
    targets = torch.randint(1, V, (B, 10), generator=g)
    tl = torch.tensor([10, 7, 4])
    # oracle: fp32 autograd through the functional restatement + torch CTC (the reference's own loss call)
    pnames = {n for n, _ in model.named_parameters()}
    sdr = {k: (v.clone().requires_grad_(True) if k in pnames else v) for k, v in sd.items()}
    logits_ref = oc.forward(x, il, sdr, 4, 2, training=True)
    loss_ref = oc.ctc_loss_torch(logits_ref, targets, il, tl)
    loss_ref.backward()
    # device
    logits = model(x.to(cuda), il)
    loss, nll, dlogits = L.ctc_loss_fwd_bwd(logits.detach(), targets.to(cuda), (il // 4).to(cuda), tl.to(cuda))
    logits.backward(dlogits)
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    worst = {}
    for name, p in model.named_parameters():
        gref = sdr[name].grad
        if gref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name  # dead norm_conv params
            continue
        assert p.grad is not None, name
        if name.endswith("depthwise_conv.bias"):
            # a bias in front of BatchNorm has an exactly-zero gradient; both sides only hold rounding noise
            wg = dict(model.named_parameters())[name.replace(".bias", ".weight")].grad
            assert float(p.grad.abs().max()) < 1e-2 * float(wg.abs().max()), name
            continue
        worst[name] = _rel(p.grad, gref)
    bad = {k: v for k, v in worst.items() if v > 6e-2}
    assert not bad, bad
    assert np.median(list(worst.values())) < 2e-2


def test_conformer_m_scale_parity(cuda):
    """BASELINE configs[2] shape family: d_model 512, 8 heads (GroupNorm 32 groups x 16 channels), 2 blocks."""
    model, sd, x, il = _make(cuda, d=512, H=8, nb=2, V=1000, B=2, T=171, seed=3)
    ref = oc.forward(x, il, sd, 8, 2, training=True)
    out = model(x.to(cuda), il)
    assert out.shape == ref.shape
    assert _rel(out, ref) < 2e-2
    g = torch.Generator().manual_seed(9)
    targets = torch.randint(1, 1000, (2, 8), generator=g)
    tl = torch.tensor([8, 5])
    pnames = {n for n, _ in model.named_parameters()}
    sdr = {k: (v.clone().requires_grad_(True) if k in pnames else v) for k, v in sd.items()}
    loss_ref = oc.ctc_loss_torch(oc.forward(x, il, sdr, 8, 2, training=True), targets, il, tl)
    loss_ref.backward()
    loss, _, dlogits = L.ctc_loss_fwd_bwd(out.detach(), targets.to(cuda), (il // 4).to(cuda), tl.to(cuda))
    out.backward(dlogits)
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    rels = []
    for name, p in model.named_parameters():
        gref = sdr[name].grad
        if gref is None or name.endswith("depthwise_conv.bias"):
            continue
        rels.append(_rel(p.grad, gref))
    assert max(rels) < 8e-2 and np.median(rels) < 2e-2


def test_long_form_inference_and_greedy_decode(cuda):
    """BASELINE configs[3] shape family: 60 s utterances (T = 6001 mel frames -> T' = 1501), eval mode, padding
    mask, greedy CTC decode; logits vs the oracle, token ids bit-exact on identical logits."""
    from turkish_asr_model_b200.utils.decoding import GreedyDecoder
    torch.manual_seed(4)
    model = TurkishASRModel(80, 256, 4, 2, 200, dropout=0.1)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(cuda).eval()
    g = torch.Generator().manual_seed(5)
    T = 6001
    x = torch.randn(2, T, 80, generator=g)
    il = torch.tensor([6001, 4321])
    x[1, 4321:] = 0
    with torch.no_grad():
        out = model(x.to(cuda), il)
    ref = oc.forward(x, il, sd, 4, 2, training=False)
    assert out.shape == ref.shape == (2, 1501, 200)
    assert oc.encoder_frames(T) == 1501 and int(il[0]) // 4 == 1500  # L' = T // 4 <= T' (SURVEY finding 5)
    assert _rel(out, ref) < 2e-2
    lengths = il // 4
    ids_ref, toks_ref = oc.greedy_ids(out.float().cpu(), lengths)
    toks = GreedyDecoder(None, blank_id=0).decode_ids_batch(out, lengths)
    assert toks == toks_ref
    with torch.no_grad():  # no mask at all (reference inference path, inference.py:117)
        out2 = model(x.to(cuda), None)
    assert _rel(out2, oc.forward(x, None, sd, 4, 2, training=False)) < 2e-2


def test_standalone_submodules_match_oracle(cuda):
    """ConformerBlock / sub-modules called on their own (reference API, model/conformer.py:90-135) with autograd."""
    from turkish_asr_model_b200.model import ConformerBlock
    torch.manual_seed(1)
    blk = ConformerBlock(256, 4, dropout=0.0)
    sd = {"b." + k: v.detach().clone() for k, v in blk.state_dict().items()}
    blk = blk.to(cuda).train()
    g = torch.Generator().manual_seed(2)
    B, T = 2, 90
    x = torch.randn(B, T, 256, generator=g)
    lens = torch.tensor([90, 51])
    mask = (torch.arange(T)[None, :] < lens[:, None])[:, None, None, :]
    pnames = {"b." + n for n, _ in blk.named_parameters()}
    sdr = {k: (v.clone().requires_grad_(True) if k in pnames else v) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    ref = oc.block(xr, sdr, "b.", 4, lens, training=True)
    dy = torch.randn(B, T, 256, generator=g)
    ref.backward(dy)
    xd = x.to(cuda).requires_grad_(True)
    out = blk(xd, mask=mask.to(cuda))
    assert _rel(out, ref.detach()) < 2e-2
    out.backward(dy.to(cuda))
    assert _rel(xd.grad, xr.grad) < 5e-2
    rels = []
    for name, p in blk.named_parameters():
        gref = sdr["b." + name].grad
        if gref is None:
            assert p.grad is None
            continue
        if name.endswith("depthwise_conv.bias"):
            continue
        rels.append(_rel(p.grad, gref))
    assert max(rels) < 8e-2 and np.median(rels) < 2e-2
    # individual modules
    y = blk.ff1(xd.detach())
    assert _rel(y, oc.swiglu_ff(x, sd, "b.ff1.")) < 2e-2
    a, w = blk.attn(xd.detach(), xd.detach(), xd.detach(), mask=mask.to(cuda))
    assert w is None and _rel(a, oc.mqa_attention(x, sd, "b.attn.", 4, lens)) < 2e-2
    n = blk.norm_ff1(xd.detach())
    assert _rel(n, oc.group_norm_tokens(x, sd["b.norm_ff1.norm.weight"], sd["b.norm_ff1.norm.bias"])) < 1e-4


def test_edge_cases_short_and_fully_masked(cuda):
    """Ragged / degenerate inputs: B = 1, T' = 3, one utterance with L' = T // 4 = 0 (every key masked -> the
    attention contributes 0 for that utterance by definition) and an infeasible CTC sample."""
    torch.manual_seed(8)
    model = TurkishASRModel(80, 128, 2, 1, 20, dropout=0.0).to(cuda).train()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 9, 80, generator=g)
    out = model(x.to(cuda), torch.tensor([9]))
    assert out.shape == (1, oc.encoder_frames(9), 20) and torch.isfinite(out.float()).all()
    x2 = torch.randn(2, 40, 80, generator=g)
    il = torch.tensor([40, 3])  # second utterance: L' = 0
    out2 = model(x2.to(cuda), il)
    assert torch.isfinite(out2.float()).all()
    targets = torch.randint(1, 20, (2, 4), generator=g)
    tl = torch.tensor([4, 4])  # sample 1 is infeasible (4 labels, 0 frames): zero loss and gradient (zero_infinity)
    loss, nll, dl = L.ctc_loss_fwd_bwd(out2.detach(), targets.to(cuda), (il // 4).to(cuda), tl.to(cuda))
    assert torch.isfinite(loss).all() and torch.isinf(nll[1]) and float(dl[1].float().abs().max()) == 0.0
    out2.backward(dl)
    for n, p in model.named_parameters():
        assert p.grad is None or torch.isfinite(p.grad).all(), n
