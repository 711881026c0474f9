"""Per-kernel parity of the C-ABI entry points against the CPU oracle (oracle/), on seeded inputs.

Tolerances: bf16 outputs are compared at bf16 resolution (relative 2^-8 of the tensor scale);
fp32 outputs at 1e-4..1e-5; CTC loss/gradient at the north_star's 1e-3 relative; integers bit-exact.
"""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import conformer as oc
from oracle import ctc as octc
from turkish_asr_model_b200 import _lib as L

pytestmark = pytest.mark.gpu


def rel_err(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------ GroupNorm
@pytest.mark.parametrize("B,T,d", [(3, 251, 256), (2, 126, 512), (1, 7, 256), (16, 45, 256)])
@pytest.mark.parametrize("out_bf16", [True, False])
def test_groupnorm_fwd_bwd(cuda, B, T, d, out_bf16):
    g = torch.Generator().manual_seed(B * 100 + T)
    x = torch.randn(B, T, d, generator=g) * 2 + 0.5
    gamma = torch.randn(d, generator=g)
    beta = torch.randn(d, generator=g)
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = oc.group_norm_tokens(xd, gd, bd, 32)
    y, stats = L.groupnorm_fwd(x.to(cuda), 32, gamma.to(cuda), beta.to(cuda), out_bf16=out_bf16)
    torch.cuda.synchronize()
    assert rel_err(y, ref.detach()) < (6e-3 if out_bf16 else 2e-5)
    dy = torch.randn(B, T, d, generator=g)
    dy_dev = bf(dy).to(cuda) if out_bf16 else dy.to(cuda)
    dy_used = dy_dev.double().cpu()
    ref.backward(dy_used)
    dres0 = torch.randn(B, T, d, generator=g)
    dres = dres0.clone().to(cuda)
    dgamma = torch.zeros(d, device=cuda)
    dbeta = torch.zeros(d, device=cuda)
    L.groupnorm_bwd(dy_dev, x.to(cuda), 32, stats, gamma.to(cuda), dres, True, dgamma, dbeta)
    torch.cuda.synchronize()
    assert rel_err(dres, dres0.double() + xd.grad) < 2e-5
    assert rel_err(dgamma, gd.grad) < 2e-5
    assert rel_err(dbeta, bd.grad) < 2e-5
    dres2 = torch.full((B, T, d), float("nan"), device=cuda)
    L.groupnorm_bwd(dy_dev, x.to(cuda), 32, stats, gamma.to(cuda), dres2, False, None, None)
    torch.cuda.synchronize()
    assert rel_err(dres2, xd.grad) < 2e-5
    # fused cast: same bits as the stand-alone cast kernel (scale + counter-based dropout mask) on the new dres
    dres3 = torch.empty(B, T, d, device=cuda)
    fused = L.groupnorm_bwd(dy_dev, x.to(cuda), 32, stats, gamma.to(cuda), dres3, False, None, None, cast=(0.5, 0.1, 1234))
    sep = L.cast_bf16(dres3, alpha=0.5, drop_p=0.1, seed=1234)
    torch.cuda.synchronize()
    assert torch.equal(fused, sep) and torch.equal(dres3, dres2)


# ------------------------------------------------------------------ depthwise conv + BatchNorm + SiLU
@pytest.mark.parametrize("B,T,d", [(3, 251, 256), (2, 70, 512), (1, 5, 256), (16, 140, 256), (12, 33, 128)])
def test_dwconv_bn_silu(cuda, B, T, d):
    g = torch.Generator().manual_seed(T)
    ab = bf(torch.randn(B, T, 2 * d, generator=g))
    u = bf(F.glu(ab.float(), dim=-1))
    w = torch.randn(d, 31, generator=g) / 5
    bias = torch.randn(d, generator=g)
    gamma, beta = torch.randn(d, generator=g), torch.randn(d, generator=g)
    rm, rv = torch.zeros(d), torch.ones(d)
    # oracle (float64 on the same bf16-rounded inputs)
    ud = u.double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), bias.double().requires_grad_(True)
    gd, btd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    conv = F.conv1d(ud.transpose(1, 2), wd.unsqueeze(1), bd, padding=15, groups=d)
    conv_q = conv.detach().to(torch.bfloat16).double()  # the kernel stores bf16; BN sees the rounded values
    conv_in = conv + (conv_q - conv.detach())
    rmd, rvd = rm.double().clone(), rv.double().clone()
    s_ref = F.silu(F.batch_norm(conv_in, rmd, rvd, gd, btd, training=True, momentum=0.1, eps=1e-5)).transpose(1, 2)
    # device
    wdev, part = L.dwconv_fwd(u.to(cuda), w.to(cuda), bias.to(cuda))
    rm_d, rv_d = rm.to(cuda), rv.to(cuda)
    nbt = torch.zeros((), dtype=torch.int64, device=cuda)
    stats = L.bn_finalize(part, d, B * T, 1e-5, 0.1, True, rm_d, rv_d, nbt)
    s = L.bn_silu_fwd(wdev, stats, gamma.to(cuda), beta.to(cuda))
    torch.cuda.synchronize()
    assert rel_err(wdev, conv.detach().transpose(1, 2)) < 6e-3
    assert rel_err(s, s_ref.detach()) < 1e-2
    assert rel_err(rm_d, rmd) < 1e-4 and rel_err(rv_d, rvd) < 1e-4 and int(nbt) == 1
    # backward
    ds = bf(torch.randn(B, T, d, generator=g))
    s_ref.backward(ds.double())
    dgamma, dbeta = torch.zeros(d, device=cuda), torch.zeros(d, device=cuda)
    dw = L.bn_silu_bwd(ds.to(cuda), wdev, stats, gamma.to(cuda), beta.to(cuda), dgamma, dbeta)
    dweight, dbias = torch.zeros(d, 31, device=cuda), torch.zeros(d, device=cuda)
    du = L.dwconv_bwd(dw, u.to(cuda), None, w.to(cuda), dweight, dbias)
    torch.cuda.synchronize()
    assert rel_err(dgamma, gd.grad) < 1e-2 and rel_err(dbeta, btd.grad) < 1e-2
    assert rel_err(du, ud.grad) < 2e-2
    assert rel_err(dweight, wd.grad) < 2e-2
    # fused GLU backward
    abd = ab.double().requires_grad_(True)
    (F.glu(abd, dim=-1) * ud.grad).sum().backward()
    dweight2, dbias2 = torch.zeros(d, 31, device=cuda), torch.zeros(d, device=cuda)
    dab = L.dwconv_bwd(dw, u.to(cuda), ab.to(cuda), w.to(cuda), dweight2, dbias2)
    torch.cuda.synchronize()
    assert rel_err(dab, abd.grad) < 2e-2


def test_bn_eval_mode(cuda):
    d, M = 256, 300
    g = torch.Generator().manual_seed(1)
    w = bf(torch.randn(M, d, generator=g))
    rm, rv = torch.randn(d, generator=g), torch.rand(d, generator=g) + 0.5
    gamma, beta = torch.randn(d, generator=g), torch.randn(d, generator=g)
    stats = L.bn_finalize(None, d, M, 1e-5, 0.1, False, rm.to(cuda), rv.to(cuda), None)
    s = L.bn_silu_fwd(w.to(cuda), stats, gamma.to(cuda), beta.to(cuda))
    ref = F.silu(F.batch_norm(w.double(), rm.double(), rv.double(), gamma.double(), beta.double(), training=False, eps=1e-5))
    assert rel_err(s, ref) < 1e-2


# ------------------------------------------------------------------ elementwise
def test_cast_colsum_rope(cuda):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1000, 257, generator=g)
    y = L.cast_bf16(x.to(cuda), alpha=0.5)
    assert torch.equal(y.cpu(), bf(x * 0.5))
    yd = L.cast_bf16(x.to(cuda), alpha=1.0, drop_p=0.25, seed=77)
    keep = (yd != 0).float().mean().item()
    assert abs(keep - 0.75) < 0.01
    nz = yd.cpu().float() != 0
    assert torch.allclose(yd.cpu().float()[nz], bf(x / 0.75).float()[nz], rtol=1e-2)
    a = bf(torch.randn(3001, 1000, generator=g))
    out = torch.ones(1000, device=cuda)
    L.colsum_add(a.to(cuda), out)
    assert rel_err(out, a.double().sum(0) + 1) < 1e-5
    # rope on (M, d+128) with T positions
    B, T, H = 2, 77, 4
    d = H * 64
    qkv = bf(torch.randn(B * T, d + 128, generator=g))
    cos, sin = oc.rope_tables(T, 64)
    cs = torch.stack([cos[:, :32], sin[:, :32]], dim=-1).contiguous()
    dev = qkv.clone().to(cuda)
    L.rope_inplace(dev, T, d + 64, cs.to(cuda))
    q = qkv[:, :d].double().view(B, T, H, 64).transpose(1, 2)
    k = qkv[:, d:d + 64].double().view(B, T, 1, 64).transpose(1, 2)
    qr = oc.apply_rope(q, cos.double(), sin.double()).transpose(1, 2).reshape(B * T, d)
    kr = oc.apply_rope(k, cos.double(), sin.double()).transpose(1, 2).reshape(B * T, 64)
    assert rel_err(dev[:, :d], qr) < 6e-3 and rel_err(dev[:, d:d + 64], kr) < 6e-3
    assert torch.equal(dev[:, d + 64:].cpu(), qkv[:, d + 64:])
    L.rope_inplace(dev, T, d + 64, cs.to(cuda), inverse=True)
    assert rel_err(dev, qkv) < 1.5e-2


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,T,H,lens", [(2, 251, 4, [250, 100]), (3, 126, 4, None), (1, 400, 8, [390]), (2, 64, 4, [64, 1])])
def test_mqa_attention_fwd_bwd(cuda, B, T, H, lens):
    d = H * 64
    g = torch.Generator().manual_seed(T + H)
    qkv = bf(torch.randn(B * T, d + 128, generator=g))
    kl = None if lens is None else torch.tensor(lens, dtype=torch.int64)
    x = qkv.double().requires_grad_(True)
    q = x[:, :d].view(B, T, H, 64).transpose(1, 2)
    k = x[:, d:d + 64].view(B, T, 1, 64).transpose(1, 2)
    v = x[:, d + 64:].view(B, T, 1, 64).transpose(1, 2)
    ref = oc.mqa_core(q, k, v, kl).transpose(1, 2).reshape(B * T, d)
    kl_dev = None if kl is None else kl.to(cuda)
    ctx, lse2 = L.mqa_fwd(qkv.to(cuda), B, T, H, d, kl_dev)
    torch.cuda.synchronize()
    assert rel_err(ctx, ref.detach()) < 1.2e-2
    dctx = bf(torch.randn(B * T, d, generator=g))
    ref.backward(dctx.double())
    dqkv = L.mqa_bwd(qkv.to(cuda), ctx, dctx.to(cuda), lse2, B, T, H, d, kl_dev, None)
    torch.cuda.synchronize()
    gref = x.grad
    assert rel_err(dqkv[:, :d], gref[:, :d]) < 2e-2
    assert rel_err(dqkv[:, d:d + 64], gref[:, d:d + 64]) < 2e-2
    assert rel_err(dqkv[:, d + 64:], gref[:, d + 64:]) < 2e-2


def test_mqa_attention_dropout_mask_consistent(cuda):
    """Dropout on the probabilities: the mask the forward applied is read back exactly (one-hot V), then forward and
    backward are checked against float64 autograd with that same mask (model/attention.py:239 dropout_p)."""
    B, T, H, p_drop, seed = 2, 128, 4, 0.25, 99
    d = H * 64
    lens = [128, 77]
    g = torch.Generator().manual_seed(5)
    qkv = bf(torch.randn(B * T, d + 128, generator=g) * 0.7)
    kl = torch.tensor(lens, dtype=torch.int64)
    kl_dev = kl.to(cuda)
    pd = torch.zeros(B, H, T, T, dtype=torch.float64)
    for blk in range(T // 64):
        probe = qkv.clone().view(B, T, d + 128)
        probe[:, :, d + 64:] = 0
        for j in range(64):
            probe[:, 64 * blk + j, d + 64 + j] = 1.0
        c, _ = L.mqa_fwd(probe.view(B * T, d + 128).to(cuda), B, T, H, d, kl_dev, p_drop, seed)
        pd[:, :, :, 64 * blk:64 * blk + 64] = c.double().cpu().view(B, T, H, 64).transpose(1, 2)
    mask = (pd > 0).double()
    frac = 1.0 - mask[0].mean().item()       # utterance 0 has every key valid
    assert abs(frac - p_drop) < 0.02
    x = qkv.double().requires_grad_(True)
    q = x[:, :d].view(B, T, H, 64).transpose(1, 2)
    k = x[:, d:d + 64].view(B, T, 1, 64).transpose(1, 2)
    v = x[:, d + 64:].view(B, T, 1, 64).transpose(1, 2)
    sc = (q @ k.transpose(-1, -2)) / 8.0
    keymask = torch.arange(T)[None, :] >= kl[:, None]
    sc = sc.masked_fill(keymask[:, None, None, :], float("-inf"))
    pr = torch.softmax(sc, dim=-1) * mask / (1.0 - p_drop)
    ref = (pr @ v).transpose(1, 2).reshape(B * T, d)
    ctx, lse2 = L.mqa_fwd(qkv.to(cuda), B, T, H, d, kl_dev, p_drop, seed)
    assert rel_err(ctx, ref.detach()) < 1.5e-2
    dctx = bf(torch.randn(B * T, d, generator=g))
    ref.backward(dctx.double())
    dqkv = L.mqa_bwd(qkv.to(cuda), ctx, dctx.to(cuda), lse2, B, T, H, d, kl_dev, None, p_drop, seed)
    torch.cuda.synchronize()
    gref = x.grad
    assert rel_err(dqkv[:, :d], gref[:, :d]) < 2.5e-2
    assert rel_err(dqkv[:, d:d + 64], gref[:, d:d + 64]) < 2.5e-2
    assert rel_err(dqkv[:, d + 64:], gref[:, d + 64:]) < 2.5e-2


# ------------------------------------------------------------------ subsampler
@pytest.mark.parametrize("B,T,d", [(2, 203, 256), (1, 130, 512), (3, 9, 256), (2, 64, 128)])
def test_implicit_conv_subsampler(cuda, B, T, d):
    """conv1_fwd + conv2_fwd / dgrad / wgrad + conv1_bwd (implicit GEMM) vs F.conv2d autograd in float64."""
    Fm = 80
    g = torch.Generator().manual_seed(T + d)
    x = torch.randn(B, T, Fm, generator=g)
    w1 = torch.randn(d, 1, 3, 3, generator=g) / 3
    b1 = torch.randn(d, generator=g) / 3
    w2 = torch.randn(d, d, 3, 3, generator=g) / (3 * d ** 0.5)
    b2 = torch.randn(d, generator=g) / 3
    T1, F1, T2, F2 = L.sub_dims(T, Fm)
    w1d, b1d = w1.double().requires_grad_(True), b1.double().requires_grad_(True)
    w2q = bf(w2).double().requires_grad_(True)
    b2d = b2.double().requires_grad_(True)
    y1r = F.silu(F.conv2d(x.double().unsqueeze(1), w1d, b1d, stride=2, padding=1))        # (B,d,T1,F1)
    y1q = y1r + (y1r.detach().to(torch.bfloat16).double() - y1r.detach())                  # bf16-rounded operand
    z2r = F.conv2d(y1q, w2q, b2d, stride=2, padding=1)                                     # (B,d,T2,F2)
    # device
    y1 = L.conv1_fwd(x.to(cuda), w1.to(cuda), b1.to(cuda))
    w2p = L.pack_weight_remap(w2.view(d, 9 * d).to(cuda), 9)
    z2, y2 = L.conv2_fwd(y1, T, Fm, w2p, b2.to(cuda))
    torch.cuda.synchronize()
    assert y1.shape == (B, T1, F1, d)
    assert rel_err(y1, y1r.detach().permute(0, 2, 3, 1)) < 6e-3
    z2_ref = z2r.detach().permute(0, 2, 3, 1).reshape(B * T2 * F2, d)
    assert rel_err(z2, z2_ref) < 1e-2
    assert rel_err(y2, F.silu(z2_ref)) < 1.2e-2
    # backward
    dz2 = bf(torch.randn(B * T2 * F2, d, generator=g))
    z2r.backward(dz2.double().view(B, T2, F2, d).permute(0, 3, 1, 2))
    dy1 = L.conv2_dgrad(dz2.to(cuda), B, T, Fm, w2p)
    dw2 = torch.zeros(d, d, 3, 3, device=cuda)
    L.conv2_wgrad(dz2.to(cuda), y1, T, Fm, dw2)
    torch.cuda.synchronize()
    assert rel_err(dw2, w2q.grad) < 1.5e-2
    # dy1 reference: gradient w.r.t. the conv2 input
    y1leaf = y1q.detach().requires_grad_(True)
    F.conv2d(y1leaf, w2q.detach(), b2d.detach(), stride=2, padding=1).backward(dz2.double().view(B, T2, F2, d).permute(0, 3, 1, 2))
    assert rel_err(dy1, y1leaf.grad.permute(0, 2, 3, 1)) < 1.2e-2
    dw1, db1 = torch.zeros(d, 1, 3, 3, device=cuda), torch.zeros(d, device=cuda)
    L.conv1_bwd(dy1, x.to(cuda), w1.to(cuda), b1.to(cuda), dw1, db1)
    torch.cuda.synchronize()
    assert rel_err(dw1, w1d.grad) < 1.5e-2 and rel_err(db1, b1d.grad) < 1.5e-2


def test_pack_weight_remap(cuda):
    g = torch.Generator().manual_seed(3)
    w = torch.randn(256, 128, 3, 3, generator=g)
    out = L.pack_weight_remap(w.view(256, -1).to(cuda), 9)
    ref = bf(w.permute(0, 2, 3, 1).reshape(256, -1))
    assert torch.equal(out.cpu(), ref)
    wi = torch.randn(64, 32 * 20, generator=g)  # (n, c*20 + f) -> (n, f*32 + c)
    out = L.pack_weight_remap(wi.to(cuda), 20)
    assert torch.equal(out.cpu(), bf(wi.view(64, 32, 20).permute(0, 2, 1).reshape(64, -1)))


# ------------------------------------------------------------------ CTC
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_ctc_loss_and_grad(cuda, dtype):
    B, T, V, Smax = 5, 60, 50, 12
    g = torch.Generator().manual_seed(4)
    logits = (torch.randn(B, T, V, generator=g) * 2).to(dtype)
    targets = torch.randint(1, V, (B, Smax), generator=g)
    targets[1, :6] = torch.tensor([7, 7, 7, 3, 3, 9])  # repeats need separating blanks
    tl = torch.tensor([12, 6, 0, 12, 3])
    il = torch.tensor([60, 31, 10, 11, 0])  # sample 3 is infeasible (12 labels, 11 frames); sample 4 empty input
    loss, nll, grad = octc.ctc_loss_and_grad(logits.double().numpy(), targets.numpy(), il.numpy(), tl.numpy())
    l_dev, nll_dev, g_dev = L.ctc_loss_fwd_bwd(logits.to(cuda), targets.to(cuda), il.to(cuda), tl.to(cuda))
    torch.cuda.synchronize()
    assert abs(l_dev.item() - loss) <= 1e-3 * abs(loss)
    fin = np.isfinite(nll)
    assert np.allclose(nll_dev.cpu().numpy()[fin], nll[fin], rtol=1e-4)
    assert np.all(np.isinf(nll_dev.cpu().numpy()[~fin]))
    gd = g_dev.double().cpu().numpy()
    tol = 1e-3 if dtype == torch.float32 else 8e-3
    assert np.abs(gd - grad).max() <= tol * np.abs(grad).max()
    assert np.all(gd[1, 31:] == 0) and np.all(gd[3] == 0) and np.all(gd[4] == 0)
    # cross-check the oracle's own convention against torch's CTCLoss (the reference's call)
    lt = oc.ctc_loss_torch(logits.double(), targets, il * 4, tl)
    assert abs(lt.item() - loss) < 1e-9


def test_ctc_c2_shape(cuda):
    B, T, V = 16, 376, 1000
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(B, T, V, generator=g).to(torch.bfloat16)
    tl = torch.randint(20, 61, (B,), generator=g)
    targets = torch.randint(1, V, (B, 60), generator=g)
    il = torch.randint(250, 377, (B,), generator=g)
    loss, nll, grad = octc.ctc_loss_and_grad(logits.double().numpy(), targets.numpy(), il.numpy(), tl.numpy())
    l_dev, nll_dev, g_dev = L.ctc_loss_fwd_bwd(logits.to(cuda), targets.to(cuda), il.to(cuda), tl.to(cuda))
    assert abs(l_dev.item() - loss) <= 1e-3 * abs(loss)
    assert np.abs(g_dev.double().cpu().numpy() - grad).max() <= 8e-3 * np.abs(grad).max()
    # rows of the gradient sum to zero (softmax minus a distribution)
    assert g_dev.float().sum(-1).abs().max().item() < 1e-3
    # north_star: CTC gradients within 1e-3 relative.  The arithmetic is fp32 for bf16 logits too; asked for unrounded,
    # the gradient meets 1e-3 (the 8e-3 above is only the bf16 storage format of the GEMM operand)
    _, _, g32 = L.ctc_loss_fwd_bwd(logits.to(cuda), targets.to(cuda), il.to(cuda), tl.to(cuda), grad_dtype=torch.float32)
    assert g32.dtype == torch.float32
    assert np.abs(g32.double().cpu().numpy() - grad).max() <= 1e-3 * np.abs(grad).max()


def test_ctc_long_targets(cuda):
    """Long-form transcripts: more than 255 labels per utterance (the one-thread-per-state kernel covers 2*255+1 states;
    beyond that the strided kernel takes over), fp32 logits, ragged lengths, one NaN-free infeasible sample."""
    B, T, V = 3, 700, 40
    g = torch.Generator().manual_seed(8)
    logits = torch.randn(B, T, V, generator=g)
    tl = torch.tensor([300, 260, 699])
    targets = torch.randint(1, V, (B, 699), generator=g)
    il = torch.tensor([700, 640, 700])  # sample 2: 699 labels with repeats cannot fit 700 frames -> infeasible
    # checker: torch's own CTCLoss + autograd in fp64 on the CPU (the reference's call; oracle/ctc.py is pinned to it by
    # tests/test_oracle_golden.py and is a pure-Python loop, too slow at this size)
    lg = logits.double().requires_grad_(True)
    lp = torch.nn.functional.log_softmax(lg, dim=-1).transpose(0, 1)
    nll = torch.nn.functional.ctc_loss(lp, targets, il, tl, blank=0, reduction="none", zero_infinity=False)
    loss_ref = torch.nn.functional.ctc_loss(lp, targets, il, tl, blank=0, reduction="mean", zero_infinity=True)
    loss_ref.backward()
    l_dev, nll_dev, g_dev = L.ctc_loss_fwd_bwd(logits.to(cuda), targets.to(cuda), il.to(cuda), tl.to(cuda))
    torch.cuda.synchronize()
    assert torch.isinf(nll[2]) and np.isinf(nll_dev[2].item())
    assert abs(l_dev.item() - loss_ref.item()) <= 1e-3 * abs(loss_ref.item())
    assert np.allclose(nll_dev.cpu().numpy()[:2], nll.detach().numpy()[:2], rtol=1e-4)
    grad = lg.grad.numpy()
    assert np.abs(g_dev.double().cpu().numpy() - grad).max() <= 1e-3 * np.abs(grad).max()
    assert np.all(g_dev[2].cpu().numpy() == 0)


# ------------------------------------------------------------------ optimizer / decode
def test_clip_adamw(cuda):
    n = 100003
    g = torch.Generator().manual_seed(6)
    p, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 3
    m, v = torch.randn(n, generator=g) * 0.1, torch.rand(n, generator=g) * 0.1
    step, lr = 7, 5e-4
    total, coef = oc.clip_coef([gr], 1.0)
    pr, mr, vr = oc.adamw_step(p.double(), gr.double() * coef, m.double(), v.double(), step, lr=lr)
    hyper = torch.tensor([lr, 0.9, 0.999, 1e-8, 1e-6, 1 - 0.9 ** step, 1 - 0.999 ** step, 1.0, 1.0])
    pd, gd_, md, vd = p.to(cuda), gr.to(cuda), m.to(cuda), v.to(cuda)
    shadow = torch.empty(n, dtype=torch.bfloat16, device=cuda)
    sumsq = torch.zeros(1, dtype=torch.float64, device=cuda)
    norm = torch.zeros(1, device=cuda)
    L.grad_sumsq(gd_, sumsq)
    L.clip_adamw(pd, gd_, md, vd, shadow, hyper.to(cuda), sumsq, norm)
    torch.cuda.synchronize()
    assert abs(norm.item() - float(total)) < 1e-3 * float(total)
    assert rel_err(pd, pr) < 1e-6 and rel_err(md, mr) < 1e-5 and rel_err(vd, vr) < 1e-5
    assert torch.equal(shadow.cpu(), bf(pd.cpu()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_argmax_collapse(cuda, dtype):
    B, T, V = 4, 1501, 1000
    g = torch.Generator().manual_seed(8)
    logits = torch.randn(B, T, V, generator=g)
    # long runs + blanks so that collapsing matters
    path = torch.randint(0, 6, (B, T), generator=g)
    path[:, 0:1500:3] = path[:, 1:1501:3]
    logits.scatter_(2, path.unsqueeze(-1), 30.0)
    logits = logits.to(dtype)
    logits[0, 5, :] = 1.0  # full tie -> first index
    lengths = torch.tensor([1501, 1000, 1, 777])
    ids_ref, toks_ref = oc.greedy_ids(logits.float(), lengths, blank=0)
    ids, tokens, out_len = L.argmax_collapse(logits.to(cuda), lengths.to(cuda), blank=0)
    torch.cuda.synchronize()
    assert torch.equal(ids.cpu(), ids_ref)  # bit-exact
    for b in range(B):
        n = int(out_len[b])
        assert tokens[b, :n].tolist() == toks_ref[b]
        assert torch.all(tokens[b, n:] == -1)
