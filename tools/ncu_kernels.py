"""One launch of every hot kernel at the bench's largest batch shape (B=64, 13.25 s: T=1325 mel frames, T'=332, d=256),
for `ncu --set full` (each kernel is replayed ~40x, so nothing is repeated here).  Prints nothing but 'done'."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turkish_asr_model_b200 import _lib as L  # noqa: E402
from turkish_asr_model_b200.data.preprocessing import AudioPreprocessor  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, T, F, d, H, G, dff, V = 64, 1325, 80, 256, 4, 32, 1024, 1000
T1, F1, Tp, F2 = L.sub_dims(T, F)
M = B * Tp
g = torch.Generator(device=dev).manual_seed(0)


def rb(*s):
    return (torch.randn(*s, generator=g, device=dev) * 0.5).to(torch.bfloat16)


# mel
pre = AudioPreprocessor(device="cuda")
wav = 0.1 * torch.randn(B, (T - 1) * 160, generator=g, device=dev)
ns = torch.full((B,), (T - 1) * 160, dtype=torch.int64)
feats, frames = pre.extract_features_batch(wav, ns, T)
# subsampler
w1 = torch.randn(d, 1, 3, 3, device=dev) * 0.3
b1 = torch.randn(d, device=dev) * 0.1
y1 = L.conv1_fwd(feats, w1, b1)
w2p = rb(d, 9 * d) * 0.05
b2 = torch.randn(d, device=dev) * 0.1
z2, y2 = L.conv2_fwd(y1, T, F, w2p, b2)
dz2 = rb(M * F2, d) * 0.02
dy1 = L.conv2_dgrad(dz2, B, T, F, w2p)
gw2 = torch.zeros(d, d, 3, 3, device=dev)
L.conv2_wgrad(dz2, y1, T, F, gw2)
dw1, db1 = torch.zeros_like(w1), torch.zeros_like(b1)
L.conv1_bwd(dy1, feats, w1, b1, dw1, db1)
del y1, z2, y2, dz2, dy1
# GroupNorm
x = torch.randn(B, Tp, d, generator=g, device=dev)
gamma, beta = torch.randn(d, device=dev), torch.randn(d, device=dev)
xn, st = L.groupnorm_fwd(x, G, gamma, beta)
dyb = rb(B, Tp, d)
dres = torch.randn(B, Tp, d, generator=g, device=dev)
dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
L.groupnorm_bwd(dyb, x, G, st, gamma, dres, True, dg, db, cast=(0.5, 0.1, 7))
# depthwise conv + BatchNorm
u = rb(B, Tp, d)
ab = rb(B, Tp, 2 * d)
dwt, dwb = torch.randn(d, 31, device=dev) * 0.2, torch.randn(d, device=dev) * 0.1
w, part = L.dwconv_fwd(u, dwt, dwb)
gdw, gdb = torch.zeros_like(dwt), torch.zeros_like(dwb)
L.dwconv_bwd(u, u, ab, dwt, gdw, gdb)
rm, rv, nb = torch.zeros(d, device=dev), torch.ones(d, device=dev), torch.zeros((), dtype=torch.long, device=dev)
bst = L.bn_finalize(part, d, M, 1e-5, 0.1, True, rm, rv, nb)
L.bn_silu_fwd(w, bst, gamma, beta)
L.bn_silu_bwd(u, w, bst, gamma, beta, dg, db)
# attention
inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, device=dev).float() / 64))
fr = torch.outer(torch.arange(Tp, device=dev).float(), inv)
cs = torch.stack([fr.cos(), fr.sin()], -1).contiguous()
xq = rb(M, d)
wqkv, bqkv = rb(d + 128, d) * 0.1, torch.randn(d + 128, device=dev) * 0.1
qkv = torch.empty(M, d + 128, dtype=torch.bfloat16, device=dev)
L.gemm(M, d + 128, d, xq, d, wqkv, d, L.EPI_ROPE, qkv, d + 128, bias=bqkv, aux=cs, n_half=Tp, remap_p0=d + 64)
klen = torch.full((B,), Tp - 1, dtype=torch.int64, device=dev)
ctx, lse2 = L.mqa_fwd(qkv, B, Tp, H, d, klen, 0.1, 3)
dctx = rb(M, d) * 0.1
L.mqa_bwd(qkv, ctx, dctx, lse2, B, Tp, H, d, klen, cs, 0.1, 3)
# FFN GEMMs of one block (forward, dgrad, wgrad with fused bias gradient) + an output projection
W1, bb1 = rb(2 * dff, d) * 0.1, torch.randn(2 * dff, device=dev)
gv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=dev)
h = torch.empty(M, dff, dtype=torch.bfloat16, device=dev)
W2, bb2 = rb(d, dff) * 0.1, torch.randn(d, device=dev)
res = torch.randn(M, d, device=dev)
out = torch.empty(M, d, device=dev)
dgv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=dev)
dW1, dB1 = torch.zeros(2 * dff, d, device=dev), torch.zeros(2 * dff, device=dev)
dW2, dB2 = torch.zeros(d, dff, device=dev), torch.zeros(d, device=dev)
dxn = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
dy = rb(M, d) * 0.1
L.gemm(M, dff, d, xq, d, W1, d, L.EPI_SWIGLU, h, dff, out2=gv, ldo2=2 * dff, bias=bb1, n_half=dff, drop_p=0.1, seed=1)
L.gemm(M, d, dff, h, dff, W2, dff, L.EPI_RESID, out, d, bias=bb2, aux=res, ldaux=d, alpha=0.5, drop_p=0.1, seed=2)
L.gemm(M, dff, d, dy, d, W2, dff, L.EPI_SWIGLU_BWD, dgv, 2 * dff, b_mn=1, aux=gv, ldaux=2 * dff, n_half=dff, drop_p=0.1, seed=1)
L.gemm(2 * dff, d, M, dgv, 2 * dff, xq, d, L.EPI_ATOMIC, dW1, d, a_mn=1, b_mn=1, split_k=9, colsum=dB1)
L.gemm(d, dff, M, dy, d, h, dff, L.EPI_ATOMIC, dW2, dff, a_mn=1, b_mn=1, split_k=37, colsum=dB2)
L.gemm(M, d, 2 * dff, dgv, 2 * dff, W1, d, L.EPI_STORE, dxn, d, b_mn=1)
L.gemm(M, d, d, dy, d, wqkv[:d], d, L.EPI_STORE, dxn, d, b_mn=1)
# CTC
Vp = (V + 7) // 8 * 8
logits = rb(B, Tp, Vp)[:, :, :V]
tl = torch.full((B,), 53, dtype=torch.int64, device=dev)
tg = torch.randint(1, V, (B, 53), generator=g, device=dev)
L.ctc_loss_fwd_bwd(logits, tg, klen, tl)
torch.cuda.synchronize()
print("done")
