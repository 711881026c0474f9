"""Warm, back-to-back CUDA-event timing of the non-GEMM kernels at the bench's shapes (B=64, T=1325 frames, T'=332, d=256).
usage: python tools/bench_kernels.py [name-substring ...]     (no args = all)"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turkish_asr_model_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, T, F, d, H, G = 64, 1325, 80, 256, 4, 8
T1, F1 = (T - 1) // 2 + 1, (F - 1) // 2 + 1
Tp = (T1 - 1) // 2 + 1
M = B * Tp
sel = sys.argv[1:]
# BK_FLUSH=1: write a 512 MB buffer between launches, so every launch sees cold operands (HBM, not L2) like inside a step
_flush = torch.empty(128 << 20, dtype=torch.float32, device=dev) if os.environ.get("BK_FLUSH", "0") == "1" else None


def timeit(name, fn, bytes_=0, iters=int(os.environ.get("BK_ITERS", "10"))):
    if sel and not any(s in name for s in sel):
        return
    for _ in range(int(os.environ.get("BK_WARMUP", "3"))):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:   # device time only (these calls are host-bound from python)
        for _ in range(iters):
            if _flush is not None:
                _flush.fill_(0.0)
            fn()
        torch.cuda.synchronize()
    evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and "at::native" not in ev.name]
    us = sum(ev.device_time for ev in evs) / iters
    parts = {}
    for ev in evs:
        k = ev.name.split("::")[-1].split("(")[0][:28]
        parts[k] = parts.get(k, 0.0) + ev.device_time / iters
    if len(parts) > 1:
        name = name + " [" + ", ".join("%s %.1f" % kv for kv in parts.items()) + "]"
    if us <= 0:  # no CUPTI (e.g. running under ncu)
        print("%-28s (no device timing available)" % name)
        return
    print("%-28s %9.1f us   %7.2f TB/s (algorithmic %.1f MB)" % (name, us, bytes_ / us / 1e6 if bytes_ else 0, bytes_ / 1e6))


from turkish_asr_model_b200.data.preprocessing import AudioPreprocessor  # noqa: E402
_pre = AudioPreprocessor()
_w = torch.randn(B, (T - 1) * 160, device=dev) * 0.1
_n = torch.full((B,), (T - 1) * 160, dtype=torch.int64, device=dev)
timeit("mel_forward", lambda: _pre.extract_features_batch(_w, _n, T), _w.numel() * 4 + B * T * F * 4 * 3)
feats = torch.randn(B, T, F, device=dev)
w1 = torch.randn(d, 1, 3, 3, device=dev) * 0.3
b1 = torch.randn(d, device=dev) * 0.1
y1 = L.conv1_fwd(feats, w1, b1)
dy1 = torch.randn_like(y1)
dw1 = torch.zeros_like(w1)
db1 = torch.zeros_like(b1)
timeit("conv1_fwd", lambda: L.conv1_fwd(feats, w1, b1), y1.numel() * 2 + feats.numel() * 4)
timeit("conv1_bwd", lambda: L.conv1_bwd(dy1, feats, w1, b1, dw1, db1), y1.numel() * 2 + feats.numel() * 4)
del y1, dy1

x = torch.randn(B, Tp, d, device=dev)
gamma = torch.randn(d, device=dev)
beta = torch.randn(d, device=dev)
y, stats = L.groupnorm_fwd(x, G, gamma, beta)
timeit("groupnorm_fwd", lambda: L.groupnorm_fwd(x, G, gamma, beta), M * d * 6)
dy = torch.randn(B, Tp, d, device=dev).bfloat16()
dres = torch.randn(B, Tp, d, device=dev)
dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
timeit("groupnorm_bwd(+cast)", lambda: L.groupnorm_bwd(dy, x, G, stats, gamma, dres, True, dg, db, cast=(0.5, 0.1, 7)), M * d * 16)

u = torch.randn(B, Tp, d, device=dev).bfloat16()
ab = torch.randn(B, Tp, 2 * d, device=dev).bfloat16()
dwt = torch.randn(d, 31, device=dev) * 0.2
dwb = torch.randn(d, device=dev) * 0.1
timeit("dwconv_fwd", lambda: L.dwconv_fwd(u, dwt, dwb), M * d * 4)
gdw, gdb = torch.zeros_like(dwt), torch.zeros_like(dwb)
timeit("dwconv_bwd(data+weight)", lambda: L.dwconv_bwd(u, u, ab, dwt, gdw, gdb), M * d * 12)
w, part = L.dwconv_fwd(u, dwt, dwb)
rm, rv, nb = torch.zeros(d, device=dev), torch.ones(d, device=dev), torch.zeros((), dtype=torch.long, device=dev)
bst = L.bn_finalize(part, d, M, 1e-5, 0.1, True, rm, rv, nb)
timeit("bn_finalize", lambda: L.bn_finalize(part, d, M, 1e-5, 0.1, True, rm, rv, nb))
timeit("bn_silu_fwd", lambda: L.bn_silu_fwd(w, bst, gamma, beta), M * d * 4)
timeit("bn_silu_bwd", lambda: L.bn_silu_bwd(u, w, bst, gamma, beta, dg, db), M * d * 8)
timeit("colsum d", lambda: L.colsum_add(u.view(M, d), dg), M * d * 2)
timeit("colsum 2d", lambda: L.colsum_add(ab.view(M, 2 * d), torch.zeros(2 * d, device=dev)), M * d * 4)

dqkv = d + 2 * (d // H)
qkv = (torch.randn(M, dqkv, device=dev) * 0.5).bfloat16()
klen = torch.full((B,), Tp, dtype=torch.int64, device=dev)
ctx, lse2 = L.mqa_fwd(qkv, B, Tp, H, d, klen, 0.1, 3)
fl = 4.0 * B * H * Tp * Tp * (d // H)
timeit("mqa_fwd", lambda: L.mqa_fwd(qkv, B, Tp, H, d, klen, 0.1, 3))
half = (d // H) // 2
ang = torch.arange(Tp, device=dev)[:, None] * (10000.0 ** (-torch.arange(half, device=dev) / half))[None]
cos_sin = torch.stack([ang.cos(), ang.sin()], -1).float().contiguous()
dctx = torch.randn_like(ctx)
try:
    timeit("mqa_bwd", lambda: L.mqa_bwd(qkv, ctx, dctx, lse2, B, Tp, H, d, klen, cos_sin, 0.1, 3))
except Exception as e:  # cos/sin layout differs from the engine's: report and go on
    print("mqa_bwd skipped:", e)
print("attention fwd flops %.1f GF" % (fl / 1e9))

V, S = 1000, 60
logits = torch.randn(B, Tp, V, device=dev).bfloat16()
tg = torch.randint(1, V, (B, S), device=dev)
il = torch.full((B,), Tp, dtype=torch.int64, device=dev)
tl = torch.full((B,), S, dtype=torch.int64, device=dev)
timeit("ctc_loss_fwd_bwd", lambda: L.ctc_loss_fwd_bwd(logits, tg, il, tl), B * Tp * V * 4)
