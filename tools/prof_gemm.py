"""Stand-alone timing of the GEMM shapes of one Conformer block (for ncu and quick A/B)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turkish_asr_model_b200 import _lib as L
from turkish_asr_model_b200.engine import _split_k

dev = torch.device("cuda:0")
M, d, dff = 21248, int(os.environ.get("PG_D", "256")), int(os.environ.get("PG_DFF", "1024"))
g = torch.Generator(device=dev).manual_seed(0)
def rb(*s): return (torch.randn(*s, generator=g, device=dev) * 0.5).to(torch.bfloat16)
x = rb(M, d); W1 = rb(2 * dff, d); b1 = torch.randn(2 * dff, device=dev)
gv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=dev); h = torch.empty(M, dff, dtype=torch.bfloat16, device=dev)
W2 = rb(d, dff); b2 = torch.randn(d, device=dev); res = torch.randn(M, d, device=dev); out = torch.empty(M, d, device=dev)
dy = rb(M, d); dgv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=dev)
dW1 = torch.zeros(2 * dff, d, device=dev); dxn = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
drop = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1

def ff1(): L.gemm(M, dff, d, x, d, W1, d, L.EPI_SWIGLU, h, dff, out2=gv, ldo2=2 * dff, bias=b1, n_half=dff, drop_p=drop, seed=1)
def ff2(): L.gemm(M, d, dff, h, dff, W2, dff, L.EPI_RESID, out, d, bias=b2, aux=res, ldaux=d, alpha=0.5, drop_p=drop, seed=2)
def ff2_dgrad(): L.gemm(M, dff, d, dy, d, W2, dff, L.EPI_SWIGLU_BWD, dgv, 2 * dff, b_mn=1, aux=gv, ldaux=2 * dff, n_half=dff, drop_p=drop, seed=1)
def ff1_wgrad(): L.gemm(2 * dff, d, M, dgv, 2 * dff, x, d, L.EPI_ATOMIC, dW1, d, a_mn=1, b_mn=1, split_k=_split_k(2 * dff, d, M))
def ff1_dgrad(): L.gemm(M, d, 2 * dff, dgv, 2 * dff, W1, d, L.EPI_STORE, dxn, d, b_mn=1)
def plain(): L.gemm(M, 2 * dff, d, x, d, W1, d, L.EPI_STORE, gv, 2 * dff, bias=b1)
fns = {"ff1": ff1, "ff2": ff2, "ff2_dgrad": ff2_dgrad, "ff1_wgrad": ff1_wgrad, "ff1_dgrad": ff1_dgrad, "plain": plain}
flops = {"ff1": 2 * M * 2 * dff * d, "ff2": 2 * M * d * dff, "ff2_dgrad": 2 * M * dff * d, "ff1_wgrad": 2 * M * 2 * dff * d,
         "ff1_dgrad": 2 * M * 2 * dff * d, "plain": 2 * M * 2 * dff * d}
if which == "once":  # one warm-up + one measured launch per shape (for ncu --set full)
    for name, fn in fns.items():
        fn()
    torch.cuda.synchronize()
    for name, fn in fns.items():
        fn()
    torch.cuda.synchronize()
    print("once done")
    sys.exit(0)
for name, fn in fns.items():
    if which != "all" and which != name:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n = 10
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:  # device time (python launch overhead excluded)
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    us = sum(ev.device_time for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and "gemm" in ev.name) / n
    print("%-10s %8.1f us  %7.1f TFLOP/s" % (name, us, flops[name] / us / 1e6))
