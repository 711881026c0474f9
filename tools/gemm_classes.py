"""Every tcgen05 GEMM launch of one training step, grouped by (epilogue, M, N, K, operand major-ness): launches per step,
stand-alone device time per launch (CUDA events, L2 flushed between launches), TFLOP/s and the class's share.
    python tools/gemm_classes.py [--conformer-m]"""
import collections
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from turkish_asr_model_b200 import _lib as L  # noqa: E402
from turkish_asr_model_b200.model import TurkishASRModel  # noqa: E402
from turkish_asr_model_b200.trainer import Trainer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
Cc = bench.CFG
if "--conformer-m" in sys.argv:
    Cc.update(d_model=512, n_heads=8, n_blocks=16)
model = TurkishASRModel(Cc["n_mels"], Cc["d_model"], Cc["n_heads"], Cc["n_blocks"], Cc["vocab"], dropout=Cc["dropout"]).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)


class Cfg:
    log_interval = 10 ** 9


tr = Trainer(model, None, opt, None, dev, Cfg(), None, use_cuda_graphs=False)
b = bench.make_batches(1, 0, 1)[0]
w = bench.synth_waves(b, dev)
args = (w, b["n_samples"].to(dev), b["targets"].to(dev), b["target_lengths"].to(dev))
tmax = 1 + int(b["n_samples"].max()) // 160
tr.train_step_waveforms(*args, tmax=tmax)
L.GEMM_PROFILE = []
tr.train_step_waveforms(*args, tmax=tmax)
torch.cuda.synchronize()
prof, L.GEMM_PROFILE = L.GEMM_PROFILE, None
names = ["STORE", "RESID", "SWIGLU", "GLU", "SILU", "SWIGLU_BWD", "GLU_BWD", "SILU_BWD", "ATOMIC", "ROPE"]
classes = collections.OrderedDict()
for flops, a, keep, shape in prof:
    classes.setdefault(shape, []).append((flops, a, keep))
flush = torch.empty(128 << 20, dtype=torch.float32, device=dev)
fn = L.lib().tasr_gemm_bf16
rows = []
for shape, items in classes.items():
    flops, a, keep = items[0]
    for _ in range(2):
        L.check(fn(C.addressof(a), L.stream_ptr()))
    tot = 0.0
    reps = 6
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(fn(C.addressof(a), L.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / reps * 1e3
    rows.append((us * len(items), shape, len(items), us, flops / us / 1e6))
total = sum(r[0] for r in rows)
print("%-11s %6s %6s %6s %4s %3s %9s %9s %9s %6s" % ("epilogue", "M", "N", "K", "maj", "n", "us/launch", "TFLOP/s", "us/step", "share"))
for tot_us, (M, N, K, epi, am, bm), n, us, tf in sorted(rows, reverse=True):
    print("%-11s %6d %6d %6d  %d%d %3d %9.1f %9.1f %9.1f %5.1f%%" % (names[epi], M, N, K, am, bm, n, us, tf, tot_us, 100 * tot_us / total))
print("total %.1f us/step (cold operands, stand-alone launches)" % total)
