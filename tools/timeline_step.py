"""Timeline of one CUDA-graph replay of the training step (torch.profiler / CUPTI chrome trace): per-stream busy time,
idle gaps between consecutive kernels of the main chain, and which kernels the largest gaps follow.
    python tools/timeline_step.py [batch_index] [--eager]"""
import collections
import json
import os
import re
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from turkish_asr_model_b200.model import TurkishASRModel  # noqa: E402
from turkish_asr_model_b200.trainer import Trainer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
C = bench.CFG
if "--conformer-m" in sys.argv:
    C.update(d_model=512, n_heads=8, n_blocks=16)
model = TurkishASRModel(C["n_mels"], C["d_model"], C["n_heads"], C["n_blocks"], C["vocab"], dropout=C["dropout"]).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)


class Cfg:
    log_interval = 10 ** 9


tr = Trainer(model, None, opt, None, dev, Cfg(), None, use_cuda_graphs="--eager" not in sys.argv)
nums = [a for a in sys.argv[1:] if a.isdigit()]
which = int(nums[0]) if nums else 0
b = bench.make_batches(which + 1, 0, 1)[which]
w = bench.synth_waves(b, dev)
args = (w, b["n_samples"].to(dev), b["targets"].to(dev), b["target_lengths"].to(dev))
tmax = 1 + int(b["n_samples"].max()) // 160
for _ in range(4):
    tr.train_step_waveforms(*args, tmax=tmax)
torch.cuda.synchronize()
NSTEP = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NSTEP):
        tr.train_step_waveforms(*args, tmax=tmax)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "tasr_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])


def short(n):
    n = re.sub(r'\(CUtensor.*', '', n)
    n = re.sub(r'\((const|int|float|long|__nv|void|unsigned).*', '', n)
    return re.sub(r'^void ', '', n).replace("(anonymous namespace)::", "")[:70]


streams = collections.defaultdict(list)
for e in ev:
    streams[e["args"].get("stream", -1)].append(e)
span = (ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]) / NSTEP
print("Tmax %d: span %.1f us/step, %d device activities/step" % (tmax, span, len(ev) // NSTEP))
main = max(streams, key=lambda s: sum(e["dur"] for e in streams[s]))
for s, lst in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    print("  stream %s: %d activities/step, busy %.1f us/step%s" % (s, len(lst) // NSTEP, sum(e["dur"] for e in lst) / NSTEP,
                                                                   "  <- main chain" if s == main else ""))
for sid in [s for s in streams if len(streams[s]) // NSTEP >= 20]:
    lst = streams[sid]
    gaps = collections.defaultdict(lambda: [0, 0.0])
    tot_gap, n_gap, hist = 0.0, 0, collections.Counter()
    for a, b2 in zip(lst[:-1], lst[1:]):
        g = b2["ts"] - (a["ts"] + a["dur"])
        if g > 2000:  # boundary between two steps (host side)
            continue
        g = max(g, 0.0)
        tot_gap += g
        n_gap += 1
        hist[min(int(g), 10)] += 1
        key = short(a["name"]) + "  ->  " + short(b2["name"])
        gaps[key][0] += 1
        gaps[key][1] += g
    print("stream %s: idle between consecutive kernels %.1f us/step over %d boundaries/step (mean %.2f us)" % (
        sid, tot_gap / NSTEP, n_gap // NSTEP, tot_gap / max(n_gap, 1)))
    print("  gap histogram (us -> count/step):", {k: v // NSTEP for k, v in sorted(hist.items())})
    for k, (n, t) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:14]:
        print("  %9.1f us/step  n=%3d  mean %6.2f  %s" % (t / NSTEP, n // NSTEP, t / n, k))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[short(e["name"])][0] += 1
    agg[short(e["name"])][1] += e["dur"]
tot = sum(v[1] for v in agg.values())
print("kernel-time sum %.1f us/step" % (tot / NSTEP))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:48]:
    print(f'{t / NSTEP:10.1f} us {100 * t / tot:5.1f}%  n={n // NSTEP:4d}  {k}')
