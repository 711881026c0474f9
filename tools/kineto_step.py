"""Per-kernel device time of one eager training step under torch.profiler (CUPTI): warm L2, real overlap, unlike the
cold/serialised ncu launch list.  Prints kernels aggregated by name, and by (name, grid) with --grid."""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from turkish_asr_model_b200.model import TurkishASRModel  # noqa: E402
from turkish_asr_model_b200.trainer import Trainer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
C = bench.CFG
model = TurkishASRModel(C["n_mels"], C["d_model"], C["n_heads"], C["n_blocks"], C["vocab"], dropout=C["dropout"]).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)


class Cfg:
    log_interval = 10 ** 9


tr = Trainer(model, None, opt, None, dev, Cfg(), None, use_cuda_graphs=False)
which = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 0
b = bench.make_batches(which + 1, 0, 1)[which]
w = bench.synth_waves(b, dev)
args = (w, b["n_samples"].to(dev), b["targets"].to(dev), b["target_lengths"].to(dev))
tmax = 1 + int(b["n_samples"].max()) // 160
for _ in range(3):
    tr.train_step_waveforms(*args, tmax=tmax)
torch.cuda.synchronize()
NSTEP = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NSTEP):
        tr.train_step_waveforms(*args, tmax=tmax)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
first, last = None, None
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = ev.name
    short = re.sub(r'\(CUtensor.*', '', name)
    short = re.sub(r'\((const|int|float|long|__nv|void|unsigned).*', '', short)
    short = re.sub(r'^void ', '', short)
    dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    agg[short][0] += 1
    agg[short][1] += dur
    tot += dur
    s = ev.time_range.start
    e = ev.time_range.end
    first = s if first is None else min(first, s)
    last = e if last is None else max(last, e)
print("Tmax %d  kernel-time sum %.1f us/step, span %.1f us/step" % (tmax, tot / NSTEP, (last - first) / NSTEP))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{t / NSTEP:10.1f} us {100 * t / tot:5.1f}%  n={n // NSTEP:4d}  {k[:120]}')
