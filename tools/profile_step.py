"""One training step of the bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

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
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
b = bench.make_batches(which + 1, 0, 1)[which]
w = bench.synth_waves(b, dev)
args = (w, b["n_samples"].to(dev), b["targets"].to(dev), b["target_lengths"].to(dev))
tmax = 1 + int(b["n_samples"].max()) // 160
for _ in range(2):
    tr.train_step_waveforms(*args, tmax=tmax)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = tr.train_step_waveforms(*args, tmax=tmax)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("Tmax", tmax, "loss", float(loss))
