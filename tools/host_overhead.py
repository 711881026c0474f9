"""Host-issue time vs GPU time of one training step (is the step launch-bound?)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from turkish_asr_model_b200.model import TurkishASRModel
from turkish_asr_model_b200.trainer import Trainer
from turkish_asr_model_b200 import _lib as L

dev = torch.device("cuda:0")
torch.manual_seed(0)
C = bench.CFG
model = TurkishASRModel(C["n_mels"], C["d_model"], C["n_heads"], C["n_blocks"], C["vocab"], dropout=C["dropout"]).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)
class Cfg: log_interval = 10 ** 9
tr = Trainer(model, None, opt, None, dev, Cfg(), None)
batches = bench.make_batches(6, 0, 1)
db = []
for b in batches:
    db.append((bench.synth_waves(b, dev), b["n_samples"].to(dev), b["targets"].to(dev), b["target_lengths"].to(dev), 1 + int(b["n_samples"].max()) // 160))
for i in range(12):
    tr.train_step_waveforms(*db[i % 6][:4], tmax=db[i % 6][4])
torch.cuda.synchronize()
N = 24
t0 = time.perf_counter()
for i in range(N):
    b = db[i % 6]
    tr.train_step_waveforms(*b[:4], tmax=b[4])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host issue ms/step %.2f   total ms/step %.2f   tmax %s" % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3, [b[4] for b in db]))
