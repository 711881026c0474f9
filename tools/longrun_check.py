import sys, os, torch
sys.path.insert(0, os.getcwd())
import bench
from turkish_asr_model_b200.model import TurkishASRModel
from turkish_asr_model_b200.trainer import Trainer
dev = torch.device("cuda:0"); torch.manual_seed(0)
C = bench.CFG
model = TurkishASRModel(C["n_mels"], C["d_model"], C["n_heads"], C["n_blocks"], C["vocab"], dropout=0.1).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)
sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=5e-4, total_steps=400, pct_start=0.1, anneal_strategy="cos")
class Cfg: log_interval = 10**9
tr = Trainer(model, None, opt, sched, dev, Cfg(), None)
bs = bench.make_batches(4, 0, 1)
db = [(bench.synth_waves(b, dev), b["n_samples"].to(dev), b["targets"].to(dev), b["target_lengths"].to(dev), 1 + int(b["n_samples"].max()) // 160) for b in bs]
losses = []
for i in range(300):
    w, n, t, tl, tmax = db[i % 4]
    l = tr.train_step_waveforms(w, n, t, tl, tmax=tmax)
    if i % 25 == 0 or i == 299:
        losses.append(round(float(l), 3)); 
print("losses every 25 steps:", losses, "grad norm", float(tr.last_grad_norm))
assert all(map(lambda v: v == v, losses)) and losses[-1] < 0.6 * losses[0]
print("LONGRUN_OK")
