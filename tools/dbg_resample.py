import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import augment as oa
from turkish_asr_model_b200.data.preprocessing import SpeedPerturbation
g = torch.Generator().manual_seed(1)
lengths = [3000, 2500, 1800]
waves = torch.zeros(3, 3000)
for i, n in enumerate(lengths):
    waves[i, :n] = torch.randn(n, generator=g)
speeds = [0.9, 1.1, 1.0]
y, new_len = SpeedPerturbation().apply_batch(waves.cuda(), torch.tensor(lengths), speeds=speeds)
for i, (n, s) in enumerate(zip(lengths, speeds)):
    o, m = (1, 1) if s == 1.0 else oa.speed_to_freqs(s)
    ref = oa.resample_sinc(waves[i, :n].numpy(), o, m)
    d = np.abs(y[i, :ref.shape[0]].cpu().numpy() - ref)
    bad = np.where(d > 1e-5)[0]
    print(i, s, o, m, d.max(), len(bad), bad[:20], [(int(j) // m, int(j) % m) for j in bad[:8]])
