"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE (/root/reference, CPU, fp32).

The reference ships no tests or golden vectors, so parity is pinned by executing it:
    python tools/make_golden.py
Only numeric inputs/outputs are stored (no reference source).  Fixtures:
  mel_golden.npz      waveforms -> AudioPreprocessor.extract_features (data/preprocessing.py:81-116)
  model_golden.npz    TurkishASRModel(80, 128, 2, 1, 32, dropout=0) seeded init: forward logits, CTC loss
                      (trainer/trainer.py:167-173), gradient norms, BatchNorm running stats, greedy ids
  sampler_golden.npz  BucketingSampler flat orders (data/dataset.py:149-167) for seeded global `random`
"""
import os
import random
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.modules.setdefault("jiwer", types.ModuleType("jiwer"))
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)

from data.preprocessing import AudioPreprocessor  # noqa: E402
from data.dataset import BucketingSampler  # noqa: E402
from model.conformer import TurkishASRModel  # noqa: E402

torch.set_num_threads(4)

# ---------------------------------------------------------------- mel
g = torch.Generator().manual_seed(1234)
lengths = [16000, 12345, 8000, 201]
waves = [0.1 * torch.randn(n, generator=g) for n in lengths]
waves[1][:4000] = 0.0  # leading silence: exercises the top_db clamp and the 1e-10 floor
pre = AudioPreprocessor()
feats = [pre.extract_features(w).numpy() for w in waves]
np.savez_compressed(os.path.join(OUT, "mel_golden.npz"), lengths=np.array(lengths),
                    **{"wave%d" % i: w.numpy() for i, w in enumerate(waves)},
                    **{"feat%d" % i: f for i, f in enumerate(feats)})

# ---------------------------------------------------------------- model
CFG = dict(n_mels=80, d_model=128, n_heads=2, n_blocks=1, n_classes=32)
torch.manual_seed(0)
model = TurkishASRModel(CFG["n_mels"], CFG["d_model"], CFG["n_heads"], CFG["n_blocks"], CFG["n_classes"], dropout=0.0).train()
g = torch.Generator().manual_seed(7)
B, T = 2, 67
x = torch.randn(B, T, 80, generator=g)
il = torch.tensor([67, 41])
x[1, 41:] = 0.0
targets = torch.randint(1, CFG["n_classes"], (B, 5), generator=g)
tl = torch.tensor([5, 3])
checksums = {k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in model.state_dict().items()}
logits = model(x, il)
log_probs = torch.nn.functional.log_softmax(logits.permute(1, 0, 2), dim=2)
loss = torch.nn.CTCLoss(blank=0, zero_infinity=True)(log_probs, targets, il // 4, tl)
loss.backward()
grad_norms = {n: float(p.grad.norm()) if p.grad is not None else -1.0 for n, p in model.named_parameters()}
bn = {k: v.numpy() for k, v in model.state_dict().items() if "running_" in k}
model.eval()
with torch.no_grad():
    logits_eval = model(x, il)
ids = torch.argmax(logits_eval, dim=-1)
np.savez_compressed(
    os.path.join(OUT, "model_golden.npz"), x=x.numpy(), input_lengths=il.numpy(), targets=targets.numpy(),
    target_lengths=tl.numpy(), logits_train=logits.detach().numpy(), logits_eval=logits_eval.numpy(), loss=float(loss),
    greedy_ids=ids.numpy(), grad_names=np.array(list(grad_norms.keys())), grad_norms=np.array(list(grad_norms.values())),
    checksum_names=np.array(list(checksums.keys())), checksums=np.stack(list(checksums.values())),
    bn_names=np.array(list(bn.keys())), **{"bn%d" % i: v for i, v in enumerate(bn.values())})

# ---------------------------------------------------------------- sampler
class _DS:
    def __init__(self, n):
        self.file_pairs = [("/nonexistent/%d.wav" % i, "") for i in range(n)]

    def __len__(self):
        return len(self.file_pairs)


rng = np.random.RandomState(5)
sizes = (44 + 2 * rng.randint(80000, 240000, size=203)).tolist()
orders = {}
for bs, seed, drop in [(8, 42, False), (16, 7, True), (64, 1234, False)]:
    s = BucketingSampler(_DS(len(sizes)), bs, shuffle=True, drop_last=drop)
    s.lengths = list(sizes)  # os.path.getsize proxy (the files do not exist)
    random.seed(seed)
    orders["order_bs%d_seed%d_drop%d" % (bs, seed, int(drop))] = np.array(list(iter(s)))
np.savez_compressed(os.path.join(OUT, "sampler_golden.npz"), sizes=np.array(sizes), **orders)
print("golden fixtures written to", OUT, [f for f in os.listdir(OUT)])
