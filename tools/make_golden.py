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

# ---------------------------------------------------------------- full-depth models (BASELINE configs[0]/[1] and [2])
# Inputs are regenerated from the seed by the tests (torch CPU generator); only checksums + outputs are stored, and the
# logits are stored on a sub-grid to keep the fixtures small.
def big_case(name, d, H, nb, V, B, T, lens, S, t_stride, v_stride):
    torch.manual_seed(0)
    model = TurkishASRModel(80, d, H, nb, V, dropout=0.0).train()
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(B, T, 80, generator=g)
    il = torch.tensor(lens)
    for b in range(B):
        x[b, il[b]:] = 0.0
    tl = torch.tensor([min(S, int(l) // 8) for l in lens])
    targets = torch.randint(1, V, (B, S), generator=g)
    logits = model(x, il)
    lp = torch.nn.functional.log_softmax(logits.permute(1, 0, 2), dim=2)
    loss = torch.nn.CTCLoss(blank=0, zero_infinity=True)(lp, targets, il // 4, tl)
    loss.backward()
    gn = {n: float(p.grad.norm()) if p.grad is not None else -1.0 for n, p in model.named_parameters()}
    np.savez_compressed(
        os.path.join(OUT, name), cfg=np.array([d, H, nb, V, B, T, S, t_stride, v_stride]), input_lengths=il.numpy(),
        target_lengths=tl.numpy(), x_checksum=np.array([float(x.double().sum()), float(x.double().abs().sum())]),
        targets=targets.numpy(), logits_sub=logits.detach()[:, ::t_stride, ::v_stride].numpy().astype(np.float32),
        logits_absmax=float(logits.detach().abs().max()), loss=float(loss), grad_names=np.array(list(gn.keys())),
        grad_norms=np.array(list(gn.values())))
    print(name, "loss", float(loss))


big_case("c1_golden.npz", 256, 4, 8, 1000, 8, 1001, [1001, 1001, 977, 900, 1001, 811, 640, 1001], 40, 5, 4)
big_case("cm_golden.npz", 512, 8, 16, 1000, 2, 301, [301, 222], 20, 3, 4)

# ---------------------------------------------------------------- Trainer.train_epoch of the reference (CPU)
from trainer.trainer import Trainer as RefTrainer  # noqa: E402


class _Log:
    def info(self, *a, **k):
        pass

    warning = error = info


class _Cfg:
    log_interval = 10 ** 9
    epochs = 1
    checkpoint_dir = "/tmp/_tasr_golden_ckpt"
    resume = False


def trainer_case(accum):
    torch.manual_seed(0)
    model = TurkishASRModel(CFG["n_mels"], CFG["d_model"], CFG["n_heads"], CFG["n_blocks"], CFG["n_classes"], dropout=0.0)
    init = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=5e-4, total_steps=100, pct_start=0.1, anneal_strategy="cos")
    g = torch.Generator().manual_seed(77)
    batches = []
    for i in range(3):
        T = 67 + 8 * i
        xb = torch.randn(2, T, 80, generator=g)
        ilb = torch.tensor([T, T - 20])
        xb[1, T - 20:] = 0.0
        batches.append((xb, torch.randint(1, CFG["n_classes"], (2, 5), generator=g), ilb, torch.tensor([5, 3])))
    tr = RefTrainer(model, batches, opt, sched, torch.device("cpu"), _Cfg(), _Log(), gradient_clip=1.0,
                    accumulation_steps=accum)
    avg = tr.train_epoch(1)
    final = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    step = opt.state[next(iter(model.parameters()))]["step"]
    return {"avg_loss_accum%d" % accum: float(avg), "global_step_accum%d" % accum: tr.global_step,
            "sched_last_epoch_accum%d" % accum: sched.last_epoch, "opt_step_accum%d" % accum: float(step),
            "delta_sub_accum%d" % accum: (final - init)[::16].numpy(), "lr_accum%d" % accum: opt.param_groups[0]["lr"]}


out = {}
out.update(trainer_case(1))
out.update(trainer_case(2))
np.savez_compressed(os.path.join(OUT, "trainer_golden.npz"), **out)
print({k: (v if np.ndim(v) == 0 else np.shape(v)) for k, v in out.items()})
