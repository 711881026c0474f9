"""GPU: the Trainer step (reference trainer/trainer.py:147-225) on the B200 kernels against
  * the reference's own Trainer.train_epoch (tests/golden/trainer_golden.npz, written by tools/make_golden.py),
  * oracle/trainer.py (fp32 CPU restatement, itself pinned to that fixture by tests/test_oracle_golden.py),
and the trainer-level behaviours: CUDA-graph replay == eager, NaN skip, gradient accumulation + leftover flush,
checkpoint save / resume in the reference's format, stale-operand protection on the autograd path."""
import os

import numpy as np
import pytest
import torch

from golden_inputs import GCFG, trainer_golden_batches
from oracle import conformer as oc
from oracle.trainer import TrainerOracle
from turkish_asr_model_b200 import _lib as L
from turkish_asr_model_b200.model import TurkishASRModel
from turkish_asr_model_b200.trainer import Trainer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class Cfg:
    log_interval = 10 ** 9
    epochs = 1
    save_interval = 1
    resume = False
    checkpoint_dir = None


def _sched(opt):
    return torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=5e-4, total_steps=100, pct_start=0.1, anneal_strategy="cos")


def _trainer(model, dev, loader=None, accum=1, graphs=False, cfg=None, **kw):
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)
    return Trainer(model, loader, opt, _sched(opt), dev, cfg if cfg is not None else Cfg(), None, gradient_clip=1.0,
                   accumulation_steps=accum, use_cuda_graphs=graphs, **kw)


def _flat_params(model):
    return torch.cat([p.detach().reshape(-1).float().cpu() for p in model.parameters()])


def _cos(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))


# ------------------------------------------------------------------ vs the reference's own Trainer
@pytest.mark.parametrize("accum", [1, 2])
def test_train_epoch_vs_reference_trainer_golden(cuda, accum):
    g = np.load(os.path.join(GOLD, "trainer_golden.npz"))
    torch.manual_seed(0)
    model = TurkishASRModel(GCFG["n_mels"], GCFG["d_model"], GCFG["n_heads"], GCFG["n_blocks"], GCFG["n_classes"], dropout=0.0)
    init = _flat_params(model)
    model = model.to(cuda)
    tr = _trainer(model, cuda, loader=trainer_golden_batches(), accum=accum)
    avg = tr.train_epoch(1)
    torch.cuda.synchronize()
    ref_avg = float(g["avg_loss_accum%d" % accum])
    assert abs(avg - ref_avg) < 2e-2 * abs(ref_avg), (avg, ref_avg)
    # integer bookkeeping is exact: optimizer steps, scheduler position, global_step (the leftover flush steps the
    # optimizer but neither the scheduler nor global_step, trainer/trainer.py:213-219)
    assert tr.global_step == int(g["global_step_accum%d" % accum])
    assert tr.scheduler.last_epoch == int(g["sched_last_epoch_accum%d" % accum])
    assert tr._opt_step == int(g["opt_step_accum%d" % accum])
    assert abs(tr.optimizer.param_groups[0]["lr"] - float(g["lr_accum%d" % accum])) < 1e-12
    delta = (_flat_params(model) - init)[::16].numpy()
    ref = g["delta_sub_accum%d" % accum]
    # AdamW's first steps move a weight by ~lr*sign(g): bf16 operand noise flips entries whose gradient is ~0
    assert _cos(delta, ref) > 0.93, _cos(delta, ref)
    big = np.abs(ref) > 0.5 * np.abs(ref).max()
    assert np.mean(np.sign(delta[big]) == np.sign(ref[big])) > 0.97
    assert abs(np.linalg.norm(delta) / np.linalg.norm(ref) - 1.0) < 5e-2


# ------------------------------------------------------------------ vs the oracle, default width
def _default_case(nb=2, V=64, B=3, T=203, seed=0):
    torch.manual_seed(seed)
    model = TurkishASRModel(80, 256, 4, nb, V, dropout=0.0)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    batches = []
    for i in range(3):
        Ti = T + 12 * i
        x = torch.randn(B, Ti, 80, generator=g)
        il = torch.tensor([Ti, Ti - 53, Ti // 2][:B])
        for b in range(B):
            x[b, il[b]:] = 0.0
        batches.append((x, torch.randint(1, V, (B, 9), generator=g), il, torch.tensor([9, 6, 4][:B])))
    return model, sd, batches


def test_three_steps_vs_oracle_adamw(cuda):
    model, sd, batches = _default_case()
    pnames = [n for n, _ in model.named_parameters()]
    init = _flat_params(model)
    orc = TrainerOracle(sd, pnames, 4, 2, scheduler_fn=_sched)
    model = model.to(cuda)
    tr = _trainer(model, cuda)
    for x, tg, il, tl in batches:
        _, ref_losses = orc.train_epoch([(x, tg, il, tl)])
        loss = float(tr.train_step(x, tg, il, tl))
        assert abs(loss - ref_losses[0]) < 2e-2 * abs(ref_losses[0]), (loss, ref_losses[0])
        # global gradient norm BEFORE clipping (clip_grad_norm_'s return value, trainer/trainer.py:189)
        assert abs(float(tr.last_grad_norm) - orc.last_grad_norm) < 3e-2 * orc.last_grad_norm
    final_ref = torch.cat([orc.sd[n].detach().reshape(-1) for n in pnames])
    d_gpu, d_ref = (_flat_params(model) - init).numpy(), (final_ref - init).numpy()
    assert _cos(d_gpu, d_ref) > 0.93
    assert abs(np.linalg.norm(d_gpu) / np.linalg.norm(d_ref) - 1.0) < 5e-2
    # BatchNorm running statistics followed the reference's momentum update for three steps
    for k, v in model.state_dict().items():
        if k.endswith(("running_mean", "running_var")):
            r = orc.sd[k]
            assert ((v.cpu() - r).abs().max() / r.abs().max()).item() < 2e-2, k
        if k.endswith("num_batches_tracked"):
            assert int(v) == 3


def _wave_case(cuda, V=64, seed=3):
    torch.manual_seed(seed)
    model = TurkishASRModel(80, 256, 4, 2, V, dropout=0.0)
    g = torch.Generator().manual_seed(seed + 1)
    steps = []
    for n0 in (16000, 20480, 16000):
        ns = torch.tensor([n0, n0 - 3655])
        w = torch.zeros(2, n0)
        for b in range(2):
            w[b, : ns[b]] = 0.1 * torch.randn(int(ns[b]), generator=g)
        steps.append((w, ns, torch.randint(1, V, (2, 6), generator=g), torch.tensor([6, 4])))
    return model, steps


def test_cuda_graph_replay_matches_eager(cuda):
    """train_step_waveforms through captured CUDA graphs (incl. a replay of an already captured shape) == the same
    steps launched eagerly: same losses, same parameter trajectory."""
    results = []
    for graphs in (False, True):
        model, steps = _wave_case(cuda)
        init = _flat_params(model)
        model = model.to(cuda)
        tr = _trainer(model, cuda, graphs=graphs)
        losses = []
        for w, ns, tg, tl in steps:
            losses.append(float(tr.train_step_waveforms(w.to(cuda), ns, tg, tl)))
        torch.cuda.synchronize()
        if graphs:
            assert len(tr._graphs) == 2 and tr.graph_kernel_launches > 0  # third step replayed the first capture
        results.append((losses, (_flat_params(model) - init).numpy(), float(tr.last_grad_norm)))
    (l0, d0, n0), (l1, d1, n1) = results
    for a, b in zip(l0, l1):
        assert abs(a - b) < 2e-3 * abs(a), (l0, l1)
    assert abs(n0 - n1) < 1e-2 * n0  # third step: the trajectories have been apart by rounding noise for two updates
    assert _cos(d0, d1) > 0.995  # fp32 atomics reorder the last bits of the gradients; sign-like early AdamW steps


def test_graph_cache_is_bounded_and_workspace_never_freed(cuda):
    model, steps = _wave_case(cuda)
    model = model.to(cuda)
    tr = _trainer(model, cuda, graphs=True, max_cached_graphs=2)
    g = torch.Generator().manual_seed(0)
    seen = []
    for n0 in (16000, 17600, 19200, 16000):
        w = 0.1 * torch.randn(2, n0, generator=g)
        loss = tr.train_step_waveforms(w.to(cuda), torch.tensor([n0, n0 - 1000]), steps[0][2], steps[0][3])
        seen.append(float(loss))
        assert len(tr._graphs) <= 2
    assert all(np.isfinite(seen))
    # growing the scratch buffer keeps the old one alive (captured graphs hold its address)
    ws0 = L.workspace(1, cuda)
    big = L.workspace(ws0.numel() + 1, cuda)
    assert big.numel() >= 2 * ws0.numel() and any(b is ws0 for b in L._ws_retired)
    assert np.isfinite(float(tr.train_step_waveforms(w.to(cuda), torch.tensor([16000, 15000]), steps[0][2], steps[0][3])))


# ------------------------------------------------------------------ NaN skip, accumulation
def test_nan_gradients_skip_the_update(cuda):
    """reference trainer/trainer.py:178-181 skips a batch whose loss is NaN; here the check is on the device: a
    non-finite global gradient norm leaves parameters, moments and bf16 operands untouched."""
    model, sd, batches = _default_case()
    model = model.to(cuda)
    tr = _trainer(model, cuda)
    x, tg, il, tl = batches[0]
    tr.train_step(x, tg, il, tl)
    eng, flat = tr._flat()
    torch.cuda.synchronize()
    p0, m0, v0, s0 = flat.params.clone(), flat.exp_avg.clone(), flat.exp_avg_sq.clone(), flat.shadow.clone()
    xb = x.clone()
    xb[0, 5, 7] = float("nan")
    loss = tr.train_step(xb, tg, il, tl)
    torch.cuda.synchronize()
    assert not np.isfinite(float(loss)) and not np.isfinite(float(tr.last_grad_norm))
    n = flat.live_numel
    assert torch.equal(flat.params[:n], p0[:n]) and torch.equal(flat.exp_avg[:n], m0[:n])
    assert torch.equal(flat.exp_avg_sq[:n], v0[:n]) and torch.equal(flat.shadow[:n], s0[:n])
    loss2 = tr.train_step(x, tg, il, tl)  # a clean batch afterwards trains normally
    torch.cuda.synchronize()
    assert np.isfinite(float(loss2)) and not torch.equal(flat.params[:n], p0[:n])


def test_gradient_accumulation_matches_oracle(cuda):
    """accumulation_steps = 2: gradients of two micro-batches, each scaled by 1/2 (trainer/trainer.py:176), are summed
    in the flat gradient buffer; one optimizer step follows the second micro-batch."""
    model, sd, batches = _default_case()
    pnames = [n for n, _ in model.named_parameters()]
    sdr = {k: (v.clone().requires_grad_(True) if k in pnames else v) for k, v in sd.items()}
    for x, tg, il, tl in batches[:2]:
        (oc.ctc_loss_torch(oc.forward(x, il, sdr, 4, 2, training=True), tg, il, tl) / 2).backward()
    model = model.to(cuda)
    tr = _trainer(model, cuda, accum=2)
    tr.train_step(*batches[0])
    assert tr._opt_step == 0 and tr._micro == 1
    tr.train_step(*batches[1])
    assert tr._opt_step == 1 and tr._micro == 0 and tr.global_step == 1
    torch.cuda.synchronize()
    eng, flat = tr._flat()
    views = flat.grad_views()
    rels = []
    for name in pnames:
        gref = sdr[name].grad
        if gref is None or name.endswith("depthwise_conv.bias"):
            continue
        rels.append(((views[name].cpu().double() - gref.double()).abs().max() / gref.abs().max()).item())
    assert max(rels) < 8e-2 and np.median(rels) < 2e-2, (max(rels), np.median(rels))


# ------------------------------------------------------------------ checkpoints
def test_checkpoint_resume_is_exact(cuda, tmp_path):
    """save -> fresh model + trainer -> load_checkpoint -> the next step gives the same loss and the same parameters
    as continuing the original run (weights, AdamW moments and step, scheduler, BatchNorm buffers all restored)."""
    model, sd, batches = _default_case()
    cfg = Cfg()
    cfg.checkpoint_dir = str(tmp_path)
    model = model.to(cuda)
    tr = _trainer(model, cuda, cfg=cfg)
    tr.train_step(*batches[0])
    tr.train_step(*batches[1])
    tr.save_checkpoint(1)
    pc = _flat_params(model)
    loss_a = float(tr.train_step(*batches[2]))
    pa = _flat_params(model)

    model2, _, _ = _default_case(seed=123)  # different init: everything must come from the file
    model2 = model2.to(cuda)
    cfg2 = Cfg()
    cfg2.checkpoint_dir, cfg2.resume = str(tmp_path), True
    tr2 = _trainer(model2, cuda, cfg=cfg2)
    tr2.load_checkpoint()
    assert tr2.start_epoch == 2 and tr2.global_step == 2 and tr2._opt_step == 2
    assert tr2.scheduler.last_epoch == 2
    loss_b = float(tr2.train_step(*batches[2]))
    pb = _flat_params(model2)
    assert abs(loss_a - loss_b) < 1e-4 * abs(loss_a), (loss_a, loss_b)
    # identical state => identical third step, up to the fp32 atomics order inside the gradient kernels (entries whose
    # gradient is pure rounding noise, e.g. the depthwise bias in front of BatchNorm, move by +-lr either way)
    da, db = (pa - pc).numpy(), (pb - pc).numpy()
    assert _cos(da, db) > 0.999
    assert np.mean(np.abs(da - db) < 1e-6) > 0.99


def test_checkpoint_interchange_with_reference_format(cuda, tmp_path):
    """(i) A checkpoint in the reference's layout (trainer/trainer.py:89-98: torch state_dicts) loads: weights, the
    per-parameter torch AdamW moments and step land in the flat buffers of the fused optimizer.
    (ii) A checkpoint written here loads into plain torch objects the way the reference does it
    (model.load_state_dict strict, optimizer.load_state_dict, scheduler.load_state_dict, GradScaler.load_state_dict)."""
    torch.manual_seed(5)
    ref_model = TurkishASRModel(80, 256, 4, 1, 40, dropout=0.0)  # same module tree / state_dict keys as the reference
    ref_opt = torch.optim.AdamW(ref_model.parameters(), lr=5e-4, weight_decay=1e-6)
    ref_sched = _sched(ref_opt)
    g = torch.Generator().manual_seed(6)
    for _ in range(3):  # synthetic gradients: this is about the container format, not the arithmetic
        for n, p in ref_model.named_parameters():
            p.grad = None if "norm_conv" in n else 1e-2 * torch.randn(p.shape, generator=g)
        ref_opt.step()
        ref_sched.step()
    state = {"epoch": 4, "global_step": 3, "model_state_dict": ref_model.state_dict(),
             "optimizer_state_dict": ref_opt.state_dict(), "scheduler_state_dict": ref_sched.state_dict(),
             "scaler_state_dict": torch.amp.GradScaler("cuda").state_dict(), "best_val_loss": 1.25,
             "config": {}}
    torch.save(state, os.path.join(tmp_path, "checkpoint_epoch_4.pt"))

    torch.manual_seed(99)
    model = TurkishASRModel(80, 256, 4, 1, 40, dropout=0.0).to(cuda)
    cfg = Cfg()
    cfg.checkpoint_dir, cfg.resume = str(tmp_path), True
    tr = _trainer(model, cuda, cfg=cfg)
    tr.load_checkpoint()
    assert tr.start_epoch == 5 and tr.global_step == 3 and tr._opt_step == 3 and tr.best_val_loss == 1.25
    eng, flat = tr._flat()
    ref_params = dict(ref_model.named_parameters())
    for name, p in model.named_parameters():
        assert torch.equal(p.detach().cpu(), ref_params[name].detach()), name
        st = ref_opt.state.get(ref_params[name])
        if st:
            assert torch.equal(flat.view(flat.exp_avg, name).cpu(), st["exp_avg"]), name
            assert torch.equal(flat.view(flat.exp_avg_sq, name).cpu(), st["exp_avg_sq"]), name
    assert abs(tr.optimizer.param_groups[0]["lr"] - ref_opt.param_groups[0]["lr"]) < 1e-12

    # (ii) write from here, read back the way the reference does
    x = torch.randn(2, 67, 80, generator=g)
    tr.train_step(x, torch.randint(1, 40, (2, 4), generator=g), torch.tensor([67, 50]), torch.tensor([4, 3]))
    tr.save_checkpoint(5)
    ck = torch.load(os.path.join(tmp_path, "checkpoint_epoch_5.pt"), map_location="cpu", weights_only=False)
    assert set(ck) >= {"epoch", "global_step", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict",
                       "scaler_state_dict", "best_val_loss", "config"}
    plain = TurkishASRModel(80, 256, 4, 1, 40, dropout=0.0)
    plain.load_state_dict(ck["model_state_dict"], strict=True)
    opt2 = torch.optim.AdamW(plain.parameters(), lr=5e-4, weight_decay=1e-6)
    opt2.load_state_dict(ck["optimizer_state_dict"])
    sch2 = _sched(opt2)
    sch2.load_state_dict(ck["scheduler_state_dict"])
    torch.amp.GradScaler("cuda").load_state_dict(ck["scaler_state_dict"])  # rejects an empty dict when enabled
    names = [n for n, _ in plain.named_parameters()]
    n_state = 0
    for idx, p in enumerate(plain.parameters()):
        st = opt2.state.get(p)
        if "norm_conv" in names[idx]:
            assert not st
            continue
        assert int(st["step"]) == 4
        assert torch.equal(st["exp_avg"], flat.view(flat.exp_avg, names[idx]).cpu()), names[idx]
        n_state += 1
    assert n_state == len(names) - 2
    opt2.step()  # the restored torch optimizer is usable (all grads None: no-op)


# ------------------------------------------------------------------ autograd path: operands follow the weights
def test_autograd_path_sees_torch_optimizer_updates(cuda):
    """model(x) -> loss.backward() -> torch.optim.AdamW.step() -> model(x): the second forward must use the updated
    weights in its GEMMs (bf16 operand copies re-cast), like any nn.Module."""
    model, sd, batches = _default_case()
    model = model.to(cuda).train()
    x, tg, il, tl = batches[0]
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2, weight_decay=0.0)  # large lr: stale operands would show
    logits = model(x.to(cuda), il)
    loss, _, dlogits = L.ctc_loss_fwd_bwd(logits.detach(), tg.to(cuda), (il // 4).to(cuda), tl.to(cuda))
    logits.backward(dlogits)
    opt.step()
    sd_new = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    out = model(x.to(cuda), il)
    ref_new = oc.forward(x, il, sd_new, 4, 2, training=True)
    ref_old = oc.forward(x, il, sd, 4, 2, training=True)
    rel = lambda a, b: ((a.float().cpu() - b).abs().max() / b.abs().max()).item()
    assert rel(out, ref_new) < 2e-2
    assert rel(out, ref_old) > 1e-1  # the step was large enough for the check to mean something
    # load_state_dict and in-place surgery through .data are picked up too
    model.load_state_dict(sd)
    assert rel(model(x.to(cuda), il), ref_old) < 2e-2
    with torch.no_grad():
        model.fc.weight.data.mul_(2.0)
        model.fc.bias.data.mul_(2.0)
    assert rel(model(x.to(cuda), il), 2.0 * ref_old) < 2e-2


# ------------------------------------------------------------------ the reference's outer loop
def test_fit_validate_checkpoint_and_resume(cuda, tmp_path):
    """Trainer.fit() as main.py drives it (reference trainer/trainer.py:284-319): epochs over a loader of collated
    batches, validation each epoch, periodic + best + final checkpoints, then a second Trainer with resume=True picks up
    at the next epoch with the same weights and optimizer state."""
    from turkish_asr_model_b200.data.dataset import collate_fn
    torch.manual_seed(21)
    model = TurkishASRModel(80, 256, 4, 1, 40, dropout=0.0).to(cuda)
    g = torch.Generator().manual_seed(22)
    items = [(torch.randn(70 + 9 * i, 80, generator=g), torch.randint(1, 40, (4 + i % 3,), generator=g)) for i in range(8)]
    train = [collate_fn(items[i: i + 2]) for i in range(0, 8, 2)] + [(None, None, None, None)]  # a failed batch is skipped
    valid = [collate_fn(items[:3])]

    class C(Cfg):
        epochs = 2
        save_interval = 1
        output_model_path = "final_model.pt"
    cfg = C()
    cfg.checkpoint_dir = str(tmp_path)
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)
    tr = Trainer(model, train, opt, _sched(opt), cuda, cfg, None, valid_loader=valid, gradient_clip=1.0)
    v0 = tr.validate(0)
    tr.fit()
    v2 = tr.validate(2)
    assert np.isfinite(v0) and np.isfinite(v2) and v2 < v0          # two epochs of training lowered the validation loss
    assert tr.global_step == 8 and tr.scheduler.last_epoch == 8
    for name in ("checkpoint_epoch_1.pt", "checkpoint_epoch_2.pt", "best_model.pt", "final_model.pt"):
        assert os.path.exists(os.path.join(tmp_path, name)), name
    assert tr.best_val_loss <= v2 + 1e-6 and tr.best_val_loss < v0  # best over the epochs (epoch 1 may beat epoch 2)

    torch.manual_seed(99)
    model2 = TurkishASRModel(80, 256, 4, 1, 40, dropout=0.0).to(cuda)
    cfg2 = C()
    cfg2.checkpoint_dir, cfg2.resume, cfg2.epochs = str(tmp_path), True, 3
    opt2 = torch.optim.AdamW(model2.parameters(), lr=5e-4, weight_decay=1e-6)
    tr2 = Trainer(model2, train, opt2, _sched(opt2), cuda, cfg2, None, valid_loader=valid, gradient_clip=1.0)
    tr2.load_checkpoint()
    assert tr2.start_epoch == 3 and tr2.global_step == 8 and tr2._opt_step == 8
    assert abs(tr2.validate(2) - v2) < 1e-4 * abs(v2)                # same weights and BatchNorm buffers
    tr2.fit()                                                        # runs epoch 3 only
    assert tr2.global_step == 12
