"""Worker of tests/test_dp_nccl_gpu.py (launched with torch.distributed.run, one rank per GPU, NCCL):
data-parallel Trainer steps on real GPUs against the sum of per-rank oracle gradients."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import conformer as oc  # noqa: E402
from oracle import mel as om  # noqa: E402
from turkish_asr_model_b200.model import TurkishASRModel  # noqa: E402
from turkish_asr_model_b200.trainer import Trainer  # noqa: E402


class Cfg:
    log_interval = 10 ** 9
    checkpoint_dir = os.environ.get("TASR_TEST_CKPT", "/tmp/tasr_dp_ckpt")


def rank_batch(r, V):
    g = torch.Generator().manual_seed(100 + r)
    ns = torch.tensor([16000 + 800 * r, 12000 + 400 * r])
    w = torch.zeros(2, int(ns.max()))
    for b in range(2):
        w[b, : ns[b]] = 0.1 * torch.randn(int(ns[b]), generator=g)
    return w, ns, torch.randint(1, V, (2, 6), generator=g), torch.tensor([6, 4])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    V = 48
    for graphs in (True, False):
        torch.manual_seed(0)
        model = TurkishASRModel(80, 256, 4, 2, V, dropout=0.0)
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        pnames = [n for n, _ in model.named_parameters()]
        model = model.to(dev).train()
        opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)
        tr = Trainer(model, None, opt, None, dev, Cfg(), None, gradient_clip=1.0, use_cuda_graphs=graphs)
        assert tr.world_size == world
        w, ns, tg, tl = rank_batch(rank, V)
        loss = float(tr.train_step_waveforms(w.to(dev), ns, tg, tl))
        torch.cuda.synchronize()
        eng, flat = tr._flat()
        # oracle: every rank recomputes ALL ranks' gradients on the CPU and sums them
        sdr = {k: (v.clone().requires_grad_(True) if k in pnames else v) for k, v in sd.items()}
        ref_losses = []
        for r in range(world):
            wr, nr, tgr, tlr = rank_batch(r, V)
            feats, frames = om.log_mel_batch(wr.numpy(), nr.tolist())
            xl = oc.forward(torch.from_numpy(feats.astype(np.float32)), torch.from_numpy(frames), sdr, 4, 2, training=True)
            l = oc.ctc_loss_torch(xl, tgr, torch.from_numpy(frames), tlr)
            l.backward()
            ref_losses.append(float(l.detach()))
        assert abs(loss - ref_losses[rank]) < 2e-2 * abs(ref_losses[rank]), (loss, ref_losses)
        views = flat.grad_views()
        rels = []
        for name in pnames:
            gref = sdr[name].grad
            if gref is None or name.endswith("depthwise_conv.bias"):
                continue
            got = views[name].detach().cpu().double()  # SUM over ranks (the mean is folded into the clip scale)
            rels.append(((got - gref.double()).abs().max() / gref.abs().max()).item())
        assert max(rels) < 8e-2 and np.median(rels) < 2e-2, (graphs, max(rels), float(np.median(rels)))
        # grad norm = norm of the MEAN gradient; identical on every rank, and so are the updated parameters
        ref_norm = float(torch.sqrt(sum((sdr[n].grad.double() ** 2).sum() for n in pnames if sdr[n].grad is not None))) / world
        assert abs(float(tr.last_grad_norm) - ref_norm) < 3e-2 * ref_norm, (float(tr.last_grad_norm), ref_norm)
        mine = flat.params[: flat.live_numel].clone()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        for other in gathered:
            assert torch.equal(other, mine), "parameters diverged across ranks"
        # a second step replays the captured graphs (or the eager bucketed path) and stays in lock-step
        loss2 = float(tr.train_step_waveforms(w.to(dev), ns, tg, tl))
        assert np.isfinite(loss2) and loss2 < loss
        # checkpoint: every rank calls, rank 0 writes, BatchNorm running statistics are reconciled
        tr.save_checkpoint(1, name="dp_%d.pt" % int(graphs))
        rm = model.blocks[0].conv.batch_norm.running_mean.clone()
        allrm = [torch.empty_like(rm) for _ in range(world)]
        dist.all_gather(allrm, rm)
        for other in allrm:
            assert torch.equal(other, rm)
        dist.barrier()
        if rank == 0:
            assert os.path.exists(os.path.join(Cfg.checkpoint_dir, "dp_%d.pt" % int(graphs)))
    if rank == 0:
        print("DP_NCCL_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
