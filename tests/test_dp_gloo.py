"""CPU, world_size 2 over gloo: the data-parallel host logic of the Trainer (gradient buckets in the order
backward completes them, asynchronous all-reduce per bucket, rank-sharded sampler)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from turkish_asr_model_b200.data.dataset import BucketingSampler
    from turkish_asr_model_b200.engine import FlatParams
    from turkish_asr_model_b200.model import TurkishASRModel
    from turkish_asr_model_b200.trainer import Trainer

    torch.manual_seed(0)
    model = TurkishASRModel(80, 128, 2, 3, 24)
    flat = FlatParams(model)

    class FakeEngine:  # stands in for the CUDA engine: writes rank-dependent gradients segment by segment
        n_blocks = 3

        def backward(self, tape, dlogits, on_segment_done=None):
            order = tape
            for k, (lo, hi) in enumerate(order):
                flat.grads[lo:hi] = float(rank + 1) * (k + 1)
                on_segment_done(k)

    class Cfg:
        log_interval = 10 ** 9

    tr = Trainer(model, None, None, None, "cpu", Cfg(), None, bucket_bytes=1 << 20, use_cuda_graphs=False)
    assert tr.world_size == world
    eng = FakeEngine()
    order = tr._buckets(eng, flat)
    # segments tile [0, live) exactly and come in backward-completion order: fc, blocks 2..0, subsampler
    assert sorted(order)[0][0] == 0 and sorted(order)[-1][1] == flat.live_numel
    assert sum(hi - lo for lo, hi in order) == flat.live_numel
    assert order[0][0] == flat.offsets["fc.weight"] and order[-1][0] == 0
    handles = tr._backward_with_allreduce(eng, flat, order, None)
    for h in handles:
        h.wait()
    ok = True
    for k, (lo, hi) in enumerate(order):
        expect = float(sum(r + 1 for r in range(world))) * (k + 1)
        ok = ok and bool(torch.all(flat.grads[lo:hi] == expect))
    ok = ok and bool(torch.all(flat.grads[flat.live_numel:] == 0))  # dead parameters are never reduced
    # rank-sharded sampler: ranks agree on the step count and do not overlap
    sizes = [(i * 7919) % 1000 + 100 for i in range(300)]
    mine = list(iter(BucketingSampler(None, 4, lengths=sizes, rank=rank, world_size=world, seed=3)))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = ok and len({len(g) for g in gathered}) == 1 and not (set(gathered[0]) & set(gathered[1]))
    q.put((rank, ok, len(handles)))
    dist.destroy_process_group()


def test_dp_buckets_and_sharding_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(n >= 2 for _, _, n in res)  # more than one bucket was all-reduced
