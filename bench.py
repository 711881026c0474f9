#!/usr/bin/env python
"""Headline benchmark: training audio-seconds/second of the waveform -> log-mel -> Conformer -> CTC
(fwd + bwd + clip + AdamW) hot path, BASELINE.json config[1]:

    default Conformer-CTC (80 mel, d_model 256, 4 heads, 8 blocks, V = 1000), bf16 operands,
    batch 64 of bucketed 5-15 s synthetic 16 kHz utterances per GPU, dropout 0.1.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  `value` = real (unpadded) audio seconds of all ranks / device time of K
steps with the waveforms already resident in HBM; `e2e` = the same through Trainer.train_step_waveforms with
pinned HOST buffers (H2D of waveforms/targets/lengths and D2H of the loss inside the timed region).
`--impl reference` times the CPU oracle port of the reference (oracle/, plain PyTorch on the host cores) on a
bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
CFG = dict(n_mels=80, d_model=256, n_heads=4, n_blocks=8, vocab=1000, dropout=0.1, batch=64)
METRIC = "train_audio_seconds_per_second"
UNIT = "audio-s/s"


def workload_config(n_gpus):
    name = "BASELINE configs[1]: default Conformer-CTC" if CFG["d_model"] == 256 else "BASELINE configs[2]: Conformer-M"
    return {
        "workload": "%s (80 mel, d_model=%d, %d heads, %d blocks, V=1000) full "
                    "training step (log-mel + fwd + CTC + bwd + clip + AdamW), batch 64 bucketed 5-15 s synthetic 16 kHz "
                    "utterances per GPU, dropout 0.1" % (name, CFG["d_model"], CFG["n_heads"], CFG["n_blocks"]),
        "per_gpu_batch": CFG["batch"], "global_batch": CFG["batch"] * n_gpus, "utterance_seconds": "U[5,15] bucketed",
        "parallelism": "dp%d" % n_gpus,
        "l2": "no explicit flush: every step streams a new batch and >2 GB of activations (>> 126 MB L2)",
    }


# --------------------------------------------------------------------------------------------- workload
def make_epoch(n_utts=4096, seed=1234):
    g = torch.Generator().manual_seed(seed)
    dur = 5.0 + 10.0 * torch.rand(n_utts, generator=g)
    n_samples = torch.round(dur * SR).to(torch.int64)
    return n_samples


def make_batches(n_batches, rank, world, seed=1234):
    """Bucketed batches exactly as the (rank-sharded) BucketingSampler yields them."""
    from turkish_asr_model_b200.data.dataset import BucketingSampler
    n_samples = make_epoch(seed=seed)
    sizes = [44 + 2 * int(n) for n in n_samples]
    sampler = BucketingSampler(None, CFG["batch"], shuffle=True, drop_last=False, rank=rank, world_size=world, seed=seed,
                               lengths=sizes)
    flat = list(iter(sampler))
    B = CFG["batch"]
    batches = []
    g = torch.Generator().manual_seed(seed + 17 + rank)
    for i in range(n_batches):
        idx = flat[(i * B) % (len(flat) - B + 1): (i * B) % (len(flat) - B + 1) + B]
        ns = n_samples[idx]
        tl = torch.round(4.0 * ns.double() / SR).to(torch.int64)
        smax = int(tl.max())
        targets = torch.zeros(B, smax, dtype=torch.int64)
        for b in range(B):
            targets[b, : int(tl[b])] = torch.randint(1, CFG["vocab"], (int(tl[b]),), generator=g)
        batches.append({"n_samples": ns, "targets": targets, "target_lengths": tl, "seed": seed * 1000 + rank * 100 + i})
    return batches


def synth_waves(batch, device):
    """0.1 * randn waveforms, every 4th utterance with 0.5 s of leading silence (exercises top_db)."""
    ns = batch["n_samples"]
    B, nmax = ns.shape[0], int(ns.max())
    g = torch.Generator(device=device).manual_seed(batch["seed"])
    w = 0.1 * torch.randn(B, nmax, generator=g, device=device)
    ar = torch.arange(nmax, device=device)[None, :]
    w = w * (ar < ns.to(device)[:, None])
    w[::4, : SR // 2] = 0.0
    return w


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step_fn(threads):
    """The reference's CPU path restated by the oracle (plain PyTorch fp32 autograd on the host cores):
    per-utterance log-mel (numpy), model forward, log_softmax + CTCLoss, backward, clip_grad_norm_, AdamW.
    The oracle is only the thing being *timed* here (cpu_baseline / --impl reference), never the product."""
    import numpy as np
    from oracle import conformer as oc
    from oracle import mel as om
    from turkish_asr_model_b200.model import TurkishASRModel
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    skel = TurkishASRModel(CFG["n_mels"], CFG["d_model"], CFG["n_heads"], CFG["n_blocks"], CFG["vocab"], dropout=0.0)
    sd = {k: v.detach().clone() for k, v in skel.state_dict().items()}
    pnames = [n for n, _ in skel.named_parameters()]
    params = {n: sd[n].requires_grad_(True) for n in pnames}
    opt = torch.optim.AdamW(list(params.values()), lr=5e-4, weight_decay=1e-6)
    fb, win = om.melscale_fbanks(), om.hann_periodic()

    def step(waves, ns, targets, tl):
        feats, frames = om.log_mel_batch(waves.numpy(), ns.tolist(), fb=fb, window=win)
        x = torch.from_numpy(feats.astype(np.float32))
        il = torch.from_numpy(frames)
        opt.zero_grad(set_to_none=True)
        logits = oc.forward(x, il, sd, CFG["n_heads"], CFG["n_blocks"], training=True)
        loss = oc.ctc_loss_torch(logits, targets, il, tl)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params.values() if p.grad is not None], 1.0)
        opt.step()
        return float(loss.detach())

    return step


def cpu_sample(batch, n_utts=8):
    """Bounded sample of the workload: the first n_utts utterances of a bucketed batch."""
    ns = batch["n_samples"][:n_utts]
    g = torch.Generator().manual_seed(batch["seed"])
    nmax = int(ns.max())
    w = 0.1 * torch.randn(n_utts, nmax, generator=g)
    tl = batch["target_lengths"][:n_utts]
    targets = batch["targets"][:n_utts, : int(tl.max())]
    return w, ns, targets, tl


def time_cpu(steps, warmup, threads, batches):
    step = cpu_reference_step_fn(threads)
    total_audio, total_t = 0.0, 0.0
    for i in range(warmup + steps):
        w, ns, targets, tl = cpu_sample(batches[i % len(batches)])
        t0 = time.perf_counter()
        step(w, ns, targets, tl)
        dt = time.perf_counter() - t0
        if i >= warmup:
            total_audio += float(ns.sum()) / SR
            total_t += dt
    return total_audio / total_t, total_t / max(steps, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batches = make_batches(4, 0, 1)
    steps = max(1, min(args.steps, 3))
    warmup = 1
    val, sec_per_step = time_cpu(steps, warmup, threads, batches)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "first 8 utterances of each bucketed batch of 64 (5-15 s), full train step, fp32, "
                                   "oracle port of the reference on torch CPU (dropout not modelled: favours the CPU arm)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch.distributed as dist
    from turkish_asr_model_b200 import _lib as L
    from turkish_asr_model_b200.model import TurkishASRModel
    from turkish_asr_model_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    L.lib()

    torch.manual_seed(0)
    model = TurkishASRModel(CFG["n_mels"], CFG["d_model"], CFG["n_heads"], CFG["n_blocks"], CFG["vocab"],
                            dropout=CFG["dropout"]).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-6)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=5e-4, total_steps=100000, pct_start=0.1, anneal_strategy="cos")

    class Cfg:
        log_interval = 10 ** 9
    trainer = Trainer(model, None, opt, sched, dev, Cfg(), None, gradient_clip=1.0, accumulation_steps=1)
    if args.no_graphs:
        trainer.use_cuda_graphs = False

    K, W = args.steps, args.warmup
    # every distinct batch shape is seen (and its CUDA graph captured) during warm-up
    n_distinct = max(1, min(K + W, 12, W))
    batches = make_batches(n_distinct, rank, world)
    dev_batches = []
    for b in batches:
        dev_batches.append({"waves": synth_waves(b, dev), "n_samples": b["n_samples"].to(dev), "targets": b["targets"].to(dev),
                            "target_lengths": b["target_lengths"].to(dev), "tmax": 1 + int(b["n_samples"].max()) // 160,
                            "audio_s": float(b["n_samples"].sum()) / SR,
                            "padded_s": float(b["n_samples"].max()) * len(b["n_samples"]) / SR})

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def resident_step(i):
        b = dev_batches[i % n_distinct]
        return trainer.train_step_waveforms(b["waves"], b["n_samples"], b["targets"], b["target_lengths"], tmax=b["tmax"])

    # ---- warm-up + timed region (device-resident inputs)
    for i in range(W):
        resident_step(i)
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = L.lib().tasr_launch_count() + trainer.graph_kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    audio = 0.0
    padded = 0.0  # audio seconds including the zero padding to the longest utterance of each batch (what is computed)
    e0.record()
    for i in range(K):
        loss = resident_step(W + i)
        audio += dev_batches[(W + i) % n_distinct]["audio_s"]
        padded += dev_batches[(W + i) % n_distinct]["padded_s"]
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = L.lib().tasr_launch_count() + trainer.graph_kernel_launches - launches0
    final_loss = float(loss.detach())

    # ---- end-to-end: pinned host buffers -> H2D -> step -> D2H of the loss, every step
    host_batches = []
    for b, db in zip(batches, dev_batches):
        host_batches.append({"waves": db["waves"].cpu().pin_memory(), "n_samples": b["n_samples"].pin_memory(),
                             "targets": b["targets"].pin_memory(), "target_lengths": b["target_lengths"].pin_memory()})
    # The loss of every step is copied to pinned host memory; the host reads it one step late (it waits for step
    # i-1 while step i runs), so logging never drains the GPU.  Every step's H2D and D2H is inside the timed region.
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
    seen = {"n": 0, "last": 0.0}

    def e2e_step(i):
        hb = host_batches[i % n_distinct]
        l = trainer.train_step_waveforms(hb["waves"], hb["n_samples"], hb["targets"], hb["target_lengths"])
        k = seen["n"] & 1
        loss_host[k].copy_(l.detach().reshape(1), non_blocking=True)
        loss_evt[k].record()
        if seen["n"] > 0:
            loss_evt[k ^ 1].synchronize()  # the user reads the previous step's loss
            seen["last"] = float(loss_host[k ^ 1])
        seen["n"] += 1
        return hb

    for i in range(min(W, 3)):
        e2e_step(i)
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    audio_e2e, h2d = 0.0, 0
    f0.record()
    for i in range(K):
        hb = e2e_step(W + i)
        audio_e2e += float(hb["n_samples"].sum()) / SR
        h2d += hb["waves"].numel() * 4 + hb["targets"].numel() * 8 + hb["n_samples"].numel() * 8 * 2
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)
    clock_info = clocks.stop() if rank == 0 else None

    # ---- max over ranks, sum of audio
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    a = torch.tensor([audio, audio_e2e, padded], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t[0]), float(t[1])
    audio, audio_e2e, padded = float(a[0]), float(a[1]), float(a[2])

    # ---- roofline of the dominant kernel (the tcgen05 GEMM), CUDA events around each launch, rank 0
    roofline = None
    cpu_baseline = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        hbm_peak = float(peaks.get("hbm_gbs", 6550.0))
        # All tcgen05 GEMM launches of one step are recorded (argument structs, operands kept alive) and then
        # replayed back to back as ONE CUDA graph between two CUDA events: device time of exactly those kernels,
        # no host launch gaps.
        trainer.use_cuda_graphs = False
        trainer.world_size = 1  # rank-local pass: no collectives (the other ranks are not in this branch)
        L.GEMM_PROFILE = []
        resident_step(0)
        torch.cuda.synchronize()
        prof, L.GEMM_PROFILE = L.GEMM_PROFILE, None
        flops = sum(p[0] for p in prof)

        def gemm_bytes(shape):
            """Algorithmic HBM bytes of one launch: both operands once + every output (+ the saved tensor its epilogue
            reads); bf16 = 2 B, fp32 residual stream / weight gradients = 4 B."""
            M_, N_, K_, epi, _, _ = shape
            ab = 2 * (M_ * K_ + N_ * K_)
            if epi == L.EPI_RESID:
                return ab + 8 * M_ * N_                      # fp32 residual read + fp32 write
            if epi in (L.EPI_SWIGLU, L.EPI_GLU):
                return ab + 2 * M_ * N_ + 2 * M_ * (N_ // 2)  # gate|up saved (bf16) + activation (bf16)
            if epi in (L.EPI_SWIGLU_BWD, L.EPI_GLU_BWD):
                return ab + 4 * M_ * N_ + 4 * M_ * N_         # gate|up read + d gate|d up written (bf16, 2 N columns each)
            if epi in (L.EPI_SILU, L.EPI_SILU_BWD):
                return ab + 4 * M_ * N_
            if epi == L.EPI_ATOMIC:
                return ab + 4 * M_ * N_
            return ab + 2 * M_ * N_
        gbytes = float(sum(gemm_bytes(p[3]) for p in prof))
        gg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gg):
            L.gemm_replay(prof)
        gg.replay()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        r0.record()
        for _ in range(reps):
            gg.replay()
        r1.record()
        torch.cuda.synchronize()
        gms = r0.elapsed_time(r1) / reps
        step_ms_this_batch = None
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        trainer.use_cuda_graphs = not args.no_graphs
        resident_step(0)
        s0.record()
        resident_step(0)
        s1.record()
        torch.cuda.synchronize()
        step_ms_this_batch = s0.elapsed_time(s1)
        tensor_view = {"achieved": flops / (gms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                       "frac": flops / (gms * 1e-3) / 1e12 / peak, "flops_per_step": flops,
                       "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"}
        # the same launches against the HBM roof (most of them have K = 256: operands + outputs dominate)
        hbm_view = {"achieved": gbytes / (gms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": gbytes / (gms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_step": gbytes,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6550 GB/s"}
        # the binding roof is the one the family is closer to (larger fraction = larger lower bound on its time)
        bound = "hbm" if hbm_view["frac"] >= tensor_view["frac"] else "tensor"
        prim = hbm_view if bound == "hbm" else tensor_view
        roofline = {"bound": bound, "kernel": "gemm_tc_kernel (persistent tcgen05 bf16 GEMM; all %d fwd/dgrad/wgrad launches "
                                              "of one step replayed back to back)" % len(prof),
                    "achieved": prim["achieved"], "peak": prim["peak"], "unit": prim["unit"], "frac": prim["frac"],
                    "traffic": None, "peak_source": prim["peak_source"],
                    "launches_per_step": len(prof), "gemm_ms_per_step": gms, "step_share": gms / step_ms_this_batch,
                    "tensor_view": tensor_view, "hbm_view": hbm_view}
        del prof
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            v, spp = time_cpu(2, 1, threads, batches)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": "first 8 utterances of 2 bucketed batches (5-15 s), full train step in fp32 on the "
                                      "oracle port of the reference (torch CPU; dropout not modelled, which favours the CPU arm), "
                                      "%.1f s/step" % spp}

    if rank == 0:
        line = {
            "metric": METRIC, "value": audio / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": audio_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d // K,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "final_loss": final_loss,
            # `value` counts real (unpadded) audio; the same run counted in padded seconds
            "padded_audio_seconds_per_second": padded / (ms * 1e-3),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graphs", action="store_true",
                    help="eager launches instead of per-shape CUDA graphs (under DP: bucketed all-reduce overlapped with backward)")
    ap.add_argument("--model", default="default", choices=["default", "conformer-m"],
                    help="default = BASELINE configs[1] (the headline); conformer-m = configs[2] (d_model 512, 8 heads, 16 blocks)")
    args = ap.parse_args()
    if args.model == "conformer-m":
        CFG.update(d_model=512, n_heads=8, n_blocks=16)
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    run_gpu(args)


if __name__ == "__main__":
    main()
