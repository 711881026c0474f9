"""Drop-in mirror of reference trainer/trainer.py (Trainer): same constructor arguments and public methods
(fit / train_epoch / validate / save_checkpoint / load_checkpoint).  The per-batch step of
trainer/trainer.py:160-200 (autocast forward, log_softmax + CTCLoss, backward, clip_grad_norm_, AdamW,
scheduler) is replaced by `train_step`, a fixed sequence of libtasr kernels on flat buffers:

    [log-mel] -> encoder forward -> fused log-softmax/CTC loss+grad -> encoder backward
              -> (DP) bucketed NCCL all-reduce overlapped with backward -> global-norm clip + AdamW

Differences kept deliberate and documented: bf16 operands without GradScaler (the reference
uses fp16 + GradScaler), no host synchronisation per step (the NaN check of :179 is done on the device: a
non-finite gradient norm skips the update), data parallelism (absent from the reference)."""
import glob
import os
import time
from typing import Optional

import torch

from .. import _lib as L
from ..data.preprocessing import AudioPreprocessor


class _NullLogger:
    def info(self, *a, **k):
        pass

    warning = error = info


class Trainer:
    def __init__(self, model, train_loader, optimizer, scheduler, device, config, logger, valid_loader=None,
                 tokenizer=None, gradient_clip: float = 1.0, accumulation_steps: int = 1, *, process_group=None,
                 bucket_bytes: int = 25 << 20, preprocessor: Optional[AudioPreprocessor] = None,
                 use_cuda_graphs: bool = True, max_graph_samples: int = 16000 * 20):
        self.model = model
        self.train_loader = train_loader
        self.valid_loader = valid_loader
        self.optimizer = optimizer
        self.scheduler = scheduler
        self.device = torch.device(device)
        self.config = config
        self.logger = logger if logger is not None else _NullLogger()
        self.tokenizer = tokenizer
        self.gradient_clip = gradient_clip
        self.accumulation_steps = max(1, int(accumulation_steps))
        self.metrics = None  # WER/CER need jiwer + the HF tokenizer (out of scope); token-id decoding is available
        if tokenizer is None:
            self.logger.warning("Tokenizer not provided! WER/CER calculation disabled.")
        self.blank = 0
        self.start_epoch = 1
        self.best_val_loss = float("inf")
        self.global_step = 0
        self.preprocessor = preprocessor
        # data parallel state
        self.pg = process_group
        self.world_size = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world_size = torch.distributed.get_world_size(process_group)
        self.bucket_bytes = bucket_bytes
        self._micro = 0
        self._opt_step = 0
        self._hyper = None
        self._sumsq = None
        self._norm = None
        self.last_grad_norm = None
        # CUDA graphs: one captured step per batch shape (bucketed batches recur every epoch), shared memory pool
        self.use_cuda_graphs = use_cuda_graphs
        self.max_graph_samples = max_graph_samples
        self._graphs = {}
        self._graph_pool = None
        self._static = None
        self._seed_dev = None
        self.graph_kernel_launches = 0  # kernels executed through graph replays (bench.py's gpu_launches)

    # ------------------------------------------------------------------ optimizer state on flat buffers
    def _flat(self):
        eng = self.model.engine()
        flat = eng.ensure_flat()
        if flat.exp_avg is None:
            flat.exp_avg = torch.zeros_like(flat.params)
            flat.exp_avg_sq = torch.zeros_like(flat.params)
            self._hyper = torch.zeros(9, dtype=torch.float32, device=flat.device)
            self._sumsq = torch.zeros(1, dtype=torch.float64, device=flat.device)
            self._norm = torch.zeros(1, dtype=torch.float32, device=flat.device)
            self._hyper_ring = [torch.zeros(9, dtype=torch.float32).pin_memory() for _ in range(16)]
            self._seed_dev = torch.zeros(1, dtype=torch.int64, device=flat.device)
            L.check(L.lib().tasr_set_dropout_seed_ptr(self._seed_dev.data_ptr()))
            views = flat.grad_views()
            for name, p in self.model.named_parameters():
                if name in views:
                    p.grad = views[name]  # user-visible .grad aliases the flat gradient buffer
        return eng, flat

    def _group(self):
        g = self.optimizer.param_groups[0] if self.optimizer is not None else {}
        lr = float(g.get("lr", 5e-4))
        b1, b2 = g.get("betas", (0.9, 0.999))
        return lr, float(b1), float(b2), float(g.get("eps", 1e-8)), float(g.get("weight_decay", 1e-6))

    def _buckets(self, eng, flat):
        """Contiguous [lo, hi) ranges of the flat gradient buffer in the order backward completes them:
        classifier, blocks n-1..0, subsampler (reverse registration order)."""
        o = flat.offsets
        marks = [o["blocks.%d.norm_ff1.norm.weight" % i] for i in range(eng.n_blocks)] + [o["fc.weight"], flat.live_numel]
        segs = [(marks[i], marks[i + 1]) for i in range(len(marks) - 1)]  # block 0..n-1, fc
        head = (0, marks[0])
        order = [segs[-1]] + segs[:-1][::-1] + [head]
        return order

    # ------------------------------------------------------------------ one training step
    def train_step(self, features, targets, input_lengths, target_lengths):
        """features (B, T, F) fp32 (device), targets (B, Smax) int64, input_lengths (B,) mel frames,
        target_lengths (B,).  Returns the (device, fp32) loss of this micro-batch; nothing synchronises."""
        eng, flat = self._flat()
        dev = flat.device
        self.model.train()
        return self._step_features(eng, flat, features, targets, input_lengths, target_lengths, host_opt=True)

    def _step_features(self, eng, flat, features, targets, input_lengths, target_lengths, host_opt, device_opt=True):
        dev = flat.device
        if self._micro == 0:
            flat.grads[: flat.live_numel].zero_()
        features = features.to(dev, non_blocking=True)
        targets = targets.to(dev, non_blocking=True)
        il = input_lengths.to(dev, dtype=torch.int64, non_blocking=True)
        tl = target_lengths.to(dev, dtype=torch.int64, non_blocking=True)
        logits, tape = eng.forward(features, il, True, self.model.dropout_p, save=True)
        loss, _, dlogits = L.ctc_loss_fwd_bwd(logits, targets, il // 4, tl, blank=self.blank,
                                              grad_scale=1.0 / self.accumulation_steps)
        last_micro = (self._micro + 1) % self.accumulation_steps == 0
        handles = []
        if self.world_size > 1 and last_micro and device_opt:
            handles = self._backward_with_allreduce(eng, flat, tape, dlogits)
        else:
            eng.backward(tape, dlogits)
        self._micro += 1
        if last_micro:
            for h in handles:
                h.wait()
            if host_opt:
                self._optimizer_host()
            if device_opt:
                self._optimizer_device(flat)
            self._micro = 0
        return loss[0]

    def train_step_waveforms(self, waves, n_samples, targets, target_lengths, tmax=None):
        """Same step starting from raw 16 kHz waveforms (B, Nmax) + lengths: the log-mel front-end runs on the
        GPU in front of the encoder (data/preprocessing.py path of the reference runs on CPU workers)."""
        if self.preprocessor is None:
            self.preprocessor = AudioPreprocessor(device="cuda")
        dev = self.model.fc.weight.device
        if tmax is None and not n_samples.is_cuda:
            tmax = 1 + int(n_samples.max()) // 160
        B, nmax = waves.shape
        if self.use_cuda_graphs and tmax is not None and self.accumulation_steps == 1 and nmax <= self.max_graph_samples:
            return self._graphed_step(waves, n_samples, targets, target_lengths, tmax)
        feats, frames = self.preprocessor.extract_features_batch(waves.to(dev, non_blocking=True), n_samples, tmax)
        return self.train_step(feats, targets, frames, target_lengths)

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def _graphed_step(self, waves, n_samples, targets, target_lengths, tmax):
        """The step is a fixed kernel sequence for a given batch shape, and bucketed batches recur every epoch,
        so each shape is captured once (torch.cuda.graph, one shared memory pool) and replayed: ~600 kernel
        launches collapse into one graph launch.  Per-step variation lives in device memory: the inputs (static
        buffers), the AdamW hyper-parameters (uploaded before the replay) and the dropout seed counter."""
        eng, flat = self._flat()
        dev = flat.device
        self.model.train()
        B, nmax = waves.shape
        smax = targets.shape[1]
        if self._static is None or self._static["B"] != B:
            cap = self.max_graph_samples
            self._static = {"B": B, "waves": torch.zeros(B * cap, dtype=torch.float32, device=dev),
                            "n": torch.zeros(B, dtype=torch.int64, device=dev),
                            "targets": torch.zeros(B * 1024, dtype=torch.int64, device=dev),
                            "tl": torch.zeros(B, dtype=torch.int64, device=dev)}
            self._graphs = {}
            self._graph_pool = torch.cuda.graph_pool_handle()
        if smax > 1024:
            self.use_cuda_graphs = False
            return self.train_step_waveforms(waves, n_samples, targets, target_lengths, tmax)
        st = self._static
        s_w = st["waves"][: B * nmax].view(B, nmax)
        s_t = st["targets"][: B * smax].view(B, smax)
        if not waves.is_cuda:
            # Host inputs: the (large) waveform H2D runs on a copy stream into one of two device staging buffers, so it
            # overlaps with the previous step's graph still executing; the main stream then only does a D2D copy.
            if "stage" not in st:
                cap = self.max_graph_samples
                st["stage"] = [torch.empty(B * cap, dtype=torch.float32, device=dev) for _ in range(2)]
                st["stage_free"] = [torch.cuda.Event(), torch.cuda.Event()]
                st["copy_stream"] = torch.cuda.Stream(device=dev)
                st["step"] = 0
                for e in st["stage_free"]:
                    e.record()
            k = st["step"] & 1
            st["step"] += 1
            stage = st["stage"][k][: B * nmax].view(B, nmax)
            cs = st["copy_stream"]
            cs.wait_event(st["stage_free"][k])  # the D2D that last read this staging buffer has run
            with torch.cuda.stream(cs):
                stage.copy_(waves, non_blocking=True)
                done = torch.cuda.Event()
                done.record(cs)
            torch.cuda.current_stream().wait_event(done)
            s_w.copy_(stage, non_blocking=True)
            st["stage_free"][k].record()
        else:
            s_w.copy_(waves, non_blocking=True)
        st["n"].copy_(n_samples, non_blocking=True)
        s_t.copy_(targets, non_blocking=True)
        st["tl"].copy_(target_lengths, non_blocking=True)
        self._optimizer_host()
        key = (B, nmax, tmax, smax)
        entry = self._graphs.get(key)
        if entry is None:
            # everything that is created lazily must exist before capture
            self.preprocessor._tables(dev)
            eng.cos_sin(4096, dev)
            eng.ensure_side_stream(dev)
            L.workspace(1, dev)
            flat.refresh_shadow()
            eng.device_seed = True
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.lib().tasr_launch_count()
            # thread_local: the NCCL watchdog thread may touch CUDA while this thread captures
            # capture on a high-priority stream: the main chain's CTAs are scheduled ahead of the weight-gradient kernels
            # that the engine runs on its (lowest-priority) side stream
            if getattr(self, "_cap_stream", None) is None:
                self._cap_stream = torch.cuda.Stream(device=dev, priority=-1)
            with torch.cuda.graph(g, pool=self._graph_pool, stream=self._cap_stream, capture_error_mode="thread_local"):
                feats, frames = self.preprocessor.extract_features_batch(s_w, st["n"], tmax)
                loss = self._step_features(eng, flat, feats, s_t, frames, st["tl"], host_opt=False,
                                           device_opt=(self.world_size == 1))
            entry = self._graphs[key] = (g, loss, L.lib().tasr_launch_count() - n0)
        entry[0].replay()
        self.graph_kernel_launches += entry[2]
        if self.world_size > 1:
            # data parallel: graph = mel + forward + CTC + backward; then ONE NCCL all-reduce over the flat
            # gradient buffer (71 MB ~ 0.3 ms on NVLink 5, a few % of the step) and the fused optimizer, eagerly.
            import torch.distributed as dist
            dist.all_reduce(flat.grads[: flat.live_numel], group=self.pg)
            self._optimizer_device(flat)
        return entry[1]

    def _backward_with_allreduce(self, eng, flat, tape, dlogits):
        """Backward with one asynchronous NCCL all-reduce per gradient bucket, issued as soon as the kernels
        producing that slice of the flat gradient buffer have been enqueued (overlaps with the rest of
        backward; ProcessGroupNCCL orders its stream after the current one)."""
        import torch.distributed as dist
        order = self._buckets(eng, flat)
        handles = []
        pending_lo, pending_hi = None, None

        def flush(force=False):
            nonlocal pending_lo, pending_hi
            if pending_lo is None:
                return
            if force or (pending_hi - pending_lo) * 4 >= self.bucket_bytes:
                handles.append(dist.all_reduce(flat.grads[pending_lo:pending_hi], group=self.pg, async_op=True))
                pending_lo = pending_hi = None

        def done(seg):
            nonlocal pending_lo, pending_hi
            lo, hi = seg
            if pending_lo is None:
                pending_lo, pending_hi = lo, hi
            elif hi == pending_lo:
                pending_lo = lo
            else:  # not adjacent (fc sits at the end of the buffer): send what we have
                flush(force=True)
                pending_lo, pending_hi = lo, hi
            flush()

        eng.backward(tape, dlogits, on_segment_done=lambda idx: done(order[idx]))
        flush(force=True)
        return handles

    def _optimizer_host(self):
        """Host half of the optimizer step: hyper-parameters of THIS step go to the device buffer the fused
        kernel reads (pinned ring so that an in-flight copy is never overwritten), counters, LR schedule."""
        lr, b1, b2, eps, wd = self._group()
        self._opt_step += 1
        t = self._opt_step
        h = self._hyper_ring[t % len(self._hyper_ring)]
        h[0], h[1], h[2], h[3], h[4] = lr, b1, b2, eps, wd
        h[5], h[6] = 1.0 - b1 ** t, 1.0 - b2 ** t
        h[7] = float(self.gradient_clip) if self.gradient_clip else 0.0
        h[8] = float(self.world_size)  # all-reduce sums over ranks; the mean is folded into the clip scale
        self._hyper.copy_(h, non_blocking=True)
        self.global_step += 1
        if self.scheduler is not None:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")  # torch warns because optimizer.step() is replaced by the fused kernel
                self.scheduler.step()

    def _optimizer_device(self, flat):
        """Device half: global grad norm + clip + AdamW + bf16 shadow refresh (graph-capturable)."""
        self._sumsq.zero_()
        n = flat.live_numel
        L.grad_sumsq(flat.grads[:n], self._sumsq)
        L.clip_adamw(flat.params[:n], flat.grads[:n], flat.exp_avg[:n], flat.exp_avg_sq[:n], flat.shadow[:n], self._hyper,
                     self._sumsq, self._norm)
        self._seed_dev.add_(0x632BE59BD9B4E019)  # fresh dropout masks for the next step / graph replay
        flat.shadow_fresh = True
        self.last_grad_norm = self._norm

    def _optimizer_step(self, flat):
        self._optimizer_host()
        self._optimizer_device(flat)

    # ------------------------------------------------------------------ epoch loops (reference API)
    def train_epoch(self, epoch: int) -> float:
        """reference trainer/trainer.py:147-225."""
        self.model.train()
        losses = []
        start = time.time()
        for batch_idx, batch in enumerate(self.train_loader):
            if batch[0] is None:
                continue
            features, targets, input_lengths, target_lengths = batch
            loss = self.train_step(features, targets, input_lengths, target_lengths)
            losses.append(loss)
            if (batch_idx + 1) % getattr(self.config, "log_interval", 10) == 0:
                lr = self._group()[0]
                self.logger.info("Epoch [%d/%s] Batch [%d/%d] Loss: %.4f LR: %.2e" % (
                    epoch, getattr(self.config, "epochs", "?"), batch_idx + 1, len(self.train_loader), float(loss), lr))
        if self._micro != 0:  # leftover accumulated gradients (reference :213-219): step without the scheduler
            sched, self.scheduler = self.scheduler, None
            eng, flat = self._flat()
            self._optimizer_step(flat)
            self.scheduler = sched
            self._micro = 0
        avg = float(torch.stack(losses).mean()) if losses else 0.0
        self.logger.info("Epoch %d Complete | Loss: %.4f | Time: %.1fs" % (epoch, avg, time.time() - start))
        return avg

    def validate(self, epoch: int) -> Optional[float]:
        """reference trainer/trainer.py:227-282 (loss only; WER/CER need the tokenizer + jiwer)."""
        if not self.valid_loader:
            return None
        self.model.eval()
        eng, flat = self._flat()
        dev = flat.device
        total, n = 0.0, 0
        with torch.no_grad():
            for batch in self.valid_loader:
                if batch[0] is None:
                    continue
                features, targets, input_lengths, target_lengths = batch
                il = input_lengths.to(dev, dtype=torch.int64)
                logits, _ = eng.forward(features.to(dev), il, False, 0.0, save=False)
                loss, _, _ = L.ctc_loss_fwd_bwd(logits, targets.to(dev), il // 4, target_lengths.to(dev, torch.int64),
                                                blank=self.blank, want_grad=False)
                total += float(loss)
                n += 1
        avg = total / max(n, 1)
        self.logger.info("Epoch %d Validation | Loss: %.4f" % (epoch, avg))
        return avg

    def save_checkpoint(self, epoch: int, name: Optional[str] = None, is_best: bool = False) -> None:
        """reference trainer/trainer.py:84-110 (same dict keys; rank 0 only under DP)."""
        if self.world_size > 1 and torch.distributed.get_rank(self.pg) != 0:
            return
        ckpt_dir = self.config.checkpoint_dir
        os.makedirs(ckpt_dir, exist_ok=True)
        flat = self.model.engine().flat
        state = {
            "epoch": epoch,
            "global_step": self.global_step,
            "model_state_dict": {k: v.detach().clone().cpu() for k, v in self.model.state_dict().items()},
            "optimizer_state_dict": self.optimizer.state_dict() if self.optimizer is not None else {},
            "scheduler_state_dict": self.scheduler.state_dict() if self.scheduler is not None else {},
            "scaler_state_dict": {},  # bf16: no GradScaler
            "fused_adamw": None if flat is None or flat.exp_avg is None else {
                "step": self._opt_step, "exp_avg": flat.exp_avg.cpu(), "exp_avg_sq": flat.exp_avg_sq.cpu()},
            "best_val_loss": self.best_val_loss,
            "config": dict(vars(self.config)) if hasattr(self.config, "__dict__") else {},
        }
        path = os.path.join(ckpt_dir, name if name is not None else "checkpoint_epoch_%d.pt" % epoch)
        torch.save(state, path)
        self.logger.info("Checkpoint saved: %s" % path)
        if is_best:
            torch.save(state, os.path.join(ckpt_dir, "best_model.pt"))

    def load_checkpoint(self) -> None:
        """reference trainer/trainer.py:112-145: resume from the newest checkpoint_epoch_*.pt by mtime."""
        if not getattr(self.config, "resume", False):
            return
        ckpts = sorted(glob.glob(os.path.join(self.config.checkpoint_dir, "checkpoint_epoch_*.pt")), key=os.path.getmtime)
        if not ckpts:
            self.logger.warning("No checkpoint found! Starting from scratch.")
            return
        ck = torch.load(ckpts[-1], map_location="cpu", weights_only=False)
        self.model.load_state_dict(ck["model_state_dict"])
        eng, flat = self._flat()
        flat.shadow_fresh = False
        if ck.get("optimizer_state_dict") and self.optimizer is not None:
            try:
                self.optimizer.load_state_dict(ck["optimizer_state_dict"])
            except Exception as e:  # param-group mismatch across versions
                self.logger.warning("optimizer state not restored: %s" % e)
        if ck.get("scheduler_state_dict") and self.scheduler is not None:
            self.scheduler.load_state_dict(ck["scheduler_state_dict"])
        fa = ck.get("fused_adamw")
        if fa is not None and fa["exp_avg"].numel() == flat.exp_avg.numel():
            flat.exp_avg.copy_(fa["exp_avg"])
            flat.exp_avg_sq.copy_(fa["exp_avg_sq"])
            self._opt_step = int(fa["step"])
        self.start_epoch = int(ck.get("epoch", 0)) + 1
        self.global_step = int(ck.get("global_step", 0))
        self.best_val_loss = ck.get("best_val_loss", float("inf"))
        self.logger.info("Loaded checkpoint. Resuming from Epoch %d" % self.start_epoch)

    def fit(self) -> None:
        """reference trainer/trainer.py:284-319."""
        self.load_checkpoint()
        epochs = self.config.epochs
        if self.start_epoch > epochs:
            self.logger.info("Training already completed.")
            return
        for epoch in range(self.start_epoch, epochs + 1):
            self.train_epoch(epoch)
            val_loss = self.validate(epoch)
            if epoch % getattr(self.config, "save_interval", 5) == 0:
                self.save_checkpoint(epoch)
            if val_loss is not None and val_loss < self.best_val_loss:
                self.best_val_loss = val_loss
                self.save_checkpoint(epoch, name="best_model.pt", is_best=True)
        self.save_checkpoint(epochs, name=getattr(self.config, "output_model_path", "final_model.pt"))
