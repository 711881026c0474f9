"""Drop-in mirror of reference trainer/trainer.py (Trainer): same constructor arguments and public methods
(fit / train_epoch / validate / save_checkpoint / load_checkpoint).  The per-batch step of
trainer/trainer.py:160-200 (autocast forward, log_softmax + CTCLoss, backward, clip_grad_norm_, AdamW,
scheduler) is replaced by `train_step`, a fixed sequence of libtasr kernels on flat buffers:

    [log-mel] -> encoder forward -> fused log-softmax/CTC loss+grad -> encoder backward
              -> (DP) bucketed NCCL all-reduce overlapped with backward -> global-norm clip + AdamW

Differences kept deliberate and documented: bf16 operands without GradScaler (the reference
uses fp16 + GradScaler), no host synchronisation per step (the NaN check of :179 is done on the device: a
non-finite gradient norm skips the update), data parallelism (absent from the reference)."""
import collections
import glob
import os
import time
import weakref
from typing import Optional

import torch

from .. import _lib as L
from ..data.preprocessing import AudioPreprocessor


class _NullLogger:
    def info(self, *a, **k):
        pass

    warning = error = info


def _clear_seed_ptr(addr):
    try:
        lib = L.lib()
        if lib.tasr_get_dropout_seed_ptr() == addr:
            lib.tasr_set_dropout_seed_ptr(None)
    except Exception:
        pass


class Trainer:
    def __init__(self, model, train_loader, optimizer, scheduler, device, config, logger, valid_loader=None,
                 tokenizer=None, gradient_clip: float = 1.0, accumulation_steps: int = 1, *, process_group=None,
                 bucket_bytes: int = 25 << 20, preprocessor: Optional[AudioPreprocessor] = None,
                 use_cuda_graphs: bool = True, max_graph_samples: int = 16000 * 20, max_cached_graphs: int = 32,
                 graph_len_quantum: int = 1):
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
        # captured steps are keyed by the exact batch shape and kept in an LRU of `max_cached_graphs` entries (real
        # variable-length data would otherwise capture without bound).  `graph_len_quantum` > 1 rounds the padded
        # waveform length up to a multiple of that many samples so that more batches share a capture; it is opt-in
        # because extra padding frames change GroupNorm / BatchNorm statistics (padding is real compute in the
        # reference, SURVEY.md finding 4).
        self.max_cached_graphs = max(1, int(max_cached_graphs))
        self.graph_len_quantum = max(1, int(graph_len_quantum))
        self._graphs = collections.OrderedDict()
        self._graph_pool = None
        self._static = None
        self._seed_dev = None
        self._seed_finalizer = None
        self.graph_kernel_launches = 0  # kernels executed through graph replays (bench.py's gpu_launches)
        self.timing = None  # set to {} to collect per-step device events (bench.py: compute vs exposed communication)

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
            self._hyper_done = [None] * len(self._hyper_ring)  # event: the H2D copy out of that pinned slot has run
            # device-side dropout counter: starts from torch's seed mixed with the rank (torch.manual_seed controls
            # it; data-parallel ranks draw different masks), advanced by the optimizer kernel sequence every step
            rank = torch.distributed.get_rank(self.pg) if self.world_size > 1 else 0
            s0 = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + (rank + 1) * 0xD1B54A32D192ED03) & 0x3FFFFFFFFFFFFFFF
            self._seed_dev = torch.full((1,), s0, dtype=torch.int64, device=flat.device)
            self._bind_seed()
            # the library keeps a raw pointer to the counter: clear it when this trainer (and its tensor) goes away
            self._seed_finalizer = weakref.finalize(self, _clear_seed_ptr, self._seed_dev.data_ptr())
            views = flat.grad_views()
            for name, p in self.model.named_parameters():
                if name in views:
                    p.grad = views[name]  # user-visible .grad aliases the flat gradient buffer
        return eng, flat

    def _bind_seed(self):
        """Point the library's dropout-seed counter at THIS trainer's device counter (several trainers / models may
        live in one process; kernels already captured in a CUDA graph keep the address they were captured with)."""
        L.check(L.lib().tasr_set_dropout_seed_ptr(self._seed_dev.data_ptr()))

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
        self._bind_seed()
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

    def train_step_waveforms(self, waves, n_samples, targets, target_lengths, tmax=None, spec_params=None):
        """Same step starting from raw 16 kHz waveforms (B, Nmax) + lengths: the log-mel front-end runs on the
        GPU in front of the encoder (data/preprocessing.py path of the reference runs on CPU workers).
        spec_params (B, n_masks, 3) int32 = (axis, start, end) per utterance applies SpecAugment to the normalised
        features (reference data/dataset.py:98-99; draw them with SpecAugment.mask_params)."""
        if self.preprocessor is None:
            self.preprocessor = AudioPreprocessor(device="cuda")
        dev = self.model.fc.weight.device
        if tmax is None and not n_samples.is_cuda:
            tmax = 1 + int(n_samples.max()) // 160
        B, nmax = waves.shape
        if self.use_cuda_graphs and tmax is not None and self.accumulation_steps == 1 and nmax <= self.max_graph_samples:
            return self._graphed_step(waves, n_samples, targets, target_lengths, tmax, spec_params)
        feats, frames = self.preprocessor.extract_features_batch(waves.to(dev, non_blocking=True), n_samples, tmax)
        if spec_params is not None:
            L.specaugment_(feats, spec_params.to(device=dev, dtype=torch.int32).contiguous(), frames.to(dev))
        return self.train_step(feats, targets, frames, target_lengths)

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def _graphed_step(self, waves, n_samples, targets, target_lengths, tmax, spec_params=None):
        """The step is a fixed kernel sequence for a given batch shape, and bucketed batches recur every epoch,
        so each shape is captured once (torch.cuda.graph, one shared memory pool) and replayed: ~600 kernel
        launches collapse into one graph launch.  Per-step variation lives in device memory: the inputs (static
        buffers), the AdamW hyper-parameters (uploaded before the replay) and the dropout seed counter.

        Data parallel: the step is captured as TWO graphs, A = log-mel + forward + CTC + backward of the classifier
        and the Conformer blocks, B = backward of input_proj and the Conv2d subsampler.  After A the all-reduce of the
        block + classifier gradient range (95 % of the bytes) is launched asynchronously on NCCL's stream and runs
        under B (~2 ms of work); the small head range follows B; the fused clip + AdamW waits for both."""
        eng, flat = self._flat()
        dev = flat.device
        self.model.train()
        B, nmax = waves.shape
        q = self.graph_len_quantum
        if q > 1 and nmax % q:
            nmax_q = min((nmax + q - 1) // q * q, self.max_graph_samples)
            if nmax_q > nmax:
                tmax = 1 + nmax_q // 160
        else:
            nmax_q = nmax
        smax = targets.shape[1]
        if self._static is None or self._static["B"] != B:
            cap = self.max_graph_samples
            self._static = {"B": B, "waves": torch.zeros(B * cap, dtype=torch.float32, device=dev),
                            "n": torch.zeros(B, dtype=torch.int64, device=dev),
                            "targets": torch.zeros(B * 1024, dtype=torch.int64, device=dev),
                            "tl": torch.zeros(B, dtype=torch.int64, device=dev)}
            self._graphs = collections.OrderedDict()
            self._graph_pool = torch.cuda.graph_pool_handle()
        if smax > 1024:  # beyond the static target buffer: this batch runs eagerly (graphs stay enabled for the others)
            feats, frames = self.preprocessor.extract_features_batch(waves.to(dev, non_blocking=True), n_samples, tmax)
            if spec_params is not None:
                L.specaugment_(feats, spec_params.to(device=dev, dtype=torch.int32).contiguous(), frames.to(dev))
            return self.train_step(feats, targets, frames, target_lengths)
        st = self._static
        s_w = st["waves"][: B * nmax_q].view(B, nmax_q)
        s_t = st["targets"][: B * smax].view(B, smax)
        if nmax_q > nmax:
            s_w[:, nmax:].zero_()
        s_wv = s_w[:, :nmax]
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
            s_wv.copy_(stage, non_blocking=True)
            st["stage_free"][k].record()
        else:
            s_wv.copy_(waves, non_blocking=True)
        st["n"].copy_(n_samples, non_blocking=True)
        s_t.copy_(targets, non_blocking=True)
        st["tl"].copy_(target_lengths, non_blocking=True)
        nmask = 0
        if spec_params is not None:
            nmask = int(spec_params.shape[1])
            if "spec" not in st or st["spec"].shape[1] != nmask:
                st["spec"] = torch.zeros(B, nmask, 3, dtype=torch.int32, device=dev)
            st["spec"].copy_(spec_params, non_blocking=True)
        self._optimizer_host()
        # weights edited behind the trainer's back (load_state_dict, manual surgery): refresh the bf16 operands eagerly
        flat.refresh_shadow()
        key = (B, nmax_q, tmax, smax, nmask)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture_step(eng, flat, key, s_w, s_t)
            self._graphs[key] = entry
            while len(self._graphs) > self.max_cached_graphs:
                self._graphs.popitem(last=False)  # least recently used; its memory returns to the shared pool
        else:
            self._graphs.move_to_end(key)
        tm = self.timing
        if tm is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        entry["graph"].replay()
        self.graph_kernel_launches += entry["launches"]
        if self.world_size > 1:
            import torch.distributed as dist
            lo = flat.offsets["blocks.0.norm_ff1.norm.weight"] if eng.n_blocks else flat.offsets["fc.weight"]
            h_body = dist.all_reduce(flat.grads[lo: flat.live_numel], group=self.pg, async_op=True)
            entry["graph_head"].replay()
            if tm is not None:
                ev[1].record()
            h_head = dist.all_reduce(flat.grads[:lo], group=self.pg, async_op=True)
            h_body.wait()
            h_head.wait()
            if tm is not None:
                ev[2].record()
            self._optimizer_device(flat)
        if tm is not None:
            ev[3].record()
            tm.setdefault("events", []).append(ev if self.world_size > 1 else [ev[0], ev[3]])
        return entry["loss"]

    def _capture_step(self, eng, flat, key, s_w, s_t):
        B, nmax, tmax, smax, nmask = key
        dev = flat.device
        st = self._static
        # everything that is created lazily must exist before capture
        self.preprocessor._tables(dev)
        eng.cos_sin(4096, dev)
        eng.ensure_side_stream(dev)
        L.workspace(eng.workspace_bytes(B, tmax, smax), dev)
        flat.refresh_shadow()
        eng.device_seed = True
        self._bind_seed()
        torch.cuda.synchronize()
        lib = L.lib()
        # capture on a high-priority stream: the main chain's CTAs are scheduled ahead of the weight-gradient kernels
        # that the engine runs on its (lowest-priority) side stream
        if getattr(self, "_cap_stream", None) is None:
            self._cap_stream = torch.cuda.Stream(device=dev, priority=-1)
        dp = self.world_size > 1
        n0 = lib.tasr_launch_count()
        g = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread may touch CUDA while this thread captures
        with torch.cuda.graph(g, pool=self._graph_pool, stream=self._cap_stream, capture_error_mode="thread_local"):
            feats, frames = self.preprocessor.extract_features_batch(s_w, st["n"], tmax)
            if nmask:
                L.specaugment_(feats, st["spec"], frames)
            if not dp:
                loss = self._step_features(eng, flat, feats, s_t, frames, st["tl"], host_opt=False, device_opt=True)
            else:
                flat.grads[: flat.live_numel].zero_()
                logits, tape = eng.forward(feats, frames, True, self.model.dropout_p, save=True)
                loss, _, dlogits = L.ctc_loss_fwd_bwd(logits, s_t, frames // 4, st["tl"], blank=self.blank, grad_scale=1.0)
                loss = loss[0]
                dx0 = eng.backward_blocks(tape, dlogits)
        entry = {"graph": g, "loss": loss, "graph_head": None}
        if dp:
            # the head graph reads dx0 and the subsampler part of the tape out of graph A's allocations: both objects
            # stay referenced until B is captured, afterwards the memory may go back to the (shared) pool
            gh = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gh, pool=self._graph_pool, stream=self._cap_stream, capture_error_mode="thread_local"):
                eng.backward_head(tape, dx0)
            entry["graph_head"] = gh
            del tape, dx0, dlogits, logits
        entry["launches"] = lib.tasr_launch_count() - n0
        return entry

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

    def _optimizer_host(self, scheduled=True):
        """Host half of the optimizer step: hyper-parameters of THIS step go to the device buffer the fused
        kernel reads (pinned ring so that an in-flight copy is never overwritten), counters, LR schedule.
        scheduled=False is the leftover-gradient flush at the end of an epoch (reference trainer/trainer.py:213-219):
        the optimizer steps, but neither the scheduler nor global_step advance."""
        lr, b1, b2, eps, wd = self._group()
        self._opt_step += 1
        t = self._opt_step
        slot = t % len(self._hyper_ring)
        if self._hyper_done[slot] is not None:
            self._hyper_done[slot].synchronize()  # the copy queued 16 steps ago has read this pinned slot
        h = self._hyper_ring[slot]
        h[0], h[1], h[2], h[3], h[4] = lr, b1, b2, eps, wd
        h[5], h[6] = 1.0 - b1 ** t, 1.0 - b2 ** t
        h[7] = float(self.gradient_clip) if self.gradient_clip else 0.0
        h[8] = float(self.world_size)  # all-reduce sums over ranks; the mean is folded into the clip scale
        self._hyper.copy_(h, non_blocking=True)
        if self._hyper_done[slot] is None:
            self._hyper_done[slot] = torch.cuda.Event()
        self._hyper_done[slot].record()
        if not scheduled:
            return
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

    def _optimizer_step(self, flat, scheduled=True):
        self._optimizer_host(scheduled)
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
            eng, flat = self._flat()
            self._optimizer_step(flat, scheduled=False)
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
                if eng.precision == "fp32":  # the reference validates in fp32 on CPU (trainer/trainer.py:227-282)
                    logits = eng.forward_f32(features.to(dev), il, False)
                else:
                    logits, _ = eng.forward(features.to(dev), il, False, 0.0, save=False)
                loss, _, _ = L.ctc_loss_fwd_bwd(logits, targets.to(dev), il // 4, target_lengths.to(dev, torch.int64),
                                                blank=self.blank, want_grad=False)
                total += float(loss)
                n += 1
        avg = total / max(n, 1)
        self.logger.info("Epoch %d Validation | Loss: %.4f" % (epoch, avg))
        return avg

    # ------------------------------------------------------------------ checkpoints (reference format)
    def _opt_params(self):
        """Parameters in the index order torch.optim uses in its state_dict (group by group)."""
        return [p for g in self.optimizer.param_groups for p in g["params"]] if self.optimizer is not None else []

    def _export_optimizer_state(self, flat):
        """optimizer_state_dict in the layout torch.optim.AdamW itself writes (reference trainer/trainer.py:93):
        {'state': {idx: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}, filled from the flat moment buffers
        of the fused kernel, so that the reference's Trainer.load_checkpoint resumes from a checkpoint of this one."""
        if self.optimizer is None:
            return {}
        sd = self.optimizer.state_dict()
        if flat is None or flat.exp_avg is None or self._opt_step == 0:
            return sd
        by_ptr = {p.data_ptr(): n for n, p in self.model.named_parameters()}
        state = {}
        for idx, p in enumerate(self._opt_params()):
            name = by_ptr.get(p.data_ptr())
            if name is None or flat.offsets[name] >= flat.live_numel:
                continue  # parameters that never receive a gradient have no optimizer state in torch either
            state[idx] = {"step": torch.tensor(float(self._opt_step)),
                          "exp_avg": flat.view(flat.exp_avg, name).detach().clone().cpu(),
                          "exp_avg_sq": flat.view(flat.exp_avg_sq, name).detach().clone().cpu()}
        sd["state"] = state
        return sd

    def _import_optimizer_state(self, flat, sd):
        """Inverse of _export_optimizer_state: per-parameter torch AdamW state -> flat moment buffers + step count.
        Returns the number of parameters restored."""
        state = sd.get("state", {}) if isinstance(sd, dict) else {}
        if not state:
            return 0
        by_ptr = {p.data_ptr(): n for n, p in self.model.named_parameters()}
        params = self._opt_params()
        restored, step = 0, 0
        for idx, st in state.items():
            idx = int(idx)
            if idx >= len(params) or "exp_avg" not in st:
                continue
            name = by_ptr.get(params[idx].data_ptr())
            if name is None or st["exp_avg"].numel() != flat.view(flat.exp_avg, name).numel():
                continue
            flat.view(flat.exp_avg, name).copy_(st["exp_avg"].to(torch.float32).view(flat.shapes[name]))
            flat.view(flat.exp_avg_sq, name).copy_(st["exp_avg_sq"].to(torch.float32).view(flat.shapes[name]))
            step = max(step, int(float(st.get("step", 0))))
            restored += 1
        if restored:
            self._opt_step = step
        return restored

    def _sync_bn_buffers(self):
        """Data parallel: BatchNorm batch statistics are rank-local during training (like un-synced DDP); at
        checkpoint time the running statistics are averaged over the ranks (in place, all ranks end up identical)."""
        if self.world_size <= 1:
            return
        import torch.distributed as dist
        bufs = [b for n, b in self.model.named_buffers() if n.endswith(("running_mean", "running_var"))]
        if not bufs:
            return
        flat = torch.cat([b.detach().float().reshape(-1) for b in bufs])
        dist.all_reduce(flat, group=self.pg)
        flat /= self.world_size
        off = 0
        for b in bufs:
            b.copy_(flat[off: off + b.numel()].view_as(b))
            off += b.numel()

    def save_checkpoint(self, epoch: int, name: Optional[str] = None, is_best: bool = False) -> None:
        """reference trainer/trainer.py:84-110 (same dict keys and value layouts).  Under data parallelism every rank
        must call it (BatchNorm buffers are reconciled collectively); rank 0 writes the file."""
        self._sync_bn_buffers()
        if self.world_size > 1 and torch.distributed.get_rank(self.pg) != 0:
            return
        ckpt_dir = self.config.checkpoint_dir
        os.makedirs(ckpt_dir, exist_ok=True)
        flat = self.model.engine().flat
        state = {
            "epoch": epoch,
            "global_step": self.global_step,
            "model_state_dict": {k: v.detach().clone().cpu() for k, v in self.model.state_dict().items()},
            "optimizer_state_dict": self._export_optimizer_state(flat),
            "scheduler_state_dict": self.scheduler.state_dict() if self.scheduler is not None else {},
            # bf16 needs no loss scaling; a neutral GradScaler state keeps the file loadable by the reference trainer
            # (its scaler.load_state_dict rejects an empty dict when AMP is enabled, trainer/trainer.py:135)
            "scaler_state_dict": {"scale": 65536.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000,
                                  "_growth_tracker": 0},
            "best_val_loss": self.best_val_loss,
            "config": dict(vars(self.config)) if hasattr(self.config, "__dict__") else {},
        }
        path = os.path.join(ckpt_dir, name if name is not None else "checkpoint_epoch_%d.pt" % epoch)
        torch.save(state, path)
        self.logger.info("Checkpoint saved: %s" % path)
        if is_best:
            torch.save(state, os.path.join(ckpt_dir, "best_model.pt"))

    def load_checkpoint(self) -> None:
        """reference trainer/trainer.py:112-145: resume from the newest checkpoint_epoch_*.pt by mtime.  Checkpoints
        written by the reference trainer load too: its per-parameter AdamW state is mapped into the flat moment
        buffers of the fused optimizer kernel."""
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
        osd = ck.get("optimizer_state_dict")
        if osd and self.optimizer is not None:
            try:
                groups = osd.get("param_groups")
                if groups:  # hyper-parameters (lr as left by the scheduler, betas, eps, weight decay)
                    for g, saved in zip(self.optimizer.param_groups, groups):
                        g.update({k: v for k, v in saved.items() if k != "params"})
            except Exception as e:  # param-group mismatch across versions
                self.logger.warning("optimizer hyper-parameters not restored: %s" % e)
            n = self._import_optimizer_state(flat, osd)
            if n == 0 and osd.get("state"):
                self.logger.warning("optimizer moments could not be mapped onto this model: AdamW state restarts from zero")
        elif self.optimizer is not None:
            self.logger.warning("checkpoint has no optimizer state: AdamW moments restart from zero")
        if ck.get("scheduler_state_dict") and self.scheduler is not None:
            self.scheduler.load_state_dict(ck["scheduler_state_dict"])
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
