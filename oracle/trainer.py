"""Training-loop oracle (plain torch on CPU, fp32).  TEST INFRASTRUCTURE ONLY: imported by tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py, never by the product package.

Restates reference trainer/trainer.py:147-225 (Trainer.train_epoch) on top of oracle.conformer.forward:
  :160-176  forward, permute, log_softmax, CTCLoss(blank 0, zero_infinity) on input_lengths // 4, loss / accumulation
  :178-181  NaN loss -> the batch is skipped (no backward, no optimizer step, counters untouched)
  :184      backward (GradScaler is the identity for fp32/bf16)
  :187-198  every `accumulation_steps` batches (by batch index): clip_grad_norm_(params, clip), AdamW.step,
            scheduler.step, zero_grad, global_step += 1
  :213-219  leftover accumulated gradients: clip + AdamW.step + zero_grad, WITHOUT scheduler.step / global_step
Pinned to the reference by tests/test_oracle_golden.py::test_trainer_oracle_vs_golden (tests/golden/trainer_golden.npz,
written by tools/make_golden.py from the reference's own Trainer on CPU).
"""
import torch

from . import conformer as oc


class TrainerOracle:
    def __init__(self, sd, param_names, n_heads, n_blocks, lr=5e-4, weight_decay=1e-6, clip=1.0, accumulation_steps=1,
                 scheduler_fn=None):
        self.sd = {k: v.detach().clone() for k, v in sd.items()}
        self.param_names = list(param_names)
        for n in self.param_names:
            self.sd[n].requires_grad_(True)
        self.params = [self.sd[n] for n in self.param_names]
        self.n_heads, self.n_blocks = n_heads, n_blocks
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)
        self.sched = scheduler_fn(self.opt) if scheduler_fn is not None else None
        self.clip = clip
        self.accum = max(1, int(accumulation_steps))
        self.global_step = 0
        self.last_grad_norm = None

    def _optimizer_step(self, with_scheduler):
        self.last_grad_norm = float(torch.nn.utils.clip_grad_norm_(self.params, self.clip))
        self.opt.step()
        if with_scheduler:
            if self.sched is not None:
                self.sched.step()
            self.global_step += 1
        self.opt.zero_grad()

    def train_epoch(self, batches):
        """batches: iterable of (features (B,T,F), targets (B,Smax), input_lengths (B,), target_lengths (B,)).
        Returns (average loss, list of per-batch losses)."""
        self.opt.zero_grad()
        losses, seen = [], 0
        for idx, (x, targets, il, tl) in enumerate(batches):
            bn_state = {}
            logits = oc.forward(x, il, self.sd, self.n_heads, self.n_blocks, training=True, bn_state=bn_state)
            loss = oc.ctc_loss_torch(logits, targets, il, tl) / self.accum
            if torch.isnan(loss):
                continue
            loss.backward()
            with torch.no_grad():  # BatchNorm running statistics advance like nn.BatchNorm1d in train mode
                for k, v in bn_state.items():
                    self.sd[k] = v
            if (idx + 1) % self.accum == 0:
                self._optimizer_step(True)
            losses.append(float(loss.detach()) * self.accum)
            seen += 1
        if seen % self.accum != 0:
            self._optimizer_step(False)
        return sum(losses) / max(seen, 1), losses
