"""Call-compatible replacement of nn.CTCLoss(blank=0, zero_infinity=True) at the reference call site
trainer/trainer.py:76,167-173, fused with the log-softmax: takes the encoder LOGITS (B, T', V) and produces
the same mean-reduced loss; the gradient w.r.t. the logits comes out of the same kernel sequence."""
import torch

from . import _lib as L


class _CTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, input_lengths, target_lengths, blank):
        dev = logits.device
        loss, nll, dlogits = L.ctc_loss_fwd_bwd(logits.detach(), targets.to(dev), input_lengths.to(dev, torch.int64),
                                                target_lengths.to(dev, torch.int64), blank=blank, grad_scale=1.0,
                                                want_grad=True)
        ctx.save_for_backward(dlogits)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g.to(dlogits.dtype), None, None, None, None


class FusedCTCLoss(torch.nn.Module):
    """loss = FusedCTCLoss(blank=0)(logits (B,T',V), targets (B,Smax), input_lengths (B,), target_lengths (B,)).
    `from_log_probs(log_probs (T',B,V), ...)` keeps the reference's (T, B, V) log-prob signature: since
    log_softmax is idempotent the same kernels apply."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = True):
        super().__init__()
        if reduction != "mean" or not zero_infinity:
            raise NotImplementedError("the reference uses reduction='mean', zero_infinity=True")
        self.blank = blank

    def forward(self, logits, targets, input_lengths, target_lengths):
        return _CTCFn.apply(logits, targets, input_lengths, target_lengths, self.blank)

    def from_log_probs(self, log_probs, targets, input_lengths, target_lengths):
        return self.forward(log_probs.permute(1, 0, 2), targets, input_lengths, target_lengths)
