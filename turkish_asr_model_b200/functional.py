"""autograd glue between the drop-in nn.Modules and the kernel engine (no arithmetic of its own)."""
import torch

from . import _lib as L


class _EncoderFn(torch.autograd.Function):
    """Whole-model forward/backward as one autograd node: forward runs ConformerEngine.forward and keeps its
    tape; backward runs ConformerEngine.backward and hands the parameter gradients (views of the flat
    gradient buffer) back to autograd."""

    @staticmethod
    def forward(ctx, model, feats, input_lengths, need_grad, *params):
        eng = model.engine()
        logits, tape = eng.forward(feats, input_lengths, model.training, model.dropout_p, save=need_grad)
        ctx.model = model
        ctx.tape = tape
        ctx.names = [n for n, _ in model.named_parameters()]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        eng = ctx.model.engine()
        flat = eng.flat
        flat.grads[: flat.live_numel].zero_()
        eng.backward(ctx.tape, dlogits)
        ctx.tape = None
        views = flat.grad_views()
        grads = tuple(views[n].clone() if n in views else None for n in ctx.names)
        return (None, None, None, None) + grads


def model_forward(model, x, input_lengths):
    L.require_cuda(x)
    eng = model.engine()
    eng.ensure_flat()
    params = tuple(model.parameters())
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _EncoderFn.apply(model, x, input_lengths, need_grad, *params)


def _not_standalone(name):
    raise L.TasrError(
        "%s.forward is only available as part of TurkishASRModel on the B200 path in this round "
        "(the fused engine owns the activations); call the model instead" % name)


def ff_module_forward(mod, x):
    _not_standalone("SwiGLUFeedForward")


def groupnorm_module_forward(mod, x):
    L.require_cuda(x)
    y, _ = L.groupnorm_fwd(x.contiguous().float(), mod.norm.num_groups, mod.norm.weight.detach().float(),
                           mod.norm.bias.detach().float(), eps=mod.norm.eps, out_bf16=False)
    return y


def conv_module_forward(mod, x):
    _not_standalone("ConformerConvModule")


def attention_module_forward(mod, x, mask):
    _not_standalone("RelativeMultiHeadAttention")


def block_module_forward(mod, x, mask):
    _not_standalone("ConformerBlock")
