"""autograd glue between the drop-in nn.Modules and the kernels (no arithmetic of its own).

* `model_forward`: the whole encoder as ONE autograd node on engine.ConformerEngine (the fast path).
* Stand-alone sub-module forwards (SwiGLUFeedForward, TransposeGroupNorm, RelativeMultiHeadAttention,
  ConformerConvModule, ConformerBlock): one autograd node per module built from the same C-ABI ops, so the
  reference's sub-module API works outside TurkishASRModel too (operands bf16, results fp32).
"""
import torch

from . import _lib as L

DH = 64


# ------------------------------------------------------------------------------------------------ whole model
class _EncoderFn(torch.autograd.Function):
    """forward runs ConformerEngine.forward and keeps its tape; backward runs ConformerEngine.backward and hands
    the parameter gradients (copies of the flat gradient buffer views) back to autograd."""

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
    flat = eng.ensure_flat()
    # Drop-in autograd path: the weights may have been edited by anything between two calls (torch.optim step,
    # load_state_dict, p.data.copy_ -- the last one is invisible to version counters), so the bf16 GEMM operands are
    # re-cast on every call (one pass over the parameters, ~0.1 GB).  The Trainer path keeps them fresh itself.
    flat.shadow_fresh = False
    params = tuple(model.parameters())
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    if eng.precision == "fp32":
        if need_grad:
            raise L.TasrError("fp32 operand mode is forward-only: call the model under torch.no_grad() (training runs "
                              "in the bf16 operand mode)")
        return eng.forward_f32(x, input_lengths, model.training)
    return _EncoderFn.apply(model, x, input_lengths, need_grad, *params)


# ------------------------------------------------------------------------------------------------ helpers
def _seed(p):
    return int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) * 64 if p > 0 else 0


def _bf(w):
    return L.cast_bf16(w.detach().float().contiguous())


def _f32(x):
    return x.detach().float().contiguous()


def _wgrad(dy, x, out_f, in_f, tokens):
    gw = torch.zeros(out_f, in_f, dtype=torch.float32, device=dy.device)
    tiles = ((out_f + 127) // 128) * ((in_f + 255) // 256)
    splits = max(1, min((tokens + 63) // 64, 148 // max(tiles, 1)))
    L.gemm(out_f, in_f, tokens, dy, dy.stride(0), x, x.stride(0), L.EPI_ATOMIC, gw, in_f, a_mn=1, b_mn=1, split_k=splits)
    return gw


def _colsum(dy):
    out = torch.zeros(dy.shape[1], dtype=torch.float32, device=dy.device)
    L.colsum_add(dy, out)
    return out


# ------------------------------------------------------------------------------------------------ GroupNorm
class _GroupNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps):
        xf = _f32(x)
        y, stats = L.groupnorm_fwd(xf, groups, _f32(weight), _f32(bias), eps=eps, out_bf16=False)
        ctx.save_for_backward(xf, stats, weight)
        ctx.groups = groups
        return y

    @staticmethod
    def backward(ctx, dy):
        xf, stats, weight = ctx.saved_tensors
        dx = torch.empty_like(xf)
        dg = torch.zeros_like(weight, dtype=torch.float32)
        db = torch.zeros_like(weight, dtype=torch.float32)
        L.groupnorm_bwd(_f32(dy), xf, ctx.groups, stats, _f32(weight), dx, False, dg, db)
        return dx, dg, db, None, None


def groupnorm_module_forward(mod, x):
    L.require_cuda(x)
    return _GroupNormFn.apply(x, mod.norm.weight, mod.norm.bias, mod.norm.num_groups, mod.norm.eps)


# ------------------------------------------------------------------------------------------------ SwiGLU FFN
class _FeedForwardFn(torch.autograd.Function):
    """y = residual + alpha * dropout2(linear2(dropout1(swiglu(linear1(x)))))   (model/conformer.py:18-26,121)"""

    @staticmethod
    def forward(ctx, x, residual, w1, b1, w2, b2, alpha, p, training):
        B, T, d = x.shape
        M, dff = B * T, w2.shape[1]
        drop = float(p) if training else 0.0
        seed = _seed(drop)
        xb = L.cast_bf16(_f32(x)).view(M, d)
        w1b, w2b = _bf(w1), _bf(w2)
        gv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=x.device)
        h = torch.empty(M, dff, dtype=torch.bfloat16, device=x.device)
        L.gemm(M, dff, d, xb, d, w1b, d, L.EPI_SWIGLU, h, dff, out2=gv, ldo2=2 * dff, bias=_f32(b1), n_half=dff,
               drop_p=drop, seed=seed)
        res = _f32(residual).view(M, d) if residual is not None else torch.zeros(M, d, dtype=torch.float32, device=x.device)
        out = torch.empty(M, d, dtype=torch.float32, device=x.device)
        L.gemm(M, d, dff, h, dff, w2b, dff, L.EPI_RESID, out, d, bias=_f32(b2), aux=res, ldaux=d, alpha=alpha,
               drop_p=drop, seed=seed + 1)
        ctx.save_for_backward(xb, gv, h, w1b, w2b)
        ctx.meta = (B, T, d, dff, alpha, drop, seed, residual is not None)
        return out.view(B, T, d)

    @staticmethod
    def backward(ctx, dy):
        xb, gv, h, w1b, w2b = ctx.saved_tensors
        B, T, d, dff, alpha, drop, seed, has_res = ctx.meta
        M = B * T
        dyf = _f32(dy).view(M, d)
        dyb = L.cast_bf16(dyf, alpha=alpha, drop_p=drop, seed=seed + 1)
        dw2 = _wgrad(dyb, h, d, dff, M)
        db2 = _colsum(dyb)
        dgv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=dy.device)
        L.gemm(M, dff, d, dyb, d, w2b, dff, L.EPI_SWIGLU_BWD, dgv, 2 * dff, b_mn=1, aux=gv, ldaux=2 * dff, n_half=dff,
               drop_p=drop, seed=seed)
        dw1 = _wgrad(dgv, xb, 2 * dff, d, M)
        db1 = _colsum(dgv)
        dx = torch.empty(M, d, dtype=torch.float32, device=dy.device)
        L.gemm(M, d, 2 * dff, dgv, 2 * dff, w1b, d, L.EPI_STORE, dx, d, b_mn=1, out_f32=1)
        return dx.view(B, T, d), (dy if has_res else None), dw1, db1, dw2, db2, None, None, None


def ff_module_forward(mod, x, residual=None, alpha=1.0):
    L.require_cuda(x)
    return _FeedForwardFn.apply(x, residual, mod.linear1.weight, mod.linear1.bias, mod.linear2.weight, mod.linear2.bias,
                                alpha, mod.dropout1.p, mod.training)


# ------------------------------------------------------------------------------------------------ attention
def _key_lengths(mask, B, T, device):
    """(B,1,1,T) boolean prefix mask of model/conformer.py:187-202 -> (B,) valid key counts."""
    if mask is None:
        return None
    m = mask.reshape(B, -1, T)[:, -1, :] != 0
    lens = m.sum(-1).to(torch.int64)
    expect = torch.arange(T, device=m.device)[None, :] < lens[:, None]
    if not torch.equal(m, expect):
        raise L.TasrError("the B200 attention kernel supports key-padding (prefix) masks only")
    return lens.to(device)


def _cos_sin(mod, T, device):
    inv_freq = mod.rotary_emb.inv_freq.to(device=device, dtype=torch.float32)
    t = torch.arange(T, device=device, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    return torch.stack([freqs.cos(), freqs.sin()], dim=-1).contiguous()


class _AttentionFn(torch.autograd.Function):
    """y = residual + linear_out(MQA(RoPE(q), RoPE(k), v))   (model/attention.py:195-251)"""

    @staticmethod
    def forward(ctx, x, residual, wq, bq, wk, bk, wv, bv, wo, bo, H, key_len, cs, p, training):
        B, T, d = x.shape
        M = B * T
        drop = float(p) if training else 0.0
        seed = _seed(drop)
        xb = L.cast_bf16(_f32(x)).view(M, d)
        wqkv = _bf(torch.cat([wq.detach(), wk.detach(), wv.detach()], 0))
        bqkv = torch.cat([bq.detach(), bk.detach(), bv.detach()], 0).float().contiguous()
        nq = d + 2 * DH
        qkv = torch.empty(M, nq, dtype=torch.bfloat16, device=x.device)
        L.gemm(M, nq, d, xb, d, wqkv, d, L.EPI_ROPE, qkv, nq, bias=bqkv, aux=cs, n_half=T, remap_p0=d + DH)
        c, lse2 = L.mqa_fwd(qkv, B, T, H, d, key_len, drop_p=drop, seed=seed)
        wob = _bf(wo)
        res = _f32(residual).view(M, d) if residual is not None else torch.zeros(M, d, dtype=torch.float32, device=x.device)
        out = torch.empty(M, d, dtype=torch.float32, device=x.device)
        L.gemm(M, d, d, c, d, wob, d, L.EPI_RESID, out, d, bias=_f32(bo), aux=res, ldaux=d, alpha=1.0)
        ctx.save_for_backward(xb, qkv, c, lse2, wqkv, wob, cs, key_len if key_len is not None else torch.empty(0))
        ctx.meta = (B, T, d, H, drop, seed, residual is not None, key_len is not None)
        return out.view(B, T, d)

    @staticmethod
    def backward(ctx, dy):
        xb, qkv, c, lse2, wqkv, wob, cs, key_len = ctx.saved_tensors
        B, T, d, H, drop, seed, has_res, has_len = ctx.meta
        M, nq = B * T, d + 2 * DH
        dyb = L.cast_bf16(_f32(dy).view(M, d))
        dwo = _wgrad(dyb, c, d, d, M)
        dbo = _colsum(dyb)
        dctx = torch.empty(M, d, dtype=torch.bfloat16, device=dy.device)
        L.gemm(M, d, d, dyb, d, wob, d, L.EPI_STORE, dctx, d, b_mn=1)
        dqkv = L.mqa_bwd(qkv, c, dctx, lse2, B, T, H, d, key_len if has_len else None, cs, drop_p=drop, seed=seed)
        dwqkv = _wgrad(dqkv, xb, nq, d, M)
        dbqkv = _colsum(dqkv)
        dx = torch.empty(M, d, dtype=torch.float32, device=dy.device)
        L.gemm(M, d, nq, dqkv, nq, wqkv, d, L.EPI_STORE, dx, d, b_mn=1, out_f32=1)
        return (dx.view(B, T, d), (dy if has_res else None), dwqkv[:d], dbqkv[:d], dwqkv[d:d + DH], dbqkv[d:d + DH],
                dwqkv[d + DH:], dbqkv[d + DH:], dwo, dbo, None, None, None, None, None)


def attention_module_forward(mod, x, mask, residual=None):
    L.require_cuda(x)
    if mod.d_head != DH:
        raise L.TasrError("the B200 attention kernel needs d_model == 64 * n_heads")
    B, T, _ = x.shape
    key_len = _key_lengths(mask, B, T, x.device)
    cs = _cos_sin(mod, T, x.device)
    return _AttentionFn.apply(x, residual, mod.linear_q.weight, mod.linear_q.bias, mod.linear_k.weight, mod.linear_k.bias,
                              mod.linear_v.weight, mod.linear_v.bias, mod.linear_out.weight, mod.linear_out.bias,
                              mod.n_heads, key_len, cs, mod.dropout_p, mod.training)


# ------------------------------------------------------------------------------------------------ conv module
class _ConvModuleFn(torch.autograd.Function):
    """y = residual + pw2(silu(bn(dwconv(glu(pw1(gn(x)))))))   (model/conformer.py:76-88)"""

    @staticmethod
    def forward(ctx, x, residual, gn_w, gn_b, pw1_w, pw1_b, dw_w, dw_b, bn_w, bn_b, pw2_w, pw2_b, mod):
        B, T, d = x.shape
        M = B * T
        xf = _f32(x)
        G = mod.norm.norm.num_groups
        xn, st = L.groupnorm_fwd(xf, G, _f32(gn_w), _f32(gn_b), eps=mod.norm.norm.eps)
        pw1b, pw2b = _bf(pw1_w.detach().view(2 * d, d)), _bf(pw2_w.detach().view(d, d))
        ab = torch.empty(M, 2 * d, dtype=torch.bfloat16, device=x.device)
        u = torch.empty(M, d, dtype=torch.bfloat16, device=x.device)
        L.gemm(M, d, d, xn.view(M, d), d, pw1b, d, L.EPI_GLU, u, d, out2=ab, ldo2=2 * d, bias=_f32(pw1_b), n_half=d)
        dww = _f32(dw_w).view(d, 31)
        w, part = L.dwconv_fwd(u.view(B, T, d), dww, _f32(dw_b), want_stats=mod.training)
        bn = mod.batch_norm
        bnst = L.bn_finalize(part, d, M, bn.eps, bn.momentum if bn.momentum is not None else 0.1, mod.training,
                             bn.running_mean, bn.running_var, bn.num_batches_tracked)
        s = L.bn_silu_fwd(w, bnst, _f32(bn_w), _f32(bn_b))
        res = _f32(residual).view(M, d) if residual is not None else torch.zeros(M, d, dtype=torch.float32, device=x.device)
        out = torch.empty(M, d, dtype=torch.float32, device=x.device)
        L.gemm(M, d, d, s.view(M, d), d, pw2b, d, L.EPI_RESID, out, d, bias=_f32(pw2_b), aux=res, ldaux=d, alpha=1.0)
        ctx.save_for_backward(xf, st, xn, ab, u, w, bnst, s, pw1b, pw2b, dww, _f32(gn_w), _f32(bn_w), _f32(bn_b))
        ctx.meta = (B, T, d, G, residual is not None)
        return out.view(B, T, d)

    @staticmethod
    def backward(ctx, dy):
        xf, st, xn, ab, u, w, bnst, s, pw1b, pw2b, dww, gn_w, bn_w, bn_b = ctx.saved_tensors
        B, T, d, G, has_res = ctx.meta
        M = B * T
        dev = dy.device
        dyb = L.cast_bf16(_f32(dy).view(M, d))
        dpw2 = _wgrad(dyb, s.view(M, d), d, d, M)
        dpw2b = _colsum(dyb)
        ds = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        L.gemm(M, d, d, dyb, d, pw2b, d, L.EPI_STORE, ds, d, b_mn=1)
        dbn_w, dbn_b = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
        dw = L.bn_silu_bwd(ds, w, bnst, bn_w, bn_b, dbn_w, dbn_b)
        ddw_w, ddw_b = torch.zeros(d, 31, device=dev), torch.zeros(d, device=dev)
        dab = L.dwconv_bwd(dw.view(B, T, d), u.view(B, T, d), ab.view(B, T, 2 * d), dww, ddw_w, ddw_b).view(M, 2 * d)
        dpw1 = _wgrad(dab, xn.view(M, d), 2 * d, d, M)
        dpw1b = _colsum(dab)
        dxn = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        L.gemm(M, d, 2 * d, dab, 2 * d, pw1b, d, L.EPI_STORE, dxn, d, b_mn=1)
        dx = torch.empty(B, T, d, dtype=torch.float32, device=dev)
        dgn_w, dgn_b = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
        L.groupnorm_bwd(dxn.view(B, T, d), xf, G, st, gn_w, dx, False, dgn_w, dgn_b)
        return (dx, (dy if has_res else None), dgn_w, dgn_b, dpw1.view(2 * d, d, 1), dpw1b, ddw_w.view(d, 1, 31), ddw_b,
                dbn_w, dbn_b, dpw2.view(d, d, 1), dpw2b, None)


def conv_module_forward(mod, x, residual=None):
    L.require_cuda(x)
    return _ConvModuleFn.apply(x, residual, mod.norm.norm.weight, mod.norm.norm.bias, mod.pointwise_conv1.weight,
                               mod.pointwise_conv1.bias, mod.depthwise_conv.weight, mod.depthwise_conv.bias,
                               mod.batch_norm.weight, mod.batch_norm.bias, mod.pointwise_conv2.weight,
                               mod.pointwise_conv2.bias, mod)


# ------------------------------------------------------------------------------------------------ block
def block_module_forward(mod, x, mask):
    """model/conformer.py:114-135; every residual add happens inside a GEMM epilogue."""
    L.require_cuda(x)
    x = ff_module_forward(mod.ff1, groupnorm_module_forward(mod.norm_ff1, x), residual=x, alpha=0.5)
    x = attention_module_forward(mod.attn, groupnorm_module_forward(mod.norm_attn, x), mask, residual=x)
    x = conv_module_forward(mod.conv, x, residual=x)
    x = ff_module_forward(mod.ff2, groupnorm_module_forward(mod.norm_ff2, x), residual=x, alpha=0.5)
    return groupnorm_module_forward(mod.final_norm, x)
