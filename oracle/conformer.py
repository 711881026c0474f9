"""Conformer-CTC encoder oracle (plain torch on CPU, any float dtype).  TEST INFRASTRUCTURE ONLY.

Functional restatement of the reference model, driven by a reference-layout state_dict
(SURVEY.md §A.2):
  model/conformer.py:8-26   SwiGLUFeedForward      -> swiglu_ff
  model/conformer.py:28-49  TransposeGroupNorm     -> group_norm_tokens
  model/conformer.py:51-88  ConformerConvModule    -> conv_module
  model/conformer.py:90-135 ConformerBlock         -> block
  model/conformer.py:137-211 TurkishASRModel       -> subsample, key_mask, forward
  model/attention.py:21-70  RotaryEmbedding/rotate_half/apply_rotary_pos_emb -> rope_tables, apply_rope
  model/attention.py:121-140 _standard_attention   -> attention core (masked_fill -1e9)
  model/attention.py:195-251 RelativeMultiHeadAttention.forward (MQA) -> mqa_attention
Dropout is not modelled (parity runs use dropout=0.0, SURVEY.md §0 finding 10).
"""
import math

import torch
import torch.nn.functional as F


def num_groups(num_channels, requested=32):
    """model/conformer.py:33-43: 32 groups, falling back to the first divisor in [32,16,8,4,2], else 1."""
    if num_channels % requested == 0:
        return requested
    for g in (32, 16, 8, 4, 2):
        if num_channels % g == 0:
            return g
    return 1


def group_norm_tokens(x, weight, bias, groups=None, eps=1e-5):
    """x (B, T, C): GroupNorm over (C/groups channels x all T) per sample (model/conformer.py:45-49)."""
    g = num_groups(x.shape[-1]) if groups is None else groups
    return F.group_norm(x.transpose(1, 2), g, weight, bias, eps).transpose(1, 2)


def swiglu_ff(x, sd, prefix):
    """model/conformer.py:18-26 (dropout omitted)."""
    y = F.linear(x, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"])
    gate, value = y.chunk(2, dim=-1)
    y = F.silu(gate) * value
    return F.linear(y, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])


def rope_tables(seq_len, dim=64, base=10000.0, dtype=torch.float32):
    """model/attention.py:36-49: inv_freq, emb = cat(freqs, freqs); returns cos, sin (T, dim)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
    t = torch.arange(seq_len, dtype=inv_freq.dtype)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def apply_rope(x, cos, sin):
    """model/attention.py:62-70. x (..., T, dim)."""
    x1, x2 = x.chunk(2, dim=-1)
    return x * cos + torch.cat((-x2, x1), dim=-1) * sin


def mqa_core(q, k, v, key_lengths=None):
    """q (B,H,T,64), k, v (B,1,T,64) -> (B,H,T,64).  model/attention.py:121-140 with the key mask of
    model/conformer.py:187-202 (mask == 0 -> -1e9)."""
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(q.shape[-1])
    if key_lengths is not None:
        t = q.shape[2]
        mask = torch.arange(t)[None, :] < key_lengths[:, None]
        scores = scores.masked_fill(~mask[:, None, None, :], -1e9)
    return torch.matmul(F.softmax(scores, dim=-1), v)


def mqa_attention(x, sd, prefix, n_heads, key_lengths=None):
    """model/attention.py:195-251 with use_mqa=True."""
    b, t, d = x.shape
    dh = d // n_heads
    q = F.linear(x, sd[prefix + "linear_q.weight"], sd[prefix + "linear_q.bias"]).view(b, t, n_heads, dh).transpose(1, 2)
    k = F.linear(x, sd[prefix + "linear_k.weight"], sd[prefix + "linear_k.bias"]).view(b, t, 1, dh).transpose(1, 2)
    v = F.linear(x, sd[prefix + "linear_v.weight"], sd[prefix + "linear_v.bias"]).view(b, t, 1, dh).transpose(1, 2)
    cos, sin = rope_tables(t, dh, dtype=x.dtype)
    q = apply_rope(q, cos, sin)
    k = apply_rope(k, cos, sin)
    ctx = mqa_core(q, k, v, key_lengths)
    ctx = ctx.transpose(1, 2).contiguous().view(b, t, d)
    return F.linear(ctx, sd[prefix + "linear_out.weight"], sd[prefix + "linear_out.bias"])


def conv_module(x, sd, prefix, training=True, bn_state=None):
    """model/conformer.py:76-88.  BatchNorm in training mode normalises with biased batch statistics over
    (B*T) incl. padded frames; bn_state (dict) receives the running-stat update (momentum 0.1)."""
    d = x.shape[-1]
    y = group_norm_tokens(x, sd[prefix + "norm.norm.weight"], sd[prefix + "norm.norm.bias"])
    y = y.transpose(1, 2)
    y = F.conv1d(y, sd[prefix + "pointwise_conv1.weight"], sd[prefix + "pointwise_conv1.bias"])
    y = F.glu(y, dim=1)
    y = F.conv1d(y, sd[prefix + "depthwise_conv.weight"], sd[prefix + "depthwise_conv.bias"], padding=15, groups=d)
    rm = sd[prefix + "batch_norm.running_mean"].clone()
    rv = sd[prefix + "batch_norm.running_var"].clone()
    y = F.batch_norm(y, rm, rv, sd[prefix + "batch_norm.weight"], sd[prefix + "batch_norm.bias"], training=training,
                     momentum=0.1, eps=1e-5)
    if bn_state is not None:
        bn_state[prefix + "batch_norm.running_mean"] = rm
        bn_state[prefix + "batch_norm.running_var"] = rv
    y = F.silu(y)
    y = F.conv1d(y, sd[prefix + "pointwise_conv2.weight"], sd[prefix + "pointwise_conv2.bias"])
    return y.transpose(1, 2)


def block(x, sd, prefix, n_heads, key_lengths=None, training=True, bn_state=None):
    """model/conformer.py:114-135."""
    def gn(name, t):
        return group_norm_tokens(t, sd[prefix + name + ".norm.weight"], sd[prefix + name + ".norm.bias"])
    x = x + 0.5 * swiglu_ff(gn("norm_ff1", x), sd, prefix + "ff1.")
    x = x + mqa_attention(gn("norm_attn", x), sd, prefix + "attn.", n_heads, key_lengths)
    x = x + conv_module(x, sd, prefix + "conv.", training, bn_state)
    x = x + 0.5 * swiglu_ff(gn("norm_ff2", x), sd, prefix + "ff2.")
    return gn("final_norm", x)


def subsample(feats, sd):
    """model/conformer.py:177-185: two stride-2 3x3 convs + SiLU, (B,C,T',F')->(B,T',C*F'), input_proj."""
    x = feats.unsqueeze(1)
    x = F.silu(F.conv2d(x, sd["subsample.0.weight"], sd["subsample.0.bias"], stride=2, padding=1))
    x = F.silu(F.conv2d(x, sd["subsample.2.weight"], sd["subsample.2.bias"], stride=2, padding=1))
    b, c, t, f = x.shape
    x = x.permute(0, 2, 1, 3).contiguous().view(b, t, c * f)
    return F.linear(x, sd["input_proj.weight"], sd["input_proj.bias"])


def encoder_frames(t_mel):
    """T' after two k3/s2/p1 convolutions (integer parity item, SURVEY.md §0 finding 5)."""
    t1 = (t_mel - 1) // 2 + 1
    return (t1 - 1) // 2 + 1


def forward(feats, input_lengths, sd, n_heads, n_blocks, training=True, bn_state=None, return_hidden=False):
    """model/conformer.py:172-211.  feats (B,T,F); input_lengths (B,) int64 or None -> logits (B,T',V)."""
    x = subsample(feats, sd)
    key_lengths = None if input_lengths is None else (input_lengths // 4)
    hidden = [x]
    for i in range(n_blocks):
        x = block(x, sd, "blocks.%d." % i, n_heads, key_lengths, training, bn_state)
        hidden.append(x)
    logits = F.linear(x, sd["fc.weight"], sd["fc.bias"])
    return (logits, hidden) if return_hidden else logits


def ctc_loss_torch(logits, targets, input_lengths, target_lengths, blank=0):
    """trainer/trainer.py:167-173: permute, log_softmax (fp32 or better), nn.CTCLoss(blank, zero_infinity=True),
    mean reduction; input lengths are mel lengths // 4."""
    log_probs = F.log_softmax(logits.permute(1, 0, 2), dim=2)
    return F.ctc_loss(log_probs, targets, input_lengths // 4, target_lengths, blank=blank, reduction="mean",
                      zero_infinity=True)


def greedy_ids(logits, lengths=None, blank=0):
    """utils/decoding.py:149,163 + data/tokenizer.py:44-54: argmax, collapse repeats, drop blank."""
    ids = torch.argmax(logits, dim=-1)
    out = []
    for b in range(ids.shape[0]):
        seq = ids[b].tolist()
        if lengths is not None:
            seq = seq[: int(lengths[b])]
        toks, last = [], None
        for c in seq:
            if c != last and c != blank:
                toks.append(c)
            last = c
        out.append(toks)
    return ids, out


def adamw_step(p, g, m, v, step, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6):
    """torch.optim.AdamW single-tensor update (main.py:106-110 hyper-parameters)."""
    p = p * (1 - lr * weight_decay)
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g * g
    bc1 = 1 - betas[0] ** step
    bc2 = 1 - betas[1] ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


def clip_coef(grads, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_ (trainer/trainer.py:190): global L2 norm, coef = min(1, max/(norm+1e-6))."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads))
    return total, min(1.0, max_norm / (float(total) + 1e-6))
