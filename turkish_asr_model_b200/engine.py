"""Forward / backward of the Conformer-CTC encoder on the libtasr_kernels C-ABI (host-side orchestration).

This is the host logic underneath the drop-in modules of model/conformer.py: it owns the flat fp32
parameter / gradient / bf16-shadow buffers and strings the kernels together (no autograd tape: every
backward formula is explicit, so the whole step is a fixed kernel sequence that can be replayed or
graph-captured).  Mirrors reference model/conformer.py:114-135 (block), :172-211 (model forward) and the
autograd backward PyTorch derives from them; citations per function below.

Layout decisions: activations token-major (M = B*T', d); residual stream fp32 (the
reference's autocast flow after the first GroupNorm); GEMM operands bf16; gradients of
weights accumulated in fp32 straight into the flat gradient buffer by the split-K wgrad epilogue.
"""
import math

import os

import torch

from . import _lib as L

DH = 64


def _align(n, a=8):
    return (n + a - 1) // a * a


class FlatParams:
    """All parameters of a TurkishASRModel re-homed into one flat fp32 buffer (+ grads, AdamW moments and a
    bf16 shadow used as GEMM operands).  q/k/v projection weights (and biases) are laid out contiguously so
    that the fused QKV GEMM and its wgrad see one (d+128, d) matrix.  Parameters that never receive a
    gradient in the reference (blocks.*.norm_conv.*, SURVEY.md §0 finding 6) sit in a tail segment that the
    optimizer and the gradient all-reduce skip."""

    def __init__(self, model):
        self.model = model
        named = dict(model.named_parameters())
        order, dead = [], []
        used = set()

        def take(name):
            order.append(name)
            used.add(name)

        for name in ("subsample.0.weight", "subsample.0.bias", "subsample.2.weight", "subsample.2.bias",
                     "input_proj.weight", "input_proj.bias"):
            take(name)
        for i in range(len(model.blocks)):
            p = "blocks.%d." % i
            for name in ("norm_ff1.norm.weight", "norm_ff1.norm.bias", "ff1.linear1.weight", "ff1.linear1.bias",
                         "ff1.linear2.weight", "ff1.linear2.bias", "norm_attn.norm.weight", "norm_attn.norm.bias",
                         "attn.linear_q.weight", "attn.linear_k.weight", "attn.linear_v.weight",
                         "attn.linear_q.bias", "attn.linear_k.bias", "attn.linear_v.bias",
                         "attn.linear_out.weight", "attn.linear_out.bias", "conv.norm.norm.weight", "conv.norm.norm.bias",
                         "conv.pointwise_conv1.weight", "conv.pointwise_conv1.bias", "conv.depthwise_conv.weight",
                         "conv.depthwise_conv.bias", "conv.batch_norm.weight", "conv.batch_norm.bias",
                         "conv.pointwise_conv2.weight", "conv.pointwise_conv2.bias", "norm_ff2.norm.weight",
                         "norm_ff2.norm.bias", "ff2.linear1.weight", "ff2.linear1.bias", "ff2.linear2.weight",
                         "ff2.linear2.bias", "final_norm.norm.weight", "final_norm.norm.bias"):
                take(p + name)
            dead += [p + "norm_conv.norm.weight", p + "norm_conv.norm.bias"]
        take("fc.weight")
        take("fc.bias")
        for name in dead:
            used.add(name)
        missing = [n for n in named if n not in used]
        if missing:
            raise L.TasrError("unexpected parameters: %s" % missing)

        self.offsets = {}
        off = 0
        for name in order:
            self.offsets[name] = off
            off += _align(named[name].numel())
        self.live_numel = off
        for name in dead:
            self.offsets[name] = off
            off += _align(named[name].numel())
        self.total = off
        dev = next(model.parameters()).device
        self.device = dev
        self.params = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(self.total, dtype=torch.bfloat16, device=dev)
        self.exp_avg = None
        self.exp_avg_sq = None
        self.names = order + dead
        self.shapes = {n: tuple(named[n].shape) for n in self.names}
        self._ptrs = {}
        with torch.no_grad():
            for name in self.names:
                p = named[name]
                view = self.params[self.offsets[name]: self.offsets[name] + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self._ptrs[name] = view.data_ptr()
        self.shadow_fresh = False
        self._plist = [named[n] for n in self.names]
        self._cast_version = -1

    def param_version(self):
        """Sum of the autograd version counters of every parameter: any in-place edit made through PyTorch
        (torch.optim step on the autograd path, load_state_dict, p.data.copy_ ...) bumps it.  The fused AdamW
        kernel writes params and shadow together through raw pointers and therefore leaves it untouched."""
        return sum(p._version for p in self._plist)

    def still_valid(self):
        p = self.model.fc.weight
        return p.data_ptr() == self._ptrs["fc.weight"] and p.device == self.device

    def view(self, buf, name, shape=None, numel=None):
        p_shape = self.shapes[name] if shape is None else shape
        n = int(math.prod(p_shape)) if numel is None else numel
        return buf[self.offsets[name]: self.offsets[name] + n].view(p_shape)

    def refresh_shadow(self):
        """Re-cast the bf16 GEMM operands when the fp32 parameters were changed behind the engine's back."""
        v = self.param_version()
        if not self.shadow_fresh or v != self._cast_version:
            if torch.cuda.is_current_stream_capturing():
                raise L.TasrError("parameters changed during CUDA-graph capture")
            L.cast_bf16(self.params, out=self.shadow)
            self.shadow_fresh = True
            self._cast_version = v

    def grad_views(self):
        """name -> view into the flat gradient buffer (dead parameters excluded)."""
        out = {}
        for name, p in self.model.named_parameters():
            if self.offsets[name] < self.live_numel:
                out[name] = self.grads[self.offsets[name]: self.offsets[name] + p.numel()].view(p.shape)
        return out


class _W:
    """Per-forward bundle of operand views for one model (fp32 params, bf16 shadows, fp32 grads)."""


def _split_k(out_f, in_f, tokens):
    """Split the token reduction of a wgrad so that one wave of ~148 CTAs covers it: more splits only multiply the
    fp32 reduce-add traffic into the (small) weight-gradient tile."""
    tiles = ((out_f + 127) // 128) * ((in_f + 255) // 256)
    kb = (tokens + 63) // 64
    return max(1, min(kb, 148 // max(tiles, 1)))  # 74 / 111 / 296 CTAs measured: slower or within noise


class ConformerEngine:
    def __init__(self, model):
        self.model = model
        self.flat = None
        self.d = model.d_model
        self.H = model.n_heads
        if self.d != self.H * DH:
            raise L.TasrError("the B200 attention kernel needs d_model == 64 * n_heads (got %d, %d)" % (self.d, self.H))
        self.dff = 4 * self.d
        self.G = model.blocks[0].norm_ff1.norm.num_groups if len(model.blocks) else 32
        self.n_blocks = len(model.blocks)
        self.V = model.fc.out_features
        self.F2 = model.input_proj.in_features // self.d
        self._cos_sin = None
        self.device_seed = False  # True: dropout seeds come from the device counter (CUDA-graph replay)
        # weight / bias gradient kernels are leaves of the backward graph: they run on a second (lower priority) stream
        # and fill the SMs that the tail of each main-chain kernel leaves idle
        self.overlap_wgrad = os.environ.get("TASR_NO_WGRAD_OVERLAP", "0") != "1"
        self._side = None
        self._side_busy = False
        self._keep = []
        self._lagging = None  # (event, tensors) of the previous block's leaf kernels, joined one block late
        # "bf16" (default: bf16 operands, the training path) or "fp32" (forward only: fp32 activations, contractions
        # evaluated from bf16 piece expansions of the fp32 operands, logits within 1e-4 of the reference's fp32 run)
        self.precision = "bf16"
        self.f32_terms = 3  # 3 = a0b0+a0b1+a1b0 (~2^-17 per product); 6 adds the 2^-24 terms

    # ------------------------------------------------------------------ parameters
    def ensure_flat(self):
        if self.flat is None or not self.flat.still_valid():
            self.flat = FlatParams(self.model)
            self._build_views()
        return self.flat

    def _build_views(self):
        f = self.flat
        d, dff = self.d, self.dff

        cache = {}

        def cached(buf_id, buf, name, shape, numel):
            key = (buf_id, name, shape, numel)
            v = cache.get(key)
            if v is None:
                v = cache[key] = f.view(buf, name, shape, numel)
            return v

        def P(name, shape=None, numel=None):
            return cached(0, f.params, name, shape, numel)

        def S(name, shape=None, numel=None):
            return cached(1, f.shadow, name, shape, numel)

        def G(name, shape=None, numel=None):
            return cached(2, f.grads, name, shape, numel)

        self.P, self.S, self.Gv = P, S, G
        self.qkv_w_shape = (d + 2 * DH, d)

    def workspace_bytes(self, B, T, smax):
        """Largest scratch request any kernel of one training step makes for a batch of B utterances of T mel frames
        and targets padded to smax (the shared scratch buffer must have this size BEFORE a CUDA-graph capture)."""
        lib = L.lib()
        _, _, T2, _ = L.sub_dims(T, 80)
        M = B * T2
        return max(lib.tasr_groupnorm_workspace_bytes(B, T2, self.d), lib.tasr_bn_bwd_workspace_bytes(M, self.d),
                   lib.tasr_mqa_attention_bwd_workspace_bytes(B, T2, self.H, self.d),
                   lib.tasr_ctc_workspace_bytes(B, T2, self.V, max(int(smax), 1)), 1)

    def cos_sin(self, T, device):
        if self._cos_sin is None or self._cos_sin.shape[0] < T or self._cos_sin.device != device:
            n = max(T, 2048)
            inv_freq = self.model.blocks[0].attn.rotary_emb.inv_freq.to(device=device, dtype=torch.float32)
            t = torch.arange(n, device=device, dtype=torch.float32)
            freqs = torch.outer(t, inv_freq)  # model/attention.py:42-44
            self._cos_sin = torch.stack([freqs.cos(), freqs.sin()], dim=-1).contiguous()
        return self._cos_sin

    # ------------------------------------------------------------------ forward
    def forward(self, feats, input_lengths, training, dropout_p, save):
        """feats (B, T, F) fp32 cuda; input_lengths (B,) int64 (any device) or None.
        Returns logits (B, T', V) bf16 and (if save) the tape needed by backward()."""
        model = self.model
        f = self.ensure_flat()
        f.refresh_shadow()
        P, S = self.P, self.S
        d, dff, H, G = self.d, self.dff, self.H, self.G
        B, T, F = feats.shape
        feats = feats.contiguous().float()
        _, _, T2, F2 = L.sub_dims(T, F)
        if F2 != self.F2:
            raise L.TasrError("n_mel_channels mismatch: input_proj expects %d frequency bins after subsampling" % self.F2)
        M = B * T2
        dev = feats.device
        key_len = None
        if input_lengths is not None:
            key_len = (input_lengths.to(device=dev, dtype=torch.int64) // 4).contiguous()  # model/conformer.py:191
        cs = self.cos_sin(T2, dev)
        drop = float(dropout_p) if training else 0.0
        seed0 = 0
        if drop > 0.0 and not self.device_seed:
            seed0 = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) * 4096
        tape = {"B": B, "T": T, "F": F, "T2": T2, "M": M, "key_len": key_len, "drop": drop, "seed0": seed0,
                "feats": feats, "blocks": []} if save else None

        # ---- subsampler (model/conformer.py:177-185)
        w2p = L.pack_weight_remap(P("subsample.2.weight").view(d, 9 * d), 9)
        winp = L.pack_weight_remap(P("input_proj.weight"), F2)
        # conv1 (K = 9) direct to NHWC bf16; conv2 as implicit GEMM on tcgen05 (4-D TMA gathers, SiLU epilogue)
        y1 = L.conv1_fwd(feats, P("subsample.0.weight"), P("subsample.0.bias"))
        z2, y2 = L.conv2_fwd(y1, T, F, w2p, P("subsample.2.bias"))
        x = torch.empty(M, d, dtype=torch.float32, device=dev)
        L.gemm(M, d, F2 * d, y2, F2 * d, winp, F2 * d, L.EPI_STORE, x, d, out_f32=1, bias=P("input_proj.bias"))
        if save:
            tape.update(y1=y1, z2=z2, y2=y2, w2p=w2p, winp=winp)
        else:
            del y1, z2

        for i in range(self.n_blocks):
            x = self._block_forward(i, x, B, T2, key_len, cs, training, drop, seed0 + i * 16, tape)

        xb = L.cast_bf16(x)
        V = self.V
        Vp = _align(V)  # row pitch of logits / dlogits: TMA needs 16-byte multiples
        logits = torch.empty(M, Vp, dtype=torch.bfloat16, device=dev) if Vp == V else \
            torch.zeros(M, Vp, dtype=torch.bfloat16, device=dev)
        L.gemm(M, V, d, xb, d, S("fc.weight"), d, L.EPI_STORE, logits, Vp, bias=P("fc.bias"))
        if save:
            tape["x_final_bf16"] = xb
        return logits.view(B, T2, Vp)[:, :, :V], tape

    def _ff_forward(self, pre, x, xn, M, drop, seed, saved, key):
        d, dff = self.d, self.dff
        P, S = self.P, self.S
        dev = x.device
        gv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=dev)
        h = torch.empty(M, dff, dtype=torch.bfloat16, device=dev)
        L.gemm(M, dff, d, xn, d, S(pre + "linear1.weight"), d, L.EPI_SWIGLU, h, dff, out2=gv, ldo2=2 * dff,
               bias=P(pre + "linear1.bias"), n_half=dff, drop_p=drop, seed=seed)
        out = torch.empty(M, d, dtype=torch.float32, device=dev)
        L.gemm(M, d, dff, h, dff, S(pre + "linear2.weight"), dff, L.EPI_RESID, out, d, bias=P(pre + "linear2.bias"),
               aux=x, ldaux=d, alpha=0.5, drop_p=drop, seed=seed + 1)
        if saved is not None:
            saved[key] = (gv, h)
        return out

    def _block_forward(self, i, x, B, T, key_len, cs, training, drop, seed, tape):
        """model/conformer.py:114-135."""
        d, H, G = self.d, self.H, self.G
        P, S = self.P, self.S
        pre = "blocks.%d." % i
        M = B * T
        dev = x.device
        sv = {} if tape is not None else None
        x3d = x.view(B, T, d)

        # 1. x + 0.5 * ff1(norm_ff1(x))
        xn1, st1 = L.groupnorm_fwd(x3d, G, P(pre + "norm_ff1.norm.weight"), P(pre + "norm_ff1.norm.bias"))
        x1 = self._ff_forward(pre + "ff1.", x, xn1.view(M, d), M, drop, seed + 0, sv, "ff1")
        # 2. x + attn(norm_attn(x) x3)   (the three evaluations are one tensor; model/conformer.py:124)
        xn2, st2 = L.groupnorm_fwd(x1.view(B, T, d), G, P(pre + "norm_attn.norm.weight"), P(pre + "norm_attn.norm.bias"))
        qkv = torch.empty(M, d + 2 * DH, dtype=torch.bfloat16, device=dev)
        wqkv = S(pre + "attn.linear_q.weight", self.qkv_w_shape)
        bqkv = P(pre + "attn.linear_q.bias", (d + 2 * DH,))
        # fused q|k|v projection; RoPE (model/attention.py:228-230) is applied to the q and k heads in the epilogue
        L.gemm(M, d + 2 * DH, d, xn2.view(M, d), d, wqkv, d, L.EPI_ROPE, qkv, d + 2 * DH, bias=bqkv, aux=cs, n_half=T,
               remap_p0=d + DH)
        ctx, lse2 = L.mqa_fwd(qkv, B, T, H, d, key_len, drop_p=drop, seed=seed + 2)
        x2 = torch.empty(M, d, dtype=torch.float32, device=dev)
        L.gemm(M, d, d, ctx, d, S(pre + "attn.linear_out.weight"), d, L.EPI_RESID, x2, d,
               bias=P(pre + "attn.linear_out.bias"), aux=x1, ldaux=d, alpha=1.0)
        # 3. x + conv(x)   (model/conformer.py:76-88)
        xn3, st3 = L.groupnorm_fwd(x2.view(B, T, d), G, P(pre + "conv.norm.norm.weight"), P(pre + "conv.norm.norm.bias"))
        ab = torch.empty(M, 2 * d, dtype=torch.bfloat16, device=dev)
        u = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        L.gemm(M, d, d, xn3.view(M, d), d, S(pre + "conv.pointwise_conv1.weight", (2 * d, d)), d, L.EPI_GLU, u, d,
               out2=ab, ldo2=2 * d, bias=P(pre + "conv.pointwise_conv1.bias"), n_half=d)
        w, part = L.dwconv_fwd(u.view(B, T, d), P(pre + "conv.depthwise_conv.weight", (d, 31)),
                               P(pre + "conv.depthwise_conv.bias"), want_stats=training)
        bn = self.model.blocks[i].conv.batch_norm
        bnst = L.bn_finalize(part, d, M, bn.eps, bn.momentum if bn.momentum is not None else 0.1, training,
                             bn.running_mean, bn.running_var, bn.num_batches_tracked)
        s = L.bn_silu_fwd(w, bnst, P(pre + "conv.batch_norm.weight"), P(pre + "conv.batch_norm.bias"))
        x3 = torch.empty(M, d, dtype=torch.float32, device=dev)
        L.gemm(M, d, d, s.view(M, d), d, S(pre + "conv.pointwise_conv2.weight", (d, d)), d, L.EPI_RESID, x3, d,
               bias=P(pre + "conv.pointwise_conv2.bias"), aux=x2, ldaux=d, alpha=1.0)
        # 4. x + 0.5 * ff2(norm_ff2(x))
        xn4, st4 = L.groupnorm_fwd(x3.view(B, T, d), G, P(pre + "norm_ff2.norm.weight"), P(pre + "norm_ff2.norm.bias"))
        x4 = self._ff_forward(pre + "ff2.", x3, xn4.view(M, d), M, drop, seed + 4, sv, "ff2")
        # 5. final_norm
        x5, st5 = L.groupnorm_fwd(x4.view(B, T, d), G, P(pre + "final_norm.norm.weight"), P(pre + "final_norm.norm.bias"),
                                  out_bf16=False)
        if tape is not None:
            sv.update(x=x, xn1=xn1, st1=st1, x1=x1, xn2=xn2, st2=st2, qkv=qkv, ctx=ctx, lse2=lse2, x2=x2, xn3=xn3,
                      st3=st3, ab=ab, u=u, w=w, bnst=bnst, s=s, x3=x3, xn4=xn4, st4=st4, x4=x4, st5=st5, seed=seed)
            tape["blocks"].append(sv)
        return x5.view(M, d)

    # ------------------------------------------------------------------ fp32 operand mode (forward only)
    def forward_f32(self, feats, input_lengths, training):
        """Same forward as forward(), for callers that want fp32 results (reference without autocast:
        trainer/trainer.py:227-282, inference.py:101-128, the CPU run of BASELINE configs[0]).  Activations stay fp32;
        every Linear / Conv is the tcgen05 GEMM on bf16 piece expansions of its fp32 operands (csrc/fp32_mode.cu).
        Returns logits (B, T', V) fp32.  No tape: backward is not available in this mode."""
        if training and self.model.dropout_p > 0.0:
            raise L.TasrError("fp32 mode models no dropout: use eval mode or dropout=0.0")
        f = self.ensure_flat()
        P = self.P
        d, dff, H, G, V, nt = self.d, self.dff, self.H, self.G, self.V, self.f32_terms
        B, T, F = feats.shape
        feats = feats.contiguous().float()
        T1, F1, T2, F2 = L.sub_dims(T, F)
        if F2 != self.F2:
            raise L.TasrError("n_mel_channels mismatch: input_proj expects %d frequency bins after subsampling" % self.F2)
        M = B * T2
        dev = feats.device
        key_len = None
        if input_lengths is not None:
            key_len = (input_lengths.to(device=dev, dtype=torch.int64) // 4).contiguous()
        cs = self.cos_sin(T2, dev)

        def linear(a_split, w2d, bias, n, k, out=None, resid=None, alpha=1.0, remap_q=0):
            """out (rows, n) fp32 = a @ w^T + bias (+ residual); a_split is the expanded A operand (rows, nt*k)."""
            wb = L.f32_split(w2d, k, 1, nt, remap_q=remap_q)
            rows = a_split.shape[0]
            ldo = (n + 3) // 4 * 4
            if out is None:
                out = torch.empty(rows, ldo, dtype=torch.float32, device=dev)
            if resid is None:
                L.gemm(rows, n, nt * k, a_split, nt * k, wb, nt * k, L.EPI_STORE, out, ldo, out_f32=1, bias=bias)
            else:
                L.gemm(rows, n, nt * k, a_split, nt * k, wb, nt * k, L.EPI_RESID, out, ldo, bias=bias, aux=resid, ldaux=n,
                       alpha=alpha)
            return out

        # ---- subsampler (model/conformer.py:177-185)
        y1 = L.f32_conv1(feats, P("subsample.0.weight"), P("subsample.0.bias"))
        col = L.f32_im2col_split(y1, T, F, nt)
        del y1
        # conv2 weight (co, ci, kh, kw) -> (co, kh, kw, ci): the im2col column order
        z2 = linear(col, P("subsample.2.weight").view(d, 9 * d), P("subsample.2.bias"), d, 9 * d, remap_q=9)
        del col
        a = L.f32_split(z2.view(M, F2 * d), F2 * d, 0, nt, act=L.ACT_SILU)  # SiLU, rows (b,t'), columns (f', c)
        del z2
        x = linear(a, P("input_proj.weight"), P("input_proj.bias"), d, F2 * d, remap_q=F2)
        del a

        for i in range(self.n_blocks):
            pre = "blocks.%d." % i

            def gn(t, name):
                y, _ = L.groupnorm_fwd(t.view(B, T2, d), G, P(pre + name + ".weight"), P(pre + name + ".bias"), out_bf16=False)
                return y.view(M, d)

            def ff(t, name, norm):
                h1 = linear(L.f32_split(gn(t, norm), d, 0, nt), P(pre + name + "linear1.weight"), P(pre + name + "linear1.bias"),
                            2 * dff, d)
                return linear(L.f32_split(h1, dff, 0, nt, act=L.ACT_SWIGLU), P(pre + name + "linear2.weight"),
                              P(pre + name + "linear2.bias"), d, dff, resid=t, alpha=0.5)

            x = ff(x, "ff1.", "norm_ff1.norm")
            # attention (model/attention.py:195-251): fused q|k|v projection, RoPE on q and k, MQA core, output projection
            qkv = linear(L.f32_split(gn(x, "norm_attn.norm"), d, 0, nt), P(pre + "attn.linear_q.weight", self.qkv_w_shape),
                         P(pre + "attn.linear_q.bias", (d + 2 * DH,)), d + 2 * DH, d)
            L.f32_rope_(qkv, T2, d + DH, cs)
            ctx = L.f32_mqa_fwd(qkv, B, T2, H, d, key_len)
            x = linear(L.f32_split(ctx, d, 0, nt), P(pre + "attn.linear_out.weight"), P(pre + "attn.linear_out.bias"), d, d,
                       resid=x, alpha=1.0)
            # conv module (model/conformer.py:76-88)
            ab = linear(L.f32_split(gn(x, "conv.norm.norm"), d, 0, nt), P(pre + "conv.pointwise_conv1.weight", (2 * d, d)),
                        P(pre + "conv.pointwise_conv1.bias"), 2 * d, d)
            u = L.f32_glu(ab, d)  # the depthwise conv needs the fp32 activation itself, not an expanded operand
            w, part = L.f32_dwconv(u.view(B, T2, d), P(pre + "conv.depthwise_conv.weight", (d, 31)),
                                   P(pre + "conv.depthwise_conv.bias"), want_stats=training)
            bn = self.model.blocks[i].conv.batch_norm
            bnst = L.bn_finalize(part, d, M, bn.eps, bn.momentum if bn.momentum is not None else 0.1, training,
                                 bn.running_mean, bn.running_var, bn.num_batches_tracked)
            sact = L.f32_bn_silu(w, bnst, P(pre + "conv.batch_norm.weight"), P(pre + "conv.batch_norm.bias"))
            x = linear(L.f32_split(sact.view(M, d), d, 0, nt), P(pre + "conv.pointwise_conv2.weight", (d, d)),
                       P(pre + "conv.pointwise_conv2.bias"), d, d, resid=x, alpha=1.0)
            x = ff(x, "ff2.", "norm_ff2.norm")
            x = gn(x, "final_norm.norm")

        Vp = (V + 3) // 4 * 4
        logits = torch.zeros(M, Vp, dtype=torch.float32, device=dev) if Vp != V else torch.empty(M, Vp, dtype=torch.float32, device=dev)
        linear(L.f32_split(x, d, 0, nt), P("fc.weight"), P("fc.bias"), V, d, out=logits)
        return logits.view(B, T2, Vp)[:, :, :V]

    # ------------------------------------------------------------------ backward
    def ensure_side_stream(self, device):
        """Create the side stream up front (must exist before a CUDA-graph capture starts)."""
        if self.overlap_wgrad and (self._side is None or self._side.device != device):
            self._side = torch.cuda.Stream(device=device, priority=0)
        return self._side

    def _leaf(self, fn, *tensors):
        """Enqueue fn() (kernels nothing downstream in backward reads: weight / bias gradients) on the side stream,
        ordered after what the current stream has enqueued so far.  `tensors` are kept alive until _join()."""
        if not self.overlap_wgrad:
            fn()
            return
        main = torch.cuda.current_stream()
        self.ensure_side_stream(main.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            fn()
        self._side_busy = True
        self._keep.extend(tensors)

    def _join(self):
        """The current stream waits for the side stream; only then may the tensors its kernels read be released."""
        if self._side_busy:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)
            self._side_busy = False
        self._keep.clear()
        self._lagging = None

    def _join_lagging(self):
        """End of a block's backward: instead of stalling the main chain until this block's last weight-gradient kernel
        has run, wait for the PREVIOUS block's leaf kernels (long finished by now) and release that block's
        temporaries; this block's are handed over to the next call (or to the final _join())."""
        if not self._side_busy:
            return
        ev = torch.cuda.Event()
        ev.record(self._side)
        prev = self._lagging
        self._lagging = (ev, list(self._keep))
        self._keep.clear()
        if prev is not None:
            torch.cuda.current_stream().wait_event(prev[0])

    def _wgrad_bias(self, dy, x, out_f, in_f, tokens, gw, gb, remap=None):
        """gw += dy^T x, gb += column sums of dy: one leaf kernel (side stream)."""
        self._leaf(lambda: self._wgrad(dy, x, out_f, in_f, tokens, gw, remap=remap, gb=gb), dy, x)

    def _wgrad(self, dy, x, out_f, in_f, tokens, gw, remap=None, gb=None):
        """gw (out_f, in_f) fp32 += dy^T x; gb (out_f) fp32 += column sums of dy, from the same pass over dy."""
        p0, p1 = remap if remap is not None else (0, 0)
        L.gemm(out_f, in_f, tokens, dy, dy.stride(0), x, x.stride(0), L.EPI_ATOMIC, gw, in_f, a_mn=1, b_mn=1,
               split_k=_split_k(out_f, in_f, tokens), remap_p0=p0, remap_p1=p1, colsum=gb)

    def _ff_backward(self, pre, dy, saved, xn, M, drop, seed):
        """Backward of x + 0.5*dropout(linear2(dropout(swiglu(linear1(xn))))); dy = bf16(0.5 * mask * dres);
        returns d xn (bf16)."""
        d, dff = self.d, self.dff
        S, Gv = self.S, self.Gv
        gv, h = saved
        self._wgrad_bias(dy, h, d, dff, M, Gv(pre + "linear2.weight"), Gv(pre + "linear2.bias"))
        dgv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=dy.device)
        L.gemm(M, dff, d, dy, d, S(pre + "linear2.weight"), dff, L.EPI_SWIGLU_BWD, dgv, 2 * dff, b_mn=1, aux=gv,
               ldaux=2 * dff, n_half=dff, drop_p=drop, seed=seed)
        self._wgrad_bias(dgv, xn, 2 * dff, d, M, Gv(pre + "linear1.weight"), Gv(pre + "linear1.bias"))
        dxn = torch.empty(M, d, dtype=torch.bfloat16, device=dy.device)
        L.gemm(M, d, 2 * dff, dgv, 2 * dff, S(pre + "linear1.weight"), d, L.EPI_STORE, dxn, d, b_mn=1)
        return dxn

    def _block_backward(self, i, dres, sv, B, T, key_len, cs, drop):
        d, H, G = self.d, self.H, self.G
        P, S, Gv = self.P, self.S, self.Gv
        pre = "blocks.%d." % i
        M = B * T
        dev = dres.device
        seed = sv["seed"]
        d3 = dres.view(B, T, d)

        def gn_bwd(dy, xin, st, name, accumulate, cast=None):
            out = L.groupnorm_bwd(dy.view(B, T, d), xin.view(B, T, d), G, st, P(pre + name + ".weight"), d3, accumulate,
                                  Gv(pre + name + ".weight"), Gv(pre + name + ".bias"), cast=cast)
            return None if out is None else out.view(M, d)

        # Every GroupNorm backward also emits the bf16 operand ("dy") of the backward GEMMs that follow it.
        # 5. final_norm (in place: dres <- d x4)
        dy = gn_bwd(dres, sv["x4"], sv["st5"], "final_norm.norm", False, cast=(0.5, drop, seed + 4 + 1))
        # 4. ff2
        dxn = self._ff_backward(pre + "ff2.", dy, sv["ff2"], sv["xn4"].view(M, d), M, drop, seed + 4)
        dy = gn_bwd(dxn, sv["x3"], sv["st4"], "norm_ff2.norm", True, cast=(1.0, 0.0, 0))
        # 3. conv module
        self._wgrad_bias(dy, sv["s"].view(M, d), d, d, M, Gv(pre + "conv.pointwise_conv2.weight", (d, d)),
                         Gv(pre + "conv.pointwise_conv2.bias"))
        ds = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        L.gemm(M, d, d, dy, d, S(pre + "conv.pointwise_conv2.weight", (d, d)), d, L.EPI_STORE, ds, d, b_mn=1)
        dw = L.bn_silu_bwd(ds, sv["w"], sv["bnst"], P(pre + "conv.batch_norm.weight"), P(pre + "conv.batch_norm.bias"),
                           Gv(pre + "conv.batch_norm.weight"), Gv(pre + "conv.batch_norm.bias"))
        # the 31-tap weight / bias gradient is a leaf: side stream; the data half (with the GLU backward) stays on the chain
        dwv, uv, wdw = dw.view(B, T, d), sv["u"].view(B, T, d), P(pre + "conv.depthwise_conv.weight", (d, 31))
        self._leaf(lambda: L.dwconv_bwd_weight(dwv, uv, wdw, Gv(pre + "conv.depthwise_conv.weight", (d, 31)),
                                               Gv(pre + "conv.depthwise_conv.bias")), dw)
        dab = L.dwconv_bwd_data(dwv, uv, sv["ab"].view(B, T, 2 * d), wdw).view(M, 2 * d)
        self._wgrad_bias(dab, sv["xn3"].view(M, d), 2 * d, d, M, Gv(pre + "conv.pointwise_conv1.weight", (2 * d, d)),
                         Gv(pre + "conv.pointwise_conv1.bias"))
        dxn = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        L.gemm(M, d, 2 * d, dab, 2 * d, S(pre + "conv.pointwise_conv1.weight", (2 * d, d)), d, L.EPI_STORE, dxn, d, b_mn=1)
        dy = gn_bwd(dxn, sv["x2"], sv["st3"], "conv.norm.norm", True, cast=(1.0, 0.0, 0))
        # 2. attention
        self._wgrad_bias(dy, sv["ctx"], d, d, M, Gv(pre + "attn.linear_out.weight"), Gv(pre + "attn.linear_out.bias"))
        dctx = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        L.gemm(M, d, d, dy, d, S(pre + "attn.linear_out.weight"), d, L.EPI_STORE, dctx, d, b_mn=1)
        dqkv = L.mqa_bwd(sv["qkv"], sv["ctx"], dctx, sv["lse2"], B, T, H, d, key_len, cs, drop_p=drop, seed=seed + 2)
        nq = d + 2 * DH
        self._wgrad_bias(dqkv, sv["xn2"].view(M, d), nq, d, M, Gv(pre + "attn.linear_q.weight", self.qkv_w_shape),
                         Gv(pre + "attn.linear_q.bias", (nq,)))
        dxn = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
        L.gemm(M, d, nq, dqkv, nq, S(pre + "attn.linear_q.weight", self.qkv_w_shape), d, L.EPI_STORE, dxn, d, b_mn=1)
        dy = gn_bwd(dxn, sv["x1"], sv["st2"], "norm_attn.norm", True, cast=(0.5, drop, seed + 0 + 1))
        # 1. ff1
        dxn = self._ff_backward(pre + "ff1.", dy, sv["ff1"], sv["xn1"].view(M, d), M, drop, seed + 0)
        # the gradient leaving block 0 feeds input_proj's backward GEMMs: emit its bf16 copy too
        out = gn_bwd(dxn, sv["x"], sv["st1"], "norm_ff1.norm", True, cast=(1.0, 0.0, 0) if i == 0 else None)
        self._join_lagging()  # weight gradients trail the main chain by up to one block (see backward_blocks)
        return out

    def backward(self, tape, dlogits, on_segment_done=None):
        """dlogits (B, T', V) bf16.  Accumulates (+=) every parameter gradient into flat.grads.
        on_segment_done(k) is called after the kernels producing gradient segment k have been enqueued
        (k = 0: classifier, 1..n: blocks n-1..0, n+1: subsampler) so that a data-parallel caller can start
        that segment's all-reduce while the rest of backward runs."""
        dx0 = self.backward_blocks(tape, dlogits, on_segment_done)
        self.backward_head(tape, dx0)
        if on_segment_done is not None:
            on_segment_done(1 + self.n_blocks)

    def backward_blocks(self, tape, dlogits, on_segment_done=None):
        """Classifier + Conformer blocks (95 % of the gradient bytes).  Returns the bf16 gradient w.r.t. the
        input_proj output, the operand of backward_head()."""
        P, S, Gv = self.P, self.S, self.Gv
        d, V = self.d, self.V
        B, T2, M = tape["B"], tape["T2"], tape["M"]
        dev = dlogits.device
        Vp = _align(V)
        if (dlogits.dtype == torch.bfloat16 and dlogits.stride(2) == 1 and dlogits.stride(1) == Vp
                and dlogits.stride(0) == T2 * Vp):
            dl = dlogits.as_strided((M, V), (Vp, 1))
        else:
            buf = torch.zeros(M, Vp, dtype=torch.bfloat16, device=dev)
            buf[:, :V] = dlogits.reshape(M, V)
            dl = buf[:, :V]
        cs = self.cos_sin(T2, dev)
        # classifier (model/conformer.py:209)
        self._wgrad_bias(dl, tape["x_final_bf16"], V, d, M, Gv("fc.weight"), Gv("fc.bias"))
        dres = torch.empty(M, d, dtype=torch.float32, device=dev)
        L.gemm(M, d, V, dl, dl.stride(0), S("fc.weight"), d, L.EPI_STORE, dres, d, b_mn=1, out_f32=1)
        self._join()
        if on_segment_done is not None:
            on_segment_done(0)
        dx0 = None
        for k, i in enumerate(reversed(range(self.n_blocks))):
            dx0 = self._block_backward(i, dres, tape["blocks"][i], B, T2, tape["key_len"], cs, tape["drop"])
            if on_segment_done is not None:
                self._join()  # the caller's all-reduce of this segment must see the finished weight gradients
                on_segment_done(1 + k)
        self._join()
        if dx0 is None:  # no blocks
            dx0 = L.cast_bf16(dres)
        return dx0

    def backward_head(self, tape, dx0):
        """input_proj + Conv2d subsampler (model/conformer.py:177-185)."""
        P, Gv = self.P, self.Gv
        d, F2 = self.d, self.F2
        B, M = tape["B"], tape["M"]
        dev = dx0.device
        y2v = tape["y2"].view(M, F2 * d)
        self._wgrad_bias(dx0, y2v, d, F2 * d, M, Gv("input_proj.weight"), Gv("input_proj.bias"), remap=(d, F2))
        Mpix = M * F2
        dz2 = torch.empty(Mpix, d, dtype=torch.bfloat16, device=dev)
        L.gemm(M, F2 * d, d, dx0, d, tape["winp"], F2 * d, L.EPI_SILU_BWD, dz2, F2 * d, b_mn=1, aux=tape["z2"],
               ldaux=F2 * d)
        def conv2_leaf():
            L.conv2_wgrad(dz2, tape["y1"], tape["T"], tape["F"], Gv("subsample.2.weight"))
            L.colsum_add(dz2, Gv("subsample.2.bias"))
        self._leaf(conv2_leaf, dz2)
        dy1 = L.conv2_dgrad(dz2, B, tape["T"], tape["F"], tape["w2p"])
        L.conv1_bwd(dy1, tape["feats"], P("subsample.0.weight"), P("subsample.0.bias"),
                    Gv("subsample.0.weight"), Gv("subsample.0.bias"))
        self._join()
