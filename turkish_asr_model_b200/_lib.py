"""ctypes binding of libtasr_kernels.so (the C-ABI declared in include/tasr_kernels.h).

There is no CPU fallback: if the shared library is missing or the device is not sm_100 every op
raises.  PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtasr_kernels.so")
_lib = None



P, I, L, F, Z, U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint64
# every exported symbol of include/tasr_kernels.h (tests check the library exports all of them)
_SIGNATURES = {
    "tasr_version": (I, []),
    "tasr_check_device": (I, []),
    "tasr_init": (I, []),
    "tasr_launch_count": (U64, []),
    "tasr_set_dropout_seed_ptr": (I, [P]),
    "tasr_get_dropout_seed_ptr": (P, []),
    "tasr_gemm_bf16": (I, [P, P]),
    "tasr_mel_filter_ranges": (I, [P, I, P, P]),
    "tasr_mel_workspace_bytes": (Z, [I, I]),
    "tasr_mel_forward": (I, [P, L, P, I, I, P, P, P, I, I, I, I, P, P, Z, P]),
    "tasr_groupnorm_workspace_bytes": (Z, [I, I, I]),
    "tasr_groupnorm_fwd": (I, [P, I, I, I, I, F, P, P, P, I, P, P, Z, P]),
    "tasr_groupnorm_bwd": (I, [P, I, P, I, I, I, I, P, P, P, I, P, P, P, F, F, U64, P, Z, P]),
    "tasr_dwconv_bn_parts": (I, [I, I]),
    "tasr_dwconv31_fwd": (I, [P, I, I, I, P, P, P, P, P]),
    "tasr_dwconv31_bwd": (I, [P, P, P, I, I, I, P, P, P, P, P, P]),
    "tasr_bn_finalize": (I, [P, I, I, L, F, F, I, P, P, P, P, P]),
    "tasr_bn_silu_fwd": (I, [P, L, I, P, P, P, P, P]),
    "tasr_bn_bwd_workspace_bytes": (Z, [L, I]),
    "tasr_bn_silu_bwd": (I, [P, P, L, I, P, P, P, P, P, P, P, Z, P]),
    "tasr_cast_f32_bf16": (I, [P, P, L, F, F, U64, P]),
    "tasr_colsum_bf16": (I, [P, L, I, L, P, P]),
    "tasr_rope_inplace": (I, [P, L, I, I, I, P, I, P]),
    "tasr_mqa_attention_fwd": (I, [P, I, I, I, I, P, F, U64, P, P, P]),
    "tasr_mqa_attention_bwd_workspace_bytes": (Z, [I, I, I, I]),
    "tasr_mqa_attention_bwd": (I, [P, P, P, P, I, I, I, I, P, F, U64, P, P, P, Z, P]),
    "tasr_pack_weight_remap": (I, [P, L, I, I, P, P]),
    "tasr_conv1_fwd": (I, [P, I, I, I, I, P, P, P, P]),
    "tasr_conv1_bwd": (I, [P, P, I, I, I, I, P, P, P, P, P]),
    "tasr_conv2_fwd": (I, [P, I, I, I, I, P, P, P, P, P]),
    "tasr_conv2_dgrad": (I, [P, I, I, I, I, P, P, P]),
    "tasr_conv2_wgrad": (I, [P, P, I, I, I, I, P, P]),
    "tasr_resample_sinc": (I, [P, L, P, P, P, I, P, L, I, P]),
    "tasr_specaugment": (I, [P, I, I, I, P, I, P, P]),
    "tasr_ctc_workspace_bytes": (Z, [I, I, I, I]),
    "tasr_ctc_loss_fwd_bwd": (I, [P, I, L, I, I, I, P, I, P, P, I, F, P, P, P, P, Z, P]),
    "tasr_grad_sumsq": (I, [P, L, P, P]),
    "tasr_clip_adamw": (I, [P, P, P, P, P, L, P, P, P, P]),
    "tasr_argmax_collapse": (I, [P, I, L, I, I, I, P, I, P, P, P, P]),
    "tasr_f32_split_terms": (I, [P, L, I, L, I, I, C.c_uint32, I, P, P]),
    "tasr_f32_conv1": (I, [P, I, I, I, I, P, P, P, P]),
    "tasr_f32_im2col_split": (I, [P, I, I, I, I, C.c_uint32, I, P, P]),
    "tasr_f32_rope": (I, [P, L, I, I, I, P, P]),
    "tasr_f32_mqa_fwd": (I, [P, I, I, I, I, P, P, P]),
    "tasr_f32_dwconv_parts": (I, [I, I]),
    "tasr_f32_dwconv31": (I, [P, I, I, I, P, P, P, P, P]),
    "tasr_f32_bn_silu": (I, [P, L, I, P, P, P, P, P]),
    "tasr_f32_glu": (I, [P, L, I, P, P]),
}


class TasrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TasrError(
                "libtasr_kernels.so is not built (run `python -m turkish_asr_model_b200.build`); "
                "there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.tasr_status_string.restype = C.c_char_p
        _lib.tasr_last_error.restype = C.c_char_p
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if torch.cuda.is_available():
            rc = _lib.tasr_init()
            if rc != 0:
                raise TasrError("tasr_init failed: %s" % _lib.tasr_status_string(rc).decode())
    return _lib


def check(rc):
    if rc != 0:
        l = lib()
        raise TasrError("tasr kernel call failed: %s (%s)" % (
            l.tasr_status_string(rc).decode(), l.tasr_last_error().decode()))


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


_ws_cache = {}
_ws_retired = []  # replaced scratch buffers: captured CUDA graphs may still hold their addresses


def workspace(nbytes, device):
    """Reusable scratch buffer (per device); kernels of one stream run in order so sharing is safe.
    A buffer that has to grow is replaced by one of at least twice the size; the old one is kept allocated for the
    life of the process because CUDA graphs captured earlier have its raw address baked in (geometric growth bounds
    the retained memory by the size of the live buffer)."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise TasrError("scratch workspace would have to grow during CUDA-graph capture; run one eager step first")
        size = max(int(nbytes), 128 << 20)
        if buf is not None:
            size = max(size, 2 * buf.numel())
            _ws_retired.append(buf)
        buf = torch.empty(size, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise TasrError("tasr kernels need CUDA tensors (no CPU fallback)")


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
GEMM_PROFILE = None
(EPI_STORE, EPI_RESID, EPI_SWIGLU, EPI_GLU, EPI_SILU, EPI_SWIGLU_BWD, EPI_GLU_BWD, EPI_SILU_BWD, EPI_ATOMIC,
 EPI_ROPE) = range(10)


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("epilogue", C.c_int32), ("out_f32", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("bias", C.c_void_p),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("alpha", C.c_float),
        ("n_half", C.c_int32),
        ("drop_p", C.c_float),
        ("seed", C.c_uint64),
        ("split_k", C.c_int32),
        ("remap_p0", C.c_int32), ("remap_p1", C.c_int32),
        ("colsum", C.c_void_p),
    ]


def gemm(M, N, K, A, lda, B, ldb, epilogue, out, ldo, a_mn=0, b_mn=0, out_f32=0, out2=None, ldo2=0,
         bias=None, aux=None, ldaux=0, alpha=1.0, n_half=0, drop_p=0.0, seed=0, split_k=1,
         remap_p0=0, remap_p1=0, colsum=None):
    require_cuda(A, B, out)
    a = GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.a_mn_major, a.b_mn_major = a_mn, b_mn
    a.A, a.lda, a.B, a.ldb = A.data_ptr(), lda, B.data_ptr(), ldb
    a.epilogue, a.out_f32 = epilogue, out_f32
    a.out, a.ldo = out.data_ptr(), ldo
    a.out2, a.ldo2 = (out2.data_ptr() if out2 is not None else 0), ldo2
    a.bias = bias.data_ptr() if bias is not None else 0
    a.aux, a.ldaux = (aux.data_ptr() if aux is not None else 0), ldaux
    a.alpha, a.n_half, a.drop_p, a.seed = alpha, n_half, drop_p, seed
    a.split_k, a.remap_p0, a.remap_p1 = split_k, remap_p0, remap_p1
    a.colsum = colsum.data_ptr() if colsum is not None else 0
    fn = lib().tasr_gemm_bf16
    if GEMM_PROFILE is not None:  # bench.py roofline pass: remember every tcgen05 GEMM launch of a step
        nb = 2 if epilogue in (EPI_SWIGLU, EPI_GLU) else 1
        GEMM_PROFILE.append((2.0 * M * N * nb * K, a, (A, B, out, out2, bias, aux), (M, N * nb, K, epilogue, a_mn, b_mn)))
    check(fn(C.addressof(a), stream_ptr()))


# ---------------------------------------------------------------------------------------------
# log-mel front-end
# ---------------------------------------------------------------------------------------------
def mel_filter_ranges(fb):
    require_cuda(fb)
    n_mels = fb.shape[1]
    ranges = torch.empty(2 * n_mels, dtype=torch.int32, device=fb.device)
    check(lib().tasr_mel_filter_ranges(ptr(fb), n_mels, ptr(ranges), stream_ptr()))
    return ranges


def mel_forward(wave, n_samples, tmax, window, fb, ranges, normalize=True):
    """wave (B, Nmax) fp32 cuda, n_samples (B,) int32 cuda -> feats (B, tmax, n_mels) fp32."""
    require_cuda(wave, n_samples, window, fb, ranges)
    B = wave.shape[0]
    n_mels = fb.shape[1]
    feats = torch.empty(B, tmax, n_mels, dtype=torch.float32, device=wave.device)
    wsb = lib().tasr_mel_workspace_bytes(B, n_mels)
    ws = torch.empty(wsb, dtype=torch.uint8, device=wave.device)
    check(lib().tasr_mel_forward(ptr(wave), wave.stride(0), ptr(n_samples), B, tmax, ptr(window), ptr(fb),
                                 ptr(ranges), n_mels, 400, 160, int(bool(normalize)), ptr(feats), ptr(ws),
                                 wsb, stream_ptr()))
    return feats


# ---------------------------------------------------------------------------------------------
# GroupNorm
# ---------------------------------------------------------------------------------------------
def groupnorm_fwd(x, G, gamma, beta, eps=1e-5, out_bf16=True):
    """x (B,T,d) fp32 -> (y, stats (B,G,2))."""
    require_cuda(x, gamma, beta)
    B, T, d = x.shape
    y = torch.empty(B, T, d, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x.device)
    stats = torch.empty(B, G, 2, dtype=torch.float32, device=x.device)
    wsb = lib().tasr_groupnorm_workspace_bytes(B, T, d)
    ws = workspace(wsb, x.device)
    check(lib().tasr_groupnorm_fwd(ptr(x), B, T, d, G, eps, ptr(gamma), ptr(beta), ptr(y), int(out_bf16), ptr(stats),
                                   ptr(ws), wsb, stream_ptr()))
    return y, stats


def groupnorm_bwd(dy, x, G, stats, gamma, dres, accumulate, dgamma, dbeta, cast=None):
    """dres (B,T,d) fp32 (+)= dx; dgamma/dbeta += .  cast = (alpha, drop_p, seed) additionally returns
    bf16(alpha * dropout_mask * dres) written in the same pass."""
    require_cuda(dy, x, dres)
    B, T, d = x.shape
    wsb = lib().tasr_groupnorm_workspace_bytes(B, T, d)
    ws = workspace(wsb, x.device)
    cast_out = None
    alpha, drop_p, seed = 1.0, 0.0, 0
    if cast is not None:
        alpha, drop_p, seed = cast
        cast_out = torch.empty(B, T, d, dtype=torch.bfloat16, device=x.device)
    check(lib().tasr_groupnorm_bwd(ptr(dy), int(dy.dtype == torch.bfloat16), ptr(x), B, T, d, G, ptr(stats), ptr(gamma),
                                   ptr(dres), int(accumulate), ptr(dgamma), ptr(dbeta), ptr(cast_out), alpha, drop_p, seed,
                                   ptr(ws), wsb, stream_ptr()))
    return cast_out


# ---------------------------------------------------------------------------------------------
# depthwise conv + BatchNorm/SiLU
# ---------------------------------------------------------------------------------------------
def dwconv_fwd(u, weight, bias, want_stats=True):
    """u (B,T,d) bf16; weight (d,31) fp32 -> (w (B,T,d) bf16, bn_partial)."""
    require_cuda(u, weight, bias)
    B, T, d = u.shape
    out = torch.empty_like(u)
    part = None
    if want_stats:
        part = torch.empty(lib().tasr_dwconv_bn_parts(B, T), d, 2, dtype=torch.float32, device=u.device)
    check(lib().tasr_dwconv31_fwd(ptr(u), B, T, d, ptr(weight), ptr(bias), ptr(out), ptr(part), stream_ptr()))
    return out, part


def dwconv_bwd(dw, u, ab, weight, dweight, dbias):
    require_cuda(dw, u, weight)
    B, T, d = u.shape
    dab = torch.empty(B, T, 2 * d, dtype=torch.bfloat16, device=u.device) if ab is not None else None
    du = torch.empty_like(u) if ab is None else None
    check(lib().tasr_dwconv31_bwd(ptr(dw), ptr(u), ptr(ab), B, T, d, ptr(weight), ptr(dab), ptr(du), ptr(dweight),
                                  ptr(dbias), stream_ptr()))
    return dab if ab is not None else du


def dwconv_bwd_data(dw, u, ab, weight):
    """Data half of dwconv_bwd: returns d(a|b) (GLU backward fused) or du when ab is None."""
    B, T, d = u.shape
    dab = torch.empty(B, T, 2 * d, dtype=torch.bfloat16, device=u.device) if ab is not None else None
    du = torch.empty_like(u) if ab is None else None
    check(lib().tasr_dwconv31_bwd(ptr(dw), ptr(u), ptr(ab), B, T, d, ptr(weight), ptr(dab), ptr(du), None, None, stream_ptr()))
    return dab if ab is not None else du


def dwconv_bwd_weight(dw, u, weight, dweight, dbias):
    """Weight / bias-gradient half of dwconv_bwd (accumulates): a leaf of the backward graph."""
    B, T, d = u.shape
    check(lib().tasr_dwconv31_bwd(ptr(dw), ptr(u), None, B, T, d, ptr(weight), None, None, ptr(dweight), ptr(dbias), stream_ptr()))


def bn_finalize(part, d, count, eps, momentum, training, running_mean, running_var, num_batches_tracked):
    stats = torch.empty(d, 2, dtype=torch.float32, device=running_mean.device)
    npart = part.shape[0] if part is not None else 0
    check(lib().tasr_bn_finalize(ptr(part), npart, d, count, eps, momentum, int(training), ptr(running_mean),
                                 ptr(running_var), ptr(num_batches_tracked), ptr(stats), stream_ptr()))
    return stats


def bn_silu_fwd(w, stats, gamma, beta):
    M, d = w.numel() // w.shape[-1], w.shape[-1]
    out = torch.empty_like(w)
    check(lib().tasr_bn_silu_fwd(ptr(w), M, d, ptr(stats), ptr(gamma), ptr(beta), ptr(out), stream_ptr()))
    return out


def bn_silu_bwd(ds, w, stats, gamma, beta, dgamma, dbeta):
    M, d = w.numel() // w.shape[-1], w.shape[-1]
    dw = torch.empty_like(w)
    wsb = lib().tasr_bn_bwd_workspace_bytes(M, d)
    ws = workspace(wsb, w.device)
    check(lib().tasr_bn_silu_bwd(ptr(ds), ptr(w), M, d, ptr(stats), ptr(gamma), ptr(beta), ptr(dw), ptr(dgamma),
                                 ptr(dbeta), ptr(ws), wsb, stream_ptr()))
    return dw


# ---------------------------------------------------------------------------------------------
# elementwise
# ---------------------------------------------------------------------------------------------
def cast_bf16(x, alpha=1.0, drop_p=0.0, seed=0, out=None):
    require_cuda(x)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib().tasr_cast_f32_bf16(ptr(x), ptr(out), x.numel(), alpha, drop_p, seed, stream_ptr()))
    return out


def colsum_add(x2d, out):
    """out (N) fp32 += column sums of x2d (M,N) bf16."""
    M, N = x2d.shape
    check(lib().tasr_colsum_bf16(ptr(x2d), M, N, x2d.stride(0), ptr(out), stream_ptr()))


def rope_inplace(qkv, T, rot_cols, cos_sin, inverse=False):
    M, ld = qkv.shape
    check(lib().tasr_rope_inplace(ptr(qkv), M, T, ld, rot_cols, ptr(cos_sin), int(inverse), stream_ptr()))


# ---------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------
def _check_key_lengths(key_lengths, B):
    if key_lengths is not None and (key_lengths.dtype != torch.int64 or key_lengths.numel() != B or not key_lengths.is_cuda):
        raise ValueError("key_lengths must be a CUDA int64 tensor of B elements")


def mqa_fwd(qkv, B, T, H, d, key_lengths, drop_p=0.0, seed=0):
    require_cuda(qkv)
    _check_key_lengths(key_lengths, B)
    ctx = torch.empty(B * T, d, dtype=torch.bfloat16, device=qkv.device)
    lse2 = torch.empty(B, H, T, dtype=torch.float32, device=qkv.device)
    check(lib().tasr_mqa_attention_fwd(ptr(qkv), B, T, H, d, ptr(key_lengths), drop_p, seed, ptr(ctx), ptr(lse2),
                                       stream_ptr()))
    return ctx, lse2


def mqa_bwd(qkv, ctx, dctx, lse2, B, T, H, d, key_lengths, cos_sin, drop_p=0.0, seed=0):
    _check_key_lengths(key_lengths, B)
    dqkv = torch.empty_like(qkv)
    wsb = lib().tasr_mqa_attention_bwd_workspace_bytes(B, T, H, d)
    ws = workspace(wsb, qkv.device)
    check(lib().tasr_mqa_attention_bwd(ptr(qkv), ptr(ctx), ptr(dctx), ptr(lse2), B, T, H, d, ptr(key_lengths), drop_p,
                                       seed, ptr(cos_sin), ptr(dqkv), ptr(ws), wsb, stream_ptr()))
    return dqkv


# ---------------------------------------------------------------------------------------------
# subsampler
# ---------------------------------------------------------------------------------------------
def sub_dims(T, F):
    T1, F1 = (T - 1) // 2 + 1, (F - 1) // 2 + 1
    return T1, F1, (T1 - 1) // 2 + 1, (F1 - 1) // 2 + 1


def pack_weight_remap(w2d, q):
    N, K = w2d.shape
    out = torch.empty(N, K, dtype=torch.bfloat16, device=w2d.device)
    check(lib().tasr_pack_weight_remap(ptr(w2d), N, K, q, ptr(out), stream_ptr()))
    return out


def conv1_fwd(x, w1, b1):
    """x (B,T,F) fp32 -> y1 (B,T1,F1,d) bf16 NHWC."""
    require_cuda(x, w1, b1)
    B, T, F = x.shape
    d = w1.shape[0]
    T1, F1, _, _ = sub_dims(T, F)
    y1 = torch.empty(B, T1, F1, d, dtype=torch.bfloat16, device=x.device)
    check(lib().tasr_conv1_fwd(ptr(x), B, T, F, d, ptr(w1), ptr(b1), ptr(y1), stream_ptr()))
    return y1


def conv1_bwd(dy1, x, w1, b1, dw1, db1):
    B, T, F = x.shape
    check(lib().tasr_conv1_bwd(ptr(dy1), ptr(x), B, T, F, w1.shape[0], ptr(w1), ptr(b1), ptr(dw1), ptr(db1), stream_ptr()))


def conv2_fwd(y1, T, F, w2p, bias):
    """y1 (B,T1,F1,d) -> (z2, y2) each (B*T2*F2, d) bf16."""
    B, _, _, d = y1.shape
    _, _, T2, F2 = sub_dims(T, F)
    z2 = torch.empty(B * T2 * F2, d, dtype=torch.bfloat16, device=y1.device)
    y2 = torch.empty_like(z2)
    check(lib().tasr_conv2_fwd(ptr(y1), B, T, F, d, ptr(w2p), ptr(bias), ptr(z2), ptr(y2), stream_ptr()))
    return z2, y2


def conv2_dgrad(dz2, B, T, F, w2p):
    d = dz2.shape[-1]
    T1, F1, _, _ = sub_dims(T, F)
    dy1 = torch.empty(B, T1, F1, d, dtype=torch.bfloat16, device=dz2.device)
    check(lib().tasr_conv2_dgrad(ptr(dz2), B, T, F, d, ptr(w2p), ptr(dy1), stream_ptr()))
    return dy1


def conv2_wgrad(dz2, y1, T, F, dw2):
    B, _, _, d = y1.shape
    check(lib().tasr_conv2_wgrad(ptr(dz2), ptr(y1), B, T, F, d, ptr(dw2), stream_ptr()))


# ---------------------------------------------------------------------------------------------
# CTC, optimizer, decode
# ---------------------------------------------------------------------------------------------
def _rows_view(logits):
    """(B,T,V) tensor whose rows are dense with a common pitch ld >= V -> (tensor, ld); copies otherwise."""
    B, T, V = logits.shape
    if logits.stride(2) == 1 and logits.stride(0) == T * logits.stride(1) and logits.stride(1) >= V:
        return logits, logits.stride(1)
    logits = logits.contiguous()
    return logits, V


def ctc_loss_fwd_bwd(logits, targets, input_lengths, target_lengths, blank=0, grad_scale=1.0, want_grad=True,
                     grad_dtype=None):
    """logits (B,T,V) bf16|fp32; targets (B,Smax) int64; lengths int64 (device) -> (loss (1), nll (B), dlogits).
    grad_dtype=torch.float32 with bf16 logits returns the gradient unrounded (default: the logits' dtype)."""
    require_cuda(logits, targets, input_lengths, target_lengths)
    B, T, V = logits.shape
    logits, ld = _rows_view(logits)
    Smax = targets.shape[1]
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    nll = torch.empty(B, dtype=torch.float32, device=logits.device)
    dlogits = None
    mode = int(logits.dtype == torch.bfloat16)
    if want_grad:  # same row pitch as the logits; padding columns (if any) are zero
        gdt = logits.dtype if grad_dtype is None else grad_dtype
        if gdt != logits.dtype:
            if not (logits.dtype == torch.bfloat16 and gdt == torch.float32):
                raise TasrError("grad_dtype: only an fp32 gradient for bf16 logits differs from the logits dtype")
            mode = 2
        dl = torch.zeros(B * T, ld, dtype=gdt, device=logits.device) if ld != V else \
            torch.empty(B * T, ld, dtype=gdt, device=logits.device)
        dlogits = dl.view(B, T, ld)[:, :, :V]
    wsb = lib().tasr_ctc_workspace_bytes(B, T, V, Smax)
    ws = workspace(wsb, logits.device)
    check(lib().tasr_ctc_loss_fwd_bwd(ptr(logits), mode, ld, B, T, V, ptr(targets), Smax,
                                      ptr(input_lengths), ptr(target_lengths), blank, grad_scale, ptr(loss), ptr(nll),
                                      ptr(dlogits), ptr(ws), wsb, stream_ptr()))
    return loss, nll, dlogits


def grad_sumsq(g, out):
    check(lib().tasr_grad_sumsq(ptr(g), g.numel(), ptr(out), stream_ptr()))


def clip_adamw(p, g, m, v, shadow, hyper, sumsq, norm_out):
    check(lib().tasr_clip_adamw(ptr(p), ptr(g), ptr(m), ptr(v), ptr(shadow), p.numel(), ptr(hyper), ptr(sumsq),
                                ptr(norm_out), stream_ptr()))


def argmax_collapse(logits, lengths=None, blank=0):
    require_cuda(logits)
    B, T, V = logits.shape
    logits, ld = _rows_view(logits)
    ids = torch.empty(B, T, dtype=torch.int64, device=logits.device)
    tokens = torch.empty(B, T, dtype=torch.int64, device=logits.device)
    out_len = torch.empty(B, dtype=torch.int32, device=logits.device)
    check(lib().tasr_argmax_collapse(ptr(logits), int(logits.dtype == torch.bfloat16), ld, B, T, V, ptr(lengths), blank,
                                     ptr(ids), ptr(tokens), ptr(out_len), stream_ptr()))
    return ids, tokens, out_len


def gemm_replay(profile):
    """Re-issue the GEMM launches recorded in GEMM_PROFILE (same argument structs; operands kept alive)."""
    fn = lib().tasr_gemm_bf16
    sp = stream_ptr()
    for _, a, _, _ in profile:
        check(fn(C.addressof(a), sp))


# ---------------------------------------------------------------------------------------------
# augmentation
# ---------------------------------------------------------------------------------------------
def resample_sinc(waves, n_in, orig, new, max_out):
    """waves (B, Nmax) fp32; n_in/orig/new (B,) int32 (frequencies divided by their gcd) -> (B, max_out) fp32."""
    require_cuda(waves, n_in, orig, new)
    B = waves.shape[0]
    y = torch.empty(B, max_out, dtype=torch.float32, device=waves.device)
    check(lib().tasr_resample_sinc(ptr(waves), waves.stride(0), ptr(n_in), ptr(orig), ptr(new), B, ptr(y), y.stride(0),
                                   max_out, stream_ptr()))
    return y


def specaugment_(feats, params, frames=None):
    """feats (B,T,F) fp32 in place; params (B, nmask, 3) int32 cuda."""
    require_cuda(feats, params)
    B, T, F = feats.shape
    check(lib().tasr_specaugment(ptr(feats), B, T, F, ptr(params), params.shape[1], ptr(frames), stream_ptr()))
    return feats


# ---------------------------------------------------------------------------------------------
# fp32 operand mode (csrc/fp32_mode.cu): bf16 piece expansion around the same tcgen05 GEMM
# ---------------------------------------------------------------------------------------------
# piece indices per product term, A side / B side: sum_t A_piece[t] * B_piece[t]
F32_TERMS = {3: ((0, 0, 1), (0, 1, 0)), 6: ((0, 0, 1, 0, 1, 2), (0, 1, 0, 2, 1, 0))}
ACT_NONE, ACT_SILU, ACT_SWIGLU, ACT_GLU = range(4)


def _pack_terms(pieces):
    v = 0
    for t, p in enumerate(pieces):
        v |= p << (4 * t)
    return v


def f32_split(x2d, K, side, nterms=3, act=ACT_NONE, remap_q=0):
    """x2d (M, K or 2K) fp32 (row pitch = stride(0)) -> (M, nterms*K) bf16 operand of the expanded GEMM.
    side 0 = A (activations), 1 = B (weights)."""
    require_cuda(x2d)
    M = x2d.shape[0]
    out = torch.empty(M, nterms * K, dtype=torch.bfloat16, device=x2d.device)
    check(lib().tasr_f32_split_terms(ptr(x2d), M, K, x2d.stride(0), act, remap_q, _pack_terms(F32_TERMS[nterms][side]), nterms,
                                     ptr(out), stream_ptr()))
    return out


def f32_conv1(x, w1, b1):
    B, T, F = x.shape
    d = w1.shape[0]
    T1, F1, _, _ = sub_dims(T, F)
    y1 = torch.empty(B, T1, F1, d, dtype=torch.float32, device=x.device)
    check(lib().tasr_f32_conv1(ptr(x), B, T, F, d, ptr(w1), ptr(b1), ptr(y1), stream_ptr()))
    return y1


def f32_im2col_split(y1, T, F, nterms=3):
    B, _, _, d = y1.shape
    _, _, T2, F2 = sub_dims(T, F)
    out = torch.empty(B * T2 * F2, nterms * 9 * d, dtype=torch.bfloat16, device=y1.device)
    check(lib().tasr_f32_im2col_split(ptr(y1), B, T, F, d, _pack_terms(F32_TERMS[nterms][0]), nterms, ptr(out), stream_ptr()))
    return out


def f32_rope_(qkv, T, rot_cols, cos_sin):
    check(lib().tasr_f32_rope(ptr(qkv), qkv.shape[0], T, qkv.stride(0), rot_cols, ptr(cos_sin), stream_ptr()))


def f32_mqa_fwd(qkv, B, T, H, d, key_lengths):
    _check_key_lengths(key_lengths, B)
    ctx = torch.empty(B * T, d, dtype=torch.float32, device=qkv.device)
    check(lib().tasr_f32_mqa_fwd(ptr(qkv), B, T, H, d, ptr(key_lengths), ptr(ctx), stream_ptr()))
    return ctx


def f32_dwconv(u, weight, bias, want_stats):
    B, T, d = u.shape
    out = torch.empty_like(u)
    part = torch.empty(lib().tasr_f32_dwconv_parts(B, T), d, 2, dtype=torch.float32, device=u.device) if want_stats else None
    check(lib().tasr_f32_dwconv31(ptr(u), B, T, d, ptr(weight), ptr(bias), ptr(out), ptr(part), stream_ptr()))
    return out, part


def f32_bn_silu(w, stats, gamma, beta):
    M, d = w.numel() // w.shape[-1], w.shape[-1]
    out = torch.empty_like(w)
    check(lib().tasr_f32_bn_silu(ptr(w), M, d, ptr(stats), ptr(gamma), ptr(beta), ptr(out), stream_ptr()))
    return out


def f32_glu(ab, d):
    M = ab.shape[0]
    if ab.stride(0) != 2 * d:
        raise TasrError("f32_glu needs a dense (M, 2d) matrix")
    u = torch.empty(M, d, dtype=torch.float32, device=ab.device)
    check(lib().tasr_f32_glu(ptr(ab), M, d, ptr(u), stream_ptr()))
    return u
