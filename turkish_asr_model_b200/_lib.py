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
        _lib.tasr_mel_workspace_bytes.restype = C.c_size_t
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
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise TasrError("tasr kernels need CUDA tensors (no CPU fallback)")


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
EPI_STORE, EPI_RESID, EPI_SWIGLU, EPI_GLU, EPI_SILU, EPI_SWIGLU_BWD, EPI_GLU_BWD, EPI_SILU_BWD, EPI_ATOMIC = range(9)


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
    ]


def gemm(M, N, K, A, lda, B, ldb, epilogue, out, ldo, a_mn=0, b_mn=0, out_f32=0, out2=None, ldo2=0,
         bias=None, aux=None, ldaux=0, alpha=1.0, n_half=0, drop_p=0.0, seed=0, split_k=1,
         remap_p0=0, remap_p1=0, debug=False):
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
    fn = lib().tasr_gemm_bf16_debug if debug else lib().tasr_gemm_bf16
    check(fn(C.byref(a), stream_ptr()))


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
    check(lib().tasr_mel_forward(ptr(wave), C.c_int64(wave.stride(0)), ptr(n_samples), B, tmax, ptr(window), ptr(fb),
                                 ptr(ranges), n_mels, 400, 160, int(bool(normalize)), ptr(feats), ptr(ws),
                                 C.c_size_t(wsb), stream_ptr()))
    return feats
