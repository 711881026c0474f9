"""tcgen05 GEMM (C-ABI tasr_gemm_bf16) against a float64 CPU contraction of the same bf16 inputs."""
import pytest
import torch

from turkish_asr_model_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _ref(A, B, a_mn, b_mn):
    a = A.double().cpu()
    b = B.double().cpu()
    if a_mn:
        a = a.t()
    if b_mn:
        b = b.t()
    return a @ b.t()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 256), (304, 200, 136), (2008, 1000, 256)])
def test_gemm_store(cuda, M, N, K, a_mn, b_mn):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g).to(torch.bfloat16).to(cuda)
    B = torch.randn((K, N) if b_mn else (N, K), generator=g).to(torch.bfloat16).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    out = torch.full((M, N), float("nan"), device=cuda, dtype=torch.float32)
    L.gemm(M, N, K, A, A.stride(0), B, B.stride(0), L.EPI_STORE, out, N, a_mn=a_mn, b_mn=b_mn, out_f32=1,
           bias=bias)
    torch.cuda.synchronize()
    ref = _ref(A, B, a_mn, b_mn) + bias.double().cpu()
    err = (out.double().cpu() - ref).abs().max().item()
    assert err < 2e-3 * (K ** 0.5), f"max err {err}"


def test_gemm_swiglu_and_bwd(cuda):
    M, d, dff = 384, 256, 1024
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(M, d, generator=g)).to(torch.bfloat16).to(cuda)
    W = (torch.randn(2 * dff, d, generator=g) / 16).to(torch.bfloat16).to(cuda)
    b = torch.randn(2 * dff, generator=g).to(cuda)
    gv = torch.empty(M, 2 * dff, device=cuda, dtype=torch.bfloat16)
    h = torch.empty(M, dff, device=cuda, dtype=torch.bfloat16)
    L.gemm(M, dff, d, x, d, W, d, L.EPI_SWIGLU, h, dff, out2=gv, ldo2=2 * dff, bias=b, n_half=dff)
    torch.cuda.synchronize()
    ref = x.double().cpu() @ W.double().cpu().t() + b.double().cpu()
    assert (gv.double().cpu() - ref).abs().max().item() < 0.05
    gq, vq = gv.double().cpu()[:, :dff], gv.double().cpu()[:, dff:]
    href = torch.nn.functional.silu(gq) * vq
    assert (h.double().cpu() - href).abs().max().item() < 0.03 * href.abs().max().item()
    # backward epilogue: dh = dy @ W2 with W2 (d, dff) used MN-major; out = dg|dv
    dy = torch.randn(M, d, generator=g).to(torch.bfloat16).to(cuda)
    W2 = (torch.randn(d, dff, generator=g) / 16).to(torch.bfloat16).to(cuda)
    dgv = torch.empty(M, 2 * dff, device=cuda, dtype=torch.bfloat16)
    L.gemm(M, dff, d, dy, d, W2, dff, L.EPI_SWIGLU_BWD, dgv, 2 * dff, b_mn=1, aux=gv, ldaux=2 * dff, n_half=dff)
    torch.cuda.synchronize()
    dh = dy.double().cpu() @ W2.double().cpu()
    s = torch.sigmoid(gq)
    dg = dh * vq * (s * (1 + gq * (1 - s)))
    dv = dh * gq * s
    ref2 = torch.cat([dg, dv], 1)
    assert (dgv.double().cpu() - ref2).abs().max().item() < 0.03 * ref2.abs().max().item()


def test_gemm_wgrad_splitk_remap(cuda):
    Mtok, N, K = 5000, 256, 512
    g = torch.Generator().manual_seed(9)
    dy = torch.randn(Mtok, N, generator=g).to(torch.bfloat16).to(cuda)
    x = torch.randn(Mtok, K, generator=g).to(torch.bfloat16).to(cuda)
    dW = torch.zeros(N, K, device=cuda)
    L.gemm(N, K, Mtok, dy, N, x, K, L.EPI_ATOMIC, dW, K, a_mn=1, b_mn=1, split_k=8)
    torch.cuda.synchronize()
    ref = dy.double().cpu().t() @ x.double().cpu()
    assert (dW.double().cpu() - ref).abs().max().item() < 0.02 * ref.abs().max().item()
    # remapped columns: n -> (n % p0) * p1 + n / p0
    dW2 = torch.zeros(N, K, device=cuda)
    p0, p1 = 128, 4
    L.gemm(N, K, Mtok, dy, N, x, K, L.EPI_ATOMIC, dW2, K, a_mn=1, b_mn=1, split_k=3, remap_p0=p0, remap_p1=p1)
    torch.cuda.synchronize()
    idx = torch.arange(K)
    dst = (idx % p0) * p1 + idx // p0
    ref2 = torch.zeros_like(ref)
    ref2[:, dst] = ref
    assert (dW2.double().cpu() - ref2).abs().max().item() < 0.02 * ref.abs().max().item()


@pytest.mark.parametrize("N,K,Mtok,split", [(256, 1024, 5000, 4), (1000, 256, 2008, 7), (384, 256, 777, 1), (2048, 256, 3000, 9)])
def test_gemm_wgrad_with_fused_bias_gradient(cuda, N, K, Mtok, split):
    """ATOMIC epilogue with `colsum`: the bias gradient sum_tokens dy[t, o] comes out of the same pass over dy (an extra
    N = 16 UMMA against a tile of ones), for wide / narrow tiles, ragged out-feature counts and any split count."""
    g = torch.Generator().manual_seed(N + K + split)
    dy = torch.randn(Mtok, N, generator=g).to(torch.bfloat16).to(cuda)
    x = torch.randn(Mtok, K, generator=g).to(torch.bfloat16).to(cuda)
    dW = torch.zeros(N, K, device=cuda)
    db = torch.full((N,), 0.5, device=cuda)  # accumulates (+=)
    L.gemm(N, K, Mtok, dy, N, x, K, L.EPI_ATOMIC, dW, K, a_mn=1, b_mn=1, split_k=split, colsum=db)
    torch.cuda.synchronize()
    ref = dy.double().cpu().t() @ x.double().cpu()
    assert (dW.double().cpu() - ref).abs().max().item() < 0.02 * ref.abs().max().item()
    ref_b = 0.5 + dy.double().cpu().sum(0)
    assert (db.double().cpu() - ref_b).abs().max().item() < 2e-3 * ref_b.abs().max().item()


def test_gemm_resid(cuda):
    M, N, K = 777, 256, 1024
    g = torch.Generator().manual_seed(11)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(cuda)
    B = (torch.randn(N, K, generator=g) / 32).to(torch.bfloat16).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    res = torch.randn(M, N, generator=g).to(cuda)
    out = torch.empty_like(res)
    L.gemm(M, N, K, A, K, B, K, L.EPI_RESID, out, N, bias=bias, aux=res, ldaux=N, alpha=0.5)
    torch.cuda.synchronize()
    ref = res.double().cpu() + 0.5 * (A.double().cpu() @ B.double().cpu().t() + bias.double().cpu())
    assert (out.double().cpu() - ref).abs().max().item() < 1e-3


def test_dropout_mask_gemm_epilogues_match_cast_kernel(cuda):
    """One dropout site = one mask over the flat element index: the forward GEMM epilogues (per-run fast hash), the
    stand-alone / GroupNorm-backward cast (generic hash) and the SwiGLU backward epilogue must all see the same mask for
    the same seed (forward RESID dropout is undone in backward by the cast; SwiGLU dropout by the SWIGLU_BWD epilogue)."""
    M, N, K, p, seed = 300, 256, 64, 0.25, 777
    g = torch.Generator().manual_seed(9)
    A = torch.zeros(M, K, dtype=torch.bfloat16, device=cuda)
    B = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).to(cuda)
    ones = torch.ones(N, device=cuda)
    res = torch.zeros(M, N, device=cuda)
    out = torch.empty(M, N, device=cuda)
    L.gemm(M, N, K, A, K, B, K, L.EPI_RESID, out, N, bias=ones, aux=res, ldaux=N, alpha=1.0, drop_p=p, seed=seed)
    cast = L.cast_bf16(torch.ones(M, N, device=cuda), alpha=1.0, drop_p=p, seed=seed)
    torch.cuda.synchronize()
    mask_gemm = out != 0
    mask_cast = cast.float() != 0
    assert torch.equal(mask_gemm, mask_cast)
    keep = mask_gemm.float().mean().item()
    assert abs(keep - (1 - p)) < 0.01
    assert torch.allclose(out[mask_gemm], torch.full_like(out[mask_gemm], 1.0 / (1 - p)), rtol=1e-6)
    # SwiGLU forward (dropout on the activation) vs SwiGLU backward epilogue (same mask applied to the incoming gradient)
    dff = 128
    W1 = torch.zeros(2 * dff, K, dtype=torch.bfloat16, device=cuda)
    b1 = torch.cat([torch.full((dff,), 3.0), torch.full((dff,), 2.0)]).to(cuda)     # gate = 3, up = 2 everywhere
    h = torch.empty(M, dff, dtype=torch.bfloat16, device=cuda)
    gv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=cuda)
    L.gemm(M, dff, K, A, K, W1, K, L.EPI_SWIGLU, h, dff, out2=gv, ldo2=2 * dff, bias=b1, n_half=dff, drop_p=p, seed=seed + 1)
    dy = torch.zeros(M, K, dtype=torch.bfloat16, device=cuda)
    dy[:, 0] = 1.0
    W2 = torch.zeros(K, dff, dtype=torch.bfloat16, device=cuda)
    W2[0, :] = 1.0                                                                   # d act = 1 everywhere
    dgv = torch.empty(M, 2 * dff, dtype=torch.bfloat16, device=cuda)
    L.gemm(M, dff, K, dy, K, W2, dff, L.EPI_SWIGLU_BWD, dgv, 2 * dff, b_mn=1, aux=gv, ldaux=2 * dff, n_half=dff,
           drop_p=p, seed=seed + 1)
    torch.cuda.synchronize()
    assert torch.equal(h.float() != 0, dgv[:, dff:].float() != 0)
    assert abs((h.float() != 0).float().mean().item() - (1 - p)) < 0.01


@pytest.mark.parametrize("d,H,B,T", [(256, 4, 3, 77), (512, 8, 2, 203)])
def test_gemm_rope_epilogue(cuda, d, H, B, T):
    """Fused q|k|v projection with the rotary embedding applied to the q and k heads in the epilogue, against
    oracle.apply_rope (model/attention.py:62-70) on the fp64 projection; the v columns are stored unrotated."""
    from oracle import conformer as oc
    g = torch.Generator().manual_seed(d + T)
    M, N = B * T, d + 128
    x = torch.randn(M, d, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, d, generator=g) / d ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    cos, sin = oc.rope_tables(T, 64)
    cs = torch.stack([cos[:, :32], sin[:, :32]], dim=-1).contiguous().to(cuda)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda)
    L.gemm(M, N, d, x.to(cuda), d, w.to(cuda), d, L.EPI_ROPE, out, N, bias=bias.to(cuda), aux=cs, n_half=T, remap_p0=d + 64)
    torch.cuda.synchronize()
    y = x.double() @ w.double().t() + bias.double()
    q = y[:, :d].reshape(B, T, H, 64).transpose(1, 2)
    k = y[:, d:d + 64].reshape(B, T, 1, 64).transpose(1, 2)
    qr = oc.apply_rope(q, cos.double(), sin.double()).transpose(1, 2).reshape(M, d)
    kr = oc.apply_rope(k, cos.double(), sin.double()).transpose(1, 2).reshape(M, 64)
    ref = torch.cat([qr, kr, y[:, d + 64:]], 1)
    err = (out.double().cpu() - ref).abs().max().item()
    assert err < 1.2e-2 * ref.abs().max().item(), err  # bf16 output rounding
