"""-m gpu: unit parity of the kernels reworked in round 2, through the C ABI, against plain fp32 torch on the same device.

* gemm_tn: every epilogue at shapes that take the B-resident path (K <= 192, one N tile per CTA for its whole life, also
  with more N tiles than... fewer CTAs than tiles) and the streaming path (K = 576 / 768), with several tiles per CTA;
* gemm_wgrad with the split count the library recommends (bulk-copy epilogue, split fold);
* conv_last forward / backward (row-walking kernels, H % 4 == 0) and the generic fallback (H % 4 != 0), widths that are not a
  multiple of the 64-pixel segment.
Tolerances: bf16 operands with fp32 accumulation against fp32 torch on the bf16-rounded operands -> output rounding only
(rel-L2 <= 4e-3); reductions in fp32 (<= 2e-3)."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2

pytestmark = pytest.mark.gpu
bf = torch.bfloat16


def _capi():
    from superresolution_def_b200 import _capi as capi
    return capi


def _gelu_ref(u):
    cdf = 0.5 * (1 + torch.erf(u / math.sqrt(2)))
    pdf = torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)
    return u * cdf, cdf + u * pdf


@pytest.mark.parametrize("M,N,K", [(128 * 150, 576, 192), (128 * 3, 768, 192), (128 * 160, 192, 768), (128 * 40, 128, 64),
                                   (128, 576, 128), (128 * 150, 192, 576)])
def test_gemm_store_resident_and_streaming(M, N, K):
    capi = _capi()
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").to(bf)
    B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf)
    C = torch.full((M, N), float("nan"), device="cuda", dtype=bf)
    capi.gemm_tn(capi.EPI_STORE, A, B, C)
    ref = A.float() @ B.float().t()
    assert not torch.isnan(C.float()).any()
    assert rel_l2(C, ref) < 4e-3, rel_l2(C, ref)


@pytest.mark.parametrize("M", [128 * 5, 128 * 310])
def test_gemm_gelu2_and_mul(M):
    """fc1 epilogue (gelu and gelu' from the packed-fp16 pair, constant 1.0 column) and the fc2 input gradient epilogue
    (multiplier streamed four boxes ahead), N = 768 = 3 x 256 resp. 4 x 192 tiles, K = 192 (B-resident)."""
    capi = _capi()
    torch.manual_seed(M)
    N, K = 768, 192
    A = torch.randn(M, K, device="cuda").to(bf)
    B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf)
    C = torch.empty(M, N, device="cuda", dtype=bf)
    C2 = torch.empty_like(C, dtype=torch.float16)   # gelu' is stored as fp16
    capi.gemm_tn(capi.EPI_GELU2, A, B, C, C2=C2, ln=capi.make_ln_args(N, 720, None))
    u = A.float() @ B.float().t()
    a, g = _gelu_ref(u)
    a[:, 720], g[:, 720] = 1.0, 0.0
    assert rel_l2(C, a) < 4e-3 and rel_l2(C2, g) < 3e-3, (rel_l2(C, a), rel_l2(C2, g))
    assert (C[:, 720].float() == 1).all() and (C2[:, 720].float() == 0).all()
    X1 = torch.randn(M, N, device="cuda").to(torch.float16)
    D = torch.empty(M, N, device="cuda", dtype=bf)
    capi.gemm_tn(capi.EPI_MUL, A, B, D, X1=X1)
    assert rel_l2(D, u.to(bf).float() * X1.float()) < 4e-3


@pytest.mark.parametrize("K", [192, 576, 768])
def test_gemm_res_ln_and_lnbwd(K):
    """Row epilogues over several tiles per CTA: residual + LayerNorm forward (row statistics), LayerNorm backward with
    the residual gradient read from global memory and the per-CTA dgamma / dbeta partials."""
    capi = _capi()
    torch.manual_seed(K)
    M, N, n = 128 * 170, 192, 180
    A = torch.randn(M, K, device="cuda").to(bf)
    B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bf)
    B[n:] = 0
    R = torch.randn(M, N, device="cuda").to(bf)
    R[:, n:] = 0
    gamma, beta = 1 + 0.3 * torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    stats = torch.zeros(M, 2, device="cuda")
    C, C2 = torch.zeros(M, N, device="cuda", dtype=bf), torch.zeros(M, N, device="cuda", dtype=bf)
    capi.gemm_tn(capi.EPI_RES_LN, A, B, C, C2=C2, X1=R, ln=capi.make_ln_args(n, 180, gamma, beta, stats=stats))
    v = ((A.float() @ B.float().t()).to(bf) + R).float()
    assert rel_l2(C, v) < 4e-3
    vv = C.float()[:, :n]
    mean, var = vv.mean(1, keepdim=True), vv.var(1, unbiased=False, keepdim=True)
    xn = torch.zeros(M, N, device="cuda")
    xn[:, :n] = (vv - mean) / torch.sqrt(var + 1e-5) * gamma + beta
    xn[:, 180] = 1
    assert rel_l2(C2, xn) < 5e-3
    assert rel_l2(stats[:, 0], mean[:, 0]) < 1e-4 and rel_l2(stats[:, 1], 1 / torch.sqrt(var[:, 0] + 1e-5)) < 1e-4
    # backward: acc = dxn, X1 = x (the LN input = C), X2 = residual gradient
    dres = torch.randn(M, N, device="cuda").to(bf)
    dres[:, n:] = 0
    parts = torch.zeros(capi.gemm_grid(M, N), 2, N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda", dtype=bf)
    capi.gemm_tn(capi.EPI_LNBWD, A, B, out, X1=C, X2=dres, ln=capi.make_ln_args(n, -1, gamma, None, stats=stats, partials=parts))
    dxn = (A.float() @ B.float().t()).to(bf).float()[:, :n]
    xhat = (vv - stats[:, 0:1]) * stats[:, 1:2]
    dxh = dxn * gamma
    dx = stats[:, 1:2] * (dxh - dxh.mean(1, keepdim=True) - xhat * (dxh * xhat).mean(1, keepdim=True))
    ref = torch.zeros(M, N, device="cuda")
    ref[:, :n] = dres.float()[:, :n] + dx.to(bf).float()
    assert not torch.isnan(out.float()).any()
    assert rel_l2(out, ref) < 5e-3, rel_l2(out, ref)
    assert (out[:, n:].float() == 0).all()
    p = parts.sum(0)
    assert rel_l2(p[0, :n], (dxn * xhat).sum(0)) < 2e-3 and rel_l2(p[1, :n], dxn.sum(0)) < 2e-3


@pytest.mark.parametrize("T,Ca,Cb", [(64 * 300, 768, 192), (64 * 300, 576, 192), (64 * 77, 192, 192), (64 * 9, 128, 64)])
def test_gemm_wgrad_with_recommended_splits(T, Ca, Cb):
    capi = _capi()
    torch.manual_seed(T + Ca)
    A, B = torch.randn(T, Ca, device="cuda").to(bf), torch.randn(T, Cb, device="cuda").to(bf)
    splits = capi.wgrad_splits(T, Ca)
    assert 1 <= splits <= T // 64
    ws = torch.empty(capi.wgrad_workspace_elems(Ca, Cb, splits), device="cuda")
    rows = (Ca + 127) // 128 * 128
    out = torch.full((rows, Cb), float("nan"), device="cuda")
    capi.gemm_wgrad(A, B, ws, splits, out)
    ref = A.float().t() @ B.float()
    assert rel_l2(out[:Ca], ref) < 2e-3, rel_l2(out[:Ca], ref)
    assert (out[Ca:] == 0).all()


@pytest.mark.parametrize("Bn,H,W", [(2, 32, 64), (1, 12, 40), (2, 10, 24), (1, 64, 200)])
def test_conv_last_forward_and_backward(Bn, H, W):
    """64 -> 1 convolution (conv_last): NHWC bf16 input, fp32 image out; backward = input gradient (bf16 NHWC), weight and
    bias gradients (fp32).  H % 4 == 0 takes the row-walking kernels, H = 10 the generic ones."""
    capi = _capi()
    torch.manual_seed(H * W)
    C = 64
    x = torch.randn(Bn * H * W, C, device="cuda").to(bf)
    w = torch.randn(1, C, 3, 3, device="cuda") / 24
    b = torch.randn(1, device="cuda")
    y = torch.full((Bn, 1, H, W), float("nan"), device="cuda")
    capi.conv_out1_fwd(x, w, b, y, Bn, H, W, C)
    xr = x.float().view(Bn, H, W, C).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr = w.to(bf).float().requires_grad_(True)          # the kernel multiplies bf16-rounded weights, like the autocast conv
    br = b.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, padding=1)
    assert rel_l2(y, ref) < 2e-3, rel_l2(y, ref)
    dy = torch.randn_like(ref)
    ref.backward(dy)
    dx = torch.full_like(x, float("nan"))
    dw, db = torch.empty_like(w), torch.empty_like(b)
    capi.conv_out1_bwd(dy.contiguous(), x, w, dx, dw, db, Bn, H, W, C)
    # the backward kernels use the fp32 weights / the fp32 image gradient (the reference's autocast backward rounds them)
    xr2 = x.float().view(Bn, H, W, C).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    w2 = w.clone().requires_grad_(True)
    F.conv2d(xr2, w2, None, padding=1).backward(dy)
    assert rel_l2(dx.float().view(Bn, H, W, C).permute(0, 3, 1, 2), xr2.grad) < 4e-3
    assert rel_l2(dw, w2.grad) < 2e-3 and rel_l2(db, dy.sum().view(1)) < 2e-3
