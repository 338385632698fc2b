"""-m gpu parity tests of the UNetDiscriminatorSN path against oracle/discriminator_oracle.py.

Layers of evidence:
  * the two data-movement kernels are bit-exact: `srk_disc_patches_k4s2` against `F.unfold` of the same bf16 image (with
    and without the LeakyReLU-backward mask, on channel-slice views), `srk_disc_fold_k4s2` against `F.fold` of the same taps
    summed in fp32 and rounded once (+ add, + both activations);
  * `srk_gemm_tn_lrelu` against fp32 matmul of the bf16 operands (rel-L2 <= 4e-3: one bf16 rounding of the output);
  * the whole network (forward, image gradient, all 12 weight gradients) against the fp32 oracle on the same weights:
    rel-L2 <= 2e-2 on the logits (and within 1.6x + 5e-3 of the autocast error); every gradient within 2x (+5e-3) of the
    error the same ATen sequence makes under bf16 autocast, rel-L2 and max-abs (the yardstick used for the generators' conv tails: LeakyReLU sign flips of near-zero pre-activations
    dominate, and the autocast run has them too);
  * the module mirror (`gan.UNetDiscriminatorSN`, train mode: one power iteration) against the oracle module holding the
    same state_dict: logits, `weight_orig` gradients, and bit-identical `weight_u` buffers after the forward.
"""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2, max_abs

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _mk(shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def _nchw(t2d, B, H, W, c0, C):
    return t2d[:B * H * W, c0:c0 + C].float().reshape(B, H, W, C).permute(0, 3, 1, 2)


@pytest.mark.parametrize("B,H,W,Ct,c0,C", [(2, 8, 8, 64, 0, 64), (1, 16, 32, 128, 64, 64), (3, 4, 4, 1024, 512, 512), (1, 2, 2, 64, 0, 64)])
def test_patches_kernel_is_unfold(B, H, W, Ct, c0, C):
    from superresolution_def_b200 import _capi as capi
    x = _mk((B * H * W, Ct), seed=1).to(BF)
    f = _mk((B * H * W, Ct), seed=2).to(BF)
    M = B * (H // 2) * (W // 2)
    for masked in (False, True):
        p = torch.full((M + 3, 16 * C), 7.0, device="cuda", dtype=BF)
        capi.disc_patches_k4s2(capi.view(x, c0, C), capi.view(f, c0, C) if masked else None, 0.2, B, H, W, p)
        src = _nchw(x, B, H, W, c0, C)
        if masked:
            fm = _nchw(f, B, H, W, c0, C)
            src = torch.where(fm > 0, src, (src * 0.2).to(BF).float())
        # F.unfold: [B, C*16, L] with channel-major (c, ky, kx) rows -> ours is (ky, kx, c)
        u = F.unfold(src, kernel_size=4, stride=2, padding=1).reshape(B, C, 16, -1).permute(0, 3, 2, 1).reshape(M, 16 * C)
        assert torch.equal(p[:M].float(), u.to(BF).float()), (masked, max_abs(p[:M], u))
        assert torch.all(p[M:] == 7.0)   # rows beyond the image are not touched


@pytest.mark.parametrize("B,Hi,Wi,C", [(2, 4, 4, 64), (1, 8, 16, 128), (2, 1, 1, 512)])
def test_fold_kernel_is_fold(B, Hi, Wi, C):
    from superresolution_def_b200 import _capi as capi
    Mi, Ho, Wo = B * Hi * Wi, 2 * Hi, 2 * Wi
    taps = _mk((Mi, 16 * C), seed=3).to(BF)
    add = _mk((B * Ho * Wo, 2 * C), seed=4).to(BF)
    f = _mk((B * Ho * Wo, 2 * C), seed=5).to(BF)
    cols = taps.float().reshape(B, Hi * Wi, 16, C).permute(0, 3, 2, 1).reshape(B, C * 16, Hi * Wi)
    want = F.fold(cols, output_size=(Ho, Wo), kernel_size=4, stride=2, padding=1)          # [B,C,Ho,Wo] fp32 sums
    a_ = _nchw(add, B, Ho, Wo, C, C)
    f_ = _nchw(f, B, Ho, Wo, 0, C)
    for act, use_add in ((capi.FOLD_NONE, False), (capi.FOLD_LRELU, False), (capi.FOLD_MASK, True), (capi.FOLD_NONE, True)):
        y = torch.zeros(B * Ho * Wo, 2 * C, device="cuda", dtype=BF)
        capi.disc_fold_k4s2(taps, B, Hi, Wi, capi.view(y, 0, C), add=capi.view(add, C, C) if use_add else None,
                            f=capi.view(f, 0, C) if act == capi.FOLD_MASK else None, act=act, slope=0.2)
        w = want + (a_ if use_add else 0)
        if act == capi.FOLD_LRELU:
            w = torch.where(w > 0, w, w * 0.2)
        elif act == capi.FOLD_MASK:
            w = torch.where(f_ > 0, w, w * 0.2)
        got = _nchw(y, B, Ho, Wo, 0, C)
        # the summation order of <= 5 fp32 terms may differ from F.fold's: allow one bf16 ulp
        assert rel_l2(got, w.to(BF)) < 2e-3 and max_abs(got, w) <= 2.0 ** -7 * w.abs().max().item(), (act, use_add)
        assert torch.all(y[:, C:] == 0)   # the other half of the level buffer is not touched


@pytest.mark.parametrize("M,N,K,ldc", [(256, 64, 1024, 128), (128, 512, 8192, 512), (384, 128, 64, 256)])
def test_gemm_lrelu_epilogue(M, N, K, ldc):
    from superresolution_def_b200 import _capi as capi
    A = _mk((M, K), seed=6).to(BF)
    Bm = _mk((N, K), K ** -0.5, seed=7).to(BF)
    C = torch.zeros(M, ldc, device="cuda", dtype=BF)
    capi.gemm_tn_lrelu(A, Bm, C[:, ldc - N:], 0.2)
    want = F.leaky_relu(A.float() @ Bm.float().t(), 0.2)
    assert rel_l2(C[:, ldc - N:], want) < 4e-3
    if ldc > N:
        assert torch.all(C[:, :ldc - N] == 0)


@pytest.mark.parametrize("P,Q", [(64, 64), (512, 256), (1024, 256)])
def test_weight_operand_forms_are_bit_exact(P, Q):
    from superresolution_def_b200 import _capi as capi
    w = _mk((P, Q, 4, 4), seed=8)
    a = torch.empty(P, 16 * Q, device="cuda", dtype=BF)
    at = torch.empty(16 * Q, P, device="cuda", dtype=BF)
    capi.disc_prep_w4(w, a, at)
    want = w.permute(0, 2, 3, 1).reshape(P, 16 * Q).to(BF)
    assert torch.equal(a, want) and torch.equal(at, want.t())


@pytest.mark.parametrize("T,R,Cb", [(256, 64, 64), (512, 128, 512), (192, 512, 1024), (2048, 64, 128)])
def test_wgrad4_is_the_unpermuted_matmul(T, R, Cb):
    from superresolution_def_b200 import _capi as capi
    A = _mk((T, 16 * R), seed=9).to(BF)
    Bm = _mk((T, 2 * Cb), seed=10).to(BF)[:, Cb:]          # a channel slice with a row pitch, as the level buffers are
    dw = torch.empty(Cb, R, 4, 4, device="cuda")
    capi.disc_wgrad4(A, Bm, R, dw)
    want = (A.float().t() @ Bm.float()).view(4, 4, R, Cb).permute(3, 2, 0, 1)
    assert rel_l2(dw, want) < 1e-5, rel_l2(dw, want)      # fp32 accumulation of exact bf16 products: summation order only


@pytest.mark.parametrize("shape,dim", [((512, 256, 4, 4), 0), ((1024, 256, 4, 4), 1), ((64, 1, 3, 3), 0), ((1, 64, 3, 3), 0),
                                       ((64, 128, 3, 3), 0), ((256, 128, 4, 4), 1)])
def test_spectral_norm_kernels(shape, dim):
    """srk_spectral_norm / srk_spectral_norm_bwd against the formulas of torch.nn.utils.spectral_norm in fp64:
    v <- normalize(Wm^T u), u <- normalize(Wm v), sigma = u . Wm v, W / sigma; backward with u, v constant."""
    from superresolution_def_b200 import _capi as capi
    w = _mk(shape, 0.05, seed=11)
    U = shape[0] if dim == 0 else shape[1]
    V = w.numel() // U
    u0 = torch.nn.functional.normalize(_mk((U,), seed=12), dim=0)
    v0 = torch.nn.functional.normalize(_mk((V,), seed=13), dim=0)
    wm = (w if dim == 0 else w.permute(1, 0, 2, 3)).reshape(U, V).double()
    for train in (True, False):
        u, v = u0.clone(), v0.clone()
        sig = torch.zeros(1, device="cuda")
        w_sn = torch.empty_like(w)
        layers = capi.sn_layers([w], [u], [v], [dim], sig, [w_sn])
        capi.spectral_norm(layers, train, 1e-12, w.device)
        ur, vr = u0.double(), v0.double()
        if train:
            vr = torch.nn.functional.normalize(wm.t() @ ur, dim=0)
            ur = torch.nn.functional.normalize(wm @ vr, dim=0)
        sr = ur @ (wm @ vr)
        assert rel_l2(u, ur) < 1e-5 and rel_l2(v, vr) < 1e-5, (train, rel_l2(u, ur), rel_l2(v, vr))
        assert abs(sig.item() / sr.item() - 1) < 1e-5
        assert rel_l2(w_sn, w.double() / sr) < 1e-5
        g = _mk(shape, seed=14)
        wd = w.double().requires_grad_(True)
        wmd = (wd if dim == 0 else wd.permute(1, 0, 2, 3)).reshape(U, V)
        ((wd / (ur @ (wmd @ vr))) * g.double()).sum().backward()
        dw = torch.empty_like(w)
        capi.spectral_norm_bwd(layers, [g], [dw], w.device)
        assert rel_l2(dw, wd.grad) < 2e-5, (train, rel_l2(dw, wd.grad))
        g2 = g.clone()
        capi.spectral_norm_bwd(layers, [g2], [g2], w.device)      # in place, as the engine calls it
        assert torch.equal(g2, dw)


def _weights(seed=0):
    nf = 64
    shapes = [(nf, 1, 3, 3), (nf, nf, 4, 4), (2 * nf, nf, 4, 4), (4 * nf, 2 * nf, 4, 4), (8 * nf, 4 * nf, 4, 4), (8 * nf, 8 * nf, 4, 4),
              (8 * nf, 8 * nf, 4, 4), (16 * nf, 4 * nf, 4, 4), (8 * nf, 2 * nf, 4, 4), (4 * nf, nf, 4, 4), (nf, 2 * nf, 3, 3), (1, nf, 3, 3)]
    ws = []
    for i, s in enumerate(shapes):
        transposed = 6 <= i <= 9
        fan_in = (s[0] if transposed else s[1]) * s[2] * s[3] / (4 if transposed else 1)
        ws.append(_mk(s, 1.6 * fan_in ** -0.5, seed=seed + i))
    return ws


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 128, 96), (2, 512, 512)])
def test_discriminator_matches_oracle(B, H, W):
    """forward + image gradient + all weight gradients; (2, 512, 512) is the shape train_swin.py runs (micro-batch 2)."""
    from oracle.discriminator_oracle import unet_discriminator_forward
    from superresolution_def_b200.disc_engine import unet_discriminator
    ws = _weights()
    x = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(99)).cuda()
    dout = _mk((B, 1, H // 2, W // 2), seed=50)

    def run(fn, autocast):
        xs = x.clone().requires_grad_(True)
        wl = [w.clone().requires_grad_(True) for w in ws]
        with torch.autocast("cuda", dtype=BF, enabled=autocast):
            out = fn(xs, wl)
        out.float().backward(dout)
        return out.detach().float(), [xs.grad] + [w.grad for w in wl]

    o_ref, g_ref = run(unet_discriminator_forward, False)
    o_ac, g_ac = run(unet_discriminator_forward, True)
    o_my, g_my = run(unet_discriminator, False)
    assert o_my.shape == o_ref.shape and o_my.dtype == torch.float32
    e_out, e_ac = rel_l2(o_my, o_ref), rel_l2(o_ac, o_ref)
    print(f"logits rel-L2 {e_out:.4f} (autocast oracle {e_ac:.4f}) max-abs {max_abs(o_my, o_ref):.4f} on {o_ref.abs().max().item():.3f}")
    assert e_out < 2e-2 and e_out < 1.6 * e_ac + 5e-3
    assert max_abs(o_my, o_ref) < 1.6 * max_abs(o_ac, o_ref) + 5e-3 * o_ref.abs().max().item()
    names = ["x", "conv0.0", "conv0.2", "conv1", "conv2", "conv3", "conv4", "up1", "up2", "up3", "up4", "final.0", "final.2"]
    bad = []
    for n, a, b, r in zip(names, g_my, g_ac, g_ref):
        assert a is not None and a.shape == r.shape and a.dtype == r.dtype, n
        ea, eb = rel_l2(a, r), rel_l2(b, r)
        ma, mb, top = max_abs(a, r), max_abs(b, r), r.abs().max().item()
        print(f"  grad {n:8s} rel-L2 {ea:.4f} (autocast oracle {eb:.4f})  max-abs {ma:.3e} ({mb:.3e}) on {top:.3e}")
        if not (ea < 2.0 * eb + 5e-3 and ma < 2.0 * mb + 5e-3 * top):
            bad.append((n, ea, eb, ma, mb))
    assert not bad, bad


def test_discriminator_frozen_and_detached_modes():
    """train_swin.py:221-243 calls D three ways: parameters trainable / input detached (D step), parameters frozen / input
    requires grad (G step), and neither (d_real for the G loss): the node returns exactly the gradients asked for."""
    from superresolution_def_b200.disc_engine import unet_discriminator
    ws = _weights(3)
    x = torch.rand(2, 1, 64, 64, device="cuda")
    wl = [w.clone().requires_grad_(True) for w in ws]
    unet_discriminator(x, wl).mean().backward()
    assert all(w.grad is not None for w in wl)
    xs = x.clone().requires_grad_(True)
    out = unet_discriminator(xs, ws)
    out.mean().backward()
    assert xs.grad is not None and xs.grad.shape == x.shape and xs.grad.abs().max() > 0
    assert not unet_discriminator(x, ws).requires_grad
    with pytest.raises(Exception):
        unet_discriminator(torch.rand(1, 1, 48, 64, device="cuda"), ws)   # not a multiple of 32: bilinear branch


def test_module_mirror_matches_oracle_module_train_mode():
    from oracle.discriminator_oracle import UNetDiscriminatorSN as OraD
    from superresolution_def_b200.gan import UNetDiscriminatorSN
    torch.manual_seed(5)
    ora = OraD(1, 64).cuda().train()
    ora_ac = OraD(1, 64).cuda().train()          # the same module under bf16 autocast: the yardstick for the gradients
    mine = UNetDiscriminatorSN(1, 64).cuda().train()
    x = torch.rand(2, 1, 128, 128, device="cuda")
    with torch.no_grad():
        for _ in range(3):      # leave the random initial u / v behind (sigma ~ 1e-3 per layer otherwise)
            ora(x)
    mine.load_state_dict(ora.state_dict(), strict=True)
    ora_ac.load_state_dict(ora.state_dict(), strict=True)
    a, b = mine(x), ora(x)
    with torch.autocast("cuda", dtype=BF):
        c = ora_ac(x)
    for k, v in ora.state_dict().items():
        if k.endswith("weight_u") or k.endswith("weight_v"):   # the same single power iteration, srk_spectral_norm vs the hook
            assert rel_l2(mine.state_dict()[k], v) < 1e-5, (k, rel_l2(mine.state_dict()[k], v))
    assert rel_l2(a, b) < 2e-2 and rel_l2(a, b) < 1.6 * rel_l2(c, b) + 5e-3, (rel_l2(a, b), rel_l2(c, b))
    g = torch.randn_like(b)
    a.backward(g); b.backward(g); c.float().backward(g)
    bad = []
    for (n, p), (_, q), (_, r) in zip(mine.named_parameters(), ora.named_parameters(), ora_ac.named_parameters()):
        assert p.grad is not None and p.grad.shape == q.grad.shape, n
        e_my, e_ac = rel_l2(p.grad, q.grad), rel_l2(r.grad, q.grad)
        if not e_my < 2.0 * e_ac + 5e-3:
            bad.append((n, e_my, e_ac))
    assert not bad, bad
    # eval mode: no power iteration, buffers unchanged
    mine.load_state_dict(ora.state_dict(), strict=True)
    mine.eval(); ora.eval()
    u0 = mine.conv1.model[0].weight_u.clone()
    with torch.no_grad():
        assert rel_l2(mine(x), ora(x)) < 2e-2
    assert torch.equal(mine.conv1.model[0].weight_u, u0)


# ----------------------------------------------------------------------------------------------------------------------
# models/discriminator_hat.py (the discriminator of train_hat.py)
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,Ct,c0,C", [(2, 4, 4, 64, 0, 64), (1, 8, 16, 512, 256, 256), (3, 1, 2, 128, 0, 128)])
def test_bilinear2x_kernels_match_interpolate(B, H, W, Ct, c0, C):
    from superresolution_def_b200 import _capi as capi
    x = _mk((B * H * W, Ct), seed=21).to(BF)
    s = _mk((B * H * W, Ct), seed=22).to(BF)
    for with_s in (False, True):
        y = torch.zeros(B * 4 * H * W, C, device="cuda", dtype=BF)
        capi.bilinear2x_fwd(capi.view(x, c0, C), capi.view(s, c0, C) if with_s else None, capi.view(y), B, H, W)
        src = _nchw(x, B, H, W, c0, C) + (_nchw(s, B, H, W, c0, C) if with_s else 0)
        want = F.interpolate(src, scale_factor=2, mode="bilinear", align_corners=False)
        got = _nchw(y, B, 2 * H, 2 * W, 0, C)
        assert rel_l2(got, want.to(BF)) < 2e-3 and max_abs(got, want) <= 2.0 ** -7 * want.abs().max().item(), with_s
    dy = _mk((B * 4 * H * W, C), seed=23).to(BF)
    dx = torch.zeros(B * H * W, Ct, device="cuda", dtype=BF)
    capi.bilinear2x_bwd(capi.view(dy), capi.view(dx, c0, C), B, H, W)
    z = torch.zeros(B, C, H, W, device="cuda", requires_grad=True)
    F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=False).backward(_nchw(dy, B, 2 * H, 2 * W, 0, C))
    got = _nchw(dx, B, H, W, c0, C)
    assert rel_l2(got, z.grad.to(BF)) < 2e-3 and max_abs(got, z.grad) <= 2.0 ** -7 * z.grad.abs().max().item()
    if c0:
        assert torch.all(dx[:, :c0] == 0)


def _hat_weights(seed=0):
    nf = 64
    shapes = [(nf, 1, 3, 3), (2 * nf, nf, 4, 4), (4 * nf, 2 * nf, 4, 4), (8 * nf, 4 * nf, 4, 4), (4 * nf, 8 * nf, 3, 3),
              (2 * nf, 4 * nf, 3, 3), (nf, 2 * nf, 3, 3), (nf, nf, 3, 3), (nf, nf, 3, 3), (1, nf, 3, 3)]
    ws = [_mk(s, 1.6 * (s[1] * s[2] * s[3]) ** -0.5, seed=seed + i) for i, s in enumerate(shapes)]
    return ws, _mk((nf,), 0.2, seed=seed + 20), _mk((1,), 0.2, seed=seed + 21)


@pytest.mark.parametrize("B,H,W,skip", [(2, 64, 64, True), (1, 32, 128, False), (1, 512, 512, True)])
def test_hat_discriminator_matches_oracle(B, H, W, skip):
    """forward + image gradient + all weight and bias gradients of models/discriminator_hat.py's network; (1, 512, 512) is
    the shape train_hat.py runs.  Same yardstick as the swin variant: the ATen sequence under bf16 autocast."""
    from oracle.discriminator_oracle import unet_discriminator_hat_forward
    from superresolution_def_b200.disc_engine import unet_discriminator_hat
    ws, b0, b9 = _hat_weights()
    x = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(98)).cuda()
    dout = _mk((B, 1, H, W), seed=51)

    def run(fn, autocast):
        xs = x.clone().requires_grad_(True)
        wl = [w.clone().requires_grad_(True) for w in ws]
        bl = [b0.clone().requires_grad_(True), b9.clone().requires_grad_(True)]
        with torch.autocast("cuda", dtype=BF, enabled=autocast):
            out = fn(xs, wl, bl[0], bl[1], skip)
        out.float().backward(dout)
        return out.detach().float(), [xs.grad, bl[0].grad, bl[1].grad] + [w.grad for w in wl]

    o_ref, g_ref = run(unet_discriminator_hat_forward, False)
    o_ac, g_ac = run(unet_discriminator_hat_forward, True)
    o_my, g_my = run(unet_discriminator_hat, False)
    assert o_my.shape == o_ref.shape and o_my.dtype == torch.float32
    e_out, e_ac = rel_l2(o_my, o_ref), rel_l2(o_ac, o_ref)
    print(f"hat logits rel-L2 {e_out:.4f} (autocast oracle {e_ac:.4f}) max-abs {max_abs(o_my, o_ref):.4f} on {o_ref.abs().max().item():.3f}")
    assert e_out < 2e-2 and e_out < 1.6 * e_ac + 5e-3
    names = ["x", "conv0.bias", "conv9.bias"] + [f"conv{i}" for i in range(10)]
    bad = []
    for n, a, b, r in zip(names, g_my, g_ac, g_ref):
        assert a is not None and a.shape == r.shape and a.dtype == r.dtype, n
        ea, eb = rel_l2(a, r), rel_l2(b, r)
        ma, mb, top = max_abs(a, r), max_abs(b, r), r.abs().max().item()
        print(f"  grad {n:10s} rel-L2 {ea:.4f} (autocast oracle {eb:.4f})  max-abs {ma:.3e} ({mb:.3e}) on {top:.3e}")
        if not (ea < 2.0 * eb + 5e-3 and ma < 2.0 * mb + 5e-3 * top):
            bad.append((n, ea, eb, ma, mb))
    assert not bad, bad


def test_hat_module_mirror_matches_oracle_module():
    from oracle.discriminator_oracle import UNetDiscriminatorSNHat as OraD
    from superresolution_def_b200.discriminator_hat import UNetDiscriminatorSN
    torch.manual_seed(6)
    ora = OraD(1, 64).cuda().train()
    ora_ac = OraD(1, 64).cuda().train()
    mine = UNetDiscriminatorSN(1, 64).cuda().train()
    x = torch.rand(1, 1, 128, 128, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            ora(x)
    mine.load_state_dict(ora.state_dict(), strict=True)
    ora_ac.load_state_dict(ora.state_dict(), strict=True)
    xs = x.clone().requires_grad_(True)
    a, b = mine(xs), ora(x)
    with torch.autocast("cuda", dtype=BF):
        c = ora_ac(x)
    for k, v in ora.state_dict().items():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            assert rel_l2(mine.state_dict()[k], v) < 1e-5, k
    assert a.shape == b.shape and rel_l2(a, b) < 2e-2 and rel_l2(a, b) < 1.6 * rel_l2(c, b) + 5e-3, (rel_l2(a, b), rel_l2(c, b))
    g = torch.randn_like(b)
    a.backward(g); b.backward(g); c.float().backward(g)
    assert xs.grad is not None and xs.grad.shape == x.shape
    bad = []
    for (n, p), (_, q), (_, r) in zip(mine.named_parameters(), ora.named_parameters(), ora_ac.named_parameters()):
        assert p.grad is not None and p.grad.shape == q.grad.shape, n
        e_my, e_ac = rel_l2(p.grad, q.grad), rel_l2(r.grad, q.grad)
        if not e_my < 2.0 * e_ac + 5e-3:
            bad.append((n, e_my, e_ac))
    assert not bad, bad
    for p in mine.parameters():      # G step of train_hat.py:223-241: frozen discriminator, gradient to the image only
        p.requires_grad = False
    xs2 = x.clone().requires_grad_(True)
    mine(xs2).mean().backward()
    assert xs2.grad is not None and xs2.grad.abs().max() > 0 and all(p.grad is not None for p in mine.parameters())
