"""CPU test of the discriminator's HOST logic (superresolution_def_b200/disc_engine.py): buffer layout per U-Net level,
channel-slice views standing in for torch.cat, weight operand layouts, orientation / un-permutation of the weight-gradient
GEMMs, which gradient feeds which fold.  Every libsrk entry point the engine calls is replaced by a torch restatement of its
documented contract (include/srk.h) in fp32 — so the autograd node must reproduce the oracle's logits and all 13 gradients
to fp32 round-off; any indexing mistake shows up as an O(1) error.  The kernels themselves are checked on the GPU
(tests/test_disc_gpu.py): this file never touches CUDA."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2


class _View:
    def __init__(self, t, c0, C):
        self.t, self.c0, self.C = t, c0, C

    def get(self, n):
        return self.t[:n, self.c0:self.c0 + self.C]

    def put(self, n, v):
        self.t[:n, self.c0:self.c0 + self.C] = v


def _nchw(v, B, H, W):
    return v.get(B * H * W).reshape(B, H, W, v.C).permute(0, 3, 1, 2)


def _rows(img):   # [B,C,H,W] -> [B*H*W, C]
    return img.permute(0, 2, 3, 1).reshape(-1, img.shape[1])


@pytest.fixture
def engine(monkeypatch):
    from superresolution_def_b200 import disc_engine as de
    capi = de.capi
    monkeypatch.setattr(de, "BF16", torch.float32)
    weights = {}

    def patches(x, f, slope, B, H, W, p):
        src = _nchw(x, B, H, W)
        if f is not None:
            src = torch.where(_nchw(f, B, H, W) > 0, src, src * slope)
        M = B * (H // 2) * (W // 2)
        p[:M] = F.unfold(src, 4, stride=2, padding=1).reshape(B, x.C, 16, -1).permute(0, 3, 2, 1).reshape(M, 16 * x.C)

    def fold(taps, B, Hi, Wi, y, add=None, f=None, act=0, slope=0.2):
        C = y.C
        cols = taps[:B * Hi * Wi].reshape(B, Hi * Wi, 16, C).permute(0, 3, 2, 1).reshape(B, C * 16, Hi * Wi)
        w = F.fold(cols, (2 * Hi, 2 * Wi), 4, stride=2, padding=1)
        if add is not None:
            w = w + _nchw(add, B, 2 * Hi, 2 * Wi)
        if act == capi.FOLD_LRELU:
            w = torch.where(w > 0, w, w * slope)
        elif act == capi.FOLD_MASK:
            w = torch.where(_nchw(f, B, 2 * Hi, 2 * Wi) > 0, w, w * slope)
        y.put(B * 4 * Hi * Wi, _rows(w))

    def gemm_tn(epi, A, B, C, **kw):
        assert epi == capi.EPI_STORE and A.shape[0] % 128 == 0 and A.shape[1] % 64 == 0 and B.shape[0] % 64 == 0
        C.copy_(A @ B.t())

    def gemm_tn_lrelu(A, B, C, slope):
        assert A.shape[0] % 128 == 0 and A.shape[1] % 64 == 0 and B.shape[0] % 64 == 0
        C.copy_(F.leaky_relu(A @ B.t(), slope))

    def view_lrelu(y, n, slope):
        y.put(n, F.leaky_relu(y.get(n), slope))

    def view_lrelu_mask(g, f, n, slope, colsum=None):
        assert g.C <= 256
        g.put(n, torch.where(f.get(n) > 0, g.get(n), g.get(n) * slope))

    def conv_in1_fwd(x, w, b, y, B, H, W, C, Cp):
        y[:B * H * W] = _rows(F.conv2d(x.reshape(B, 1, H, W), w, b, 1, 1))

    @torch.enable_grad()
    def conv_in1_wgrad(x, dy, dw, db, B, H, W, C, Cp):
        w = torch.zeros(C, 1, 3, 3, requires_grad=True)
        d = dy[:B * H * W].reshape(B, H, W, C).permute(0, 3, 1, 2)
        F.conv2d(x.reshape(B, 1, H, W), w, None, 1, 1).backward(d)
        dw.copy_(w.grad)
        db.copy_(d.sum((0, 2, 3)))

    def conv_out1_fwd(x, w, b, y, B, H, W, C):
        y.copy_(F.conv2d(x[:B * H * W].reshape(B, H, W, C).permute(0, 3, 1, 2), w, b, 1, 1))

    @torch.enable_grad()
    def conv_out1_bwd(dy, x, w, dx, dw, db, B, H, W, C):
        xi = x[:B * H * W].reshape(B, H, W, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
        ww = w.clone().requires_grad_(True)
        F.conv2d(xi, ww, None, 1, 1).backward(dy)
        dx[:B * H * W] = _rows(xi.grad)
        dw.copy_(ww.grad)
        db.copy_(dy.sum().reshape(1))

    def prep(w, b, Cout_p, Cin_p, ps, wf, wt, bp):
        assert b is None and not ps and tuple(w.shape) == (Cout_p, Cin_p, 3, 3) and w.is_contiguous()
        weights[wf.data_ptr()] = ("fwd", w.clone())
        weights[wt.data_ptr()] = ("dgrad", w.clone())
        bp.zero_()

    def igemm_v(epi, B, H, W, Cin_p, Cout_p, n_real, x, wk, bias, y, r=None, slope=0.01, alpha=1.0, y32=None):
        kind, w = weights[wk.data_ptr()]
        assert Cin_p <= 256 and Cout_p <= 256 and H % 8 == 0 and W % 16 == 0 and x.C == Cin_p and y.C == Cout_p
        assert tuple(w.shape[:2]) == ((Cout_p, Cin_p) if kind == "fwd" else (Cin_p, Cout_p))
        xi = _nchw(x, B, H, W)
        o = F.conv2d(xi, w, None, 1, 1) if kind == "fwd" else F.conv_transpose2d(xi, w, None, 1, 1)
        if epi == capi.CEPI_BIAS_LRELU:
            o = F.leaky_relu(o, slope)
        elif epi == capi.CEPI_BIAS_RES:
            o = alpha * o + _nchw(r, B, H, W)
        else:
            assert epi == capi.CEPI_BIAS
        y.put(B * H * W, _rows(o))

    def igemm(epi, B, H, W, Cin_p, Cout_p, n_real, x, wk, bias, y, x_ps=False, y_ps=False, y2=None, r=None, slope=0.01):
        igemm_v(epi, B, H, W, Cin_p, Cout_p, n_real, _View(x, 0, x.shape[1]), wk, bias, _View(y, 0, y.shape[1]), slope=slope)

    @torch.enable_grad()
    def wgrad_v(B, H, W, Cin, Cout, Cin_p, Cout_p, dy, x, dw):
        assert H % 4 == 0 and W % 16 == 0 and Cin_p <= 256 and x.C == Cin and dy.C == Cout
        w = torch.zeros(Cout, Cin, 3, 3, requires_grad=True)
        F.conv2d(_nchw(x, B, H, W), w, None, 1, 1).backward(_nchw(dy, B, H, W))
        dw.copy_(w.grad)

    def conv3x3_wgrad(B, H, W, Cin, Cout, Cin_p, Cout_p, ps, dy, x, dw):
        wgrad_v(B, H, W, Cin, Cout, Cin_p, Cout_p, _View(dy, 0, dy.shape[1]), _View(x, 0, x.shape[1]), dw)

    def bilinear2x_fwd(x, s_, y, B, H, W):
        src = _nchw(x, B, H, W) + (0 if s_ is None else _nchw(s_, B, H, W))
        y.put(B * 4 * H * W, _rows(F.interpolate(src, scale_factor=2, mode="bilinear", align_corners=False)))

    @torch.enable_grad()
    def bilinear2x_bwd(dy, dx, B, H, W):
        z = torch.zeros(B, dx.C, H, W, requires_grad=True)
        F.interpolate(z, scale_factor=2, mode="bilinear", align_corners=False).backward(_nchw(dy, B, 2 * H, 2 * W))
        dx.put(B * H * W, _rows(z.grad))

    def view_axpy(y, a, x, npix, alpha):
        y.put(npix, alpha * a.get(npix) + (0 if x is None else x.get(npix)))

    def disc_prep_w4(w, a, at=None, sigma=None):
        P, Q = w.shape[0], w.shape[1]
        assert P % 32 == 0 and Q % 32 == 0 and w.dtype == torch.float32 and w.is_contiguous()
        a.copy_(w.permute(0, 2, 3, 1).reshape(P, 16 * Q) / (1.0 if sigma is None else sigma))
        if at is not None:
            at.copy_(a.t())

    def disc_wgrad4(A, B, R, dw):
        T, Cb = B.shape
        assert T % 64 == 0 and tuple(A.shape) == (T, 16 * R) and (Cb in (64, 128, 192, 256) or Cb % 256 == 0)
        dw.copy_((A.t() @ B).view(4, 4, R, Cb).permute(3, 2, 0, 1))

    # srk_spectral_norm / _bwd restated from their contract in include/srk.h (NOT from torch's hook: the test below compares
    # the result with the hook-driven oracle module)
    def sn_layers(ws, us, vs, dims, sigmas, w_sn=None):
        return [(w, u, v, dm, sigmas, i, None if w_sn is None else w_sn[i]) for i, (w, u, v, dm) in enumerate(zip(ws, us, vs, dims))]

    def _mat(w, dm):
        return w.reshape(w.shape[0], -1) if dm == 0 else w.permute(1, 0, 2, 3).reshape(w.shape[1], -1)

    def spectral_norm(layers, power_iteration, eps, device):
        for w, u, v, dm, sigmas, i, out in layers:
            wm = _mat(w, dm)
            if power_iteration:
                t = wm.t() @ u
                v.copy_(t / t.norm().clamp_min(eps))
            s_ = wm @ v
            if power_iteration:
                u.copy_(s_ / s_.norm().clamp_min(eps))
            sigmas[i] = u @ s_
            if out is not None:
                out.copy_(w / sigmas[i])

    def spectral_norm_bwd(layers, dw_sn, dw, device):
        for (w, u, v, dm, sigmas, i, _), g, o in zip(layers, dw_sn, dw):
            if g is None:
                continue
            sg = sigmas[i]
            uv = torch.outer(u, v)
            uv = uv.reshape(w.shape) if dm == 0 else uv.reshape(w.shape[1], w.shape[0], *w.shape[2:]).permute(1, 0, 2, 3)
            o.copy_(g / sg - ((g * w).sum() / sg ** 2) * uv)

    for name, fn in dict(sn_layers=sn_layers, spectral_norm=spectral_norm, spectral_norm_bwd=spectral_norm_bwd, view=lambda t, c0=0, C=None: _View(t, c0, t.shape[1] - c0 if C is None else C),
                         disc_patches_k4s2=patches, disc_fold_k4s2=fold, gemm_tn=gemm_tn, gemm_tn_lrelu=gemm_tn_lrelu,
                         view_lrelu=view_lrelu, view_lrelu_mask=view_lrelu_mask, conv_in1_fwd=conv_in1_fwd,
                         conv_in1_wgrad=conv_in1_wgrad, conv_out1_fwd=conv_out1_fwd, conv_out1_bwd=conv_out1_bwd,
                         conv3x3_prep_weights=prep, conv3x3_igemm=igemm, conv3x3_igemm_v=igemm_v, conv3x3_wgrad=conv3x3_wgrad,
                         conv3x3_wgrad_v=wgrad_v, bilinear2x_fwd=bilinear2x_fwd, bilinear2x_bwd=bilinear2x_bwd, view_axpy=view_axpy,
                         disc_prep_w4=disc_prep_w4,
                         disc_wgrad4=disc_wgrad4).items():
        monkeypatch.setattr(capi, name, fn)
    return de


def _weights():
    nf = 64
    shapes = [(nf, 1, 3, 3), (nf, nf, 4, 4), (2 * nf, nf, 4, 4), (4 * nf, 2 * nf, 4, 4), (8 * nf, 4 * nf, 4, 4), (8 * nf, 8 * nf, 4, 4),
              (8 * nf, 8 * nf, 4, 4), (16 * nf, 4 * nf, 4, 4), (8 * nf, 2 * nf, 4, 4), (4 * nf, nf, 4, 4), (nf, 2 * nf, 3, 3), (1, nf, 3, 3)]
    g = torch.Generator().manual_seed(0)
    ws = []
    for i, s in enumerate(shapes):
        transposed = 6 <= i <= 9
        fan_in = (s[0] if transposed else s[1]) * s[2] * s[3] / (4 if transposed else 1)
        ws.append(torch.randn(s, generator=g) * 1.6 * fan_in ** -0.5)
    return ws


@pytest.mark.parametrize("B,H,W", [(1, 64, 96), (3, 32, 32)])
def test_engine_host_logic_reproduces_the_oracle_in_fp32(engine, B, H, W):
    from oracle.discriminator_oracle import unet_discriminator_forward
    ws = _weights()
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, 1, H, W, generator=g)
    dout = torch.randn(B, 1, H // 2, W // 2, generator=g)

    def run(fn):
        xs = x.clone().requires_grad_(True)
        wl = [w.clone().requires_grad_(True) for w in ws]
        out = fn(xs, wl)
        out.backward(dout)
        return out.detach(), [xs.grad] + [w.grad for w in wl]

    o_ref, g_ref = run(unet_discriminator_forward)
    o_my, g_my = run(lambda xs, wl: engine.UNetDiscriminatorFunction.apply(xs, None, *wl))
    assert rel_l2(o_my, o_ref) < 1e-5
    for i, (a, b) in enumerate(zip(g_my, g_ref)):
        assert a.shape == b.shape and rel_l2(a, b) < 1e-4, (i, rel_l2(a, b))


def test_engine_returns_only_the_requested_gradients(engine):
    ws = _weights()
    x = torch.rand(1, 1, 32, 32)
    xs = x.clone().requires_grad_(True)
    out = engine.UNetDiscriminatorFunction.apply(xs, None, *ws)          # G step: frozen parameters
    out.mean().backward()
    assert xs.grad is not None and xs.grad.abs().max() > 0
    wl = [w.clone().requires_grad_(True) for w in ws]
    engine.UNetDiscriminatorFunction.apply(x, None, *wl).mean().backward()   # D step: detached image
    assert all(w.grad is not None and w.grad.shape == w.shape for w in wl)
    assert not engine.UNetDiscriminatorFunction.apply(x, None, *ws).requires_grad
    with pytest.raises(Exception):
        engine.UNetDiscriminatorFunction.apply(torch.rand(1, 1, 48, 64), None, *ws)


@pytest.mark.parametrize("training", [True, False])
def test_spectral_norm_path_matches_the_hook_driven_oracle_module(engine, training):
    """weight_orig / weight_u / weight_v in, as gan.UNetDiscriminatorSN.forward passes them: logits, weight_orig gradients
    and the in-place buffer update equal the oracle module driven by torch.nn.utils.spectral_norm's own hook — across TWO
    forwards before the backward (the D step runs D(hr) and D(sr) first: the second power iteration must not disturb the
    first forward's backward)."""
    from oracle.discriminator_oracle import UNetDiscriminatorSN as OraD
    torch.manual_seed(3)
    ora, twin = OraD(1, 64), OraD(1, 64)
    with torch.no_grad():   # a few power iterations first: with the random initial u / v, sigma is ~1e-3 per layer and the
        for _ in range(3):  # logits of the 12-layer stack overflow any meaningful comparison
            ora(torch.rand(1, 1, 32, 32))
    twin.load_state_dict(ora.state_dict())
    ora.train(training); twin.train(training)
    convs = twin.convs()
    x1, x2 = torch.rand(1, 1, 32, 32), torch.rand(1, 1, 32, 32)

    def mine(x):
        sn = dict(u=[m.weight_u for m in convs], v=[m.weight_v for m in convs], training=training, eps=1e-12)
        return engine.UNetDiscriminatorFunction.apply(x, sn, *[m.weight_orig for m in convs])

    a1, a2 = mine(x1), mine(x2)
    b1, b2 = ora(x1), ora(x2)
    assert rel_l2(a1, b1) < 1e-5 and rel_l2(a2, b2) < 1e-5
    for k, v in ora.state_dict().items():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            assert rel_l2(twin.state_dict()[k], v) < 1e-5, k
    (a1.mean() + 2 * a2.mean()).backward()
    (b1.mean() + 2 * b2.mean()).backward()
    for (n, p), (_, q) in zip(twin.named_parameters(), ora.named_parameters()):
        assert p.grad is not None and rel_l2(p.grad, q.grad) < 2e-4, (n, rel_l2(p.grad, q.grad))


def test_public_entry_refuses_cpu_tensors():
    from superresolution_def_b200 import disc_engine as de
    from superresolution_def_b200._capi import SrkError
    with pytest.raises(SrkError):
        de.unet_discriminator(torch.rand(1, 1, 32, 32), _weights())


@pytest.mark.parametrize("skip,training", [(True, True), (False, True), (True, False)])
def test_hat_variant_matches_the_hook_driven_oracle_module(engine, skip, training):
    """models/discriminator_hat.py: the whole module path (weights + biases + spectral-norm buffers in, as
    superresolution_def_b200.discriminator_hat.UNetDiscriminatorSN.forward passes them) against the oracle module, two
    forwards before the backward, fp32 round-off."""
    from oracle.discriminator_oracle import UNetDiscriminatorSNHat as OraD
    torch.manual_seed(4)
    ora, twin = OraD(1, 64, skip), OraD(1, 64, skip)
    with torch.no_grad():
        for _ in range(3):
            ora(torch.rand(1, 1, 32, 64))
    twin.load_state_dict(ora.state_dict())
    ora.train(training); twin.train(training)
    sn = twin.convs()[1:9]
    x1, x2 = torch.rand(1, 1, 32, 64), torch.rand(2, 1, 64, 64)
    x2g = x2.clone().requires_grad_(True)
    x2o = x2.clone().requires_grad_(True)

    def mine(x):
        d = dict(u=[m.weight_u for m in sn], v=[m.weight_v for m in sn], training=training, eps=1e-12)
        ws = [twin.conv0.weight] + [m.weight_orig for m in sn] + [twin.conv9.weight]
        return engine.UNetDiscriminatorHatFunction.apply(x, d, skip, twin.conv0.bias, twin.conv9.bias, *ws)

    a1, a2 = mine(x1), mine(x2g)
    b1, b2 = ora(x1), ora(x2o)
    assert a1.shape == b1.shape and rel_l2(a1, b1) < 1e-5 and rel_l2(a2, b2) < 1e-5
    for k, v in ora.state_dict().items():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            assert rel_l2(twin.state_dict()[k], v) < 1e-5, k
    g1, g2 = torch.randn_like(b1), torch.randn_like(b2)
    torch.autograd.backward([a1, a2], [g1, g2])
    torch.autograd.backward([b1, b2], [g1, g2])
    assert rel_l2(x2g.grad, x2o.grad) < 1e-2     # see below
    # Layers downstream of every LeakyReLU agree to 1e-6.  Upstream, ONE activation out of ~1e6 whose pre-activation is within
    # fp32 round-off of zero takes the other slope (summation order differs between a fold and a direct convolution) and
    # moves every gradient behind it by ~1e-3 rel-L2 (sqrt(1e-6) x 0.8); an indexing mistake is O(1).
    for (n, p), (_, q) in zip(twin.named_parameters(), ora.named_parameters()):
        assert p.grad is not None and p.grad.shape == q.grad.shape and rel_l2(p.grad, q.grad) < 1e-2, (n, rel_l2(p.grad, q.grad))
    for n in ("conv7.weight_orig", "conv8.weight_orig", "conv9.weight", "conv9.bias"):
        assert rel_l2(dict(twin.named_parameters())[n].grad, dict(ora.named_parameters())[n].grad) < 1e-5, n
