"""-m gpu parity tests of the convolutional head/tail kernels against plain torch fp32 on the same (bf16-rounded)
inputs.  Tolerances: outputs rel-L2 <= 1e-2; every gradient tensor within 1.6x (+5e-3) of the error the same ATen sequence
makes under bf16 autocast (the reference's training arithmetic), rel-L2 and max-abs.  Gradients upstream of
conv_before_upsample are dominated by LeakyReLU sign flips (a pre-activation within bf16 rounding distance of zero gets slope
0.01 instead of 1 on one side of the comparison) — the autocast run has the same flips, which is why it is the yardstick
rather than a blanket bound; layers after the LeakyReLU agree to <= 1e-2 (asserted separately)."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2, max_abs

pytestmark = pytest.mark.gpu


def _mk(shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (1, 24, 48), (1, 128, 128)])
def test_tail_matches_torch(B, H, W):
    from superresolution_def_b200.conv_engine import SwinIRTailFunction
    C, Cp = 180, 192
    T = B * H * W
    body = torch.zeros(T, Cp, device="cuda"); body[:, :C] = _mk((T, C), seed=1); body[:, C] = 1.0
    first = torch.zeros(T, Cp, device="cuda"); first[:, :C] = _mk((T, C), seed=2)
    body, first = body.to(torch.bfloat16), first.to(torch.bfloat16)
    ws = [_mk((180, 180, 3, 3), 0.03, 3), _mk((180,), 0.1, 4), _mk((64, 180, 3, 3), 0.03, 5), _mk((64,), 0.1, 6),
          _mk((256, 64, 3, 3), 0.05, 7), _mk((256,), 0.1, 8), _mk((256, 64, 3, 3), 0.05, 9), _mk((256,), 0.1, 10),
          _mk((1, 64, 3, 3), 0.05, 11), _mk((1,), 0.1, 12)]
    mine = [w.clone().requires_grad_(True) for w in ws]
    ref = [w.clone().requires_grad_(True) for w in ws]
    bm, fm = body.clone().requires_grad_(True), first.clone().requires_grad_(True)
    torch.manual_seed(1234)
    out = SwinIRTailFunction.apply(bm, fm, (B, H, W), C, *mine)

    def nchw(t):
        return t.float()[:, :C].reshape(B, H, W, C).permute(0, 3, 1, 2)

    w = None

    def run_ref(autocast):
        """The reference's ATen sequence (architecture_swin.py:249-255) in fp32, or under the bf16 autocast the scripts
        train with: the latter's distance to fp32 calibrates the per-tensor bound (LeakyReLU sign flips included)."""
        nonlocal w
        ps = [t.detach().clone().requires_grad_(True) for t in ws]
        br, fr = body.float().clone().requires_grad_(True), first.float().clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            x = F.conv2d(nchw(br), ps[0], ps[1], padding=1) + nchw(fr)
            x = F.leaky_relu(F.conv2d(x, ps[2], ps[3], padding=1), 0.01)
            x = F.pixel_shuffle(F.conv2d(x, ps[4], ps[5], padding=1), 2)
            x = F.pixel_shuffle(F.conv2d(x, ps[6], ps[7], padding=1), 2)
            ro = F.conv2d(x, ps[8], ps[9], padding=1)
        if w is None:
            w = torch.randn_like(ro.float())
        (ro.float() * w).sum().backward()
        g = {i: p.grad for i, p in enumerate(ps)}
        g["body"], g["first"] = br.grad[:, :C], fr.grad[:, :C]
        return ro.float().detach(), g

    ro, g32 = run_ref(False)
    r16, g16 = run_ref(True)
    assert out.shape == ro.shape
    assert rel_l2(out, ro) < 1e-2 and rel_l2(out, ro) < 1.6 * rel_l2(r16, ro) + 2e-3, (rel_l2(out, ro), rel_l2(r16, ro))
    assert max_abs(out, ro) < 3 * max_abs(r16, ro) + 1e-3 * ro.abs().max().item(), (max_abs(out, ro), max_abs(r16, ro))
    (out * w).sum().backward()
    gm = {i: a.grad for i, a in enumerate(mine)}
    gm["body"], gm["first"] = bm.grad[:, :C], fm.grad[:, :C]
    errs = {k: (rel_l2(gm[k], g32[k]), rel_l2(g16[k], g32[k])) for k in gm}
    print({k: (round(a, 4), round(b, 4)) for k, (a, b) in errs.items()})
    # per tensor: within 1.6x (+5e-3) of what the reference's own bf16-autocast arithmetic delivers for that tensor
    bad = {k: v for k, v in errs.items() if v[0] > 1.6 * v[1] + 5e-3}
    assert not bad, str(bad)
    for k in gm:
        scale = g32[k].abs().max().item()
        assert max_abs(gm[k], g32[k]) < 3 * max_abs(g16[k], g32[k]) + 2e-2 * scale, (k, max_abs(gm[k], g32[k]), scale)
    assert all(errs[i][0] < 1e-2 for i in range(4, 10)), str(errs)
    assert bm.grad[:, C:].abs().max() == 0 and fm.grad[:, C:].abs().max() == 0


def test_conv_first_matches_torch():
    from superresolution_def_b200.conv_engine import ConvFirstFunction
    B, H, W, C, Cp = 2, 16, 24, 180, 192
    x = torch.rand(B, 1, H, W, device="cuda")
    w, b = _mk((C, 1, 3, 3), 0.3, 1), _mk((C,), 0.1, 2)
    wm, bm = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = ConvFirstFunction.apply(x, wm, bm, Cp)
    ref = F.conv2d(x, wr, br, padding=1).permute(0, 2, 3, 1).reshape(B * H * W, C)
    assert rel_l2(y[:, :C], ref) < 5e-3 and y[:, C:].abs().max() == 0
    g = torch.zeros(B * H * W, Cp, device="cuda"); g[:, :C] = torch.randn(B * H * W, C, device="cuda")
    g = g.to(torch.bfloat16)
    (y.float() * g.float()).sum().backward()
    (ref * g.float()[:, :C]).sum().backward()
    assert rel_l2(wm.grad, wr.grad) < 1e-2 and rel_l2(bm.grad, br.grad) < 1e-2
