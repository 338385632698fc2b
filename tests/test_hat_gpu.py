"""-m gpu parity tests of the HAT path: CUDA kernels (through the C ABI) vs the oracle (oracle/hat_oracle.py, itself
pinned to the unmodified reference) on the same seeded inputs.
Tolerances (bf16 tensor-core math vs fp32 oracle): outputs rel-L2 <= 2e-2, gradients rel-L2 <= 4e-2."""
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


@pytest.mark.parametrize("shift", [0, 8])
def test_attn16_self_core_matches_torch(shift):
    """srk_win_attn16_fwd/bwd (SELF) on packed qkv vs torch: shifted 16x16 windows + the 0/-100 mask, 2 images 32x48."""
    from superresolution_def_b200 import _capi as capi
    from oracle import swinir_oracle as so
    ho = _ho()
    B, H, W, heads, dh, ds = 2, 32, 48, 6, 30, 32
    T = B * H * W
    qkv = _packed_qkv(T, heads, dh, ds, 1)
    table = torch.randn(961, heads, device="cuda")
    out = torch.zeros(T, heads * ds, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(heads, T, device="cuda")
    geom = capi.SrkGeom(B, H, W, 16, shift)
    capi.win_attn16_fwd(geom, capi.ATTN_SELF, heads, qkv, table, out, lse, ones_col=dh)
    torch.cuda.synchronize()
    idx = ho.rpi_sa(16).cuda()
    mask = ho.shift_mask(H, W, 16, 8).cuda() if shift else None
    q = qkv.float().view(B, H, W, 3, heads, ds).requires_grad_(True)
    tab = table.clone().requires_grad_(True)

    def ref_attn(qv, tb):
        x = torch.roll(qv, shifts=(-shift, -shift), dims=(1, 2)).reshape(B, H, W, -1)
        win = so.window_partition(x, 16).reshape(-1, 256, 3, heads, ds).permute(2, 0, 3, 1, 4)
        logits = win[0] @ win[1].transpose(-2, -1)
        logits = logits + tb[idx.reshape(-1)].reshape(256, 256, heads).permute(2, 0, 1)[None]
        if mask is not None:
            nw = mask.shape[0]
            logits = (logits.reshape(B, nw, heads, 256, 256) + mask[None, :, None]).reshape(-1, heads, 256, 256)
        p = torch.softmax(logits, -1)
        y = (p @ win[2]).transpose(1, 2).reshape(-1, 16, 16, heads * ds)
        y = so.window_reverse(y, 16, H, W)
        return torch.roll(y, shifts=(shift, shift), dims=(1, 2)).reshape(T, heads * ds), logits

    ref, logits = ref_attn(q, tab)
    ref_out = ref.detach().clone()
    ref_out[:, dh] = 1.0
    assert rel_l2(out, ref_out) < 1e-2, (rel_l2(out, ref_out), max_abs(out, ref_out))
    dout = torch.zeros(T, heads, ds, device="cuda")
    dout[..., :dh] = torch.randn(T, heads, dh, device="cuda")
    dout = dout.view(T, heads * ds).to(torch.bfloat16)
    ref.backward(dout.float())
    dqkv = torch.zeros_like(qkv)
    dtab = torch.zeros(961, heads, device="cuda")
    ws = torch.empty(capi.attn16_bwd_ws_bytes(geom, capi.ATTN_SELF, heads), device="cuda", dtype=torch.uint8)
    capi.win_attn16_bwd(geom, capi.ATTN_SELF, heads, qkv, table, out, dout, lse, dqkv, ws, dtab)
    torch.cuda.synchronize()
    e = rel_l2(dqkv, q.grad.reshape(T, -1))
    assert e < 2e-2, e
    e = rel_l2(dtab, tab.grad)
    assert e < 2e-2, e


def test_attn16_oca_core_matches_torch():
    """srk_win_attn16_fwd/bwd (OCA): 24x24 zero-padded halo key windows, wrap-around bias indices, overlapping dK/dV."""
    from superresolution_def_b200 import _capi as capi
    from oracle import swinir_oracle as so
    ho = _ho()
    B, H, W, heads, dh, ds = 2, 32, 48, 6, 30, 32
    T = B * H * W
    qkv = _packed_qkv(T, heads, dh, ds, 2)
    table = torch.randn(1521, heads, device="cuda")
    out = torch.zeros(T, heads * ds, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(heads, T, device="cuda")
    geom = capi.SrkGeom(B, H, W, 16, 0)
    capi.win_attn16_fwd(geom, capi.ATTN_OCA, heads, qkv, table, out, lse, ones_col=dh)
    torch.cuda.synchronize()
    idx = ho.rpi_oca(16).cuda()
    q = qkv.float().view(B, H, W, 3, heads * ds).requires_grad_(True)
    tab = table.clone().requires_grad_(True)

    def ref_attn(qv, tb):
        qq = so.window_partition(qv[:, :, :, 0], 16).reshape(-1, 256, heads, ds).permute(0, 2, 1, 3)
        kv = torch.cat((qv[:, :, :, 1], qv[:, :, :, 2]), dim=-1).permute(0, 3, 1, 2)           # b, 2c, h, w
        kvw = F.unfold(kv, kernel_size=(24, 24), stride=16, padding=4)
        nw = kvw.shape[-1]
        c = heads * ds
        kvw = kvw.reshape(B, 2, c, 576, nw).permute(1, 0, 4, 3, 2).reshape(2, B * nw, 576, heads, ds).permute(0, 1, 3, 2, 4)
        logits = qq @ kvw[0].transpose(-2, -1)
        logits = logits + tb[idx.reshape(-1)].reshape(256, 576, heads).permute(2, 0, 1)[None]
        p = torch.softmax(logits, -1)
        y = (p @ kvw[1]).transpose(1, 2).reshape(-1, 16, 16, c)
        return so.window_reverse(y, 16, H, W).reshape(T, c)

    ref = ref_attn(q, tab)
    ref_out = ref.detach().clone()
    ref_out[:, dh] = 1.0
    assert rel_l2(out, ref_out) < 1e-2, (rel_l2(out, ref_out), max_abs(out, ref_out))
    dout = torch.zeros(T, heads, ds, device="cuda")
    dout[..., :dh] = torch.randn(T, heads, dh, device="cuda")
    dout = dout.view(T, heads * ds).to(torch.bfloat16)
    ref.backward(dout.float())
    dqkv = torch.zeros_like(qkv)
    dtab = torch.zeros(1521, heads, device="cuda")
    ws = torch.empty(capi.attn16_bwd_ws_bytes(geom, capi.ATTN_OCA, heads), device="cuda", dtype=torch.uint8)
    capi.win_attn16_bwd(geom, capi.ATTN_OCA, heads, qkv, table, out, dout, lse, dqkv, ws, dtab)
    torch.cuda.synchronize()
    gq = q.grad.reshape(T, 3, heads * ds)
    got = dqkv.float().view(T, 3, heads * ds)
    for s, name in enumerate("qkv"):
        e = rel_l2(got[:, s], gq[:, s])
        assert e < 2e-2, (name, e)
    e = rel_l2(dtab, tab.grad)
    assert e < 2e-2, e


def _sd_of(mod):
    return {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in mod.state_dict().items()}


@pytest.mark.parametrize("shift", [0, 8])
def test_hab_matches_oracle(shift):
    from superresolution_def_b200.hat_arch import HAB
    ho = _ho()
    torch.manual_seed(3)
    B, R, C, heads = 2, 32, 180, 6
    blk = randomize_(HAB(C, (R, R), heads, window_size=16, shift_size=shift), seed=4).cuda()
    x = torch.randn(B, R * R, C, device="cuda")
    xr, xm = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    sd = _sd_of(blk)
    ref = ho.hab(xr, sd, "", (R, R), heads, 16, shift, ho.rpi_sa(16).cuda(), ho.shift_mask(R, R, 16, 8).cuda())
    got = blk(xm, (R, R), None, None)
    assert got.shape == ref.shape and rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (got * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL, rel_l2(xm.grad, xr.grad)
    worst = {n: rel_l2(p.grad, sd[n].grad) for n, p in blk.named_parameters()}
    bad = {k: round(v, 4) for k, v in worst.items() if v > GRAD_TOL}
    assert not bad, bad


def test_hab_stochastic_depth_matches_oracle():
    """Training-mode HAB with explicit per-sample drop-path factors (one sample dropped in each branch)."""
    from superresolution_def_b200.hat_arch import HAB
    ho = _ho()
    torch.manual_seed(9)
    B, R, C, heads = 3, 32, 180, 6
    blk = randomize_(HAB(C, (R, R), heads, window_size=16, shift_size=8, drop_path=0.25), seed=10).cuda().train()
    x = torch.randn(B, R * R, C, device="cuda")
    xr, xm = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    drop = (torch.tensor([1 / 0.75, 0.0, 1 / 0.75], device="cuda"), torch.tensor([0.0, 1 / 0.75, 1 / 0.75], device="cuda"))
    sd = _sd_of(blk)
    ref = ho.hab(xr, sd, "", (R, R), heads, 16, 8, ho.rpi_sa(16).cuda(), ho.shift_mask(R, R, 16, 8).cuda(), drop=drop)
    got = blk(xm, (R, R), None, None, drop=drop)
    assert rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (got * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL, rel_l2(xm.grad, xr.grad)
    bad = {n: round(rel_l2(p.grad, sd[n].grad), 4) for n, p in blk.named_parameters() if rel_l2(p.grad, sd[n].grad) > GRAD_TOL}
    assert not bad, bad
    # and the module draws its own factors in training mode without error
    out = blk(x, (R, R), None, None)
    assert torch.isfinite(out).all()


def test_ocab_matches_oracle():
    from superresolution_def_b200.hat_arch import OCAB
    ho = _ho()
    torch.manual_seed(5)
    B, R, C, heads = 2, 32, 180, 6
    blk = randomize_(OCAB(C, (R, R), 16, 0.5, heads, mlp_ratio=4), seed=6).cuda()
    x = torch.randn(B, R * R, C, device="cuda")
    xr, xm = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    sd = _sd_of(blk)
    ref = ho.ocab(xr, sd, "", (R, R), heads, 16, ho.rpi_oca(16).cuda())
    got = blk(xm, (R, R), None)
    assert rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (got * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL, rel_l2(xm.grad, xr.grad)
    worst = {n: rel_l2(p.grad, sd[n].grad) for n, p in blk.named_parameters()}
    bad = {k: round(v, 4) for k, v in worst.items() if v > GRAD_TOL}
    assert not bad, bad


def test_hat_small_matches_oracle():
    """Whole HAT generator (C=180, window 16, one RHAG of 2 HAB + OCAB) forward + every parameter gradient; the oracle
    under bf16 autocast calibrates what bf16 arithmetic can deliver (ours must stay within 1.6x of it + 1e-2)."""
    from superresolution_def_b200.hat_arch import HAT
    ho = _ho()
    torch.manual_seed(7)
    kw = dict(window_size=16, depths=(2,), num_heads=(6,))
    net = randomize_(HAT(img_size=32, in_chans=1, embed_dim=180, upscale=4, upsampler="pixelshuffle", drop_path_rate=0.0,
                         **kw), seed=8, table_std=0.5).cuda()
    x = torch.rand(2, 1, 32, 32, device="cuda")
    w = torch.randn(2, 1, 128, 128, device="cuda")

    def run_oracle(autocast):
        sd = _sd_of(net)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = ho.hat_forward(x, sd, upscale=4, **kw)
        (out.float() * w).mean().backward()
        return out, sd

    ref, sd32 = run_oracle(False)
    r16, sd16 = run_oracle(True)
    got = net(x)
    assert got.shape == ref.shape == (2, 1, 128, 128)
    assert rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    assert rel_l2(got, ref) < 1.6 * rel_l2(r16, ref) + 1e-2
    (got.float() * w).mean().backward()
    bad = {}
    for n, p in net.named_parameters():
        mine, auto = rel_l2(p.grad, sd32[n].grad), rel_l2(sd16[n].grad, sd32[n].grad)
        if mine > 1.6 * auto + 1e-2:
            bad[n] = (round(mine, 4), round(auto, 4))
    assert not bad, bad


def test_standalone_channel_attention_module():
    """ChannelAttention.forward(x NCHW) on its own (reference hat_arch.py:40-58): output, input gradient and the four
    squeeze / excite parameter gradients vs the oracle."""
    from superresolution_def_b200.hat_arch import ChannelAttention
    ho = _ho()
    torch.manual_seed(21)
    ca = randomize_(ChannelAttention(180, 30), seed=22).cuda()
    x = torch.randn(2, 180, 16, 24, device="cuda")
    xm, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    sd = _sd_of(ca)
    got, ref = ca(xm), ho.channel_attention(xr, sd, "")
    assert got.shape == ref.shape and got.dtype == ref.dtype and rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (got * w).sum().backward(); (ref * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL, rel_l2(xm.grad, xr.grad)
    bad = {n: round(rel_l2(p.grad, sd[n].grad), 4) for n, p in ca.named_parameters() if rel_l2(p.grad, sd[n].grad) > GRAD_TOL}
    assert not bad, bad


def test_standalone_cab_and_window_attention_modules():
    """CAB.forward(x NCHW) and HAT's WindowAttention.forward(x, rpi, mask=None) called on their own
    (reference hat_arch.py:61-74, :129-196): outputs + every gradient vs the oracle."""
    from superresolution_def_b200.hat_arch import CAB, WindowAttention
    ho = _ho()
    torch.manual_seed(11)
    cab = randomize_(CAB(180, 3, 30), seed=12).cuda()
    x = torch.randn(2, 180, 16, 32, device="cuda")
    xm, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    sd = _sd_of(cab)
    got, ref = cab(xm), ho.cab(xr, sd, "")
    assert got.shape == ref.shape and rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (got * w).sum().backward(); (ref * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL, rel_l2(xm.grad, xr.grad)
    bad = {n: round(rel_l2(p.grad, sd[n].grad), 4) for n, p in cab.named_parameters() if rel_l2(p.grad, sd[n].grad) > GRAD_TOL}
    assert not bad, bad
    att = randomize_(WindowAttention(180, (16, 16), 6), seed=13).cuda()
    x = torch.randn(3, 256, 180, device="cuda")
    xm, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    sd = _sd_of(att)
    got, ref = att(xm, None, None), ho.window_attention(xr, sd, "", 6, ho.rpi_sa(16).cuda(), None)
    assert rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (got * w).sum().backward(); (ref * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL
    bad = {n: round(rel_l2(p.grad, sd[n].grad), 4) for n, p in att.named_parameters() if rel_l2(p.grad, sd[n].grad) > GRAD_TOL}
    assert not bad, bad
