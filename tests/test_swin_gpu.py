"""-m gpu parity tests: CUDA path (through the C ABI) vs the oracle on the same seeded inputs.
Tolerances (bf16 tensor-core math vs fp32 oracle): outputs rel-L2 <= 2e-2, gradients rel-L2 <= 4e-2."""
import math

import pytest
import torch

from tests.util import rel_l2, max_abs, randomize_

pytestmark = pytest.mark.gpu
OUT_TOL = 2e-2
GRAD_TOL = 4e-2


def _oracle():
    from oracle import swinir_oracle as o
    return o


def test_attention_core_matches_torch():
    """srk_win_attn_fwd/bwd on packed qkv vs a direct torch evaluation (shifted windows, 2 images)."""
    from superresolution_def_b200 import _capi as capi
    torch.manual_seed(1)
    B, H, W, heads, dh, ds, shift = 2, 16, 24, 6, 30, 32, 4
    T = B * H * W
    qkv = torch.zeros(T, 3 * heads * ds, device="cuda")
    real = torch.randn(T, 3, heads, dh, device="cuda")
    qkv.view(T, 3, heads, ds)[..., :dh] = real
    qkv = qkv.to(torch.bfloat16)
    table = torch.randn(225, heads, device="cuda")
    out = torch.zeros(T, heads * ds, device="cuda", dtype=torch.bfloat16)
    geom = capi.SrkGeom(B, H, W, 8, shift)
    capi.win_attn_fwd(geom, heads, qkv, table, out, ones_col=dh)
    torch.cuda.synchronize()

    o = _oracle()
    q = qkv.float().view(B, H, W, 3, heads, ds).requires_grad_(True)

    def ref_attn(qv):
        x = torch.roll(qv, shifts=(-shift, -shift), dims=(1, 2)).reshape(B, H, W, -1)
        win = o.window_partition(x, 8).reshape(-1, 64, 3, heads, ds).permute(2, 0, 3, 1, 4)
        logits = win[0] @ win[1].transpose(-2, -1)
        idx = o.relative_position_index(8).to("cuda")
        logits = logits + table[idx.reshape(-1)].reshape(64, 64, heads).permute(2, 0, 1)[None]
        p = torch.softmax(logits, -1)
        y = (p @ win[2]).transpose(1, 2).reshape(-1, 8, 8, heads * ds)
        y = o.window_reverse(y, 8, H, W)
        return torch.roll(y, shifts=(shift, shift), dims=(1, 2)).reshape(T, heads * ds)

    ref = ref_attn(q)
    ref_out = ref.detach().clone()
    ref_out[:, dh] = 1.0
    assert rel_l2(out, ref_out) < 1e-2, (rel_l2(out, ref_out), max_abs(out, ref_out))

    dout = torch.zeros(T, heads, ds, device="cuda")
    dout[..., :dh] = torch.randn(T, heads, dh, device="cuda")
    dout = dout.view(T, heads * ds).to(torch.bfloat16)
    table_r = table.clone().requires_grad_(True)
    table_saved, table = table, table_r
    ref = ref_attn(q)
    ref.backward(dout.float())
    dqkv = torch.zeros_like(qkv)
    dtab = torch.zeros(225, heads, device="cuda")
    capi.win_attn_bwd(geom, heads, qkv, table_saved, dout, dqkv, dtab)
    torch.cuda.synchronize()
    e = rel_l2(dqkv, q.grad.reshape(T, -1))
    assert e < 2e-2, e
    e = rel_l2(dtab, table_r.grad)
    assert e < 2e-2, e


@pytest.mark.parametrize("shift", [0, 4])
def test_swin_block_matches_oracle(shift):
    from superresolution_def_b200.architecture_swin import SwinTransformerBlock
    o = _oracle()
    torch.manual_seed(2)
    B, R, C, heads = 2, 16, 180, 6
    blk = randomize_(SwinTransformerBlock(C, (R, R), heads, window_size=8, shift_size=shift), seed=3).cuda()
    x = torch.randn(B, R * R, C, device="cuda")
    xr = x.clone().requires_grad_(True)
    xm = x.clone().requires_grad_(True)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in blk.state_dict().items()}
    ref = o.swin_block(xr, sd, "", (R, R), heads, 8, shift)
    got = blk(xm)
    assert got.shape == ref.shape and got.dtype == x.dtype
    assert rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (got * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL, rel_l2(xm.grad, xr.grad)
    worst = {}
    for name, p in blk.named_parameters():
        worst[name] = rel_l2(p.grad, sd[name].grad)
    bad = {k: v for k, v in worst.items() if v > GRAD_TOL}
    assert not bad, worst


def test_swinir_small_matches_oracle():
    """Whole generator, forward + every parameter gradient.  Two pins: the fp32 oracle (absolute tolerance) and the
    same oracle under bf16 autocast — the reference's own training dtype — whose distance to fp32 calibrates what
    bf16 arithmetic can deliver through this depth: ours must stay within 1.6x of it (+1e-2)."""
    from superresolution_def_b200.architecture_swin import SwinIR
    o = _oracle()
    torch.manual_seed(4)
    kw = dict(img_size=16, window_size=8, depths=[2, 2], num_heads=[6, 6])
    net = randomize_(SwinIR(upscale=4, in_chans=1, embed_dim=180, mlp_ratio=2, **kw), seed=5, table_std=0.5).cuda()
    x = torch.rand(2, 1, 16, 16, device="cuda")
    w = torch.randn(2, 1, 64, 64, device="cuda")

    def run_oracle(autocast):
        sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in net.state_dict().items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = o.swinir_forward(x, sd, upscale=4, **kw)
        (out.float() * w).mean().backward()
        return out, sd

    ref, sd32 = run_oracle(False)
    r16, sd16 = run_oracle(True)
    got = net(x)
    assert got.shape == ref.shape == (2, 1, 64, 64)
    assert rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    assert rel_l2(got, ref) < 1.6 * rel_l2(r16, ref) + 1e-2
    (got.float() * w).mean().backward()
    bad = {}
    for n, p in net.named_parameters():
        mine, auto = rel_l2(p.grad, sd32[n].grad), rel_l2(sd16[n].grad, sd32[n].grad)
        if mine > 1.6 * auto + 1e-2:
            bad[n] = (round(mine, 4), round(auto, 4))
    assert not bad, bad


def test_standalone_window_attention_and_mlp_modules():
    """The module-level interfaces WindowAttention.forward(x, mask=None) and Mlp.forward(x) (reference
    architecture_swin.py:71-96, :19-25) called on their own, outputs + all gradients vs the oracle."""
    from superresolution_def_b200.architecture_swin import WindowAttention, Mlp
    o = _oracle()
    torch.manual_seed(7)
    att = randomize_(WindowAttention(180, (8, 8), 6), seed=11).cuda()
    x = torch.randn(5, 64, 180, device="cuda")
    xm, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in att.state_dict().items()}
    got, ref = att(xm), o.window_attention(xr, sd, "", 6, 8)
    assert rel_l2(got, ref) < OUT_TOL
    w = torch.randn_like(ref)
    (got * w).sum().backward(); (ref * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL
    for n, p in att.named_parameters():
        assert rel_l2(p.grad, sd[n].grad) < GRAD_TOL, n
    mlp = randomize_(Mlp(180, 720), seed=12).cuda()
    x = torch.randn(3, 50, 180, device="cuda")
    xm, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in mlp.state_dict().items()}
    got, ref = mlp(xm), o.mlp(xr, sd, "")
    assert rel_l2(got, ref) < OUT_TOL
    w = torch.randn_like(ref)
    (got * w).sum().backward(); (ref * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < GRAD_TOL
    for n, p in mlp.named_parameters():
        assert rel_l2(p.grad, sd[n].grad) < GRAD_TOL, n
