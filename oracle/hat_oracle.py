"""oracle/hat_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, plain-PyTorch (fp32) restatement of the reference's HAT hot path (models/hat_arch/hat_arch.py — the
file the scripts actually import), used only as the checker in tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg.  The product path (superresolution_def_b200) never imports this package.

Parity pin: every function here is checked against outputs and gradients produced by the UNMODIFIED reference
modules imported from /root/reference in the build container (tools/make_golden.py -> tests/golden/hat_*.pt,
tests/test_oracle_golden.py) and live at the real hot-path dimensions in tests/test_oracle_vs_reference.py.
The reference ships no tests or golden vectors of its own, so these are the pins.  The reference
needs three arithmetic-free symbols of the un-vendored `basicsr` package (registry decorator, to_2tuple,
trunc_normal_); tools/ref_shim.py supplies them so the reference file itself is imported unmodified.

Functions take a flat `state_dict`-style mapping with the reference's own key names.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import torch
import torch.nn.functional as F

from .swinir_oracle import mlp, window_partition, window_reverse, upsample_x4

Tensor = torch.Tensor


# --------------------------------------------------------------------------- index / mask construction
def rpi_sa(ws: int) -> Tensor:
    """(ws^2, ws^2) index into the (2ws-1)^2 table.  Reference: HAT.calculate_rpi_sa, hat_arch.py:882-894."""
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    return (ys[:, None] - ys[None, :] + ws - 1) * (2 * ws - 1) + (xs[:, None] - xs[None, :] + ws - 1)


def rpi_oca(ws: int, overlap_ratio: float = 0.5) -> Tensor:
    """(ws^2, wse^2) index into the (ws+wse-1)^2 OCAB table.  NOTE the reference's offset `ws - wse + 1` leaves
    negative entries, which PyTorch indexing wraps around the table end; reproduced as is.
    Reference: HAT.calculate_rpi_oca, hat_arch.py:896-919."""
    wse = ws + int(overlap_ratio * ws)
    oy, ox = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    ey, ex = torch.meshgrid(torch.arange(wse), torch.arange(wse), indexing="ij")
    oy, ox, ey, ex = oy.reshape(-1), ox.reshape(-1), ey.reshape(-1), ex.reshape(-1)
    ry = ey[None, :] - oy[:, None] + (ws - wse + 1)
    rx = ex[None, :] - ox[:, None] + (ws - wse + 1)
    return ry * (ws + wse - 1) + rx


def shift_mask(h: int, w: int, ws: int, shift: int) -> Tensor:
    """(nW, ws^2, ws^2) additive mask, 0 / -100 (not -inf).  Reference: HAT.calculate_mask, hat_arch.py:921-940."""
    img = torch.zeros(1, h, w, 1)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).reshape(-1, ws * ws)
    diff = mw[:, None, :] - mw[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


# --------------------------------------------------------------------------- modules, functional
def window_attention(x: Tensor, p: Mapping[str, Tensor], prefix: str, num_heads: int, rpi: Tensor,
                     mask: Tensor | None = None) -> Tensor:
    """Reference: WindowAttention.forward(x, rpi, mask), hat_arch.py:165-196."""
    b_, n, c = x.shape
    d = c // num_heads
    qkv = F.linear(x, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"]).reshape(b_, n, 3, num_heads, d)
    qkv = qkv.permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (d ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = p[prefix + "relative_position_bias_table"][rpi.reshape(-1)].reshape(n, n, num_heads).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.reshape(b_ // nw, nw, num_heads, n, n) + mask[None, :, None]
        attn = attn.reshape(-1, num_heads, n, n)
    attn = torch.softmax(attn, dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(b_, n, c)
    return F.linear(out, p[prefix + "proj.weight"], p[prefix + "proj.bias"])


def channel_attention(x: Tensor, p: Mapping[str, Tensor], prefix: str) -> Tensor:
    """x * sigmoid(W2 relu(W1 avgpool(x))).  Reference: ChannelAttention, hat_arch.py:40-58."""
    y = x.mean(dim=(2, 3), keepdim=True)
    y = F.relu(F.conv2d(y, p[prefix + "attention.1.weight"], p[prefix + "attention.1.bias"]))
    y = torch.sigmoid(F.conv2d(y, p[prefix + "attention.3.weight"], p[prefix + "attention.3.bias"]))
    return x * y


def cab(x: Tensor, p: Mapping[str, Tensor], prefix: str) -> Tensor:
    """conv3x3 -> GELU -> conv3x3 -> channel attention on NCHW.  Reference: CAB, hat_arch.py:61-74."""
    y = F.gelu(F.conv2d(x, p[prefix + "cab.0.weight"], p[prefix + "cab.0.bias"], padding=1))
    y = F.conv2d(y, p[prefix + "cab.2.weight"], p[prefix + "cab.2.bias"], padding=1)
    return channel_attention(y, p, prefix + "cab.3.")


def hab(x: Tensor, p: Mapping[str, Tensor], prefix: str, x_size: tuple[int, int], num_heads: int, ws: int, shift: int,
        rpi: Tensor, attn_mask: Tensor | None, conv_scale: float = 0.01, drop=None) -> Tensor:
    """Hybrid attention block.  drop: None (eval mode / rate 0) or (attn_factors, mlp_factors), each [B]: the per-sample
    stochastic-depth factors floor(keep + U)/keep that the reference's drop_path draws (hat_arch.py:11-23).
    Reference: HAB.forward, hat_arch.py:266-309 (window clamp :239-241)."""
    h, w = x_size
    b, _, c = x.shape
    if min(x_size) <= ws:  # constructor-time clamp on input_resolution; callers pass matching sizes
        shift, ws = 0, min(x_size)
    shortcut = x
    y = F.layer_norm(x, (c,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], 1e-5).reshape(b, h, w, c)
    conv_x = cab(y.permute(0, 3, 1, 2), p, prefix + "conv_block.").permute(0, 2, 3, 1).reshape(b, h * w, c)
    if shift > 0:
        ys = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
        mask = attn_mask
    else:
        ys, mask = y, None
    win = window_partition(ys, ws).reshape(-1, ws * ws, c)
    win = window_attention(win, p, prefix + "attn.", num_heads, rpi, mask)
    ys = window_reverse(win.reshape(-1, ws, ws, c), ws, h, w)
    if shift > 0:
        ys = torch.roll(ys, shifts=(shift, shift), dims=(1, 2))
    sa = 1.0 if drop is None else drop[0].reshape(b, 1, 1).to(x.dtype)
    sm = 1.0 if drop is None else drop[1].reshape(b, 1, 1).to(x.dtype)
    x = shortcut + ys.reshape(b, h * w, c) * sa + conv_x * conv_scale
    z = F.layer_norm(x, (c,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], 1e-5)
    return x + mlp(z, p, prefix + "mlp.") * sm


def ocab(x: Tensor, p: Mapping[str, Tensor], prefix: str, x_size: tuple[int, int], num_heads: int, ws: int,
         rpi: Tensor, overlap_ratio: float = 0.5) -> Tensor:
    """Overlapping cross-attention block: queries from ws x ws windows, keys/values from the (1+r)ws halo window,
    zero-padded AFTER the qkv projection, no mask.  Reference: OCAB.forward, hat_arch.py:392-438."""
    h, w = x_size
    b, _, c = x.shape
    wse = int(ws * overlap_ratio) + ws
    d = c // num_heads
    shortcut = x
    y = F.layer_norm(x, (c,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], 1e-5).reshape(b, h, w, c)
    qkv = F.linear(y, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"]).reshape(b, h, w, 3, c).permute(3, 0, 4, 1, 2)
    q = qkv[0].permute(0, 2, 3, 1)
    kv = torch.cat((qkv[1], qkv[2]), dim=1)
    qw = window_partition(q, ws).reshape(-1, ws * ws, c)
    kvw = F.unfold(kv, kernel_size=(wse, wse), stride=ws, padding=(wse - ws) // 2)      # b, 2c*wse*wse, nw
    nw = kvw.shape[-1]
    kvw = kvw.reshape(b, 2, c, wse * wse, nw).permute(1, 0, 4, 3, 2).reshape(2, b * nw, wse * wse, c)
    kw_, vw = kvw[0], kvw[1]
    b_, nq, _ = qw.shape
    n = kw_.shape[1]
    qh = qw.reshape(b_, nq, num_heads, d).permute(0, 2, 1, 3) * (d ** -0.5)
    kh = kw_.reshape(b_, n, num_heads, d).permute(0, 2, 1, 3)
    vh = vw.reshape(b_, n, num_heads, d).permute(0, 2, 1, 3)
    attn = qh @ kh.transpose(-2, -1)
    bias = p[prefix + "relative_position_bias_table"][rpi.reshape(-1)].reshape(ws * ws, wse * wse, num_heads)
    attn = torch.softmax(attn + bias.permute(2, 0, 1).unsqueeze(0), dim=-1)
    out = (attn @ vh).transpose(1, 2).reshape(b_, nq, c)
    out = window_reverse(out.reshape(-1, ws, ws, c), ws, h, w).reshape(b, h * w, c)
    x = F.linear(out, p[prefix + "proj.weight"], p[prefix + "proj.bias"]) + shortcut
    z = F.layer_norm(x, (c,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], 1e-5)
    return x + mlp(z, p, prefix + "mlp.")


def rhag(x: Tensor, p: Mapping[str, Tensor], prefix: str, x_size: tuple[int, int], depth: int, num_heads: int, ws: int,
         rpi_sa_: Tensor, rpi_oca_: Tensor, attn_mask: Tensor, overlap_ratio: float = 0.5) -> Tensor:
    """6 x HAB (odd ones shifted by ws/2) + OCAB, then conv3x3 + residual.
    Reference: AttenBlocks.forward hat_arch.py:526-534, RHAG.forward :618-619."""
    h, w = x_size
    b, _, c = x.shape
    y = x
    for j in range(depth):
        y = hab(y, p, f"{prefix}residual_group.blocks.{j}.", x_size, num_heads, ws, 0 if j % 2 == 0 else ws // 2,
                rpi_sa_, attn_mask)
    y = ocab(y, p, f"{prefix}residual_group.overlap_attn.", x_size, num_heads, ws, rpi_oca_, overlap_ratio)
    img = y.transpose(1, 2).reshape(b, c, h, w)
    img = F.conv2d(img, p[prefix + "conv.weight"], p[prefix + "conv.bias"], padding=1)
    return img.flatten(2).transpose(1, 2) + x


def hat_forward(x: Tensor, p: Mapping[str, Tensor], *, window_size: int, depths: Sequence[int],
                num_heads: Sequence[int], upscale: int = 4, overlap_ratio: float = 0.5) -> Tensor:
    """Whole HAT generator, upsampler='pixelshuffle', in_chans=1 (mean = 0), img_range 1, patch_norm=True, no ape.
    Reference: HAT.forward_features hat_arch.py:950-969 and HAT.forward :971-984."""
    first = F.conv2d(x, p["conv_first.weight"], p["conv_first.bias"], padding=1)
    b, c, h, w = first.shape
    ws = window_size
    idx_sa = p.get("relative_position_index_SA", rpi_sa(ws)).to(x.device)
    idx_oca = p.get("relative_position_index_OCA", rpi_oca(ws, overlap_ratio)).to(x.device)
    mask = shift_mask(h, w, ws, ws // 2).to(x.device)
    t = first.flatten(2).transpose(1, 2)
    t = F.layer_norm(t, (c,), p["patch_embed.norm.weight"], p["patch_embed.norm.bias"], 1e-5)
    for i, depth in enumerate(depths):
        t = rhag(t, p, f"layers.{i}.", (h, w), depth, num_heads[i], ws, idx_sa, idx_oca, mask, overlap_ratio)
    t = F.layer_norm(t, (c,), p["norm.weight"], p["norm.bias"], 1e-5)
    body = t.transpose(1, 2).reshape(b, c, h, w)
    res = F.conv2d(body, p["conv_after_body.weight"], p["conv_after_body.bias"], padding=1) + first
    out = F.leaky_relu(F.conv2d(res, p["conv_before_upsample.0.weight"], p["conv_before_upsample.0.bias"], padding=1),
                       0.01)
    if upscale == 4:
        out = upsample_x4(out, p)
    elif upscale == 2:
        out = F.pixel_shuffle(F.conv2d(out, p["upsample.0.weight"], p["upsample.0.bias"], padding=1), 2)
    else:
        raise ValueError("oracle restates the x2 / x4 PixelShuffle heads the scripts use")
    return F.conv2d(out, p["conv_last.weight"], p["conv_last.bias"], padding=1)


# --------------------------------------------------------------------------- hybrid wrapper (RRDB trunk)
def rdb(x: Tensor, p: Mapping[str, Tensor], prefix: str) -> Tensor:
    """Residual dense block.  Reference: ResidualDenseBlock.forward, models/hybridmodels_hat.py:38-44."""
    def c(i, t):
        return F.conv2d(t, p[f"{prefix}conv{i}.weight"], p[f"{prefix}conv{i}.bias"], padding=1)
    x1 = F.leaky_relu(c(1, x), 0.2)
    x2 = F.leaky_relu(c(2, torch.cat((x, x1), 1)), 0.2)
    x3 = F.leaky_relu(c(3, torch.cat((x, x1, x2), 1)), 0.2)
    x4 = F.leaky_relu(c(4, torch.cat((x, x1, x2, x3), 1)), 0.2)
    x5 = c(5, torch.cat((x, x1, x2, x3, x4), 1))
    return x5 * 0.2 + x


def rrdb(x: Tensor, p: Mapping[str, Tensor], prefix: str) -> Tensor:
    """Reference: RRDBBlock.forward, models/hybridmodels_hat.py:54-58."""
    out = rdb(rdb(rdb(x, p, prefix + "rdb1."), p, prefix + "rdb2."), p, prefix + "rdb3.")
    return out * 0.2 + x


def hybrid_forward(x: Tensor, p: Mapping[str, Tensor], *, window_size: int, depths: Sequence[int],
                   num_heads: Sequence[int], num_rrdb: int) -> Tensor:
    """HAT(x2) -> conv_adapt -> RRDB trunk -> nearest x2 -> convs.
    Reference: HybridHATRealESRGAN.forward, models/hybridmodels_hat.py:117-131."""
    hp = {k[len("hat."):]: v for k, v in p.items() if k.startswith("hat.")}
    hat_out = hat_forward(x, hp, window_size=window_size, depths=depths, num_heads=num_heads, upscale=2)
    feat = F.leaky_relu(F.conv2d(hat_out, p["conv_adapt.weight"], p["conv_adapt.bias"], padding=1), 0.2)
    body = feat
    for i in range(num_rrdb):
        body = rrdb(body, p, f"rrdb_trunk.{i}.")
    feat = feat + F.conv2d(body, p["conv_body.weight"], p["conv_body.bias"], padding=1)
    feat = F.interpolate(feat, scale_factor=2, mode="nearest")
    feat = F.leaky_relu(F.conv2d(feat, p["conv_up.weight"], p["conv_up.bias"], padding=1), 0.2)
    feat = F.leaky_relu(F.conv2d(feat, p["conv_hr.weight"], p["conv_hr.bias"], padding=1), 0.2)
    return F.conv2d(feat, p["conv_last.weight"], p["conv_last.bias"], padding=1)
