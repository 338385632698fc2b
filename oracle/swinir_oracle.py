"""oracle/swinir_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, plain-PyTorch (fp32) restatement of the reference's SwinIR hot path, used only as the
checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The
product path (superresolution_def_b200) never imports this package.

Parity pin: every function here is checked against outputs and gradients produced by the UNMODIFIED
reference modules (models/architecture_swin.py, imported from /root/reference in the build
container) — see tools/make_golden.py, tests/golden/*.pt and tests/test_oracle_golden.py, plus the
live comparison tests/test_oracle_vs_reference.py that runs wherever /root/reference exists.  The
reference ships no tests or golden vectors of its own, so these are the pins.

All functions take a flat `state_dict`-style mapping with the reference's own key names, so a
reference checkpoint drives the oracle directly.
"""
from __future__ import annotations

import math
from typing import Mapping, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- window helpers
def window_partition(x: Tensor, ws: int) -> Tensor:
    """(B,H,W,C) -> (B*nW, ws, ws, C).  Reference: models/architecture_swin.py:27-31."""
    b, h, w, c = x.shape
    t = x.reshape(b, h // ws, ws, w // ws, ws, c)
    return t.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, c)


def window_reverse(win: Tensor, ws: int, h: int, w: int) -> Tensor:
    """Inverse of window_partition.  Reference: models/architecture_swin.py:33-37."""
    nwin = (h // ws) * (w // ws)
    b = win.shape[0] // nwin
    t = win.reshape(b, h // ws, w // ws, ws, ws, -1)
    return t.permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, -1)


def relative_position_index(ws: int) -> Tensor:
    """(ws*ws, ws*ws) int64 index into the (2ws-1)^2 bias table.
    Reference: WindowAttention.__init__, models/architecture_swin.py:51-61."""
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    dy = ys[:, None] - ys[None, :] + (ws - 1)
    dx = xs[:, None] - xs[None, :] + (ws - 1)
    return dy * (2 * ws - 1) + dx


# --------------------------------------------------------------------------- modules, functional
def window_attention(x: Tensor, p: Mapping[str, Tensor], prefix: str, num_heads: int, ws: int,
                     mask: Tensor | None = None) -> Tensor:
    """softmax(scale*Q K^T + B[idx] (+mask)) V, then proj.  x: (B_, N, C).
    Reference: WindowAttention.forward, models/architecture_swin.py:71-96."""
    b_, n, c = x.shape
    d = c // num_heads
    qkv = F.linear(x, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"])
    qkv = qkv.reshape(b_, n, 3, num_heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (d ** -0.5), qkv[1], qkv[2]
    logits = q @ k.transpose(-2, -1)
    table = p[prefix + "relative_position_bias_table"]
    idx = p.get(prefix + "relative_position_index")
    if idx is None:
        idx = relative_position_index(ws).to(table.device)
    bias = table[idx.reshape(-1)].reshape(n, n, num_heads).permute(2, 0, 1)
    logits = logits + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        logits = logits.reshape(b_ // nw, nw, num_heads, n, n) + mask[None, :, None]
        logits = logits.reshape(b_, num_heads, n, n)
    probs = torch.softmax(logits, dim=-1)
    out = (probs @ v).transpose(1, 2).reshape(b_, n, c)
    return F.linear(out, p[prefix + "proj.weight"], p[prefix + "proj.bias"])


def mlp(x: Tensor, p: Mapping[str, Tensor], prefix: str) -> Tensor:
    """fc1 -> exact (erf) GELU -> fc2.  Reference: Mlp.forward, models/architecture_swin.py:19-25."""
    h = F.gelu(F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return F.linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def swin_block(x: Tensor, p: Mapping[str, Tensor], prefix: str, res: tuple[int, int], num_heads: int,
               ws: int, shift: int) -> Tensor:
    """x + Attn(LN1(x)) with cyclic shift, then + Mlp(LN2(.)).  x: (B, H*W, C).
    NOTE the reference never builds a shift mask for SwinIR (mask=None always).
    Reference: SwinTransformerBlock.forward, models/architecture_swin.py:123-151 (clamp at :110-112)."""
    h, w = res
    if min(res) <= ws:
        shift, ws = 0, min(res)
    b, l, c = x.shape
    y = F.layer_norm(x, (c,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], 1e-5).reshape(b, h, w, c)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    win = window_partition(y, ws).reshape(-1, ws * ws, c)
    win = window_attention(win, p, prefix + "attn.", num_heads, ws, mask=None)
    y = window_reverse(win.reshape(-1, ws, ws, c), ws, h, w)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    x = x + y.reshape(b, l, c)
    z = F.layer_norm(x, (c,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], 1e-5)
    return x + mlp(z, p, prefix + "mlp.")


def upsample_x4(x: Tensor, p: Mapping[str, Tensor], prefix: str = "upsample.") -> Tensor:
    """2 x [conv3x3 F->4F, PixelShuffle(2)].  Reference: Upsample, models/architecture_swin.py:175-190."""
    x = F.pixel_shuffle(F.conv2d(x, p[prefix + "0.weight"], p[prefix + "0.bias"], padding=1), 2)
    return F.pixel_shuffle(F.conv2d(x, p[prefix + "2.weight"], p[prefix + "2.bias"], padding=1), 2)


def swinir_forward(x: Tensor, p: Mapping[str, Tensor], *, img_size: int, window_size: int, depths: Sequence[int],
                   num_heads: Sequence[int], upscale: int = 4) -> Tensor:
    """Whole generator.  Reference: SwinIR.forward, models/architecture_swin.py:232-256."""
    assert upscale == 4, "the oracle restates the x4 PixelShuffle head used by the scripts"
    h0, w0 = x.shape[2], x.shape[3]
    ph = (window_size - h0 % window_size) % window_size
    pw = (window_size - w0 % window_size) % window_size
    if ph or pw:
        x = F.pad(x, (0, pw, 0, ph), mode="reflect")
    first = F.conv2d(x, p["conv_first.weight"], p["conv_first.bias"], padding=1)
    b, c, h, w = first.shape
    t = first.flatten(2).transpose(1, 2)
    for i, depth in enumerate(depths):
        for j in range(depth):
            t = swin_block(t, p, f"layers.{i}.{j}.", (img_size, img_size), num_heads[i], window_size,
                           0 if j % 2 == 0 else window_size // 2)
    t = F.layer_norm(t, (c,), p["norm.weight"], p["norm.bias"], 1e-5)
    body = t.transpose(1, 2).reshape(b, c, h, w)
    res = F.conv2d(body, p["conv_after_body.weight"], p["conv_after_body.bias"], padding=1) + first
    out = F.leaky_relu(F.conv2d(res, p["conv_before_upsample.0.weight"], p["conv_before_upsample.0.bias"], padding=1),
                       0.01)
    out = upsample_x4(out, p)
    out = F.conv2d(out, p["conv_last.weight"], p["conv_last.bias"], padding=1)
    return out[:, :, : h0 * upscale, : w0 * upscale]


# --------------------------------------------------------------------------- metrics (the PSNR judge)
def psnr(sr: Tensor, hr: Tensor) -> float:
    """10*log10(1/(mse+1e-8)).  Reference: TrainMetrics, utils/metrics_swin.py:15-26."""
    mse = torch.mean((sr.float() - hr.float()) ** 2).item()
    return 10.0 * math.log10(1.0 / (mse + 1e-8))


# --------------------------------------------------------------------------- parameter construction
def init_state_dict(*, img_size: int, window_size: int, embed_dim: int, depths: Sequence[int],
                    num_heads: Sequence[int], in_chans: int = 1, mlp_ratio: float = 4.0, seed: int = 0):
    """A state_dict with the reference's keys/shapes (SwinIR.__init__, models/architecture_swin.py:193-230) and
    PyTorch-default-like initial values (uniform +-1/sqrt(fan_in) for conv/linear, LN = (1, 0), bias tables
    trunc-normal 0.02).  Used where the oracle must run without a reference checkpoint (bench CPU baseline)."""
    g = torch.Generator().manual_seed(seed)
    sd: dict[str, Tensor] = {}

    def dense(name, *shape):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        bound = 1.0 / math.sqrt(fan_in)
        sd[name + ".weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(shape[0], generator=g) * 2 - 1) * bound

    def ln(name, c):
        sd[name + ".weight"] = torch.ones(c)
        sd[name + ".bias"] = torch.zeros(c)

    c, hidden = embed_dim, int(embed_dim * mlp_ratio)
    dense("conv_first", c, in_chans, 3, 3)
    for i, depth in enumerate(depths):
        for j in range(depth):
            p = f"layers.{i}.{j}."
            ln(p + "norm1", c)
            sd[p + "attn.relative_position_bias_table"] = (
                torch.randn((2 * window_size - 1) ** 2, num_heads[i], generator=g) * 0.02).clamp(-0.04, 0.04)
            sd[p + "attn.relative_position_index"] = relative_position_index(window_size)
            dense(p + "attn.qkv", 3 * c, c)
            dense(p + "attn.proj", c, c)
            ln(p + "norm2", c)
            dense(p + "mlp.fc1", hidden, c)
            dense(p + "mlp.fc2", c, hidden)
    ln("norm", c)
    dense("conv_after_body", c, c, 3, 3)
    dense("conv_before_upsample.0", 64, c, 3, 3)
    dense("upsample.0", 256, 64, 3, 3)
    dense("upsample.2", 256, 64, 3, 3)
    dense("conv_last", in_chans, 64, 3, 3)
    return sd
