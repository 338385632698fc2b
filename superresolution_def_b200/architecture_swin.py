"""Drop-in mirror of the reference's `models/architecture_swin.py` module interface on top of libsrk.

Same class names, constructor arguments, forward signatures, parameter/buffer names, shapes and registration
order as the reference (so `state_dict()` round-trips with strict=True and EMA / optimizer code that walks
`named_parameters()` sees the same sequence) — but every forward runs the sm_100a kernels:

  SwinTransformerBlock.forward  -> srk_swin_block_fwd/bwd      (reference: models/architecture_swin.py:123-151)
  WindowAttention.forward       -> tcgen05 qkv/proj GEMMs + srk_win_attn_fwd/bwd          (:71-96)
  Mlp.forward                   -> tcgen05 fc1(+GELU) / fc2 GEMMs                          (:19-25)
  SwinIR.forward                -> conv head/tail + 6 fused block stacks + fused final norm (:232-256)

Compute dtype is bf16 with fp32 accumulation regardless of the autocast state (the reference scripts train
under autocast); inputs/outputs keep the caller's dtype.  There is no CPU path: calling these modules without
CUDA tensors raises.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _capi as capi
from . import swin_engine as eng
from . import conv_engine as cv


def window_partition(x, window_size):
    """(b,h,w,c) -> (b*nW, ws, ws, c).  Kept for API parity (reference :27-31); the fused kernels never call it."""
    b, h, w, c = x.shape
    return (x.reshape(b, h // window_size, window_size, w // window_size, window_size, c)
            .transpose(2, 3).reshape(-1, window_size, window_size, c))


def window_reverse(windows, window_size, h, w):
    """Inverse of window_partition (reference :33-37)."""
    b = windows.shape[0] // ((h // window_size) * (w // window_size))
    return (windows.reshape(b, h // window_size, w // window_size, window_size, window_size, -1)
            .transpose(2, 3).reshape(b, h, w, -1))


def _rel_pos_index(wh: int, ww: int) -> torch.Tensor:
    ys, xs = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    return (ys[:, None] - ys[None, :] + wh - 1) * (2 * ww - 1) + (xs[:, None] - xs[None, :] + ww - 1)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        eng.track_weight_changes(self)
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        if not isinstance(self.act, nn.GELU) or drop != 0.:
            raise capi.SrkError("libsrk Mlp implements exact-erf GELU without dropout (the reference's only use)")

    def forward(self, x):
        return eng.mlp_forward(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        eng.track_weight_changes(self)
        self.dim = dim
        self.window_size = window_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads))
        self.register_buffer("relative_position_index", _rel_pos_index(window_size[0], window_size[1]))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)
        if not qkv_bias or qk_scale is not None or attn_drop != 0. or proj_drop != 0.:
            raise capi.SrkError("libsrk WindowAttention: qkv_bias=True, default scale, no dropout (reference usage)")

    def forward(self, x, mask=None):
        if mask is not None:
            raise capi.SrkError("SwinIR never passes a mask (reference :138); the masked variant lives in the HAT ops")
        return eng.window_attention_forward(x, self.window_size[0], self.num_heads, self.relative_position_bias_table,
                                            self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        eng.track_weight_changes(self)
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        if min(self.input_resolution) <= self.window_size:  # reference :110-112
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=(self.window_size, self.window_size), num_heads=num_heads,
                                    qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def block_cfg(self) -> eng.BlockCfg:
        return eng.BlockCfg.for_model(self.dim, self.num_heads, self.mlp.fc1.out_features, self.window_size)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        if L != H * W:
            raise capi.SrkError(f"token count {L} does not match input_resolution {H}x{W} (reference :128 would fail)")
        cfg = self.block_cfg()
        tok = eng.pack_tokens(x, cfg.Cp)
        xn, stats = eng.layernorm_tokens(tok, self.norm1.weight, self.norm1.bias, C)
        ident_w, ident_b = eng.identity_norm(C, x.device)
        out, _, _ = eng.SwinStackFunction.apply(tok, xn, stats, cfg, (B, H, W), (self.shift_size,),
                                                *eng.block_params_of(self), ident_w, ident_b)
        return eng.unpack_tokens(out, B, C, x.dtype)


class PatchEmbed(nn.Module):
    def __init__(self, embed_dim=96, norm_layer=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.norm = norm_layer(embed_dim) if norm_layer else None

    def forward(self, x):
        x = x.flatten(2).transpose(1, 2)
        return self.norm(x) if self.norm else x


class PatchUnEmbed(nn.Module):
    def __init__(self, embed_dim=96):
        super().__init__()
        self.embed_dim = embed_dim

    def forward(self, x, x_size):
        B, HW, C = x.shape
        return x.transpose(1, 2).view(B, self.embed_dim, x_size[0], x_size[1])


class Upsample(nn.Sequential):
    """conv3x3(F -> 4F) + PixelShuffle(2), log2(scale) times (reference :175-190).  When called through
    SwinIR.forward the pair runs as one implicit-GEMM kernel whose epilogue stores in shuffled order."""

    def __init__(self, scale, num_feat):
        m = []
        if (scale & (scale - 1)) == 0:
            for _ in range(int(math.log(scale, 2))):
                m.append(nn.Conv2d(num_feat, 4 * num_feat, 3, 1, 1))
                m.append(nn.PixelShuffle(2))
        elif scale == 3:
            m.append(nn.Conv2d(num_feat, 9 * num_feat, 3, 1, 1))
            m.append(nn.PixelShuffle(3))
        else:
            raise ValueError(f"scale {scale} is not supported (powers of 2, or 3)")
        super().__init__(*m)


class SwinIR(nn.Module):
    def __init__(self, img_size=64, in_chans=1, embed_dim=96, depths=[6, 6, 6], num_heads=[6, 6, 6], window_size=7,
                 upscale=2, **kwargs):
        super().__init__()
        eng.track_weight_changes(self)
        # NOTE: like the reference (:193-194) extra kwargs (mlp_ratio, img_range, upsampler, ...) are swallowed:
        # the effective MLP ratio is the block default 4.0.
        self.upscale = upscale
        self.window_size = window_size
        self.embed_dim = embed_dim
        self.img_size = img_size
        self.conv_first = nn.Conv2d(in_chans, embed_dim, 3, 1, 1)
        self.patch_embed = PatchEmbed(embed_dim=embed_dim)
        self.patch_unembed = PatchUnEmbed(embed_dim=embed_dim)
        self.layers = nn.ModuleList()
        for i in range(len(depths)):
            self.layers.append(nn.ModuleList([
                SwinTransformerBlock(dim=embed_dim, input_resolution=(img_size, img_size), num_heads=num_heads[i],
                                     window_size=window_size, shift_size=0 if (j % 2 == 0) else window_size // 2)
                for j in range(depths[i])]))
        self.norm = nn.LayerNorm(embed_dim)
        self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, 64, 3, 1, 1), nn.LeakyReLU(inplace=True))
        self.upsample = Upsample(upscale, 64)
        self.conv_last = nn.Conv2d(64, in_chans, 3, 1, 1)

    def forward(self, x):
        eng.check_precision(x)
        H0, W0 = x.shape[2], x.shape[3]
        ws = self.window_size
        pad_h, pad_w = (ws - H0 % ws) % ws, (ws - W0 % ws) % ws
        if pad_h or pad_w:
            x = F.pad(x, (0, pad_w, 0, pad_h), mode="reflect")
        B, _, H, W = x.shape
        blocks = [blk for layer in self.layers for blk in layer]
        if (H, W) != tuple(blocks[0].input_resolution):
            raise capi.SrkError(f"padded input {H}x{W} != img_size {blocks[0].input_resolution} (reference :128 fails too)")
        cfg = blocks[0].block_cfg()
        C = self.embed_dim
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype

        # conv_first -> token-major bf16 residual stream [T, Cp]
        first = cv.conv3x3_tokens(x, self.conv_first.weight, self.conv_first.bias, cfg.Cp)          # [T, Cp]
        xn, stats = eng.layernorm_tokens(first, blocks[0].norm1.weight, blocks[0].norm1.bias, C)
        t = first
        idx = 0
        for li, layer in enumerate(self.layers):
            params = []
            for blk in layer:
                params += eng.block_params_of(blk)
            nxt = blocks[idx + len(layer)].norm1 if idx + len(layer) < len(blocks) else self.norm
            shifts = tuple(blk.shift_size for blk in layer)
            t, xn, stats = eng.SwinStackFunction.apply(t, xn, stats, cfg, (B, H, W), shifts, *params, nxt.weight,
                                                       nxt.bias)
            idx += len(layer)
        body = eng.FusedNormOutput.apply(t, xn, stats, self.norm.weight, self.norm.bias, C)          # [T, Cp]
        out = cv.swinir_tail(body, first, (B, H, W), C, self.conv_after_body, self.conv_before_upsample[0],
                             self.upsample, self.conv_last)
        out = out[:, :, :H0 * self.upscale, :W0 * self.upscale]
        return out.to(out_dtype)
