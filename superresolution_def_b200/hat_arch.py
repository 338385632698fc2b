"""Drop-in mirror of the reference's `models/hat_arch/hat_arch.py` module interface on top of libsrk.

Same class names, constructor arguments, forward signatures, parameter/buffer names, shapes and registration order
as the reference (state_dict round-trips with strict=True; EMA / optimizer code walking `named_parameters()` sees the
same sequence) — but the forwards run the sm_100a kernels:

  HAB.forward(x, x_size, rpi_sa, attn_mask)  -> CAB convs (implicit GEMM) + channel attention + srk_hat_block_fwd/bwd
                                               (16x16 window attention with in-kernel shift mask)   (reference :266-309)
  OCAB.forward(x, x_size, rpi)               -> srk_hat_block_fwd/bwd with the 24x24 halo key window (:392-438)
  RHAG.forward / HAT.forward                 -> block chain + implicit-GEMM convs + fused PixelShuffle tail (:618-619,:950-984)

`rpi_*` and `attn_mask` arguments are accepted for signature parity; the kernels derive the relative-position
indices and the 0/-100 shift mask from coordinates (they are pure functions of window size / shift / image size), so
a caller-supplied mask that differs from HAT.calculate_mask's is not supported.
Compute dtype is bf16 with fp32 accumulation; there is no CPU path.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _capi as capi
from . import swin_engine as eng
from . import hat_engine as heng
from . import conv_engine as cv
from .architecture_swin import Mlp, Upsample, window_partition, window_reverse  # same code in both reference files


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def drop_path(x, drop_prob: float = 0., training: bool = False):
    """Per-sample stochastic depth (reference :11-23); identity when not training or rate 0."""
    if drop_prob == 0. or not training:
        return x
    keep = 1 - drop_prob
    mask = (keep + torch.rand((x.shape[0],) + (1,) * (x.ndim - 1), dtype=x.dtype, device=x.device)).floor_()
    return x.div(keep) * mask


class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)


class ChannelAttention(nn.Module):
    def __init__(self, num_feat, squeeze_factor=16):
        super().__init__()
        self.attention = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(num_feat, num_feat // squeeze_factor, 1, padding=0),
                                       nn.ReLU(inplace=True), nn.Conv2d(num_feat // squeeze_factor, num_feat, 1, padding=0),
                                       nn.Sigmoid())

    def forward(self, x):
        """x: (B, C, H, W) -> x * attention(x) (:56-58).  Inside CAB / HAB the same kernels run fused on token-major data."""
        if not x.is_cuda:
            raise capi.SrkError("libsrk ChannelAttention runs on CUDA only")
        a = self.attention
        return heng.ChannelAttentionFunction.apply(x, a[1].weight, a[1].bias, a[3].weight, a[3].bias)


class CAB(nn.Module):
    def __init__(self, num_feat, compress_ratio=3, squeeze_factor=30):
        super().__init__()
        eng.track_weight_changes(self)
        self.cab = nn.Sequential(nn.Conv2d(num_feat, num_feat // compress_ratio, 3, 1, 1), nn.GELU(),
                                 nn.Conv2d(num_feat // compress_ratio, num_feat, 3, 1, 1),
                                 ChannelAttention(num_feat, squeeze_factor))

    def forward(self, x):
        """x: (B, C, H, W) as in the reference (:73); inside HAB.forward the same kernels run on token-major data."""
        if not x.is_cuda:
            raise capi.SrkError("libsrk CAB runs on CUDA only")
        ca = self.cab[3].attention
        return heng.CabFunction.apply(x, self.cab[0].weight, self.cab[0].bias, self.cab[2].weight, self.cab[2].bias,
                                      ca[1].weight, ca[1].bias, ca[3].weight, ca[3].bias)


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
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)
        if not qkv_bias or qk_scale is not None or attn_drop != 0. or proj_drop != 0.:
            raise capi.SrkError("libsrk WindowAttention: qkv_bias=True, default scale, no dropout (reference usage)")

    def forward(self, x, rpi=None, mask=None):
        """x: (num_windows*b, 256, c) already-partitioned 16x16 windows (reference :165).  `rpi` is accepted for signature
        parity (the index is a pure function of the window size).  An explicit `mask` tensor is not supported here: the
        shift mask only exists inside HAB.forward, where the kernel derives it from coordinates."""
        if mask is not None:
            raise capi.SrkError("libsrk WindowAttention.forward: pass mask=None; the shifted-window mask is applied "
                                "inside HAB.forward (computed in-kernel from coordinates)")
        if not x.is_cuda:
            raise capi.SrkError("libsrk WindowAttention runs on CUDA only")
        return eng.window_attention_forward(x, self.window_size[0], self.num_heads, self.relative_position_bias_table,
                                            self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias)


class HAB(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, compress_ratio=3, squeeze_factor=30,
                 conv_scale=0.01, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        eng.track_weight_changes(self)
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        if min(self.input_resolution) <= self.window_size:
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, 'shift_size must in 0-window_size'
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=to_2tuple(self.window_size), num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.conv_scale = conv_scale
        self.conv_block = CAB(num_feat=dim, compress_ratio=compress_ratio, squeeze_factor=squeeze_factor)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def block_cfg(self) -> eng.BlockCfg:
        return heng.hat_block_cfg(self.dim, self.num_heads, self.mlp.fc1.out_features, self.window_size)

    def _drop_factors(self, b, device):
        """Per-sample stochastic-depth factors of the two residual branches (reference drop_path :11-23, applied at
        :306-307): 0 or 1/keep_prob, drawn in the reference's order (attention branch first, then the MLP branch)."""
        p = getattr(self.drop_path, "drop_prob", None) or 0.
        if not self.training or p == 0.:
            return None
        keep = 1.0 - p
        draw = lambda: ((keep + torch.rand(b, device=device)).floor_() / keep).float().contiguous()  # noqa: E731
        return (draw(), draw())

    def run(self, t, xn, stats, geom, next_norm, drop=None):
        """Token-major entry used by the model-level forward: (x, xn, stats) -> (x_out, xn_out, stats_out).
        drop: optional explicit (attn_factors, mlp_factors) [B] tensors (tests); default: drawn when training."""
        if drop is None:
            drop = self._drop_factors(geom[0], t.device)
        return heng.HatBlockFunction.apply(t, xn, stats, self.block_cfg(), geom, "hab", self.shift_size,
                                           self.conv_scale, drop, *heng.hab_params_of(self), next_norm[0], next_norm[1])

    def forward(self, x, x_size, rpi_sa=None, attn_mask=None, drop=None):
        h, w = x_size
        b, _, c = x.shape
        cfg = self.block_cfg()
        tok = eng.pack_tokens(x, cfg.Cp)
        xn, stats = eng.layernorm_tokens(tok, self.norm1.weight, self.norm1.bias, c)
        out, _, _ = self.run(tok, xn, stats, (b, h, w), eng.identity_norm(c, x.device), drop=drop)
        return eng.unpack_tokens(out, b, c, x.dtype)


class OCAB(nn.Module):
    def __init__(self, dim, input_resolution, window_size, overlap_ratio, num_heads, qkv_bias=True, qk_scale=None,
                 mlp_ratio=2, norm_layer=nn.LayerNorm):
        super().__init__()
        eng.track_weight_changes(self)
        self.dim = dim
        self.input_resolution = input_resolution
        self.window_size = window_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.overlap_win_size = int(window_size * overlap_ratio) + window_size
        self.norm1 = norm_layer(dim)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.unfold = nn.Unfold(kernel_size=(self.overlap_win_size, self.overlap_win_size), stride=window_size,
                                padding=(self.overlap_win_size - window_size) // 2)
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((window_size + self.overlap_win_size - 1) * (window_size + self.overlap_win_size - 1), num_heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)
        self.proj = nn.Linear(dim, dim)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=nn.GELU)
        if (window_size, self.overlap_win_size) not in ((16, 24), (8, 12)) or not qkv_bias or qk_scale is not None:
            raise capi.SrkError("libsrk OCAB: window 16 or 8 with overlap_ratio 0.5 (24x24 / 12x12 key window), qkv_bias=True")

    def block_cfg(self) -> eng.BlockCfg:
        return heng.hat_block_cfg(self.dim, self.num_heads, self.mlp.fc1.out_features, self.window_size)

    def run(self, t, xn, stats, geom, next_norm):
        return heng.HatBlockFunction.apply(t, xn, stats, self.block_cfg(), geom, "ocab", 0, 0.0, None,
                                           *heng.ocab_params_of(self), next_norm[0], next_norm[1])

    def forward(self, x, x_size, rpi=None):
        h, w = x_size
        b, _, c = x.shape
        cfg = self.block_cfg()
        tok = eng.pack_tokens(x, cfg.Cp)
        xn, stats = eng.layernorm_tokens(tok, self.norm1.weight, self.norm1.bias, c)
        out, _, _ = self.run(tok, xn, stats, (b, h, w), eng.identity_norm(c, x.device))
        return eng.unpack_tokens(out, b, c, x.dtype)


class AttenBlocks(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, compress_ratio, squeeze_factor, conv_scale,
                 overlap_ratio, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            HAB(dim=dim, input_resolution=input_resolution, num_heads=num_heads, window_size=window_size,
                shift_size=0 if (i % 2 == 0) else window_size // 2, compress_ratio=compress_ratio,
                squeeze_factor=squeeze_factor, conv_scale=conv_scale, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path, norm_layer=norm_layer)
            for i in range(depth)])
        self.overlap_attn = OCAB(dim=dim, input_resolution=input_resolution, window_size=window_size,
                                 overlap_ratio=overlap_ratio, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                 mlp_ratio=mlp_ratio, norm_layer=norm_layer)
        if downsample is not None:
            raise capi.SrkError("PatchMerging downsample is never instantiated by the reference (downsample=None, :849)")
        self.downsample = None

    def run(self, t, geom):
        """Token-major: t [T, Cp] -> AttenBlocks output [T, Cp]."""
        C = self.dim
        blocks = list(self.blocks)
        xn, stats = eng.layernorm_tokens(t, blocks[0].norm1.weight, blocks[0].norm1.bias, C)
        for i, blk in enumerate(blocks):
            nxt = blocks[i + 1].norm1 if i + 1 < len(blocks) else self.overlap_attn.norm1
            t, xn, stats = blk.run(t, xn, stats, geom, (nxt.weight, nxt.bias))
        t, _, _ = self.overlap_attn.run(t, xn, stats, geom, eng.identity_norm(C, t.device))
        return t

    def forward(self, x, x_size, params=None):
        b, _, c = x.shape
        cfg = self.blocks[0].block_cfg()
        t = self.run(eng.pack_tokens(x, cfg.Cp), (b, x_size[0], x_size[1]))
        return eng.unpack_tokens(t, b, c, x.dtype)


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.patches_resolution = [img_size[0] // patch_size[0], img_size[1] // patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        x = x.flatten(2).transpose(1, 2)
        return self.norm(x) if self.norm is not None else x


class PatchUnEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.patches_resolution = [img_size[0] // patch_size[0], img_size[1] // patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim

    def forward(self, x, x_size):
        return x.transpose(1, 2).contiguous().view(x.shape[0], self.embed_dim, x_size[0], x_size[1])


class RHAG(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, compress_ratio, squeeze_factor, conv_scale,
                 overlap_ratio, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False, img_size=224, patch_size=4,
                 resi_connection='1conv'):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.residual_group = AttenBlocks(dim=dim, input_resolution=input_resolution, depth=depth, num_heads=num_heads,
                                          window_size=window_size, compress_ratio=compress_ratio,
                                          squeeze_factor=squeeze_factor, conv_scale=conv_scale, overlap_ratio=overlap_ratio,
                                          mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop,
                                          attn_drop=attn_drop, drop_path=drop_path, norm_layer=norm_layer,
                                          downsample=downsample, use_checkpoint=use_checkpoint)
        if resi_connection != '1conv':
            raise capi.SrkError("libsrk RHAG implements resi_connection='1conv' (the reference scripts' setting)")
        self.conv = nn.Conv2d(dim, dim, 3, 1, 1)
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=0, embed_dim=dim, norm_layer=None)
        self.patch_unembed = PatchUnEmbed(img_size=img_size, patch_size=patch_size, in_chans=0, embed_dim=dim,
                                          norm_layer=None)

    def run(self, t, geom):
        y = self.residual_group.run(t, geom)
        return heng.RhagConvFunction.apply(y, t, geom, self.dim, self.conv.weight, self.conv.bias)

    def forward(self, x, x_size, params=None):
        b, _, c = x.shape
        cfg = self.residual_group.blocks[0].block_cfg()
        t = self.run(eng.pack_tokens(x, cfg.Cp), (b, x_size[0], x_size[1]))
        return eng.unpack_tokens(t, b, c, x.dtype)


class HAT(nn.Module):
    def __init__(self, img_size=64, patch_size=1, in_chans=3, embed_dim=96, depths=(6, 6, 6, 6), num_heads=(6, 6, 6, 6),
                 window_size=7, compress_ratio=3, squeeze_factor=30, conv_scale=0.01, overlap_ratio=0.5, mlp_ratio=4.,
                 qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1,
                 norm_layer=nn.LayerNorm, ape=False, patch_norm=True, use_checkpoint=False, upscale=2, img_range=1.,
                 upsampler='', resi_connection='1conv', **kwargs):
        super().__init__()
        eng.track_weight_changes(self)
        self.window_size = window_size
        self.shift_size = window_size // 2
        self.overlap_ratio = overlap_ratio
        num_feat = 64
        self.img_range = img_range
        if in_chans != 1 or upsampler != 'pixelshuffle' or ape or not patch_norm or patch_size != 1 or drop_rate != 0.:
            raise capi.SrkError("libsrk HAT implements the reference scripts' configuration: in_chans=1, "
                                "upsampler='pixelshuffle', patch_size=1, patch_norm=True, ape=False, drop_rate=0")
        self.mean = torch.zeros(1, 1, 1, 1)
        self.upscale = upscale
        self.upsampler = upsampler
        self.register_buffer('relative_position_index_SA', self.calculate_rpi_sa())
        self.register_buffer('relative_position_index_OCA', self.calculate_rpi_oca())
        self.conv_first = nn.Conv2d(in_chans, embed_dim, 3, 1, 1)
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.ape = ape
        self.patch_norm = patch_norm
        self.num_features = embed_dim
        self.mlp_ratio = mlp_ratio
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=embed_dim, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None)
        self.patches_resolution = self.patch_embed.patches_resolution
        self.patch_unembed = PatchUnEmbed(img_size=img_size, patch_size=patch_size, in_chans=embed_dim, embed_dim=embed_dim,
                                          norm_layer=norm_layer if self.patch_norm else None)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(RHAG(dim=embed_dim, input_resolution=tuple(self.patches_resolution), depth=depths[i],
                                    num_heads=num_heads[i], window_size=window_size, compress_ratio=compress_ratio,
                                    squeeze_factor=squeeze_factor, conv_scale=conv_scale, overlap_ratio=overlap_ratio,
                                    mlp_ratio=self.mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                                    attn_drop=attn_drop_rate, drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])],
                                    norm_layer=norm_layer, downsample=None, use_checkpoint=use_checkpoint,
                                    img_size=img_size, patch_size=patch_size, resi_connection=resi_connection))
        self.norm = norm_layer(self.num_features)
        self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, num_feat, 3, 1, 1), nn.LeakyReLU(inplace=True))
        self.upsample = Upsample(upscale, num_feat)
        self.conv_last = nn.Conv2d(num_feat, in_chans, 3, 1, 1)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def calculate_rpi_sa(self):
        ws = self.window_size
        ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
        ys, xs = ys.flatten(), xs.flatten()
        return (ys[:, None] - ys[None, :] + ws - 1) * (2 * ws - 1) + (xs[:, None] - xs[None, :] + ws - 1)

    def calculate_rpi_oca(self):
        ws = self.window_size
        wse = ws + int(self.overlap_ratio * ws)
        oy, ox = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
        ey, ex = torch.meshgrid(torch.arange(wse), torch.arange(wse), indexing="ij")
        oy, ox, ey, ex = oy.flatten(), ox.flatten(), ey.flatten(), ex.flatten()
        ry = ey[None, :] - oy[:, None] + ws - wse + 1
        rx = ex[None, :] - ox[:, None] + ws - wse + 1
        return ry * (ws + wse - 1) + rx

    def calculate_mask(self, x_size):
        """Kept for API parity (reference :921-940); the kernels compute the same 0/-100 mask from coordinates."""
        h, w = x_size
        img_mask = torch.zeros((1, h, w, 1))
        cnt = 0
        for hs in (slice(0, -self.window_size), slice(-self.window_size, -self.shift_size), slice(-self.shift_size, None)):
            for ws_ in (slice(0, -self.window_size), slice(-self.window_size, -self.shift_size), slice(-self.shift_size, None)):
                img_mask[:, hs, ws_, :] = cnt
                cnt += 1
        mw = window_partition(img_mask, self.window_size).reshape(-1, self.window_size * self.window_size)
        m = mw.unsqueeze(1) - mw.unsqueeze(2)
        return m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0))

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'absolute_pos_embed'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table'}

    def forward(self, x):
        eng.check_precision(x)
        B, _, H, W = x.shape
        ws = self.window_size
        if H % ws or W % ws:
            raise capi.SrkError(f"input {H}x{W} is not a multiple of the window size {ws} (the reference fails too)")
        C = self.embed_dim
        cfg = self.layers[0].residual_group.blocks[0].block_cfg()
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
        geom = (B, H, W)
        # (x - mean) * img_range is the identity for in_chans == 1, img_range == 1 (reference :972-973)
        xin = x if self.img_range == 1. else x * self.img_range
        first = cv.conv3x3_tokens(xin, self.conv_first.weight, self.conv_first.bias, cfg.Cp)
        t = heng.LayerNormTokensFunction.apply(first, self.patch_embed.norm.weight, self.patch_embed.norm.bias, C, -1)
        for layer in self.layers:
            t = layer.run(t, geom)
        body = heng.LayerNormTokensFunction.apply(t, self.norm.weight, self.norm.bias, C, C)
        if self.upscale not in (2, 4):
            raise capi.SrkError("libsrk HAT tail implements upscale 2 / 4 (one / two fused conv+PixelShuffle stages)")
        out = cv.swinir_tail(body, first, geom, C, self.conv_after_body, self.conv_before_upsample[0], self.upsample,
                             self.conv_last)
        if self.img_range != 1.:
            out = out / self.img_range
        return out.to(out_dtype)
