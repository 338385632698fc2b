"""Host-side engine for HAT's blocks on libsrk: HAB (window attention 16x16 with shift mask + CAB branch), OCAB
(overlapping cross-attention), the RHAG tail convolution and differentiable token LayerNorm.

Same HBM layout and (x, xn, stats) hand-over protocol as swin_engine: the residual stream is token-major bf16
[T, Cp]; every block receives its own LayerNorm-1 output `xn` (+ row statistics) from the epilogue of the previous
block's fc2 GEMM and returns the next one.  A block is one autograd node; PyTorch provides memory, streams and the
autograd graph only — all arithmetic is in the library and there is no fallback path.

Reference being replaced (models/hat_arch/hat_arch.py): HAB.forward :266-309, CAB/ChannelAttention :40-74,
WindowAttention.forward :165-196, OCAB.forward :392-438, RHAG.forward :618-619, calculate_mask :921-940.
"""
from __future__ import annotations

import torch

from . import _capi as capi
from . import swin_engine as eng
from . import conv_engine as cv

BF16 = torch.bfloat16
HAB_KEYS = ("norm1.weight", "norm1.bias", "attn.relative_position_bias_table", "attn.qkv.weight", "attn.qkv.bias",
            "attn.proj.weight", "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias",
            "mlp.fc2.weight", "mlp.fc2.bias")
CAB_KEYS = ("conv_block.cab.0.weight", "conv_block.cab.0.bias", "conv_block.cab.2.weight", "conv_block.cab.2.bias",
            "conv_block.cab.3.attention.1.weight", "conv_block.cab.3.attention.1.bias",
            "conv_block.cab.3.attention.3.weight", "conv_block.cab.3.attention.3.bias")
OCAB_KEYS = ("norm1.weight", "norm1.bias", "relative_position_bias_table", "qkv.weight", "qkv.bias", "proj.weight",
             "proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight",
             "mlp.fc2.bias")


def hat_block_cfg(C: int, heads: int, hidden: int, ws: int) -> eng.BlockCfg:
    if not (C < 192 and heads * 32 == 192 and C % heads == 0 and C // heads < 32 and ws in (8, 16)):
        raise capi.SrkError(f"libsrk HAT kernels are specialised for heads=6, head_dim<32, embed_dim<192, window 8 or 16; "
                            f"got C={C} heads={heads} ws={ws}")
    return eng.BlockCfg(C=C, heads=heads, hidden=hidden, ws=ws, Cp=192, ds=32, Hp=((hidden + 1 + 255) // 256) * 256)


def hab_params_of(blk) -> list[torch.Tensor]:
    sd = dict(blk.named_parameters())
    return [sd[k] for k in HAB_KEYS + CAB_KEYS]


def ocab_params_of(blk) -> list[torch.Tensor]:
    sd = dict(blk.named_parameters())
    return [sd[k] for k in OCAB_KEYS]


_attn_ws_cache: dict = {}


def _attn_ws(geom, mode, heads, device):
    n = capi.attn16_bwd_ws_bytes(geom, mode, heads)
    k = (mode, str(device))
    t = _attn_ws_cache.get(k)
    if t is None or t.numel() < n:
        capi.retire(t)   # grow-only: a captured graph may still address the old buffer
        t = _attn_ws_cache[k] = torch.empty(n, device=device, dtype=torch.uint8)
    return t


class HatBlockFunction(torch.autograd.Function):
    """(x, xn, stats) -> (x_out, xn_out, stats_out) through one HAB (kind 'hab') or OCAB (kind 'ocab').

    tensors = 13 block parameters (HAB_KEYS / OCAB_KEYS order) [+ 8 CAB parameters for 'hab'] + (next_norm_weight,
    next_norm_bias): the affine of the LayerNorm that consumes x_out (its gradient belongs to the consumer).
    The gradient returned for `x` is the complete dL/dx; `xn`/`stats` are derived data without gradient."""

    @staticmethod
    def forward(ctx, x, xn, stats, cfg, geom, kind, shift, conv_scale, drop, *tensors):
        """drop: None, or (attn_factors, mlp_factors): fp32 [B] stochastic-depth factors (0 or 1/keep_prob) per sample."""
        B, H, W = geom
        T = B * H * W
        dev = x.device
        for t in tensors:
            eng._check_param(t)
        assert x.dtype == BF16 and x.shape == (T, cfg.Cp) and x.is_contiguous() and xn.is_contiguous()
        need_grad = any(ctx.needs_input_grad)
        dims = cfg.dims()
        params = [t.detach() for t in tensors[:13]]
        nw, nb = tensors[-2].detach(), tensors[-1].detach()
        pdict = dict(zip(capi.PARAM_NAMES, params))
        weights = eng._weights_for(cfg, params, refresh=True)
        acts = eng._alloc_acts(cfg, T, dev)
        acts.update(x_in=x, xn1=xn, stats1=stats)
        g = capi.SrkGeom(B, H, W, cfg.ws, shift)
        lse = torch.empty(cfg.heads, T, device=dev, dtype=torch.float32)
        extra = None
        if kind == "hab":
            c1w, c1b, c2w, c2b, s1w, s1b, s2w, s2b = [t.detach() for t in tensors[13:21]]
            Cm, S = c1w.shape[0], s1w.shape[0]
            Cm_p = cv._pad64(Cm)
            wf1, _, bp1 = cv.conv_weights(c1w, c1b, Cm_p, cfg.Cp)
            wf2, _, bp2 = cv.conv_weights(c2w, c2b, cfg.Cp, Cm_p)
            c1 = torch.empty(T, Cm_p, device=dev, dtype=BF16)
            dc1 = torch.empty_like(c1)
            capi.conv3x3_igemm(capi.CEPI_BIAS_GELU, B, H, W, cfg.Cp, Cm_p, Cm, xn, wf1, bp1, c1, y2=dc1)
            c2 = torch.empty(T, cfg.Cp, device=dev, dtype=BF16)
            capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, Cm_p, cfg.Cp, cfg.C, c1, wf2, bp2, c2)
            pool = torch.empty(B, cfg.C, device=dev, dtype=torch.float32)
            hidden = torch.empty(B, S, device=dev, dtype=torch.float32)
            scale = torch.empty(B, cfg.C, device=dev, dtype=torch.float32)
            xr = torch.empty_like(x)
            capi.cab_se_fwd(c2, x, B, H * W, cfg.C, S, s1w, s1b, s2w, s2b, float(conv_scale), pool, hidden, scale, xr)
            capi.hat_block_fwd(dims, g, weights, pdict, nw, nb, acts, capi.ATTN_SELF, xr, lse, drop=drop or (None, None))
            extra = (c1, dc1, c2, pool, hidden, scale, Cm, Cm_p, S)
        elif kind == "ocab":
            capi.hat_block_fwd(dims, g, weights, pdict, nw, nb, acts, capi.ATTN_OCA, x, lse)
        else:
            raise capi.SrkError(f"unknown HAT block kind {kind!r}")
        outs = (acts["x_out"], acts["xn_out"], acts["stats_out"])
        if need_grad:
            # the node's own outputs are not needed by its backward: keeping them on ctx would close an
            # output -> grad_fn -> ctx -> output reference cycle
            keep = {k: v for k, v in acts.items() if k not in ("x_out", "xn_out", "stats_out")}
            ctx.saved = (keep, lse, extra, weights)
            ctx.meta = (cfg, geom, kind, shift, float(conv_scale), drop)
            ctx.params = tensors
        ctx.mark_non_differentiable(outs[1], outs[2])
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, g_x, _g_xn, _g_stats):
        tensors = ctx.params
        if g_x is None:   # the block's output did not reach the loss
            return (None,) * (9 + len(tensors))
        if ctx.saved is None:
            raise capi.SrkError("HatBlockFunction.backward ran twice: saved activations are released by the first "
                                "backward (retain_graph / double backward is not supported)")
        acts, lse, extra, weights = ctx.saved
        cfg, (B, H, W), kind, shift, conv_scale, drop = ctx.meta
        T = B * H * W
        dev = g_x.device
        dims = cfg.dims()
        g = g_x.contiguous()
        if g.dtype != BF16:
            g = g.to(BF16)
        params = [t.detach() for t in tensors[:13]]
        pdict = dict(zip(capi.PARAM_NAMES, params))
        scratch = eng._bwd_scratch(cfg, B, H, W, dev)
        geom = capi.SrkGeom(B, H, W, cfg.ws, shift)
        gdict = {n: torch.empty_like(p) for n, p in zip(capi.PARAM_NAMES, params)}
        g_in = torch.empty(T, cfg.Cp, device=dev, dtype=BF16)
        grads: list = [None] * len(tensors)
        if kind == "ocab":
            ws = _attn_ws(geom, capi.ATTN_OCA, cfg.heads, dev)
            capi.hat_block_bwd(dims, geom, weights, pdict, acts, g, scratch, g_in, gdict, capi.ATTN_OCA, lse, ws)
        else:
            c1, dc1, c2, pool, hidden, scale, Cm, Cm_p, S = extra
            c1w, c1b, c2w, c2b, s1w, s1b, s2w, s2b = tensors[13:21]
            ws = _attn_ws(geom, capi.ATTN_SELF, cfg.heads, dev)
            dxn1 = torch.empty(T, cfg.Cp, device=dev, dtype=BF16)
            gs_buf = torch.empty(T, cfg.Cp, device=dev, dtype=BF16) if drop else None
            capi.hat_block_bwd(dims, geom, weights, pdict, acts, g, scratch, None, gdict, capi.ATTN_SELF, lse, ws, d_xn1=dxn1,
                               drop=drop or (None, None), gs_buf=gs_buf)
            g_mid = scratch["g_mid"]
            # CAB backward: channel attention, conv2, GELU, conv1 (input gradient accumulated onto the attention path)
            d_c2 = torch.empty(T, cfg.Cp, device=dev, dtype=BF16)
            ds1w, ds1b, ds2w, ds2b = (torch.empty_like(t) for t in (s1w, s1b, s2w, s2b))
            capi.cab_se_bwd(g_mid, c2, B, H * W, cfg.C, S, s1w.detach(), s2w.detach(), conv_scale, pool, hidden, scale,
                            d_c2, ds1w, ds1b, ds2w, ds2b)
            _, wt2, _ = cv.conv_weights(c2w, c2b, cfg.Cp, Cm_p, refresh=False)
            d_c1 = torch.empty(T, Cm_p, device=dev, dtype=BF16)
            capi.conv3x3_igemm(capi.CEPI_MUL, B, H, W, cfg.Cp, Cm_p, Cm_p, d_c2, wt2, None, d_c1, r=dc1)
            dc2w, dc2b = torch.empty_like(c2w), torch.empty_like(c2b)
            capi.conv3x3_wgrad(B, H, W, Cm, cfg.C, Cm_p, cfg.Cp, False, d_c2, c1, dc2w)
            db_full = torch.empty(cfg.Cp, device=dev, dtype=torch.float32)
            capi.bias_grad_nhwc(d_c2, B, H, W, cfg.Cp, False, db_full)
            dc2b.copy_(db_full[:cfg.C])
            _, wt1, _ = cv.conv_weights(c1w, c1b, Cm_p, cfg.Cp, refresh=False)
            dxn1_tot = torch.empty(T, cfg.Cp, device=dev, dtype=BF16)
            capi.conv3x3_igemm(capi.CEPI_BIAS_RES, B, H, W, Cm_p, cfg.Cp, cfg.Cp, d_c1, wt1, None, dxn1_tot, r=dxn1)
            dc1w, dc1b = torch.empty_like(c1w), torch.empty_like(c1b)
            capi.conv3x3_wgrad(B, H, W, cfg.C, Cm, cfg.Cp, Cm_p, False, d_c1, acts["xn1"], dc1w)
            db1_full = torch.empty(Cm_p, device=dev, dtype=torch.float32)
            capi.bias_grad_nhwc(d_c1, B, H, W, Cm_p, False, db1_full)
            dc1b.copy_(db1_full[:Cm])
            # LayerNorm-1 backward over the summed branch gradients, plus the residual path g_mid
            capi.layernorm_bwd(dxn1_tot, acts["x_in"], acts["stats1"], params[0], g_mid, g_in, gdict["norm1_w"],
                               gdict["norm1_b"], cfg.C)
            grads[13:21] = [dc1w, dc1b, dc2w, dc2b, ds1w, ds1b, ds2w, ds2b]
        for j, n in enumerate(capi.PARAM_NAMES):
            grads[j] = gdict[n]
        ctx.saved = None
        return (g_in, None, None, None, None, None, None, None, None, *grads)


class RhagConvFunction(torch.autograd.Function):
    """y = conv3x3(x) + x_group_in on token-major bf16 (RHAG.forward, hat_arch.py:618-619; patch_unembed/embed are free
    in this layout)."""

    @staticmethod
    def forward(ctx, x, x_in, geom, C, weight, bias):
        B, H, W = geom
        Cp = x.shape[1]
        wf, _, bp = cv.conv_weights(weight, bias, Cp, Cp)
        y = torch.empty_like(x)
        capi.conv3x3_igemm(capi.CEPI_BIAS_RES, B, H, W, Cp, Cp, C, x, wf, bp, y, r=x_in)
        ctx.save_for_backward(x)
        ctx.meta = (geom, C, weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        (B, H, W), C, weight, bias = ctx.meta
        Cp = x.shape[1]
        dy = dy.contiguous()
        _, wt, _ = cv.conv_weights(weight, bias, Cp, Cp, refresh=False)
        dx = torch.empty_like(x)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, Cp, Cp, Cp, dy, wt, None, dx)
        dw, db = torch.empty_like(weight), torch.empty_like(bias)
        capi.conv3x3_wgrad(B, H, W, weight.shape[1], weight.shape[0], Cp, Cp, False, dy, x, dw)
        db_full = torch.empty(Cp, device=dy.device, dtype=torch.float32)
        capi.bias_grad_nhwc(dy, B, H, W, Cp, False, db_full)
        db.copy_(db_full[:C])
        return dx, dy, None, None, dw, db


class LayerNormTokensFunction(torch.autograd.Function):
    """Differentiable LayerNorm over token-major bf16 rows (HAT.patch_embed.norm :801-806, HAT.norm :966)."""

    @staticmethod
    def forward(ctx, x, weight, bias, C: int, ones_col: int):
        y = torch.empty_like(x)
        stats = torch.empty(x.shape[0], 2, device=x.device, dtype=torch.float32)
        capi.layernorm_fwd(x, y, stats, weight.detach(), bias.detach(), C, ones_col=ones_col)
        ctx.save_for_backward(x, stats, weight)
        ctx.C = C
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, weight = ctx.saved_tensors
        dy = dy.contiguous().to(BF16)
        dx = torch.empty_like(x)
        dgamma, dbeta = torch.empty_like(weight), torch.empty_like(weight)
        capi.layernorm_bwd(dy, x, stats, weight.detach(), None, dx, dgamma, dbeta, ctx.C)
        return dx, dgamma, dbeta, None, None


class ChannelAttentionFunction(torch.autograd.Function):
    """Stand-alone ChannelAttention.forward on an NCHW tensor (hat_arch.py:40-58): x * sigmoid(W2 relu(W1 avgpool(x) + b1) + b2)
    on the kernels of the fused path (srk_cab_se_fwd / _bwd with a zero shortcut and alpha = 1); the NCHW <-> token-major
    packing is torch indexing.  Inside CAB / HAB the same kernels run without this detour."""

    @staticmethod
    def forward(ctx, x, s1w, s1b, s2w, s2b):
        B, C, H, W = x.shape
        if C >= 192:
            raise capi.SrkError("libsrk ChannelAttention: fewer than 192 channels")
        dev, T, Cp = x.device, B * H * W, 192
        S = s1w.shape[0]
        for t in (s1w, s1b, s2w, s2b):
            eng._check_param(t)
        xt = torch.zeros(T, Cp, device=dev, dtype=BF16)
        xt[:, :C] = x.permute(0, 2, 3, 1).reshape(T, C).to(BF16)
        pool = torch.empty(B, C, device=dev, dtype=torch.float32)
        hidden = torch.empty(B, S, device=dev, dtype=torch.float32)
        scale = torch.empty(B, C, device=dev, dtype=torch.float32)
        zero = torch.zeros(T, Cp, device=dev, dtype=BF16)
        out = torch.empty(T, Cp, device=dev, dtype=BF16)
        capi.cab_se_fwd(xt, zero, B, H * W, C, S, s1w.detach(), s1b.detach(), s2w.detach(), s2b.detach(), 1.0, pool, hidden,
                        scale, out)
        ctx.saved = (xt, pool, hidden, scale)
        ctx.meta = (B, C, H, W, S, x.dtype)
        ctx.params = (s1w, s1b, s2w, s2b)
        return out[:, :C].reshape(B, H, W, C).permute(0, 3, 1, 2).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        xt, pool, hidden, scale = ctx.saved
        B, C, H, W, S, dtype = ctx.meta
        s1w, s1b, s2w, s2b = ctx.params
        dev, T, Cp = dy.device, B * H * W, 192
        g = torch.zeros(T, Cp, device=dev, dtype=BF16)
        g[:, :C] = dy.permute(0, 2, 3, 1).reshape(T, C).to(BF16)
        dxt = torch.empty(T, Cp, device=dev, dtype=BF16)
        ds1w, ds1b, ds2w, ds2b = (torch.empty_like(t) for t in (s1w, s1b, s2w, s2b))
        capi.cab_se_bwd(g, xt, B, H * W, C, S, s1w.detach(), s2w.detach(), 1.0, pool, hidden, scale, dxt, ds1w, ds1b, ds2w,
                        ds2b)
        return dxt[:, :C].reshape(B, H, W, C).permute(0, 3, 1, 2).to(dtype), ds1w, ds1b, ds2w, ds2b


class CabFunction(torch.autograd.Function):
    """Stand-alone CAB.forward on an NCHW tensor (reference CAB / ChannelAttention, hat_arch.py:40-74): conv3x3 -> GELU ->
    conv3x3 -> x * sigmoid(W2 relu(W1 avgpool(x))).  Same kernels as the fused HAB path; only the NCHW <-> token-major
    packing is done with torch indexing."""

    @staticmethod
    def forward(ctx, x, c1w, c1b, c2w, c2b, s1w, s1b, s2w, s2b):
        B, C, H, W = x.shape
        if C >= 192 or H % 8 or W % 16:
            raise capi.SrkError("libsrk CAB: fewer than 192 channels, H % 8 == 0, W % 16 == 0")
        dev, T, Cp = x.device, B * H * W, 192
        Cm, S = c1w.shape[0], s1w.shape[0]
        Cm_p = cv._pad64(Cm)
        for t in (c1w, c1b, c2w, c2b, s1w, s1b, s2w, s2b):
            eng._check_param(t)
        xt = torch.zeros(T, Cp, device=dev, dtype=BF16)
        xt[:, :C] = x.permute(0, 2, 3, 1).reshape(T, C).to(BF16)
        wf1, _, bp1 = cv.conv_weights(c1w, c1b, Cm_p, Cp)
        wf2, _, bp2 = cv.conv_weights(c2w, c2b, Cp, Cm_p)
        c1 = torch.empty(T, Cm_p, device=dev, dtype=BF16)
        dc1 = torch.empty_like(c1)
        capi.conv3x3_igemm(capi.CEPI_BIAS_GELU, B, H, W, Cp, Cm_p, Cm, xt, wf1, bp1, c1, y2=dc1)
        c2 = torch.empty(T, Cp, device=dev, dtype=BF16)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, Cm_p, Cp, C, c1, wf2, bp2, c2)
        pool = torch.empty(B, C, device=dev, dtype=torch.float32)
        hidden = torch.empty(B, S, device=dev, dtype=torch.float32)
        scale = torch.empty(B, C, device=dev, dtype=torch.float32)
        zero = torch.zeros(T, Cp, device=dev, dtype=BF16)
        out = torch.empty(T, Cp, device=dev, dtype=BF16)
        capi.cab_se_fwd(c2, zero, B, H * W, C, S, s1w.detach(), s1b.detach(), s2w.detach(), s2b.detach(), 1.0, pool, hidden,
                        scale, out)
        ctx.saved = (xt, c1, dc1, c2, pool, hidden, scale)
        ctx.meta = (B, C, H, W, Cm, Cm_p, S, x.dtype)
        ctx.params = (c1w, c1b, c2w, c2b, s1w, s1b, s2w, s2b)
        return out[:, :C].reshape(B, H, W, C).permute(0, 3, 1, 2).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        xt, c1, dc1, c2, pool, hidden, scale = ctx.saved
        B, C, H, W, Cm, Cm_p, S, dtype = ctx.meta
        c1w, c1b, c2w, c2b, s1w, s1b, s2w, s2b = ctx.params
        dev, T, Cp = dy.device, B * H * W, 192
        g = torch.zeros(T, Cp, device=dev, dtype=BF16)
        g[:, :C] = dy.permute(0, 2, 3, 1).reshape(T, C).to(BF16)
        d_c2 = torch.empty(T, Cp, device=dev, dtype=BF16)
        ds1w, ds1b, ds2w, ds2b = (torch.empty_like(t) for t in (s1w, s1b, s2w, s2b))
        capi.cab_se_bwd(g, c2, B, H * W, C, S, s1w.detach(), s2w.detach(), 1.0, pool, hidden, scale, d_c2, ds1w, ds1b,
                        ds2w, ds2b)
        _, wt2, _ = cv.conv_weights(c2w, c2b, Cp, Cm_p, refresh=False)
        d_c1 = torch.empty(T, Cm_p, device=dev, dtype=BF16)
        capi.conv3x3_igemm(capi.CEPI_MUL, B, H, W, Cp, Cm_p, Cm_p, d_c2, wt2, None, d_c1, r=dc1)
        dc2w, dc2b = torch.empty_like(c2w), torch.empty_like(c2b)
        capi.conv3x3_wgrad(B, H, W, Cm, C, Cm_p, Cp, False, d_c2, c1, dc2w)
        db_full = torch.empty(Cp, device=dev, dtype=torch.float32)
        capi.bias_grad_nhwc(d_c2, B, H, W, Cp, False, db_full)
        dc2b.copy_(db_full[:C])
        _, wt1, _ = cv.conv_weights(c1w, c1b, Cm_p, Cp, refresh=False)
        dxt = torch.empty(T, Cp, device=dev, dtype=BF16)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, Cm_p, Cp, Cp, d_c1, wt1, None, dxt)
        dc1w, dc1b = torch.empty_like(c1w), torch.empty_like(c1b)
        capi.conv3x3_wgrad(B, H, W, C, Cm, Cp, Cm_p, False, d_c1, xt, dc1w)
        db1_full = torch.empty(Cm_p, device=dev, dtype=torch.float32)
        capi.bias_grad_nhwc(d_c1, B, H, W, Cm_p, False, db1_full)
        dc1b.copy_(db1_full[:Cm])
        dx = dxt[:, :C].reshape(B, H, W, C).permute(0, 3, 1, 2).to(dtype)
        return dx, dc1w, dc1b, dc2w, dc2b, ds1w, ds1b, ds2w, ds2b
