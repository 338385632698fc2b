"""Convolutional head / tail of the SR generators on NHWC bf16 activations (libsrk implicit-GEMM kernels).

Activations stay token-major ([B*H*W, C] == NHWC) end to end, so the NCHW<->NHWC shuffles the reference pays in
PatchEmbed/PatchUnEmbed (models/architecture_swin.py:153-171) disappear; PixelShuffle is folded into the store
addresses of the producing convolution; LeakyReLU and the residual add are GEMM epilogues.

Reference: SwinIR.forward head/tail, models/architecture_swin.py:241-256; Upsample :175-190.
"""
from __future__ import annotations

import torch

from . import _capi as capi

BF16 = torch.bfloat16


def _require_cuda(t):
    if not t.is_cuda:
        raise capi.SrkError("superresolution_def_b200 runs on CUDA (sm_100a) only; there is no CPU path")


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


class _ConvWeightCache:
    """bf16 implicit-GEMM operands of one Conv2d(3x3); re-derived from the fp32 master weights on every forward
    (see swin_engine._WeightCache for why no staleness test is trusted)."""

    def __init__(self):
        self.wf = self.wt = self.bias = None

    def get(self, w, b, Cout_p, Cin_p, ps, refresh):
        if self.wf is None or self.wf.numel() != Cout_p * 9 * Cin_p:
            self.wf = torch.empty(Cout_p * 9 * Cin_p, device=w.device, dtype=BF16)
            self.wt = torch.empty(Cin_p * 9 * Cout_p, device=w.device, dtype=BF16)
            self.bias = torch.empty(Cout_p, device=w.device, dtype=torch.float32)
            refresh = True
        if refresh:
            capi.conv3x3_prep_weights(w.detach(), None if b is None else b.detach(), Cout_p, Cin_p, ps, self.wf, self.wt,
                                      self.bias)
        return self.wf, self.wt, self.bias


_conv_caches: dict = {}


def conv_weights(w, b, Cout_p, Cin_p, ps=False, refresh=True):
    """refresh=True in every forward; the backward of the same step passes refresh=False to reuse them."""
    k = (w.data_ptr(), Cout_p, Cin_p, ps)
    c = _conv_caches.get(k)
    if c is None:
        c = _conv_caches[k] = _ConvWeightCache()
    return c.get(w, b, Cout_p, Cin_p, ps, refresh)


class ConvFirstFunction(torch.autograd.Function):
    """conv_first: (B,1,H,W) fp32 image -> token-major bf16 [B*H*W, Cp] (architecture_swin.py:202,241)."""

    @staticmethod
    def forward(ctx, x, weight, bias, Cp: int):
        _require_cuda(x)
        B, cin, H, W = x.shape
        if cin != 1:
            raise capi.SrkError("conv_first kernel handles in_chans == 1 (the reference's only configuration)")
        C = weight.shape[0]
        xf = x.detach().contiguous().float()
        y = torch.empty(B * H * W, Cp, device=x.device, dtype=BF16)
        capi.conv_in1_fwd(xf, weight.detach(), bias.detach(), y, B, H, W, C, Cp)
        ctx.save_for_backward(xf)
        ctx.meta = (B, H, W, C, Cp, weight.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xf,) = ctx.saved_tensors
        B, H, W, C, Cp, wshape = ctx.meta
        dw = torch.empty(wshape, device=dy.device, dtype=torch.float32)
        db = torch.empty(C, device=dy.device, dtype=torch.float32)
        capi.conv_in1_wgrad(xf, dy.contiguous(), dw, db, B, H, W, C, Cp)
        return None, dw, db, None


class SwinIRTailFunction(torch.autograd.Function):
    """conv_after_body(body)+first -> conv_before_upsample+LeakyReLU -> 2 x [conv 64->256 + PixelShuffle(2)] ->
    conv_last, all NHWC bf16 (architecture_swin.py:249-255).  One autograd node; 5 forward and 13 backward launches."""

    @staticmethod
    def forward(ctx, body, first, geom, C, w_ab, b_ab, w_bu, b_bu, w_u0, b_u0, w_u2, b_u2, w_last, b_last):
        B, H, W = geom
        dev = body.device
        Cp = body.shape[1]
        F = w_bu.shape[0]  # 64
        if F != 64 or w_u0.shape[0] != 256 or (w_u2 is not None and w_u2.shape[0] != 256) or w_last.shape[0] != 1:
            raise capi.SrkError("SwinIR tail kernels are specialised for num_feat=64, x4 PixelShuffle, 1 output channel")
        wf_ab, _, bp_ab = conv_weights(w_ab, b_ab, Cp, Cp)
        wf_bu, _, bp_bu = conv_weights(w_bu, b_bu, 64, Cp)
        wf_u0, _, bp_u0 = conv_weights(w_u0, b_u0, 256, 64, ps=True)
        two = w_u2 is not None   # x4 = two conv+PixelShuffle(2) stages, x2 = one (Upsample, architecture_swin.py:175-190)
        up = 4 if two else 2
        res = torch.empty(B * H * W, Cp, device=dev, dtype=BF16)
        capi.conv3x3_igemm(capi.CEPI_BIAS_RES, B, H, W, Cp, Cp, C, body, wf_ab, bp_ab, res, r=first)
        t64 = torch.empty(B * H * W, 64, device=dev, dtype=BF16)
        capi.conv3x3_igemm(capi.CEPI_BIAS_LRELU, B, H, W, Cp, 64, 64, res, wf_bu, bp_bu, t64, slope=0.01)
        u0 = torch.empty(B * 2 * H * 2 * W, 64, device=dev, dtype=BF16)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, 64, 256, 256, t64, wf_u0, bp_u0, u0, y_ps=True)
        u1 = u0
        if two:
            wf_u2, _, bp_u2 = conv_weights(w_u2, b_u2, 256, 64, ps=True)
            u1 = torch.empty(B * 4 * H * 4 * W, 64, device=dev, dtype=BF16)
            capi.conv3x3_igemm(capi.CEPI_BIAS, B, 2 * H, 2 * W, 64, 256, 256, u0, wf_u2, bp_u2, u1, y_ps=True)
        out = torch.empty(B, 1, up * H, up * W, device=dev, dtype=torch.float32)
        # conv_last (64 -> 1): one useful output column makes the tensor-core form (N = 16) instruction-rate bound (613 us at
        # 512^2 x 16); the row-walking CUDA-core kernel streams the input once
        capi.conv_out1_fwd(u1, w_last.detach(), b_last.detach(), out, B, up * H, up * W, 64)
        if any(ctx.needs_input_grad):
            ctx.saved = (body, res, t64, u0, u1)
            ctx.params = (w_ab, b_ab, w_bu, b_bu, w_u0, b_u0, w_u2, b_u2, w_last, b_last)
            ctx.meta = (B, H, W, C, Cp)
        return out

    @staticmethod
    def backward(ctx, dout):
        body, res, t64, u0, u1 = ctx.saved
        w_ab, b_ab, w_bu, b_bu, w_u0, b_u0, w_u2, b_u2, w_last, b_last = ctx.params
        B, H, W, C, Cp = ctx.meta
        dev = dout.device
        f32 = torch.float32
        dout = dout.contiguous().float()
        two = w_u2 is not None
        up = 4 if two else 2
        # conv_last
        d_u1 = torch.empty_like(u1)
        dw_last, db_last = torch.empty_like(w_last), torch.empty_like(b_last)
        capi.conv_out1_bwd(dout, u1, w_last.detach(), d_u1, dw_last, db_last, B, up * H, up * W, 64)
        dw_u2 = db_u2 = None
        d_u0 = d_u1
        if two:
            # upsample.2 (input u0 at 2H x 2W, gradient arrives pixel-shuffled at 4H x 4W)
            _, wt_u2, _ = conv_weights(w_u2, b_u2, 256, 64, ps=True, refresh=False)
            d_u0 = torch.empty_like(u0)
            capi.conv3x3_igemm(capi.CEPI_BIAS, B, 2 * H, 2 * W, 256, 64, 64, d_u1, wt_u2, None, d_u0, x_ps=True)
            dw_u2, db_u2 = torch.empty_like(w_u2), torch.empty_like(b_u2)
            capi.conv3x3_wgrad(B, 2 * H, 2 * W, 64, 256, 64, 256, True, d_u1, u0, dw_u2)
            capi.bias_grad_nhwc(d_u1, B, 2 * H, 2 * W, 256, True, db_u2)
        # upsample.0 (+ LeakyReLU backward fused into the input-gradient epilogue)
        _, wt_u0, _ = conv_weights(w_u0, b_u0, 256, 64, ps=True, refresh=False)
        d_t64 = torch.empty_like(t64)
        capi.conv3x3_igemm(capi.CEPI_MASK_LRELU, B, H, W, 256, 64, 64, d_u0, wt_u0, None, d_t64, x_ps=True, r=t64,
                           slope=0.01)
        dw_u0, db_u0 = torch.empty_like(w_u0), torch.empty_like(b_u0)
        capi.conv3x3_wgrad(B, H, W, 64, 256, 64, 256, True, d_u0, t64, dw_u0)
        capi.bias_grad_nhwc(d_u0, B, H, W, 256, True, db_u0)
        # conv_before_upsample
        _, wt_bu, _ = conv_weights(w_bu, b_bu, 64, Cp, refresh=False)
        d_res = torch.empty_like(res)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, 64, Cp, Cp, d_t64, wt_bu, None, d_res)
        dw_bu, db_bu = torch.empty_like(w_bu), torch.empty_like(b_bu)
        capi.conv3x3_wgrad(B, H, W, w_bu.shape[1], 64, Cp, 64, False, d_t64, res, dw_bu)
        capi.bias_grad_nhwc(d_t64, B, H, W, 64, False, db_bu)
        # conv_after_body (+ residual: d_first = d_res)
        _, wt_ab, _ = conv_weights(w_ab, b_ab, Cp, Cp, refresh=False)
        d_body = torch.empty_like(body)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, Cp, Cp, Cp, d_res, wt_ab, None, d_body)
        dw_ab, db_ab = torch.empty_like(w_ab), torch.empty_like(b_ab)
        capi.conv3x3_wgrad(B, H, W, w_ab.shape[1], w_ab.shape[0], Cp, Cp, False, d_res, body, dw_ab)
        db_full = torch.empty(Cp, device=dev, dtype=f32)
        capi.bias_grad_nhwc(d_res, B, H, W, Cp, False, db_full)
        db_ab.copy_(db_full[:C])
        ctx.saved = None
        return (d_body, d_res, None, None, dw_ab, db_ab, dw_bu, db_bu, dw_u0, db_u0, dw_u2, db_u2, dw_last, db_last)


def conv3x3_tokens(x, weight, bias, Cp: int):
    return ConvFirstFunction.apply(x, weight, bias, Cp)


def swinir_tail(body, first, geom, C, conv_after_body, conv_before_up, upsample, conv_last):
    two = len(upsample) > 2   # x4: Sequential(conv, PixelShuffle, conv, PixelShuffle); x2: (conv, PixelShuffle)
    return SwinIRTailFunction.apply(body, first, geom, C, conv_after_body.weight, conv_after_body.bias,
                                    conv_before_up.weight, conv_before_up.bias, upsample[0].weight, upsample[0].bias,
                                    upsample[2].weight if two else None, upsample[2].bias if two else None,
                                    conv_last.weight, conv_last.bias)
