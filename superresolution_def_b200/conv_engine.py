"""Convolutional head / tail of SwinIR on token-major (NHWC) bf16 activations.

INTERIM (round 1, first slice): the 3x3 convolutions below still go through torch's cuDNN path so that the
transformer-block kernels can be validated end to end; they are being replaced by the implicit-GEMM tcgen05
kernels in csrc/conv3x3.cuh (bias / LeakyReLU / residual / PixelShuffle epilogues).  Nothing here runs on CPU.
Reference: SwinIR.forward head/tail, models/architecture_swin.py:241-256; Upsample :175-190.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _capi as capi

BF16 = torch.bfloat16


def _require_cuda(t):
    if not t.is_cuda:
        raise capi.SrkError("superresolution_def_b200 runs on CUDA (sm_100a) only; there is no CPU path")


def conv3x3_tokens(x, weight, bias, Cp: int):
    """conv_first: (B,Cin,H,W) -> token-major bf16 [B*H*W, Cp] (pad channels zero)."""
    _require_cuda(x)
    with torch.autocast("cuda", dtype=BF16):
        y = F.conv2d(x, weight, bias, padding=1)
    B, C, H, W = y.shape
    out = y.new_zeros((B * H * W, Cp), dtype=BF16)
    out[:, :C] = y.permute(0, 2, 3, 1).reshape(B * H * W, C)
    return out


def swinir_tail(body, first, geom, C, conv_after_body, conv_before_up, upsample, conv_last):
    """conv_after_body(body)+first -> conv_before_upsample+LeakyReLU -> 2x[conv, PixelShuffle] -> conv_last."""
    B, H, W = geom
    to_nchw = lambda t: t[:, :C].reshape(B, H, W, C).permute(0, 3, 1, 2)  # noqa: E731
    with torch.autocast("cuda", dtype=BF16):
        res = F.conv2d(to_nchw(body), conv_after_body.weight, conv_after_body.bias, padding=1) + to_nchw(first)
        out = F.leaky_relu(F.conv2d(res, conv_before_up.weight, conv_before_up.bias, padding=1), 0.01)
        for m in upsample:
            out = m(out)
        out = F.conv2d(out, conv_last.weight, conv_last.bias, padding=1)
    return out
