"""Both U-Net discriminators forward / backward on libsrk: models/discriminator_swin.py:43-84
(`UNetDiscriminatorFunction`, described first) and models/discriminator_hat.py:8-49 (`UNetDiscriminatorHatFunction`, second
half of this file).

One autograd node for the whole U-Net, NHWC bf16 activations, fp32 accumulation:

* the five `Conv2d(.., 4, 2, 1)` layers (:52, :10) and the four `ConvTranspose2d(.., 4, 2, 1)` layers (:25) are GEMMs on the
  persistent tcgen05 kernel (`srk_gemm_tn` / `srk_gemm_tn_lrelu`) over a patch matrix (`srk_disc_patches_k4s2`) or followed
  by a fold (`srk_disc_fold_k4s2`); LeakyReLU(0.2) is the GEMM epilogue or part of the fold, its backward mask rides on
  the gather / fold that moves the gradient anyway; weight gradients are `srk_gemm_wgrad` (MN-major tcgen05);
* `torch.cat((x, skip), 1)` (:40) never happens: every level owns ONE [pixels, C_up + C_skip] buffer, the encoder writes
  its half through the GEMM's row pitch and the decoder folds into the other half; the concat's backward is the `add`
  operand of the fold;
* the 1 -> 64 head (:49), the 128 -> 64 (:67) and 64 -> 1 (:69) 3x3 layers reuse the generators' kernels
  (`srk_conv_in1_*`, `srk_conv3x3_igemm` / `_wgrad`, `srk_conv_out1_*`).

Spectral normalisation (torch.nn.utils.spectral_norm, :10,25,49-52,67-69) runs inside the node too when the caller hands in
the `weight_orig` parameters and the `weight_u` / `weight_v` buffers (`unet_discriminator_sn`, what gan.py does):
`srk_spectral_norm` performs the hook's power iteration in place on the buffers, 1 / sigma is folded into the bf16 operand
packing, and `srk_spectral_norm_bwd` maps the gradients back to `weight_orig`.  `unet_discriminator` takes already
normalised weights (tests).

The bilinear resize of UNetUpBlock (:36-38) only triggers when an encoder level has odd size; inputs must therefore be
multiples of 32 pixels (the reference scripts feed 512^2) — anything else raises.
"""
from __future__ import annotations

import torch

from . import _capi as capi

BF16 = torch.bfloat16
SLOPE = 0.2


def _rows(m: int) -> int:
    return (m + 127) // 128 * 128


def _alloc(m: int, c: int, dev) -> torch.Tensor:
    """[ceil128(m), c] bf16; the GEMMs work on 128-row tiles, rows beyond m are kept zero."""
    r = _rows(m)
    t = torch.empty(r, c, device=dev, dtype=BF16)
    if r > m:
        t[m:].zero_()
    return t


def _ops4(w, sigma=None):
    """4x4 weight [P,Q,4,4] fp32 -> (a [P,16Q], a^T [16Q,P]) bf16 of W / sigma, column (ky*4+kx)*Q + q.
    Conv2d [Cout,Cin,4,4]: a = Wf (forward operand), a^T = Wt (input-gradient operand);
    ConvTranspose2d [Cin,Cout,4,4]: a = Wd (input-gradient operand), a^T = Wu (forward operand)."""
    P, Q = w.shape[0], w.shape[1]
    a = torch.empty(P, 16 * Q, device=w.device, dtype=BF16)
    at = torch.empty(16 * Q, P, device=w.device, dtype=BF16)
    capi.disc_prep_w4(w, a, at, sigma)
    return a, at


# the `dim` torch.nn.utils.spectral_norm picks per layer: 0 for nn.Conv2d, 1 for nn.ConvTranspose2d (the four up blocks)
SN_DIMS = (0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0)
_SMALL = (0, 10, 11)   # the 3x3 layers: their kernels take fp32 filters, so W / sigma is materialised for them


def _wgrad(A, B2d, R):
    """dw [Cb,R,4,4] fp32, dw[c][r][ky][kx] = sum_t A[t][(ky*4+kx)*R + r] * B[t][c]: the parameter's own layout."""
    dw = torch.empty(B2d.shape[1], R, 4, 4, device=A.device, dtype=torch.float32)
    capi.disc_wgrad4(A, B2d, R, dw)
    return dw


class UNetDiscriminatorFunction(torch.autograd.Function):
    """x [B,1,H,W] -> logits [B,1,H/2,W/2] (UNetDiscriminatorSN.forward, discriminator_swin.py:72-84)."""

    @staticmethod
    def forward(ctx, x, sn, *weights):
        """weights: the 12 convolution weights in forward order.  sn is None: they are used as given (already normalised);
        sn = dict(u=[...], v=[...], training=bool, eps=float): they are the `weight_orig` parameters and spectral
        normalisation (one power iteration on u / v in place when training, sigma, W / sigma) runs here on libsrk."""
        B, cin, H, W = x.shape
        if len(weights) != 12:
            raise capi.SrkError("UNetDiscriminatorSN has 12 convolutions")
        dev = x.device
        weights = [w.detach().float().contiguous() for w in weights]
        sig = [None] * 12
        small = {i: weights[i] for i in _SMALL}
        sn_ctx = None
        if sn is not None:
            sigmas = torch.empty(12, device=dev, dtype=torch.float32)
            w_sn = [torch.empty_like(weights[i]) if i in _SMALL else None for i in range(12)]
            capi.spectral_norm(capi.sn_layers(weights, sn["u"], sn["v"], SN_DIMS, sigmas, w_sn), sn["training"], sn["eps"], dev)
            sig = [sigmas[i:i + 1] for i in range(12)]
            small = {i: w_sn[i] for i in _SMALL}
            if any(ctx.needs_input_grad[2:]):
                # u and v as this forward left them (the hook clones them for the same reason: a second forward before
                # the backward runs another power iteration in place)
                sizes = [t.numel() for t in sn["u"]] + [t.numel() for t in sn["v"]]
                flat = torch.cat([t.reshape(-1) for t in list(sn["u"]) + list(sn["v"])])
                parts = list(torch.split(flat, sizes))
                sn_ctx = (parts[:12], parts[12:], sigmas)
        w0a, w0b, w1, w2, w3, w4, u1, u2, u3, u4, wf1, wf2 = weights
        nf = w0a.shape[0]
        if cin != 1 or nf != 64 or wf2.shape[0] != 1:
            raise capi.SrkError("discriminator kernels are specialised for num_in_ch=1, num_feat=64 (the scripts' only configuration)")
        if H % 32 or W % 32:
            raise capi.SrkError("UNetDiscriminatorSN: H and W must be multiples of 32 (the bilinear resize of UNetUpBlock, "
                                "discriminator_swin.py:36-38, is not implemented)")
        f32 = torch.float32
        xf = x.detach().contiguous().float().reshape(B, H, W)
        Hs = [H >> i for i in range(6)]
        Ws = [W >> i for i in range(6)]
        M = [B * Hs[i] * Ws[i] for i in range(6)]
        V = capi.view
        zeros64 = torch.zeros(64, device=dev, dtype=f32)

        # conv0: 1 -> 64 3x3 + LeakyReLU, then 64 -> 64 4x4 s2 + LeakyReLU (:48-53)
        a0 = _alloc(M[0], 64, dev)
        capi.conv_in1_fwd(xf, small[0], zeros64, a0, B, H, W, 64, 64)
        capi.view_lrelu(V(a0), M[0], SLOPE)
        # one buffer per level: [decoder output | encoder skip] = torch.cat((x, skip_input), 1) (:40)
        cat4 = _alloc(M[1], 128, dev)    # d4 (64)  | x0 (64)
        cat3 = _alloc(M[2], 256, dev)    # d3 (128) | x1 (128)
        cat2 = _alloc(M[3], 512, dev)    # d2 (256) | x2 (256)
        cat1 = _alloc(M[4], 1024, dev)   # d1 (512) | x3 (512)
        x4 = _alloc(M[5], 512, dev)
        ops_d = [_ops4(weights[i], sig[i]) for i in range(1, 6)]         # (Wf, Wt)
        ops_u = [_ops4(weights[i], sig[i])[::-1] for i in range(6, 10)]  # (Wu, Wd)
        keep_patches = any(ctx.needs_input_grad[2:])   # the weight gradients read the forward's patch matrices again
        saved_patches = []

        def down(src, c0, cin_, lvl, wf, dst, d0):
            """level lvl [B,Hs,Ws,cin_] (channels c0.. of src) -> level lvl+1, written into channels d0.. of dst"""
            p = _alloc(M[lvl + 1], 16 * cin_, dev)
            capi.disc_patches_k4s2(V(src, c0, cin_), None, SLOPE, B, Hs[lvl], Ws[lvl], p)
            capi.gemm_tn_lrelu(p, wf, dst[:, d0:d0 + wf.shape[0]], SLOPE)
            if keep_patches:
                saved_patches.append(p)

        down(a0, 0, 64, 0, ops_d[0][0], cat4, 64)       # x0
        down(cat4, 64, 64, 1, ops_d[1][0], cat3, 128)   # x1 (conv1 :55)
        down(cat3, 128, 128, 2, ops_d[2][0], cat2, 256)  # x2
        down(cat2, 256, 256, 3, ops_d[3][0], cat1, 512)  # x3
        down(cat1, 512, 512, 4, ops_d[4][0], x4, 0)      # x4

        def up(src, lvl, wu, dst):
            """level lvl (all channels of src) -> level lvl-1, folded into channels [0, Cout) of dst"""
            cout = wu.shape[0] // 16
            taps = torch.empty(src.shape[0], 16 * cout, device=dev, dtype=BF16)
            capi.gemm_tn(capi.EPI_STORE, src, wu, taps)
            capi.disc_fold_k4s2(taps, B, Hs[lvl], Ws[lvl], V(dst, 0, cout), act=capi.FOLD_LRELU, slope=SLOPE)

        up(x4, 5, ops_u[0][0], cat1)    # d1 (up1 :60)
        up(cat1, 4, ops_u[1][0], cat2)  # d2
        up(cat2, 3, ops_u[2][0], cat3)  # d3
        up(cat3, 2, ops_u[3][0], cat4)  # d4

        # final_conv: 128 -> 64 3x3 + LeakyReLU, 64 -> 1 3x3 (:66-70)
        wk_f1 = torch.empty(64 * 9 * 128, device=dev, dtype=BF16)
        wt_f1 = torch.empty(128 * 9 * 64, device=dev, dtype=BF16)
        b_f1 = torch.empty(64, device=dev, dtype=f32)
        capi.conv3x3_prep_weights(small[10], None, 64, 128, False, wk_f1, wt_f1, b_f1)
        f1a = _alloc(M[1], 64, dev)
        capi.conv3x3_igemm(capi.CEPI_BIAS_LRELU, B, Hs[1], Ws[1], 128, 64, 64, cat4, wk_f1, b_f1, f1a, slope=SLOPE)
        out = torch.empty(B, 1, Hs[1], Ws[1], device=dev, dtype=f32)
        w_last = small[11]
        capi.conv_out1_fwd(f1a, w_last, torch.zeros(1, device=dev, dtype=f32), out, B, Hs[1], Ws[1], 64)

        if any(ctx.needs_input_grad):
            ctx.acts = (xf, a0, cat4, cat3, cat2, cat1, x4, f1a)
            ctx.ops = (ops_d, ops_u, wt_f1, w_last, small[0])
            ctx.sn = (sn_ctx, weights if sn_ctx is not None else None)
            ctx.patches = saved_patches
            ctx.meta = (B, Hs, Ws, M, x.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.acts is None:
            raise capi.SrkError("UNetDiscriminatorFunction: second backward through the same forward (activations were released)")
        xf, a0, cat4, cat3, cat2, cat1, x4, f1a = ctx.acts
        ops_d, ops_u, wt_f1, w_last, w0a = ctx.ops
        patches = ctx.patches
        B, Hs, Ws, M, x_dtype = ctx.meta
        need = ctx.needs_input_grad
        need_x, need_w = need[0], any(need[2:])
        dev = dout.device
        f32 = torch.float32
        V = capi.view
        dout = dout.contiguous().float()

        # final_conv[2] (64 -> 1) and the LeakyReLU before it
        d_f1 = _alloc(M[1], 64, dev)
        dw_f2 = torch.empty(1, 64, 3, 3, device=dev, dtype=f32)
        db = torch.empty(1, device=dev, dtype=f32)
        capi.conv_out1_bwd(dout, f1a, w_last, d_f1, dw_f2, db, B, Hs[1], Ws[1], 64)
        capi.view_lrelu_mask(V(d_f1), V(f1a), M[1], SLOPE)
        # final_conv[0] (128 -> 64)
        d_cat4 = _alloc(M[1], 128, dev)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, Hs[1], Ws[1], 64, 128, 128, d_f1, wt_f1, None, d_cat4)
        dw_f1 = None
        if need_w:
            dw_f1 = torch.empty(64, 128, 3, 3, device=dev, dtype=f32)
            capi.conv3x3_wgrad(B, Hs[1], Ws[1], 128, 64, 128, 64, False, d_f1, cat4, dw_f1)

        # decoder, last block first: gradient of level lvl-1's first `cout` channels -> gradient of all of level lvl
        def up_bwd(g_dst, f_dst, lvl, src, wd):
            cin_ = wd.shape[0]
            cout = wd.shape[1] // 16
            g = _alloc(M[lvl], 16 * cout, dev)
            capi.disc_patches_k4s2(V(g_dst, 0, cout), V(f_dst, 0, cout), SLOPE, B, Hs[lvl - 1], Ws[lvl - 1], g)
            d_src = _alloc(M[lvl], cin_, dev)
            capi.gemm_tn(capi.EPI_STORE, g, wd, d_src)
            dw = None
            if need_w:   # dW[ci,co,ky,kx] = sum_m src[m,ci] g[m,(ky,kx,co)]
                dw = _wgrad(g, src, cout)
            return d_src, dw

        d_cat3, dw_u4 = up_bwd(d_cat4, cat4, 2, cat3, ops_u[3][1])
        d_cat2, dw_u3 = up_bwd(d_cat3, cat3, 3, cat2, ops_u[2][1])
        d_cat1, dw_u2 = up_bwd(d_cat2, cat2, 4, cat1, ops_u[1][1])
        d_x4, dw_u1 = up_bwd(d_cat1, cat1, 5, x4, ops_u[0][1])
        for c0 in (0, 256):   # LeakyReLU backward of conv4's output (the mask kernel handles <= 256 channels per call)
            capi.view_lrelu_mask(V(d_x4, c0, 256), V(x4, c0, 256), M[5], SLOPE)

        # encoder, deepest first: d_pre of level lvl+1 -> masked gradient of the source channels at level lvl
        def down_bwd(d_pre, lvl, src, c0, cin_, wt, add, a0_):
            dw = None
            if need_w:   # dW[co,ci,ky,kx] = sum_m d_pre[m,co] patches[m,(ky,kx,ci)]
                dw = _wgrad(patches[lvl], d_pre, cin_)   # the forward's patch matrix of this level (kept, not re-gathered)
                patches[lvl] = None
            taps = torch.empty(d_pre.shape[0], 16 * cin_, device=dev, dtype=BF16)
            capi.gemm_tn(capi.EPI_STORE, d_pre, wt, taps)
            d_src = _alloc(M[lvl], cin_, dev)
            capi.disc_fold_k4s2(taps, B, Hs[lvl + 1], Ws[lvl + 1], V(d_src), add=None if add is None else V(add, a0_, cin_),
                                f=V(src, c0, cin_), act=capi.FOLD_MASK, slope=SLOPE)
            return d_src, dw

        d_pre3, dw4 = down_bwd(d_x4, 4, cat1, 512, 512, ops_d[4][1], d_cat1, 512)
        d_pre2, dw3 = down_bwd(d_pre3, 3, cat2, 256, 256, ops_d[3][1], d_cat2, 256)
        d_pre1, dw2 = down_bwd(d_pre2, 2, cat3, 128, 128, ops_d[2][1], d_cat3, 128)
        d_pre0, dw1 = down_bwd(d_pre1, 1, cat4, 64, 64, ops_d[1][1], d_cat4, 64)
        d_a0, dw0b = down_bwd(d_pre0, 0, a0, 0, 64, ops_d[0][1], None, 0)

        # conv0[0] (1 -> 64): weight gradient, and the image gradient as a 64 -> 1 convolution with the flipped filter
        dw0a = None
        if need_w:
            dw0a = torch.empty(64, 1, 3, 3, device=dev, dtype=f32)
            db0 = torch.empty(64, device=dev, dtype=f32)
            capi.conv_in1_wgrad(xf, d_a0, dw0a, db0, B, Hs[0], Ws[0], 64, 64)
        dx = None
        if need_x:
            w_flip = w0a.flip(2, 3).permute(1, 0, 2, 3).contiguous()   # [1,64,3,3]
            dx = torch.empty(B, 1, Hs[0], Ws[0], device=dev, dtype=f32)
            capi.conv_out1_fwd(d_a0, w_flip, torch.zeros(1, device=dev, dtype=f32), dx, B, Hs[0], Ws[0], 64)
            dx = dx.to(x_dtype)
        grads = [dw0a, dw0b, dw1, dw2, dw3, dw4, dw_u1, dw_u2, dw_u3, dw_u4, dw_f1, dw_f2 if need_w else None]
        grads = [g if n else None for g, n in zip(grads, need[2:])]
        sn_ctx, w_orig = ctx.sn
        if sn_ctx is not None and need_w:
            # through W / sigma(W) back to weight_orig, in place on the gradients just computed
            us, vs, sigmas = sn_ctx
            capi.spectral_norm_bwd(capi.sn_layers(w_orig, us, vs, SN_DIMS, sigmas), grads, grads, dev)
        ctx.acts = ctx.patches = ctx.sn = None
        return (dx, None) + tuple(grads)


def unet_discriminator(x, weights):
    """weights: the 12 (already spectrally normalised) conv weights in forward order."""
    if not x.is_cuda:
        raise capi.SrkError("UNetDiscriminatorSN runs on CUDA (sm_100a) only; there is no CPU path")
    return UNetDiscriminatorFunction.apply(x, None, *weights)


def unet_discriminator_sn(x, weight_orig, weight_u, weight_v, training: bool, eps: float = 1e-12):
    """The module path: `weight_orig` parameters + the spectral-norm buffers, normalised on libsrk (srk_spectral_norm); u and v
    are updated in place when `training`, exactly as torch.nn.utils.spectral_norm's forward pre-hook does."""
    if not x.is_cuda:
        raise capi.SrkError("UNetDiscriminatorSN runs on CUDA (sm_100a) only; there is no CPU path")
    sn = dict(u=list(weight_u), v=list(weight_v), training=bool(training), eps=float(eps))
    return UNetDiscriminatorFunction.apply(x, sn, *weight_orig)


# ======================================================================================================================
# models/discriminator_hat.py:8-49 — the Real-ESRGAN style U-Net discriminator `train_hat.py:26,138` builds: three 4x4
# stride-2 spectral-norm convolutions down, then bilinear x2 + 3x3 spectral-norm convolution three times up with additive
# skips, two more 3x3 layers and a 64 -> 1 head; conv0 (1 -> 64) and conv9 (64 -> 1) are plain convolutions with bias.
# ======================================================================================================================
_HAT_SN = tuple(range(1, 9))          # conv1 .. conv8 carry spectral_norm (all nn.Conv2d: dim 0)


def _prep3(w):
    """3x3 weight [Cout,Cin,3,3] fp32 -> (forward operand, input-gradient operand) of srk_conv3x3_igemm."""
    co, ci = w.shape[0], w.shape[1]
    wf = torch.empty(co * 9 * ci, device=w.device, dtype=BF16)
    wt = torch.empty(ci * 9 * co, device=w.device, dtype=BF16)
    bias = torch.empty(co, device=w.device, dtype=torch.float32)
    capi.conv3x3_prep_weights(w, None, co, ci, False, wf, wt, bias)
    return wf, wt, bias


class UNetDiscriminatorHatFunction(torch.autograd.Function):
    """x [B,1,H,W] -> logits [B,1,H,W] (models/discriminator_hat.py:25-49)."""

    @staticmethod
    def forward(ctx, x, sn, skip, b0, b9, *weights):
        B, cin, H, W = x.shape
        if len(weights) != 10:
            raise capi.SrkError("discriminator_hat.UNetDiscriminatorSN has 10 convolutions")
        dev = x.device
        f32 = torch.float32
        weights = [w.detach().float().contiguous() for w in weights]
        if cin != 1 or weights[0].shape[0] != 64 or weights[9].shape[0] != 1:
            raise capi.SrkError("discriminator kernels are specialised for num_in_ch=1, num_feat=64 (the scripts' only configuration)")
        if H % 32 or W % 64:
            raise capi.SrkError("discriminator_hat.UNetDiscriminatorSN: H must be a multiple of 32 and W of 64 (tile sizes of the 3x3 "
                                "implicit-GEMM kernel at a quarter of the resolution)")
        sig = [None] * 10
        wn = list(weights)                 # normalised fp32 filters of the 3x3 layers (conv4..conv8)
        sn_ctx = None
        if sn is not None:
            sigmas = torch.empty(8, device=dev, dtype=f32)
            w_sn = [torch.empty_like(weights[i]) if i >= 4 else None for i in _HAT_SN]
            capi.spectral_norm(capi.sn_layers([weights[i] for i in _HAT_SN], sn["u"], sn["v"], (0,) * 8, sigmas, w_sn),
                               sn["training"], sn["eps"], dev)
            for k, i in enumerate(_HAT_SN):
                sig[i] = sigmas[k:k + 1]
                if i >= 4:
                    wn[i] = w_sn[k]
            if any(ctx.needs_input_grad[5:]):
                sizes = [t.numel() for t in sn["u"]] + [t.numel() for t in sn["v"]]
                parts = list(torch.split(torch.cat([t.reshape(-1) for t in list(sn["u"]) + list(sn["v"])]), sizes))
                sn_ctx = (parts[:8], parts[8:], sigmas)
        Hs = [H >> i for i in range(4)]
        Ws = [W >> i for i in range(4)]
        M = [B * Hs[i] * Ws[i] for i in range(4)]
        V = capi.view
        xf = x.detach().contiguous().float().reshape(B, H, W)
        keep = any(ctx.needs_input_grad[5:])

        a0 = _alloc(M[0], 64, dev)                                                     # x0 (:26)
        capi.conv_in1_fwd(xf, weights[0], b0.detach().float().contiguous(), a0, B, H, W, 64, 64)
        capi.view_lrelu(V(a0), M[0], SLOPE)
        ops = {i: _ops4(weights[i], sig[i]) for i in (1, 2, 3)}                        # (Wf, Wt)
        patches = {}

        def down(src, lvl, i):
            cin_, cout = src.shape[1], ops[i][0].shape[0]
            p = _alloc(M[lvl + 1], 16 * cin_, dev)
            capi.disc_patches_k4s2(V(src), None, SLOPE, B, Hs[lvl], Ws[lvl], p)
            dst = _alloc(M[lvl + 1], cout, dev)
            capi.gemm_tn_lrelu(p, ops[i][0], dst, SLOPE)
            if keep:
                patches[i] = p
            return dst

        x1 = down(a0, 0, 1)            # :27
        x2 = down(x1, 1, 2)
        x3 = down(x2, 2, 3)            # :29
        u3 = _alloc(M[2], 512, dev)
        capi.bilinear2x_fwd(V(x3), None, V(u3), B, Hs[3], Ws[3])                        # :31
        # conv4 512 -> 256: the implicit-GEMM kernel takes <= 256 input channels, so the K dimension is walked in two halves
        w4a, w4b = wn[4][:, :256].contiguous(), wn[4][:, 256:].contiguous()
        p4a, p4b = _prep3(w4a), _prep3(w4b)
        tmp = _alloc(M[2], 256, dev)
        act4 = _alloc(M[2], 256, dev)
        capi.conv3x3_igemm_v(capi.CEPI_BIAS, B, Hs[2], Ws[2], 256, 256, 256, V(u3, 0, 256), p4a[0], p4a[2], V(tmp))
        capi.conv3x3_igemm_v(capi.CEPI_BIAS_RES, B, Hs[2], Ws[2], 256, 256, 256, V(u3, 256, 256), p4b[0], p4b[2], V(act4), r=V(tmp))
        del tmp
        capi.view_lrelu(V(act4), M[2], SLOPE)                                           # :32
        u4 = _alloc(M[1], 256, dev)
        capi.bilinear2x_fwd(V(act4), V(x2) if skip else None, V(u4), B, Hs[2], Ws[2])   # :34-36
        p5 = _prep3(wn[5])
        act5 = _alloc(M[1], 128, dev)
        capi.conv3x3_igemm_v(capi.CEPI_BIAS, B, Hs[1], Ws[1], 256, 128, 128, V(u4), p5[0], p5[2], V(act5))
        capi.view_lrelu(V(act5), M[1], SLOPE)                                           # :37
        u5 = _alloc(M[0], 128, dev)
        capi.bilinear2x_fwd(V(act5), V(x1) if skip else None, V(u5), B, Hs[1], Ws[1])   # :39-41
        p6, p7, p8 = _prep3(wn[6]), _prep3(wn[7]), _prep3(wn[8])
        act6 = _alloc(M[0], 64, dev)
        capi.conv3x3_igemm(capi.CEPI_BIAS_LRELU, B, H, W, 128, 64, 64, u5, p6[0], p6[2], act6, slope=SLOPE)   # :42
        x6 = act6
        if skip:                                                                        # :44-45
            x6 = _alloc(M[0], 64, dev)
            capi.view_axpy(V(x6), V(act6), V(a0), M[0], 1.0)
        act7 = _alloc(M[0], 64, dev)
        capi.conv3x3_igemm(capi.CEPI_BIAS_LRELU, B, H, W, 64, 64, 64, x6, p7[0], p7[2], act7, slope=SLOPE)    # :47
        act8 = _alloc(M[0], 64, dev)
        capi.conv3x3_igemm(capi.CEPI_BIAS_LRELU, B, H, W, 64, 64, 64, act7, p8[0], p8[2], act8, slope=SLOPE)  # :48
        out = torch.empty(B, 1, H, W, device=dev, dtype=f32)
        capi.conv_out1_fwd(act8, weights[9], b9.detach().float().contiguous(), out, B, H, W, 64)             # :49
        if any(ctx.needs_input_grad):
            ctx.acts = (xf, a0, x1, x2, x3, u3, act4, u4, act5, u5, act6, x6, act7, act8)
            ctx.ops = (ops, p4a[1], p4b[1], p5[1], p6[1], p7[1], p8[1], weights[0], weights[9])
            ctx.patches = patches
            ctx.sn = (sn_ctx, [weights[i] for i in _HAT_SN] if sn_ctx is not None else None)
            ctx.meta = (B, Hs, Ws, M, x.dtype, skip)
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.acts is None:
            raise capi.SrkError("UNetDiscriminatorHatFunction: second backward through the same forward (activations were released)")
        xf, a0, x1, x2, x3, u3, act4, u4, act5, u5, act6, x6, act7, act8 = ctx.acts
        ops, wt4a, wt4b, wt5, wt6, wt7, wt8, w0, w9 = ctx.ops
        patches = ctx.patches
        B, Hs, Ws, M, x_dtype, skip = ctx.meta
        need = ctx.needs_input_grad
        need_x, need_w = need[0], any(need[3:])
        dev = dout.device
        f32 = torch.float32
        V = capi.view
        H, W = Hs[0], Ws[0]
        dout = dout.contiguous().float()

        def masked(g, f, n):
            """g * (f > 0 ? 1 : slope) into a fresh buffer (g itself is also the skip connection's gradient)"""
            o = g.clone()
            for c0 in range(0, o.shape[1], 256):
                c = min(256, o.shape[1] - c0)
                capi.view_lrelu_mask(V(o, c0, c), V(f, c0, c), n, SLOPE)
            return o

        def wgrad3(dy, xin, cin_, cout):
            dw = torch.empty(cout, cin_, 3, 3, device=dev, dtype=f32)
            capi.conv3x3_wgrad(B, H, W, cin_, cout, cin_, cout, False, dy, xin, dw)
            return dw

        # conv9, conv8, conv7 (:47-49)
        d8 = _alloc(M[0], 64, dev)
        dw9 = torch.empty(1, 64, 3, 3, device=dev, dtype=f32)
        db9 = torch.empty(1, device=dev, dtype=f32)
        capi.conv_out1_bwd(dout, act8, w9, d8, dw9, db9, B, H, W, 64)
        capi.view_lrelu_mask(V(d8), V(act8), M[0], SLOPE)
        dw8 = wgrad3(d8, act7, 64, 64) if need_w else None
        d7 = _alloc(M[0], 64, dev)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, 64, 64, 64, d8, wt8, None, d7)
        capi.view_lrelu_mask(V(d7), V(act7), M[0], SLOPE)
        dw7 = wgrad3(d7, x6, 64, 64) if need_w else None
        d_x6 = _alloc(M[0], 64, dev)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, 64, 64, 64, d7, wt7, None, d_x6)
        del d7, d8
        # conv6 (:42) — x6 = act6 + x0: d_x6 is the gradient of both
        d6 = masked(d_x6, act6, M[0])
        dw6 = wgrad3(d6, u5, 128, 64) if need_w else None
        d_u5 = _alloc(M[0], 128, dev)
        capi.conv3x3_igemm(capi.CEPI_BIAS, B, H, W, 64, 128, 128, d6, wt6, None, d_u5)
        del d6
        d_s5 = _alloc(M[1], 128, dev)                   # gradient of (act5 + x1)
        capi.bilinear2x_bwd(V(d_u5), V(d_s5), B, Hs[1], Ws[1])
        del d_u5
        # conv5 (:37)
        d5 = masked(d_s5, act5, M[1])
        dw5 = None
        if need_w:
            dw5 = torch.empty(128, 256, 3, 3, device=dev, dtype=f32)
            capi.conv3x3_wgrad(B, Hs[1], Ws[1], 256, 128, 256, 128, False, d5, u4, dw5)
        d_u4 = _alloc(M[1], 256, dev)
        capi.conv3x3_igemm_v(capi.CEPI_BIAS, B, Hs[1], Ws[1], 128, 256, 256, V(d5), wt5, None, V(d_u4))
        del d5
        d_s4 = _alloc(M[2], 256, dev)                   # gradient of (act4 + x2)
        capi.bilinear2x_bwd(V(d_u4), V(d_s4), B, Hs[2], Ws[2])
        del d_u4
        # conv4 (:32), two 256-channel halves of its 512 input channels
        d4 = masked(d_s4, act4, M[2])
        dw4 = None
        if need_w:
            halves = []
            for c0 in (0, 256):
                h = torch.empty(256, 256, 3, 3, device=dev, dtype=f32)
                capi.conv3x3_wgrad_v(B, Hs[2], Ws[2], 256, 256, 256, 256, V(d4), V(u3, c0, 256), h)
                halves.append(h)
            dw4 = torch.cat(halves, 1)
        d_u3 = _alloc(M[2], 512, dev)
        capi.conv3x3_igemm_v(capi.CEPI_BIAS, B, Hs[2], Ws[2], 256, 256, 256, V(d4), wt4a, None, V(d_u3, 0, 256))
        capi.conv3x3_igemm_v(capi.CEPI_BIAS, B, Hs[2], Ws[2], 256, 256, 256, V(d4), wt4b, None, V(d_u3, 256, 256))
        del d4
        d3 = _alloc(M[3], 512, dev)
        capi.bilinear2x_bwd(V(d_u3), V(d3), B, Hs[3], Ws[3])
        del d_u3
        for c0 in (0, 256):                              # LeakyReLU backward of conv3's output
            capi.view_lrelu_mask(V(d3, c0, 256), V(x3, c0, 256), M[3], SLOPE)

        # encoder: d_pre of level lvl+1 -> masked gradient at level lvl (+ the skip connection's gradient)
        def down_bwd(d_pre, lvl, i, src, add):
            cin_ = src.shape[1]
            dw = None
            if need_w:
                dw = _wgrad(patches[i], d_pre, cin_)
                patches[i] = None
            taps = torch.empty(d_pre.shape[0], 16 * cin_, device=dev, dtype=BF16)
            capi.gemm_tn(capi.EPI_STORE, d_pre, ops[i][1], taps)
            d_src = _alloc(M[lvl], cin_, dev)
            capi.disc_fold_k4s2(taps, B, Hs[lvl + 1], Ws[lvl + 1], V(d_src), add=None if add is None else V(add), f=V(src),
                                act=capi.FOLD_MASK, slope=SLOPE)
            return d_src, dw

        d2, dw3 = down_bwd(d3, 2, 3, x2, d_s4 if skip else None)
        d1, dw2 = down_bwd(d2, 1, 2, x1, d_s5 if skip else None)
        d0, dw1 = down_bwd(d1, 0, 1, a0, d_x6 if skip else None)
        # conv0 (:26): weight and bias gradient, image gradient as a 64 -> 1 convolution with the flipped filter
        dw0 = db0 = None
        if need_w or need[3]:
            dw0 = torch.empty(64, 1, 3, 3, device=dev, dtype=f32)
            db0 = torch.empty(64, device=dev, dtype=f32)
            capi.conv_in1_wgrad(xf, d0, dw0, db0, B, H, W, 64, 64)
        dx = None
        if need_x:
            w_flip = w0.flip(2, 3).permute(1, 0, 2, 3).contiguous()
            dx = torch.empty(B, 1, H, W, device=dev, dtype=f32)
            capi.conv_out1_fwd(d0, w_flip, torch.zeros(1, device=dev, dtype=f32), dx, B, H, W, 64)
            dx = dx.to(x_dtype)
        grads = [dw0, dw1, dw2, dw3, dw4, dw5, dw6, dw7, dw8, dw9]
        grads = [g if n else None for g, n in zip(grads, need[5:])]
        sn_ctx, w_orig = ctx.sn
        if sn_ctx is not None and need_w:
            us, vs, sigmas = sn_ctx
            sub = [grads[i] for i in _HAT_SN]
            capi.spectral_norm_bwd(capi.sn_layers(w_orig, us, vs, (0,) * 8, sigmas), sub, sub, dev)
        ctx.acts = ctx.patches = ctx.sn = None
        return (dx, None, None, db0 if need[3] else None, db9 if need[4] else None) + tuple(grads)


def unet_discriminator_hat(x, weights, b0, b9, skip=True):
    """weights: conv0 .. conv9 with conv1 .. conv8 already spectrally normalised."""
    if not x.is_cuda:
        raise capi.SrkError("UNetDiscriminatorSN runs on CUDA (sm_100a) only; there is no CPU path")
    return UNetDiscriminatorHatFunction.apply(x, None, bool(skip), b0, b9, *weights)


def unet_discriminator_hat_sn(x, weights, b0, b9, weight_u, weight_v, training: bool, eps: float = 1e-12, skip=True):
    """weights: conv0.weight, conv1.weight_orig .. conv8.weight_orig, conv9.weight; weight_u / weight_v of conv1 .. conv8."""
    if not x.is_cuda:
        raise capi.SrkError("UNetDiscriminatorSN runs on CUDA (sm_100a) only; there is no CPU path")
    sn = dict(u=list(weight_u), v=list(weight_v), training=bool(training), eps=float(eps))
    return UNetDiscriminatorHatFunction.apply(x, sn, bool(skip), b0, b9, *weights)
