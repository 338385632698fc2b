"""UNetDiscriminatorSN forward / backward on libsrk (models/discriminator_swin.py:43-84).

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

Spectral normalisation (torch.nn.utils.spectral_norm, :10,25,49-52,67-69) is parameter preparation: the caller (gan.py) runs
the reference's own power-iteration hook and hands the normalised weights in; autograd carries the weight gradients
returned here back through W / sigma to `weight_orig`.

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
