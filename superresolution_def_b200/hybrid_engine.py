"""Host-side engine for the hybrid generator's convolutional part on libsrk: conv_adapt, the RRDB trunk (residual dense
blocks with *virtual* concatenation), conv_body + skip, nearest x2, conv_up / conv_hr / conv_last.

Reference being replaced: models/hybridmodels_hat.py — ResidualDenseBlock.forward :38-44, RRDBBlock.forward :54-58,
HybridHATRealESRGAN.forward :117-131 (everything after `self.hat(x)`).

Layout: NHWC bf16, pixels x channels.  Each residual dense block owns ONE buffer `cat` [pixels, nf + 4*gc]; channels
[0, nf) are the block input, [nf + (k-1)*gc, nf + k*gc) the output of conv_k.  conv_k reads the channel slice
[0, nf + (k-1)*gc) — exactly torch.cat((x, x1, ..)) of the reference — and writes its slice in place: the five
concatenations per block (180 per forward at the script's 12 RRDBs) are never materialised.  The implicit-GEMM kernel
(csrc/conv3x3.cuh) clips its 64-channel TMA boxes to the slice, LeakyReLU(0.2), bias and the 0.2-scaled residual are
epilogues, and conv5 stores straight into channels [0, nf) of the NEXT block's buffer.  The backward mirrors this with
one gradient buffer `dcat`: the input-gradient convolutions accumulate into channel slices in place.
There is no CPU / PyTorch fallback.
"""
from __future__ import annotations

import torch

from . import _capi as capi
from . import conv_engine as cv

BF16 = torch.bfloat16
V = capi.view
SLOPE = 0.2   # nn.LeakyReLU(negative_slope=0.2), hybridmodels_hat.py:29,95
RES = 0.2     # residual scaling `x5 * 0.2 + x`, :44 and :58


def _p64(c: int) -> int:
    return (c + 63) // 64 * 64


def _check_feat(nf: int, gc: int):
    if nf % 8 or gc % 8 or nf + 4 * gc > 256 or nf > 64 or gc > 64:
        raise capi.SrkError(f"libsrk dense blocks need num_feat % 8 == 0, num_grow_ch % 8 == 0 (16-byte channel slices), "
                            f"num_feat, num_grow_ch <= 64 and num_feat + 4*num_grow_ch <= 256; got {nf}, {gc}")


def _conv(epi, B, H, W, w, b, x, y, cin, cout, r=None, slope=SLOPE, alpha=1.0):
    """y = epilogue(conv3x3(x)) with the layer's forward operand (cached per parameter, refreshed every forward)."""
    wf, _, bp = cv.conv_weights(w, b, _p64(cout), _p64(cin))
    capi.conv3x3_igemm_v(epi, B, H, W, _p64(cin), _p64(cout), cout, x, wf, bp, y, r, slope=slope, alpha=alpha)


def _dgrad(epi, B, H, W, w, b, dy, dx, cin, cout, r=None, slope=SLOPE):
    """dx = epilogue(conv3x3(dy, flipped/transposed weights)): the input gradient of a layer prepared by _conv."""
    _, wt, _ = cv.conv_weights(w, b, _p64(cout), _p64(cin), refresh=False)
    capi.conv3x3_igemm_v(epi, B, H, W, _p64(cout), _p64(cin), cin, dy, wt, None, dx, r, slope=slope, alpha=1.0)


def _wgrads(B, H, W, w, b, dy, x, cin, cout, db=None):
    """Weight and bias gradient of one layer; db: already computed (fused into the LeakyReLU-mask pass) or None."""
    dw = torch.empty_like(w)
    capi.conv3x3_wgrad_v(B, H, W, cin, cout, _p64(cin), _p64(cout), dy, x, dw)
    if db is None:
        db = torch.empty_like(b)
        capi.bias_grad_v(dy, B * H * W, db)
    return dw, db


def rdb_fwd(geom, nf, gc, cat, out_view, p):
    """One residual dense block.  cat[:, :nf] holds the input; out_view receives conv5 * 0.2 + input.  p: 10 tensors."""
    B, H, W = geom
    for k in range(4):
        cin = nf + k * gc
        _conv(capi.CEPI_BIAS_LRELU, B, H, W, p[2 * k], p[2 * k + 1], V(cat, 0, cin), V(cat, cin, gc), cin, gc)
    cc = nf + 4 * gc
    _conv(capi.CEPI_BIAS_RES, B, H, W, p[8], p[9], V(cat, 0, cc), out_view, cc, nf, r=V(cat, 0, nf), alpha=RES)


def rdb_bwd(geom, nf, gc, cat, dcat, g5, p):
    """dcat[:, :nf] holds the gradient of the block output on entry and of the block input on exit.  Returns the 10
    parameter gradients (conv1.weight, conv1.bias, ..., conv5.bias)."""
    B, H, W = geom
    T = B * H * W
    cc = nf + 4 * gc
    grads = [None] * 10
    capi.view_axpy(V(g5), V(dcat, 0, nf), None, T, RES)                                   # d(conv5) = 0.2 * dy
    grads[8], grads[9] = _wgrads(B, H, W, p[8], p[9], V(g5), V(cat, 0, cc), cc, nf)
    # d(cat) = conv5^T(g5) (+ dy on the input slice: the residual path)
    _dgrad(capi.CEPI_BIAS_RES, B, H, W, p[8], p[9], V(g5), V(dcat, 0, cc), cc, nf, r=V(dcat, 0, nf))
    for k in (3, 2, 1, 0):
        s = nf + k * gc
        db = torch.empty_like(p[2 * k + 1])
        capi.view_lrelu_mask(V(dcat, s, gc), V(cat, s, gc), T, SLOPE, colsum=db)           # through LeakyReLU of conv_{k+1}
        grads[2 * k], grads[2 * k + 1] = _wgrads(B, H, W, p[2 * k], p[2 * k + 1], V(dcat, s, gc), V(cat, 0, s), s, gc, db=db)
        _dgrad(capi.CEPI_BIAS_RES, B, H, W, p[2 * k], p[2 * k + 1], V(dcat, s, gc), V(dcat, 0, s), s, gc, r=V(dcat, 0, s))
    return grads


def trunk_fwd(geom, nf, gc, x0_cat, rdb_params, group3, dev):
    """Chain of residual dense blocks.  x0_cat: the first block's buffer with its input in [:, :nf].  group3: blocks come
    in RRDB groups of three with the outer `out * 0.2 + x` (:58).  Returns (cats, out [T, nf])."""
    B, H, W = geom
    T = B * H * W
    cc = nf + 4 * gc
    n = len(rdb_params)
    cats = [x0_cat] + [torch.empty(T, cc, device=dev, dtype=BF16) for _ in range(n - 1)]
    out = torch.empty(T, nf, device=dev, dtype=BF16)
    t3 = torch.empty(T, nf, device=dev, dtype=BF16) if group3 else None
    for j, p in enumerate(rdb_params):
        last = j == n - 1
        if group3 and j % 3 == 2:
            rdb_fwd(geom, nf, gc, cats[j], V(t3), p)
            dst = V(out) if last else V(cats[j + 1], 0, nf)
            capi.view_axpy(dst, V(t3), V(cats[j - 2], 0, nf), T, RES)                      # RRDB: out * 0.2 + x
        else:
            rdb_fwd(geom, nf, gc, cats[j], V(out) if last else V(cats[j + 1], 0, nf), p)
    return cats, out


def trunk_bwd(geom, nf, gc, cats, rdb_params, group3, G):
    """G [T, nf]: gradient of the trunk output on entry, of the trunk input on exit.  Returns per-block gradient lists."""
    B, H, W = geom
    T = B * H * W
    dev = G.device
    cc = nf + 4 * gc
    dcat = torch.empty(T, cc, device=dev, dtype=BF16)
    g5 = torch.empty(T, nf, device=dev, dtype=BF16)
    n = len(rdb_params)
    grads = [None] * n
    if group3:
        for i in reversed(range(n // 3)):
            capi.view_axpy(V(dcat, 0, nf), V(G), None, T, RES)                              # into rdb3: 0.2 * G
            for j in (3 * i + 2, 3 * i + 1, 3 * i):
                grads[j] = rdb_bwd(geom, nf, gc, cats[j], dcat, g5, rdb_params[j])
            capi.view_axpy(V(G), V(dcat, 0, nf), V(G), T, 1.0)                              # + identity path
    else:
        capi.view_axpy(V(dcat, 0, nf), V(G), None, T, 1.0)
        for j in reversed(range(n)):
            grads[j] = rdb_bwd(geom, nf, gc, cats[j], dcat, g5, rdb_params[j])
        capi.view_axpy(V(G), V(dcat, 0, nf), None, T, 1.0)
    return grads


class DenseTrunkFunction(torch.autograd.Function):
    """Stand-alone ResidualDenseBlock (n = 1, group3 False) / RRDBBlock (n = 3, group3 True) / nn.Sequential of RRDBs on
    an NCHW tensor; NCHW <-> NHWC packing is torch indexing, all arithmetic is libsrk."""

    @staticmethod
    def forward(ctx, x, nf, gc, group3, *params):
        if not x.is_cuda:
            raise capi.SrkError("libsrk dense blocks run on CUDA only")
        _check_feat(nf, gc)
        B, C, H, W = x.shape
        if C != nf or H % 8 or W % 16:
            raise capi.SrkError("libsrk dense blocks: input channels == num_feat, H % 8 == 0, W % 16 == 0")
        T = B * H * W
        rdb_params = [params[10 * j:10 * j + 10] for j in range(len(params) // 10)]
        cat0 = torch.empty(T, nf + 4 * gc, device=x.device, dtype=BF16)
        cat0[:, :nf] = x.detach().permute(0, 2, 3, 1).reshape(T, nf).to(BF16)
        cats, out = trunk_fwd((B, H, W), nf, gc, cat0, [[t.detach() for t in p] for p in rdb_params], group3, x.device)
        ctx.saved = cats
        ctx.meta = (B, H, W, nf, gc, group3, x.dtype)
        ctx.params = params
        return out.reshape(B, H, W, nf).permute(0, 3, 1, 2).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        B, H, W, nf, gc, group3, dtype = ctx.meta
        T = B * H * W
        params = ctx.params
        rdb_params = [[t.detach() for t in params[10 * j:10 * j + 10]] for j in range(len(params) // 10)]
        G = dy.permute(0, 2, 3, 1).reshape(T, nf).to(BF16).contiguous()
        grads = trunk_bwd((B, H, W), nf, gc, ctx.saved, rdb_params, group3, G)
        ctx.saved = None
        dx = G.reshape(B, H, W, nf).permute(0, 3, 1, 2).to(dtype)
        return (dx, None, None, None, *[g for gl in grads for g in gl])


class HybridTailFunction(torch.autograd.Function):
    """Everything of HybridHATRealESRGAN.forward after the HAT stage (hybridmodels_hat.py:120-131), one autograd node.
    params: conv_adapt (w, b), 10 per residual dense block in module order, conv_body, conv_up, conv_hr, conv_last."""

    @staticmethod
    def forward(ctx, hat_out, nf, gc, *params):
        if not hat_out.is_cuda:
            raise capi.SrkError("libsrk hybrid generator runs on CUDA only")
        _check_feat(nf, gc)
        B, cin, H, W = hat_out.shape
        if cin != 1 or H % 8 or W % 16:
            raise capi.SrkError("libsrk hybrid tail: single-channel HAT output with H % 8 == 0 and W % 16 == 0")
        dev = hat_out.device
        T = B * H * W
        geom, geom2 = (B, H, W), (B, 2 * H, 2 * W)
        p = [t.detach() for t in params]
        n_rdb = (len(p) - 10) // 10
        wa, ba = p[0], p[1]
        rdb_params = [p[2 + 10 * j:12 + 10 * j] for j in range(n_rdb)]
        wb, bb, wu, bu, wh, bh, wl, bl = p[2 + 10 * n_rdb:]
        img8 = torch.empty(T, 8, device=dev, dtype=BF16)
        capi.img1_pack(hat_out.detach().contiguous().float(), img8)
        cat0 = torch.empty(T, nf + 4 * gc, device=dev, dtype=BF16)
        _conv(capi.CEPI_BIAS_LRELU, B, H, W, wa, ba, V(img8), V(cat0, 0, nf), 1, nf)        # feat = lrelu(conv_adapt(hat_out))
        cats, body = trunk_fwd(geom, nf, gc, cat0, rdb_params, True, dev)
        fsum = torch.empty(T, nf, device=dev, dtype=BF16)
        _conv(capi.CEPI_BIAS_RES, B, H, W, wb, bb, V(body), V(fsum), nf, nf, r=V(cat0, 0, nf))   # trunk_feat + conv_body(.)
        up = torch.empty(4 * T, nf, device=dev, dtype=BF16)
        capi.nearest2_fwd(V(fsum), V(up), B, H, W)
        u = torch.empty(4 * T, nf, device=dev, dtype=BF16)
        _conv(capi.CEPI_BIAS_LRELU, B, 2 * H, 2 * W, wu, bu, V(up), V(u), nf, nf)
        hr = torch.empty(4 * T, nf, device=dev, dtype=BF16)
        _conv(capi.CEPI_BIAS_LRELU, B, 2 * H, 2 * W, wh, bh, V(u), V(hr), nf, nf)
        out = torch.empty(B, 1, 2 * H, 2 * W, device=dev, dtype=torch.float32)
        wf_l, _, bp_l = cv.conv_weights(wl, bl, 16, 64)     # N = 16 instance: one real output column, fp32 from TMEM
        capi.conv3x3_igemm_v(capi.CEPI_OUT1, B, 2 * H, 2 * W, 64, 16, 1, V(hr), wf_l, bp_l, None, y32=out)
        if any(ctx.needs_input_grad):
            ctx.saved = (img8, cats, body, up, u, hr)
            ctx.meta = (geom, geom2, nf, gc, n_rdb, hat_out.dtype)
            ctx.params = params
        return out

    @staticmethod
    def backward(ctx, dout):
        img8, cats, body, up, u, hr = ctx.saved
        (B, H, W), _, nf, gc, n_rdb, in_dtype = ctx.meta
        params = ctx.params
        p = [t.detach() for t in params]
        dev = dout.device
        T = B * H * W
        H2, W2 = 2 * H, 2 * W
        wa, ba = p[0], p[1]
        rdb_params = [p[2 + 10 * j:12 + 10 * j] for j in range(n_rdb)]
        wb, bb, wu, bu, wh, bh, wl, bl = p[2 + 10 * n_rdb:]
        dout8 = torch.empty(4 * T, 8, device=dev, dtype=BF16)
        capi.img1_pack(dout.contiguous().float(), dout8)
        # conv_last (nf -> 1) backward through conv_hr's LeakyReLU; operands prepared here (forward used the N = 16 form)
        cv.conv_weights(wl, bl, 64, 64)
        d_hr = torch.empty(4 * T, nf, device=dev, dtype=BF16)
        _dgrad(capi.CEPI_MASK_LRELU, B, H2, W2, wl, bl, V(dout8), V(d_hr), nf, 1, r=V(hr))
        dwl, dbl = _wgrads(B, H2, W2, wl, bl, V(dout8), V(hr), nf, 1)
        # conv_hr
        d_u = torch.empty(4 * T, nf, device=dev, dtype=BF16)
        _dgrad(capi.CEPI_MASK_LRELU, B, H2, W2, wh, bh, V(d_hr), V(d_u), nf, nf, r=V(u))
        dwh, dbh = _wgrads(B, H2, W2, wh, bh, V(d_hr), V(u), nf, nf)
        # conv_up and the nearest x2
        d_up = d_hr   # reuse
        _dgrad(capi.CEPI_BIAS, B, H2, W2, wu, bu, V(d_u), V(d_up), nf, nf)
        dwu, dbu = _wgrads(B, H2, W2, wu, bu, V(d_u), V(up), nf, nf)
        d_fsum = torch.empty(T, nf, device=dev, dtype=BF16)
        capi.nearest2_bwd(V(d_up), V(d_fsum), B, H, W)
        # conv_body (+ skip: d_fsum also reaches the trunk input)
        G = torch.empty(T, nf, device=dev, dtype=BF16)
        _dgrad(capi.CEPI_BIAS, B, H, W, wb, bb, V(d_fsum), V(G), nf, nf)
        dwb, dbb = _wgrads(B, H, W, wb, bb, V(d_fsum), V(body), nf, nf)
        rgrads = trunk_bwd((B, H, W), nf, gc, cats, rdb_params, True, G)
        capi.view_axpy(V(G), V(G), V(d_fsum), T, 1.0)
        # conv_adapt (1 -> nf) through its LeakyReLU
        dba = torch.empty_like(ba)
        capi.view_lrelu_mask(V(G), V(cats[0], 0, nf), T, SLOPE, colsum=dba)
        dwa, dba = _wgrads(B, H, W, wa, ba, V(G), V(img8), 1, nf, db=dba)
        d_hat = None
        if ctx.needs_input_grad[0]:
            dimg8 = torch.empty(T, 8, device=dev, dtype=BF16)
            _dgrad(capi.CEPI_BIAS, B, H, W, wa, ba, V(G), V(dimg8), 1, nf)
            d_hat = torch.empty(B, 1, H, W, device=dev, dtype=torch.float32)
            capi.img1_unpack(dimg8, d_hat)
            d_hat = d_hat.to(in_dtype)
        ctx.saved = None
        flat = [dwa, dba] + [g for gl in rgrads for g in gl] + [dwb, dbb, dwu, dbu, dwh, dbh, dwl, dbl]
        return (d_hat, None, None, *flat)
