"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (imported from /root/reference) on
seeded inputs.  Run in the build container only (the reference does not travel to the GPU box):

    python tools/make_golden.py

Each fixture holds: the constructor kwargs, the full state_dict, the input, the output, and the gradients of
sum(output * w) w.r.t. the input and every parameter (fp32, CPU).  Dimensions are kept tiny: the oracle
(oracle/swinir_oracle.py) is dimension-generic, so small shapes pin it as firmly as large ones.
"""
import os
import sys

import torch

REF = os.environ.get("SR_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
from tests.util import randomize_  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def grads_of(module, out, w, x):
    (out * w).sum().backward()
    g = {n: p.grad.clone() for n, p in module.named_parameters()}
    return g, x.grad.clone()


def swin_fixtures():
    from models.architecture_swin import SwinIR, SwinTransformerBlock, WindowAttention, Mlp, window_partition, window_reverse
    torch.manual_seed(11)
    # window helpers
    x = torch.randn(2, 8, 12, 5)
    win = window_partition(x, 4)
    torch.save({"x": x, "ws": 4, "windows": win, "reversed": window_reverse(win, 4, 8, 12)}, f"{OUT}/swin_window_helpers.pt")
    # attention (with and without an explicit mask: the module supports one although SwinIR never passes it)
    att = randomize_(WindowAttention(24, (4, 4), 3), seed=1)
    xa = torch.randn(6, 16, 24, requires_grad=True)
    mask = torch.where(torch.rand(3, 16, 16) > 0.7, -100.0, 0.0)
    ya = att(xa)
    ga, gxa = grads_of(att, ya, torch.randn_like(ya), xa)
    torch.save({"kw": dict(dim=24, window=4, heads=3), "sd": att.state_dict(), "x": xa.detach(), "y": ya.detach(),
                "y_masked": att(xa, mask).detach(), "mask": mask, "grads": ga, "gx": gxa,
                "w_seed": 0}, f"{OUT}/swin_window_attention.pt")
    # mlp
    m = randomize_(Mlp(24, 96), seed=2)
    xm = torch.randn(3, 10, 24, requires_grad=True)
    ym = m(xm)
    torch.save({"sd": m.state_dict(), "x": xm.detach(), "y": ym.detach()}, f"{OUT}/swin_mlp.pt")
    # blocks, shifted and not
    for shift in (0, 2):
        torch.manual_seed(20 + shift)
        blk = randomize_(SwinTransformerBlock(36, (8, 12), 6, window_size=4, shift_size=shift), seed=3 + shift)
        xb = torch.randn(2, 96, 36, requires_grad=True)
        yb = blk(xb)
        w = torch.randn_like(yb)
        gb, gxb = grads_of(blk, yb, w, xb)
        torch.save({"kw": dict(dim=36, res=(8, 12), heads=6, ws=4, shift=shift), "sd": blk.state_dict(),
                    "x": xb.detach(), "y": yb.detach(), "w": w, "grads": gb, "gx": gxb}, f"{OUT}/swin_block_shift{shift}.pt")
    # whole generator (tiny): also exercises the reflect-pad / crop path (input 7x6 -> padded 8x8)
    torch.manual_seed(30)
    kw = dict(upscale=4, in_chans=1, img_size=8, window_size=4, embed_dim=24, depths=[2, 2], num_heads=[3, 3], mlp_ratio=2)
    net = randomize_(SwinIR(**kw), seed=7, table_std=0.5)
    xi = torch.rand(2, 1, 8, 8, requires_grad=True)
    yo = net(xi)
    w = torch.randn_like(yo)
    gn, gxi = grads_of(net, yo, w, xi)
    xpad = torch.rand(1, 1, 7, 6)
    torch.save({"kw": kw, "sd": net.state_dict(), "x": xi.detach(), "y": yo.detach(), "w": w, "grads": gn, "gx": gxi,
                "x_pad": xpad, "y_pad": net(xpad).detach()}, f"{OUT}/swinir_tiny.pt")


def hat_fixtures():
    """HAT (models/hat_arch/hat_arch.py, imported unmodified through tools/ref_shim.py) and the hybrid wrapper."""
    from tools import ref_shim
    m = ref_shim.hat_module()
    ws, heads, dim = 4, 3, 36
    # index / mask builders are methods of HAT: instantiate a tiny one
    torch.manual_seed(40)
    kw = dict(img_size=8, in_chans=1, embed_dim=dim, depths=(2, 2), num_heads=(heads, heads), window_size=ws,
              compress_ratio=3, squeeze_factor=6, upscale=4, upsampler="pixelshuffle", drop_path_rate=0.0)
    net = randomize_(m.HAT(**kw), seed=41, table_std=0.5)
    torch.save({"ws": ws, "rpi_sa": net.relative_position_index_SA, "rpi_oca": net.relative_position_index_OCA,
                "mask_8x12": net.calculate_mask((8, 12)), "mask_8x8": net.calculate_mask((8, 8))},
               f"{OUT}/hat_rpi_mask.pt")
    rpi_sa, rpi_oca = net.relative_position_index_SA, net.relative_position_index_OCA
    mask = net.calculate_mask((8, 12))
    # window attention with rpi + mask
    att = randomize_(m.WindowAttention(dim, (ws, ws), heads), seed=42)
    xa = torch.randn(12, ws * ws, dim, requires_grad=True)
    ya = att(xa, rpi_sa, mask)
    ga, gxa = grads_of(att, ya, torch.ones_like(ya) * torch.linspace(-1, 1, dim), xa)
    torch.save({"sd": att.state_dict(), "x": xa.detach(), "y": ya.detach(), "y_nomask": att(xa, rpi_sa, None).detach(),
                "grads": ga, "gx": gxa, "heads": heads}, f"{OUT}/hat_window_attention.pt")
    # CAB
    cab = randomize_(m.CAB(dim, 3, 6), seed=43)
    xc = torch.randn(2, dim, 8, 12, requires_grad=True)
    yc = cab(xc)
    wc = torch.randn_like(yc)
    gc, gxc = grads_of(cab, yc, wc, xc)
    torch.save({"sd": cab.state_dict(), "x": xc.detach(), "y": yc.detach(), "w": wc, "grads": gc, "gx": gxc},
               f"{OUT}/hat_cab.pt")
    # HAB shifted / unshifted, OCAB
    for shift in (0, 2):
        torch.manual_seed(44 + shift)
        blk = randomize_(m.HAB(dim, (8, 12), heads, window_size=ws, shift_size=shift, compress_ratio=3, squeeze_factor=6),
                         seed=45 + shift)
        xb = torch.randn(2, 96, dim, requires_grad=True)
        yb = blk(xb, (8, 12), rpi_sa, mask)
        w = torch.randn_like(yb)
        gb, gxb = grads_of(blk, yb, w, xb)
        torch.save({"kw": dict(dim=dim, res=(8, 12), heads=heads, ws=ws, shift=shift), "sd": blk.state_dict(),
                    "x": xb.detach(), "y": yb.detach(), "w": w, "grads": gb, "gx": gxb}, f"{OUT}/hat_hab_shift{shift}.pt")
    torch.manual_seed(48)
    oc = randomize_(m.OCAB(dim, (8, 12), ws, 0.5, heads, mlp_ratio=2), seed=49)
    xo = torch.randn(2, 96, dim, requires_grad=True)
    yo = oc(xo, (8, 12), rpi_oca)
    w = torch.randn_like(yo)
    go, gxo = grads_of(oc, yo, w, xo)
    torch.save({"kw": dict(dim=dim, res=(8, 12), heads=heads, ws=ws), "sd": oc.state_dict(), "x": xo.detach(),
                "y": yo.detach(), "w": w, "grads": go, "gx": gxo}, f"{OUT}/hat_ocab.pt")
    # whole HAT (tiny), eval-equivalent (drop_path_rate 0)
    xi = torch.rand(2, 1, 8, 8, requires_grad=True)
    yo = net(xi)
    w = torch.randn_like(yo)
    gn, gxi = grads_of(net, yo, w, xi)
    torch.save({"kw": {k: v for k, v in kw.items()}, "sd": net.state_dict(), "x": xi.detach(), "y": yo.detach(), "w": w,
                "grads": gn, "gx": gxi}, f"{OUT}/hat_tiny.pt")
    # hybrid wrapper (HAT x2 -> RRDB trunk -> nearest x2 -> convs)
    h = ref_shim.hybrid_module()
    torch.manual_seed(50)
    hkw = dict(img_size=8, in_chans=1, embed_dim=36, depths=(2,), num_heads=(3,), window_size=4, upscale=4, num_rrdb=2,
               num_feat=8, num_grow_ch=4)
    hyb = randomize_(h.HybridHATRealESRGAN(**hkw), seed=51, table_std=0.5).eval()  # eval: HAT default drop_path 0.1
    xh = torch.rand(1, 1, 8, 8, requires_grad=True)
    yh = hyb(xh)
    w = torch.randn_like(yh)
    gh, gxh = grads_of(hyb, yh, w, xh)
    torch.save({"kw": hkw, "sd": hyb.state_dict(), "x": xh.detach(), "y": yh.detach(), "w": w, "grads": gh, "gx": gxh},
               f"{OUT}/hybrid_tiny.pt")


def disc_fixtures():
    """Both U-Net discriminators (models/discriminator_swin.py, models/discriminator_hat.py) at num_feat = 8, in train mode
    (one power iteration inside the forward) after three warm-up forwards; the fixture holds the state BEFORE the recorded
    forward, the logits, the state's spectral-norm buffers AFTER it and every gradient."""
    from tools import ref_shim
    ref_shim.install()
    from models.discriminator_swin import UNetDiscriminatorSN as SwinD
    from models.discriminator_hat import UNetDiscriminatorSN as HatD
    for name, cls, kw in (("disc_swin_tiny", SwinD, dict(num_in_ch=1, num_feat=8)), ("disc_hat_tiny", HatD, dict(num_in_ch=1, num_feat=8))):
        torch.manual_seed(31)
        d = cls(**kw).train()
        with torch.no_grad():
            for _ in range(3):
                d(torch.rand(1, 1, 32, 32))
        sd0 = {k: v.clone() for k, v in d.state_dict().items()}
        x = torch.rand(2, 1, 64, 32, requires_grad=True)
        y = d(x)
        w = torch.randn_like(y)
        g, gx = grads_of(d, y, w, x)
        sd1 = {k: v.clone() for k, v in d.state_dict().items() if k.endswith("weight_u") or k.endswith("weight_v")}
        torch.save({"kw": kw, "sd": sd0, "x": x.detach(), "y": y.detach(), "w": w, "grads": g, "gx": gx, "sn_after": sd1},
                   f"{OUT}/{name}.pt")


if __name__ == "__main__":
    which = sys.argv[1:] or ["swin", "hat", "disc"]
    if "disc" in which:
        disc_fixtures()
    if "swin" in which:
        swin_fixtures()
    if "hat" in which:
        hat_fixtures()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
