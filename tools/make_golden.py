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


if __name__ == "__main__":
    swin_fixtures()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
