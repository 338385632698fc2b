"""Probe: whole-HAT (C=90, ws 8, x2) parity vs oracle for several seeds; prints per-parameter errors above threshold."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import rel_l2, randomize_
from oracle import hat_oracle as ho
from superresolution_def_b200.hat_arch import HAT

kw = dict(window_size=8, depths=(2,), num_heads=(6,))
for seed in (8, 18, 28):
    torch.manual_seed(seed)
    net = randomize_(HAT(img_size=32, in_chans=1, embed_dim=90, upscale=2, upsampler="pixelshuffle", drop_path_rate=0.0,
                         **kw), seed=seed, table_std=0.5).cuda()
    x = torch.rand(2, 1, 32, 32, device="cuda")
    w = torch.randn(2, 1, 64, 64, device="cuda")

    def run_oracle(autocast):
        sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in net.state_dict().items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = ho.hat_forward(x, sd, upscale=2, **kw)
        (out.float() * w).mean().backward()
        return out, sd

    ref, sd32 = run_oracle(False)
    r16, sd16 = run_oracle(True)
    got = net(x)
    (got.float() * w).mean().backward()
    print("seed", seed, "out", rel_l2(got, ref), "autocast", rel_l2(r16, ref))
    for n, p in net.named_parameters():
        mine, auto = rel_l2(p.grad, sd32[n].grad), rel_l2(sd16[n].grad, sd32[n].grad)
        if mine > 1.6 * auto + 1e-2 or mine > 0.05:
            print("   ", n, round(mine, 4), round(auto, 4))
