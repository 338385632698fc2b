"""Diagnostic: per-parameter gradient error of the CUDA path vs the fp32 oracle, next to the error of the
oracle itself under bf16 autocast (the reference's own training dtype)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import rel_l2, randomize_
from oracle import swinir_oracle as o
from superresolution_def_b200.architecture_swin import SwinIR

torch.manual_seed(4)
depths = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 2]
kw = dict(img_size=16, window_size=8, depths=depths, num_heads=[6] * len(depths))
net = randomize_(SwinIR(upscale=4, in_chans=1, embed_dim=180, mlp_ratio=2, **kw), seed=5, table_std=0.5).cuda()
x = torch.rand(2, 1, 16, 16, device="cuda")
w = torch.randn(2, 1, 64, 64, device="cuda")

def run_oracle(autocast):
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in net.state_dict().items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = o.swinir_forward(x, sd, upscale=4, **kw)
    (out.float() * w).mean().backward()
    return out, sd

ref, sd32 = run_oracle(False)
r16, sd16 = run_oracle(True)
got = net(x)
(got.float() * w).mean().backward()
print("out  mine %.4f  autocast %.4f" % (rel_l2(got, ref), rel_l2(r16, ref)))
for n, p in net.named_parameters():
    print("%-50s mine %.4f  autocast %.4f" % (n, rel_l2(p.grad, sd32[n].grad), rel_l2(sd16[n].grad, sd32[n].grad)))
