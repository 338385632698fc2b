"""Host-side profile (cProfile) of an eager SwinIR forward + backward at the GAN micro-batch (2 x 128^2, train_swin.py:147-149):
where the Python time goes when the step is launched eagerly (BASELINE configs[3]).  python tools/gpu_swinir_hostprof.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_def_b200.architecture_swin import SwinIR   # noqa: E402

torch.manual_seed(0)
net = SwinIR(upscale=4, in_chans=1, img_size=128, window_size=8, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6, mlp_ratio=2).cuda().train()
x = torch.rand(2, 1, 128, 128, device="cuda")


def step():
    with torch.autocast("cuda"):
        out = net(x)
    out.float().mean().backward()
    net.zero_grad(set_to_none=True)


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
t_host = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 20 * 1e3
print(f"eager fwd+bwd: host {t_host:.2f} ms, with device {t_all:.2f} ms per step")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
pstats.Stats(pr).sort_stats("cumtime").print_stats(25)
