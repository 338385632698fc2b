"""Host-side profile (cProfile) of the discriminator step of train_swin.py's shape: where the Python time of 50 D steps goes.
python tools/gpu_disc_hostprof.py"""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_def_b200.gan import UNetDiscriminatorSN   # noqa: E402

torch.manual_seed(0)
net = UNetDiscriminatorSN(1, 64).cuda().train()
x = torch.rand(2, 1, 512, 512, device="cuda")


def step():
    with torch.autocast("cuda"):
        out = net(x)
    out.float().mean().backward()
    net.zero_grad(set_to_none=True)


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
