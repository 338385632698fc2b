"""N discriminator steps (forward + backward with weight gradients, fp16 autocast as train_swin.py:221-233) of the product
mirror at the script's shape (micro-batch 2, 512^2), for `ncu`: python tools/gpu_disc_step.py [steps]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from superresolution_def_b200.gan import UNetDiscriminatorSN

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
net = UNetDiscriminatorSN(1, 64).cuda().train()
x = torch.rand(2, 1, 512, 512, device="cuda")
for _ in range(steps):
    with torch.autocast("cuda"):
        out = net(x)
    out.float().mean().backward()
    net.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("ok", float(out.float().mean()))
