"""Does the CUDA-graph replay of the training step follow the eager trajectory?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200.architecture_swin import SwinIR
from superresolution_def_b200.synth import synthetic_pairs
from superresolution_def_b200.graphs import GraphedStep
from superresolution_def_b200.dp import BucketedGradReducer, swinir_grad_groups

B = 4
kw = dict(upscale=4, in_chans=1, img_size=128, window_size=8, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6)
lr_, hr_ = synthetic_pairs(B, seed=1234)
lr_, hr_ = lr_.cuda(), hr_.cuda()
res = {}
for mode in ("eager", "graph"):
    torch.manual_seed(0)
    net = SwinIR(mlp_ratio=2, **kw).cuda()
    red = BucketedGradReducer(swinir_grad_groups(net), 1)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True, capturable=True)
    def step(l, h):
        red.zero_grad()
        loss = torch.nn.functional.l1_loss(net(l).float(), h)
        loss.backward()
        red.finish()
        opt.step()
        return loss
    losses = []
    if mode == "eager":
        for s in range(12):
            losses.append(step(lr_, hr_).item())
    else:
        g = GraphedStep(step, (lr_.clone(), hr_.clone()), warmup=2)   # 3 steps consumed (2 warm-up + capture)
        losses = [float("nan")] * 3
        for s in range(9):
            losses.append(g(lr_, hr_).item())
    res[mode] = losses
for s in range(12):
    print(f"step {s:2d} eager {res['eager'][s]:.5f} graph {res['graph'][s]:.5f}")
