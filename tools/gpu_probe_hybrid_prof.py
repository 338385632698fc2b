"""Per-kernel time table of one eager HybridHATRealESRGAN training step (torch.profiler, CUDA activities only).
Usage: gpu_probe_hybrid_prof.py [batch]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from superresolution_def_b200.hybridmodels_hat import HybridHATRealESRGAN
from superresolution_def_b200.synth import synthetic_pairs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
net = HybridHATRealESRGAN(img_size=128, in_chans=1, embed_dim=90, depths=(6,) * 4, num_heads=(6,) * 4, window_size=8,
                          upscale=4, num_rrdb=12, num_feat=48, num_grow_ch=24).cuda().train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True)
lr_, hr_ = synthetic_pairs(4, seed=1234)
lr_, hr_ = lr_.repeat(B // 4, 1, 1, 1).cuda(), hr_.repeat(B // 4, 1, 1, 1).cuda()


def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.l1_loss(net(lr_).float(), hr_)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as p:
    step()
    torch.cuda.synchronize()
rows = sorted(p.key_averages(), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print(f"total kernel time {tot / 1e3:.2f} ms")
for e in rows[:40]:
    print(f"{e.self_device_time_total / 1e3:8.3f} ms {100 * e.self_device_time_total / tot:5.1f}% n={e.count:4d} avg={e.self_device_time_total / e.count:8.1f}us  {e.key[:110]}")
