"""Two eager SwinIR training steps at the bench configuration (batch 16, 128^2 -> 512^2), nothing else: the command the ncu
launch list of a step is taken from (profiles/rNN_launches_step.csv = the second half of the launches).
Usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/gpu_step_eager.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200.architecture_swin import SwinIR
from superresolution_def_b200.synth import synthetic_pairs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
os.environ.setdefault("SRK_FP32_POLICY", "allow")
torch.manual_seed(0)
net = SwinIR(upscale=4, in_chans=1, img_size=128, window_size=8, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6, mlp_ratio=2).cuda()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True)
lr_, hr_ = synthetic_pairs(4, seed=1234)
lr_, hr_ = lr_.repeat(B // 4, 1, 1, 1).cuda(), hr_.repeat(B // 4, 1, 1, 1).cuda()
torch.cuda.synchronize()
print("STEP-MARK begin", flush=True)
for _ in range(2):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.l1_loss(net(lr_).float(), hr_)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("done", float(loss))
