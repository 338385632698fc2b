"""HAT x4 (BASELINE configs[2]: window 16, OCAB, CAB, C=180, 6x6) training step timing on one B200.
Usage: gpu_probe_hat.py [batch] [steps] [--prof]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200.hat_arch import HAT
from superresolution_def_b200.synth import synthetic_pairs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
prof = "--prof" in sys.argv
torch.manual_seed(0)
net = HAT(img_size=128, in_chans=1, embed_dim=180, depths=(6,) * 6, num_heads=(6,) * 6, window_size=16, upscale=4,
          upsampler="pixelshuffle", drop_path_rate=0.0).cuda()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True)
lr_, hr_ = synthetic_pairs(min(B, 4), seed=1234)
reps = (B + lr_.shape[0] - 1) // lr_.shape[0]
lr_, hr_ = lr_.repeat(reps, 1, 1, 1)[:B].cuda(), hr_.repeat(reps, 1, 1, 1)[:B].cuda()


def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.l1_loss(net(lr_).float(), hr_)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    l = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    l = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"HAT cfg3 B={B}: {ms:.2f} ms/step  {B / ms * 1e3:.1f} patches/s  loss {l.item():.4f}  "
      f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GB  step TFLOP/s {B * 3026.031 / ms:.1f}", flush=True)
if prof:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as p:
        step()
        torch.cuda.synchronize()
    print(p.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
