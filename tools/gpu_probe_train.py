"""Loss trajectory of the CUDA path vs the oracle under bf16 autocast (identical init, data, optimizer)."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import swinir_oracle as o
from superresolution_def_b200.architecture_swin import SwinIR
from superresolution_def_b200.synth import synthetic_pairs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
kw = dict(upscale=4, in_chans=1, img_size=128, window_size=8, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6)
torch.manual_seed(0)
net = SwinIR(mlp_ratio=2, **kw).cuda()
sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in net.state_dict().items()}
lr_, hr_ = synthetic_pairs(B, seed=1234)
lr_, hr_ = lr_.cuda(), hr_.cuda()
opt_m = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.9, 0.99))
opt_o = torch.optim.AdamW([v for v in sd.values() if v.is_floating_point()], lr=1e-4, betas=(0.9, 0.99))
okw = {k: kw[k] for k in ("img_size", "window_size", "depths", "num_heads", "upscale")}
for s in range(steps):
    opt_m.zero_grad(set_to_none=True)
    lm = torch.nn.functional.l1_loss(net(lr_).float(), hr_)
    lm.backward(); opt_m.step()
    opt_o.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = o.swinir_forward(lr_, sd, **okw)
    lo = torch.nn.functional.l1_loss(out.float(), hr_)
    lo.backward(); opt_o.step()
    print(f"step {s:2d}  loss ours {lm.item():.5f}   oracle(autocast) {lo.item():.5f}", flush=True)
