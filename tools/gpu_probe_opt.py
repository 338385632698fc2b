import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200.architecture_swin import SwinIR
from superresolution_def_b200.synth import synthetic_pairs
from superresolution_def_b200.dp import BucketedGradReducer, swinir_grad_groups
B = 2
kw = dict(upscale=4, in_chans=1, img_size=128, window_size=8, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6)
lr_, hr_ = synthetic_pairs(B, seed=1234); lr_, hr_ = lr_.cuda(), hr_.cuda()
for mode in ("plain", "reducer+plain", "none+fused", "none+fused_capt", "reducer+fused_capt"):
    torch.manual_seed(0)
    net = SwinIR(mlp_ratio=2, **kw).cuda()
    red = BucketedGradReducer(swinir_grad_groups(net), 1) if "reducer" in mode else None
    okw = dict(lr=1e-4, betas=(0.9, 0.99))
    if "fused" in mode: okw["fused"] = True
    if "capt" in mode: okw["capturable"] = True
    opt = torch.optim.AdamW(net.parameters(), **okw)
    out = []
    for s in range(5):
        if red: red.zero_grad()
        else: opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.l1_loss(net(lr_).float(), hr_)
        loss.backward(); opt.step(); out.append(round(loss.item(), 5))
    gn = sum(p.grad.float().norm().item() ** 2 for p in net.parameters()) ** 0.5
    print(f"{mode:22s} {out}  |g|={gn:.4f}", flush=True)
