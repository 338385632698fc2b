"""-m gpu: the training-script conventions of the reference run on top of the mirrored modules:
fp16 autocast + GradScaler (train_swin.py:169,217-259), DDP(find_unused_parameters=True) (:152), requires_grad toggling
(:214-215,237-238), no_grad / inference_mode forwards (:217-219,281-286), EMA over named_parameters (:53-74)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_swinir_generator_step_like_train_swin():
    from superresolution_def_b200.architecture_swin import SwinIR
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        torch.manual_seed(0)
        net = SwinIR(upscale=4, in_chans=1, img_size=16, window_size=8, embed_dim=180, depths=[2], num_heads=[6],
                     mlp_ratio=2).cuda()
        ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[0], find_unused_parameters=True)
        opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4, betas=(0.9, 0.99))
        scaler = torch.amp.GradScaler("cuda")
        shadow = {n: p.detach().clone() for n, p in ddp.module.named_parameters()}
        lr, hr = torch.rand(2, 1, 16, 16, device="cuda"), torch.rand(2, 1, 64, 64, device="cuda")
        # D-step style: generator frozen, forward under no_grad + autocast
        for p in ddp.parameters():
            p.requires_grad = False
        with torch.no_grad(), torch.autocast("cuda"):
            sr0 = ddp(lr)
        assert sr0.dtype == torch.float16 and sr0.shape == (2, 1, 64, 64)
        for p in ddp.parameters():
            p.requires_grad = True
        losses = []
        for _ in range(3):
            with torch.autocast("cuda"):
                sr = ddp(lr)
                loss = torch.nn.functional.l1_loss(sr, hr)
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            opt.zero_grad(set_to_none=True)
            losses.append(loss.item())
            for n, p in ddp.module.named_parameters():   # EMA update as train_swin.py:60-64
                shadow[n].mul_(0.999).add_(p.detach(), alpha=0.001)
        assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0]
        with torch.inference_mode(), torch.autocast("cuda"):
            out = ddp.module(lr)
        assert torch.isfinite(out.float()).all()
    finally:
        dist.destroy_process_group()


def test_gan_micro_steps_like_train_swin():
    """train_swin.py:214-259 on the mirrors: D step with the generator frozen and run under no_grad, G step with the
    discriminator frozen, fp16 autocast + one GradScaler, accumulation over 2 micro-steps, EMA; DDP-wrapped like the
    script.  Losses stay finite, both networks' parameters move, frozen-phase gradients do not leak."""
    from superresolution_def_b200.architecture_swin import SwinIR
    from superresolution_def_b200.gan import UNetDiscriminatorSN, GanTrainer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        torch.manual_seed(0)
        g = SwinIR(upscale=4, in_chans=1, img_size=32, window_size=8, embed_dim=180, depths=[2], num_heads=[6], mlp_ratio=2).cuda()
        d = UNetDiscriminatorSN(num_in_ch=1, num_feat=16).cuda()
        DDP = torch.nn.parallel.DistributedDataParallel
        g = DDP(g, device_ids=[0], find_unused_parameters=True)
        d = DDP(d, device_ids=[0], find_unused_parameters=False)
        tr = GanTrainer(g, d, accum=2)
        g0 = [p.detach().clone() for p in g.parameters()]
        d0 = [p.detach().clone() for p in d.parameters()]
        lr, hr = torch.rand(2, 1, 32, 32, device="cuda"), torch.rand(2, 1, 128, 128, device="cuda")
        out = [tr.micro_step(lr, hr) for _ in range(4)]
        vals = torch.tensor([[float(a), float(b)] for a, b in out])
        assert torch.isfinite(vals).all(), vals
        assert any(not torch.equal(a, b.detach()) for a, b in zip(g0, g.parameters()))
        assert any(not torch.equal(a, b.detach()) for a, b in zip(d0, d.parameters()))
        assert all(torch.isfinite(v).all() for v in tr.ema.shadow.values())
    finally:
        dist.destroy_process_group()
