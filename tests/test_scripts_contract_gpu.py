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
        d = UNetDiscriminatorSN(num_in_ch=1, num_feat=64).cuda()
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


def _detect_hybrid_params(sd):
    """The shape sniffing infer_hat.py:52-112 applies to a checkpoint (re-stated: embed_dim from hat.conv_first, num_feat from
    conv_adapt, growth from rdb1.conv1, RRDB count and HAT stage count from the key indices)."""
    p = dict(img_size=128, in_chans=1, embed_dim=90, depths=(6, 6, 6, 6), num_heads=(6, 6, 6, 6), window_size=8, upscale=4,
             num_rrdb=12, num_feat=48, num_grow_ch=24)
    p["embed_dim"] = sd["hat.conv_first.weight"].shape[0]
    p["num_feat"] = sd["conv_adapt.weight"].shape[0]
    p["num_grow_ch"] = sd["rrdb_trunk.0.rdb1.conv1.weight"].shape[0]
    p["num_rrdb"] = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("rrdb_trunk."))
    stages = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("hat.layers."))
    p["depths"], p["num_heads"] = (6,) * stages, (6,) * stages
    return p


def test_hybrid_train_and_infer_like_train_hat_and_infer_hat(tmp_path, monkeypatch):
    """train_hat.py:128-149,222-266: fp32 (NO autocast), DDP(find_unused_parameters=False), gradient accumulation, EMA copy
    updated by zipping .parameters(), D frozen during the G step and fed sr.detach() afterwards; then infer_hat.py:160-177:
    checkpoint with 'module.' prefixes -> shape auto-detection -> strict load -> eval forward -> 16-bit TIFF."""
    import warnings
    from superresolution_def_b200.hybridmodels_hat import HybridHATRealESRGAN
    from superresolution_def_b200.discriminator_hat import UNetDiscriminatorSN      # train_hat.py:26
    from superresolution_def_b200.gan import CombinedGANLoss, DiscriminatorLoss
    from superresolution_def_b200.input_pipeline import save_as_tiff16, read_tiff_u16
    from superresolution_def_b200 import swin_engine as eng
    monkeypatch.setattr(eng, "_fp32_warned", False)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    kw = dict(img_size=32, in_chans=1, embed_dim=90, depths=(6,), num_heads=(6,), window_size=8, upscale=4, num_rrdb=2,
              num_feat=48, num_grow_ch=24)
    try:
        torch.manual_seed(0)
        net_g = HybridHATRealESRGAN(**kw).cuda()
        net_ema = HybridHATRealESRGAN(**kw).cuda()
        for p in net_ema.parameters():
            p.requires_grad = False
        net_d = UNetDiscriminatorSN(num_in_ch=1, num_feat=64).cuda()
        DDP = torch.nn.parallel.DistributedDataParallel
        net_g = DDP(net_g, device_ids=[0], find_unused_parameters=False)
        net_d = DDP(net_d, device_ids=[0])
        opt_g = torch.optim.Adam(net_g.parameters(), lr=1e-4, betas=(0.9, 0.99))
        opt_d = torch.optim.Adam(net_d.parameters(), lr=1e-4, betas=(0.9, 0.99))
        crit_g, crit_d = CombinedGANLoss().cuda(), DiscriminatorLoss().cuda()
        lr, hr = torch.rand(1, 1, 32, 32, device="cuda"), torch.rand(1, 1, 128, 128, device="cuda")
        accum, losses = 2, []
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            for i in range(4):
                warm = i < 2                                   # warm-up epochs train on L1 only (train_hat.py:236-238)
                for p in net_d.parameters():
                    p.requires_grad = False
                sr = net_g(lr)                                  # fp32 in, no autocast
                assert sr.dtype == torch.float32
                l1 = torch.nn.functional.l1_loss(sr, hr)
                loss_g = l1 if warm else crit_g(sr, hr, net_d(hr).detach(), net_d(sr))[0]
                (loss_g / accum).backward()
                if (i + 1) % accum == 0:
                    opt_g.step(); opt_g.zero_grad()
                    with torch.no_grad():
                        for ps, pe in zip(net_g.module.parameters(), net_ema.parameters()):
                            pe.data.mul_(0.999).add_(ps.data, alpha=0.001)
                if not warm:
                    for p in net_d.parameters():
                        p.requires_grad = True
                    loss_d, _ = crit_d(net_d(hr), net_d(sr.detach()))
                    (loss_d / accum).backward()
                    if (i + 1) % accum == 0:
                        opt_d.step(); opt_d.zero_grad()
                losses.append(float(loss_g))
        assert all(torch.isfinite(torch.tensor(losses))), losses
        # the fp32 -> bf16 downgrade is announced exactly once (SRK_FP32_POLICY=warn is the default)
        assert len([m for m in w if issubclass(m.category, RuntimeWarning) and "fp32 input outside autocast" in str(m.message)]) == 1
        ckpt = tmp_path / "best_hybrid_model.pth"
        torch.save({"model_state_dict": net_g.state_dict()}, ckpt)             # DDP state_dict: 'module.' prefixes
        sd = torch.load(ckpt, map_location="cpu")["model_state_dict"]
        clean = {k.replace("module.", ""): v for k, v in sd.items()}
        params = _detect_hybrid_params(clean)
        assert (params["embed_dim"], params["num_feat"], params["num_grow_ch"], params["num_rrdb"], params["depths"]) == (90, 48, 24, 2, (6,))
        params["img_size"] = 32
        model = HybridHATRealESRGAN(**params).cuda()
        model.load_state_dict(clean, strict=True)
        model.eval()
        with torch.no_grad():
            out = torch.clamp(model(lr), 0, 1)
            ref = torch.clamp(net_g.module.eval()(lr), 0, 1)
        assert torch.equal(out, ref)                      # same weights, same kernels: bit-identical
        save_as_tiff16(out, tmp_path / "sr.tiff")
        back = read_tiff_u16(tmp_path / "sr.tiff")
        assert back.shape == (128, 128) and back.dtype.name == "uint16"
        assert abs(float(back.astype("float32").mean()) / 65535.0 - float(out.mean())) < 1e-4
    finally:
        dist.destroy_process_group()
