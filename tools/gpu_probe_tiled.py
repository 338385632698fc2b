"""BASELINE configs[4]: HAT x4 tiled inference of a synthetic 4096x4096 frame (32x32 tiles of 128^2), one GPU or
torchrun (tiles round-robin over ranks, final gather on rank 0).  Also prints forward-only patches/s of SwinIR and HAT.
Usage: [torchrun ...] gpu_probe_tiled.py [frame_side] [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from superresolution_def_b200.hat_arch import HAT
from superresolution_def_b200.architecture_swin import SwinIR
from superresolution_def_b200 import swin_engine
from superresolution_def_b200.tiling import sr_frame_tiled

side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
torch.manual_seed(0)
hat = HAT(img_size=128, in_chans=1, embed_dim=180, depths=(6,) * 6, num_heads=(6,) * 6, window_size=16, upscale=4,
          upsampler="pixelshuffle").to(dev).eval()
swin_engine.freeze_weights(True)   # inference: keep the prepared bf16 operands between forwards
g = torch.Generator().manual_seed(1)
frame = torch.rand(side, side, generator=g)
with torch.no_grad():
    for _ in range(2):
        hat(torch.rand(batch, 1, 128, 128, device=dev))
    # warm-up frame: creates the NCCL point-to-point channels the gather uses and the pinned result buffer pool
    sr_frame_tiled(hat, torch.rand(1024, 1024, generator=g), tile=128, batch=batch, rank=rank, world=world, device=dev)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
out = sr_frame_tiled(hat, frame, tile=128, batch=batch, rank=rank, world=world, device=dev)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    n = (side // 128) ** 2
    print(f"tiled HAT x4 inference: {side}x{side} frame, {n} tiles on {world} GPU(s): {dt:.2f} s/frame = {1 / dt:.3f} frames/s, "
          f"{n / dt:.1f} tiles/s  (output {tuple(out.shape)}, finite={bool(torch.isfinite(out).all())})", flush=True)
if world == 1:
    x = torch.rand(batch, 1, 128, 128, device=dev)
    for name, net in (("HAT", hat), ("SwinIR", SwinIR(upscale=4, in_chans=1, img_size=128, window_size=8, embed_dim=180,
                                                      depths=[6] * 6, num_heads=[6] * 6).to(dev).eval())):
        with torch.no_grad():
            for _ in range(3):
                net(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                net(x)
            e1.record()
            e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name} forward only (eval, no_grad), batch {batch}: {ms:.2f} ms = {batch / ms * 1e3:.1f} patches/s", flush=True)
if world > 1:
    dist.destroy_process_group()
