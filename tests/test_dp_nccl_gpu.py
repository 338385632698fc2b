"""Multi-rank gradient equality on hardware (SURVEY §4 level 4; train_swin.py:152, train_hat.py:148): two ranks over
NCCL, each with its shard of the batch, must end up with the gradients a single process computes for the concatenated
batch — through BucketedGradReducer (overlapped hooks, and the serial reduce_all mode), through the reducer under
gradient accumulation, and through torch's own DistributedDataParallel(find_unused_parameters=True) wrapped around the
mirror, which is how the reference scripts would run it.  Needs two GPUs: run with `gpurun --gpus 2`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

KW = dict(upscale=4, in_chans=1, img_size=16, window_size=8, embed_dim=180, depths=[2, 2], num_heads=[6, 6], mlp_ratio=2)
PER_RANK = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(dev):
    from superresolution_def_b200.architecture_swin import SwinIR
    from tests.util import randomize_
    torch.manual_seed(0)
    return randomize_(SwinIR(**KW), seed=5).to(dev)


def _batch(world):
    g = torch.Generator().manual_seed(11)
    lr = torch.rand(world * PER_RANK * 2, 1, 16, 16, generator=g)
    hr = torch.rand(world * PER_RANK * 2, 1, 64, 64, generator=g)
    return lr, hr


def _worker(rank, world, port, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from superresolution_def_b200.dp import BucketedGradReducer, swinir_grad_groups
    net = _build(dev)
    lr, hr = _batch(world)
    # two micro-batches per rank (gradient accumulation modes use both; the others see them as one batch)
    sl = [slice((2 * rank + m) * PER_RANK, (2 * rank + m + 1) * PER_RANK) for m in range(2)]
    l1 = torch.nn.functional.l1_loss
    if mode == "ddp":
        ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[rank], find_unused_parameters=True)
        idx = slice(2 * rank * PER_RANK, (2 * rank + 2) * PER_RANK)
        l1(ddp(lr[idx].to(dev)).float(), hr[idx].to(dev)).backward()
    else:
        red = BucketedGradReducer(swinir_grad_groups(net), world, overlap=(mode != "serial"))
        red.zero_grad()
        if mode in ("overlap", "serial"):
            idx = slice(2 * rank * PER_RANK, (2 * rank + 2) * PER_RANK)
            l1(net(lr[idx].to(dev)).float(), hr[idx].to(dev)).backward()
            red.finish() if mode == "overlap" else red.reduce_all()
        elif mode == "accum":       # DDP semantics: every micro-step's backward is reduced
            for m in range(2):
                (l1(net(lr[sl[m]].to(dev)).float(), hr[sl[m]].to(dev)) / 2).backward()
                red.finish()
        elif mode == "accum_no_sync":
            with red.no_sync():
                (l1(net(lr[sl[0]].to(dev)).float(), hr[sl[0]].to(dev)) / 2).backward()
            (l1(net(lr[sl[1]].to(dev)).float(), hr[sl[1]].to(dev)) / 2).backward()
            red.finish()
    torch.cuda.synchronize()
    torch.save({n: p.grad.detach().float().cpu() for n, p in net.named_parameters()}, os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["overlap", "serial", "accum", "accum_no_sync", "ddp"])
def test_world2_nccl_gradients_equal_single_process(mode, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)
    dev = torch.device("cuda", 0)
    net = _build(dev)
    lr, hr = _batch(world)
    torch.nn.functional.l1_loss(net(lr.to(dev)).float(), hr.to(dev)).backward()
    ref = {n: p.grad.detach().float().cpu() for n, p in net.named_parameters()}
    worst = (0.0, "")
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"g{r}.pt"))
        assert got.keys() == ref.keys()
        for n, g in got.items():
            den = ref[n].norm().item() + 1e-12
            rel = (g - ref[n]).norm().item() / den
            mx = (g - ref[n]).abs().max().item()
            worst = max(worst, (rel, f"rank {r} {n} max_abs {mx:.3e}"))
            # the shards' gradients are accumulated in fp32 and the 1/2 loss scaling is exact in bf16, so the only
            # difference from the single-process batch is fp32 summation order and the bf16 rounding of averaged
            # activations' gradients at shard boundaries: 2e-3 relative, max-abs within 1e-2 of the tensor's own scale
            assert rel < 2e-3 and mx <= 1e-2 * ref[n].abs().max().item() + 1e-7, (mode, r, n, rel, mx)
    print(f"[{mode}] worst rel-L2 {worst[0]:.3e} ({worst[1]})")
