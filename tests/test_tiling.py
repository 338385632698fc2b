"""Tiled full-frame inference (BASELINE configs[4]): host logic on CPU (single process and world-size-2 gloo), and on GPU
the assembled frame equals model(lr_tile) tile by tile."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from superresolution_def_b200.tiling import sr_frame_tiled, tile_origins, extract_tiles


def _fake_model(x):  # local operator: x4 nearest upsampling commutes with tiling exactly
    return F.interpolate(x, scale_factor=4, mode="nearest")


def test_tiling_matches_direct_for_a_local_operator():
    torch.manual_seed(0)
    frame = torch.rand(64, 96)
    out = sr_frame_tiled(_fake_model, frame, tile=32, batch=4)
    assert out.shape == (1, 1, 256, 384)
    assert torch.equal(out, _fake_model(frame[None, None]))
    out_h = sr_frame_tiled(_fake_model, frame, tile=32, halo=8, batch=5)
    assert torch.equal(out_h, out)
    # cross-faded overlaps: a local operator gives every overlapping tile the same values, so the weighted mean is exact
    out_b = sr_frame_tiled(_fake_model, frame, tile=32, halo=8, batch=5, blend=True)
    assert out_b.shape == out.shape and torch.allclose(out_b, out, atol=1e-6)
    # and it IS a blend: tiles that disagree by a constant are cross-faded monotonically across the seam
    seam = sr_frame_tiled(lambda x: F.interpolate(x * 0 + x[:, :, :1, :1], scale_factor=4, mode="nearest"), frame, tile=32,
                          halo=8, batch=2, blend=True)
    row = seam[0, 0, 8, 96:160]                       # crosses the seam between tile columns 0 and 1 (at x = 128)
    lo, hi = float(min(row[0], row[-1])), float(max(row[0], row[-1]))
    assert bool(((row >= lo - 1e-6) & (row <= hi + 1e-6)).all()) and float((row[1:] - row[:-1]).abs().max()) <= (hi - lo) / 32 + 1e-6
    with pytest.raises(ValueError):
        sr_frame_tiled(_fake_model, frame, tile=32, blend=True)
    with pytest.raises(ValueError):
        tile_origins(60, 96, 32)
    assert extract_tiles(frame[None, None], [(0, 0), (32, 64)], 32, 4).shape == (2, 1, 40, 40)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    frame = torch.rand(96, 96)
    out = sr_frame_tiled(_fake_model, frame, tile=32, batch=2, rank=rank, world=world)
    # by value (numpy pickles its bytes): a torch tensor would travel as a shared-memory handle that dies with this worker
    q.put((rank, None if out is None else out.numpy().copy()))
    dist.destroy_process_group()


def test_tiles_sharded_over_two_ranks_gather_on_rank0():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    torch.manual_seed(0)
    frame = torch.rand(96, 96)
    assert res[1] is None
    assert torch.equal(torch.from_numpy(res[0]), _fake_model(frame[None, None]))   # 9 tiles: ranks hold 5 and 4


@pytest.mark.gpu
def test_tiled_hat_equals_per_tile_model_output():
    from superresolution_def_b200.hat_arch import HAT
    from tests.util import randomize_
    torch.manual_seed(1)
    net = randomize_(HAT(img_size=32, in_chans=1, embed_dim=180, depths=(1,), num_heads=(6,), window_size=16, upscale=4,
                         upsampler="pixelshuffle"), seed=2, table_std=0.5).cuda().eval()
    frame = torch.rand(64, 96)
    out = sr_frame_tiled(net, frame, tile=32, batch=4, device="cuda")
    assert out.shape == (1, 1, 256, 384)
    with torch.no_grad():
        for (y, x) in tile_origins(64, 96, 32):
            ref = net(frame[None, None, y:y + 32, x:x + 32].cuda()).float().cpu()
            got = out[:, :, 4 * y:4 * y + 128, 4 * x:4 * x + 128]
            assert (got - ref).abs().max() < 2e-2 * ref.abs().max(), (y, x)   # batch-composition independent up to bf16 noise
