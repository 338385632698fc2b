"""Tiled x4 inference of a full observatory frame, tiles sharded across the GPUs of one box (BASELINE configs[4]).

The reference promises tiled inference for large frames (README.md:158) but `infer_hat.py` only runs the model on each
pre-cut 128x128 test patch (infer_hat.py:219-226), so the contract here is defined tile by tile: the
assembled frame equals `model(lr_tile)` for every tile (exactly, with halo == 0; with a halo the centre crop of the
larger tile is kept, which removes the seams a zero-padded convolution leaves at tile borders).

Tiles are independent: rank r processes tiles r, r + world, ... in batches; there is no data-path collective, only the
final gather of the finished tiles on rank 0.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist
import torch.nn.functional as F


def tile_origins(H: int, W: int, tile: int) -> list[tuple[int, int]]:
    """Top-left corners of the non-overlapping tile grid (the frame must be a multiple of the tile size)."""
    if H % tile or W % tile:
        raise ValueError(f"frame {H}x{W} is not a multiple of the tile size {tile}")
    return [(y, x) for y in range(0, H, tile) for x in range(0, W, tile)]


def extract_tiles(frame: torch.Tensor, origins, tile: int, halo: int) -> torch.Tensor:
    """frame (1,1,H,W) -> (n,1,tile+2*halo,tile+2*halo); the frame border is reflect-padded by `halo`."""
    if halo:
        frame = F.pad(frame, (halo, halo, halo, halo), mode="reflect")
    size = tile + 2 * halo
    return torch.cat([frame[:, :, y:y + size, x:x + size] for y, x in origins], dim=0)


@torch.no_grad()
def sr_frame_tiled(model: Callable[[torch.Tensor], torch.Tensor], frame: torch.Tensor, *, tile: int = 128, halo: int = 0,
                   scale: int = 4, batch: int = 16, rank: int = 0, world: int = 1, device=None):
    """Super-resolve `frame` ((H,W) or (1,1,H,W), float in [0,1]) tile by tile.

    Returns the (1,1,scale*H,scale*W) result on rank 0 (CPU tensor) and None on the other ranks.  `model` maps
    (b,1,t,t) -> (b,1,scale*t,scale*t) with t = tile + 2*halo."""
    if frame.dim() == 2:
        frame = frame[None, None]
    H, W = frame.shape[-2:]
    origins = tile_origins(H, W, tile)
    mine = list(range(rank, len(origins), world))
    device = device or frame.device
    out_tile = tile * scale
    results = torch.empty(len(mine), 1, out_tile, out_tile, dtype=torch.float32)
    for s in range(0, len(mine), batch):
        idx = mine[s:s + batch]
        lr = extract_tiles(frame, [origins[i] for i in idx], tile, halo).to(device, non_blocking=True)
        sr = model(lr).float()
        if halo:
            sr = sr[:, :, halo * scale:halo * scale + out_tile, halo * scale:halo * scale + out_tile]
        results[s:s + len(idx)] = sr.cpu()
    if world > 1:
        # final gather only: every rank contributes the same number of tiles up to one; pad to the maximum
        per = (len(origins) + world - 1) // world
        buf = torch.zeros(per, 1, out_tile, out_tile)
        buf[:len(mine)] = results
        backend = dist.get_backend()
        send = buf.to(device) if backend == "nccl" else buf
        gathered = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
        dist.gather(send, gathered, dst=0)
        if rank != 0:
            return None
        parts = [g.cpu() for g in gathered]
    else:
        parts = [results]
    out = torch.empty(1, 1, H * scale, W * scale, dtype=torch.float32)
    for r, part in enumerate(parts):
        for k, i in enumerate(range(r, len(origins), world)):
            y, x = origins[i]
            out[0, 0, y * scale:(y + tile) * scale, x * scale:(x + tile) * scale] = part[k, 0]
    return out
