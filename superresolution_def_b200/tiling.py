"""Tiled x4 inference of a full observatory frame, tiles sharded across the GPUs of one box (BASELINE configs[4]).

The reference promises tiled inference for large frames (README.md:158) but `infer_hat.py` only runs the model on each
pre-cut 128x128 test patch (infer_hat.py:219-226), so the contract here is defined tile by tile: the
assembled frame equals `model(lr_tile)` for every tile (exactly, with halo == 0; with a halo the centre crop of the
larger tile is kept, which removes the seams a zero-padded convolution leaves at tile borders).

Tiles are independent: rank r processes tiles r, r + world, ... in batches; there is no data-path collective, only the
final gather of the finished tiles on rank 0 (device to device), after which the frame is one permute of the tile stack.
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


def _tile_view(frame: torch.Tensor, tile: int, halo: int) -> torch.Tensor:
    """(1,1,H,W) -> (ny, nx, t, t) strided view of all tiles, t = tile + 2*halo (the frame border is reflect-padded by
    `halo`); no copy until a batch of tiles is gathered."""
    if halo:
        frame = F.pad(frame, (halo, halo, halo, halo), mode="reflect")
    size = tile + 2 * halo
    return frame[0, 0].unfold(0, size, tile).unfold(1, size, tile)    # (ny, nx, size, size), indexed per batch below


def _blend_tiles(tiles: torch.Tensor, ny: int, nx: int, step: int, ov: int) -> torch.Tensor:
    """Overlap-add of (n,1,S,S) tiles, S = step + 2*ov, placed every `step` pixels on a frame padded by `ov`: each tile is
    weighted by a separable trapezoid (linear ramps over the 2*ov overlap), the sum is normalised by the summed weights and
    the padding is cropped.  One F.fold for the tiles and one for the weights: no Python loop over tiles."""
    S = step + 2 * ov
    ramp = torch.ones(S, device=tiles.device, dtype=torch.float32)
    up = (torch.arange(2 * ov, device=tiles.device, dtype=torch.float32) + 0.5) / (2 * ov)
    ramp[:2 * ov] = up
    ramp[S - 2 * ov:] = up.flip(0)
    w = (ramp[:, None] * ramp[None, :]).reshape(1, S * S, 1)
    cols = tiles.reshape(ny * nx, S * S).t().unsqueeze(0)                     # (1, S*S, n), tiles in row-major grid order
    size = (ny * step + 2 * ov, nx * step + 2 * ov)
    num = F.fold(cols * w, size, kernel_size=S, stride=step)
    den = F.fold(w.expand(1, S * S, ny * nx), size, kernel_size=S, stride=step)
    return (num / den)[:, :, ov:ov + ny * step, ov:ov + nx * step]


@torch.no_grad()
def sr_frame_tiled(model: Callable[[torch.Tensor], torch.Tensor], frame: torch.Tensor, *, tile: int = 128, halo: int = 0,
                   scale: int = 4, batch: int = 16, rank: int = 0, world: int = 1, device=None, blend: bool = False):
    """Super-resolve `frame` ((H,W) or (1,1,H,W), float in [0,1]) tile by tile.

    Returns the (1,1,scale*H,scale*W) result on rank 0 (CPU tensor) and None on the other ranks.  `model` maps
    (b,1,t,t) -> (b,1,scale*t,scale*t) with t = tile + 2*halo.

    The frame is moved to `device` once, tiles are strided views of it, finished tiles stay on the device, the gather
    runs device to device (NCCL over NVLink when the process group is NCCL) and the frame is assembled with one
    permute on rank 0, so that the only host traffic is the frame in and the result out.

    blend=True (needs halo > 0): instead of keeping each haloed tile's centre crop, the overlapping borders of neighbouring
    tiles are cross-faded with a separable linear ramp over the 2*halo*scale overlap (weights sum to 1 after normalisation),
    which also hides the low-frequency disagreement between tiles that a centre crop leaves as a faint grid."""
    if frame.dim() == 2:
        frame = frame[None, None]
    H, W = frame.shape[-2:]
    origins = tile_origins(H, W, tile)
    ny, nx = H // tile, W // tile
    n = len(origins)
    mine = list(range(rank, n, world))
    device = torch.device(device) if device is not None else frame.device
    out_tile = tile * scale
    fdev = frame.to(device, non_blocking=True).float()
    tiles = _tile_view(fdev, tile, halo)                               # (ny, nx, t, t) view
    per = (n + world - 1) // world
    if blend and not halo:
        raise ValueError("blend=True needs halo > 0 (the overlap that is cross-faded)")
    keep = (tile + 2 * halo) * scale if blend else out_tile
    results = torch.zeros(per, 1, keep, keep, dtype=torch.float32, device=device)
    for s in range(0, len(mine), batch):
        idx = torch.tensor(mine[s:s + batch], device=device)
        lr = tiles[idx // nx, idx % nx].unsqueeze(1).contiguous()
        sr = model(lr).float()
        if halo and not blend:
            sr = sr[:, :, halo * scale:halo * scale + out_tile, halo * scale:halo * scale + out_tile]
        results[s:s + idx.numel()] = sr
    if world > 1:
        # final gather only (no data-path collective): every rank contributes `per` tile slots
        gathered = [torch.empty_like(results) for _ in range(world)] if rank == 0 else None
        dist.gather(results, gathered, dst=0)
        if rank != 0:
            return None
        # slot k of rank r is tile r + k*world: interleave the ranks back into row-major tile order
        allt = torch.stack(gathered, dim=1).reshape(per * world, 1, keep, keep)[:n]
    else:
        allt = results[:n]
    if blend:
        out = _blend_tiles(allt, ny, nx, out_tile, halo * scale)
    else:
        out = allt.reshape(ny, nx, out_tile, out_tile).permute(0, 2, 1, 3).reshape(1, 1, ny * out_tile, nx * out_tile)
    if out.is_cuda:
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=False)
        return host
    return out.contiguous()
