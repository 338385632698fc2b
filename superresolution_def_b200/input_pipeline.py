"""16-bit TIFF pairs -> augmented float32 batches on the GPU, and 16-bit TIFF output.

What the reference does per sample on the host (dataset/astronomical_dataset_swin.py:25-69, identical in the HAT copy):
PIL decode -> np.float32 -> /65535 -> torch.flip / torch.flip / torch.rot90 (three more copies) -> pinned collate -> H2D of
4 bytes per pixel.  At 1.8 k patches/s on 8 GPUs that loader is the next wall.  Here the host only decodes the TIFF into a
pinned uint16 staging buffer; the batch crosses PCIe at 2 bytes per pixel on a side stream while the previous step runs,
and ONE kernel (srk_u16_to_f32_aug) does the /65535 conversion with the flips / rotation as address arithmetic.

Dataset semantics are the reference's: the split file is a JSON list of {ground_path (LR), hubble_path (HR)}; paths
containing '/data/' are re-rooted under base_path/data (:19-23); an unreadable pair is replaced by a random other index
(:52-54); the augmentation draws are, per sample and in this order, random() > 0.5 (flip -1), random() > 0.5 (flip -2),
randint(0, 3) (rot90 k) (:58-67), from a `random.Random` the caller can seed.
"""
from __future__ import annotations

import json
import random
from pathlib import Path

import numpy as np
import torch

from . import _capi as capi


def read_tiff_u16(path) -> np.ndarray | None:
    """(H, W) uint16 array of a 16-bit greyscale TIFF, or None if it cannot be read (the reference returns None too)."""
    from PIL import Image
    try:
        p = Path(path)
        if not p.exists():
            return None
        img = Image.open(p)
        img.load()
        arr = np.array(img)   # a writable copy (torch.from_numpy refuses read-only views)
        if arr.ndim != 2:
            return None
        if arr.dtype != np.uint16:
            # np.array(img, dtype=float32) / 65535 of the reference accepts any integer mode; values beyond 16 bits
            # cannot be staged as uint16, so such files are rejected instead of silently wrapped
            if arr.min() < 0 or arr.max() > 65535:
                return None
            arr = arr.astype(np.uint16)
        return arr
    except Exception:  # noqa: BLE001 - same contract as the reference loader
        return None


def save_as_tiff16(tensor: torch.Tensor, path) -> None:
    """infer_hat.py:42-50 / infer_swin.py: clip to [0,1], * 65535, truncate to uint16, 16-bit TIFF.  The quantisation runs
    on the device (srk_f32_to_u16) when the tensor is on one, so the D2H copy is 2 bytes per pixel."""
    from PIL import Image
    t = tensor.detach().squeeze()
    if t.is_cuda:
        src = t.float().contiguous()
        dst = torch.empty(src.shape, dtype=torch.uint16, device=src.device)
        capi.f32_to_u16(src, dst)
        arr = dst.cpu().numpy()
    else:
        arr = (np.clip(t.float().numpy(), 0, 1) * 65535).astype(np.uint16)
    Image.fromarray(arr).save(str(path))   # a uint16 array is written as a 16-bit ("I;16") TIFF


def draw_aug_code(rng: random.Random) -> int:
    """fh | fv << 1 | k << 2 with the reference's draw order (astronomical_dataset_swin.py:58-67)."""
    fh = rng.random() > 0.5
    fv = rng.random() > 0.5
    k = rng.randint(0, 3)
    return int(fh) | (int(fv) << 1) | (k << 2)


def apply_aug_reference(t: torch.Tensor, code: int) -> torch.Tensor:
    """The reference's torch ops for one (1,H,W) / (H,W) tensor — used by the tests as the definition of `code`."""
    if code & 1:
        t = torch.flip(t, [-1])
    if code & 2:
        t = torch.flip(t, [-2])
    k = (code >> 2) & 3
    if k:
        t = torch.rot90(t, k, [-2, -1])
    return t.contiguous()


class TiffPairDataset:
    """Index -> (lr_u16, hr_u16) numpy planes with the reference dataset's path handling and fallback."""

    def __init__(self, split_file, base_path, rng: random.Random | None = None):
        self.base_path = Path(base_path)
        with open(split_file, "r") as f:
            self.pairs = json.load(f)
        self.rng = rng or random.Random()

    def __len__(self):
        return len(self.pairs)

    def _fix_path(self, path_str: str) -> Path:
        if "/data/" in path_str:
            return self.base_path / "data" / path_str.split("/data/", 1)[1]
        return self.base_path / path_str

    def load(self, idx: int):
        for _ in range(64):
            pair = self.pairs[idx]
            lr = read_tiff_u16(self._fix_path(str(pair["ground_path"])))
            hr = read_tiff_u16(self._fix_path(str(pair["hubble_path"])))
            if lr is not None and hr is not None:
                return lr, hr
            idx = self.rng.randint(0, len(self.pairs) - 1)
        raise RuntimeError("no readable LR/HR pair found in 64 attempts")


class GpuBatchLoader:
    """Iterates `{'lr': (B,1,h,h) float32, 'hr': (B,1,H,H) float32}` CUDA batches over a TiffPairDataset.

    Double-buffered: while the consumer trains on batch i, batch i+1 is decoded into pinned uint16 buffers, copied on a
    side stream and converted + augmented by one kernel launch per tensor; the consumer's stream waits on an event only.
    `indices` is the (already sharded / shuffled) sample order of this rank, e.g. list(DistributedSampler(...))."""

    def __init__(self, dataset: TiffPairDataset, indices, batch_size: int, device, augment: bool = True,
                 drop_last: bool = True, seed: int | None = None):
        self.ds, self.indices, self.B = dataset, list(indices), batch_size
        self.device = torch.device(device)
        self.augment, self.drop_last = augment, drop_last
        self.rng = random.Random(seed)
        self.stream = torch.cuda.Stream(device=self.device)
        self._bufs = None

    def __len__(self):
        n = len(self.indices)
        return n // self.B if self.drop_last else (n + self.B - 1) // self.B

    def _alloc(self, b, h, H):
        mk = lambda n: (torch.empty(b, n, n, dtype=torch.uint16).pin_memory(),                       # noqa: E731
                        torch.empty(b, n, n, dtype=torch.uint16, device=self.device),
                        torch.empty(b, 1, n, n, dtype=torch.float32, device=self.device))
        return {"lr": mk(h), "hr": mk(H), "codes_h": torch.empty(b, dtype=torch.int32).pin_memory(),
                "codes_d": torch.empty(b, dtype=torch.int32, device=self.device), "event": torch.cuda.Event()}

    def _stage(self, slot, idxs):
        pairs = [self.ds.load(i) for i in idxs]
        h, H = pairs[0][0].shape[-1], pairs[0][1].shape[-1]
        b = len(pairs)
        if self._bufs is None:
            self._bufs = [self._alloc(self.B, h, H), self._alloc(self.B, h, H)]
        buf = self._bufs[slot]
        buf["event"].synchronize()   # the previous H2D copies out of this slot's pinned buffers have completed
        for j, (lr, hr) in enumerate(pairs):
            if lr.shape != (h, h) or hr.shape != (H, H):
                raise capi.SrkError("GpuBatchLoader needs square planes of one size per tensor (n % 32 == 0)")
            buf["lr"][0][j].copy_(torch.from_numpy(lr))
            buf["hr"][0][j].copy_(torch.from_numpy(hr))
            buf["codes_h"][j] = draw_aug_code(self.rng) if self.augment else 0
        # the previous consumer of this slot's device tensors must be done before they are overwritten
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            buf["codes_d"].copy_(buf["codes_h"], non_blocking=True)
            for k in ("lr", "hr"):
                host, dev, out = buf[k]
                dev[:b].copy_(host[:b], non_blocking=True)
                capi.u16_to_f32_aug(dev[:b], out[:b], buf["codes_d"][:b] if self.augment else None)
            buf["event"].record(self.stream)
        return buf, b

    def __iter__(self):
        nb = len(self)
        batches = [self.indices[i * self.B:(i + 1) * self.B] for i in range(nb)]
        nxt = self._stage(0, batches[0]) if nb else None
        for i in range(nb):
            buf, b = nxt
            nxt = self._stage((i + 1) & 1, batches[i + 1]) if i + 1 < nb else None
            torch.cuda.current_stream(self.device).wait_event(buf["event"])
            yield {"lr": buf["lr"][2][:b], "hr": buf["hr"][2][:b]}
