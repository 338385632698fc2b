"""Input / output formats around the path.  CPU: host logic of the TIFF pair dataset (path re-rooting,
fallback on unreadable pairs, augmentation draw order, 16-bit TIFF round trip) — against the reference's own
AstronomicalDataset where /root/reference exists.  GPU: srk_u16_to_f32_aug bit-exact against the reference's torch ops for
all 16 augmentation codes, srk_f32_to_u16 bit-exact against numpy's clip/scale/astype, and the double-buffered loader."""
import json
import os
import random
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"


def _make_tree(tmp, n=6, h=32, H=128, seed=0):
    from PIL import Image
    rng = np.random.default_rng(seed)
    root = tmp / "proj"
    d = root / "data" / "m33" / "8_dataset_split"
    d.mkdir(parents=True)
    pairs = []
    for i in range(n):
        lr = rng.integers(0, 65536, size=(h, h), dtype=np.uint16)
        hr = rng.integers(0, 65536, size=(H, H), dtype=np.uint16)
        pl, ph = d / f"lr_{i}.tiff", d / f"hr_{i}.tiff"
        Image.fromarray(lr).save(str(pl))
        Image.fromarray(hr).save(str(ph))
        # the split files carry absolute paths of the machine that built them: '/data/' is the re-rooting marker
        pairs.append({"ground_path": f"/somewhere/else/data/m33/8_dataset_split/lr_{i}.tiff",
                      "hubble_path": f"/somewhere/else/data/m33/8_dataset_split/hr_{i}.tiff", "_lr": lr, "_hr": hr})
    split = root / "train.json"
    split.write_text(json.dumps([{k: v for k, v in p.items() if not k.startswith("_")} for p in pairs]))
    return root, split, pairs


def test_dataset_paths_fallback_and_tiff16_roundtrip(tmp_path):
    from superresolution_def_b200.input_pipeline import TiffPairDataset, read_tiff_u16, save_as_tiff16
    root, split, pairs = _make_tree(tmp_path)
    ds = TiffPairDataset(split, root, rng=random.Random(3))
    assert len(ds) == 6
    lr, hr = ds.load(2)
    assert lr.dtype == np.uint16 and np.array_equal(lr, pairs[2]["_lr"]) and np.array_equal(hr, pairs[2]["_hr"])
    os.remove(root / "data" / "m33" / "8_dataset_split" / "hr_4.tiff")      # unreadable pair -> some other, readable pair
    lr4, hr4 = ds.load(4)
    assert any(np.array_equal(lr4, p["_lr"]) and np.array_equal(hr4, p["_hr"]) for i, p in enumerate(pairs) if i != 4)
    x = torch.rand(1, 1, 64, 64) * 1.2 - 0.1
    save_as_tiff16(x, tmp_path / "o.tiff")
    back = read_tiff_u16(tmp_path / "o.tiff")
    assert np.array_equal(back, (np.clip(x.squeeze().numpy(), 0, 1) * 65535).astype(np.uint16))


def test_aug_code_draw_order_matches_the_reference_ops():
    from superresolution_def_b200.input_pipeline import draw_aug_code, apply_aug_reference
    # same draws, consumed in the reference's order (random() > .5, random() > .5, randint(0, 3))
    a, b = random.Random(11), random.Random(11)
    t = torch.arange(16.0).reshape(1, 4, 4)
    for _ in range(50):
        code = draw_aug_code(a)
        ref = t
        if b.random() > 0.5:
            ref = torch.flip(ref, [-1])
        if b.random() > 0.5:
            ref = torch.flip(ref, [-2])
        k = b.randint(0, 3)
        if k > 0:
            ref = torch.rot90(ref, k, [-2, -1])
        assert torch.equal(apply_aug_reference(t, code), ref.contiguous())


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_dataset_equals_the_reference_dataset_without_augmentation(tmp_path):
    from superresolution_def_b200.input_pipeline import TiffPairDataset
    root, split, _ = _make_tree(tmp_path)
    sys.path.insert(0, REF)
    try:
        from dataset.astronomical_dataset_swin import AstronomicalDataset
        ref = AstronomicalDataset(str(split), base_path=root, augment=False)
    finally:
        sys.path.remove(REF)
    ds = TiffPairDataset(split, root)
    assert len(ref) == len(ds)
    for i in range(len(ds)):
        r = ref[i]
        lr, hr = ds.load(i)
        # the staged uint16 planes divided in float32 are exactly the reference's tensors
        assert torch.equal(torch.from_numpy(lr.astype(np.float32) / np.float32(65535.0))[None], r["lr"])
        assert torch.equal(torch.from_numpy(hr.astype(np.float32) / np.float32(65535.0))[None], r["hr"])


@pytest.mark.gpu
def test_u16_aug_kernel_is_bit_exact_for_all_codes():
    from superresolution_def_b200 import _capi as capi
    from superresolution_def_b200.input_pipeline import apply_aug_reference
    g = torch.Generator().manual_seed(0)
    for n in (32, 128):
        src = torch.randint(0, 65536, (16, n, n), generator=g, dtype=torch.int32).to(torch.uint16)
        codes = torch.arange(16, dtype=torch.int32)
        out = torch.empty(16, 1, n, n, device="cuda")
        capi.u16_to_f32_aug(src.cuda(), out, codes.cuda())
        ref = torch.stack([apply_aug_reference(src[b].to(torch.float32) / 65535.0, int(codes[b])) for b in range(16)])[:, None]
        assert torch.equal(out.cpu(), ref), n
        capi.u16_to_f32_aug(src.cuda(), out, None)
        assert torch.equal(out.cpu(), (src.to(torch.float32) / 65535.0)[:, None])


@pytest.mark.gpu
def test_f32_to_u16_kernel_matches_numpy_quantisation():
    from superresolution_def_b200 import _capi as capi
    x = torch.rand(3, 515, generator=torch.Generator().manual_seed(1)) * 1.4 - 0.2
    x[0, :4] = torch.tensor([0.0, 1.0, 0.99999, 1e-6])
    dst = torch.empty(x.shape, dtype=torch.uint16, device="cuda")
    capi.f32_to_u16(x.cuda(), dst)
    assert np.array_equal(dst.cpu().numpy(), (np.clip(x.numpy(), 0, 1) * 65535).astype(np.uint16))


@pytest.mark.gpu
def test_gpu_batch_loader_matches_reference_ops(tmp_path):
    from superresolution_def_b200.input_pipeline import TiffPairDataset, GpuBatchLoader, draw_aug_code, apply_aug_reference
    root, split, pairs = _make_tree(tmp_path, n=8, h=32, H=128)
    ds = TiffPairDataset(split, root)
    order = [5, 0, 3, 7, 1, 2, 6, 4]
    loader = GpuBatchLoader(ds, order, batch_size=3, device="cuda", augment=True, seed=123)
    assert len(loader) == 2
    rng = random.Random(123)
    seen = 0
    for bi, batch in enumerate(loader):
        assert batch["lr"].shape == (3, 1, 32, 32) and batch["hr"].shape == (3, 1, 128, 128) and batch["lr"].is_cuda
        lr, hr = batch["lr"].cpu(), batch["hr"].cpu()   # consume before the slot is reused
        for j in range(3):
            p = pairs[order[bi * 3 + j]]
            code = draw_aug_code(rng)
            for got, raw in ((lr[j], p["_lr"]), (hr[j], p["_hr"])):
                want = apply_aug_reference(torch.from_numpy(raw.astype(np.float32) / np.float32(65535.0))[None], code)
                assert torch.equal(got, want)
            seen += 1
    assert seen == 6
