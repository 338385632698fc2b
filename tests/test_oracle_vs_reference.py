"""Live pin of the oracle against the unmodified reference at the REAL hot-path dimensions (C=180, 6 heads,
ws=8).  Runs only where the reference checkout exists (the build container); skipped on the GPU box."""
import os
import sys

import pytest
import torch

from oracle import swinir_oracle as o
from tests.util import rel_l2, randomize_

REF = os.environ.get("SR_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout absent")


def _ref_mod():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models.architecture_swin as m
    return m


def test_block_c180_matches_reference():
    m = _ref_mod()
    torch.manual_seed(0)
    blk = randomize_(m.SwinTransformerBlock(180, (16, 16), 6, window_size=8, shift_size=4), seed=9)
    x = torch.randn(1, 256, 180, requires_grad=True)
    y = blk(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    sd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in blk.state_dict().items()}
    x2 = x.detach().clone().requires_grad_(True)
    y2 = o.swin_block(x2, sd, "", (16, 16), 6, 8, 4)
    (y2 * w).sum().backward()
    assert rel_l2(y2, y) < 1e-5 and rel_l2(x2.grad, x.grad) < 1e-4
    for n, p in blk.named_parameters():
        assert rel_l2(sd[n].grad, p.grad) < 1e-4, n


def test_product_module_schema_matches_reference():
    m = _ref_mod()
    from superresolution_def_b200.architecture_swin import SwinIR
    kw = dict(upscale=4, in_chans=1, img_size=128, window_size=8, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6, mlp_ratio=2)
    ref, mine = m.SwinIR(**kw), SwinIR(**kw)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
    mine.load_state_dict(a, strict=True)
    assert all(torch.equal(a[k], b[k]) for k in a if "index" in k)
