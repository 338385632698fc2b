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


def test_hab_and_ocab_c180_ws16_match_reference():
    """HAT blocks at the BASELINE configs[2] dimensions (C=180, 6 heads, window 16, OCAB 24x24 keys), 32x32 tokens."""
    from tools import ref_shim
    from oracle import hat_oracle as ho
    m = ref_shim.hat_module()
    torch.manual_seed(1)
    hat = m.HAT(img_size=32, in_chans=1, embed_dim=180, depths=(1,), num_heads=(6,), window_size=16, upscale=4,
                upsampler="pixelshuffle", drop_path_rate=0.0)
    assert torch.equal(ho.rpi_sa(16), hat.relative_position_index_SA)
    assert torch.equal(ho.rpi_oca(16), hat.relative_position_index_OCA)
    mask = hat.calculate_mask((32, 32))
    assert torch.equal(ho.shift_mask(32, 32, 16, 8), mask)
    for kind in ("hab", "ocab"):
        if kind == "hab":
            blk = randomize_(m.HAB(180, (32, 32), 6, window_size=16, shift_size=8), seed=2)
        else:
            blk = randomize_(m.OCAB(180, (32, 32), 16, 0.5, 6, mlp_ratio=4), seed=3)
        x = torch.randn(1, 1024, 180, requires_grad=True)
        y = blk(x, (32, 32), hat.relative_position_index_SA, mask) if kind == "hab" else \
            blk(x, (32, 32), hat.relative_position_index_OCA)
        w = torch.randn_like(y)
        (y * w).sum().backward()
        sd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v)
              for k, v in blk.state_dict().items()}
        x2 = x.detach().clone().requires_grad_(True)
        y2 = ho.hab(x2, sd, "", (32, 32), 6, 16, 8, ho.rpi_sa(16), mask) if kind == "hab" else \
            ho.ocab(x2, sd, "", (32, 32), 6, 16, ho.rpi_oca(16))
        (y2 * w).sum().backward()
        assert rel_l2(y2, y) < 1e-5 and rel_l2(x2.grad, x.grad) < 1e-4, kind
        for n, p in blk.named_parameters():
            assert rel_l2(sd[n].grad, p.grad) < 2e-4, (kind, n)


def test_hat_module_schema_matches_reference():
    from tools import ref_shim
    m = ref_shim.hat_module()
    from superresolution_def_b200.hat_arch import HAT
    kw = dict(img_size=128, in_chans=1, embed_dim=180, depths=(6,) * 2, num_heads=(6,) * 2, window_size=16, upscale=4,
              upsampler="pixelshuffle")
    ref, mine = m.HAT(**kw), HAT(**kw)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
    mine.load_state_dict(a, strict=True)
    assert all(torch.equal(a[k], b[k]) for k in a if "index" in k)
    assert torch.equal(ref.calculate_mask((64, 48)), mine.calculate_mask((64, 48)))


def test_hab_drop_path_matches_reference_rng_order():
    """Training-mode HAB with stochastic depth: the oracle, fed the factors obtained by replaying the reference's two
    torch.rand draws (attention branch, then MLP branch), reproduces the reference output and gradients."""
    from tools import ref_shim
    from oracle import hat_oracle as ho
    m = ref_shim.hat_module()
    blk = randomize_(m.HAB(36, (8, 8), 3, window_size=4, shift_size=2, compress_ratio=3, squeeze_factor=6, drop_path=0.4),
                     seed=5).train()
    hat = m.HAT(img_size=8, in_chans=1, embed_dim=36, depths=(1,), num_heads=(3,), window_size=4, squeeze_factor=6,
                upsampler="pixelshuffle")
    rpi, mask = hat.relative_position_index_SA, hat.calculate_mask((8, 8))
    x = torch.randn(6, 64, 36, requires_grad=True)
    torch.manual_seed(123)
    y = blk(x, (8, 8), rpi, mask)
    torch.manual_seed(123)
    keep = 0.6
    drop = tuple(((keep + torch.rand((6, 1, 1))).floor_() / keep).reshape(6) for _ in range(2))
    assert 0 < sum(int((d == 0).sum()) for d in drop) < 12      # both kept and dropped samples occur
    w = torch.randn_like(y)
    (y * w).sum().backward()
    sd = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in blk.state_dict().items()}
    x2 = x.detach().clone().requires_grad_(True)
    y2 = ho.hab(x2, sd, "", (8, 8), 3, 4, 2, rpi, mask, drop=drop)
    (y2 * w).sum().backward()
    assert rel_l2(y2, y) < 1e-5 and rel_l2(x2.grad, x.grad) < 1e-4
    for n, p in blk.named_parameters():
        assert rel_l2(sd[n].grad, p.grad) < 2e-4, n


def test_hybrid_module_schema_and_oracle_match_reference():
    """HybridHATRealESRGAN at the script's configuration (train_hat.py:132-136, fewer blocks): the product mirror has the
    reference's state_dict / parameter order, and the oracle reproduces the reference's forward at C=90, window 8,
    OCAB 12x12, nf 48 / gc 24."""
    from tools import ref_shim
    from oracle import hat_oracle as ho
    h = ref_shim.hybrid_module()
    from superresolution_def_b200.hybridmodels_hat import HybridHATRealESRGAN
    kw = dict(img_size=16, in_chans=1, embed_dim=90, depths=(2,), num_heads=(6,), window_size=8, upscale=4, num_rrdb=1,
              num_feat=48, num_grow_ch=24)
    torch.manual_seed(0)
    ref, mine = h.HybridHATRealESRGAN(**kw), HybridHATRealESRGAN(**kw)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
    mine.load_state_dict(a, strict=True)
    randomize_(ref, seed=3, table_std=0.5).eval()
    x = torch.rand(1, 1, 16, 16)
    with torch.no_grad():
        want = ref(x)
        got = ho.hybrid_forward(x, ref.state_dict(), window_size=8, depths=(2,), num_heads=(6,), num_rrdb=1)
    assert rel_l2(got, want) < 2e-5, rel_l2(got, want)


def test_discriminator_oracle_equals_the_reference_module():
    """oracle.discriminator_oracle.UNetDiscriminatorSN vs models/discriminator_swin.py:43-84 imported unmodified, and the
    product mirror's schema: same state_dict keys / shapes / order (spectral-norm weight_orig, weight_u, weight_v
    included), strict load in every direction, identical logits and gradients on CPU in eval mode (no power iteration)
    and in train mode (one power iteration on both sides); the functional oracle (what the libsrk path is compared with on
    the GPU) equals the module form on the hook-normalised weights."""
    _ref_mod()
    from models.discriminator_swin import UNetDiscriminatorSN as RefD
    from oracle.discriminator_oracle import UNetDiscriminatorSN as OraD, unet_discriminator_forward
    from superresolution_def_b200.gan import UNetDiscriminatorSN
    torch.manual_seed(0)
    ref = RefD(num_in_ch=1, num_feat=16)
    ora = OraD(num_in_ch=1, num_feat=16)
    mine = UNetDiscriminatorSN(num_in_ch=1, num_feat=16)
    sd = ref.state_dict()
    for m in (ora, mine):
        assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert [n for n, _ in ref.named_parameters()] == [n for n, _ in m.named_parameters()]
        m.load_state_dict(sd, strict=True)
        ref.load_state_dict(m.state_dict(), strict=True)
    x = torch.rand(2, 1, 64, 64)
    ref.eval(); ora.eval(); mine.eval()
    with torch.no_grad():
        want = ref(x)
        assert torch.equal(ora(x), want)
        # the product mirror hands exactly these weights to libsrk (its own forward needs the GPU)
        from superresolution_def_b200.gan import _sn_weight
        assert torch.equal(unet_discriminator_forward(x, [_sn_weight(m) for m in mine._convs()]), want)
    ref.train(); ora.train()
    a, b = ora(x), ref(x)
    assert torch.equal(a, b)
    a.mean().backward(); b.mean().backward()
    for (n, p), (_, q) in zip(ora.named_parameters(), ref.named_parameters()):
        assert torch.equal(p.grad, q.grad), n
    # train mode: the mirror's hook call performs the same single power iteration as the reference's forward did
    mine.train()
    w_mine = [_sn_weight(m) for m in mine._convs()]
    for k in sd:
        if k.endswith("weight_u"):
            assert torch.equal(mine.state_dict()[k], ref.state_dict()[k]), k
    assert torch.equal(unet_discriminator_forward(x, w_mine), b)


def test_hat_discriminator_oracle_equals_the_reference_module():
    """oracle.discriminator_oracle.UNetDiscriminatorSNHat vs models/discriminator_hat.py:8-49 imported unmodified (basicsr's
    registry decorator supplied by tools/ref_shim.py), and the product mirror's schema: same state_dict keys / shapes /
    order, strict load in every direction, identical logits and gradients on CPU in eval and train mode, with and without
    the skip connections; the functional oracle equals the module form on the hook-normalised weights."""
    from tools import ref_shim
    ref_shim.install()
    from models.discriminator_hat import UNetDiscriminatorSN as RefD
    from oracle.discriminator_oracle import UNetDiscriminatorSNHat as OraD, unet_discriminator_hat_forward
    from superresolution_def_b200.discriminator_hat import UNetDiscriminatorSN
    for skip in (True, False):
        torch.manual_seed(0)
        ref = RefD(num_in_ch=1, num_feat=16, skip_connection=skip)
        ora = OraD(num_in_ch=1, num_feat=16, skip_connection=skip)
        mine = UNetDiscriminatorSN(num_in_ch=1, num_feat=16, skip_connection=skip)
        sd = ref.state_dict()
        for m in (ora, mine):
            assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
            assert [n for n, _ in ref.named_parameters()] == [n for n, _ in m.named_parameters()]
            m.load_state_dict(sd, strict=True)
            ref.load_state_dict(m.state_dict(), strict=True)
        x = torch.rand(2, 1, 32, 32)
        ref.eval(); ora.eval()
        with torch.no_grad():
            want = ref(x)
            assert torch.equal(ora(x), want)
            w = [c.weight for c in ora.convs()]   # after the forward: the hook-normalised weights of conv1 .. conv8
            assert torch.equal(unet_discriminator_hat_forward(x, w, ora.conv0.bias, ora.conv9.bias, skip), want)
        ref.train(); ora.train()
        a, b = ora(x), ref(x)
        assert torch.equal(a, b)
        a.mean().backward(); b.mean().backward()
        for (n, p), (_, q) in zip(ora.named_parameters(), ref.named_parameters()):
            assert torch.equal(p.grad, q.grad), n
