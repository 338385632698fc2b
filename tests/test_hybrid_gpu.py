"""-m gpu parity tests of the hybrid generator's convolutional part (models/hybridmodels_hat.py:21-131): residual dense
blocks with virtual concatenation, RRDB, and the whole HybridHATRealESRGAN at the script's channel configuration
(embed_dim 90, window 8, num_feat 48, num_grow_ch 24; fewer blocks), CUDA kernels through the C ABI vs the oracle.
Tolerances: outputs rel-L2 <= 2e-2, gradients <= 6e-2 (the conv-tail bound of tests/test_conv_gpu.py: a dense block
accumulates five bf16 gradient contributions per channel slice, and bias gradients are cancelling sums over pixels) AND
within 1.6x (+1e-2) of the error of the oracle itself under bf16 autocast, parameter by parameter."""
import pytest
import torch

from tests.util import rel_l2, randomize_

pytestmark = pytest.mark.gpu
OUT_TOL = 2e-2
GRAD_TOL = 6e-2


def _ho():
    from oracle import hat_oracle as ho
    return ho


def _sd_of(mod):
    return {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in mod.state_dict().items()}


def _cmp_module(mod, ref_fn, x, out_tol=OUT_TOL, grad_tol=GRAD_TOL):
    xr, xa, xm = (x.clone().requires_grad_(True) for _ in range(3))
    sd, sda = _sd_of(mod), _sd_of(mod)
    ref = ref_fn(xr, sd)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        r16 = ref_fn(xa, sda)
    got = mod(xm)
    assert got.shape == ref.shape and rel_l2(got, ref) < out_tol, rel_l2(got, ref)
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (r16.float() * w).sum().backward()
    (got * w).sum().backward()
    assert rel_l2(xm.grad, xr.grad) < grad_tol, rel_l2(xm.grad, xr.grad)
    bad = {}
    for n, p in mod.named_parameters():
        mine, auto = rel_l2(p.grad, sd[n].grad), rel_l2(sda[n].grad, sd[n].grad)
        if mine > grad_tol or mine > 1.6 * auto + 1e-2:
            bad[n] = (round(mine, 4), round(auto, 4))
    assert not bad, bad


@pytest.mark.parametrize("nf,gc", [(48, 24), (64, 32)])
def test_rdb_matches_oracle(nf, gc):
    from superresolution_def_b200.hybridmodels_hat import ResidualDenseBlock
    ho = _ho()
    torch.manual_seed(1)
    blk = randomize_(ResidualDenseBlock(nf, gc), seed=2).cuda()
    x = torch.randn(2, nf, 24, 32, device="cuda")     # 24 rows: per-tap conv kernel everywhere (H % 16 != 0)
    _cmp_module(blk, lambda t, sd: ho.rdb(t, sd, ""), x)


@pytest.mark.parametrize("mode", ["1", "3", "5", "6"])
def test_rdb_halo_conv_matches_oracle(mode):
    """Same block at a shape the halo-resident and role-swapped conv kernels take (H % 32 == 0): SRK_CONV_HALO = 1
    production heuristic, 3 = halo-resident kernel for every layer the swapped kernel does not take, 5 = no role-swapped
    kernel, 6 = role-swapped kernel for every eligible layer.  Run in a subprocess: the mode is read once per process."""
    import os, subprocess, sys
    code = ("import torch; from tests.test_hybrid_gpu import _cmp_module, _ho; from tests.util import randomize_;"
            "from superresolution_def_b200.hybridmodels_hat import ResidualDenseBlock; ho = _ho(); torch.manual_seed(1);"
            "blk = randomize_(ResidualDenseBlock(48, 24), seed=2).cuda();"
            "_cmp_module(blk, lambda t, sd: ho.rdb(t, sd, ''), torch.randn(2, 48, 32, 48, device='cuda'));"
            "blk = randomize_(ResidualDenseBlock(64, 32), seed=3).cuda();"      # the class defaults (num_feat 64 / grow 32)
            "_cmp_module(blk, lambda t, sd: ho.rdb(t, sd, ''), torch.randn(1, 64, 64, 32, device='cuda')); print('ok')")
    env = dict(os.environ, SRK_CONV_HALO=mode)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_rrdb_matches_oracle():
    from superresolution_def_b200.hybridmodels_hat import RRDBBlock
    ho = _ho()
    torch.manual_seed(3)
    blk = randomize_(RRDBBlock(48, 24), seed=4).cuda()
    x = torch.randn(2, 48, 16, 32, device="cuda")
    _cmp_module(blk, lambda t, sd: ho.rrdb(t, sd, ""), x)


def test_hybrid_small_matches_oracle():
    """HybridHATRealESRGAN at the script's widths (train_hat.py:132-136) with one RHAG of 2 HAB + OCAB and 2 RRDBs."""
    from superresolution_def_b200.hybridmodels_hat import HybridHATRealESRGAN
    ho = _ho()
    torch.manual_seed(28)
    kw = dict(window_size=8, depths=(2,), num_heads=(6,))
    net = randomize_(HybridHATRealESRGAN(img_size=32, in_chans=1, embed_dim=90, upscale=4, num_rrdb=2, num_feat=48,
                                         num_grow_ch=24, **kw), seed=28, table_std=0.5).cuda().eval()
    x = torch.rand(2, 1, 32, 32, device="cuda")
    w = torch.randn(2, 1, 128, 128, device="cuda")

    def run_oracle(autocast):
        sd = _sd_of(net)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = ho.hybrid_forward(x, sd, num_rrdb=2, **kw)
        (out.float() * w).mean().backward()
        return out, sd

    ref, sd32 = run_oracle(False)
    r16, sd16 = run_oracle(True)
    got = net(x)
    assert got.shape == ref.shape == (2, 1, 128, 128)
    assert rel_l2(got, ref) < OUT_TOL, rel_l2(got, ref)
    assert rel_l2(got, ref) < 1.6 * rel_l2(r16, ref) + 1e-2
    (got.float() * w).mean().backward()
    bad = {}
    for n, p in net.named_parameters():
        mine, auto = rel_l2(p.grad, sd32[n].grad), rel_l2(sd16[n].grad, sd32[n].grad)
        if auto > 0.2:
            continue  # ill-conditioned under bf16 (a cancelling scalar sum such as hat.conv_last.bias): no information
        if mine > 1.6 * auto + 1e-2:
            bad[n] = (round(mine, 4), round(auto, 4))
    assert not bad, bad
