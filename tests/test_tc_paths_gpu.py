"""-m gpu: the Blackwell-native variants that are not the default dispatch run the SAME parity tests as the default kernels.

SRK_ATTN_TC=1  window attention (8x8) on tcgen05 / TMEM / TMA (csrc/attn_tc8.cuh: TMA-staged window quadrants, QK^T, PV and
               the five backward products as tcgen05.mma, softmax on tcgen05.ld rows) instead of the mma.sync kernels;
SRK_FUSED_MLP=1 the fused fc1 -> GELU -> fc2 -> residual -> LayerNorm kernel (csrc/mlp_fused.cuh) instead of two GEMM launches.
Both switches are read once per process, so the existing test modules are re-run in a subprocess with the switches set; the
parity bars (oracle, tolerances) are exactly those of the default path.  The fused MLP is additionally compared bit for bit
with the two-kernel path by tools/gpu_probe_mlp_fused.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra, timeout=900):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_tcgen05_attention_and_fused_mlp_pass_the_swin_and_hat8_parity_suites():
    r = _run(["-m", "pytest", "tests/test_swin_gpu.py", "tests/test_hat8_gpu.py", "-m", "gpu", "-x", "-q"],
             {"SRK_ATTN_TC": "1", "SRK_FUSED_MLP": "1"})
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


def test_tcgen05_attention_is_the_kernel_that_ran():
    """The switch must actually change the dispatch: with SRK_ATTN_TC=1 the shifted-window forward/backward at an eligible
    shape agree with the default kernels to bf16 rounding but not bit for bit (different accumulation order), and the library
    reports the same launch count."""
    code = ("import torch, sys; from superresolution_def_b200 import _capi as capi;"
            "torch.manual_seed(0); B,H,W,heads=2,16,32,6; T=B*H*W;"
            "qkv=torch.zeros(T,3,heads,32,device='cuda'); qkv[...,:30]=torch.randn(T,3,heads,30,device='cuda'); qkv=qkv.view(T,576).to(torch.bfloat16);"
            "tab=torch.randn(225,heads,device='cuda'); out=torch.empty(T,192,device='cuda',dtype=torch.bfloat16);"
            "capi.win_attn_fwd(capi.SrkGeom(B,H,W,8,4),heads,qkv,tab,out,ones_col=30); torch.cuda.synchronize();"
            "torch.save(out.cpu(), sys.argv[1])")
    import tempfile
    import torch
    with tempfile.TemporaryDirectory() as d:
        a, b = os.path.join(d, "a.pt"), os.path.join(d, "b.pt")
        r0 = _run(["-c", code, a], {"SRK_ATTN_TC": "0"})
        r1 = _run(["-c", code, b], {"SRK_ATTN_TC": "1"})
        assert r0.returncode == 0 and r1.returncode == 0, (r0.stderr[-1500:], r1.stderr[-1500:])
        x, y = torch.load(a).float(), torch.load(b).float()
    rel = ((x - y).norm() / x.norm()).item()
    assert rel < 2e-3, rel
    assert not torch.equal(x, y), "SRK_ATTN_TC=1 produced bit-identical output: the tcgen05 kernel did not run"


def test_fused_mlp_is_bit_identical_to_the_two_kernel_path():
    r = _run(["tools/gpu_probe_mlp_fused.py", "128", "1024", "18944"], {})
    assert r.returncode == 0 and "ALL OK" in r.stdout, (r.stdout[-3000:], r.stderr[-1500:])


def test_measured_alternatives_of_round_2_pass_the_kernel_unit_tests():
    """SRK_GEMM_BRES=0 (weight tile streamed per output tile instead of resident), SRK_WGRAD_AT=2 (two accumulators per
    weight-gradient CTA) and SRK_STORE_DACT=0 (gelu' recomputed in the fc2 input-gradient kernel) are A/B switches whose
    results must stay correct: the unit and block parity suites are re-run with them set."""
    r = _run(["-m", "pytest", "tests/test_kernels_r2_gpu.py", "tests/test_swin_gpu.py", "-m", "gpu", "-x", "-q"],
             {"SRK_GEMM_BRES": "0", "SRK_WGRAD_AT": "2", "SRK_STORE_DACT": "0"})
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


# ---- window 16 (HAT): the tcgen05 / TMEM / TMA forward kernel (csrc/attn_tc16.cuh) IS the default dispatch; the mma.sync
# kernel (SRK_ATTN16_TC=0, read per call) stays as the A/B partner.  test_hat_gpu.py compares the default against torch / the
# oracle; here the two kernels are compared with each other on the shapes the HAT models produce, incl. ragged image sizes,
# the masked last window row / column of a shifted block and OCAB windows whose 24x24 halo leaves the image on every side.
@pytest.mark.parametrize("mode,shift", [("self", 0), ("self", 8), ("oca", 0)])
@pytest.mark.parametrize("shape", [(1, 16, 16), (2, 32, 48), (1, 64, 64), (3, 48, 16)])
def test_window16_forward_tcgen05_matches_mma_sync(mode, shift, shape):
    import torch
    from superresolution_def_b200 import _capi as capi
    B, H, W = shape
    heads, T = 6, B * H * W
    g = torch.Generator(device="cuda").manual_seed(11 + shift + H)
    qkv = torch.zeros(T, 3, heads, 32, device="cuda")
    qkv[..., :30] = torch.randn(T, 3, heads, 30, device="cuda", generator=g) * 1.5
    qkv = qkv.view(T, 576).to(torch.bfloat16)
    m = capi.ATTN_SELF if mode == "self" else capi.ATTN_OCA
    table = torch.randn(961 if mode == "self" else 1521, heads, device="cuda", generator=g)
    res = {}
    old = os.environ.get("SRK_ATTN16_TC")
    try:
        for tc in ("0", "1"):
            os.environ["SRK_ATTN16_TC"] = tc
            out = torch.full((T, heads * 32), float("nan"), device="cuda", dtype=torch.bfloat16)
            lse = torch.full((heads, T), float("nan"), device="cuda")
            n0 = capi.launch_count()
            capi.win_attn16_fwd(capi.SrkGeom(B, H, W, 16, shift), m, heads, qkv, table, out, lse, ones_col=30)
            torch.cuda.synchronize()
            assert capi.launch_count() == n0 + 1
            res[tc] = (out.float(), lse)
    finally:
        if old is None:
            os.environ.pop("SRK_ATTN16_TC", None)
        else:
            os.environ["SRK_ATTN16_TC"] = old
    (o0, l0), (o1, l1) = res["0"], res["1"]
    assert torch.isfinite(o1).all() and torch.isfinite(l1).all()
    rel = ((o1 - o0).norm() / o0.norm()).item()
    # both kernels round P to bf16 before P V (different summation order): agreement to bf16 rounding, tolerance 6e-3 rel-L2
    # on the output, 2e-3 absolute on the row log-sum-exp (fp32 in both)
    assert rel < 6e-3, rel
    assert (l1 - l0).abs().max().item() < 2e-3
    assert (o1[:, 30] == 1).all(), "bias-folding column of head 0"
    assert not torch.equal(o0, o1), "SRK_ATTN16_TC did not change the dispatch"


def test_window16_forward_tcgen05_large_logits_and_wide_bias_range():
    """The tcgen05 kernel subtracts an UPPER BOUND of the row maximum (raw-logit row max + table max) instead of the exact
    maximum; logits of +-150 with a bias table spanning +-30 (far beyond anything a trained HAT produces) must neither
    overflow nor flush a row to zero: output and log-sum-exp stay finite and agree with the mma.sync kernel."""
    import torch
    from superresolution_def_b200 import _capi as capi
    B, H, W, heads = 1, 32, 32, 6
    T = B * H * W
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.zeros(T, 3, heads, 32, device="cuda")
    qkv[..., :30] = torch.randn(T, 3, heads, 30, device="cuda", generator=g)
    qkv[:, :2] *= 5.0   # q, k: logits ~ N(0, 137^2)
    qkv = qkv.view(T, 576).to(torch.bfloat16)
    table = torch.randn(961, heads, device="cuda", generator=g) * 10.0
    res = {}
    try:
        for tc in ("0", "1"):
            os.environ["SRK_ATTN16_TC"] = tc
            out = torch.full((T, heads * 32), float("nan"), device="cuda", dtype=torch.bfloat16)
            lse = torch.full((heads, T), float("nan"), device="cuda")
            capi.win_attn16_fwd(capi.SrkGeom(B, H, W, 16, 8), capi.ATTN_SELF, heads, qkv, table, out, lse, ones_col=-1)
            torch.cuda.synchronize()
            res[tc] = (out.float(), lse)
    finally:
        os.environ.pop("SRK_ATTN16_TC", None)
    (o0, l0), (o1, l1) = res["0"], res["1"]
    assert torch.isfinite(o1).all() and torch.isfinite(l1).all()
    assert ((o1 - o0).norm() / o0.norm()).item() < 1e-2
    assert ((l1 - l0).abs() / l0.abs().clamp_min(1.0)).max().item() < 1e-4
