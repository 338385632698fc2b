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
