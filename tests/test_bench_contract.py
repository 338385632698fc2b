"""bench.py contract checks that need no GPU: the reference arm (the oracle port of the reference's CPU path) runs and
prints ONE JSON line with the keys the driver reads; the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "patches/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("SwinIR x4 train patches/s") and line["value"] > 0
    assert line["vs_baseline"] is None and line["scaling"] == "weak" and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_product_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return  # on a GPU box the product arm is exercised by the driver itself
    r = _run("--steps", "1", "--warmup", "0", timeout=300)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)


def test_hybrid_tail_parameter_order_matches_the_module():
    """HybridTailFunction receives the parameters in module order: conv_adapt, 30 per RRDB (rdb1..3 x conv1..5 x w,b),
    conv_body, conv_up, conv_hr, conv_last — the order its returned gradients are matched against."""
    from superresolution_def_b200.hybridmodels_hat import HybridHATRealESRGAN
    net = HybridHATRealESRGAN(img_size=16, embed_dim=90, depths=(1,), num_heads=(6,), window_size=8, num_rrdb=2, num_feat=48,
                              num_grow_ch=24)
    names = {id(p): n for n, p in net.named_parameters()}
    got = [names[id(p)] for p in net.tail_params()]
    want = ["conv_adapt.weight", "conv_adapt.bias"]
    for i in range(2):
        for j in (1, 2, 3):
            for k in range(1, 6):
                want += [f"rrdb_trunk.{i}.rdb{j}.conv{k}.weight", f"rrdb_trunk.{i}.rdb{j}.conv{k}.bias"]
    for m in ("conv_body", "conv_up", "conv_hr", "conv_last"):
        want += [f"{m}.weight", f"{m}.bias"]
    assert got == want
    assert set(got) | {n for n, _ in net.hat.named_parameters(prefix="hat")} == {n for n, _ in net.named_parameters()}
