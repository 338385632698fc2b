"""The C-ABI library loads on a CPU-only box and exports every symbol include/srk.h declares (no compute calls)."""
import ctypes

import pytest
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "srk.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srk_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from superresolution_def_b200 import _build
    lib = ctypes.CDLL(str(_build.build()))
    syms = _declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.srk_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.srk_version()


def test_binding_rejects_cpu_tensors():
    import pytest
    import torch
    from superresolution_def_b200 import _capi as capi
    from superresolution_def_b200.architecture_swin import SwinTransformerBlock
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    blk = SwinTransformerBlock(180, (16, 16), 6, window_size=8)
    with pytest.raises((capi.SrkError, AssertionError, RuntimeError)):
        blk(torch.randn(1, 256, 180))  # no CPU fallback: must fail loudly


def test_shape_helpers_without_gpu():
    from superresolution_def_b200 import _capi as capi
    d = capi.SrkBlockDims(180, 192, 6, 30, 32, 720, 768)
    assert capi.block_weight_elems(d) == [110592, 110592, 36864, 36864, 147456, 147456, 147456, 147456]
    assert capi.block_bwd_scratch_floats(d, capi.SrkGeom(16, 128, 128, 8, 0)) > 0


def test_fp32_policy_is_loud_and_controllable(monkeypatch):
    """train_hat.py:222-251 trains in fp32 without autocast; the mirrors compute in bf16.  The downgrade warns once by default,
    raises under SRK_FP32_POLICY=error, and is silent under autocast."""
    import warnings
    import torch
    from superresolution_def_b200 import swin_engine as eng, _capi as capi
    x = torch.zeros(1, 1, 8, 8)
    monkeypatch.setattr(eng, "_fp32_warned", False)
    monkeypatch.setenv("SRK_FP32_POLICY", "warn")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        eng.check_precision(x)
        eng.check_precision(x)
    assert len([m for m in w if issubclass(m.category, RuntimeWarning)]) == 1
    monkeypatch.setenv("SRK_FP32_POLICY", "error")
    with pytest.raises(capi.SrkError):
        eng.check_precision(x)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        eng.check_precision(x)          # autocast on: the reference computes in reduced precision too
    eng.check_precision(x.to(torch.bfloat16))


def _prototypes():
    """{symbol: parameter count} parsed from include/srk.h (comments stripped; `void` = 0 parameters)."""
    text = (ROOT / "include" / "srk.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    protos = {}
    for m in re.finditer(r"\b(srk_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        protos[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return protos


def test_ctypes_signatures_match_the_header_arity():
    """Every entry point bound in _capi.py with explicit argtypes passes as many arguments as its prototype in include/srk.h
    declares: a prototype changed on one side only (an extra pointer, a dropped flag) would otherwise shift every following
    argument silently — ctypes cannot see C prototypes."""
    from superresolution_def_b200 import _capi as capi
    protos = _prototypes()
    assert len(protos) >= 50, len(protos)
    checked, bad = 0, []
    for name, n in protos.items():
        fn = getattr(capi.lib, name)
        if fn.argtypes is None:
            continue
        checked += 1
        if len(fn.argtypes) != n:
            bad.append((name, len(fn.argtypes), n))
    assert not bad, bad
    assert checked >= 45, checked
