"""CPU tests: oracle/hat_oracle.py reproduces the golden vectors minted from the UNMODIFIED reference HAT modules
(tools/make_golden.py hat) — outputs and every gradient, fp32."""
import os

import pytest
import torch

from oracle import hat_oracle as o
from tests.util import rel_l2

G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5
GTOL = 2e-4


def _load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


def _req(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}


def _check_grads(sd, x, f, tol=GTOL):
    assert rel_l2(x.grad, f["gx"]) < tol
    for n, g in f["grads"].items():
        assert rel_l2(sd[n].grad, g) < tol, n


def test_rpi_and_mask_builders():
    f = _load("hat_rpi_mask.pt")
    assert torch.equal(o.rpi_sa(f["ws"]), f["rpi_sa"])
    assert torch.equal(o.rpi_oca(f["ws"], 0.5), f["rpi_oca"])
    assert (f["rpi_oca"] < 0).any()  # the reference's wrap-around quirk is present in the pin
    assert torch.equal(o.shift_mask(8, 12, f["ws"], f["ws"] // 2), f["mask_8x12"])
    assert torch.equal(o.shift_mask(8, 8, f["ws"], f["ws"] // 2), f["mask_8x8"])


def test_window_attention_rpi_mask():
    f = _load("hat_window_attention.pt")
    r = _load("hat_rpi_mask.pt")
    sd = _req(f["sd"])
    x = f["x"].clone().requires_grad_(True)
    y = o.window_attention(x, sd, "", f["heads"], r["rpi_sa"], r["mask_8x12"])
    assert rel_l2(y, f["y"]) < TOL
    assert rel_l2(o.window_attention(x, sd, "", f["heads"], r["rpi_sa"], None), f["y_nomask"]) < TOL
    (y * (torch.ones_like(y) * torch.linspace(-1, 1, y.shape[-1]))).sum().backward()
    _check_grads(sd, x, f)


def test_cab():
    f = _load("hat_cab.pt")
    sd = _req(f["sd"])
    x = f["x"].clone().requires_grad_(True)
    y = o.cab(x, sd, "")
    assert rel_l2(y, f["y"]) < TOL
    (y * f["w"]).sum().backward()
    _check_grads(sd, x, f)


@pytest.mark.parametrize("shift", [0, 2])
def test_hab(shift):
    f = _load(f"hat_hab_shift{shift}.pt")
    r = _load("hat_rpi_mask.pt")
    kw = f["kw"]
    sd = _req(f["sd"])
    x = f["x"].clone().requires_grad_(True)
    y = o.hab(x, sd, "", tuple(kw["res"]), kw["heads"], kw["ws"], kw["shift"], r["rpi_sa"], r["mask_8x12"])
    assert rel_l2(y, f["y"]) < TOL
    (y * f["w"]).sum().backward()
    _check_grads(sd, x, f)


def test_ocab():
    f = _load("hat_ocab.pt")
    r = _load("hat_rpi_mask.pt")
    kw = f["kw"]
    sd = _req(f["sd"])
    x = f["x"].clone().requires_grad_(True)
    y = o.ocab(x, sd, "", tuple(kw["res"]), kw["heads"], kw["ws"], r["rpi_oca"])
    assert rel_l2(y, f["y"]) < TOL
    (y * f["w"]).sum().backward()
    _check_grads(sd, x, f)


def test_hat_tiny():
    f = _load("hat_tiny.pt")
    kw = f["kw"]
    sd = _req(f["sd"])
    x = f["x"].clone().requires_grad_(True)
    y = o.hat_forward(x, sd, window_size=kw["window_size"], depths=kw["depths"], num_heads=kw["num_heads"], upscale=4)
    assert y.shape == f["y"].shape and rel_l2(y, f["y"]) < TOL
    (y * f["w"]).sum().backward()
    _check_grads(sd, x, f, 5e-4)


def test_hybrid_tiny():
    f = _load("hybrid_tiny.pt")
    kw = f["kw"]
    sd = _req(f["sd"])
    x = f["x"].clone().requires_grad_(True)
    y = o.hybrid_forward(x, sd, window_size=kw["window_size"], depths=kw["depths"], num_heads=kw["num_heads"],
                         num_rrdb=kw["num_rrdb"])
    assert y.shape == f["y"].shape == (1, 1, 32, 32) and rel_l2(y, f["y"]) < TOL
    (y * f["w"]).sum().backward()
    _check_grads(sd, x, f, 5e-4)
