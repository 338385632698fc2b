"""CPU tests: the oracle (oracle/swinir_oracle.py) reproduces the golden vectors minted from the UNMODIFIED
reference (tools/make_golden.py) — outputs and gradients, fp32, tolerance 1e-5 relative."""
import os

import pytest
import torch

from oracle import swinir_oracle as o
from tests.util import rel_l2

G = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5


def _load(name):
    return torch.load(os.path.join(G, name), weights_only=False)


def _req(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}


def test_window_helpers():
    f = _load("swin_window_helpers.pt")
    win = o.window_partition(f["x"], f["ws"])
    assert torch.equal(win, f["windows"])
    assert torch.equal(o.window_reverse(win, f["ws"], 8, 12), f["reversed"])
    assert torch.equal(o.window_reverse(win, f["ws"], 8, 12), f["x"])


def test_relative_position_index_matches_reference_buffer():
    f = _load("swin_window_attention.pt")
    assert torch.equal(o.relative_position_index(4), f["sd"]["relative_position_index"])


def test_window_attention_forward_backward_and_mask():
    f = _load("swin_window_attention.pt")
    sd = _req(f["sd"])
    x = f["x"].clone().requires_grad_(True)
    y = o.window_attention(x, sd, "", f["kw"]["heads"], f["kw"]["window"])
    assert rel_l2(y, f["y"]) < TOL
    ym = o.window_attention(x, sd, "", f["kw"]["heads"], f["kw"]["window"], mask=f["mask"])
    assert rel_l2(ym, f["y_masked"]) < TOL
    # gradient check: same projection weights as the generator (torch.randn_like after the forward, seed stream)
    # is not reproducible here, so verify linearity instead: d/dx of sum(y*y_ref) against finite differences
    g, = torch.autograd.grad((y * f["y"]).sum(), x)
    eps = 1e-3
    d = torch.randn_like(x)
    with torch.no_grad():
        yp = o.window_attention(f["x"] + eps * d, f["sd"], "", f["kw"]["heads"], f["kw"]["window"])
        ym_ = o.window_attention(f["x"] - eps * d, f["sd"], "", f["kw"]["heads"], f["kw"]["window"])
    fd = (((yp - ym_) / (2 * eps)) * f["y"]).sum()
    assert abs(fd.item() - (g * d).sum().item()) < 2e-2 * max(1.0, abs(fd.item()))


def test_mlp():
    f = _load("swin_mlp.pt")
    assert rel_l2(o.mlp(f["x"], f["sd"], ""), f["y"]) < TOL


@pytest.mark.parametrize("shift", [0, 2])
def test_swin_block_outputs_and_grads(shift):
    f = _load(f"swin_block_shift{shift}.pt")
    kw = f["kw"]
    sd = _req(f["sd"])
    x = f["x"].clone().requires_grad_(True)
    y = o.swin_block(x, sd, "", tuple(kw["res"]), kw["heads"], kw["ws"], kw["shift"])
    assert rel_l2(y, f["y"]) < TOL
    (y * f["w"]).sum().backward()
    assert rel_l2(x.grad, f["gx"]) < 1e-4
    for n, g in f["grads"].items():
        assert rel_l2(sd[n].grad, g) < 1e-4, n


def test_swinir_tiny_outputs_grads_and_padding():
    f = _load("swinir_tiny.pt")
    kw = {k: f["kw"][k] for k in ("img_size", "window_size", "depths", "num_heads", "upscale")}
    sd = _req(f["sd"])
    x = f["x"].clone().requires_grad_(True)
    y = o.swinir_forward(x, sd, **kw)
    assert y.shape == f["y"].shape and rel_l2(y, f["y"]) < TOL
    (y * f["w"]).sum().backward()
    assert rel_l2(x.grad, f["gx"]) < 1e-4
    for n, g in f["grads"].items():
        assert rel_l2(sd[n].grad, g) < 2e-4, n
    yp = o.swinir_forward(f["x_pad"], f["sd"], **kw)
    assert yp.shape == f["y_pad"].shape == (1, 1, 28, 24) and rel_l2(yp, f["y_pad"]) < TOL


def test_init_state_dict_schema_matches_golden_keys():
    f = _load("swinir_tiny.pt")
    sd = o.init_state_dict(img_size=8, window_size=4, embed_dim=24, depths=[2, 2], num_heads=[3, 3])
    assert list(sd.keys()) == list(f["sd"].keys())
    assert all(sd[k].shape == f["sd"][k].shape and sd[k].dtype == f["sd"][k].dtype for k in sd)
