"""CPU tests: oracle/discriminator_oracle.py reproduces the golden vectors minted from the UNMODIFIED reference
discriminators (tools/make_golden.py disc: models/discriminator_swin.py and models/discriminator_hat.py at num_feat = 8,
train mode) — logits, the spectral-norm buffers after the forward's power iteration, and every gradient.  These fixtures
travel with the repository; tests/test_oracle_vs_reference.py repeats the comparison live (bit-exact) where /root/reference
exists."""
import os

import pytest
import torch

from oracle import discriminator_oracle as do
from tests.util import rel_l2

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,cls", [("disc_swin_tiny.pt", do.UNetDiscriminatorSN), ("disc_hat_tiny.pt", do.UNetDiscriminatorSNHat)])
def test_discriminator_oracle_reproduces_the_reference_fixture(name, cls):
    f = torch.load(os.path.join(G, name), weights_only=False)
    d = cls(**f["kw"]).train()
    assert [(k, tuple(v.shape)) for k, v in d.state_dict().items()] == [(k, tuple(v.shape)) for k, v in f["sd"].items()]
    d.load_state_dict(f["sd"], strict=True)
    x = f["x"].clone().requires_grad_(True)
    y = d(x)
    assert y.shape == f["y"].shape and rel_l2(y, f["y"]) < 2e-5, rel_l2(y, f["y"])
    for k, v in f["sn_after"].items():                       # one power iteration, in place on the buffers
        assert rel_l2(d.state_dict()[k], v) < 1e-6, k
    (y * f["w"]).sum().backward()
    assert rel_l2(x.grad, f["gx"]) < 5e-4
    for n, p in d.named_parameters():
        assert rel_l2(p.grad, f["grads"][n]) < 5e-4, (n, rel_l2(p.grad, f["grads"][n]))
    # the functional form (what the GPU tests compare libsrk with) on the weights the hook just produced
    ws = [c.weight.detach() for c in d.convs()]
    with torch.no_grad():
        if cls is do.UNetDiscriminatorSN:
            y2 = do.unet_discriminator_forward(f["x"], ws)
        else:
            y2 = do.unet_discriminator_hat_forward(f["x"], ws, d.conv0.bias, d.conv9.bias, d.skip_connection)
    assert rel_l2(y2, f["y"]) < 2e-5
