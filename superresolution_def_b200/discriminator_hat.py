"""`models/discriminator_hat.py` of the reference (the discriminator `train_hat.py:26,138` builds next to the hybrid generator):
same class name, constructor, attribute names, registration order and spectral-norm parameters / buffers — so `state_dict()`
round-trips with strict=True — with forward and backward on libsrk (`disc_engine.UNetDiscriminatorHatFunction`): conv1..conv3
(4x4, stride 2) as tcgen05 GEMMs over a patch matrix, the bilinear x2 resizes as `srk_bilinear2x_fwd/bwd` (the additive skip
rides on the resize), conv4..conv8 on the implicit-GEMM 3x3 kernel, conv0 / conv9 on the single-channel kernels, spectral
normalisation as `srk_spectral_norm`.  The reference registers the class with basicsr's ARCH_REGISTRY (:7); nothing in the
scripts looks it up there, so the decorator is not reproduced.  There is no CPU path."""
from __future__ import annotations

from torch import nn as nn
from torch.nn.utils import spectral_norm


class UNetDiscriminatorSN(nn.Module):
    """U-Net discriminator with spectral normalisation (discriminator_hat.py:8-49): per-pixel logits at the input resolution."""

    def __init__(self, num_in_ch, num_feat=64, skip_connection=True):
        super().__init__()
        self.skip_connection = skip_connection
        norm = spectral_norm
        self.conv0 = nn.Conv2d(num_in_ch, num_feat, kernel_size=3, stride=1, padding=1)
        self.conv1 = norm(nn.Conv2d(num_feat, num_feat * 2, 4, 2, 1, bias=False))
        self.conv2 = norm(nn.Conv2d(num_feat * 2, num_feat * 4, 4, 2, 1, bias=False))
        self.conv3 = norm(nn.Conv2d(num_feat * 4, num_feat * 8, 4, 2, 1, bias=False))
        self.conv4 = norm(nn.Conv2d(num_feat * 8, num_feat * 4, 3, 1, 1, bias=False))
        self.conv5 = norm(nn.Conv2d(num_feat * 4, num_feat * 2, 3, 1, 1, bias=False))
        self.conv6 = norm(nn.Conv2d(num_feat * 2, num_feat, 3, 1, 1, bias=False))
        self.conv7 = norm(nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=False))
        self.conv8 = norm(nn.Conv2d(num_feat, num_feat, 3, 1, 1, bias=False))
        self.conv9 = nn.Conv2d(num_feat, 1, 3, 1, 1)

    def _sn_convs(self):
        return [self.conv1, self.conv2, self.conv3, self.conv4, self.conv5, self.conv6, self.conv7, self.conv8]

    def forward(self, x):
        from . import disc_engine
        sn = self._sn_convs()
        if len({m.training for m in sn}) != 1:
            raise RuntimeError("UNetDiscriminatorSN: mixed train / eval modes across the spectral-norm layers are not supported")
        eps = next(iter(sn[0]._forward_pre_hooks.values())).eps
        weights = [self.conv0.weight] + [m.weight_orig for m in sn] + [self.conv9.weight]
        return disc_engine.unet_discriminator_hat_sn(x, weights, self.conv0.bias, self.conv9.bias, [m.weight_u for m in sn],
                                                     [m.weight_v for m in sn], sn[0].training, eps, self.skip_connection)
