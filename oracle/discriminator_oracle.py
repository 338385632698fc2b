"""ORACLE (test infrastructure, never imported by the product): ATen restatement of the reference's U-Net discriminator.

Follows models/discriminator_swin.py of GDev96/SuperResolution_Def:
  UNetConv2        :6-19   spectral_norm(Conv2d(in, out, 4, 2, 1, bias=False)) + LeakyReLU(0.2)
  UNetUpBlock      :21-41  spectral_norm(ConvTranspose2d(in, out, 4, 2, 1, bias=False)) + LeakyReLU(0.2), bilinear resize
                           (align_corners=True) to the skip's size when they differ, torch.cat((x, skip), 1)
  UNetDiscriminatorSN :43-84

Two forms: `unet_discriminator_forward` is the functional restatement on already-normalised weights (what
superresolution_def_b200.disc_engine computes); `UNetDiscriminatorSN` is the module form (same tree, same spectral-norm
hooks) whose state_dict is interchangeable with the reference's and with the product mirror's.

Parity pin: tests/test_oracle_vs_reference.py::test_discriminator_oracle_equals_the_reference_module runs this module and the
unmodified reference on the same weights and input on CPU and requires identical logits and gradients (torch.equal), in
eval mode and in train mode (one power iteration on both sides); the functional form is pinned against the module form in
the same test.  tests/golden/disc_swin_tiny.pt / disc_hat_tiny.pt (tools/make_golden.py disc) hold reference outputs and gradients
for machines without the reference (tests/test_disc_oracle_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

SLOPE = 0.2


def unet_discriminator_forward(x, w):
    """x [B,1,H,W]; w = the 12 normalised weights in forward order:
    conv0[0] 3x3, conv0[2] 4x4s2, conv1..conv4 4x4s2, up1..up4 transposed 4x4s2, final_conv[0] 3x3, final_conv[2] 3x3."""
    lr = lambda t: F.leaky_relu(t, SLOPE)
    a0 = lr(F.conv2d(x, w[0], None, 1, 1))                      # :49-50
    x0 = lr(F.conv2d(a0, w[1], None, 2, 1))                     # :51-52
    x1 = lr(F.conv2d(x0, w[2], None, 2, 1))                     # :74
    x2 = lr(F.conv2d(x1, w[3], None, 2, 1))
    x3 = lr(F.conv2d(x2, w[4], None, 2, 1))
    x4 = lr(F.conv2d(x3, w[5], None, 2, 1))                     # :77

    def up(t, skip, wt):                                        # :33-41
        t = lr(F.conv_transpose2d(t, wt, None, 2, 1))
        if t.shape[-2:] != skip.shape[-2:]:
            t = F.interpolate(t, size=skip.shape[-2:], mode="bilinear", align_corners=True)
        return torch.cat((t, skip), 1)

    d = up(x4, x3, w[6])
    d = up(d, x2, w[7])
    d = up(d, x1, w[8])
    d = up(d, x0, w[9])
    return F.conv2d(lr(F.conv2d(d, w[10], None, 1, 1)), w[11], None, 1, 1)   # :66-70, :83


class UNetConv2(nn.Module):
    def __init__(self, in_size, out_size, dropout=0.0):
        super().__init__()
        layers = [spectral_norm(nn.Conv2d(in_size, out_size, 4, 2, 1, bias=False)), nn.LeakyReLU(SLOPE, inplace=True)]
        if dropout > 0:
            layers.append(nn.Dropout(dropout))
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class UNetUpBlock(nn.Module):
    def __init__(self, in_size, out_size, dropout=0.0):
        super().__init__()
        layers = [spectral_norm(nn.ConvTranspose2d(in_size, out_size, 4, 2, 1, bias=False)), nn.LeakyReLU(SLOPE, inplace=True)]
        if dropout > 0:
            layers.append(nn.Dropout(dropout))
        self.model = nn.Sequential(*layers)

    def forward(self, x, skip_input):
        x = self.model(x)
        if x.shape[-2:] != skip_input.shape[-2:]:
            x = F.interpolate(x, size=skip_input.shape[-2:], mode="bilinear", align_corners=True)
        return torch.cat((x, skip_input), 1)


class UNetDiscriminatorSN(nn.Module):
    def __init__(self, num_in_ch=1, num_feat=64, skip_connection=True):
        super().__init__()
        self.skip_connection = skip_connection
        nf = num_feat
        self.conv0 = nn.Sequential(spectral_norm(nn.Conv2d(num_in_ch, nf, 3, 1, 1, bias=False)), nn.LeakyReLU(SLOPE, inplace=True),
                                   spectral_norm(nn.Conv2d(nf, nf, 4, 2, 1, bias=False)), nn.LeakyReLU(SLOPE, inplace=True))
        self.conv1 = UNetConv2(nf, nf * 2)
        self.conv2 = UNetConv2(nf * 2, nf * 4)
        self.conv3 = UNetConv2(nf * 4, nf * 8)
        self.conv4 = UNetConv2(nf * 8, nf * 8)
        self.up1 = UNetUpBlock(nf * 8, nf * 8)
        self.up2 = UNetUpBlock(nf * 16, nf * 4)
        self.up3 = UNetUpBlock(nf * 8, nf * 2)
        self.up4 = UNetUpBlock(nf * 4, nf)
        self.final_conv = nn.Sequential(spectral_norm(nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=False)), nn.LeakyReLU(SLOPE, inplace=True),
                                        spectral_norm(nn.Conv2d(nf, 1, 3, 1, 1, bias=False)))

    def convs(self):
        return [self.conv0[0], self.conv0[2], self.conv1.model[0], self.conv2.model[0], self.conv3.model[0],
                self.conv4.model[0], self.up1.model[0], self.up2.model[0], self.up3.model[0], self.up4.model[0],
                self.final_conv[0], self.final_conv[2]]

    def forward(self, x):
        x0 = self.conv0(x)
        x1 = self.conv1(x0)
        x2 = self.conv2(x1)
        x3 = self.conv3(x2)
        x4 = self.conv4(x3)
        d = self.up1(x4, x3)
        d = self.up2(d, x2)
        d = self.up3(d, x1)
        d = self.up4(d, x0)
        return self.final_conv(d)


# ----------------------------------------------------------------------------------------------------------------------
# models/discriminator_hat.py:8-49 (the discriminator of train_hat.py): functional restatement on normalised weights and the
# module form.  Pinned by tests/test_oracle_vs_reference.py::test_hat_discriminator_oracle_equals_the_reference_module
# (the reference file imported unmodified through tools/ref_shim.py, which supplies basicsr's arithmetic-free registry).
# ----------------------------------------------------------------------------------------------------------------------
def unet_discriminator_hat_forward(x, w, b0, b9, skip=True):
    """w = conv0 .. conv9 weights (conv1 .. conv8 normalised), b0 / b9 the biases of the two plain convolutions."""
    lr = lambda t: F.leaky_relu(t, SLOPE)
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    x0 = lr(F.conv2d(x, w[0], b0, 1, 1))          # :26
    x1 = lr(F.conv2d(x0, w[1], None, 2, 1))
    x2 = lr(F.conv2d(x1, w[2], None, 2, 1))
    x3 = lr(F.conv2d(x2, w[3], None, 2, 1))       # :29
    x4 = lr(F.conv2d(up(x3), w[4], None, 1, 1))   # :31-32
    if skip:
        x4 = x4 + x2
    x5 = lr(F.conv2d(up(x4), w[5], None, 1, 1))   # :36-37
    if skip:
        x5 = x5 + x1
    x6 = lr(F.conv2d(up(x5), w[6], None, 1, 1))   # :41-42
    if skip:
        x6 = x6 + x0
    out = lr(F.conv2d(x6, w[7], None, 1, 1))      # :47
    out = lr(F.conv2d(out, w[8], None, 1, 1))
    return F.conv2d(out, w[9], b9, 1, 1)          # :49


class UNetDiscriminatorSNHat(nn.Module):
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

    def convs(self):
        return [self.conv0, self.conv1, self.conv2, self.conv3, self.conv4, self.conv5, self.conv6, self.conv7, self.conv8,
                self.conv9]

    def forward(self, x):
        lr = lambda t: F.leaky_relu(t, negative_slope=SLOPE, inplace=True)
        up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
        x0 = lr(self.conv0(x))
        x1 = lr(self.conv1(x0))
        x2 = lr(self.conv2(x1))
        x3 = lr(self.conv3(x2))
        x4 = lr(self.conv4(up(x3)))
        if self.skip_connection:
            x4 = x4 + x2
        x5 = lr(self.conv5(up(x4)))
        if self.skip_connection:
            x5 = x5 + x1
        x6 = lr(self.conv6(up(x5)))
        if self.skip_connection:
            x6 = x6 + x0
        out = lr(self.conv7(x6))
        out = lr(self.conv8(out))
        return self.conv9(out)
