"""GAN side of `train_swin.py` (BASELINE configs[3]): the discriminator and the loss modules the script builds, plus its
micro-step, so that the SwinIR generator mirror can be exercised and timed in the exact training arrangement —
DDP(find_unused_parameters=True) generator, DDP discriminator, fp16 autocast + GradScaler, requires_grad toggling, RaGAN
losses, gradient accumulation, EMA.

`UNetDiscriminatorSN` keeps the reference's module tree (`models/discriminator_swin.py:43-84`: same
attribute names, registration order and spectral-norm hooks, so `state_dict()` round-trips with strict=True and a reference
checkpoint's `net_d` loads); its forward and backward run on libsrk (`disc_engine.py`: the 4x4 stride-2 convolutions and
transposed convolutions as tcgen05 GEMMs over a patch matrix / followed by a fold, the 3x3 layers on the generators'
kernels).  Spectral normalisation (one power iteration on the `weight_u` / `weight_v` buffers when training, sigma, W / sigma
and its backward) is `srk_spectral_norm` / `srk_spectral_norm_bwd`: same buffers, same update rule as the hook that
`spectral_norm()` installs (which only fires inside `Conv2d.forward`, never called here).  There is no CPU path: the ATen restatement lives in
`oracle/discriminator_oracle.py` and is test infrastructure only.

The loss modules are interface mirrors (losses are out of scope).  The VGG feature extractor of the
perceptual loss (`utils/losses_train_swin.py:6-43`) cannot download its ImageNet weights here: it is seeded random, which
leaves the arithmetic (and the cost) of the loss unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm


def _sn_weight(m):
    """W / sigma(W) of a spectral_norm-wrapped layer, computed by the reference's own forward pre-hook
    (torch.nn.utils.spectral_norm: one power iteration in training mode, none in eval mode; u is updated in place) —
    the hook normally fires inside `m.forward`, which the libsrk path never calls."""
    for hook in m._forward_pre_hooks.values():
        hook(m, None)
    return m.weight


class UNetConv2(nn.Module):
    """4x4 stride-2 spectral-norm conv + LeakyReLU(0.2) (discriminator_swin.py:6-19)."""

    def __init__(self, in_size, out_size, dropout=0.0):
        super().__init__()
        layers = [spectral_norm(nn.Conv2d(in_size, out_size, 4, 2, 1, bias=False)), nn.LeakyReLU(0.2, inplace=True)]
        if dropout > 0:
            layers.append(nn.Dropout(dropout))
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        raise RuntimeError("UNetConv2 is executed by UNetDiscriminatorSN.forward (disc_engine); it has no stand-alone path")


class UNetUpBlock(nn.Module):
    """4x4 stride-2 spectral-norm transposed conv + LeakyReLU(0.2), bilinear resize to the skip's size if needed, then
    channel concat with the skip (discriminator_swin.py:21-41)."""

    def __init__(self, in_size, out_size, dropout=0.0):
        super().__init__()
        layers = [spectral_norm(nn.ConvTranspose2d(in_size, out_size, 4, 2, 1, bias=False)), nn.LeakyReLU(0.2, inplace=True)]
        if dropout > 0:
            layers.append(nn.Dropout(dropout))
        self.model = nn.Sequential(*layers)

    def forward(self, x, skip_input):
        raise RuntimeError("UNetUpBlock is executed by UNetDiscriminatorSN.forward (disc_engine); it has no stand-alone path")


class UNetDiscriminatorSN(nn.Module):
    """U-Net discriminator with spectral normalisation (discriminator_swin.py:43-84): per-pixel real/fake logits at half
    the input resolution.  Forward / backward = one libsrk autograd node (disc_engine.UNetDiscriminatorFunction)."""

    def __init__(self, num_in_ch=1, num_feat=64, skip_connection=True):
        super().__init__()
        self.skip_connection = skip_connection
        nf = num_feat
        self.conv0 = nn.Sequential(spectral_norm(nn.Conv2d(num_in_ch, nf, 3, 1, 1, bias=False)), nn.LeakyReLU(0.2, inplace=True),
                                   spectral_norm(nn.Conv2d(nf, nf, 4, 2, 1, bias=False)), nn.LeakyReLU(0.2, inplace=True))
        self.conv1 = UNetConv2(nf, nf * 2)
        self.conv2 = UNetConv2(nf * 2, nf * 4)
        self.conv3 = UNetConv2(nf * 4, nf * 8)
        self.conv4 = UNetConv2(nf * 8, nf * 8)
        self.up1 = UNetUpBlock(nf * 8, nf * 8)
        self.up2 = UNetUpBlock(nf * 16, nf * 4)
        self.up3 = UNetUpBlock(nf * 8, nf * 2)
        self.up4 = UNetUpBlock(nf * 4, nf)
        self.final_conv = nn.Sequential(spectral_norm(nn.Conv2d(nf * 2, nf, 3, 1, 1, bias=False)), nn.LeakyReLU(0.2, inplace=True),
                                        spectral_norm(nn.Conv2d(nf, 1, 3, 1, 1, bias=False)))

    def _convs(self):
        """the 12 spectrally normalised layers in forward order (discriminator_swin.py:72-84)"""
        return [self.conv0[0], self.conv0[2], self.conv1.model[0], self.conv2.model[0], self.conv3.model[0],
                self.conv4.model[0], self.up1.model[0], self.up2.model[0], self.up3.model[0], self.up4.model[0],
                self.final_conv[0], self.final_conv[2]]

    def forward(self, x):
        from . import disc_engine
        convs = self._convs()
        if len({m.training for m in convs}) != 1:
            raise RuntimeError("UNetDiscriminatorSN: mixed train / eval modes across the spectral-norm layers are not supported")
        eps = next(iter(convs[0]._forward_pre_hooks.values())).eps
        return disc_engine.unet_discriminator_sn(x, [m.weight_orig for m in convs], [m.weight_u for m in convs],
                                                 [m.weight_v for m in convs], convs[0].training, eps)


class RelativeGANLoss(nn.Module):
    """Relativistic average GAN loss on logits (gan_losses_swin.py:28-42)."""

    def __init__(self):
        super().__init__()
        self.loss = nn.BCEWithLogitsLoss()

    def forward(self, real_pred, fake_pred, for_discriminator=True):
        r, f = real_pred - fake_pred.mean(), fake_pred - real_pred.mean()
        if for_discriminator:
            return (self.loss(r, torch.ones_like(r)) + self.loss(f, torch.zeros_like(f))) / 2
        return (self.loss(f, torch.ones_like(f)) + self.loss(r, torch.zeros_like(r))) / 2


class VGGLoss(nn.Module):
    """L1 between VGG-19 features[:36] of the (grey -> 3 channel, ImageNet-normalised) images
    (losses_train_swin.py:6-43).  `seed` replaces the ImageNet download: same layers, same cost, random filters."""

    def __init__(self, feature_layer=35, seed=1234):
        super().__init__()
        import torchvision.models as models
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(seed)
            vgg19 = models.vgg19(weights=None)
        self.features = nn.Sequential(*list(vgg19.features.children())[:feature_layer + 1])
        for p in self.features.parameters():
            p.requires_grad = False
        self.register_buffer("mean", torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def forward(self, x, y):
        if x.shape[1] == 1:
            x = x.repeat(1, 3, 1, 1)
        if y.shape[1] == 1:
            y = y.repeat(1, 3, 1, 1)
        x, y = (x - self.mean) / self.std, (y - self.mean) / self.std
        return F.l1_loss(self.features(x), self.features(y).detach())


class CombinedGANLoss(nn.Module):
    """pixel L1 + perceptual + adversarial (RaGAN) with the script's weights (gan_losses_swin.py:77-118,
    train_swin.py:166)."""

    def __init__(self, pixel_weight=1.0, perceptual_weight=0.5, adversarial_weight=0.005):
        super().__init__()
        self.gan_loss = RelativeGANLoss()
        self.pixel_loss = nn.L1Loss()
        self.perceptual_loss = VGGLoss()
        self.pixel_weight, self.perceptual_weight, self.adversarial_weight = pixel_weight, perceptual_weight, adversarial_weight

    def forward(self, pred, target, real_pred=None, fake_pred=None):
        losses = {"pixel": self.pixel_loss(pred, target) * self.pixel_weight,
                  "perceptual": self.perceptual_loss(pred, target) * self.perceptual_weight}
        if fake_pred is not None and real_pred is not None:
            losses["adversarial"] = self.gan_loss(real_pred, fake_pred, for_discriminator=False) * self.adversarial_weight
        losses["total"] = sum(losses.values())
        return losses["total"], losses


class DiscriminatorLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.gan_loss = RelativeGANLoss()

    def forward(self, real_pred, fake_pred):
        total = self.gan_loss(real_pred, fake_pred, for_discriminator=True)
        return total, {"adversarial": total, "total": total}


class ModelEMA:
    """Exponential moving average over named_parameters (train_swin.py:45-74)."""

    def __init__(self, model, decay=0.999):
        self.model, self.decay = model, decay
        self.shadow = {n: p.data.clone() for n, p in model.named_parameters() if p.requires_grad}

    def update(self):
        for n, p in self.model.named_parameters():
            if p.requires_grad:
                self.shadow[n].mul_(self.decay).add_(p.data, alpha=1.0 - self.decay)


class GanTrainer:
    """The loop body of train_swin.py:214-259 for one rank: D micro-step (generator frozen, forward under no_grad),
    G micro-step (discriminator frozen), fp16 autocast + one shared GradScaler, optimizer steps every `accum` micro-steps,
    EMA after each generator step.  `net_g` / `net_d` may be DDP-wrapped exactly as the script wraps them."""

    def __init__(self, net_g, net_d, accum=4, lr_g=1e-4, lr_d=1e-4, ema_decay=0.999):
        self.net_g, self.net_d, self.accum = net_g, net_d, accum
        self.opt_g = torch.optim.AdamW(net_g.parameters(), lr=lr_g, weight_decay=0, betas=(0.9, 0.99))
        self.opt_d = torch.optim.AdamW(net_d.parameters(), lr=lr_d, weight_decay=0, betas=(0.9, 0.99))
        dev = next(net_g.parameters()).device
        self.criterion_g = CombinedGANLoss().to(dev)
        self.criterion_d = DiscriminatorLoss().to(dev)
        self.scaler = torch.amp.GradScaler("cuda")
        self.ema = ModelEMA(net_g.module if hasattr(net_g, "module") else net_g, ema_decay)
        self.i = 0
        self.opt_g.zero_grad()
        self.opt_d.zero_grad()

    def micro_step(self, lr_img, hr_img):
        net_g, net_d, scaler = self.net_g, self.net_d, self.scaler
        last = (self.i + 1) % self.accum == 0
        for p in net_d.parameters():
            p.requires_grad = True
        for p in net_g.parameters():
            p.requires_grad = False
        with torch.autocast("cuda"):
            with torch.no_grad():
                sr = net_g(lr_img)
            d_real = net_d(hr_img)
            d_fake = net_d(sr.detach())
            loss_d, _ = self.criterion_d(d_real, d_fake)
            loss_d = loss_d / self.accum
        scaler.scale(loss_d).backward()
        if last:
            scaler.step(self.opt_d)
            self.opt_d.zero_grad()
        for p in net_d.parameters():
            p.requires_grad = False
        for p in net_g.parameters():
            p.requires_grad = True
        with torch.autocast("cuda"):
            sr_g = net_g(lr_img)
            d_fake_g = net_d(sr_g)
            d_real_g = net_d(hr_img).detach()
            loss_g_total, _ = self.criterion_g(sr_g, hr_img, d_real_g, d_fake_g)
            loss_g = loss_g_total / self.accum
        scaler.scale(loss_g).backward()
        if last:
            scaler.step(self.opt_g)
            scaler.update()
            self.opt_g.zero_grad()
            self.ema.update()
        self.i += 1
        return loss_g_total.detach(), loss_d.detach() * self.accum
