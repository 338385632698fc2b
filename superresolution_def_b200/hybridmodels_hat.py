"""Drop-in mirror of the reference's `models/hybridmodels_hat.py` module interface on top of libsrk.

Same class names, constructor arguments, parameter names / shapes / registration order and initialisation as the
reference (`load_state_dict(strict=True)` round-trips; infer_hat.py's checkpoint sniffing of `hat.conv_first.weight`,
`conv_adapt.weight`, `rrdb_trunk.N.rdb1.conv1.weight`, `hat.layers.N.*` sees the same keys), but the forwards run the
sm_100a kernels: HAT through hat_arch (window-8 attention cores, OCAB 12x12, CAB), and everything after it through
hybrid_engine (implicit-GEMM convolutions over channel slices — the dense blocks' torch.cat is never materialised).

Reference: models/hybridmodels_hat.py:21-131.  Compute dtype bf16 with fp32 accumulation; no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _capi as capi
from . import hybrid_engine as hyb
from . import swin_engine as eng
from .hat_arch import HAT

RDB_KEYS = tuple(f"conv{i}.{s}" for i in range(1, 6) for s in ("weight", "bias"))


def _rdb_params(rdb) -> list[torch.Tensor]:
    sd = dict(rdb.named_parameters())
    return [sd[k] for k in RDB_KEYS]


def _kaiming_zero_bias(convs):
    """The reference's initialisation of its plain convolutions (:31-36, :109-115): He-normal weights, zero bias."""
    for m in convs:
        nn.init.kaiming_normal_(m.weight, a=0, mode="fan_in")
        if m.bias is not None:
            m.bias.data.zero_()


class ResidualDenseBlock(nn.Module):
    """conv_k sees the block input and the k-1 earlier growth outputs; parameters conv1..conv5 as in the reference."""

    def __init__(self, num_feat=64, num_grow_ch=32):
        super().__init__()
        self.num_feat, self.num_grow_ch = num_feat, num_grow_ch
        for k in range(1, 6):   # registration order conv1 .. conv5 (state_dict / RNG order of the reference)
            cout = num_grow_ch if k < 5 else num_feat
            setattr(self, f"conv{k}", nn.Conv2d(num_feat + (k - 1) * num_grow_ch, cout, 3, 1, 1))
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        _kaiming_zero_bias(getattr(self, f"conv{k}") for k in range(1, 6))

    def forward(self, x):
        """x: (B, num_feat, H, W) as in the reference (:38-44)."""
        return hyb.DenseTrunkFunction.apply(x, self.num_feat, self.num_grow_ch, False, *_rdb_params(self))


class RRDBBlock(nn.Module):
    def __init__(self, num_feat=64, num_grow_ch=32):
        super().__init__()
        self.rdb1 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb2 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb3 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.num_feat, self.num_grow_ch = num_feat, num_grow_ch

    def rdb_params(self) -> list[torch.Tensor]:
        return _rdb_params(self.rdb1) + _rdb_params(self.rdb2) + _rdb_params(self.rdb3)

    def forward(self, x):
        """x: (B, num_feat, H, W) (:54-58)."""
        return hyb.DenseTrunkFunction.apply(x, self.num_feat, self.num_grow_ch, True, *self.rdb_params())


class HybridHATRealESRGAN(nn.Module):
    def __init__(self, img_size=128, in_chans=1, embed_dim=180, depths=(6, 6, 6, 6, 6, 6), num_heads=(6, 6, 6, 6, 6, 6),
                 window_size=8, upscale=4, num_rrdb=23, num_feat=64, num_grow_ch=32):
        super().__init__()
        eng.track_weight_changes(self)
        self.upscale = upscale
        self.img_size = img_size
        if in_chans != 1 or upscale != 4:
            raise capi.SrkError("libsrk HybridHATRealESRGAN: in_chans=1, upscale=4 (HAT x2 + nearest x2; the reference "
                                "hard-codes both stages, hybridmodels_hat.py:87,127)")
        self.hat = HAT(img_size=img_size, in_chans=in_chans, embed_dim=embed_dim, depths=depths, num_heads=num_heads,
                       window_size=window_size, upscale=2, upsampler='pixelshuffle', img_range=1.0,
                       resi_connection='1conv')
        self.conv_adapt = nn.Conv2d(in_chans, num_feat, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        self.rrdb_trunk = nn.Sequential(*(RRDBBlock(num_feat, num_grow_ch) for _ in range(num_rrdb)))
        for name in ("conv_body", "conv_up", "conv_hr"):
            setattr(self, name, nn.Conv2d(num_feat, num_feat, 3, 1, 1))
        self.conv_last = nn.Conv2d(num_feat, in_chans, 3, 1, 1)
        self.num_feat, self.num_grow_ch = num_feat, num_grow_ch
        _kaiming_zero_bias((self.conv_adapt, self.conv_body, self.conv_up, self.conv_hr, self.conv_last))

    def tail_params(self) -> list[torch.Tensor]:
        ps = [self.conv_adapt.weight, self.conv_adapt.bias]
        for blk in self.rrdb_trunk:
            ps += blk.rdb_params()
        for m in (self.conv_body, self.conv_up, self.conv_hr, self.conv_last):
            ps += [m.weight, m.bias]
        return ps

    def forward(self, x):
        hat_out = self.hat(x)
        out = hyb.HybridTailFunction.apply(hat_out, self.num_feat, self.num_grow_ch, *self.tail_params())
        return out.to(hat_out.dtype)

    def load_pretrained_hat(self, hat_path):
        """Same contract as the reference helper (:133-143): load a HAT checkpoint non-strictly, never raise."""
        try:
            state = torch.load(hat_path, map_location="cpu")
            state = state.get("model_state_dict", state)
            self.hat.load_state_dict({k.replace("module.", ""): v for k, v in state.items()}, strict=False)
        except Exception as err:  # the reference reports and carries on with random HAT weights
            print(f"load_pretrained_hat: could not load {hat_path}: {err}")
            return
        print(f"load_pretrained_hat: HAT weights loaded from {hat_path}")
