import math
import torch


def rel_l2(got: torch.Tensor, ref: torch.Tensor) -> float:
    got = got.detach().float(); ref = ref.detach().float()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


def max_abs(got: torch.Tensor, ref: torch.Tensor) -> float:
    return (got.detach().float() - ref.detach().float()).abs().max().item()


def randomize_(module: torch.nn.Module, seed: int = 0, table_std: float = 1.0):
    """Make bias / LayerNorm / rpb-table bugs visible: random LN affine, sigma=1 bias tables, non-zero biases."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("relative_position_bias_table"):
                p.copy_(torch.randn(p.shape, generator=g) * table_std)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=g))
            elif name.endswith("attention.1.bias"):
                # squeeze layer of HAT's ChannelAttention (6 ReLU units fed by a global mean): keep the pre-activations
                # away from the ReLU kink, where bf16-vs-fp32 rounding would flip a unit on/off and make the
                # comparison of that layer's gradients meaningless (both signs are exercised)
                sign = torch.where(torch.arange(p.numel()) % 2 == 0, 1.0, -1.0)
                p.copy_(sign * (0.5 + 0.1 * torch.rand(p.shape, generator=g)))
            elif name.endswith("bias"):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan_in))
    return module
