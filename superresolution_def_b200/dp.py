"""Data-parallel gradient exchange for one process per GPU (NCCL over NVLink 5 / NVSwitch).

Replaces the reference's `DDP(net_g)` wrap (train_swin.py:152, train_hat.py:148): parameters' gradients live in
flat fp32 buckets laid out in *reverse execution order* (tail convs first, then layers 5..0, then conv_first), and a
bucket's all-reduce (mean) is launched asynchronously the moment its last gradient has been accumulated, so the
exchange of layer group k overlaps the backward kernels of group k-1.  Unlike the reference launchers
(start_swin.py:131-135) NCCL P2P/NVLS stay enabled.  Patches are independent, so there is no other collective.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class BucketedGradReducer:
    def __init__(self, groups: list[list[torch.nn.Parameter]], world_size: int, overlap: bool = True):
        """groups: parameter groups in the order their gradients become ready during backward.
        overlap=True: each bucket's all-reduce is launched from a post-accumulate hook while backward continues.
        overlap=False: no hooks; call reduce_all() after backward (used when forward+backward is replayed as a CUDA
        graph, so that no NCCL call is ever issued inside a stream capture)."""
        self.world = world_size
        self.buckets = []
        self.handles = []
        self._pending = []
        for params in groups:
            params = [p for p in params if p.requires_grad]
            if not params:
                continue
            n = sum(p.numel() for p in params)
            flat = torch.zeros(n, device=params[0].device, dtype=torch.float32)
            off = 0
            for p in params:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            bi = len(self.buckets)
            self.buckets.append(flat)
            self._pending.append(len(params))
            if world_size > 1 and overlap:
                for p in params:
                    p.register_post_accumulate_grad_hook(self._make_hook(bi))
        self._count = list(self._pending)

    def _make_hook(self, bi: int):
        def hook(_p):
            self._count[bi] -= 1
            if self._count[bi] == 0:
                self.handles.append(dist.all_reduce(self.buckets[bi], op=dist.ReduceOp.AVG, async_op=True))
        return hook

    def zero_grad(self):
        for b in self.buckets:
            b.zero_()
        self._count = list(self._pending)

    def reduce_all(self):
        """Average every bucket across ranks on the current stream (non-overlapped mode)."""
        if self.world > 1:
            for b in self.buckets:
                dist.all_reduce(b, op=dist.ReduceOp.AVG)

    def finish(self):
        """Block the current stream on every outstanding bucket exchange (call after backward())."""
        for h in self.handles:
            h.wait()
        self.handles.clear()

    @property
    def nbytes(self) -> int:
        return sum(b.numel() * 4 for b in self.buckets)


def swinir_grad_groups(net) -> list[list[torch.nn.Parameter]]:
    """Reverse-execution-order parameter groups of a SwinIR- or HAT-shaped generator: tail convs + final norm, then
    the residual groups last to first, then the head."""
    if hasattr(net, "rrdb_trunk") and hasattr(net, "hat"):
        # HybridHATRealESRGAN: tail convs, the RRDB trunk last to first (+ conv_adapt), then the HAT stage
        groups = [[p for m in (net.conv_last, net.conv_hr, net.conv_up, net.conv_body) for p in m.parameters()]]
        groups += [list(blk.parameters()) for blk in reversed(list(net.rrdb_trunk))]
        groups.append(list(net.conv_adapt.parameters()))
        return groups + swinir_grad_groups(net.hat)
    groups = [list(net.conv_last.parameters()) + list(net.upsample.parameters())
              + list(net.conv_before_upsample.parameters()) + list(net.conv_after_body.parameters())
              + list(net.norm.parameters())]
    for layer in reversed(list(net.layers)):
        groups.append(list(layer.parameters()))
    head = list(net.conv_first.parameters())
    pe = getattr(net, "patch_embed", None)
    if pe is not None:
        head += list(pe.parameters())
    groups.append(head)
    seen = {id(p) for g in groups for p in g}
    rest = [p for p in net.parameters() if id(p) not in seen]
    if rest:
        groups.append(rest)
    return groups
