"""Data-parallel gradient exchange for one process per GPU (NCCL over NVLink 5 / NVSwitch).

Replaces the reference's `DDP(net_g)` wrap (train_swin.py:152, train_hat.py:148): parameters' gradients live in
flat fp32 buckets laid out in *reverse execution order* (tail convs first, then layers 5..0, then conv_first), and a
bucket's all-reduce (mean) is launched asynchronously the moment its last gradient has been accumulated, so the
exchange of layer group k overlaps the backward kernels of group k-1.  Unlike the reference launchers
(start_swin.py:131-135) NCCL P2P/NVLS stay enabled.  Patches are independent, so there is no other collective.

Semantics match DDP's, including under gradient accumulation (both reference loops run several backward() calls per
optimizer step: train_swin.py ACCUM_STEPS, train_hat.py GRADIENT_ACCUMULATION):

* every backward() reduces every bucket exactly once — a bucket re-arms itself the moment it fires, and finish()
  reduces the buckets that did not fire because some of their parameters received no gradient in that backward
  (DDP's find_unused_parameters=True behaviour).  Because the mean is linear and an already-averaged bucket is identical
  on all ranks, averaging `avg(g_1) + g_2_local` yields `avg(g_1) + avg(g_2)`: accumulation stays exact;
* `with reducer.no_sync():` skips the exchange for the micro-steps inside it (DDP.no_sync), so that only the last
  micro-step of an accumulation window pays for communication.
"""
from __future__ import annotations

import contextlib

import torch
import torch.distributed as dist


class _DividedWork:
    """SUM all-reduce handle that turns into a mean when waited on (backends without ReduceOp.AVG, i.e. gloo)."""

    def __init__(self, work, bucket, world):
        self.work, self.bucket, self.world = work, bucket, world

    def wait(self):
        self.work.wait()
        self.bucket.div_(self.world)


class BucketedGradReducer:
    def __init__(self, groups: list[list[torch.nn.Parameter]], world_size: int, overlap: bool = True, group=None):
        """groups: parameter groups in the order their gradients become ready during backward.
        overlap=True: each bucket's all-reduce is launched from a post-accumulate hook while backward continues; call
        finish() after every backward().  overlap=False: no hooks; call reduce_all() after backward()."""
        self.world = world_size
        self.group = group
        self.overlap = overlap
        self.buckets: list[torch.Tensor] = []
        self.handles: list = []
        self._pending: list[int] = []
        self._sync = True
        self.launched = 0      # all-reduces issued since construction (tests / bench bookkeeping)
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
            if world_size > 1:   # hooks are always installed; they are inert while self.overlap is False
                for p in params:
                    p.register_post_accumulate_grad_hook(self._make_hook(bi))
        self._count = list(self._pending)
        self._fired = [False] * len(self.buckets)
        self._avg = None

    # ------------------------------------------------------------------ exchange
    def _has_avg(self) -> bool:
        if self._avg is None:
            self._avg = dist.get_backend(self.group) == "nccl"
        return self._avg

    def _launch(self, bi: int, async_op: bool):
        b = self.buckets[bi]
        self.launched += 1
        if self._has_avg():
            return dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        w = dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            return _DividedWork(w, b, self.world)
        b.div_(self.world)
        return None

    def _make_hook(self, bi: int):
        def hook(_p):
            if not self.overlap:
                return
            self._count[bi] -= 1
            if self._count[bi] == 0:
                self._count[bi] = self._pending[bi]   # re-arm for the next backward() (gradient accumulation)
                if self._sync:
                    self._fired[bi] = True
                    self.handles.append(self._launch(bi, async_op=True))
        return hook

    @contextlib.contextmanager
    def no_sync(self):
        """Backward passes inside this context accumulate local gradients without exchanging them (DDP.no_sync)."""
        prev, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = prev
            self._count = list(self._pending)

    def zero_grad(self):
        for b in self.buckets:
            b.zero_()
        self._count = list(self._pending)
        self._fired = [False] * len(self.buckets)

    def reduce_all(self):
        """Average every bucket across ranks on the current stream (non-overlapped mode)."""
        if self.world > 1:
            for bi in range(len(self.buckets)):
                self._launch(bi, async_op=False)

    def finish(self):
        """Call after every backward(): reduces the buckets whose countdown did not complete (parameters without a
        gradient in this backward), then blocks the current stream on every outstanding exchange.  On return every
        bucket has been averaged exactly once for this backward."""
        if self.world > 1 and self.overlap and self._sync:
            for bi, fired in enumerate(self._fired):
                if not fired:
                    self.handles.append(self._launch(bi, async_op=True))
        for h in self.handles:
            h.wait()
        self.handles.clear()
        self._count = list(self._pending)
        self._fired = [False] * len(self.buckets)

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
