"""world_size-2 gloo tests of the bucketed gradient reducer (host logic of the N>1 path, train_swin.py:152 /
train_hat.py:148): the averaged bucket gradients on both ranks equal the single-process gradient of the concatenated
batch — for one backward per step, for gradient accumulation (several backward() calls per optimizer step, reduced on
every micro-step as DDP does, or only on the last one under no_sync()), with a parameter that receives no gradient, and
in the non-overlapped mode.  The real async all-reduce handles are exercised (no monkey-patching)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 5)
        self.b = torch.nn.Linear(5, 3)
        self.unused = torch.nn.Linear(3, 3)   # never called: its bucket's countdown cannot complete on its own

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data():
    return torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10.0


def _loss(net, x):
    return net(x).pow(2).mean()


def _worker(rank, world, port, q, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from superresolution_def_b200.dp import BucketedGradReducer
    torch.manual_seed(0)
    net = _Tiny()
    x = _data()
    groups = [list(net.b.parameters()) + list(net.unused.parameters()), list(net.a.parameters())]
    red = BucketedGradReducer(groups, world, overlap=(mode != "serial"))
    micro = 2  # micro-batches per optimizer step; rank r owns rows [4r, 4r+4), micro-step m rows [4r+2m, 4r+2m+2)
    for _ in range(2):  # two optimizer steps: zero_grad / finish must re-arm the buckets
        red.zero_grad()
        for m in range(micro):
            xs = x[rank * 4 + 2 * m: rank * 4 + 2 * m + 2]
            if mode == "no_sync" and m + 1 < micro:
                with red.no_sync():
                    (_loss(net, xs) / micro).backward()
                continue
            (_loss(net, xs) / micro).backward()
            if mode == "serial":
                if m + 1 == micro:
                    red.reduce_all()
            else:
                red.finish()
    q.put((rank, red.launched, [p.grad.numpy().copy() for p in net.parameters()]))   # by value: no shared-memory handles that die with the worker
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["every_backward", "no_sync", "serial"])
def test_bucketed_reducer_world2_matches_single_process(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    res = {r: g for r, _, g in got}
    launched = {r: n for r, n, _ in got}
    torch.manual_seed(0)
    net = _Tiny()
    x = _data()
    # mean over ranks of (mean over micro-steps of per-micro-batch mean losses) == mean loss over the 4 equal micro-batches
    sum(_loss(net, x[2 * i: 2 * i + 2]) / 4 for i in range(4)).backward()
    for r in (0, 1):
        for g, (name, p) in zip(res[r], net.named_parameters()):
            g = torch.from_numpy(g)
            ref = p.grad if p.grad is not None else torch.zeros_like(p)
            assert torch.allclose(g, ref, atol=1e-6), (mode, r, name, (g - ref).abs().max())
    # 2 buckets x 2 steps x (2 reducing backwards | 1): every bucket is exchanged exactly once per reducing backward,
    # including the one whose countdown cannot complete (unused parameters)
    assert launched[0] == launched[1] == (8 if mode == "every_backward" else 4), launched
