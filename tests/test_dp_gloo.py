"""world_size-2 gloo test of the bucketed gradient reducer (host logic of the N>1 path): the averaged bucket
gradients on both ranks equal the single-process gradient of the concatenated batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 5)
        self.b = torch.nn.Linear(5, 3)

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from superresolution_def_b200.dp import BucketedGradReducer
    torch.manual_seed(0)
    net = _Tiny()
    x = torch.arange(4 * 6, dtype=torch.float32).reshape(4, 6) / 10.0
    red = BucketedGradReducer([list(net.b.parameters()), list(net.a.parameters())], world)
    # gloo has no AVG: emulate with SUM / world through the same hook path
    orig = dist.all_reduce

    def avg(t, op=None, async_op=False):
        h = orig(t, op=dist.ReduceOp.SUM, async_op=False)
        t /= world

        class _H:
            def wait(self):
                return None
        return _H()
    dist.all_reduce = avg
    for _ in range(2):  # two steps: zero_grad must re-arm the buckets
        red.zero_grad()
        net(x[rank * 2:(rank + 1) * 2]).pow(2).mean().backward()
        red.finish()
    q.put((rank, [p.grad.numpy().copy() for p in net.parameters()]))   # by value: no shared-memory handles that die with the worker
    dist.destroy_process_group()


def test_bucketed_reducer_world2_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    torch.manual_seed(0)
    net = _Tiny()
    x = torch.arange(4 * 6, dtype=torch.float32).reshape(4, 6) / 10.0
    # mean over ranks of per-rank mean losses == mean loss of the full batch (equal shard sizes)
    net(x).pow(2).mean().backward()
    for r in (0, 1):
        for g, p in zip(res[r], net.parameters()):
            g = torch.from_numpy(g)
            assert torch.allclose(g, p.grad, atol=1e-6), (r, (g - p.grad).abs().max())
