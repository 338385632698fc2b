"""HBM stream rates on this B200 with plain torch ops (CUDA events, best of 5): write-only, read-only, copy."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
def t(fn, nbytes, name):
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:12s} {nbytes / best / 1e6:8.0f} GB/s  ({best:.3f} ms)")
t(lambda: a.zero_(), 2 * n, "write-only")
t(lambda: a.fill_(1.5), 2 * n, "fill")
t(lambda: a.view(torch.int16).sum(dtype=torch.int64), 2 * n, "read-only")
t(lambda: b.copy_(a), 4 * n, "copy")
# mixed read/write ratios (round 2): how far below the copy rate a write-heavy stream sits
c = torch.empty(n, dtype=torch.bfloat16, device="cuda")
t(lambda: torch.add(a, b, out=c), 6 * n, "2r+1w add")
m = n // 8
t(lambda: c.view(8, m).copy_(a[:m].unsqueeze(0).expand(8, m)), 2 * n + 2 * m, "1r+8w bcast")
t(lambda: torch.add(a[:m].unsqueeze(0).expand(8, m), b.view(8, m), out=c.view(8, m)), 4 * n + 2 * m, "9r+8w")
f = torch.empty(n // 2, dtype=torch.float32, device="cuda")
t(lambda: f.zero_(), 2 * n, "write f32")
