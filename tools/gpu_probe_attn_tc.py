"""Window-attention core (ws 8) at the bench shape: times srk_win_attn_fwd / srk_win_attn_bwd and, when a dump of the other
kernel family exists, compares against it.  Run once with SRK_ATTN_TC=0 (mma.sync kernels, writes the dump) and once with
the default (tcgen05 kernels, compares).  Usage: python tools/gpu_probe_attn_tc.py [batch] [shift] [mask]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200 import _capi as capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev, bf = "cuda", torch.bfloat16
H = W = 128
heads, dh, ds = 6, 30, 32
T = B * H * W
g = torch.Generator(device=dev).manual_seed(7)
qkv = torch.zeros(T, 3, heads, ds, device=dev)
qkv[..., :dh] = torch.randn(T, 3, heads, dh, device=dev, generator=g)
qkv = qkv.view(T, 3 * heads * ds).to(bf)
dout = torch.zeros(T, heads, ds, device=dev)
dout[..., :dh] = torch.randn(T, heads, dh, device=dev, generator=g)
dout = dout.view(T, heads * ds).to(bf)
table = torch.randn(225, heads, device=dev, generator=g)
out = torch.empty(T, heads * ds, device=dev, dtype=bf)
dqkv = torch.empty_like(qkv)
dtab = torch.empty(225, heads, device=dev)
geom = capi.SrkGeom(B, H, W, 8, shift)
tc = os.environ.get("SRK_ATTN_TC", "1") != "0"


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


f = lambda: capi.win_attn_fwd(geom, heads, qkv, table, out, ones_col=dh)  # noqa: E731
b = lambda: capi.win_attn_bwd(geom, heads, qkv, table, dout, dqkv, dtab)  # noqa: E731
tf, tb = timed(f), timed(b)
torch.cuda.synchronize()
nb_f, nb_b = T * (576 + 192) * 2, T * (2 * 576 + 192) * 2
print(f"{'tcgen05' if tc else 'mma.sync'} B={B} shift={shift}: fwd {tf:.1f} us ({nb_f / tf / 1e3:.0f} GB/s)  bwd {tb:.1f} us ({nb_b / tb / 1e3:.0f} GB/s)")
dump = f"/tmp/attn_dump_{B}_{shift}.pt"
if not tc:
    torch.save({"out": out.cpu(), "dqkv": dqkv.cpu(), "dtab": dtab.cpu()}, dump)
elif os.path.exists(dump):
    ref = torch.load(dump)
    for k, v in (("out", out), ("dqkv", dqkv), ("dtab", dtab)):
        a, r = v.float().cpu(), ref[k].float()
        print(f"   {k}: rel-L2 vs mma.sync kernels {((a - r).norm() / r.norm()).item():.3e} max-abs {(a - r).abs().max().item():.3e}")
