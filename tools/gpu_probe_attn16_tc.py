"""A/B of the window-16 attention forward: tcgen05 / TMEM / TMA kernel (attn_tc16.cuh, default) against the mma.sync kernel
(attn_win16.cuh, SRK_ATTN16_TC=0) — parity of out / lse on seeded inputs, then CUDA-event timings at the HAT bench shape.
Usage: python tools/gpu_probe_attn16_tc.py [--time] [--B 8] [--hw 128]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superresolution_def_b200 import _capi as capi  # noqa: E402

dev = "cuda"


def packed_qkv(T, heads, seed, scale=1.0):
    g = torch.Generator(device=dev).manual_seed(seed)
    qkv = torch.zeros(T, 3, heads, 32, device=dev)
    qkv[..., :30] = torch.randn(T, 3, heads, 30, device=dev, generator=g) * scale
    return qkv.reshape(T, 3 * heads * 32).to(torch.bfloat16)


def run(mode, B, H, W, shift, heads, qkv, table, tc):
    os.environ["SRK_ATTN16_TC"] = "1" if tc else "0"
    T = B * H * W
    out = torch.full((T, heads * 32), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.full((heads, T), float("nan"), device=dev)
    capi.win_attn16_fwd(capi.SrkGeom(B, H, W, 16, shift), mode, heads, qkv, table, out, lse, ones_col=30)
    torch.cuda.synchronize()
    return out, lse


def parity():
    ok = True
    for mode, name, tbl in ((capi.ATTN_SELF, "SELF", 961), (capi.ATTN_OCA, "OCA", 1521)):
        for shift in ((0, 8) if mode == capi.ATTN_SELF else (0,)):
            for (B, H, W) in ((1, 16, 16), (2, 32, 48), (1, 64, 64)):
                heads = 6
                qkv = packed_qkv(B * H * W, heads, 1 + shift, scale=1.5)
                table = torch.randn(tbl, heads, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
                o0, l0 = run(mode, B, H, W, shift, heads, qkv, table, tc=False)
                o1, l1 = run(mode, B, H, W, shift, heads, qkv, table, tc=True)
                eo = ((o1.float() - o0.float()).norm() / o0.float().norm()).item()
                el = (l1 - l0).abs().max().item()
                bad = not (eo < 6e-3 and el < 2e-3) or not torch.isfinite(o1.float()).all()
                ok &= not bad
                print(f"{name} shift {shift} B{B} {H}x{W}: out rel-L2 {eo:.2e}  max|dlse| {el:.2e}  max|dout| "
                      f"{(o1.float() - o0.float()).abs().max().item():.3e} {'FAIL' if bad else 'ok'}", flush=True)
    return ok


def timing(B, hw):
    heads = 6
    T = B * hw * hw
    for mode, name, tbl, shift in ((capi.ATTN_SELF, "SELF", 961, 8), (capi.ATTN_SELF, "SELF", 961, 0), (capi.ATTN_OCA, "OCA", 1521, 0)):
        qkvs = [packed_qkv(T, heads, s) for s in range(3)]   # 3 x 151 MB > L2
        table = torch.randn(tbl, heads, device=dev) * 0.5
        out = torch.empty(T, heads * 32, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(heads, T, device=dev)
        geom = capi.SrkGeom(B, hw, hw, 16, shift)
        for tc in (0, 1):
            os.environ["SRK_ATTN16_TC"] = str(tc)
            for _ in range(3):
                capi.win_attn16_fwd(geom, mode, heads, qkvs[0], table, out, lse, ones_col=30)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            n = 12
            for k in range(n):
                capi.win_attn16_fwd(geom, mode, heads, qkvs[k % 3], table, out, lse, ones_col=30)
            e1.record()
            torch.cuda.synchronize()
            print(f"{name} shift {shift} B{B} {hw}^2 {'tcgen05' if tc else 'mma.sync'}: {e0.elapsed_time(e1) / n * 1e3:8.1f} us", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--hw", type=int, default=128)
    args = ap.parse_args()
    good = parity()
    if args.time:
        timing(args.B, args.hw)
    sys.exit(0 if good else 1)
