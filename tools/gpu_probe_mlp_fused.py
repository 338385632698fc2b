"""Fused MLP forward (srk_mlp_fwd) against the two-kernel path (gemm_tn<GELU2> + gemm_tn<RES_LN>) on the same random
operands: outputs compared element-wise, then both timed at the bench shape (batch 16: T = 262144, Hp = 768).
Usage: python tools/gpu_probe_mlp_fused.py [T ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200 import _capi as capi

dev, bf = "cuda", torch.bfloat16


def run(T, Hp=768, C=180, hidden=720, time_it=False, drop=False):
    g = torch.Generator(device=dev).manual_seed(T + Hp)
    Cp = 192
    xn2 = torch.zeros(T, Cp, device=dev); xn2[:, :C] = torch.randn(T, C, device=dev, generator=g); xn2[:, C] = 1.0
    res = torch.zeros(T, Cp, device=dev); res[:, :C] = torch.randn(T, C, device=dev, generator=g)
    xn2, res = xn2.to(bf), res.to(bf)
    w1 = torch.zeros(Hp, Cp, device=dev); w1[:hidden, :C + 1] = torch.randn(hidden, C + 1, device=dev, generator=g) / C ** 0.5
    w2 = torch.zeros(Cp, Hp, device=dev); w2[:C, :hidden + 1] = torch.randn(C, hidden + 1, device=dev, generator=g) / hidden ** 0.5
    w1, w2 = w1.to(bf), w2.to(bf)
    gam, bet = 1 + 0.3 * torch.randn(C, device=dev, generator=g), 0.2 * torch.randn(C, device=dev, generator=g)
    rs = (torch.rand(max(1, T // 256), device=dev, generator=g) > 0.3).float() / 0.7 if drop else None
    e = lambda w: torch.full((T, w), float("nan"), device=dev, dtype=bf)  # noqa: E731
    eh = lambda w: torch.full((T, w), float("nan"), device=dev, dtype=torch.float16)  # noqa: E731  (gelu' is stored as fp16)
    ref = dict(act=e(Hp), dact=eh(Hp), x=e(Cp), xn=e(Cp), st=torch.empty(T, 2, device=dev))
    got = dict(act=e(Hp), dact=eh(Hp), x=e(Cp), xn=e(Cp), st=torch.empty(T, 2, device=dev))

    def unfused(o):
        capi.gemm_tn(capi.EPI_GELU2, xn2, w1, o["act"], C2=o["dact"], ln=capi.make_ln_args(Hp, hidden, None))
        capi.gemm_tn(capi.EPI_RES_LN, o["act"], w2, o["x"], C2=o["xn"], X1=res,
                     ln=capi.make_ln_args(C, C, gam, bet, stats=o["st"], row_scale=rs, rows_per_scale=256))

    def fused(o, store=True):
        capi.mlp_fwd(xn2, w1, w2, res, o["act"] if store else None, o["dact"] if store else None, o["x"], o["xn"], hidden,
                     capi.make_ln_args(C, C, gam, bet, stats=o["st"], row_scale=rs, rows_per_scale=256))

    unfused(ref); fused(got)
    torch.cuda.synchronize()
    ok = True
    for k in ("act", "dact", "x", "xn", "st"):
        a, b = got[k].float(), ref[k].float()
        bad = ~torch.isfinite(a)
        err = (a - b).abs().max().item() if not bad.any() else float("nan")
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        exact = torch.equal(got[k], ref[k])
        print(f"T={T} Hp={Hp} drop={drop} {k:5s} max-abs {err:.3e} rel-L2 {rel:.3e} bit-exact {exact} nonfinite {int(bad.sum())}")
        ok &= (not bad.any()) and rel < 2e-3
    inf = dict(x=e(Cp), xn=e(Cp), st=torch.empty(T, 2, device=dev))
    fused(inf, store=False)
    torch.cuda.synchronize()
    print("   inference (no act store): x bit-exact", torch.equal(inf["x"], got["x"]), "xn bit-exact", torch.equal(inf["xn"], got["xn"]))
    ok &= torch.equal(inf["x"], got["x"])
    if time_it:
        for name, fn in (("unfused", lambda: unfused(ref)), ("fused", lambda: fused(got)), ("fused-infer", lambda: fused(inf, False))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record(); e1.synchronize()
            us = e0.elapsed_time(e1) * 100
            nbytes = T * 2 * ((4 * Cp + 2 * Hp) if name != "fused-infer" else 4 * Cp) + (T * Hp * 2 * 2 if name == "unfused" else 0) * 0
            print(f"   {name:12s} {us:8.1f} us/launch-pair   algorithmic {nbytes / 1e6:.0f} MB -> {nbytes / us / 1e3:.0f} GB/s")
    return ok


if __name__ == "__main__":
    print(capi.version(), torch.cuda.get_device_name(0), flush=True)
    sizes = [int(a) for a in sys.argv[1:]] or [128, 512, 148 * 128 + 384, 262144]
    allok = True
    for T in sizes:
        allok &= run(T, time_it=(T >= 100000))
    allok &= run(1024, Hp=256, C=90, hidden=180)
    allok &= run(2048, Hp=512, C=180, hidden=360, drop=True)
    print("ALL OK" if allok else "MISMATCH")
    sys.exit(0 if allok else 1)
