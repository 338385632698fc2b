"""Times the fc1 / fc2-dgrad kernel variants at the bench shapes: GELU2 vs GELU1 (forward), MUL vs MULG (backward)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200 import _capi as capi

T, C, HP = 16 * 16384, 192, 768
bf, dev = torch.bfloat16, "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g).to(bf)
xn2, gout = rnd(T, C), rnd(T, C)
w1 = (torch.randn(HP, C, device=dev, generator=g) / 14).to(bf)
w2t = (torch.randn(HP, C, device=dev, generator=g) / 28).to(bf)
act, du, du2 = (torch.empty(T, HP, device=dev, dtype=bf) for _ in range(3))
dact = torch.empty(T, HP, device=dev, dtype=torch.float16)   # gelu' is stored as fp16
ln = capi.make_ln_args(HP, 720, None)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print("GELU2 us", t(lambda: capi.gemm_tn(capi.EPI_GELU2, xn2, w1, act, C2=dact, ln=ln)))
print("GELU1 us", t(lambda: capi.gemm_tn(capi.EPI_GELU1, xn2, w1, act, ln=ln)))
print("MUL   us", t(lambda: capi.gemm_tn(capi.EPI_MUL, gout, w2t, du, X1=dact)))
print("MULG  us", t(lambda: capi.gemm_tn(capi.EPI_MULG, gout, w2t, du2, X1=xn2, X2=w1, ln=ln)))
torch.cuda.synchronize()
print("max |MUL - MULG|", (du.float() - du2.float()).abs().max().item(), "rel", ((du.float() - du2.float()).norm() / du.float().norm()).item())
