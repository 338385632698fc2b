"""Error of the fc1 epilogue outputs (gelu, gelu') against the exact functions of the fp32 accumulator, in fp64:
python tools/gpu_probe_gelu_err.py   (SRK_LIB=... selects another build of the library)"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200 import _capi as capi
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
dev, bf = "cuda", torch.bfloat16
M, N, K = 8192, 768, 192
for scale in (1.0, 3.0):
    A = (torch.randn(M, K, device=dev) * scale).to(bf)
    B = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf)
    C = torch.zeros(M, N, device=dev, dtype=bf); C2 = torch.zeros_like(C, dtype=torch.float16)
    capi.gemm_tn(capi.EPI_GELU2, A, B, C, C2=C2, ln=capi.make_ln_args(N, -1, None))
    torch.cuda.synchronize()
    u = (A.double() @ B.double().t())
    a = u * 0.5 * (1 + torch.erf(u / math.sqrt(2)))
    g = 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)
    rl = lambda x, y: ((x.double() - y).norm() / y.norm()).item()
    # the reference chain under autocast: u -> bf16, gelu in fp32 on the bf16 value, result -> bf16
    ub = u.float().to(bf).float()
    a_ref = torch.nn.functional.gelu(ub).to(bf)
    print(f"scale {scale}: act rel-L2 {rl(C, a):.3e} (autocast chain {rl(a_ref, a):.3e}; bf16 rounding alone {rl(a.float().to(bf), a):.3e})  "
          f"gelu' rel-L2 {rl(C2, g):.3e} (bf16 rounding alone {rl(g.float().to(bf), g):.3e})  max|act err| {(C.double() - a).abs().max().item():.3e}")
