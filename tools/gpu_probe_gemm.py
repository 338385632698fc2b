"""GPU probe for the tcgen05 GEMM kernels: prints error statistics per case (no asserts) so one
gpurun call tells us which descriptor/layout variants are right.  Usage: python tools/gpu_probe_gemm.py <case>"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200 import _capi as capi

torch.manual_seed(0)
dev = "cuda"
bf = torch.bfloat16


def stat(name, got, ref, tol=2e-2):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    mx = err.max().item()
    idx = err.argmax().item()
    r, c = divmod(idx, got.shape[1])
    bad = (err > tol * (ref.abs() + 1)).float().mean().item()
    print(f"[{name}] max_abs_err={mx:.4g} (ref max {denom:.4g}) at ({r},{c}) got={got[r,c].item():.4g} ref={ref[r,c].item():.4g} "
          f"frac_bad={bad:.4g} nan={torch.isnan(got).any().item()} -> {'OK' if bad == 0 and not torch.isnan(got).any() else 'FAIL'}", flush=True)
    return bad == 0


def case_store():
    for (M, N, K) in [(128, 64, 64), (256, 128, 128), (1024, 192, 192), (2048, 576, 192), (128 * 300, 192, 768), (4096, 768, 192), (512, 256, 64)]:
        A = torch.randn(M, K, device=dev).to(bf)
        B = torch.randn(N, K, device=dev).to(bf)
        C = torch.full((M, N), float("nan"), device=dev, dtype=bf)
        capi.gemm_tn(capi.EPI_STORE, A, B, C)
        torch.cuda.synchronize()
        ref = A.float() @ B.float().t()
        stat(f"store M{M} N{N} K{K}", C, ref)
    # strided views (ld != cols)
    M, N, K = 512, 192, 192
    Abig = torch.randn(M, 256, device=dev).to(bf); A = Abig[:, :K]
    Bbig = torch.randn(N, 320, device=dev).to(bf); B = Bbig[:, 64:64 + K]
    Cbig = torch.zeros(M, 576, device=dev, dtype=bf); C = Cbig[:, 192:384]
    capi.gemm_tn(capi.EPI_STORE, A, B, C)
    torch.cuda.synchronize()
    stat("store strided", C, A.float() @ B.float().t())
    print("untouched outside:", (Cbig[:, :192] == 0).all().item() and (Cbig[:, 384:] == 0).all().item())


def gelu_ref(u):
    a = torch.nn.functional.gelu(u)
    cdf = 0.5 * (1 + torch.erf(u / math.sqrt(2)))
    pdf = torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)
    return a, cdf + u * pdf


def case_gelu_mul():
    M, N, K = 1024, 768, 192
    A = torch.randn(M, K, device=dev).to(bf)
    B = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf)
    C = torch.zeros(M, N, device=dev, dtype=bf); C2 = torch.zeros_like(C, dtype=torch.float16)
    ln = capi.make_ln_args(N, 720, None)
    capi.gemm_tn(capi.EPI_GELU2, A, B, C, C2=C2, ln=ln)
    torch.cuda.synchronize()
    u = (A.float() @ B.float().t()).to(bf).float()
    a, g = gelu_ref(u)
    a[:, 720] = 1; g[:, 720] = 0
    stat("gelu2.a", C, a); stat("gelu2.g", C2, g)
    X1 = torch.randn(M, N, device=dev).to(torch.float16)
    capi.gemm_tn(capi.EPI_MUL, A, B, C, X1=X1)
    torch.cuda.synchronize()
    stat("mul", C, u * X1.float())


def case_ln():
    for K in (192, 768):
        M, N, n = 128 * 160, 192, 180
        A = torch.randn(M, K, device=dev).to(bf)
        B = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(bf); B[n:] = 0
        R = torch.randn(M, N, device=dev).to(bf); R[:, n:] = 0
        gamma = torch.randn(n, device=dev); beta = torch.randn(n, device=dev)
        stats = torch.zeros(M, 2, device=dev)
        C = torch.zeros(M, N, device=dev, dtype=bf); C2 = torch.zeros_like(C)
        ln = capi.make_ln_args(n, 180, gamma, beta, stats=stats)
        capi.gemm_tn(capi.EPI_RES_LN, A, B, C, C2=C2, X1=R, ln=ln)
        torch.cuda.synchronize()
        v = ((A.float() @ B.float().t()).to(bf) + R).float()  # bf16 + bf16 -> bf16
        v = v.to(bf).float()
        stat(f"res_ln.v K{K}", C, v)
        vv = C.float()[:, :n]  # use the kernel's own v for the LN check (isolates LN math)
        mean = vv.mean(1, keepdim=True); var = vv.var(1, unbiased=False, keepdim=True)
        xn = torch.zeros(M, N, device=dev)
        xn[:, :n] = (vv - mean) / torch.sqrt(var + 1e-5) * gamma + beta
        xn[:, 180] = 1
        stat(f"res_ln.xn K{K}", C2, xn, tol=3e-2)
        stat(f"res_ln.mean K{K}", stats[:, :1], mean, tol=1e-3); stat(f"res_ln.rstd K{K}", stats[:, 1:], 1 / torch.sqrt(var + 1e-5), tol=1e-3)
        # LNBWD: acc = dxn = A2 @ B^T ; X1 = x (= C above), X2 = dres
        dres = torch.randn(M, N, device=dev).to(bf); dres[:, n:] = 0
        grid = capi.gemm_grid(M, N)
        partials = torch.zeros(grid, 2, N, device=dev)
        out = torch.zeros(M, N, device=dev, dtype=bf)
        ln2 = capi.make_ln_args(n, -1, gamma, None, stats=stats, partials=partials)
        capi.gemm_tn(capi.EPI_LNBWD, A, B, out, X1=C, X2=dres, ln=ln2)
        torch.cuda.synchronize()
        dxn = (A.float() @ B.float().t()).to(bf).float()[:, :n]
        x = C.float()[:, :n]
        rstd = stats[:, 1:2]; mu = stats[:, 0:1]
        xhat = (x - mu) * rstd
        dxh = dxn * gamma
        dx = rstd * (dxh - dxh.mean(1, keepdim=True) - xhat * (dxh * xhat).mean(1, keepdim=True))
        ref = torch.zeros(M, N, device=dev)
        ref[:, :n] = dres.float()[:, :n] + dx.to(bf).float()
        stat(f"lnbwd.out K{K}", out, ref, tol=3e-2)
        dg = (dxn * xhat).sum(0); db = dxn.sum(0)
        p = partials.sum(0)
        stat(f"lnbwd.dgamma K{K}", p[0:1, :n], dg[None], tol=1e-2); stat(f"lnbwd.dbeta K{K}", p[1:2, :n], db[None], tol=1e-2)


def case_wgrad():
    for (T, Ca, Cb, splits) in [(256, 128, 64, 1), (4096, 576, 192, 4), (8192, 768, 192, 8), (2048, 192, 192, 2)]:
        A = torch.randn(T, Ca, device=dev).to(bf)
        B = torch.randn(T, Cb, device=dev).to(bf)
        ws = torch.zeros(capi.wgrad_workspace_elems(Ca, Cb, splits), device=dev)
        rows = ((Ca + 127) // 128) * 128
        out = torch.full((rows, Cb), float("nan"), device=dev)
        capi.gemm_wgrad(A, B, ws, splits, out)
        torch.cuda.synchronize()
        ref = torch.zeros(rows, Cb, device=dev)
        ref[:Ca] = A.float().t() @ B.float()
        stat(f"wgrad T{T} Ca{Ca} Cb{Cb} s{splits}", out, ref, tol=1e-2)


if __name__ == "__main__":
    print(capi.version(), torch.cuda.get_device_name(0), flush=True)
    case = sys.argv[1]
    if case == "store": case_store()
    elif case == "gelu_mul": case_gelu_mul()
    elif case == "ln": case_ln()
    elif case == "wgrad": case_wgrad()
    torch.cuda.synchronize()
    print("done", case, flush=True)
