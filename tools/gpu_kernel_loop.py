"""Launch one libsrk kernel a few times at the bench shapes (for `ncu --set full`).  Usage: gpu_kernel_loop.py <which>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200 import _capi as capi

which = sys.argv[1] if len(sys.argv) > 1 else "gelu2"
B = 16
T = B * 16384
bf = torch.bfloat16
dev = "cuda"
torch.manual_seed(0)
reps = 3
if which == "gelu2":
    A = torch.randn(T, 192, device=dev).to(bf); W = (torch.randn(768, 192, device=dev) / 14).to(bf)
    C = torch.empty(T, 768, device=dev, dtype=bf); C2 = torch.empty_like(C, dtype=torch.float16)
    ln = capi.make_ln_args(768, 720, None)
    for _ in range(reps): capi.gemm_tn(capi.EPI_GELU2, A, W, C, C2=C2, ln=ln)
elif which == "mul":
    A = torch.randn(T, 192, device=dev).to(bf); W = (torch.randn(768, 192, device=dev) / 14).to(bf)
    X1 = torch.randn(T, 768, device=dev).to(torch.float16); C = torch.empty(T, 768, device=dev, dtype=bf)
    for _ in range(reps): capi.gemm_tn(capi.EPI_MUL, A, W, C, X1=X1)
elif which == "qkv":
    A = torch.randn(T, 192, device=dev).to(bf); W = (torch.randn(576, 192, device=dev) / 14).to(bf)
    C = torch.empty(T, 576, device=dev, dtype=bf)
    for _ in range(reps): capi.gemm_tn(capi.EPI_STORE, A, W, C)
elif which == "res_ln":
    A = torch.randn(T, 768, device=dev).to(bf); W = (torch.randn(192, 768, device=dev) / 28).to(bf)
    R = torch.randn(T, 192, device=dev).to(bf); C = torch.empty(T, 192, device=dev, dtype=bf); C2 = torch.empty_like(C)
    st = torch.empty(T, 2, device=dev); g = torch.ones(180, device=dev); b = torch.zeros(180, device=dev)
    ln = capi.make_ln_args(180, 180, g, b, stats=st)
    for _ in range(reps): capi.gemm_tn(capi.EPI_RES_LN, A, W, C, C2=C2, X1=R, ln=ln)
elif which == "attn_fwd":
    qkv = torch.randn(T, 576, device=dev).to(bf); tab = torch.randn(225, 6, device=dev)
    out = torch.empty(T, 192, device=dev, dtype=bf)
    for _ in range(reps): capi.win_attn_fwd(capi.SrkGeom(B, 128, 128, 8, 4), 6, qkv, tab, out, ones_col=30)
elif which == "attn_bwd":
    qkv = torch.randn(T, 576, device=dev).to(bf); tab = torch.randn(225, 6, device=dev)
    do = torch.randn(T, 192, device=dev).to(bf); dq = torch.empty_like(qkv); dt = torch.empty_like(tab)
    for _ in range(reps): capi.win_attn_bwd(capi.SrkGeom(B, 128, 128, 8, 4), 6, qkv, tab, do, dq, dt)
elif which == "wgrad":
    A = torch.randn(T, 768, device=dev).to(bf); Bm = torch.randn(T, 192, device=dev).to(bf)
    ws = torch.empty(148 * 256 * 256, device=dev); out = torch.empty(768 * 256, device=dev)
    for _ in range(reps): capi.gemm_wgrad(A, Bm, ws, capi.wgrad_splits(T, 768), out)
elif which == "lnbwd":
    A = torch.randn(T, 768, device=dev).to(bf); W = (torch.randn(192, 768, device=dev) / 28).to(bf)
    X = torch.randn(T, 192, device=dev).to(bf); R = torch.randn(T, 192, device=dev).to(bf)
    C = torch.empty(T, 192, device=dev, dtype=bf); st = torch.empty(T, 2, device=dev)
    g = torch.ones(180, device=dev); b = torch.zeros(180, device=dev)
    capi.layernorm_fwd(X, C, st, g, b, 180, ones_col=180)
    parts = torch.empty(capi.gemm_grid(T, 192) * 2 * 192, device=dev)
    ln = capi.make_ln_args(180, -1, g, None, stats=st, partials=parts)
    for _ in range(reps): capi.gemm_tn(capi.EPI_LNBWD, A, W, C, X1=X, X2=R, ln=ln)
elif which in ("attn16_fwd", "attn16_bwd", "oca_fwd", "oca_bwd"):
    B8 = 8
    T8 = B8 * 16384
    mode = capi.ATTN_SELF if which.startswith("attn16") else capi.ATTN_OCA
    rows = 961 if mode == capi.ATTN_SELF else 1521
    geom = capi.SrkGeom(B8, 128, 128, 16, 8 if mode == capi.ATTN_SELF else 0)
    qkv = torch.randn(T8, 576, device=dev).to(bf); tab = torch.randn(rows, 6, device=dev)
    out = torch.empty(T8, 192, device=dev, dtype=bf); lse = torch.empty(6, T8, device=dev)
    capi.win_attn16_fwd(geom, mode, 6, qkv, tab, out, lse, ones_col=30)
    if which.endswith("fwd"):
        for _ in range(reps): capi.win_attn16_fwd(geom, mode, 6, qkv, tab, out, lse, ones_col=30)
    else:
        do = torch.randn(T8, 192, device=dev).to(bf); dq = torch.empty_like(qkv); dt = torch.empty_like(tab)
        ws = torch.empty(capi.attn16_bwd_ws_bytes(geom, mode, 6), device=dev, dtype=torch.uint8)
        for _ in range(reps): capi.win_attn16_bwd(geom, mode, 6, qkv, tab, out, do, lse, dq, ws, dt)
elif which in ("rdb_fwd", "rdb_dgrad", "rdb_wgrad"):
    # conv3 of a residual dense block at the hybrid bench shapes: batch 8, 256 x 256, nf 48, gc 24 (cin 96 -> 24)
    from superresolution_def_b200 import conv_engine as cv
    Bh, Hh, Wh, nf, gc, k = 8, 256, 256, 48, 24, 2
    Th = Bh * Hh * Wh
    cin = nf + k * gc
    cat = torch.randn(Th, nf + 4 * gc, device=dev).to(bf); dcat = torch.randn(Th, nf + 4 * gc, device=dev).to(bf)
    w = torch.randn(gc, cin, 3, 3, device=dev) / 30; b = torch.zeros(gc, device=dev)
    wf, wt, bp = cv.conv_weights(w, b, 64, 128)
    V = capi.view
    if which == "rdb_fwd":
        for _ in range(reps): capi.conv3x3_igemm_v(capi.CEPI_BIAS_LRELU, Bh, Hh, Wh, 128, 64, gc, V(cat, 0, cin), wf, bp, V(cat, cin, gc), slope=0.2)
    elif which == "rdb_dgrad":
        for _ in range(reps): capi.conv3x3_igemm_v(capi.CEPI_BIAS_RES, Bh, Hh, Wh, 64, 128, cin, V(dcat, cin, gc), wt, None, V(dcat, 0, cin), V(dcat, 0, cin))
    else:
        dw = torch.empty_like(w)
        for _ in range(reps): capi.conv3x3_wgrad_v(Bh, Hh, Wh, cin, gc, 128, 64, V(dcat, cin, gc), V(cat, 0, cin), dw)
torch.cuda.synchronize()
print("done", which)
