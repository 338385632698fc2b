"""ctypes binding of libsrk.so (include/srk.h).  No torch types cross the boundary: only raw device
pointers, sizes and the current CUDA stream handle.  The library is required: importing this module
on a machine without the built .so raises, and every call raises on a non-zero return code.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_longlong, c_void_p
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "_lib" / "libsrk.so"
if os.environ.get("SRK_LIB"):   # tools only: an instrumented build of the same sources (e.g. -DSRK_TC8_PROFILE)
    _LIB_PATH = Path(os.environ["SRK_LIB"])


class SrkError(RuntimeError):
    pass


class SrkLnArgs(Structure):
    _fields_ = [
        ("n_real", c_int),
        ("ones_col", c_int),
        ("gamma", c_void_p),
        ("beta", c_void_p),
        ("stats", c_void_p),
        ("partials", c_void_p),
        ("eps", c_float),
        ("row_scale", c_void_p),
        ("rows_per_scale", c_int),
    ]


def _load() -> ctypes.CDLL:
    if not _LIB_PATH.exists():
        raise SrkError(
            f"{_LIB_PATH} is missing: build it with `python -m superresolution_def_b200._build` "
            "(there is no CPU or PyTorch fallback for the kernels)")
    return ctypes.CDLL(str(_LIB_PATH))


lib = _load()
lib.srk_version.restype = c_char_p
lib.srk_launch_count.restype = ctypes.c_longlong


def _sig(name, argtypes):
    fn = getattr(lib, name)
    fn.restype = c_int
    fn.argtypes = argtypes
    return fn


_gemm_tn = _sig("srk_gemm_tn", [c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                 c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, POINTER(SrkLnArgs), c_void_p])
_gemm_grid = _sig("srk_gemm_grid", [c_int, c_int])
_gemm_wgrad = _sig("srk_gemm_wgrad", [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int,
                                       c_void_p, c_void_p])
_wgrad_splits = _sig("srk_gemm_wgrad_splits", [c_int, c_int])
_wgrad_ws_elems = _sig("srk_gemm_wgrad_workspace_elems", [c_int, c_int, c_int])
_wgrad_ws_elems.restype = c_longlong

_mlp_fwd = _sig("srk_mlp_fwd", [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int, POINTER(SrkLnArgs), c_void_p])

EPI_STORE, EPI_GELU2, EPI_MUL, EPI_RES_LN, EPI_LNBWD, EPI_GELU1, EPI_MULG = range(7)


def launch_count() -> int:
    return int(lib.srk_launch_count())


def version() -> str:
    return lib.srk_version().decode()


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise SrkError(f"{what} failed with code {rc} (see stderr)")


def _stream() -> int:
    # the raw handle of torch's current stream; torch.cuda.current_stream().cuda_stream builds a Stream object per call
    # (device-index resolution, lazy-init and availability checks: ~26 % of the host time of an eager discriminator step,
    # tools/gpu_disc_hostprof.py), and every libsrk call needs it
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "libsrk operates on CUDA tensors only"
    return t.data_ptr()


def _ld(t):
    if t is None:
        return 0
    assert t.dim() == 2 and t.stride(1) == 1
    return t.stride(0)


def gemm_grid(M: int, N: int) -> int:
    return _gemm_grid(M, N)


def gemm_tn(epi, A, B, C, C2=None, X1=None, X2=None, ln: SrkLnArgs | None = None):
    """C[M,N] = epilogue(A[M,K] @ B[N,K]^T); all bf16 2-D row-major (last stride 1)."""
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K and tuple(C.shape) == (M, N)
    # gelu' travels as fp16: the second output of EPI_GELU2 and the multiplier of EPI_MUL (include/srk.h)
    f16 = {EPI_GELU2: ("C2",), EPI_MUL: ("X1",)}.get(epi, ())
    for name, t in (("A", A), ("B", B), ("C", C), ("C2", C2), ("X1", X1), ("X2", X2)):
        assert t is None or t.dtype == (torch.float16 if name in f16 else torch.bfloat16), (name, t.dtype)
    rc = _gemm_tn(epi, M, N, K, _ptr(A), _ld(A), _ptr(B), _ld(B), _ptr(C), _ld(C), _ptr(C2), _ld(C2),
                  _ptr(X1), _ld(X1), _ptr(X2), _ld(X2), ctypes.byref(ln) if ln is not None else None, _stream())
    _check(rc, "srk_gemm_tn")


def make_ln_args(n_real, ones_col, gamma, beta=None, stats=None, partials=None, eps=1e-5, row_scale=None,
                 rows_per_scale=1) -> SrkLnArgs:
    for t in (gamma, beta, stats, partials, row_scale):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    return SrkLnArgs(n_real, ones_col, _ptr(gamma), _ptr(beta), _ptr(stats), _ptr(partials), eps, _ptr(row_scale),
                     rows_per_scale)


def mlp_fwd(xn2, w1, w2, resid, act, dact, x_out, xn_out, hid_ones_col: int, ln: SrkLnArgs):
    """Fused fc1 -> GELU -> fc2 -> residual -> LayerNorm (srk_mlp_fwd).  act / dact may be None (inference)."""
    T, Cp = xn2.shape
    Hp = w1.shape[0]
    assert tuple(w1.shape) == (Hp, Cp) and tuple(w2.shape) == (Cp, Hp)
    for t in (xn2, w1, w2, resid, act, x_out, xn_out):
        assert t is None or (t.dtype == torch.bfloat16 and t.is_contiguous())
    assert dact is None or (dact.dtype == torch.float16 and dact.is_contiguous())   # gelu' is stored as fp16
    rc = _mlp_fwd(T, Cp, Hp, _ptr(xn2), _ptr(w1), _ptr(w2), _ptr(resid), _ptr(act), _ptr(dact), _ptr(x_out), _ptr(xn_out),
                  hid_ones_col, ctypes.byref(ln), _stream())
    _check(rc, "srk_mlp_fwd")


_u16_to_f32_aug = _sig("srk_u16_to_f32_aug", [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p])
_f32_to_u16 = _sig("srk_f32_to_u16", [c_void_p, c_void_p, c_longlong, c_void_p])


def u16_to_f32_aug(src: torch.Tensor, dst: torch.Tensor, codes: torch.Tensor | None):
    """src [B, n, n] uint16 (CUDA), dst [B, 1, n, n] or [B, n, n] float32, codes int32 [B] (fh | fv<<1 | k<<2) or None."""
    B, n = src.shape[0], src.shape[-1]
    assert src.is_cuda and src.dtype == torch.uint16 and src.is_contiguous() and src.shape[-2] == n
    assert dst.dtype == torch.float32 and dst.is_contiguous() and dst.numel() == src.numel()
    assert codes is None or (codes.dtype == torch.int32 and codes.is_cuda and codes.numel() == B)
    _check(_u16_to_f32_aug(_ptr(src), _ptr(dst), _ptr(codes), B, n, _stream()), "srk_u16_to_f32_aug")


def f32_to_u16(src: torch.Tensor, dst: torch.Tensor):
    assert src.is_cuda and src.dtype == torch.float32 and src.is_contiguous()
    assert dst.dtype == torch.uint16 and dst.is_contiguous() and dst.numel() == src.numel()
    _check(_f32_to_u16(_ptr(src), _ptr(dst), src.numel(), _stream()), "srk_f32_to_u16")


def wgrad_workspace_elems(Ca: int, Cb: int, splits: int) -> int:
    return int(_wgrad_ws_elems(Ca, Cb, splits))


def wgrad_splits(T: int, Ca: int) -> int:
    """Token ranges that fill the GPU for this shape (the library pairs 128-channel tiles of A per CTA)."""
    return int(_wgrad_splits(T, Ca))


def gemm_wgrad(A, B, workspace, splits, out):
    """out[ceil128(Ca), Cb] (fp32) = A[T,Ca]^T @ B[T,Cb]."""
    T, Ca = A.shape
    Cb = B.shape[1]
    assert B.shape[0] == T and A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert workspace.dtype == torch.float32 and workspace.numel() >= wgrad_workspace_elems(Ca, Cb, splits)
    assert out.dtype == torch.float32 and out.numel() >= ((Ca + 127) // 128) * 128 * Cb
    rc = _gemm_wgrad(T, Ca, Cb, _ptr(A), _ld(A), _ptr(B), _ld(B), _ptr(workspace), splits, _ptr(out), _stream())
    _check(rc, "srk_gemm_wgrad")


# ---------------------------------------------------------------------------------------------------
# Block-level API (include/srk.h, second half)
# ---------------------------------------------------------------------------------------------------
from ctypes import c_longlong  # noqa: E402


class SrkBlockDims(Structure):
    _fields_ = [(n, c_int) for n in ("C", "Cp", "heads", "dh", "ds", "hidden", "Hp")]


class SrkGeom(Structure):
    _fields_ = [(n, c_int) for n in ("B", "H", "W", "ws", "shift")]


PARAM_NAMES = ("norm1_w", "norm1_b", "rpb_table", "qkv_w", "qkv_b", "proj_w", "proj_b", "norm2_w", "norm2_b",
               "fc1_w", "fc1_b", "fc2_w", "fc2_b")
WEIGHT_NAMES = ("qkv_f", "qkv_t", "proj_f", "proj_t", "fc1_f", "fc1_t", "fc2_f", "fc2_t")
ACT_NAMES = ("x_in", "xn1", "stats1", "qkv", "ao", "x_mid", "xn2", "stats2", "act", "dact", "x_out", "xn_out",
             "stats_out")
SCRATCH_NAMES = ("d_act", "d_ao", "d_qkv", "g_mid", "wg_ws")


class SrkBlockParams(Structure):
    _fields_ = [(n, c_void_p) for n in PARAM_NAMES]


class SrkBlockGrads(Structure):
    _fields_ = [(n, c_void_p) for n in PARAM_NAMES]


class SrkBlockWeights(Structure):
    _fields_ = [(n, c_void_p) for n in WEIGHT_NAMES]


class SrkBlockActs(Structure):
    _fields_ = [(n, c_void_p) for n in ACT_NAMES]


class SrkBlockScratch(Structure):
    _fields_ = [(n, c_void_p) for n in SCRATCH_NAMES]


lib.srk_block_weight_elems.restype = None
lib.srk_block_weight_elems.argtypes = [POINTER(SrkBlockDims), POINTER(c_longlong * 8)]
lib.srk_block_bwd_scratch_floats.restype = c_longlong
lib.srk_block_bwd_scratch_floats.argtypes = [POINTER(SrkBlockDims), POINTER(SrkGeom)]
lib.srk_win_attn_bwd_ws_floats.restype = c_longlong
lib.srk_win_attn_bwd_ws_floats.argtypes = [c_int]
lib.srk_layernorm_bwd_ws_floats.restype = c_longlong
lib.srk_layernorm_bwd_ws_floats.argtypes = [c_int]
_block_prep = _sig("srk_block_prep_weights", [POINTER(SrkBlockDims), POINTER(SrkBlockParams), POINTER(SrkBlockWeights),
                                              c_void_p])
_block_fwd = _sig("srk_swin_block_fwd", [POINTER(SrkBlockDims), POINTER(SrkGeom), POINTER(SrkBlockWeights),
                                         POINTER(SrkBlockParams), c_void_p, c_void_p, POINTER(SrkBlockActs), c_void_p])
_block_bwd = _sig("srk_swin_block_bwd", [POINTER(SrkBlockDims), POINTER(SrkGeom), POINTER(SrkBlockWeights),
                                         POINTER(SrkBlockParams), POINTER(SrkBlockActs), c_void_p,
                                         POINTER(SrkBlockScratch), c_void_p, POINTER(SrkBlockGrads), c_int, c_void_p])
_attn_fwd = _sig("srk_win_attn_fwd", [POINTER(SrkGeom), c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                      c_void_p])
_attn_bwd = _sig("srk_win_attn_bwd", [POINTER(SrkGeom), c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p])
_ln_fwd = _sig("srk_layernorm_fwd", [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                     c_int, c_int, c_float, c_void_p])
_ln_bwd = _sig("srk_layernorm_bwd", [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                     c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p])


def block_weight_elems(dims: SrkBlockDims):
    out = (c_longlong * 8)()
    lib.srk_block_weight_elems(ctypes.byref(dims), ctypes.byref(out))
    return list(out)


def block_bwd_scratch_floats(dims: SrkBlockDims, geom: SrkGeom) -> int:
    return int(lib.srk_block_bwd_scratch_floats(ctypes.byref(dims), ctypes.byref(geom)))


def _fill(struct_cls, names, tensors: dict):
    """struct of device pointers in field order (positional construction: one C call instead of a setattr per field)"""
    get = tensors.get
    return struct_cls(*[None if (t := get(n)) is None else t.data_ptr() for n in names])


def block_prep_weights(dims, params: dict, weights: dict):
    rc = _block_prep(ctypes.byref(dims), ctypes.byref(_fill(SrkBlockParams, PARAM_NAMES, params)),
                     ctypes.byref(_fill(SrkBlockWeights, WEIGHT_NAMES, weights)), _stream())
    _check(rc, "srk_block_prep_weights")


def swin_block_fwd(dims, geom, weights: dict, params: dict, next_norm_w, next_norm_b, acts: dict):
    rc = _block_fwd(ctypes.byref(dims), ctypes.byref(geom), ctypes.byref(_fill(SrkBlockWeights, WEIGHT_NAMES, weights)),
                    ctypes.byref(_fill(SrkBlockParams, PARAM_NAMES, params)), _ptr(next_norm_w), _ptr(next_norm_b),
                    ctypes.byref(_fill(SrkBlockActs, ACT_NAMES, acts)), _stream())
    _check(rc, "srk_swin_block_fwd")


def swin_block_bwd(dims, geom, weights: dict, params: dict, acts: dict, g_out, scratch: dict, g_in, grads: dict,
                   accumulate=False):
    rc = _block_bwd(ctypes.byref(dims), ctypes.byref(geom), ctypes.byref(_fill(SrkBlockWeights, WEIGHT_NAMES, weights)),
                    ctypes.byref(_fill(SrkBlockParams, PARAM_NAMES, params)),
                    ctypes.byref(_fill(SrkBlockActs, ACT_NAMES, acts)), _ptr(g_out),
                    ctypes.byref(_fill(SrkBlockScratch, SCRATCH_NAMES, scratch)), _ptr(g_in),
                    ctypes.byref(_fill(SrkBlockGrads, PARAM_NAMES, grads)), int(accumulate), _stream())
    _check(rc, "srk_swin_block_bwd")


def win_attn_fwd(geom, heads, qkv, rpb_table, out, ones_col=-1):
    rc = _attn_fwd(ctypes.byref(geom), heads, _ptr(qkv), _ld(qkv), _ptr(rpb_table), _ptr(out), _ld(out), ones_col,
                   _stream())
    _check(rc, "srk_win_attn_fwd")


def win_attn_bwd(geom, heads, qkv, rpb_table, d_out, d_qkv, d_rpb_table=None):
    ws = torch.empty(int(lib.srk_win_attn_bwd_ws_floats(heads)), device=qkv.device, dtype=torch.float32)
    rc = _attn_bwd(ctypes.byref(geom), heads, _ptr(qkv), _ld(qkv), _ptr(rpb_table), _ptr(d_out), _ld(d_out),
                   _ptr(d_qkv), _ptr(ws), _ptr(d_rpb_table), _stream())
    _check(rc, "srk_win_attn_bwd")


def layernorm_fwd(x, y, stats, gamma, beta, C, ones_col=-1, eps=1e-5):
    rows, Cp = x.shape
    rc = _ln_fwd(_ptr(x), _ld(x), _ptr(y), _ld(y), _ptr(stats), _ptr(gamma), _ptr(beta), rows, C, Cp, ones_col, eps,
                 _stream())
    _check(rc, "srk_layernorm_fwd")


def layernorm_bwd(dy, x, stats, gamma, dres, dx, dgamma, dbeta, C):
    rows, Cp = x.shape
    ws = torch.empty(int(lib.srk_layernorm_bwd_ws_floats(Cp)), device=x.device, dtype=torch.float32)
    rc = _ln_bwd(_ptr(dy), _ld(dy), _ptr(x), _ld(x), _ptr(stats), _ptr(gamma), _ptr(dres), _ld(dres) if dres is not None else 0,
                 _ptr(dx), _ld(dx), _ptr(ws), _ptr(dgamma), _ptr(dbeta), rows, C, Cp, _stream())
    _check(rc, "srk_layernorm_bwd")


# ---------------------------------------------------------------------------------------------------
# Convolution API (include/srk.h, third part)
# ---------------------------------------------------------------------------------------------------
CEPI_BIAS, CEPI_BIAS_LRELU, CEPI_BIAS_RES, CEPI_MASK_LRELU, CEPI_BIAS_GELU, CEPI_MUL, CEPI_OUT1 = range(7)
_conv_prep = _sig("srk_conv3x3_prep_weights", [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                               c_void_p, c_void_p, c_void_p])
_conv_igemm = _sig("srk_conv3x3_igemm", [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                         c_void_p, c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p])
_conv_wgrad = _sig("srk_conv3x3_wgrad", [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p])
lib.srk_conv3x3_wgrad_ws_floats.restype = c_longlong
lib.srk_conv3x3_wgrad_ws_floats.argtypes = [c_int, c_int]
lib.srk_small_ws_floats.restype = c_longlong
_bias_grad = _sig("srk_bias_grad_nhwc", [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                         c_void_p])
_conv_in1_fwd = _sig("srk_conv_in1_fwd", [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p])
_conv_in1_wgrad = _sig("srk_conv_in1_wgrad", [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                              c_int, c_int, c_void_p])
_conv_out1_fwd = _sig("srk_conv_out1_fwd", [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                            c_void_p])
_conv_out1_bwd = _sig("srk_conv_out1_bwd", [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_int, c_int, c_int, c_int, c_void_p])

_ws_cache: dict = {}
_retired: list = []   # outgrown workspaces: kept alive because a captured CUDA graph may still hold their addresses


def retire(t) -> None:
    """Keep an outgrown scratch tensor allocated for the life of the process.  A step captured by graphs.GraphedStep
    has raw scratch pointers baked into its kernel nodes, so a workspace may be replaced by a larger one but must
    never be returned to the allocator (a later replay would write into whatever reuses that memory)."""
    if t is not None:
        _retired.append(t)


def _ws(n: int, device) -> torch.Tensor:
    """A reusable fp32 workspace of at least n floats (same-stream reuse is ordered by the stream).  Grow-only:
    an outgrown buffer is retired, never freed (see retire())."""
    k = str(device)
    t = _ws_cache.get(k)
    if t is None or t.numel() < n:
        retire(t)
        t = _ws_cache[k] = torch.empty(max(n, 1 << 22), device=device, dtype=torch.float32)
    return t


def conv3x3_prep_weights(w, bias, Cout_p, Cin_p, ps, wf, wt, bias_packed):
    Cout, Cin = w.shape[0], w.shape[1]
    rc = _conv_prep(_ptr(w), _ptr(bias), Cout, Cin, Cout_p, Cin_p, int(ps), _ptr(wf), _ptr(wt), _ptr(bias_packed),
                    _stream())
    _check(rc, "srk_conv3x3_prep_weights")


def conv3x3_igemm(epi, B, H, W, Cin_p, Cout_p, n_real, x, wk, bias, y, x_ps=False, y_ps=False, y2=None, r=None,
                  slope=0.01):
    rc = _conv_igemm(epi, B, H, W, Cin_p, Cout_p, n_real, _ptr(x), int(x_ps), _ptr(wk), _ptr(bias), slope, _ptr(y),
                     int(y_ps), _ptr(y2), _ptr(r), _stream())
    _check(rc, "srk_conv3x3_igemm")


def conv3x3_wgrad(B, H, W, Cin, Cout, Cin_p, Cout_p, ps, dy, x, dw):
    ws = _ws(int(lib.srk_conv3x3_wgrad_ws_floats(Cin_p, Cout_p)), x.device)
    rc = _conv_wgrad(B, H, W, Cin, Cout, Cin_p, Cout_p, int(ps), _ptr(dy), _ptr(x), _ptr(ws), _ptr(dw), _stream())
    _check(rc, "srk_conv3x3_wgrad")


def bias_grad_nhwc(dy, B, H, W, C, ps, db):
    ws = _ws(int(lib.srk_small_ws_floats()), dy.device)
    rc = _bias_grad(_ptr(dy), B, H, W, C, int(ps), _ptr(ws), _ptr(db), db.numel(), _stream())
    _check(rc, "srk_bias_grad_nhwc")


def conv_in1_fwd(x, w, bias, y, B, H, W, C, Cp):
    _check(_conv_in1_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), B, H, W, C, Cp, _stream()), "srk_conv_in1_fwd")


def conv_in1_wgrad(x, dy, dw, db, B, H, W, C, Cp):
    ws = _ws(int(lib.srk_small_ws_floats()), dy.device)
    _check(_conv_in1_wgrad(_ptr(x), _ptr(dy), _ptr(ws), _ptr(dw), _ptr(db), B, H, W, C, Cp, _stream()),
           "srk_conv_in1_wgrad")


def conv_out1_fwd(x, w, bias, y, B, H, W, C):
    _check(_conv_out1_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), B, H, W, C, _stream()), "srk_conv_out1_fwd")


def conv_out1_bwd(dy, x, w, dx, dw, db, B, H, W, C):
    ws = _ws(int(lib.srk_small_ws_floats()), dy.device)
    _check(_conv_out1_bwd(_ptr(dy), _ptr(x), _ptr(w), _ptr(dx), _ptr(ws), _ptr(dw), _ptr(db), B, H, W, C, _stream()),
           "srk_conv_out1_bwd")


# ---------------------------------------------------------------------------------------------------
# HAT API (include/srk.h, fourth part): 16x16-window attention cores, HAB/OCAB orchestration, channel attention
# ---------------------------------------------------------------------------------------------------
ATTN_SELF, ATTN_OCA = 0, 1


class SrkHatExtra(Structure):
    _fields_ = [("mode", c_int), ("res_in", c_void_p), ("lse", c_void_p), ("attn_ws", c_void_p), ("d_xn1", c_void_p),
                ("drop_attn", c_void_p), ("drop_mlp", c_void_p), ("gs_buf", c_void_p)]


_attn16_fwd = _sig("srk_win_attn16_fwd", [POINTER(SrkGeom), c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                          c_void_p, c_int, c_void_p])
_attn16_bwd = _sig("srk_win_attn16_bwd", [POINTER(SrkGeom), c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                          c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p])
lib.srk_win_attn16_bwd_ws_bytes.restype = c_longlong
lib.srk_win_attn16_bwd_ws_bytes.argtypes = [POINTER(SrkGeom), c_int, c_int]
_hat_fwd = _sig("srk_hat_block_fwd", [POINTER(SrkBlockDims), POINTER(SrkGeom), POINTER(SrkBlockWeights),
                                      POINTER(SrkBlockParams), c_void_p, c_void_p, POINTER(SrkBlockActs),
                                      POINTER(SrkHatExtra), c_void_p])
_hat_bwd = _sig("srk_hat_block_bwd", [POINTER(SrkBlockDims), POINTER(SrkGeom), POINTER(SrkBlockWeights),
                                      POINTER(SrkBlockParams), POINTER(SrkBlockActs), c_void_p, POINTER(SrkBlockScratch),
                                      c_void_p, POINTER(SrkBlockGrads), POINTER(SrkHatExtra), c_void_p])
_cab_se_fwd = _sig("srk_cab_se_fwd", [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p])
_cab_se_bwd = _sig("srk_cab_se_bwd", [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p])


def attn16_bwd_ws_bytes(geom, mode, heads) -> int:
    return int(lib.srk_win_attn16_bwd_ws_bytes(ctypes.byref(geom), mode, heads))


def win_attn16_fwd(geom, mode, heads, qkv, rpb_table, out, lse, ones_col=-1):
    rc = _attn16_fwd(ctypes.byref(geom), mode, heads, _ptr(qkv), _ld(qkv), _ptr(rpb_table), _ptr(out), _ld(out),
                     _ptr(lse), ones_col, _stream())
    _check(rc, "srk_win_attn16_fwd")


def win_attn16_bwd(geom, mode, heads, qkv, rpb_table, out, d_out, lse, d_qkv, ws, d_rpb_table=None):
    rc = _attn16_bwd(ctypes.byref(geom), mode, heads, _ptr(qkv), _ld(qkv), _ptr(rpb_table), _ptr(out), _ptr(d_out),
                     _ld(d_out), _ptr(lse), _ptr(d_qkv), _ptr(ws), _ptr(d_rpb_table), _stream())
    _check(rc, "srk_win_attn16_bwd")


def _hat_extra(mode, res_in, lse, attn_ws=None, d_xn1=None, drop=(None, None), gs_buf=None):
    return SrkHatExtra(mode, _ptr(res_in), _ptr(lse), _ptr(attn_ws), _ptr(d_xn1), _ptr(drop[0]), _ptr(drop[1]),
                       _ptr(gs_buf))


def hat_block_fwd(dims, geom, weights: dict, params: dict, next_norm_w, next_norm_b, acts: dict, mode, res_in, lse,
                  drop=(None, None)):
    x = _hat_extra(mode, res_in, lse, drop=drop)
    rc = _hat_fwd(ctypes.byref(dims), ctypes.byref(geom), ctypes.byref(_fill(SrkBlockWeights, WEIGHT_NAMES, weights)),
                  ctypes.byref(_fill(SrkBlockParams, PARAM_NAMES, params)), _ptr(next_norm_w), _ptr(next_norm_b),
                  ctypes.byref(_fill(SrkBlockActs, ACT_NAMES, acts)), ctypes.byref(x), _stream())
    _check(rc, "srk_hat_block_fwd")


def hat_block_bwd(dims, geom, weights: dict, params: dict, acts: dict, g_out, scratch: dict, g_in, grads: dict, mode, lse,
                  attn_ws, d_xn1=None, drop=(None, None), gs_buf=None):
    x = _hat_extra(mode, None, lse, attn_ws, d_xn1, drop, gs_buf)
    rc = _hat_bwd(ctypes.byref(dims), ctypes.byref(geom), ctypes.byref(_fill(SrkBlockWeights, WEIGHT_NAMES, weights)),
                  ctypes.byref(_fill(SrkBlockParams, PARAM_NAMES, params)),
                  ctypes.byref(_fill(SrkBlockActs, ACT_NAMES, acts)), _ptr(g_out),
                  ctypes.byref(_fill(SrkBlockScratch, SCRATCH_NAMES, scratch)), _ptr(g_in),
                  ctypes.byref(_fill(SrkBlockGrads, PARAM_NAMES, grads)), ctypes.byref(x), _stream())
    _check(rc, "srk_hat_block_bwd")


def cab_se_fwd(y, x, B, HW, C, S, w1, b1, w2, b2, alpha, pool, hidden, scale, out):
    ws = _ws(int(lib.srk_small_ws_floats()), y.device)
    rc = _cab_se_fwd(_ptr(y), _ptr(x), B, HW, C, y.shape[1], S, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), alpha, _ptr(ws),
                     _ptr(pool), _ptr(hidden), _ptr(scale), _ptr(out), _stream())
    _check(rc, "srk_cab_se_fwd")


def cab_se_bwd(g, y, B, HW, C, S, w1, w2, alpha, pool, hidden, scale, dy, dw1, db1, dw2, db2):
    ws = _ws(int(lib.srk_small_ws_floats()), y.device)
    rc = _cab_se_bwd(_ptr(g), _ptr(y), B, HW, C, y.shape[1], S, _ptr(w1), _ptr(w2), alpha, _ptr(pool), _ptr(hidden),
                     _ptr(scale), _ptr(ws), _ptr(dy), _ptr(dw1), _ptr(db1), _ptr(dw2), _ptr(db2), _stream())
    _check(rc, "srk_cab_se_bwd")


# ---------------------------------------------------------------------------------------------------
# Channel-slice ("view") API (include/srk.h): dense blocks / tail of the hybrid generator
# ---------------------------------------------------------------------------------------------------
class SrkView(Structure):
    _fields_ = [("ptr", c_void_p), ("C", c_int), ("pitch", c_int)]


def view(t: torch.Tensor, c0: int = 0, C: int | None = None) -> SrkView:
    """Channel slice [c0, c0 + C) of a token-major bf16 tensor [pixels, pitch]."""
    if t.dtype != torch.bfloat16 or t.dim() != 2 or not t.is_contiguous() or not t.is_cuda:
        raise SrkError("view: contiguous CUDA bf16 [pixels, channels] tensor expected")
    C = t.shape[1] - c0 if C is None else C
    if c0 % 8 or C % 8 or C <= 0 or c0 + C > t.shape[1]:
        raise SrkError(f"view: channel slice [{c0}, {c0 + C}) of {t.shape[1]} must be 16-byte aligned")
    return SrkView(t.data_ptr() + 2 * c0, C, t.shape[1])


_conv_igemm_v = _sig("srk_conv3x3_igemm_v", [c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(SrkView), c_void_p,
                                             c_void_p, c_float, c_float, POINTER(SrkView), POINTER(SrkView), c_void_p,
                                             c_void_p])
_conv_wgrad_v = _sig("srk_conv3x3_wgrad_v", [c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(SrkView),
                                             POINTER(SrkView), c_void_p, c_void_p, c_void_p])
_bias_grad_v = _sig("srk_bias_grad_v", [POINTER(SrkView), c_longlong, c_void_p, c_void_p, c_int, c_void_p])
_view_lrelu_mask = _sig("srk_view_lrelu_mask", [POINTER(SrkView), POINTER(SrkView), c_longlong, c_float, c_void_p, c_void_p,
                                                c_void_p])
_view_axpy = _sig("srk_view_axpy", [POINTER(SrkView), POINTER(SrkView), POINTER(SrkView), c_longlong, c_float, c_void_p])
_nearest2_fwd = _sig("srk_nearest2_fwd", [POINTER(SrkView), POINTER(SrkView), c_int, c_int, c_int, c_void_p])
_nearest2_bwd = _sig("srk_nearest2_bwd", [POINTER(SrkView), POINTER(SrkView), c_int, c_int, c_int, c_void_p])
_img1_pack = _sig("srk_img1_pack", [c_void_p, c_void_p, c_longlong, c_void_p])
_img1_unpack = _sig("srk_img1_unpack", [c_void_p, c_void_p, c_longlong, c_void_p])


def _vref(v):
    return ctypes.byref(v) if v is not None else None


def conv3x3_igemm_v(epi, B, H, W, Cin_p, Cout_p, n_real, x: SrkView, wk, bias, y: SrkView | None, r: SrkView | None = None,
                    slope=0.01, alpha=1.0, y32=None):
    rc = _conv_igemm_v(epi, B, H, W, Cin_p, Cout_p, n_real, _vref(x), _ptr(wk), _ptr(bias), slope, alpha, _vref(y),
                       _vref(r), _ptr(y32), _stream())
    _check(rc, "srk_conv3x3_igemm_v")


def conv3x3_wgrad_v(B, H, W, Cin, Cout, Cin_p, Cout_p, dy: SrkView, x: SrkView, dw):
    ws = _ws(int(lib.srk_conv3x3_wgrad_ws_floats(Cin_p, Cout_p)), dw.device)
    rc = _conv_wgrad_v(B, H, W, Cin, Cout, Cin_p, Cout_p, _vref(dy), _vref(x), _ptr(ws), _ptr(dw), _stream())
    _check(rc, "srk_conv3x3_wgrad_v")


def bias_grad_v(dy: SrkView, npix, db):
    ws = _ws(int(lib.srk_small_ws_floats()), db.device)
    _check(_bias_grad_v(_vref(dy), npix, _ptr(ws), _ptr(db), db.numel(), _stream()), "srk_bias_grad_v")


def view_lrelu_mask(g: SrkView, f: SrkView, npix, slope, colsum=None):
    """colsum: optional fp32 [g.C] tensor receiving the column sums of the masked gradient (the bias gradient)."""
    ws = _ws(int(lib.srk_small_ws_floats()), colsum.device) if colsum is not None else None
    _check(_view_lrelu_mask(_vref(g), _vref(f), npix, slope, _ptr(ws), _ptr(colsum), _stream()), "srk_view_lrelu_mask")


def view_axpy(y: SrkView, a: SrkView, x: SrkView | None, npix, alpha):
    _check(_view_axpy(_vref(y), _vref(a), _vref(x), npix, alpha, _stream()), "srk_view_axpy")


def nearest2_fwd(x: SrkView, y: SrkView, B, H, W):
    _check(_nearest2_fwd(_vref(x), _vref(y), B, H, W, _stream()), "srk_nearest2_fwd")


def nearest2_bwd(dy: SrkView, dx: SrkView, B, H, W):
    _check(_nearest2_bwd(_vref(dy), _vref(dx), B, H, W, _stream()), "srk_nearest2_bwd")


def img1_pack(x, y8):
    _check(_img1_pack(_ptr(x), _ptr(y8), x.numel(), _stream()), "srk_img1_pack")


def img1_unpack(x8, y):
    _check(_img1_unpack(_ptr(x8), _ptr(y), y.numel(), _stream()), "srk_img1_unpack")


# ---------------------------------------------------------------------------------------------------
# UNetDiscriminatorSN helpers (include/srk.h): patch gather / fold around the tcgen05 GEMMs
# ---------------------------------------------------------------------------------------------------
EPI_LRELU = 7
FOLD_NONE, FOLD_LRELU, FOLD_MASK = 0, 1, 2

_gemm_tn_lrelu = _sig("srk_gemm_tn_lrelu", [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_float,
                                            c_void_p])
_disc_patches = _sig("srk_disc_patches_k4s2", [POINTER(SrkView), POINTER(SrkView), c_float, c_int, c_int, c_int, c_void_p,
                                               c_void_p])
_disc_fold = _sig("srk_disc_fold_k4s2", [c_void_p, c_int, c_int, c_int, POINTER(SrkView), POINTER(SrkView), c_int, c_float,
                                         POINTER(SrkView), c_void_p])
_view_lrelu = _sig("srk_view_lrelu", [POINTER(SrkView), c_longlong, c_float, c_void_p])


def gemm_tn_lrelu(A, B, C, slope):
    """C[M,N] = bf16(leaky_relu(A[M,K] @ B[N,K]^T, slope)); bf16 2-D row-major operands (last stride 1)."""
    M, K = A.shape
    N = B.shape[0]
    assert B.shape[1] == K and tuple(C.shape) == (M, N)
    for t in (A, B, C):
        assert t.dtype == torch.bfloat16
    _check(_gemm_tn_lrelu(M, N, K, _ptr(A), _ld(A), _ptr(B), _ld(B), _ptr(C), _ld(C), slope, _stream()), "srk_gemm_tn_lrelu")


def disc_patches_k4s2(x: SrkView, f: SrkView | None, slope, B, H, W, patches):
    """patches [>= B*(H/2)*(W/2), 16*C] bf16 contiguous <- 4x4 / stride-2 / pad-1 patches of the NHWC view x (masked by f)."""
    assert patches.dtype == torch.bfloat16 and patches.is_contiguous() and patches.shape[1] == 16 * x.C
    assert patches.shape[0] >= B * (H // 2) * (W // 2)
    _check(_disc_patches(_vref(x), _vref(f), slope, B, H, W, _ptr(patches), _stream()), "srk_disc_patches_k4s2")


def disc_fold_k4s2(taps, B, Hi, Wi, y: SrkView, add: SrkView | None = None, f: SrkView | None = None, act=FOLD_NONE,
                   slope=0.2):
    """NHWC view y [B, 2Hi, 2Wi, C] <- act(fold(taps [>= B*Hi*Wi, 16*C]) + add)."""
    assert taps.dtype == torch.bfloat16 and taps.is_contiguous() and taps.shape[1] == 16 * y.C
    assert taps.shape[0] >= B * Hi * Wi
    _check(_disc_fold(_ptr(taps), B, Hi, Wi, _vref(add), _vref(f), act, slope, _vref(y), _stream()), "srk_disc_fold_k4s2")


def view_lrelu(y: SrkView, npix, slope):
    _check(_view_lrelu(_vref(y), npix, slope, _stream()), "srk_view_lrelu")


_disc_prep_w4 = _sig("srk_disc_prep_w4", [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p])
_disc_wgrad4 = _sig("srk_disc_wgrad4", [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p])
lib.srk_disc_wgrad4_ws_floats.restype = c_longlong
lib.srk_disc_wgrad4_ws_floats.argtypes = [c_int, c_int, c_int]


def disc_prep_w4(w, a, at=None, sigma=None):
    """w [P,Q,4,4] fp32 -> a [P,16Q] bf16 (column (ky*4+kx)*Q + q) and at = a^T; sigma (1-element fp32 CUDA tensor or
    None): operands of W / sigma."""
    P, Q = w.shape[0], w.shape[1]
    assert w.dtype == torch.float32 and w.is_contiguous() and tuple(w.shape[2:]) == (4, 4)
    assert a.dtype == torch.bfloat16 and a.is_contiguous() and a.numel() == 16 * P * Q
    assert at is None or (at.dtype == torch.bfloat16 and at.is_contiguous() and at.numel() == 16 * P * Q)
    assert sigma is None or (sigma.dtype == torch.float32 and sigma.numel() == 1)
    _check(_disc_prep_w4(_ptr(w), P, Q, _ptr(sigma), _ptr(a), _ptr(at), _stream()), "srk_disc_prep_w4")


def disc_wgrad4(A, B, R, dw):
    """dw [Cb,R,4,4] fp32 = un-permuted A[T,16R]^T @ B[T,Cb]."""
    T, Cb = B.shape
    assert tuple(A.shape) == (T, 16 * R) and A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == 16 * R * Cb
    ws = _ws(int(lib.srk_disc_wgrad4_ws_floats(T, R, Cb)), A.device)
    _check(_disc_wgrad4(T, R, Cb, _ptr(A), _ld(A), _ptr(B), _ld(B), _ptr(ws), _ptr(dw), _stream()), "srk_disc_wgrad4")


class SrkSnLayer(Structure):
    _fields_ = [("w", c_void_p), ("u", c_void_p), ("v", c_void_p), ("A", c_int), ("B", c_int), ("KK", c_int), ("dim", c_int),
                ("sigma", c_void_p), ("w_sn", c_void_p)]


_spectral_norm = _sig("srk_spectral_norm", [POINTER(SrkSnLayer), c_int, c_int, c_float, c_void_p, c_void_p])
_spectral_norm_bwd = _sig("srk_spectral_norm_bwd", [POINTER(SrkSnLayer), c_int, POINTER(c_void_p), POINTER(c_void_p), c_void_p,
                                                    c_void_p])
lib.srk_spectral_norm_ws_floats.restype = c_longlong


def sn_layers(ws, us, vs, dims, sigmas, w_sn=None):
    """HOST descriptor array for srk_spectral_norm: ws / us / vs = weight_orig / weight_u / weight_v tensors, dims[i] = the
    `dim` spectral_norm used (0 Conv2d, 1 ConvTranspose2d), sigmas = fp32 CUDA tensor [n], w_sn[i] = optional fp32 output."""
    n = len(ws)
    arr = (SrkSnLayer * n)()
    for i, (w, u, v, dm) in enumerate(zip(ws, us, vs, dims)):
        assert w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 4 and u.dtype == torch.float32 and v.dtype == torch.float32
        A, B, KK = w.shape[0], w.shape[1], w.shape[2] * w.shape[3]
        assert u.is_contiguous() and v.is_contiguous() and u.numel() == (A if dm == 0 else B) and v.numel() == (B if dm == 0 else A) * KK
        o = None if w_sn is None else w_sn[i]
        assert o is None or (o.dtype == torch.float32 and o.is_contiguous() and o.numel() == w.numel())
        arr[i] = SrkSnLayer(_ptr(w), _ptr(u), _ptr(v), A, B, KK, dm, sigmas.data_ptr() + 4 * i, _ptr(o))
    return arr


def spectral_norm(layers, power_iteration: bool, eps: float, device):
    ws = _ws(int(lib.srk_spectral_norm_ws_floats()), device)
    _check(_spectral_norm(layers, len(layers), int(power_iteration), eps, _ptr(ws), _stream()), "srk_spectral_norm")


def spectral_norm_bwd(layers, dw_sn, dw, device):
    """dw_sn / dw: lists of fp32 CUDA tensors or None (layer skipped); dw[i] may be dw_sn[i]."""
    n = len(layers)
    a, b = (c_void_p * n)(), (c_void_p * n)()
    for i in range(n):
        a[i], b[i] = _ptr(dw_sn[i]), _ptr(dw[i])
    ws = _ws(int(lib.srk_spectral_norm_ws_floats()), device)
    _check(_spectral_norm_bwd(layers, n, a, b, _ptr(ws), _stream()), "srk_spectral_norm_bwd")


_bilinear2x_fwd = _sig("srk_bilinear2x_fwd", [POINTER(SrkView), POINTER(SrkView), POINTER(SrkView), c_int, c_int, c_int, c_void_p])
_bilinear2x_bwd = _sig("srk_bilinear2x_bwd", [POINTER(SrkView), POINTER(SrkView), c_int, c_int, c_int, c_void_p])


def bilinear2x_fwd(x: SrkView, s: SrkView | None, y: SrkView, B, H, W):
    """y [B,2H,2W,C] = bilinear x2 (align_corners=False) of (x + s) on NHWC views."""
    _check(_bilinear2x_fwd(_vref(x), _vref(s), _vref(y), B, H, W, _stream()), "srk_bilinear2x_fwd")


def bilinear2x_bwd(dy: SrkView, dx: SrkView, B, H, W):
    _check(_bilinear2x_bwd(_vref(dy), _vref(dx), B, H, W, _stream()), "srk_bilinear2x_bwd")
