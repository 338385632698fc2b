"""ctypes binding of libsrk.so (include/srk.h).  No torch types cross the boundary: only raw device
pointers, sizes and the current CUDA stream handle.  The library is required: importing this module
on a machine without the built .so raises, and every call raises on a non-zero return code.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_void_p
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "_lib" / "libsrk.so"


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
    ]


def _load() -> ctypes.CDLL:
    if not _LIB_PATH.exists():
        raise SrkError(
            f"{_LIB_PATH} is missing: build it with `python -m superresolution_def_b200._build` "
            "(there is no CPU or PyTorch fallback for the kernels)")
    return ctypes.CDLL(str(_LIB_PATH))


lib = _load()
lib.srk_version.restype = c_char_p


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
_gemm_wgrad_dbg = _sig("srk_gemm_wgrad_dbg", [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                               c_int, c_void_p, c_int, c_int, c_void_p])

EPI_STORE, EPI_GELU2, EPI_MUL, EPI_RES_LN, EPI_LNBWD = range(5)


def version() -> str:
    return lib.srk_version().decode()


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise SrkError(f"{what} failed with code {rc} (see stderr)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


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
    for t in (A, B, C, C2, X1, X2):
        assert t is None or t.dtype == torch.bfloat16
    rc = _gemm_tn(epi, M, N, K, _ptr(A), _ld(A), _ptr(B), _ld(B), _ptr(C), _ld(C), _ptr(C2), _ld(C2),
                  _ptr(X1), _ld(X1), _ptr(X2), _ld(X2), ctypes.byref(ln) if ln is not None else None, _stream())
    _check(rc, "srk_gemm_tn")


def make_ln_args(n_real, ones_col, gamma, beta=None, stats=None, partials=None, eps=1e-5) -> SrkLnArgs:
    for t in (gamma, beta, stats, partials):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    return SrkLnArgs(n_real, ones_col, _ptr(gamma), _ptr(beta), _ptr(stats), _ptr(partials), eps)


def wgrad_workspace_elems(Ca: int, Cb: int, splits: int) -> int:
    return splits * ((Ca + 127) // 128) * 128 * Cb


def gemm_wgrad(A, B, workspace, splits, out, dbg=None):
    """out[ceil128(Ca), Cb] (fp32) = A[T,Ca]^T @ B[T,Cb]."""
    T, Ca = A.shape
    Cb = B.shape[1]
    assert B.shape[0] == T and A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    assert workspace.dtype == torch.float32 and workspace.numel() >= wgrad_workspace_elems(Ca, Cb, splits)
    assert out.dtype == torch.float32 and out.numel() >= ((Ca + 127) // 128) * 128 * Cb
    if dbg is None:
        rc = _gemm_wgrad(T, Ca, Cb, _ptr(A), _ld(A), _ptr(B), _ld(B), _ptr(workspace), splits, _ptr(out), _stream())
    else:
        rc = _gemm_wgrad_dbg(T, Ca, Cb, _ptr(A), _ld(A), _ptr(B), _ld(B), _ptr(workspace), splits, _ptr(out),
                             dbg[0], dbg[1], _stream())
    _check(rc, "srk_gemm_wgrad")
