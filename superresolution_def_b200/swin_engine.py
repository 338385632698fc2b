"""Host-side engine for runs of Swin transformer blocks on libsrk (tcgen05 GEMMs + fused window attention).

The residual stream lives in HBM as token-major bf16 [T, Cp] (T = B*H*W tokens, Cp = 192 for C = 180: the 12 pad
channels are zero, and normalised copies carry 1.0 in column C so that every Linear bias is a weight column).
`SwinStackFunction` is one autograd node for a run of consecutive blocks: it owns the saved activations and calls
srk_swin_block_fwd / srk_swin_block_bwd once per block.  PyTorch is used for memory, streams and autograd plumbing
only — all arithmetic happens in the library, and there is no fallback path.

Reference being replaced: SwinTransformerBlock.forward / WindowAttention.forward / Mlp.forward
(models/architecture_swin.py:123-151, :71-96, :19-25) and their autograd backward.
"""
from __future__ import annotations

from dataclasses import dataclass

import os

import torch

from . import _capi as capi

BF16 = torch.bfloat16
BLOCK_PARAM_KEYS = ("norm1.weight", "norm1.bias", "attn.relative_position_bias_table", "attn.qkv.weight",
                    "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias", "norm2.weight", "norm2.bias",
                    "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")
N_BLOCK_PARAMS = len(BLOCK_PARAM_KEYS)


@dataclass(frozen=True)
class BlockCfg:
    C: int = 180
    heads: int = 6
    hidden: int = 720
    ws: int = 8
    Cp: int = 192
    ds: int = 32
    Hp: int = 768

    @property
    def dh(self) -> int:
        return self.C // self.heads

    @property
    def QW(self) -> int:
        return 3 * self.heads * self.ds

    @property
    def AW(self) -> int:
        return self.heads * self.ds

    def dims(self) -> capi.SrkBlockDims:
        return capi.SrkBlockDims(self.C, self.Cp, self.heads, self.dh, self.ds, self.hidden, self.Hp)

    @staticmethod
    def for_model(C: int, heads: int, hidden: int, ws: int) -> "BlockCfg":
        if C % heads != 0:
            raise capi.SrkError(f"embed_dim {C} not divisible by heads {heads}")
        cfg = BlockCfg(C=C, heads=heads, hidden=hidden, ws=ws, Cp=192, ds=32, Hp=((hidden + 1 + 255) // 256) * 256)
        if not (C < 192 and heads * 32 == 192 and C // heads < 32 and ws == 8):
            raise capi.SrkError(
                f"libsrk block kernels are specialised for heads=6, head_dim<32, embed_dim<192, window 8; got "
                f"C={C} heads={heads} ws={ws}")
        return cfg


_BLOCK_PARAM_PATHS = tuple(tuple(k.split(".")) for k in BLOCK_PARAM_KEYS)


def block_params_of(block: torch.nn.Module) -> list[torch.Tensor]:
    """The 13 parameters of a SwinTransformerBlock-shaped module, in the C ABI's order, resolved by attribute path
    (`dict(block.named_parameters())` walks the module tree in Python on every forward of every block).  Measured with
    tools/gpu_swinir_hostprof.py: the eager step at micro-batch 2 is device-bound (12.5 ms of kernels behind 11.6 ms of
    host time), so this only matters for the launch queue's slack, not for the step time."""
    out = []
    for path in _BLOCK_PARAM_PATHS:
        o = block
        for a in path:
            o = getattr(o, a)
        out.append(o)
    return out


class _WeightCache:
    """bf16 GEMM operands of one block.  They are re-derived from the fp32 master parameters on every forward
    (one small kernel per block): fused / foreach optimizers update parameters without bumping the autograd version
    counter, so no host-side staleness test is reliable.  `freeze_weights(True)` opts into reuse for pure inference."""

    def __init__(self, cfg: BlockCfg, device):
        self.cfg = cfg
        elems = capi.block_weight_elems(cfg.dims())
        self.t = {n: torch.empty(e, device=device, dtype=BF16) for n, e in zip(capi.WEIGHT_NAMES, elems)}
        self.ready = False
        self.epoch = -1

    def get(self, params: list[torch.Tensor], refresh: bool) -> dict:
        if refresh or not self.ready or not _frozen or self.epoch != _epoch:
            capi.block_prep_weights(self.cfg.dims(), dict(zip(capi.PARAM_NAMES, params)), self.t)
            self.ready = True
            self.epoch = _epoch
        return self.t


_weight_caches: dict = {}
_frozen = False
_epoch = 0   # bumped whenever a mirror module is built or loads a state_dict: frozen operands older than that are stale


def weights_changed(*_a, **_k):
    """Invalidate every frozen bf16 operand copy.  The caches are keyed by parameter address, and an address can be
    recycled by a model built after another one was freed, so the mirrors call this from __init__ and from their
    load_state_dict post-hook; training never relies on it (operands are re-derived in every forward)."""
    global _epoch
    _epoch += 1


def track_weight_changes(module: torch.nn.Module):
    weights_changed()
    module.register_load_state_dict_post_hook(weights_changed)


def freeze_weights(flag: bool = True):
    """Inference-only opt-in: keep the prepared bf16 operands between forwards (call again with False, or after
    loading new weights, to drop them)."""
    global _frozen
    _frozen = bool(flag)
    for wc in _weight_caches.values():
        wc.ready = False


def _weights_for(cfg: BlockCfg, params: list[torch.Tensor], refresh: bool = True) -> dict:
    k = (cfg, params[3].data_ptr())  # keyed by the qkv weight storage
    wc = _weight_caches.get(k)
    if wc is None:
        wc = _weight_caches[k] = _WeightCache(cfg, params[3].device)
    return wc.get(params, refresh)


_fp32_warned = False


def check_precision(x: torch.Tensor) -> None:
    """Precision policy of the mirrors.  Every libsrk kernel computes in bf16 with fp32 accumulation, which is what the
    reference gets under the autocast of train_swin.py:217-243.  train_hat.py:222-251 however trains in plain fp32 (no
    autocast): running it on the mirrors silently changes the arithmetic, so the downgrade is made loud and controllable:
      SRK_FP32_POLICY=warn  (default) one RuntimeWarning per process when an fp32 input arrives outside autocast;
      SRK_FP32_POLICY=error refuse (SrkError) — for runs that must not leave fp32;
      SRK_FP32_POLICY=allow silent.
    The parity this repository asserts for that case is stated in tests/test_fullsize_gpu.py: outputs and every parameter
    gradient against the fp32 oracle, bounded relative to the oracle's own bf16-autocast error."""
    global _fp32_warned
    if x.dtype != torch.float32 or torch.is_autocast_enabled(x.device.type):
        return
    mode = os.environ.get("SRK_FP32_POLICY", "warn").lower()
    if mode == "error":
        raise capi.SrkError("fp32 input outside autocast and SRK_FP32_POLICY=error: libsrk kernels compute in bf16 "
                            "(fp32 accumulate); wrap the call in torch.autocast or set SRK_FP32_POLICY=warn/allow")
    if mode == "warn" and not _fp32_warned:
        _fp32_warned = True
        import warnings
        warnings.warn("superresolution_def_b200: fp32 input outside autocast — the sm_100a kernels compute in bf16 with fp32 "
                      "accumulation (the reference's own autocast precision); outputs keep the caller's dtype. "
                      "Set SRK_FP32_POLICY=error to refuse, =allow to silence.", RuntimeWarning, stacklevel=3)


def _check_param(p: torch.Tensor):
    if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
        raise capi.SrkError("libsrk needs contiguous fp32 CUDA parameters (no CPU / fallback path exists)")


STORE_GELU_GRAD = bool(int(os.environ.get("SRK_STORE_DACT", "1")))


def _alloc_acts(cfg: BlockCfg, T: int, device) -> dict:
    """Activations a block saves for its backward.  gelu'(u) ("dact", [T, Hp]) is stored by default.  SRK_STORE_DACT=0
    drops it: the backward then recomputes u = xn2 W1^T inside the fc2 input-gradient kernel (SRK_EPI_MULG, bit-identical
    dU).  Measured on B200 (tools/gpu_probe_mlp.py): forward 200 -> 144 us, backward 194 -> 265 us per block — the
    epilogues are instruction-issue bound, so recomputing the GELU polynomial costs more than the 0.8 GB of traffic it
    saves; the variant is kept as the memory-saving mode (-14.5 GB at batch 16), not as the default."""
    e = lambda w: torch.empty(T, w, device=device, dtype=BF16)  # noqa: E731
    acts = {"qkv": e(cfg.QW), "ao": e(cfg.AW), "x_mid": e(cfg.Cp), "xn2": e(cfg.Cp),
            "stats2": torch.empty(T, 2, device=device, dtype=torch.float32), "act": e(cfg.Hp),
            "x_out": e(cfg.Cp), "xn_out": e(cfg.Cp),
            "stats_out": torch.empty(T, 2, device=device, dtype=torch.float32)}
    if STORE_GELU_GRAD:
        acts["dact"] = torch.empty(T, cfg.Hp, device=device, dtype=torch.float16)   # gelu'(u): fp16, read only by EPI_MUL
    return acts


def layernorm_tokens(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, C: int):
    """xn, stats = LN(x[:, :C]) with the ones column at C.  x: [T, Cp] bf16."""
    xn = torch.empty_like(x)
    stats = torch.empty(x.shape[0], 2, device=x.device, dtype=torch.float32)
    capi.layernorm_fwd(x, xn, stats, weight.detach(), bias.detach(), C, ones_col=C)
    return xn, stats


class SwinStackFunction(torch.autograd.Function):
    """(x, xn, stats) -> (x_out, xn_out, stats_out) through `len(shifts)` consecutive Swin blocks.

    tensors = 13 params per block (BLOCK_PARAM_KEYS order) followed by (next_norm_weight, next_norm_bias): the
    affine of the LayerNorm that consumes the stack's output (its gradient belongs to the consumer, not to us).
    The gradient returned for `x` is the complete dL/dx (residual path + LayerNorm-1 path); `xn`/`stats` are
    derived data and carry no gradient of their own.
    """

    @staticmethod
    def forward(ctx, x, xn, stats, cfg: BlockCfg, geom: tuple, shifts: tuple, *tensors):
        B, H, W = geom
        T = B * H * W
        nb = len(shifts)
        assert len(tensors) == nb * N_BLOCK_PARAMS + 2
        for t in tensors:
            _check_param(t)
        assert x.dtype == BF16 and x.shape == (T, cfg.Cp) and x.is_contiguous() and xn.is_contiguous()
        need_grad = any(ctx.needs_input_grad)  # grad mode is always off inside Function.forward
        dims = cfg.dims()
        saved, saved_w = [], []
        cur_x, cur_xn, cur_stats = x, xn, stats
        # inference: two activation sets are ping-ponged (block i reads set i-1's outputs while writing set i)
        pingpong = None if need_grad else [_alloc_acts(cfg, T, x.device) for _ in range(min(nb, 2))]
        for i in range(nb):
            params = [t.detach() for t in tensors[i * N_BLOCK_PARAMS:(i + 1) * N_BLOCK_PARAMS]]
            if i + 1 < nb:
                nw, nbias = tensors[(i + 1) * N_BLOCK_PARAMS].detach(), tensors[(i + 1) * N_BLOCK_PARAMS + 1].detach()
            else:
                nw, nbias = tensors[-2].detach(), tensors[-1].detach()
            weights = _weights_for(cfg, params, refresh=need_grad)
            acts = dict(_alloc_acts(cfg, T, x.device) if need_grad else pingpong[i & 1])
            acts.update(x_in=cur_x, xn1=cur_xn, stats1=cur_stats)
            g = capi.SrkGeom(B, H, W, cfg.ws, shifts[i])
            capi.swin_block_fwd(dims, g, weights, dict(zip(capi.PARAM_NAMES, params)), nw, nbias, acts)
            if need_grad:
                saved.append(acts)
                saved_w.append(weights)
            cur_x, cur_xn, cur_stats = acts["x_out"], acts["xn_out"], acts["stats_out"]
        if saved:
            # the node's own outputs are not needed by its backward; keeping them on ctx would close an
            # output -> grad_fn -> ctx -> output cycle that only the cyclic GC could free (tens of GB per abandoned graph)
            saved[-1] = {k: v for k, v in saved[-1].items() if k not in ("x_out", "xn_out", "stats_out")}
        ctx.cfg, ctx.geom, ctx.shifts, ctx.saved_acts, ctx.saved_w = cfg, geom, shifts, saved, saved_w
        ctx.params = tensors
        ctx.mark_non_differentiable(cur_xn, cur_stats)
        ctx.set_materialize_grads(False)
        return cur_x, cur_xn, cur_stats

    @staticmethod
    def backward(ctx, g_x, _g_xn, _g_stats):
        cfg, (B, H, W), shifts, tensors = ctx.cfg, ctx.geom, ctx.shifts, ctx.params
        T = B * H * W
        if g_x is None:   # set_materialize_grads(False): the stack's output did not reach the loss
            return (None,) * (6 + len(tensors))
        if ctx.saved_acts and ctx.saved_acts[-1] is None:
            raise capi.SrkError("SwinStackFunction.backward ran twice: the saved activations are released block by block "
                                "during the first backward (retain_graph / double backward is not supported)")
        dev = g_x.device
        dims = cfg.dims()
        nb = len(shifts)
        g = g_x.contiguous()
        if g.dtype != BF16:
            g = g.to(BF16)
        scratch = _bwd_scratch(cfg, B, H, W, dev)
        grads: list = [None] * len(tensors)
        bufs = [torch.empty(T, cfg.Cp, device=dev, dtype=BF16), torch.empty(T, cfg.Cp, device=dev, dtype=BF16)]
        for i in reversed(range(nb)):
            params = [t.detach() for t in tensors[i * N_BLOCK_PARAMS:(i + 1) * N_BLOCK_PARAMS]]
            weights = ctx.saved_w[i]  # the operand buffers this step's forward prepared
            acts = ctx.saved_acts[i]
            gdict = {n: torch.empty_like(p) for n, p in zip(capi.PARAM_NAMES, params)}
            g_in = bufs[i & 1]
            geom = capi.SrkGeom(B, H, W, cfg.ws, shifts[i])
            capi.swin_block_bwd(dims, geom, weights, dict(zip(capi.PARAM_NAMES, params)), acts, g, scratch, g_in, gdict)
            for j, n in enumerate(capi.PARAM_NAMES):
                grads[i * N_BLOCK_PARAMS + j] = gdict[n]
            ctx.saved_acts[i] = None  # release this block's activations as soon as they are consumed
            g = g_in
        return (g, None, None, None, None, None, *grads)


_scratch_cache: dict = {}


def _bwd_scratch(cfg: BlockCfg, B: int, H: int, W: int, device) -> dict:
    key = (cfg, B, H, W, str(device))
    s = _scratch_cache.get(key)
    if s is None:
        T = B * H * W
        n = capi.block_bwd_scratch_floats(cfg.dims(), capi.SrkGeom(B, H, W, cfg.ws, 0))
        s = {"d_act": torch.empty(T, cfg.Hp, device=device, dtype=BF16),
             "d_ao": torch.empty(T, cfg.AW, device=device, dtype=BF16),
             "d_qkv": torch.empty(T, cfg.QW, device=device, dtype=BF16),
             "g_mid": torch.empty(T, cfg.Cp, device=device, dtype=BF16),
             "wg_ws": torch.empty(n, device=device, dtype=torch.float32)}
        # one entry per geometry, never evicted: a captured CUDA graph (graphs.GraphedStep) has these addresses baked
        # into its kernel nodes, so a validation pass at another batch size must not free the training step's scratch
        _scratch_cache[key] = s
    return s


class FusedNormOutput(torch.autograd.Function):
    """Identity on the already-computed xn = LayerNorm(x) that makes it differentiable w.r.t. x, weight and bias.
    The forward value was produced by the epilogue of the last block's fc2 GEMM (SwinIR.norm, :247)."""

    @staticmethod
    def forward(ctx, x, xn, stats, weight, bias, C: int):
        ctx.save_for_backward(x, stats, weight)
        ctx.C = C
        return xn.detach()

    @staticmethod
    def backward(ctx, g_xn):
        x, stats, weight = ctx.saved_tensors
        g_xn = g_xn.contiguous().to(BF16)
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(weight)
        dbeta = torch.empty_like(weight)
        capi.layernorm_bwd(g_xn, x, stats, weight.detach(), None, dx, dgamma, dbeta, ctx.C)
        return dx, None, None, dgamma, dbeta, None


def pack_tokens(t: torch.Tensor, Cp: int) -> torch.Tensor:
    """(B, L, C) any float dtype -> [B*L, Cp] bf16 with zero pad channels (plumbing; differentiable)."""
    B, L, C = t.shape
    out = t.new_zeros((B * L, Cp), dtype=BF16)
    out[:, :C] = t.reshape(B * L, C).to(BF16)
    return out


def unpack_tokens(t: torch.Tensor, B: int, C: int, dtype) -> torch.Tensor:
    return t[:, :C].reshape(B, -1, C).to(dtype)


_ident_cache: dict = {}


def identity_norm(C: int, device):
    """(ones, zeros) affine used when a stack's output is not consumed by a LayerNorm (stand-alone block)."""
    k = (C, str(device))
    if k not in _ident_cache:
        _ident_cache[k] = (torch.ones(C, device=device), torch.zeros(C, device=device))
    return _ident_cache[k]


# ---------------------------------------------------------------------------------------------------
# Stand-alone module forwards (WindowAttention.forward / Mlp.forward called outside a block).
# Same kernels as the fused block path; only the packing / gradient unpacking is done with torch indexing.
# ---------------------------------------------------------------------------------------------------
_zero_cache: dict = {}


def _zeros(shape, device):
    k = (tuple(shape), str(device))
    if k not in _zero_cache:
        _zero_cache[k] = torch.zeros(shape, device=device)
    return _zero_cache[k]


def _partial_weights(cfg: BlockCfg, device, **given) -> dict:
    """Prepared operands when only some of a block's parameters exist (missing ones are zeros)."""
    C, hid, T2 = cfg.C, cfg.hidden, (2 * cfg.ws - 1) ** 2
    shapes = {"norm1_w": (C,), "norm1_b": (C,), "rpb_table": (T2, cfg.heads), "qkv_w": (3 * C, C), "qkv_b": (3 * C,),
              "proj_w": (C, C), "proj_b": (C,), "norm2_w": (C,), "norm2_b": (C,), "fc1_w": (hid, C), "fc1_b": (hid,),
              "fc2_w": (C, hid), "fc2_b": (C,)}
    params = {n: (given[n].detach() if n in given else _zeros(s, device)) for n, s in shapes.items()}
    for p in params.values():
        _check_param(p)
    elems = capi.block_weight_elems(cfg.dims())
    w = {n: torch.empty(e, device=device, dtype=BF16) for n, e in zip(capi.WEIGHT_NAMES, elems)}
    capi.block_prep_weights(cfg.dims(), params, w)
    return w


def _pack_rows(x2d: torch.Tensor, Cp: int, ones_col: int, rows_pad: int) -> torch.Tensor:
    T, C = x2d.shape
    out = x2d.new_zeros((rows_pad, Cp), dtype=BF16)
    out[:T, :C] = x2d.to(BF16)
    if ones_col >= 0:
        out[:T, ones_col] = 1.0
    return out


def _wgrad(A, B, rows_a):
    """fp32 [ceil128(Ca), Cb] = A^T @ B (tcgen05, MN-major operands)."""
    T, Ca = A.shape
    Cb = B.shape[1]
    tiles = (Ca + 127) // 128
    splits = capi.wgrad_splits(T, Ca)
    ws = torch.empty(capi.wgrad_workspace_elems(Ca, Cb, splits), device=A.device, dtype=torch.float32)
    out = torch.empty(tiles * 128, Cb, device=A.device, dtype=torch.float32)
    capi.gemm_wgrad(A, B, ws, splits, out)
    return out[:rows_a]


class MlpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fc1_w, fc1_b, fc2_w, fc2_b):
        C, hid = fc1_w.shape[1], fc1_w.shape[0]
        cfg = BlockCfg.for_model(C, 6, hid, 8)
        lead = x.shape[:-1]
        x2 = x.reshape(-1, C)
        T = x2.shape[0]
        Tp = (T + 127) // 128 * 128
        w = _partial_weights(cfg, x.device, fc1_w=fc1_w, fc1_b=fc1_b, fc2_w=fc2_w, fc2_b=fc2_b)
        xp = _pack_rows(x2, cfg.Cp, C, Tp)
        act = torch.empty(Tp, cfg.Hp, device=x.device, dtype=BF16)
        dact = torch.empty_like(act, dtype=torch.float16)   # gelu'(u) is stored as fp16 (EPI_GELU2 / EPI_MUL contract)
        capi.gemm_tn(capi.EPI_GELU2, xp, w["fc1_f"].view(cfg.Hp, cfg.Cp), act, C2=dact,
                     ln=capi.make_ln_args(cfg.Hp, hid, None))
        y = torch.empty(Tp, cfg.Cp, device=x.device, dtype=BF16)
        capi.gemm_tn(capi.EPI_STORE, act, w["fc2_f"].view(cfg.Cp, cfg.Hp), y)
        ctx.saved = (xp, act, dact, w, cfg, T, lead, x.dtype)
        return y[:T, :C].reshape(*lead, C).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        xp, act, dact, w, cfg, T, lead, dtype = ctx.saved
        C, hid = cfg.C, cfg.hidden
        Tp = xp.shape[0]
        dyp = _pack_rows(dy.reshape(-1, C), cfg.Cp, -1, Tp)
        dU = torch.empty(Tp, cfg.Hp, device=dy.device, dtype=BF16)
        capi.gemm_tn(capi.EPI_MUL, dyp, w["fc2_t"].view(cfg.Hp, cfg.Cp), dU, X1=dact)
        dx = torch.empty(Tp, cfg.Cp, device=dy.device, dtype=BF16)
        capi.gemm_tn(capi.EPI_STORE, dU, w["fc1_t"].view(cfg.Cp, cfg.Hp), dx)
        e2 = _wgrad(act, dyp, cfg.Hp)     # [Hp, Cp]: rows hidden (+ones row), cols out channels
        e1 = _wgrad(dU, xp, cfg.Hp)       # [Hp, Cp]: rows hidden, cols in channels (+ones col)
        return (dx[:T, :C].reshape(*lead, C).to(dtype), e1[:hid, :C].contiguous(), e1[:hid, C].contiguous(),
                e2[:hid, :C].t().contiguous(), e2[hid, :C].contiguous())


def mlp_forward(x, fc1_w, fc1_b, fc2_w, fc2_b):
    if not x.is_cuda:
        raise capi.SrkError("libsrk Mlp runs on CUDA only")
    return MlpFunction.apply(x, fc1_w, fc1_b, fc2_w, fc2_b)


class WindowAttentionFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ws, heads, table, qkv_w, qkv_b, proj_w, proj_b):
        B_, N, C = x.shape
        if (ws, N) not in ((8, 64), (16, 256)):
            raise capi.SrkError("stand-alone WindowAttention kernels: windows 8x8 (64 tokens) or 16x16 (256 tokens)")
        if ws == 8:
            cfg = BlockCfg.for_model(C, heads, 4 * C, ws)
        else:  # HAT's WindowAttention (hat_arch.py:129-196), mask=None
            if not (C < 192 and heads * 32 == 192 and C % heads == 0):
                raise capi.SrkError("libsrk window attention: heads=6, head_dim<32, embed_dim<192")
            cfg = BlockCfg(C=C, heads=heads, hidden=4 * C, ws=16, Cp=192, ds=32, Hp=((4 * C + 1 + 255) // 256) * 256)
        Bp = B_ + (B_ & 1) if ws == 8 else B_  # 128-row GEMM tiles
        w = _partial_weights(cfg, x.device, rpb_table=table, qkv_w=qkv_w, qkv_b=qkv_b, proj_w=proj_w, proj_b=proj_b)
        xp = _pack_rows(x.reshape(-1, C), cfg.Cp, C, Bp * N)
        qkv = torch.empty(Bp * N, cfg.QW, device=x.device, dtype=BF16)
        capi.gemm_tn(capi.EPI_STORE, xp, w["qkv_f"].view(cfg.QW, cfg.Cp), qkv)
        geom = capi.SrkGeom(Bp, ws, ws, ws, 0)  # already-partitioned windows: one ws x ws "image" per window
        ao = torch.empty(Bp * N, cfg.AW, device=x.device, dtype=BF16)
        lse = None
        if ws == 8:
            capi.win_attn_fwd(geom, heads, qkv, table.detach(), ao, ones_col=cfg.dh)
        else:
            lse = torch.empty(heads, Bp * N, device=x.device, dtype=torch.float32)
            capi.win_attn16_fwd(geom, capi.ATTN_SELF, heads, qkv, table.detach(), ao, lse, ones_col=cfg.dh)
        y = torch.empty(Bp * N, cfg.Cp, device=x.device, dtype=BF16)
        capi.gemm_tn(capi.EPI_STORE, ao, w["proj_f"].view(cfg.Cp, cfg.AW), y)
        ctx.saved = (xp, qkv, ao, w, cfg, table.detach(), B_, Bp, x.dtype, N, lse)
        return y[:B_ * N, :C].reshape(B_, N, C).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        xp, qkv, ao, w, cfg, table, B_, Bp, dtype, N, lse = ctx.saved
        C, heads, dh, ds = cfg.C, cfg.heads, cfg.dh, cfg.ds
        dev = dy.device
        dyp = _pack_rows(dy.reshape(-1, C), cfg.Cp, -1, Bp * N)
        d_ao = torch.empty(Bp * N, cfg.AW, device=dev, dtype=BF16)
        capi.gemm_tn(capi.EPI_STORE, dyp, w["proj_t"].view(cfg.AW, cfg.Cp), d_ao)
        d_qkv = torch.empty_like(qkv)
        d_table = torch.empty_like(table)
        geom = capi.SrkGeom(Bp, cfg.ws, cfg.ws, cfg.ws, 0)
        if cfg.ws == 8:
            capi.win_attn_bwd(geom, heads, qkv, table, d_ao, d_qkv, d_table)
        else:
            aws = torch.empty(capi.attn16_bwd_ws_bytes(geom, capi.ATTN_SELF, heads), device=dev, dtype=torch.uint8)
            capi.win_attn16_bwd(geom, capi.ATTN_SELF, heads, qkv, table, ao, d_ao, lse, d_qkv, aws, d_table)
        dx = torch.empty(Bp * N, cfg.Cp, device=dev, dtype=BF16)
        capi.gemm_tn(capi.EPI_STORE, d_qkv, w["qkv_t"].view(cfg.Cp, cfg.QW), dx)
        ep = _wgrad(dyp, ao, cfg.Cp)       # [Cp, AW]
        eq = _wgrad(d_qkv, xp, cfg.QW)     # [QW, Cp]
        scale = torch.ones(3, 1, 1, 1, device=dev)
        scale[0] = dh ** -0.5               # q rows were pre-scaled in the forward operand
        eq = eq.view(3, heads, ds, cfg.Cp)[:, :, :dh] * scale
        d_qkv_w = eq[..., :C].reshape(3 * C, C).contiguous()
        d_qkv_b = eq[..., C].reshape(3 * C).contiguous()
        epv = ep[:C].view(C, heads, ds)
        d_proj_w = epv[:, :, :dh].reshape(C, C).contiguous()
        d_proj_b = epv[:, 0, dh].contiguous()
        return (dx[:B_ * N, :C].reshape(B_, N, C).to(dtype), None, None, d_table, d_qkv_w, d_qkv_b, d_proj_w,
                d_proj_b)


def window_attention_forward(x, ws, heads, table, qkv_w, qkv_b, proj_w, proj_b):
    if not x.is_cuda:
        raise capi.SrkError("libsrk WindowAttention runs on CUDA only")
    return WindowAttentionFunction.apply(x, ws, heads, table, qkv_w, qkv_b, proj_w, proj_b)
