#!/usr/bin/env python
"""bench.py — SwinIR x4 training throughput (patches/s, 128^2 -> 512^2, bf16, batch 16/GPU) on B200.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line on rank 0.
  * our arm      : superresolution_def_b200.architecture_swin.SwinIR (libsrk kernels) — one training step =
                   forward, L1 loss, backward, (bucketed NCCL all-reduce if N>1), AdamW step; workload = BASELINE
                   configs[1].  `value` is device-timed with inputs resident in HBM; `e2e` repeats the measurement
                   with pinned-host inputs copied H2D and the loss read back D2H inside the timed region.
  * reference arm: the reference's own CPU path restated in oracle/ (the reference is pure PyTorch and cannot
                   travel to the GPU box), fp32, all host threads, bounded sample, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SwinIR x4 train patches/s (128x128 -> 512x512)"
UNIT = "patches/s"
MODEL_KW = dict(upscale=4, in_chans=1, img_size=128, window_size=8, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6,
                mlp_ratio=2)  # train_swin.py:147-149 (mlp_ratio is swallowed by the reference: effective 4.0)
GFLOP_PER_PATCH_TRAIN = 1569.826  # SURVEY.md §8d (FlopCounterMode on the reference, fwd+bwd)
# --workload hat: BASELINE configs[2] (a parity-test configuration; timed on request, not the default bench line)
HAT_KW = dict(img_size=128, in_chans=1, embed_dim=180, depths=(6,) * 6, num_heads=(6,) * 6, window_size=16, upscale=4,
              upsampler="pixelshuffle")   # drop_path_rate default 0.1: stochastic depth active, as train_hat.py would
HAT_GFLOP_PER_PATCH_TRAIN = 3026.031
# --workload hybrid: the generator train_hat.py actually trains (train_hat.py:132-136; SURVEY.md section 8a row a17, cfgH)
HYBRID_KW = dict(img_size=128, in_chans=1, embed_dim=90, depths=(6, 6, 6, 6), num_heads=(6, 6, 6, 6), window_size=8,
                 upscale=4, num_rrdb=12, num_feat=48, num_grow_ch=24)
HYBRID_GFLOP_PER_PATCH_TRAIN = 3 * 819.171


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), float(p["bf16_tflops_sustained"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_arm(steps: int, warmup: int, threads: int | None = None):
    """The reference's CPU path (oracle port): SwinIR fp32 training step, batch 1 (a bounded sample of the
    batch-16 workload), L1 loss + AdamW(1e-4, betas (0.9, 0.99)) as train_swin.py:160."""
    from oracle import swinir_oracle as o
    from superresolution_def_b200.synth import synthetic_pairs
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = o.init_state_dict(**{k: MODEL_KW[k] for k in ("img_size", "window_size", "embed_dim", "depths", "num_heads")})
    params = [v.requires_grad_(True) for v in sd.values() if v.is_floating_point()]
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.99))
    lr, hr = synthetic_pairs(1, seed=1234)
    kw = dict(img_size=128, window_size=8, depths=MODEL_KW["depths"], num_heads=MODEL_KW["num_heads"], upscale=4)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.l1_loss(o.swinir_forward(lr, sd, **kw), hr)
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps, threads


def gpu_eager_baseline(batch: int, steps: int = 3, warmup: int = 2):
    """The GPU bar: stock PyTorch eager — the oracle's ATen call sequence, i.e. exactly what the
    reference modules execute — under torch.autocast(bf16) on this same B200, same training step (fwd + L1 + bwd + fused
    AdamW), CUDA-event timed.  A reported baseline like cpu_baseline: it never touches the product path.  The largest of
    (batch, batch/2, batch/4) that fits is used and stated."""
    from oracle import swinir_oracle as o
    from superresolution_def_b200.synth import synthetic_pairs
    dev = torch.device("cuda", torch.cuda.current_device())
    kw = dict(img_size=128, window_size=8, depths=MODEL_KW["depths"], num_heads=MODEL_KW["num_heads"], upscale=4)
    for b in (batch, max(1, batch // 2), max(1, batch // 4)):
        try:
            torch.manual_seed(0)
            sd = o.init_state_dict(**{k: MODEL_KW[k] for k in ("img_size", "window_size", "embed_dim", "depths", "num_heads")})
            sd = {k: v.to(dev) for k, v in sd.items()}
            params = [v.requires_grad_(True) for v in sd.values() if v.is_floating_point()]
            opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.99), fused=True)
            lr, hr = synthetic_pairs(min(b, 4), seed=99)
            reps = (b + lr.shape[0] - 1) // lr.shape[0]
            lr, hr = lr.repeat(reps, 1, 1, 1)[:b].to(dev), hr.repeat(reps, 1, 1, 1)[:b].to(dev)

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    sr = o.swinir_forward(lr, sd, **kw)
                loss = torch.nn.functional.l1_loss(sr.float(), hr)
                loss.backward()
                opt.step()
                return loss

            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / steps
            mem = torch.cuda.max_memory_allocated() / 2 ** 30
            del sd, params, opt
            torch.cuda.empty_cache()
            return {"value": b / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": b, "dtype": "bf16 autocast",
                    "impl": "oracle/swinir_oracle.py (the reference's ATen sequence: cuBLAS / cuDNN / eager elementwise), "
                            "eager, fused AdamW", "steps": steps, "warmup": warmup, "peak_mem_gb": mem}
        except torch.OutOfMemoryError:
            torch.cuda.empty_cache()
            continue
    return {"unavailable": "out of memory at every tried batch size"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    v, spp, threads = cpu_reference_arm(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SwinIR x4 training step, L1, 128^2->512^2 (BASELINE configs[1]); CPU sample: batch 1"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{steps} fp32 training step(s) of batch 1 after {warmup} warm-up, oracle/swinir_oracle.py"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def roofline_probe(batch: int, peaks):
    """Times every per-block kernel of the SwinIR step alone at the step's shapes (CUDA events on the launching stream,
    operands larger than L2) and reports each against the HBM roofline: achieved = algorithmic bytes per launch /
    mean launch duration.  Algorithmic bytes = each operand / result tensor moved once.  The
    `roofline` object is the kernel function with the largest share of the step (launches x duration); the full table
    goes to `roofline_kernels`."""
    from superresolution_def_b200 import _capi as capi
    hbm, _, _, kind = peaks
    M, C, QW, HP, heads = batch * 16384, 192, 576, 768, 6
    bf = torch.bfloat16
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *shape: torch.randn(*shape, device=dev, generator=g).to(bf)  # noqa: E731
    x192, y192, z192, o192 = rnd(M, C), rnd(M, C), rnd(M, C), torch.empty(M, C, device=dev, dtype=bf)
    x576, o576 = rnd(M, QW), torch.empty(M, QW, device=dev, dtype=bf)
    x768, o768 = rnd(M, HP), torch.empty(M, HP, device=dev, dtype=bf)
    # gelu' is stored as fp16 (second output of GELU2, multiplier of MUL)
    y768, o768b = torch.randn(M, HP, device=dev, generator=g).to(torch.float16), torch.empty(M, HP, device=dev, dtype=torch.float16)
    w = lambda n, k: (torch.randn(n, k, device=dev, generator=g) / k ** 0.5).to(bf)  # noqa: E731
    w_qkv, w_proj, w_fc1, w_fc2, w_fc1t, w_qkvt = w(QW, C), w(C, C), w(HP, C), w(C, HP), w(C, HP), w(C, QW)
    stats = torch.empty(M, 2, device=dev)
    gam, bet = torch.ones(180, device=dev), torch.zeros(180, device=dev)
    table = torch.randn(225, heads, device=dev, generator=g)
    geom = capi.SrkGeom(batch, 128, 128, 8, 4)
    parts = torch.empty(capi.gemm_grid(M, C) * 2 * C, device=dev)
    capi.layernorm_fwd(x192, o192, stats, gam, bet, 180, ones_col=180)  # valid (mean, rstd) for the LNBWD runs
    wg_ws = torch.empty(148 * 256 * 256, device=dev)
    wg_out = torch.empty(768 * 256, device=dev)
    dtab = torch.empty_like(table)
    ln_res = capi.make_ln_args(180, 180, gam, bet, stats=stats)
    ln_bwd = capi.make_ln_args(180, -1, gam, None, stats=stats, partials=parts)
    ln_gelu = capi.make_ln_args(HP, 720, None)
    E = 2  # bytes per element

    def splits(T, ca):
        return capi.wgrad_splits(T, ca)

    cases = [
        ("gemm_tn<192,STORE> qkv", "gemm_tn_kernel<STORE>", 1, lambda: capi.gemm_tn(capi.EPI_STORE, x192, w_qkv, o576), M * (C + QW) * E),
        ("win_attn_ws8_fwd", "win_attn_ws8_fwd_kernel", 1, lambda: capi.win_attn_fwd(geom, heads, x576, table, o192, ones_col=30), M * (QW + C) * E),
        ("gemm_tn<192,RES_LN> proj", "gemm_tn_kernel<RES_LN>", 1, lambda: capi.gemm_tn(capi.EPI_RES_LN, x192, w_proj, o192, C2=z192, X1=y192, ln=ln_res), M * 4 * C * E),
        ("gemm_tn<256,GELU2> fc1", "gemm_tn_kernel<GELU2>", 1, lambda: capi.gemm_tn(capi.EPI_GELU2, x192, w_fc1, o768, C2=o768b, ln=ln_gelu), M * (C + 2 * HP) * E),
        ("gemm_tn<192,RES_LN> fc2", "gemm_tn_kernel<RES_LN>", 1, lambda: capi.gemm_tn(capi.EPI_RES_LN, x768, w_fc2, o192, C2=z192, X1=y192, ln=ln_res), M * (HP + 3 * C) * E),
        ("gemm_tn<192,MUL> fc2 dgrad", "gemm_tn_kernel<MUL>", 1, lambda: capi.gemm_tn(capi.EPI_MUL, x192, w_fc1, o768, X1=y768), M * (C + 2 * HP) * E),
        ("gemm_wgrad<192> fc1/fc2", "gemm_wgrad_kernel", 2, lambda: capi.gemm_wgrad(x768, x192, wg_ws, splits(M, HP), wg_out), M * (HP + C) * E),
        ("gemm_tn<192,LNBWD> fc1 dgrad", "gemm_tn_kernel<LNBWD>", 1, lambda: capi.gemm_tn(capi.EPI_LNBWD, x768, w_fc1t, o192, X1=x192, X2=y192, ln=ln_bwd), M * (HP + 3 * C) * E),
        ("gemm_tn<192,STORE> d_ao", "gemm_tn_kernel<STORE>", 1, lambda: capi.gemm_tn(capi.EPI_STORE, x192, w_proj, o192), M * 2 * C * E),
        ("gemm_wgrad<192> proj", "gemm_wgrad_kernel", 1, lambda: capi.gemm_wgrad(x192, y192, wg_ws, splits(M, C), wg_out), M * 2 * C * E),
        ("win_attn_ws8_bwd", "win_attn_ws8_bwd_kernel", 1, lambda: capi.win_attn_bwd(geom, heads, x576, table, y192, o576, dtab), M * (2 * QW + C) * E),
        ("gemm_tn<192,LNBWD> qkv dgrad", "gemm_tn_kernel<LNBWD>", 1, lambda: capi.gemm_tn(capi.EPI_LNBWD, x576, w_qkvt, o192, X1=x192, X2=y192, ln=ln_bwd), M * (QW + 3 * C) * E),
        ("gemm_wgrad<192> qkv", "gemm_wgrad_kernel", 1, lambda: capi.gemm_wgrad(x576, x192, wg_ws, splits(M, QW), wg_out), M * (QW + C) * E),
    ]
    rows = []
    reps = 5
    for name, fn_name, per_block, call, nbytes in cases:
        for _ in range(2):
            call()
        # every case streams >= 200 MB of operands/results per launch (> the 126 MB L2), so back-to-back launches
        # cannot be served from cache; 5 launches between one event pair amortise the launch gap
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rows.append({"kernel": name, "function": fn_name, "launches_per_step": 36 * per_block, "ms_per_launch": ms,
                     "algorithmic_bytes": nbytes, "achieved": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / hbm})
    total_ms = sum(r["ms_per_launch"] * r["launches_per_step"] for r in rows)
    top = max(rows, key=lambda r: r["ms_per_launch"] * r["launches_per_step"])
    # measured DRAM traffic of that launch shape (ncu --set full, one capture per round under profiles/), keyed by the same
    # case name, so traffic and algorithmic bytes always refer to the SAME launch shape
    traffic, tsrc = None, None
    if batch == 16:
        for fn in ("r02_traffic.json", "r01_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", fn)) as f:
                    tj = json.load(f)
                t = tj.get(top["kernel"]) or tj.get(top["function"])
                if t and abs(t.get("algorithmic_bytes", top["algorithmic_bytes"]) - top["algorithmic_bytes"]) <= 0.02 * top["algorithmic_bytes"]:
                    traffic, tsrc = t["traffic"], f"profiles/{fn}: {t.get('capture', '')}"
                    break
            except Exception:
                continue
    roof = {"kernel": top["kernel"], "function": top["function"], "bound": "hbm", "achieved": top["achieved"], "peak": hbm,
            "unit": "GB/s", "frac": top["frac"], "traffic": traffic, "traffic_source": tsrc, "peak_kind": kind,
            "ms_per_launch": top["ms_per_launch"], "algorithmic_bytes": top["algorithmic_bytes"],
            "launches_per_step": top["launches_per_step"],
            "share_of_block_kernels": top["ms_per_launch"] * top["launches_per_step"] / total_ms,
            "note": "dominant per-block kernel launch shape by launches x duration, timed alone at the step's shape (operands "
                    "larger than L2); achieved = algorithmic bytes of that shape / mean launch duration; traffic = ncu dram "
                    "read+write bytes of the same shape"}
    return roof, rows


def run_gan(args):
    """BASELINE configs[3]: SwinIR generator (libsrk mirror) + UNetDiscriminatorSN + RaGAN / perceptual losses, in the
    arrangement of train_swin.py:147-259 — DDP(G, find_unused_parameters=True), DDP(D), fp16 autocast + GradScaler,
    requires_grad toggling, micro-batch 2 x 4 accumulation steps per optimizer step, EMA — eager, like the script.
    One bench "step" = one optimizer step = `accum` micro-steps; value = patches/s over all ranks."""
    import torch.distributed as dist
    from superresolution_def_b200 import _capi as capi
    from superresolution_def_b200.architecture_swin import SwinIR
    from superresolution_def_b200.gan import UNetDiscriminatorSN, GanTrainer
    from superresolution_def_b200.synth import synthetic_pairs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for the product arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    mb, accum = args.batch or 2, 4
    torch.manual_seed(0)
    net_g = SwinIR(**MODEL_KW).to(dev)
    net_d = UNetDiscriminatorSN(num_in_ch=1, num_feat=64).to(dev)
    DDP = torch.nn.parallel.DistributedDataParallel
    net_g = DDP(net_g, device_ids=[local], output_device=local, find_unused_parameters=True)
    net_d = DDP(net_d, device_ids=[local], output_device=local, find_unused_parameters=False)
    net_g.train(); net_d.train()
    tr = GanTrainer(net_g, net_d, accum=accum)
    lr_h, hr_h = synthetic_pairs(4, seed=77 + rank)
    host = [(lr_h[i:i + mb].contiguous().pin_memory(), hr_h[i:i + mb].contiguous().pin_memory()) for i in (0, 2)]
    devs = [(a.to(dev), b.to(dev)) for a, b in host]

    def opt_step(e2e, k):
        out = None
        for m in range(accum):
            if e2e:
                a, b = host[(k + m) & 1]
                lg, ld = tr.micro_step(a.to(dev, non_blocking=True), b.to(dev, non_blocking=True))
                out = (lg.item(), ld.item())     # the script reads both losses back every micro-step (train_swin.py:257-258)
            else:
                a, b = devs[(k + m) & 1]
                out = tr.micro_step(a, b)
        return out

    def timed(n, e2e):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for k in range(n):
            out = opt_step(e2e, k)
        e1.record(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out

    timed(args.warmup, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = capi.launch_count()
    ms, out = timed(args.steps, False)
    launches = capi.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    timed(1, True)
    ms_e2e, out_e = timed(args.steps, True)
    per_step = world * mb * accum
    line = {"metric": "SwinIR + UNetDiscriminatorSN GAN train patches/s (128x128 -> 512x512)", "value": per_step * args.steps / (ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (generator and discriminator kernels) / fp16 autocast (VGG loss network)",
            "data": "synthetic",
            "config": {"workload": f"train_swin.py micro-step semantics: D step + G step, RaGAN + L1 + VGG-perceptual (seeded VGG-19), DDP over NCCL, "
                                   f"micro-batch {mb} x {accum} accumulation per optimizer step (BASELINE configs[3])",
                       "global_batch": per_step, "parallelism": f"dp{world}", "launch": "eager (as the script)",
                       "l2": "two input batches rotated; activations of every forward exceed L2",
                       "discriminator": "libsrk (disc_engine: 4x4 stride-2 convolutions / transposed convolutions as tcgen05 GEMMs over patch matrices); generator = libsrk; VGG-19 loss network = stock ATen/cuDNN"},
            "e2e": {"value": per_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": accum * mb * (128 * 128 + 512 * 512) * 4, "d2h_bytes_per_step": accum * 8},
            "gpu_launches": launches, "clocks": clocks, "loss_g": float(out_e[0]), "loss_d": float(out_e[1]),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(0)


def _dbg(msg):
    if os.environ.get("SRK_BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')} +{time.perf_counter() - _T0:.1f}s] {msg}", file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def run_ours(args):
    import torch.distributed as dist
    from superresolution_def_b200 import _capi as capi
    from superresolution_def_b200.architecture_swin import SwinIR
    from superresolution_def_b200.dp import BucketedGradReducer, swinir_grad_groups
    from superresolution_def_b200.synth import synthetic_pairs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for the product arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    _dbg("process group ready")
    hybrid = args.workload == "hybrid"
    hat = args.workload in ("hat", "hybrid")   # not the default bench line: no CPU leg, HAT metric name
    B = args.batch or (8 if hat else 16)
    torch.manual_seed(0)
    if hybrid:
        from superresolution_def_b200.hybridmodels_hat import HybridHATRealESRGAN
        net = HybridHATRealESRGAN(**HYBRID_KW).to(dev)
    elif hat:
        from superresolution_def_b200.hat_arch import HAT
        net = HAT(**HAT_KW).to(dev)
    else:
        net = SwinIR(**MODEL_KW).to(dev)
    net.train()
    gflop = HYBRID_GFLOP_PER_PATCH_TRAIN if hybrid else HAT_GFLOP_PER_PATCH_TRAIN if hat else GFLOP_PER_PATCH_TRAIN
    if hybrid:
        wl_name = ("HybridHATRealESRGAN x4 (HAT C=90 window 8 x2 + 12 RRDB nf 48 + nearest x2; train_hat.py:132-136) training "
                   f"step (fwd + L1 + bwd + AdamW), bf16, batch {B}/GPU, 128^2->512^2")
    elif hat:
        wl_name = ("HAT x4 (window 16, OCAB, CAB, stochastic depth 0.1) training step (fwd + L1 + bwd + AdamW), bf16, "
                   f"batch {B}/GPU, 128^2->512^2 (BASELINE configs[2])")
    else:
        wl_name = f"SwinIR x4 training step (fwd + L1 + bwd + AdamW), bf16, batch {B}/GPU, 128^2->512^2 (BASELINE configs[1])"
    if world > 1:
        for p in net.parameters():
            dist.broadcast(p.data, 0)
    # N > 1: flat fp32 gradient buckets + NCCL all-reduce(AVG), launched bucket by bucket from post-accumulate hooks so that
    # the exchange of layer group k overlaps the backward of group k-1.  --dp-mode overlap (default): the hooks run inside
    # the stream capture too, so the per-bucket ncclAllReduce nodes sit on a forked branch of the ONE step graph next to
    # the backward kernels (NCCL is capture-safe).  --dp-mode serial: graph(fwd+bwd) -> all-reduces -> graph(AdamW), the
    # round-1 arrangement, kept for comparison.  N == 1: no exchange.
    dp_overlap = args.dp_mode == "overlap" or not args.graph
    reducer = BucketedGradReducer(swinir_grad_groups(net), world, overlap=dp_overlap) if world > 1 else None
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, betas=(0.9, 0.99), fused=True, capturable=args.graph)
    nsets = 4
    lr_h, hr_h = synthetic_pairs(min(B, 4), seed=1234 + rank)
    reps = (B + lr_h.shape[0] - 1) // lr_h.shape[0]
    host = []
    gen = torch.Generator().manual_seed(rank)
    for s in range(nsets):  # distinct batches: shuffled / flipped copies of the generated fields
        perm = torch.randperm(lr_h.shape[0] * reps, generator=gen)[:B]
        l = lr_h.repeat(reps, 1, 1, 1)[perm]
        h = hr_h.repeat(reps, 1, 1, 1)[perm]
        if s & 1:
            l, h = l.flip(-1), h.flip(-1)
        host.append((l.contiguous().pin_memory(), h.contiguous().pin_memory()))
    dev_sets = [(l.to(dev), h.to(dev)) for l, h in host]

    def fwd_bwd(lr, hr):
        if reducer is not None:
            reducer.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        sr = net(lr)
        loss = torch.nn.functional.l1_loss(sr.float(), hr)
        loss.backward()
        return loss

    def step(lr, hr):
        loss = fwd_bwd(lr, hr)
        if reducer is not None:
            reducer.finish()
        opt.step()
        return loss

    _dbg("model + data ready")
    if args.graph:
        from superresolution_def_b200.graphs import GraphedStep
        static_lr, static_hr = dev_sets[0][0].clone(), dev_sets[0][1].clone()
        l_before = capi.launch_count()
        graphed = None
        if reducer is None or dp_overlap:
            try:
                graphed = GraphedStep(step, (static_lr, static_hr), warmup=2)
                launches_per_step = (capi.launch_count() - l_before) // 3  # 2 warm-up runs + 1 capture run
                step = graphed  # replaying the graph re-issues exactly the captured launches
            except Exception as e:  # noqa: BLE001
                if reducer is None:
                    raise
                print(f"[bench rank {rank}] capturing the overlapped exchange failed ({e!r}); falling back to --dp-mode serial",
                      file=sys.stderr, flush=True)
                dp_overlap, reducer.overlap = False, False
                reducer.handles.clear()
                l_before = capi.launch_count()
        if graphed is None:
            g_fb = GraphedStep(fwd_bwd, (static_lr, static_hr), warmup=2)
            launches_per_step = (capi.launch_count() - l_before) // 3
            g_opt = GraphedStep(lambda: opt.step(), (), warmup=2)

            def step(lr, hr):  # graph(fwd+bwd) -> NCCL all-reduce of the buckets -> graph(AdamW)
                loss = g_fb(lr, hr)
                reducer.reduce_all()
                g_opt()
                return loss

    def timed(nsteps, e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(nsteps):
            if e2e:
                l, h = host[i % nsets]  # pinned host memory -> device inside the timed region
                if args.graph:
                    loss = step(l, h)
                else:
                    loss = step(l.to(dev, non_blocking=True), h.to(dev, non_blocking=True))
                last = loss.item()  # D2H read of the step's result
            else:
                l, h = dev_sets[i % nsets]
                last = step(l, h)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), (last if e2e else last.item())

    _dbg("graph captured" if args.graph else "eager mode")
    timed(args.warmup, False)
    _dbg("warm-up done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = capi.launch_count()
    ms, loss_v = timed(args.steps, False)
    launches = (capi.launch_count() - l0) if not args.graph else launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    _dbg("timed region done")
    timed(1, True)
    ms_e2e, _ = timed(args.steps, True)
    _dbg("e2e region done")
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30

    value = world * B * args.steps / (ms * 1e-3)
    e2e_v = world * B * args.steps / (ms_e2e * 1e-3)
    peaks = load_peaks()
    line = {"metric": METRIC.replace("SwinIR", "HybridHAT" if hybrid else "HAT") if hat else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl_name, "global_batch": world * B, "parallelism": f"dp{world}",
                       "l2": "activations saved per step (~60 GB) exceed the 126 MB L2; 4 input batches rotated",
                       "grad_allreduce_mb": reducer.nbytes / 2 ** 20 if reducer is not None else 0,
                       "launch": "whole step replayed as one CUDA graph" if args.graph else "eager"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": B * (128 * 128 + 512 * 512) * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "loss": loss_v, "peak_mem_gb": mem_gb,
            "step_tflops": value / world * gflop / 1e3,
            "step_frac_of_bf16_sustained": value / world * gflop / 1e3 / peaks[2]}
    if reducer is not None:
        # what the exchange costs: (a) all buckets reduced back to back on an otherwise idle GPU, (b) the step with the
        # exchange minus the same step without it (eager steps, CUDA events, max over ranks) = the part backward does not hide
        def ev_ms(fn, n):
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record(); torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / n], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        ev_ms(reducer.reduce_all, 2)
        allreduce_ms = ev_ms(reducer.reduce_all, 5)
        line["comm"] = {"mode": ("one graph, per-bucket all-reduce overlapped with backward" if dp_overlap and args.graph
                                 else "graph(fwd+bwd) -> all-reduce -> graph(AdamW)" if args.graph else "eager, overlapped hooks"),
                        "buckets": len(reducer.buckets), "allreduce_ms_alone": allreduce_ms,
                        "bus_gbs_alone": reducer.nbytes * 2 * (world - 1) / world / (allreduce_ms * 1e-3) / 1e9}
        if args.single_gpu_ms:
            line["comm"]["exposed_comm_ms"] = ms / args.steps - args.single_gpu_ms
            line["comm"]["exposed_note"] = "this run's ms_per_step minus the N=1 ms_per_step passed by --single-gpu-ms"
    if rank == 0:
        cs = ClockSampler(local)
        cs.start()
        line["roofline"], line["roofline_kernels"] = roofline_probe(B, peaks)
        rc = cs.stop()
        # the probe runs right after the timed steps, i.e. on a GPU that sits at its power cap: compute- / shared-memory-bound
        # kernels (the attention cores) scale with this clock, HBM-bound ones do not
        line["roofline"]["clocks_during_probe"] = {"sm_mhz": rc.get("sm_mhz"), "sm_max_mhz": rc.get("sm_max_mhz"), "reasons": rc.get("reasons")}
        # BASELINE.json's metric also names the attention tensor-core utilisation.  Live figure: the MMA work the attention
        # kernels issue (padded head_dim 32, per window 64x64 logits) / CUDA-event time / measured dense bf16 peak; the ncu
        # tensor-pipe counters of this round's captures are quoted next to it.
        att = {}
        for r in line["roofline_kernels"]:
            if r["function"].startswith("win_attn"):
                nmm = 2 if "fwd" in r["function"] else 5   # QK^T, PV | + dP, dV, dK, dQ (S recomputed)
                fl = nmm * 2.0 * 64 * 32 * (B * 16384) * 6
                att[r["function"]] = {"mma_tflops": fl / (r["ms_per_launch"] * 1e-3) / 1e12,
                                      "frac_of_bf16_burst_peak": fl / (r["ms_per_launch"] * 1e-3) / 1e12 / peaks[1],
                                      "hbm_frac": r["frac"]}
        for fn in ("r02_traffic.json", "r01_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", fn)) as f:
                    tp = json.load(f)["_attention_tensor_pipe"]
                att["ncu_tensor_pipe_active"] = {k: v for k, v in tp.items() if not k.startswith("_")}
                att["ncu_source"] = f"profiles/{fn} (ncu --set full captures of this tree's kernels)"
                break
            except Exception:
                continue
        line["attn_tensor_pipe_util"] = att
        if world == 1 and not hat and not args.no_gpu_baseline:
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_baseline(B)
            if "value" in line["gpu_eager_baseline"]:
                line["gpu_eager_baseline"]["ours_over_eager"] = value / line["gpu_eager_baseline"]["value"] * \
                    (1.0 if line["gpu_eager_baseline"]["batch"] == B else 1.0)
        if world == 1 and not hat and not args.no_sub:
            # the other two generators of the path, a few steps each (separate processes: their own graphs and memory)
            line["sub_records"] = {}
            for wl in ("hat", "hybrid", "gan"):
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", wl, "--steps", "3" if wl == "gan" else "5", "--warmup", "3",
                                        "--no-cpu-baseline", "--no-sub", "--no-gpu-baseline"], capture_output=True, text=True,
                                       timeout=420, cwd=ROOT)
                    js = [l for l in r.stdout.splitlines() if l.startswith("{")]
                    if r.returncode == 0 and js:
                        d = json.loads(js[-1])
                        line["sub_records"][wl] = {k: d[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "config",
                                                                      "gpu_launches", "step_tflops", "step_frac_of_bf16_sustained",
                                                                      "peak_mem_gb", "clocks", "steps", "warmup") if k in d}
                    else:
                        line["sub_records"][wl] = {"unavailable": (r.stderr or r.stdout)[-300:]}
                except Exception as e:  # noqa: BLE001
                    line["sub_records"][wl] = {"unavailable": repr(e)[:300]}
        if world == 1 and not args.no_cpu_baseline and not hat:
            v, spp, threads = cpu_reference_arm(2, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "same_config": False,
                                    "sample": "2 fp32 training steps of batch 1 after 1 warm-up (oracle/swinir_oracle.py); "
                                              "the GPU arm's batch is 16: same model and patch shape, smaller batch"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # The step graph holds captured NCCL kernels; tearing the communicator down underneath it made
        # destroy_process_group() block until the NCCL watchdog fired (observed on 2 x B200).  Everything is measured and
        # printed at this point: synchronise, flush and leave without running the communicator's destructor.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="patches per GPU per step (default: 16 SwinIR = BASELINE configs[1], 8 HAT)")
    ap.add_argument("--workload", default="swinir", choices=["swinir", "hat", "hybrid", "gan"],
                    help="swinir = the bench line (configs[1]); hat = configs[2]; hybrid = train_hat.py's generator; "
                         "gan = configs[3] (SwinIR + UNetDiscriminatorSN, train_swin.py step semantics)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch-eager leg (gpu_eager_baseline)")
    ap.add_argument("--no-sub", action="store_true", help="skip the HAT / hybrid sub-records")
    ap.add_argument("--dp-mode", default="overlap", choices=["overlap", "serial"],
                    help="N>1: per-bucket all-reduce captured inside the step graph next to backward, or serial between graphs")
    ap.add_argument("--single-gpu-ms", type=float, default=0.0, help="N>1: ms_per_step of the N=1 run, to report exposed_comm_ms")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="issue the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "gan":
        run_gan(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
