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
    """Times the dominant kernel alone (CUDA events, current stream) at the step's shapes: the fc1 GEMM with the
    fused GELU epilogue, tcgen05, M = batch*16384 tokens, K = 192, N = 768 (two bf16 outputs).  It is HBM-bound:
    algorithmic bytes per launch = read xn2 (M*192*2) + write act and dact (2*M*768*2) + weights."""
    from superresolution_def_b200 import _capi as capi
    hbm, _, _, kind = peaks
    M, K, N = batch * 16384, 192, 768
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Bw = (torch.randn(N, K, device="cuda") / 14).to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    C2 = torch.empty_like(C)
    ln = capi.make_ln_args(N, 720, None)
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    for _ in range(3):
        capi.gemm_tn(capi.EPI_GELU2, A, Bw, C, C2=C2, ln=ln)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.gemm_tn(capi.EPI_GELU2, A, Bw, C, C2=C2, ln=ln)
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    bytes_alg = M * K * 2 + 2 * M * N * 2 + N * K * 2
    achieved = bytes_alg / (ms * 1e-3) / 1e9
    return {"kernel": "gemm_tn_kernel<256,EPI_GELU2> (fc1+GELU, tcgen05)", "bound": "hbm", "achieved": achieved,
            "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_kind": kind,
            "ms_per_launch": ms, "algorithmic_bytes": bytes_alg,
            "tflops_of_kernel": 2.0 * M * N * K / (ms * 1e-3) / 1e12}


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
    hat = args.workload == "hat"
    B = args.batch or (8 if hat else 16)
    torch.manual_seed(0)
    if hat:
        from superresolution_def_b200.hat_arch import HAT
        net = HAT(**HAT_KW).to(dev)
    else:
        net = SwinIR(**MODEL_KW).to(dev)
    net.train()
    gflop = HAT_GFLOP_PER_PATCH_TRAIN if hat else GFLOP_PER_PATCH_TRAIN
    wl_name = ("HAT x4 (window 16, OCAB, CAB, stochastic depth 0.1) training step (fwd + L1 + bwd + AdamW), bf16, "
               f"batch {B}/GPU, 128^2->512^2 (BASELINE configs[2])") if hat else \
              (f"SwinIR x4 training step (fwd + L1 + bwd + AdamW), bf16, batch {B}/GPU, 128^2->512^2 (BASELINE configs[1])")
    if world > 1:
        for p in net.parameters():
            dist.broadcast(p.data, 0)
    # N > 1: flat fp32 gradient buckets + NCCL all-reduce (overlapped with backward from hooks when run eagerly; issued
    # between the two graph replays otherwise, so no NCCL call sits inside a stream capture).  N == 1: no exchange,
    # gradients are handed to the optimizer as produced.
    reducer = BucketedGradReducer(swinir_grad_groups(net), world, overlap=not args.graph) if world > 1 else None
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
        if reducer is None:
            graphed = GraphedStep(step, (static_lr, static_hr), warmup=2)
            launches_per_step = (capi.launch_count() - l_before) // 3  # 2 warm-up runs + 1 capture run
            step = graphed  # replaying the graph re-issues exactly the captured launches
        else:
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
    line = {"metric": METRIC.replace("SwinIR", "HAT") if hat else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
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
    if rank == 0:
        line["roofline"] = roofline_probe(B, peaks)
        if world == 1 and not args.no_cpu_baseline and not hat:
            v, spp, threads = cpu_reference_arm(2, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "2 fp32 training steps of batch 1 after 1 warm-up (oracle/swinir_oracle.py)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="patches per GPU per step (default: 16 SwinIR = BASELINE configs[1], 8 HAT)")
    ap.add_argument("--workload", default="swinir", choices=["swinir", "hat"], help="swinir = the bench line (configs[1]); hat = configs[2]")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="issue the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
