"""Timing probe (not a test; run on the GPU box: `python -m tests.probe_disc_timing`): UNetDiscriminatorSN forward + backward
at train_swin.py's shape (micro-batch 2, 512^2 -> 256^2 logits) — the libsrk path (gan.UNetDiscriminatorSN) against the ATen
restatement in oracle/discriminator_oracle.py under the script's fp16 autocast (cuDNN, NCHW and channels_last).  CUDA events,
3 warm-ups, 10 iterations; D step = forward + weight gradients, G step = forward + image gradient with frozen parameters."""
import json
import sys

import torch


def _time(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    from oracle.discriminator_oracle import UNetDiscriminatorSN as OraD
    from superresolution_def_b200.gan import UNetDiscriminatorSN
    from superresolution_def_b200 import _capi as capi
    if len(sys.argv) > 1 and sys.argv[1] == "hat":
        return main_hat()
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    torch.manual_seed(0)
    mine = UNetDiscriminatorSN(1, 64).cuda().train()
    x = torch.rand(B, 1, 512, 512, device="cuda")
    res = {"shape": [B, 1, 512, 512]}

    def steps(net, tag, autocast):
        def d_step():
            for p in net.parameters():
                p.requires_grad = True
            with torch.autocast("cuda", enabled=autocast):
                out = net(x)
            out.float().mean().backward()
            net.zero_grad(set_to_none=True)

        xs = x.clone().requires_grad_(True)

        def g_step():
            for p in net.parameters():
                p.requires_grad = False
            with torch.autocast("cuda", enabled=autocast):
                out = net(xs)
            out.float().mean().backward()
            xs.grad = None

        def fwd():
            with torch.no_grad(), torch.autocast("cuda", enabled=autocast):
                net(x)

        res[tag] = {"d_step_ms": _time(d_step), "g_step_ms": _time(g_step), "fwd_ms": _time(fwd)}

    def gpu_busy(net, tag):
        """device-side time of one D step (forward + all gradients): sum of kernel durations from the torch profiler —
        the eager step is launch-bound on both sides, so the wall/event time above says little about the kernels"""
        from torch.profiler import profile, ProfilerActivity
        for p in net.parameters():
            p.requires_grad = True
        xs = x.clone().requires_grad_(True)

        def one():
            with torch.autocast("cuda"):
                out = net(xs)
            out.float().mean().backward()
            net.zero_grad(set_to_none=True)
            xs.grad = None
        for _ in range(3):
            one()
        torch.cuda.synchronize()
        l0 = capi.launch_count()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            one()
            torch.cuda.synchronize()
        rows = sorted(((e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0),
                      key=lambda r: -r[1])
        res[tag + "_gpu_busy"] = {"total_us": sum(r[1] for r in rows), "kernels": sum(r[2] for r in rows),
                                  "libsrk_launches": capi.launch_count() - l0,
                                  "top": [[k[:90], round(t, 1), c] for k, t, c in rows[:14]]}

    steps(mine, "libsrk", True)
    gpu_busy(mine, "libsrk")
    ora = OraD(1, 64).cuda().train()
    ora.load_state_dict(mine.state_dict())
    steps(ora, "aten_fp16_autocast_nchw", True)
    ora_cl = OraD(1, 64).cuda().train().to(memory_format=torch.channels_last)
    ora_cl.load_state_dict(mine.state_dict())
    steps(ora_cl, "aten_fp16_autocast_channels_last", True)
    gpu_busy(ora_cl, "aten_fp16_autocast_channels_last")
    print(json.dumps(res))


def main_hat():
    """models/discriminator_hat.py at train_hat.py's shape (BATCH_SIZE = 1, 512^2, fp32 training script: no autocast)."""
    from oracle.discriminator_oracle import UNetDiscriminatorSNHat as OraD
    from superresolution_def_b200.discriminator_hat import UNetDiscriminatorSN
    torch.manual_seed(0)
    mine = UNetDiscriminatorSN(1, 64).cuda().train()
    x = torch.rand(1, 1, 512, 512, device="cuda")
    res = {"shape": [1, 1, 512, 512], "variant": "models/discriminator_hat.py"}

    def steps(net, tag, autocast):
        def d_step():
            for p in net.parameters():
                p.requires_grad = True
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                out = net(x)
            out.float().mean().backward()
            net.zero_grad(set_to_none=True)

        xs = x.clone().requires_grad_(True)

        def g_step():
            for p in net.parameters():
                p.requires_grad = False
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                out = net(xs)
            out.float().mean().backward()
            xs.grad = None

        res[tag] = {"d_step_ms": _time(d_step), "g_step_ms": _time(g_step)}

    steps(mine, "libsrk", False)
    ora = OraD(1, 64).cuda().train()
    ora.load_state_dict(mine.state_dict())
    steps(ora, "aten_fp32_as_the_script", False)
    steps(ora, "aten_bf16_autocast", True)
    ora_cl = OraD(1, 64).cuda().train().to(memory_format=torch.channels_last)
    ora_cl.load_state_dict(mine.state_dict())
    steps(ora_cl, "aten_bf16_autocast_channels_last", True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
