"""Timing probe (not a test; `python -m tests.probe_gan_phases` on the GPU box): where one train_swin.py micro-step
(BASELINE configs[3], micro-batch 2, 128^2 -> 512^2) spends its time.  Every phase is bracketed by a device synchronise, so a
phase costs max(host launch time, device time); `busy_ms` is the device-side kernel time of the same phase from CUDA events
recorded around it without host syncs in a second pass (the difference is host / launch overhead)."""
import json
import time

import torch


def main():
    from superresolution_def_b200.architecture_swin import SwinIR
    from superresolution_def_b200.gan import UNetDiscriminatorSN, CombinedGANLoss, DiscriminatorLoss
    torch.manual_seed(0)
    kw = dict(upscale=4, in_chans=1, img_size=128, window_size=8, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6,
              mlp_ratio=2)   # train_swin.py:147-149 (= bench.MODEL_KW)
    g = SwinIR(**kw).cuda().train()
    d = UNetDiscriminatorSN(1, 64).cuda().train()
    crit_g, crit_d = CombinedGANLoss().cuda(), DiscriminatorLoss().cuda()
    opt_g = torch.optim.AdamW(g.parameters(), lr=1e-4, betas=(0.9, 0.99), weight_decay=0)
    opt_d = torch.optim.AdamW(d.parameters(), lr=1e-4, betas=(0.9, 0.99), weight_decay=0)
    scaler = torch.amp.GradScaler("cuda")
    lr, hr = torch.rand(2, 1, 128, 128, device="cuda"), torch.rand(2, 1, 512, 512, device="cuda")
    phases = {}

    def run(sync):
        st = {}

        def ph(name, fn):
            if sync:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = fn()
                torch.cuda.synchronize()
                st[name] = (time.perf_counter() - t0) * 1e3
            else:
                t0 = time.perf_counter()
                out = fn()
                st[name] = (time.perf_counter() - t0) * 1e3
            return out
        for p in d.parameters():
            p.requires_grad = True
        for p in g.parameters():
            p.requires_grad = False
        with torch.autocast("cuda"):
            def f():
                with torch.no_grad():
                    return g(lr)
            sr = ph("G forward (no_grad)", f)
            d_real = ph("D(hr) forward", lambda: d(hr))
            d_fake = ph("D(sr.detach()) forward", lambda: d(sr.detach()))
            loss_d = ph("D loss", lambda: crit_d(d_real, d_fake)[0])
        ph("D backward (2 passes)", lambda: scaler.scale(loss_d).backward())
        ph("opt_d step", lambda: (scaler.step(opt_d), opt_d.zero_grad()))
        for p in d.parameters():
            p.requires_grad = False
        for p in g.parameters():
            p.requires_grad = True
        with torch.autocast("cuda"):
            sr_g = ph("G forward", lambda: g(lr))
            d_fake_g = ph("D(sr) forward (frozen)", lambda: d(sr_g))
            d_real_g = ph("D(hr) forward (frozen, detached)", lambda: d(hr).detach())
            loss_g = ph("G loss (L1 + VGG-19 + RaGAN)", lambda: crit_g(sr_g, hr, d_real_g, d_fake_g)[0])
        ph("G backward (VGG + D image gradient + G)", lambda: scaler.scale(loss_g).backward())
        ph("opt_g step + scaler.update", lambda: (scaler.step(opt_g), scaler.update(), opt_g.zero_grad()))
        return st

    for _ in range(3):
        run(True)
    acc = {}
    n = 5
    for _ in range(n):
        for k, v in run(True).items():
            acc[k] = acc.get(k, 0.0) + v / n
    host = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        for k, v in run(False).items():
            host[k] = host.get(k, 0.0) + v / n
    t_host = (time.perf_counter() - t0) * 1e3 / n
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) * 1e3 / n
    print(json.dumps({"phase_ms_with_sync": {k: round(v, 2) for k, v in acc.items()}, "sum_with_sync_ms": round(sum(acc.values()), 2),
                      "host_ms_no_sync": {k: round(v, 2) for k, v in host.items()}, "host_total_ms": round(t_host, 2),
                      "micro_step_ms_async": round(t_all, 2)}))


if __name__ == "__main__":
    main()
