"""Per-block kernel table of the SwinIR step (the `roofline_kernels` leg of bench.py alone): python tools/gpu_probe_roofline.py [batch] [filter]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
flt = sys.argv[2] if len(sys.argv) > 2 else ""
roof, rows = bench.roofline_probe(batch, bench.load_peaks())
tot = 0.0
for r in rows:
    if flt and flt not in r["kernel"]:
        continue
    tot += r["ms_per_launch"] * r["launches_per_step"] / 36
    print(f"{r['kernel']:34s} {r['ms_per_launch'] * 1e3:8.1f} us  {r['achieved']:7.0f} GB/s  frac {r['frac']:.3f}  x{r['launches_per_step'] // 36}")
print(f"per-block sum {tot * 1e3:.1f} us")
