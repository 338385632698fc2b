"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log of tools/gpu_step_eager.py: keeps the launches of the LAST
step (second half of the srk/optimizer launches after the model was built), writes index,kernel,duration_us and prints the
per-kernel table.  Usage: python tools/ncu_launch_summary.py ncu_log.csv out_step.csv > out_summary.txt"""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    if len(r) != len(hdr):
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    rows.append((r[ik], v))
# the two steps: find the first srk::conv_in1_fwd_kernel launches (one per forward)
starts = [i for i, (k, _) in enumerate(rows) if "conv_in1_fwd_kernel" in k]
assert len(starts) >= 2, f"expected two steps, found {len(starts)} forwards"
step = rows[starts[-1]:]
with open(sys.argv[2], "w") as f:
    f.write("index,kernel,duration_us\n")
    for i, (k, v) in enumerate(step):
        f.write(f'{i},"{k[:120]}",{v:.3f}\n')
tot = sum(v for _, v in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in step:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += 1
    agg[k][1] += v
print(f"one eager SwinIR training step (B=16), ncu gpu__time_duration.sum, {len(step)} launches, {tot / 1e3:.2f} ms of kernel time")
for k, (n, v) in sorted(agg.items(), key=lambda t: -t[1][1])[:34]:
    print(f"{v / 1e3:9.3f} ms {100 * v / tot:6.1f}% n={n:4d} avg={v / n:8.1f}us  {k[:100]}")
