"""Top stall-sampled SASS instructions of an .ncu-rep (first kernel): python tools/ncu_top_sass.py file.ncu-rep [N]"""
import csv, subprocess, sys
f = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", f, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[isamp].isdigit()]
# (the report repeats the header block per kernel launch: only the first launch is kept)
seen = set(); first = []
for r in body:
    if r[ia] in seen: break
    seen.add(r[ia]); first.append(r)
body = first
tot = sum(int(r[isamp] or 0) for r in body)
print(f"{rows[0][1][:100]}  total samples {tot}")
for k, r in enumerate(body):
    r.append(k)
for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:n]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}%  #{r[-1]:5d} {r[isrc].strip()[:70]:70s} {st}")
