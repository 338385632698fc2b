"""Opcode mix + stall reasons from `ncu -i X.ncu-rep --page source --csv --print-source sass`.  Usage: ncu_sass_mix.py X.ncu-rep [kernel-index]"""
import csv, collections, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernels may follow each other: take the first block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr = b["hdr"]; ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter(); n = 0
for r in b["rows"]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip().split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    e = int(r[ix["Instructions Executed"]] or 0)
    tot[op] += e; n += e
    samp[op] += int(r[ix["# Samples"]] or 0)
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h] += int(r[ix[h]] or 0)
print(b["name"][:90], "total warp-instructions", n)
for k, v in tot.most_common(28):
    print(f"  {k:10s} {v:12d} {100*v/n:5.1f}%  samples {samp[k]}")
s = sum(stalls.values())
print({k: round(100 * v / s, 1) for k, v in stalls.most_common(10)})
