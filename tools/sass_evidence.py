"""Static SASS evidence for the library as built: per kernel, counts of the mnemonics that prove the Blackwell path
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA) next to the legacy tensor path (HMMA = mma.sync)
and cp.async (LDGSTS).  Usage: python tools/sass_evidence.py [libsrk.so] > profiles/r02_sass_mix.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent.parent / "superresolution_def_b200" / "_lib" / "libsrk.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
keys = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "LDGSTS", "LDSM", "STSM", "MUFU"]
res, cur, fn_i = [], None, -1
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn_i += 1
        cur = collections.Counter()
        res.append((names[fn_i] if fn_i < len(names) else m.group(1), cur))
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for k in keys:
            if op.startswith(k):
                cur[k] += 1
print(f"# cuobjdump -sass {Path(lib).name}: static instruction counts per kernel (sm_100a)")
print(f"# {'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in keys))
for name, c in sorted(res, key=lambda t: t[0]):
    short = re.sub(r"\(.*", "", name).replace("srk::", "")
    if c["_total"] < 64:
        continue
    print(f"{short[:72]:72s} {c['_total']:7d} " + " ".join(f"{c[k]:7d}" for k in keys))
tc = [n for n, c in res if c["UTCHMMA"]]
print(f"# kernels with tcgen05.mma (UTCHMMA): {len(tc)}; with TMA loads (UTMALDG): {len([1 for n, c in res if c['UTMALDG']])}; "
      f"with mma.sync (HMMA): {len([1 for n, c in res if c['HMMA']])}")
