"""Build superresolution_def_b200/_lib/libsrk_wd.so: the same sources with -DSRK_WATCHDOG=1 (every mbarrier wait traps after
2^26 polls instead of hanging the GPU).  Use it for the first run of a new kernel: SRK_LIB=<path> python tools/...  (tools only)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import superresolution_def_b200._build as b  # noqa: E402

d = b.LIBDIR / "obj_wd"
d.mkdir(parents=True, exist_ok=True)
objs, procs = [], []
for src in b.sources():
    o = d / (src.stem + ".o")
    objs.append(str(o))
    procs.append(subprocess.Popen(["nvcc", *b.NVCC_FLAGS, "-DSRK_WATCHDOG=1", "-c", str(src), "-o", str(o)]))
for p in procs:
    assert p.wait() == 0
subprocess.check_call(["nvcc", "-shared", "-o", str(b.LIBDIR / "libsrk_wd.so"), *objs])
print(b.LIBDIR / "libsrk_wd.so")
