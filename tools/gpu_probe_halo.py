"""Times the dense-block conv3 forward (cin 96 -> 24, batch 8, 256^2) under the conv modes given by SRK_CONV_HALO."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200 import _capi as capi, conv_engine as cv
Bh, Hh, Wh, nf, gc, k = 8, 256, 256, 48, 24, 2
Th = Bh * Hh * Wh
cin = nf + k * gc
bf = torch.bfloat16
cat = torch.randn(Th, nf + 4 * gc, device="cuda").to(bf)
w = torch.randn(gc, cin, 3, 3, device="cuda") / 30; b = torch.zeros(gc, device="cuda")
wf, wt, bp = cv.conv_weights(w, b, 64, 128)
V = capi.view
def run():
    capi.conv3x3_igemm_v(capi.CEPI_BIAS_LRELU, Bh, Hh, Wh, 128, 64, gc, V(cat, 0, cin), wf, bp, V(cat, cin, gc), slope=0.2)
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); e1.synchronize()
print("SRK_CONV_HALO", os.environ.get("SRK_CONV_HALO"), "us/launch", e0.elapsed_time(e1) * 100)
