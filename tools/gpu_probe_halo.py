"""Times a dense-block style conv (cin -> 24, batch 8, 256^2) under the conv modes given by SRK_CONV_HALO, for an input
view that ends inside a 64-channel chunk (TMA zero-fills the rest) and one that fills its chunks exactly."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200 import _capi as capi, conv_engine as cv
Bh, Hh, Wh, gc = 8, 256, 256, 24
Th = Bh * Hh * Wh
bf = torch.bfloat16
V = capi.view
for pitch, cin, cout_p in ((144, 96, 64), (192, 128, 64), (192, 128, 128), (144, 64, 64)):
    cat = torch.randn(Th, pitch, device="cuda").to(bf)
    out = torch.empty(Th, 64, device="cuda", dtype=bf)
    w = torch.randn(gc, cin, 3, 3, device="cuda") / 30; b = torch.zeros(gc, device="cuda")
    cin_p = (cin + 63) // 64 * 64
    wf, wt, bp = cv.conv_weights(w, b, cout_p, cin_p)
    def run():
        capi.conv3x3_igemm_v(capi.CEPI_BIAS_LRELU, Bh, Hh, Wh, cin_p, cout_p, gc, V(cat, 0, cin), wf, bp, V(out, 0, gc), slope=0.2)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); e1.synchronize()
    print("SRK_CONV_HALO", os.environ.get("SRK_CONV_HALO"), f"pitch {pitch} cin {cin} Cout_p {cout_p}: us/launch", round(e0.elapsed_time(e1) * 100, 1))
