"""Times win_attn_ws8_bwd at the bench shapes (batch 16, 128^2, 6 heads, shift 4)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_def_b200 import _capi as capi
B, T = 16, 16 * 16384
bf = torch.bfloat16
qkv = torch.randn(T, 576, device="cuda").to(bf); tab = torch.randn(225, 6, device="cuda")
do = torch.randn(T, 192, device="cuda").to(bf); dq = torch.empty_like(qkv); dt = torch.empty_like(tab)
g = capi.SrkGeom(B, 128, 128, 8, 4)
for _ in range(3): capi.win_attn_bwd(g, 6, qkv, tab, do, dq, dt)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): capi.win_attn_bwd(g, 6, qkv, tab, do, dq, dt)
e1.record(); e1.synchronize()
print("SRK_ATTN_BWD_CTAS", os.environ.get("SRK_ATTN_BWD_CTAS"), "us/launch", round(e0.elapsed_time(e1) * 100, 1))
