"""Timeline of CTA 0 of the attention kernel (globaltimer stamps): where do the microseconds of one unit go?"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_pytorch_b200 import _lib  # noqa: E402

L = _lib.lib()
L.b200vit_debug_set_trace.argtypes = [ctypes.c_void_p]
B, N, H, dh = 512, 197, 12, 64
qkv = torch.randn(B * N, 3 * H * dh, device="cuda").bfloat16()
out = torch.zeros(B * N, H * dh, device="cuda", dtype=torch.bfloat16)
names = {0: "mma:before full wait", 1: "mma:full ok", 2: "mma:S0 issued", 3: "mma:p_ready0 seen", 4: "mma:PV0 issued",
         8: "wg0:before s_full wait", 9: "wg0:s_full seen", 10: "wg0:p_ready arrive", 11: "wg0:o_full seen",
         12: "wg0:stores issued"}
for mode in (0, 1):
    L.b200vit_debug_set(1, mode)
    for _ in range(3):
        _lib.attention(qkv, out, B, N, H, dh, 0.125)
    tr = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
    L.b200vit_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    _lib.attention(qkv, out, B, N, H, dh, 0.125)
    torch.cuda.synchronize()
    L.b200vit_debug_set_trace(None)
    t = tr.cpu()
    t0 = int(t[0, 0])
    print(f"== mode {mode} ({'two CTAs/SM, 1 WG each' if mode == 0 else 'one CTA/SM, 2 WGs, 2 stages'}) ==")
    for it in (2, 3, 4, 10, 20):
        row = t[it]
        base = int(row[0])
        print(f"it {it:2d} start +{(base - t0) / 1e3:8.2f} us: " + "  ".join(
            f"{names[k].split(':')[1]}={(int(row[k]) - base) / 1e3:6.2f}" for k in sorted(names) if int(row[k]) > 0))
    per = (int(t[40, 0]) - int(t[8, 0])) / 32 / 1e3
    print(f"   mean period per iteration (it 8..40): {per:.2f} us")
