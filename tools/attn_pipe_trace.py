"""%globaltimer timeline of CTA 0 of attention_pipe_kernel (test hook b200vit_debug_set_trace): per pipeline tile the
MMA thread's and the softmax warps' time stamps, printed as microseconds relative to the first stamp."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_pytorch_b200 import _lib  # noqa: E402

L = _lib.lib()
L.b200vit_debug_set_trace.argtypes = [ctypes.c_void_p]
B, N, H = 512, 197, 12
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 0.5).bfloat16()
o = torch.zeros(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
L.b200vit_debug_set(1, 2)
for _ in range(2):
    _lib.attention(qkv, o, B, N, H, 64, 0.125)
tr = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
L.b200vit_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
_lib.attention(qkv, o, B, N, H, 64, 0.125)
torch.cuda.synchronize()
L.b200vit_debug_set_trace(None)
t = tr.cpu().view(64, 16)
t0 = int(t[t > 0].min())
names = {0: "S.ops", 1: "S.iss", 2: "PV.p", 3: "PV.o", 4: "PV.iss", 8: "w0.wait", 9: "w0.S", 5: "w0.ld", 10: "w0.bar",
         6: "w0.exp", 11: "w0.P", 14: "w13.ld", 15: "w13.bar", 7: "w13.P", 12: "ep.O", 13: "ep.done"}
order = [0, 1, 8, 9, 5, 10, 6, 11, 14, 15, 7, 2, 3, 4, 12, 13]
print("tile " + " ".join(f"{names[k]:>8}" for k in order))
for j in range(28, 40):
    print(f"{j:4d} " + " ".join(f"{(int(t[j, k]) - t0) / 1e3:8.2f}" if t[j, k] > 0 else "       -" for k in order))
