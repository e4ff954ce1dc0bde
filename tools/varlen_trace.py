"""Timeline of CTA 0 of the pipelined varlen attention kernel (%globaltimer stamps) on the NaViT config-5 batch."""
import ctypes
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_pytorch_b200 import _lib  # noqa: E402

L = _lib.lib()
L.b200vit_debug_set_trace.argtypes = [ctypes.c_void_p]
H, dh = 16, 64
random.seed(0)
lengths = [random.randrange(4, 33) * random.randrange(4, 33) for _ in range(256)]
T = sum(lengths)
qkv = torch.randn(T, 3 * H * dh, device="cuda").bfloat16()
out = torch.zeros(T, H * dh, device="cuda", dtype=torch.bfloat16)
cu, tp, tiles = _lib.varlen_index(lengths, "cuda")
for _ in range(3):
    _lib.attention_varlen(qkv, out, cu, tp, tiles, H, dh, 1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
_lib.attention_varlen(qkv, out, cu, tp, tiles, H, dh, 1.0)
e1.record()
torch.cuda.synchronize()
print(f"T={T} tiles={tiles} units={tiles * H}  kernel {e0.elapsed_time(e1) * 1e3:.1f} us")
tr = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
L.b200vit_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
_lib.attention_varlen(qkv, out, cu, tp, tiles, H, dh, 1.0)
torch.cuda.synchronize()
L.b200vit_debug_set_trace(None)
t = tr.cpu()
t0 = int(t[0, 0])
names = {0: "sm:start", 1: "sm:located", 2: "sm:ph1 S0", 3: "sm:ph1 end", 4: "sm:ph2 S0", 5: "sm:ph2 end", 6: "sm:o_full",
         14: "sm:stored", 8: "mma:start", 9: "mma:q_full", 10: "mma:lookahead issued", 11: "mma:end", 12: "tma:start",
         13: "tma:q_empty"}
for it in range(2, 40):
    row = t[it]
    base = int(row[0])
    if base == 0:
        break
    print(f"unit {it:2d} nb={int(row[7]):2d} start +{(base - t0) / 1e3:8.2f} us | " + " ".join(
        f"{names[k].split(':')[1] if k < 8 or k == 14 else names[k]}={(int(row[k]) - base) / 1e3:6.2f}"
        for k in (1, 2, 3, 4, 5, 6, 14, 8, 9, 10, 11, 12, 13) if int(row[k]) > 0))
