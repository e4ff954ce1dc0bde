"""Bring-up probe of b200vit_patch_embed_tma: identity weights and unit statistics make the kernel print the A operand
exactly as the tensor core saw it; every output column is matched against the pixels of its patch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_pytorch_b200 import _lib  # noqa: E402

L = _lib.lib()
B, C, H, W = 2, 3, 32, 48
K = C * 256
torch.manual_seed(0)
# distinct small integers per pixel (exact in bf16 up to 256): value = position code
img = torch.arange(B * C * H * W, dtype=torch.float32).reshape(B, C, H, W) % 251
img = img.bfloat16().cuda()
w = torch.eye(K, dtype=torch.bfloat16, device="cuda")
bias = torch.zeros(K, device="cuda")
col_s = torch.zeros(K, device="cuda")
n = (H // 16) * (W // 16)
stats = torch.zeros(B * n, 2, device="cuda")
stats[:, 1] = float(K)      # mean 0, variance 1
y = torch.full((B * n, K), float("nan"), device="cuda")
rc = L.b200vit_patch_embed_tma(img.data_ptr(), w.data_ptr(), bias.data_ptr(), col_s.data_ptr(), stats.data_ptr(), 0.0,
                               y.data_ptr(), K, B, C, H, W, K, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("rc", rc)
pix = img.float().cpu().view(B, C, H // 16, 16, W // 16, 16).permute(0, 2, 4, 1, 3, 5).reshape(B * n, K)   # (c p1 p2)
got = y.cpu()
print("exact rows:", int((got == pix).all(1).sum()), "of", B * n)
for r in (0, 1, n):
    bad = (got[r] != pix[r]).nonzero().flatten().tolist()
    print(f"row {r}: {len(bad)} wrong columns; first {bad[:12]}")
    for k in bad[:6]:
        v = got[r, k].item()
        where = [(rr, kk) for rr in range(B * n) for kk in (pix[rr] == v).nonzero().flatten().tolist()][:4]
        print(f"   col {k} (c{k // 256} p1 {k % 256 // 16} p2 {k % 16}) = {v}: expected {pix[r, k].item()}; value found at (row, col) {where}")
