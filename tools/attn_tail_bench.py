"""b200vit_attention at the ViT-H/14 shape (batch 128, N = 257, 16 heads, dim_head 64 / 80): tile-only path against
the key-tail path (256-key S tile, two CTAs per SM).  L2 flushed between launches."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_pytorch_b200 import _lib  # noqa: E402


def main():
    dev = "cuda"
    L = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    for (B, N, H) in ((128, 257, 16), (128, 256, 16)):
        for dh in (64, 80):
            qkv = torch.randn(B * N, 3 * H * dh, device=dev).bfloat16()
            out = torch.zeros(B * N, H * dh, device=dev, dtype=torch.bfloat16)
            for tails in (0, 1):
                L.b200vit_debug_set(16, tails)
                ts = []
                for i in range(13):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                    e0.record()
                    _lib.attention(qkv, out, B, N, H, dh, dh ** -0.5)
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 3:
                        ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                res[f"B{B}_N{N}_dh{dh}_tails{tails}"] = {"us_median": round(ts[len(ts) // 2], 1), "us_best": round(ts[0], 1)}
            L.b200vit_debug_set(16, 1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
