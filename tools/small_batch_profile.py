"""Per-kernel device time of one ViT-B/16 forward at small batches (CUDA events around every library call)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_pytorch_b200 import ViT, _lib  # noqa: E402

CFG = dict(image_size=224, patch_size=16, num_classes=1000, dim=768, depth=12, heads=12, mlp_dim=3072)


def main():
    torch.manual_seed(0)
    m = ViT(**CFG).eval().to("cuda", torch.bfloat16)
    out = {}
    for B in (1, 2, 4, 8, 16):
        img = torch.randn(B, 3, 224, 224, device="cuda").bfloat16()
        with torch.inference_mode():
            for _ in range(3):
                m(img)
            _lib.profile_start()
            m(img)
            rec = _lib.profile_stop()
        by = {}
        for name, meta, t in rec:
            key = name if name != "gemm" else f"gemm N{meta['N']} K{meta['K']} f{meta['flags']}"
            by.setdefault(key, [0, 0.0])
            by[key][0] += 1
            by[key][1] += t
        out[B] = {k: [n, round(t * 1e3, 1)] for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])}
        out[B]["total_us"] = round(sum(t for _, _, t in rec) * 1e3, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
