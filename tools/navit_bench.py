"""NaViT (BASELINE.json configs[4]) timing on one GPU: the synthetic packed batch of SURVEY.md 8d -- 256 images,
H, W = 16 * randrange(4, 33), dim 1024, depth 6, heads 16, mlp 4096 -- through the padding-free fused path."""
import json
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_pytorch_b200 import NaViT, _lib  # noqa: E402


def main():
    dev = "cuda"
    mode = int(os.environ.get("VARLEN_MODE", "0"))      # varlen attention kernel (b200vit_debug_set key 11)
    _lib.lib().b200vit_debug_set(11, mode)
    kwargs = dict(image_size=512, patch_size=16, num_classes=1000, dim=1024, depth=6, heads=16, mlp_dim=4096)
    torch.manual_seed(0)
    m = NaViT(**kwargs).eval().to(dev, torch.bfloat16)
    random.seed(0)
    sizes = [(16 * random.randrange(4, 33), 16 * random.randrange(4, 33)) for _ in range(256)]
    torch.manual_seed(1)
    imgs = [torch.randn(3, h, w, device=dev).bfloat16() for h, w in sizes]
    tokens = sum((h // 16) * (w // 16) for h, w in sizes)
    with torch.inference_mode():
        assert m.fused_reason(imgs) is None
        for _ in range(3):
            out = m(imgs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            out = m(imgs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        _lib.profile_start()
        m(imgs)
        rec = _lib.profile_stop()
    by = {}
    for name, meta, t in rec:
        by[name] = by.get(name, 0.0) + t
    gemm_flops = sum(meta.get("flops", 0.0) for name, meta, t in rec if name == "gemm")
    attn_flops = 6 * sum(4.0 * 16 * ((h // 16) * (w // 16)) ** 2 * 64 for h, w in sizes)
    print(json.dumps({"varlen_mode": mode, "workload": "NaViT config 5: 256 images, %d tokens, padding-free" % tokens, "ms": ms,
                      "images_per_s": 256 / ms * 1e3, "tokens_per_s": tokens / ms * 1e3,
                      "tflops_algorithmic": (gemm_flops + attn_flops) / ms / 1e9,
                      "finite": bool(torch.isfinite(out.float()).all()),
                      "breakdown_ms": {k: round(v, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])}}))


if __name__ == "__main__":
    main()
