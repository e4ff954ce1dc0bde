"""Small-batch latency of the ViT-B/16 forward: fused path call by call, fused path as a CUDA graph
(vit_pytorch_b200.graph.GraphedForward), and the unmodified reference in eager bf16 on the same GPU.
Wall-clock per forward including the host side (synchronise, time N calls, synchronise)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
from vit_pytorch_b200 import ViT  # noqa: E402
from vit_pytorch_b200.graph import GraphedForward  # noqa: E402

CFG = dict(image_size=224, patch_size=16, num_classes=1000, dim=768, depth=12, heads=12, mlp_dim=3072)


def wall_ms(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    dev = "cuda"
    torch.manual_seed(0)
    model = ViT(**CFG).eval().to(dev, torch.bfloat16)
    try:
        import vit_pytorch
        ref = vit_pytorch.ViT(**CFG).eval().to(dev, torch.bfloat16)
        ref.load_state_dict(model.state_dict())
    except Exception as e:  # noqa: BLE001
        ref = None
        print("reference unavailable:", e, file=sys.stderr)
    out = {}
    for B in (1, 8, 64, 512):
        img = torch.randn(B, 3, 224, 224, device=dev).bfloat16()
        n = 200 if B <= 64 else 20
        with torch.inference_mode():
            os.environ["B200VIT_HOST_LOOP"] = "python"          # one ctypes call per kernel (the round-2 mid state)
            fused_py = wall_ms(lambda: model(img), n)
            os.environ["B200VIT_HOST_LOOP"] = "c"               # all layers in one b200vit_encoder_blocks call (default)
            fused = wall_ms(lambda: model(img), n)
            g = GraphedForward(model, img)
            same = bool(torch.equal(g(img), model(img)))
            graphed = wall_ms(lambda: g(img), n)
            eager = wall_ms(lambda: ref(img), n) if ref is not None else None
        out[B] = {"fused_ms": round(fused, 3), "fused_python_loop_ms": round(fused_py, 3), "fused_graph_ms": round(graphed, 3),
                  "reference_eager_ms": None if eager is None else round(eager, 3), "graph_bit_identical": same}
        print(B, out[B], file=sys.stderr, flush=True)
    print(json.dumps({"model": "ViT-B/16 224^2 bf16 forward, wall clock per call", "batches": out}))


if __name__ == "__main__":
    main()
