#!/bin/bash
# compute-sanitizer memcheck over the NaViT kernels (p = 16 patchify, embed_varlen, GEMM + head-RMSNorm epilogue in both
# GEMM kernels, pipelined varlen attention with multi-block sequences, pooling) and the 16x16x3 ViT patchify.
mkdir -p gpurun_out
cat > /tmp/san_navit.py <<'PY'
import sys, random, torch
sys.path.insert(0, ".")
from vit_pytorch_b200 import _lib, NaViT, ViT
dev = "cuda"
torch.manual_seed(0)
kw = dict(image_size=256, patch_size=16, num_classes=10, dim=256, depth=1, heads=4, mlp_dim=512)
m = NaViT(**kw).eval().to(dev, torch.bfloat16)
random.seed(2)
sizes = [(16 * random.randrange(1, 17), 16 * random.randrange(1, 17)) for _ in range(12)] + [(256, 256), (16, 16)]
imgs = [torch.randn(3, h, w, device=dev).bfloat16() for h, w in sizes]
print("tokens", sum((h // 16) * (w // 16) for h, w in sizes))
for mode in ("fold", "exact"):
    import os
    os.environ["B200VIT_LN_MODE"] = mode
    with torch.inference_mode():
        assert m.fused_reason(imgs) is None
        out = m(imgs)
    print(mode, "finite", bool(torch.isfinite(out.float()).all()))
v = ViT(image_size=64, patch_size=16, num_classes=10, dim=128, depth=1, heads=2, mlp_dim=256).eval().to(dev, torch.bfloat16)
with torch.inference_mode():
    o = v(torch.randn(3, 3, 64, 64, device=dev).bfloat16())
print("vit finite", bool(torch.isfinite(o.float()).all()))
torch.cuda.synchronize()
print("SANITIZE_CASES_DONE")
PY
timeout ${SAN_TIMEOUT:-200} compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san_navit.py > gpurun_out/sanitizer_navit_memcheck.log 2>&1
grep -E "ERROR SUMMARY|SANITIZE_CASES_DONE|finite|tokens|Invalid|Error" gpurun_out/sanitizer_navit_memcheck.log | head -12
