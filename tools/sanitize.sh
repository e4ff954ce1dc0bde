#!/bin/bash
# compute-sanitizer over small invocations of every kernel (memcheck + synccheck + racecheck).
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, torch, math
sys.path.insert(0, ".")
from vit_pytorch_b200 import _lib, ViT, SimpleViT
import os
def load_golden(name):
    return torch.load(os.path.join("tests", "golden", name + ".pt"), weights_only=False)
torch.manual_seed(0)
dev = "cuda"
# GEMMs: single-CTA kernel and CTA-pair kernel in all epilogue modes (small shapes)
for force, M, N, K in ((1, 200, 264, 128), (2, 512, 512, 128)):
    _lib.lib().b200vit_debug_set(4, force)
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev); x = torch.randn(M, N, device=dev)
    ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    _lib.gemm(a, w, out_bf16=ob, bias=b, gelu=True)
    _lib.gemm(a, w, out_f32=x, bias=b, resid=x)
    if N % 64 == 0:
        st = torch.zeros(M, _lib.stats_parts(N), 2, device=dev)
        _lib.gemm(a, w, out_f32=x, out_bf16=ob, bias=b, resid=x, stats_out=st)
        sums = torch.zeros(M, 1, 2, device=dev); xb = torch.zeros(M, K, device=dev, dtype=torch.bfloat16)
        _lib.rowstats_cast(a.float(), xb, sums)
        _lib.gemm(xb, w, out_bf16=ob, bias=b, ln_sums=sums, col_s=w.float().sum(1).contiguous())
_lib.lib().b200vit_debug_set(4, 0)
# attention variants
for mode, (B, N, H) in ((0, (2, 197, 2)), (1, (2, 197, 2)), (0, (1, 300, 1)), (0, (2, 50, 3))):
    _lib.lib().b200vit_debug_set(1, mode)
    qkv = torch.randn(B * N, 3 * H * 64, device=dev).bfloat16(); o = torch.zeros(B * N, H * 64, device=dev, dtype=torch.bfloat16)
    _lib.attention(qkv, o, B, N, H, 64, 0.125)
_lib.lib().b200vit_debug_set(1, 0)
# whole models (all row kernels)
for name in ("simplevit_tiny", "vit_tiny_cls", "vit_tiny_mean_nonsquare"):
    g = load_golden(name)
    m = (ViT if g["kind"] == "vit" else SimpleViT)(**g["kwargs"]).eval(); m.load_state_dict(g["state_dict"]); m = m.to(dev, torch.bfloat16)
    with torch.inference_mode():
        out = m(g["input"].to(dev))
    err = (out.float().cpu() - g["logits_fp32"]).abs().max().item()
    print(name, "max err", err)
torch.cuda.synchronize()
print("SANITIZE_CASES_DONE")
PY
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASES_DONE|max err|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
