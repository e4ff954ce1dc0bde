"""Is the CTA-pair GEMM bound by the L2 -> SM operand feed?   (experiment; results are WRONG on purpose)

Times the four GEMMs of a ViT-B/16 block at batch 512 with the weight-tile (B operand) loads thinned out through
debug key 8: feed_skip = s loads the B tile only on every (s+1)-th k block, so s=1 removes 25 % of the operand
traffic (what a 4-CTA cluster with B multicast would remove) and s=1000 removes 50 %.  If the time does not move,
the operand feed is not the limiter.   Usage: python tools/gemm_feed_probe.py   -> gpurun_out/gemm_feed_probe.json
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


# 0..9999: feed_skip value (debug key 8);  < 0: L2 prefetch distance in k blocks (debug key 9);
# 10000 + s: use only s stages of the operand ring (debug key 10).  The last two give correct results.
MODES = [int(a) for a in sys.argv[1:]] or [0, 1000, 3000]


def main():
    import torch
    from vit_pytorch_b200 import _lib
    L = _lib.lib()
    dev = "cuda"
    M = int(os.environ.get("PROBE_M", 512 * 197))
    do_flush = os.environ.get("PROBE_FLUSH", "1") == "1"   # 0: operands stay L2-resident between launches
    torch.manual_seed(0)
    x768 = torch.randn(M, 768, device=dev).bfloat16()
    x3072 = torch.randn(M, 3072, device=dev).bfloat16()
    shapes = {
        "qkv": dict(N=2304, K=768, kind="fold_bf16"),
        "fc1": dict(N=3072, K=768, kind="fold_gelu"),
        "proj": dict(N=768, K=768, kind="dual"),
        "fc2": dict(N=768, K=3072, kind="dual"),
    }
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    results = {}
    for name, sh in shapes.items():
        N, K = sh["N"], sh["K"]
        a = x768 if K == 768 else x3072
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        bias = torch.randn(N, device=dev)
        af = a.float()
        ln = torch.stack([af.sum(1), (af * af).sum(1)], 1).contiguous()
        del af
        col_s = w.float().sum(1).contiguous()
        if sh["kind"] == "dual":
            out_f = torch.randn(M, N, device=dev)
            out_b = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            st = torch.zeros(M, _lib.stats_parts(N), 2, device=dev)
            call = lambda: _lib.gemm(a, w, out_bf16=out_b, out_f32=out_f, bias=bias, resid=out_f, stats_out=st)
        else:
            out_b = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            call = lambda: _lib.gemm(a, w, out_bf16=out_b, bias=bias, gelu=sh["kind"] == "fold_gelu", ln_sums=ln,
                                     col_s=col_s)
        row = {}
        for skip in MODES:
            L.b200vit_debug_set(8, skip if 0 <= skip < 10000 else 0)
            L.b200vit_debug_set(9, -skip if skip < 0 else 0)
            L.b200vit_debug_set(10, skip - 10000 if skip >= 10000 else 0)
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                if do_flush:
                    flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                call()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            us = ts[len(ts) // 2]
            row[f"skip{skip}_us"] = round(us, 1)
            row[f"skip{skip}_tflops"] = round(2.0 * M * N * K / us / 1e6, 1)
        L.b200vit_debug_set(8, 0)
        L.b200vit_debug_set(9, 0)
        L.b200vit_debug_set(10, 0)
        results[name] = row
        print(name, row, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_feed_probe.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
