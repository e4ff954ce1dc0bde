"""Per-kernel timing at the ViT-B/16 batch-512 shapes (or --model vit_l16 / vit_h14 shapes): CUDA events around
single launches, L2 flushed (a 256 MB memset) before every timed launch, median of --reps.

    python tools/kernel_bench.py [--batch 512] [--reps 20] [--only attention,fc2]

Prints one JSON object: {kernel: {us, tflops | gbps, variant...}}.  A/B switches use the library's test hooks
(b200vit_debug_set: attention kernel choice, epilogue-warp count); PDL is a process-wide environment switch
(B200VIT_PDL=0) and does not matter for isolated launches."""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_pytorch_b200 import _lib  # noqa: E402

SHAPES = {"vit_b16": (197, 768, 12, 3072), "vit_l16": (197, 1024, 16, 4096), "vit_h14": (257, 1280, 16, 5120)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="vit_b16", choices=sorted(SHAPES))
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    N, D, H, MLP = SHAPES[args.model]
    B = args.batch
    M, I = B * N, H * 64
    dev = "cuda"
    L = _lib.lib()
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    only = set(filter(None, args.only.split(",")))

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return statistics.median(ts), min(ts)

    out = {}
    rnd = lambda *s: (torch.randn(*s, device=dev) * 0.5).bfloat16()
    xb = rnd(M, D)
    stats1 = torch.zeros(M, 1, 2, device=dev)
    _lib.rowstats_cast(xb.float(), xb, stats1)
    x = torch.randn(M, D, device=dev)
    qkv = rnd(M, 3 * I)
    o = rnd(M, I)
    h = rnd(M, MLP)
    sa = torch.zeros(M, _lib.stats_parts(D), 2, device=dev)

    def add(name, fn, flops=None, nbytes=None, **kw):
        if only and not any(name.startswith(k) for k in only):
            return
        med, best = timed(fn)
        d = {"us": round(med, 1), "best_us": round(best, 1), **kw}
        if flops:
            d["tflops"] = round(flops / med / 1e6, 1)
        if nbytes:
            d["gbps"] = round(nbytes / med / 1e3, 1)
        out[name] = d
        print(name, d, file=sys.stderr, flush=True)

    # ---- attention
    attn_bytes = (qkv.numel() + o.numel()) * 2
    for mode, tag in ((2, "pipe"), (1, "round1")):
        if N > 224 and mode == 2:
            continue
        L.b200vit_debug_set(1, mode)
        add(f"attention_{tag}", lambda: _lib.attention(qkv, o, B, N, H, 64, 0.125), flops=4.0 * B * H * N * N * 64,
            nbytes=attn_bytes)
    L.b200vit_debug_set(1, 0)

    # ---- the four GEMMs of a layer (LN-fold schedule)
    wq, sq, tq = rnd(3 * I, D), torch.randn(3 * I, device=dev), torch.randn(3 * I, device=dev)
    add("qkv", lambda: _lib.gemm(xb, wq, out_bf16=qkv, bias=tq, ln_sums=stats1, col_s=sq), flops=2.0 * M * 3 * I * D)
    wo, bo = rnd(D, I), torch.randn(D, device=dev)
    add("outproj", lambda: _lib.gemm(o, wo, out_f32=x, out_bf16=xb, bias=bo, resid=x, stats_out=sa),
        flops=2.0 * M * D * I, nbytes=M * (I * 2 + D * 10))
    w1, s1, t1 = rnd(MLP, D), torch.randn(MLP, device=dev), torch.randn(MLP, device=dev)
    add("fc1", lambda: _lib.gemm(xb, w1, out_bf16=h, bias=t1, gelu=True, ln_sums=stats1, col_s=s1), flops=2.0 * M * MLP * D)
    w2, b2 = rnd(D, MLP), torch.randn(D, device=dev)
    for ew in (4, 8):
        L.b200vit_debug_set(12, ew)
        add(f"fc2_ew{ew}", lambda: _lib.gemm(h, w2, out_f32=x, out_bf16=xb, bias=b2, resid=x, stats_out=sa),
            flops=2.0 * M * D * MLP, nbytes=M * (MLP * 2 + D * 10))
        add(f"outproj_ew{ew}", lambda: _lib.gemm(o, wo, out_f32=x, out_bf16=xb, bias=bo, resid=x, stats_out=sa),
            flops=2.0 * M * D * I)
    L.b200vit_debug_set(12, 0)
    # cuBLAS on the same shapes (library bar, plain GEMM without the epilogues)
    add("cublas_qkv", lambda: torch.matmul(xb, wq.t()), flops=2.0 * M * 3 * I * D)
    add("cublas_fc2", lambda: torch.matmul(h, w2.t()), flops=2.0 * M * D * MLP)
    print(json.dumps({"model": args.model, "batch": B, "kernels": out}))


if __name__ == "__main__":
    main()
