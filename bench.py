#!/usr/bin/env python
"""Benchmark of the hot path: bf16 forward of ViT-B/16 224^2 (BASELINE.json configs[1]), images/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 512]            # this repo's fused sm_100a path
    python bench.py --impl reference [...]                                       # the reference algorithm on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...             # one rank per GPU (weak scaling)

One "step" = one forward of a synthetic [batch, 3, 224, 224] bf16 batch per GPU (+ the all-gather of the logits when
N > 1).  Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VIT_B16 = dict(image_size=224, patch_size=16, num_classes=1000, dim=768, depth=12, heads=12, mlp_dim=3072)
METRIC = "images/sec ViT-B/16 224^2 bf16 fwd"
# other BASELINE.json configs, selectable with --model (the headline metric stays ViT-B/16)
MODELS = {
    "vit_b16": VIT_B16,
    "vit_l16": dict(image_size=224, patch_size=16, num_classes=1000, dim=1024, depth=24, heads=16, mlp_dim=4096),
    "vit_h14": dict(image_size=224, patch_size=14, num_classes=1000, dim=1280, depth=32, heads=16, mlp_dim=5120),
}


class stdout_to_stderr:
    """fd-level redirect: library chatter (e.g. NCCL's version banner) must not pollute the one-JSON-line stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index = index
        self.proc = None
        self.path = os.path.join(ROOT, "gpurun_out", f"clocks_{index}.csv")

    def start(self) -> None:
        try:
            os.makedirs(os.path.dirname(self.path), exist_ok=True)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()          # exact PID we started
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_forward_factory(sample_batch: int, cfg: dict = None):
    """Returns (fn, dtype name, threads): fn() runs one forward of `sample_batch` images with the oracle port.

    Be generous to the baseline: pick the faster arithmetic (bf16 wins on AMX hosts) and the best intra-op thread
    count for this host (all cores is often NOT the fastest for a 16-image batch)."""
    from oracle import vit_oracle as O
    from vit_pytorch_b200 import ViT
    cfg = cfg or VIT_B16
    torch.manual_seed(0)
    model = ViT(**cfg).eval()                           # identical init stream to the reference constructor
    torch.manual_seed(1)
    img = torch.randn(sample_batch, 3, cfg["image_size"], cfg["image_size"])
    cands = {}
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        sd = O.upcast(model.state_dict(), dt)
        x = img.to(dt)
        cands[name] = (lambda sd=sd, x=x: O.vit_forward(sd, cfg, x))
    cores = os.cpu_count() or 1
    threads = sorted({t for t in (8, 16, 32, 64, cores) if t <= cores})
    best = (None, None, None)
    budget_end = time.perf_counter() + 60.0
    with torch.inference_mode():
        for nt in threads:
            torch.set_num_threads(nt)
            for name, fn in cands.items():
                if time.perf_counter() > budget_end and best[0] is not None:
                    break
                fn()
                t0 = time.perf_counter(); fn(); t = time.perf_counter() - t0
                if best[0] is None or t < best[0]:
                    best = (t, name, nt)
    torch.set_num_threads(best[2])
    return cands[best[1]], best[1], best[2]


def run_reference_arm(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    sb = args.cpu_batch
    fn, dt, cores = cpu_forward_factory(sb)
    with torch.inference_mode():
        for _ in range(max(1, min(args.warmup, 2))):
            fn()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        dt_s = time.perf_counter() - t0
    val = sb * args.steps / dt_s
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": {"workload": "ViT-B/16 224^2 forward, batch 512 per GPU (BASELINE.json configs[1])",
                   "sample": f"{sb} images per step on the host CPU"},
        "cpu_baseline": {"value": val, "unit": "images/sec", "cores": cores, "kind": "port",
                         "sample": f"oracle/vit_oracle.py (torch CPU ops, {dt}) on {sb}-image batches, "
                                   f"{args.steps} steps, {cores} threads (best of 8..{os.cpu_count()})"},
        "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def run_gpu_arm(args, rank: int, local_rank: int, world: int) -> None:
    import torch.distributed as dist
    from oracle import vit_oracle as O
    from vit_pytorch_b200 import ViT, _lib
    from vit_pytorch_b200.parallel import all_gather_logits

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()                       # NCCL prints its version banner on the first collective
    if not _lib.device_ok(local_rank):
        raise SystemExit("bench.py: libb200vit.so cannot run on this device: " +
                         _lib.lib().b200vit_last_error().decode())

    B = args.batch
    CFG = MODELS[args.model]
    torch.manual_seed(0)
    model = ViT(**CFG).eval().to(dev, torch.bfloat16)
    torch.manual_seed(1 + rank)
    img = torch.randn(B, 3, 224, 224, device=dev).bfloat16()
    with torch.inference_mode():
        assert model.fused_reason(img) is None, model.fused_reason(img)

    def step(x):
        with torch.inference_mode():
            out = model(x)
            return all_gather_logits(out) if world > 1 else out

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W = max(args.warmup, 3)
    for _ in range(W):
        step(img)
    sync_all()

    # ---- device-resident timing (the `value`) ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        out = step(img)
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    sync_all()
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end through the public API with host buffers (H2D of the batch + D2H of the logits every step) ----
    from vit_pytorch_b200.io import DeviceFeeder
    host_img = torch.empty(B, 3, 224, 224, dtype=torch.bfloat16).pin_memory()
    host_img.copy_(img)
    host_out = torch.empty(world * B, CFG["num_classes"], dtype=torch.bfloat16).pin_memory()
    feeder = DeviceFeeder(tuple(img.shape), torch.bfloat16, dev)

    def e2e_step():
        # every step copies its own batch host -> device (the copy of step i+1 overlaps the forward of step i)
        x = feeder.push(host_img, next_host=host_img)
        o = step(x)
        feeder.done(x)
        host_out.copy_(o, non_blocking=True)

    for _ in range(3):
        e2e_step()
    sync_all()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    sync_all()
    e2e = {"value": world * B * args.steps / (ms_e2e / 1e3), "unit": "images/sec",
           "h2d_bytes_per_step": host_img.numel() * 2 * world, "d2h_bytes_per_step": host_out.numel() * 2 * world}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): CUDA events around every launch of one more K steps ----
    peaks, peak_src = read_peaks()
    _lib.profile_start()
    for _ in range(min(args.steps, 5)):
        with torch.inference_mode():
            model(img)
    rec = _lib.profile_stop()
    nprof = min(args.steps, 5)
    by = {}
    for name, meta, ms in rec:
        d = by.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += ms; d["flops"] += meta.get("flops", 0.0); d["bytes"] += meta.get("bytes", 0.0); d["launches"] += 1
    tot_ms = sum(d["ms"] for d in by.values())
    breakdown = {k: {"ms_per_step": v["ms"] / nprof, "share": v["ms"] / tot_ms, "launches_per_step": v["launches"] // nprof,
                     **({"tflops": v["flops"] / v["ms"] / 1e9} if v["flops"] else {}),
                     **({"gbps": v["bytes"] / v["ms"] / 1e6} if v["bytes"] and not v["flops"] else {})}
                 for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])}
    g = by["gemm"]
    achieved = g["flops"] / g["ms"] / 1e9                     # TFLOP/s over the GEMM launches of the step
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))   # kernel timed inside a long step
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "gemm2_kernel (tcgen05 cta_group::2)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "peak_source": f"{peak_src} sustained",
                "avg_launch_ms": g["ms"] / g["launches"], "launches_per_step": g["launches"] // nprof}

    flops_img = O.flops_per_image(**CFG)
    tf = value / world * flops_img / 1e12
    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        fn, dt, cores = cpu_forward_factory(args.cpu_batch)
        with torch.inference_mode():
            t0 = time.perf_counter(); n = 0
            while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 50):
                fn(); n += 1
            dt_s = time.perf_counter() - t0
        cpu_baseline = {"value": args.cpu_batch * n / dt_s, "unit": "images/sec", "cores": cores, "kind": "port",
                        "sample": f"oracle/vit_oracle.py (torch CPU ops, {dt}), {n} forwards of {args.cpu_batch} images"}

    line = {
        "metric": METRIC if args.model == "vit_b16" else METRIC.replace("ViT-B/16", args.model), "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("ViT-B/16 224^2 forward (BASELINE.json configs[1])" if args.model == "vit_b16"
                                else f"{args.model} 224^2 forward, dim_head 64"), "batch_per_gpu": B,
                   "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2_policy": "inputs (154 MB/step) and activations (GBs/step) exceed the 126 MB L2",
                   "ln_mode": os.environ.get("B200VIT_LN_MODE", "fold"),
                   "weights": "random init, torch.manual_seed(0)"},
        "tflops_per_gpu": tf,
        "frac_of_bf16_burst_peak": tf / float(peaks["bf16_tflops"]),
        "frac_of_bf16_sustained_peak": tf / peak,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "breakdown": breakdown,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=512, help="images per GPU per step (weak scaling)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=16, help="images per CPU-baseline forward (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--model", default="vit_b16", choices=sorted(MODELS), help="vit_b16 is the headline config")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
