set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== latency"; timeout 600 python tools/latency_bench.py > gpurun_out/latency_r02w.json 2> gpurun_out/latency_r02w.err; cat gpurun_out/latency_r02w.json; tail -3 gpurun_out/latency_r02w.err
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02w_vit_b16.json 2> gpurun_out/bench_r02w.err; cut -c1-900 gpurun_out/bench_r02w_vit_b16.json; tail -2 gpurun_out/bench_r02w.err
