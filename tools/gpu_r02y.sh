set -u
mkdir -p gpurun_out
echo "== pytest (navit, variants, model)"; timeout 900 python -m pytest tests/test_gpu_navit.py tests/test_gpu_variants.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -15
for m in 0 2 0 2; do
  echo "== navit_bench varlen mode $m"; VARLEN_MODE=$m timeout 300 python tools/navit_bench.py 2>&1 | tail -1
done
echo "== bench navit"; timeout 600 python bench.py --model navit --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02y_navit.json 2> gpurun_out/bench_r02y_navit.err; cut -c1-1200 gpurun_out/bench_r02y_navit.json; tail -2 gpurun_out/bench_r02y_navit.err
