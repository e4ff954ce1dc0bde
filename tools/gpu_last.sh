set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_r02_last.log 2>&1; tail -3 gpurun_out/pytest_r02_last.log
grep -E "fused max|vs fp32 oracle|trained-like|navit config-5|dh80 B" gpurun_out/pytest_r02_last.log > gpurun_out/parity_r02_last.txt; wc -l gpurun_out/parity_r02_last.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench default (driver defaults)"; timeout 900 python bench.py > gpurun_out/bench_r02_last.json 2> gpurun_out/bench_r02_last.err; cut -c1-200 gpurun_out/bench_r02_last.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_last.json'))
print('e2e', round(d['e2e']['value']), 'traffic', d['roofline']['traffic'], 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], d['clocks'])
print('eager', d['gpu_eager_baseline']['value'], 'cpu', d['cpu_baseline']['value'])
PY
tail -2 gpurun_out/bench_r02_last.err
