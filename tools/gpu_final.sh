set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== kernel bench"; timeout 600 python tools/kernel_bench.py > gpurun_out/kernels_r02z.json 2>/dev/null; cat gpurun_out/kernels_r02z.json
for spec in "vit_b16 64" "vit_l16 64" "vit_h14 64" "vit_h14 80" "navit 64"; do
  set -- $spec
  echo "== bench $1 dim_head $2"
  timeout 900 python bench.py --model $1 --dim-head $2 --steps 20 --warmup 5 > gpurun_out/bench_r02z_$1_dh$2.json 2> gpurun_out/bench_r02z_$1_dh$2.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r02z_$1_dh$2.json'))
    print(round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']), ' frac burst', round(d['frac_of_bf16_burst_peak'], 4), 'roofline', round(d['roofline']['frac'], 3), d['clocks'])
    for k, v in d['breakdown'].items(): print('   ', k, round(v['ms_per_step'], 3), round(v.get('tflops', v.get('gbps', 0))))
    print('  eager', d.get('gpu_eager_baseline')); print('  cpu', d.get('cpu_baseline'))
except Exception as e: print('bench failed', e)
PY
  tail -2 gpurun_out/bench_r02z_$1_dh$2.err
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_r02z_reference.json 2>/dev/null; cut -c1-1500 gpurun_out/bench_r02z_reference.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02z.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-eager > gpurun_out/ncu_bench_r02z.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm2|attention|gemm_bf16" -s 6 -c 8 -f -o gpurun_out/prof_r02z python bench.py --steps 1 --warmup 3 --no-cpu --no-eager > gpurun_out/ncu_full_r02z.log 2>&1
tail -2 gpurun_out/ncu_full_r02z.log
