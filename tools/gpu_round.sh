#!/bin/bash
# One GPU session: parity tests, smoke, per-kernel timings, bench, ncu launch list, ncu full capture of the hot kernels.
#   usage: tools/gpu_round.sh TAG        env: NCU=0 skips the ncu passes, MODELS="vit_b16 navit" extra bench models
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== kernel bench"; timeout 600 python tools/kernel_bench.py > gpurun_out/kernels_${TAG}.json 2> gpurun_out/kernels_${TAG}.err; cat gpurun_out/kernels_${TAG}.json; tail -3 gpurun_out/kernels_${TAG}.err
for m in ${MODELS:-vit_b16}; do
  echo "== bench $m"
  timeout 900 python bench.py --model $m --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_$m.json 2> gpurun_out/bench_${TAG}_$m.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_${TAG}_$m.json'))
    print(round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']), ' frac burst', round(d['frac_of_bf16_burst_peak'], 4), d['clocks'])
    for k, v in d['breakdown'].items(): print('   ', k, round(v['ms_per_step'], 3), round(v.get('tflops', v.get('gbps', 0))))
    print('  eager', d.get('gpu_eager_baseline')); print('  cpu', d.get('cpu_baseline'))
except Exception as e: print('bench failed', e)
PY
  tail -3 gpurun_out/bench_${TAG}_$m.err
done
if [ "${NCU:-1}" = "1" ]; then
[ "${NCU_LIST:-1}" = "1" ] && echo "== ncu launch list" && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-eager > gpurun_out/ncu_bench_${TAG}.log 2>&1
tail -3 gpurun_out/launches_${TAG}.csv | cut -c1-300
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-"gemm2|attention"} -s ${KSKIP:-10} -c ${KCOUNT:-6} -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu --no-eager > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
fi
ls -la gpurun_out/ | tail -15
