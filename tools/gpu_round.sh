#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list, ncu full capture of the dominant kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench_${TAG}.log 2>&1
tail -3 gpurun_out/launches_${TAG}.csv | cut -c1-300
echo "== ncu full (gemm)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-gemm} -s ${KSKIP:-8} -c ${KCOUNT:-5} -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out/
fi
