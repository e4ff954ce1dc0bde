#!/bin/bash
# 2-GPU session: parity tests, 1-GPU bench in both LN modes, 2-GPU torchrun bench, reference arm.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
for mode in exact fold; do
  echo "== bench 1 GPU ($mode)"
  B200VIT_LN_MODE=$mode timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_1gpu_$mode.json 2> gpurun_out/bench_1gpu_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu_$mode.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']); [print('  ',k,v) for k,v in d['breakdown'].items()]"
  tail -2 gpurun_out/bench_1gpu_$mode.err
done
echo "== bench 2 GPUs (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 1500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-700
