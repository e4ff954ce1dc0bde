#!/bin/bash
# 8-GPU weak-scaling runs of BASELINE.json configs[1..3] (one process per GPU under torchrun, NCCL all-gather of logits)
mkdir -p gpurun_out
run() {  # model dim_head steps
  local out=gpurun_out/scale8_$1_dh$2.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps $3 --warmup 5 --model $1 --dim-head $2 > $out 2> gpurun_out/scale8_$1_dh$2.err
  python - <<PY
import json
try:
    d = json.load(open('$out'))
    pr = d['per_rank_ms_per_step']
    print('$1 dh$2:', round(d['value']), 'img/s on', d['n_gpus'], 'GPUs;', round(d['ms_per_step'], 3), 'ms/step; per rank', [round(x, 2) for x in pr],
          '; all-gather', round(d['allgather_ms'], 4), 'ms; e2e', round(d['e2e']['value']))
except Exception as e:
    print('$1 dh$2 failed:', e)
PY
  tail -2 gpurun_out/scale8_$1_dh$2.err
}
run vit_b16 64 20
run vit_l16 64 20
run vit_h14 64 10
run vit_h14 80 10
