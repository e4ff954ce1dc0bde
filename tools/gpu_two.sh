#!/bin/bash
# 2-GPU sanity of the last build: torchrun bench of the default config and of NaViT (driver-style launch lines).
set -u
mkdir -p gpurun_out
for m in vit_b16 navit; do
echo "== bench 2 GPUs $m"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --model $m --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r02_end_2gpu_$m.json 2> gpurun_out/bench_r02_end_2gpu_$m.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r02_end_2gpu_$m.json'))
    print(round(d['value']), d['n_gpus'], round(d['ms_per_step'],2), 'per rank', d.get('per_rank_ms_per_step'), 'allgather', d.get('allgather_ms'), 'traffic', d['roofline'].get('traffic'), d['clocks']['reasons'])
except Exception as e: print('failed', e)
PY
tail -2 gpurun_out/bench_r02_end_2gpu_$m.err
done
echo "== reference arm under torchrun (rank 0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
