#!/bin/bash
# quick perf/correctness loop: selected bring-up cases, gpu tests, 1-GPU bench in both LN modes
mkdir -p gpurun_out
python tools/bringup.py ${CASES:-g2_ gemm_big attn_tmem_big} 2>&1 | python tools/bringup_brief.py
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
for mode in exact fold; do
  B200VIT_LN_MODE=$mode timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${TAG:-q}_$mode.json 2> gpurun_out/bench_${TAG:-q}_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG:-q}_$mode.json')); print('$mode', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['power_w_max']); [print('  ',k, round(v['ms_per_step'],3), round(v.get('tflops',v.get('gbps',0)))) for k,v in d['breakdown'].items()]"
  tail -2 gpurun_out/bench_${TAG:-q}_$mode.err
done
