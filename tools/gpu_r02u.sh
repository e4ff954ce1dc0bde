set -u
mkdir -p gpurun_out
echo "== pytest attention kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" 2>&1 | tail -5
echo "== attn tail bench"; timeout 300 python tools/attn_tail_bench.py 2>&1 | tail -2
echo "== small batch profile"; timeout 300 python tools/small_batch_profile.py 2>&1 | tail -1 > gpurun_out/small_batch_r02u.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/small_batch_r02u.json'))
for B,v in d.items():
    print(B, v['total_us'], {k:x for k,x in list(v.items())[:9]})
PY
