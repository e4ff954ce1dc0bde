set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== attn tail bench"; timeout 300 python tools/attn_tail_bench.py 2>&1 | tail -2
echo "== small batch profile"; timeout 300 python tools/small_batch_profile.py 2>&1 | tail -1 > gpurun_out/small_batch_r02t.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/small_batch_r02t.json'))
for B,v in d.items():
    print(B, v['total_us'], {k:x for k,x in list(v.items())[:6]})
PY
echo "== latency"; timeout 600 python tools/latency_bench.py > gpurun_out/latency_r02t.json 2> gpurun_out/latency_r02t.err; cat gpurun_out/latency_r02t.json
for dh in 64 80; do
echo "== bench vit_h14 dh $dh"; timeout 600 python bench.py --model vit_h14 --dim-head $dh --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02t_vit_h14_dh$dh.json 2> gpurun_out/bench_r02t_h14_$dh.err; cut -c1-200 gpurun_out/bench_r02t_vit_h14_dh$dh.json; tail -2 gpurun_out/bench_r02t_h14_$dh.err
done
