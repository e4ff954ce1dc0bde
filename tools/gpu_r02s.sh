set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for dh in 64 80; do
echo "== bench vit_h14 dh $dh"; timeout 600 python bench.py --model vit_h14 --dim-head $dh --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02s_vit_h14_dh$dh.json 2> gpurun_out/bench_r02s_h14_$dh.err; cut -c1-200 gpurun_out/bench_r02s_vit_h14_dh$dh.json; tail -2 gpurun_out/bench_r02s_h14_$dh.err
done
