set -u
mkdir -p gpurun_out
echo "== pytest attention kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" 2>&1 | tail -12
echo "== attn tail bench"; timeout 300 python tools/attn_tail_bench.py 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
for dh in 64 80; do
echo "== bench vit_h14 dh $dh"; timeout 600 python bench.py --model vit_h14 --dim-head $dh --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02v_vit_h14_dh$dh.json 2> gpurun_out/bench_r02v_h14_$dh.err; cut -c1-400 gpurun_out/bench_r02v_vit_h14_dh$dh.json; tail -2 gpurun_out/bench_r02v_h14_$dh.err
done
