set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench_r02_end.json 2> gpurun_out/bench_r02_end.err; cut -c1-300 gpurun_out/bench_r02_end.json; tail -2 gpurun_out/bench_r02_end.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_r02_end_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r02_end_reference.json
for m in vit_l16 navit; do
echo "== bench $m"; timeout 900 python bench.py --model $m --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r02_end_$m.json 2> gpurun_out/bench_r02_end_$m.err; cut -c1-250 gpurun_out/bench_r02_end_$m.json; tail -1 gpurun_out/bench_r02_end_$m.err
done
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_end.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-eager > gpurun_out/ncu_bench_r02_end.log 2>&1
echo "== ncu full vit_b16"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm2|attention|gemm_bf16" -s 6 -c 8 -f -o gpurun_out/prof_r02_end python bench.py --steps 1 --warmup 3 --no-cpu --no-eager > gpurun_out/ncu_full_r02_end.log 2>&1; tail -1 gpurun_out/ncu_full_r02_end.log
echo "== ncu full navit varlen"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"varlen2|attn_pool" -s 2 -c 3 -f -o gpurun_out/prof_r02_end_navit python bench.py --model navit --steps 1 --warmup 3 --no-cpu --no-eager > gpurun_out/ncu_full_r02_end_navit.log 2>&1; tail -1 gpurun_out/ncu_full_r02_end_navit.log
echo "== ncu full h14 attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attention_kernel" -s 2 -c 2 -f -o gpurun_out/prof_r02_end_h14 python bench.py --model vit_h14 --dim-head 80 --steps 1 --warmup 3 --no-cpu --no-eager > gpurun_out/ncu_full_r02_end_h14.log 2>&1; tail -1 gpurun_out/ncu_full_r02_end_h14.log
ls -la gpurun_out/*.ncu-rep | tail -4
