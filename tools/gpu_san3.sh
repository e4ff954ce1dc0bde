set -u
mkdir -p gpurun_out
echo "== pytest attention"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -k "attention or large_configs or h14 or H14" 2>&1 | tail -3
log=gpurun_out/sanitizer_r02end2_racecheck_attention.log
timeout 300 compute-sanitizer --tool racecheck --launch-timeout 120 --print-limit 20 python tools/sanitize_cases.py attention > $log 2>&1
grep -E "RACECHECK SUMMARY|cases done|Race reported" $log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head -8
