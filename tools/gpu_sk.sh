set -u
mkdir -p gpurun_out
echo "== pytest split-k + new tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x -k "split or trained_like or one_call or gemm" -s 2>&1 | grep -E "passed|failed|split-K B=|trained-like|Error|assert" | head -20
echo "== latency default"; timeout 300 python tools/latency_bench.py 2>/dev/null | tail -1 | cut -c1-600
echo "== latency split-k"; B200VIT_SPLITK=1 timeout 300 python tools/latency_bench.py 2>/dev/null | tail -1 | cut -c1-600
