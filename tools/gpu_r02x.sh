set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for m in 0 0; do
  echo "== navit_bench varlen mode $m"; VARLEN_MODE=$m timeout 300 python tools/navit_bench.py 2>&1 | tail -1
done
echo "== sanitizer"
run() {
  local log=gpurun_out/sanitizer_r02x_$1_$2.log
  local t0=$(date +%s)
  timeout $3 compute-sanitizer --tool $1 --launch-timeout 120 --print-limit 20 python tools/sanitize_cases.py $2 > $log 2>&1
  echo "== $1 $2: exit $? in $(( $(date +%s) - t0 )) s"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|cases done|Error|hazard" $log | head -8
}
run memcheck attention 200
run memcheck navit 200
run racecheck attention 300
run synccheck attention 400
run synccheck navit 300
