set -u
mkdir -p gpurun_out
run() {
  local log=gpurun_out/sanitizer_r02end_$1_$2.log
  local t0=$(date +%s)
  timeout $3 compute-sanitizer --tool $1 --launch-timeout 120 --print-limit 20 python tools/sanitize_cases.py $2 > $log 2>&1
  echo "== $1 $2: exit $? in $(( $(date +%s) - t0 )) s"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|cases done|Error|hazard" $log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head -8
}
run memcheck "" 300
run racecheck attention 300
run racecheck models 300
run synccheck "" 300
