#!/bin/bash
# memcheck over every case group + racecheck per group (synccheck: tools/sanitize.sh).  usage: sanitize_quick.sh [timeout]
T=${1:-300}
mkdir -p gpurun_out
run() {
  local log=gpurun_out/sanitizer_$1_$2.log
  local args=$2
  [ "$2" = all ] && args=""
  local t0=$(date +%s)
  timeout $T compute-sanitizer --tool $1 --launch-timeout 120 --print-limit 20 python tools/sanitize_cases.py $args > $log 2>&1
  echo "== $1 $2: exit $? in $(( $(date +%s) - t0 )) s"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASES_DONE|max err|cases done|finite|Error|hazard" $log | head -10
}
run memcheck all
for g in gemm attention models navit; do run racecheck $g; done
