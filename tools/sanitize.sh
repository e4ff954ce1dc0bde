#!/bin/bash
# compute-sanitizer over tools/sanitize_cases.py: memcheck on everything, racecheck and synccheck per kernel group (so
# that one slow group cannot starve the others of the time budget).  Logs -> gpurun_out/sanitizer_<tool>_<group>.log
#   usage: tools/sanitize.sh [per-run timeout seconds, default 600]
T=${1:-600}
mkdir -p gpurun_out
run() {  # tool group
  local log=gpurun_out/sanitizer_$1_$2.log
  local args=$2
  [ "$2" = all ] && args=""
  timeout $T compute-sanitizer --tool $1 --launch-timeout 120 --print-limit 20 python tools/sanitize_cases.py $args > $log 2>&1
  echo "== $1 $2: exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASES_DONE|max err|cases done|finite|Error|hazard" $log | head -12
}
run memcheck all
for g in gemm attention models navit; do run racecheck $g; done
for g in gemm attention; do run synccheck $g; done
