#!/bin/bash
# compute-sanitizer passes over tools/sanitize_target.py (run on the GPU box: gpurun -- 'bash tools/sanitize.sh').
# One log per tool under gpurun_out/sanitizer/; the summaries that matter are copied to profiles/ by hand.
set -u
OUT=${1:-gpurun_out/sanitizer}
mkdir -p "$OUT"
CS=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
rc_all=0
TOOLS=${TOOLS:-"memcheck racecheck synccheck initcheck"}
TMO=${TMO:-1500}
for tool in $TOOLS; do
  extra=""
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  :
  args=""
  # racecheck serialises shared-memory accesses: keep its workload to the small shapes
  [ "$tool" = racecheck ] && [ -n "${SMALL:-}" ] && args="--small"
  start=$(date +%s)
  timeout "$TMO" "$CS" --tool "$tool" $extra --print-limit 40 --error-exitcode 7 python tools/sanitize_target.py $args \
      > "$OUT/$tool.log" 2>&1
  rc=$?
  echo "$tool rc=$rc seconds=$(( $(date +%s) - start ))" | tee -a "$OUT/summary.txt"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok" "$OUT/$tool.log" | tee -a "$OUT/summary.txt"
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
