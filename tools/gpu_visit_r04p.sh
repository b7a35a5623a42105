#!/bin/bash
# the bucket select under the sanitizer, on a state built to reject the candidate walk
set -u
OUT=gpurun_out; mkdir -p $OUT; LOG=$OUT/r04p_sanitize.txt; : > $LOG
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 120 python tools/sanitize_targets.py bucket bucket_umma 2>&1 | tail -5 | tee -a $LOG
for leg in "memcheck bucket bucket_umma" "racecheck bucket" "synccheck bucket bucket_umma"; do
  set -- $leg; tool=$1; shift
  echo "== $tool: $*" | tee -a $LOG
  timeout 400 $CS --tool $tool --error-exitcode 9 python tools/sanitize_targets.py "$@" > $OUT/.san.tmp 2>&1
  rc=$?
  grep -E "^ok |library launches|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error|Traceback|assert" $OUT/.san.tmp | head -20 | tee -a $LOG
  echo "rc=$rc" | tee -a $LOG
done
rm -f $OUT/.san.tmp
