#!/bin/bash
# memcheck of what changed last: the head_dim template of the general kernel, the bucket select's classify / rank code
set -u
OUT=gpurun_out; mkdir -p $OUT; LOG=$OUT/r04h_sanitize.txt; : > $LOG
CS=/usr/local/cuda/bin/compute-sanitizer
for fam in "general" "cluster decode_umma"; do
  echo "== memcheck: $fam" | tee -a $LOG
  timeout 700 $CS --tool memcheck --error-exitcode 9 python tools/sanitize_targets.py $fam > $OUT/.san.tmp 2>&1
  rc=$?
  grep -E "^ok |library launches|ERROR SUMMARY|Error|error|Traceback|assert" $OUT/.san.tmp | head -30 | tee -a $LOG
  echo "rc=$rc" | tee -a $LOG
done
rm -f $OUT/.san.tmp
