#!/bin/bash
# synccheck of the final library on the families whose tails changed last (warp shuffles in the rank count, PDL)
set -u
OUT=gpurun_out; mkdir -p $OUT; LOG=$OUT/r04o_sanitize.txt; : > $LOG
CS=/usr/local/cuda/bin/compute-sanitizer
echo "== synccheck: cluster decode_umma general persistent" | tee -a $LOG
timeout 900 $CS --tool synccheck --error-exitcode 9 python tools/sanitize_targets.py cluster decode_umma general persistent > $OUT/.san.tmp 2>&1
rc=$?
grep -E "^ok |library launches|ERROR SUMMARY|Error|error|Traceback|assert" $OUT/.san.tmp | head -30 | tee -a $LOG
echo "rc=$rc" | tee -a $LOG
rm -f $OUT/.san.tmp
