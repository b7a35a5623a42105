#!/bin/bash
# racecheck of ONE tcgen05 GQA decode case (the family timed out as a whole earlier)
set -u
OUT=gpurun_out; mkdir -p $OUT; LOG=$OUT/r04q_sanitize.txt; : > $LOG
echo "== racecheck: bucket_umma" | tee -a $LOG
timeout 400 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_targets.py bucket_umma > $OUT/.san.tmp 2>&1
rc=$?
grep -E "^ok |library launches|RACECHECK SUMMARY|hazard|Error|error|Traceback|assert" $OUT/.san.tmp | head -20 | tee -a $LOG
echo "rc=$rc" | tee -a $LOG
rm -f $OUT/.san.tmp
