#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (memcheck + racecheck + synccheck).
# Run as: gpurun --timeout 2400 -- 'bash tools/sanitize.sh'   -> gpurun_out/r02_sanitize.txt
# Every leg prints how many library launches it ran; a leg that ran nothing FAILS (round 1's racecheck leg selected
# zero tests and read as clean).
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
LOG=$OUT/r02_sanitize.txt
: > $LOG
fail=0
leg() {   # leg <tool> <family...>
  local tool=$1; shift
  echo "== $tool: $*" | tee -a $LOG
  timeout 900 $CS --tool $tool --error-exitcode 9 python tools/sanitize_targets.py "$@" > $OUT/.san.tmp 2>&1
  local rc=$?
  grep -E "^ok |library launches|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error" $OUT/.san.tmp | head -40 | tee -a $LOG
  echo "rc=$rc" | tee -a $LOG
  if [ $rc -ne 0 ] || ! grep -q "library launches in this run: [1-9]" $OUT/.san.tmp; then fail=1; echo "LEG FAILED" | tee -a $LOG; fi
}
leg memcheck persistent steady cluster decode_umma chunk_umma chunk_tc general
leg racecheck persistent steady
leg racecheck cluster
leg racecheck decode_umma
leg racecheck chunk_umma
leg racecheck chunk_tc
leg synccheck persistent steady cluster decode_umma chunk_umma chunk_tc
echo "== memcheck: smoke" | tee -a $LOG
timeout 600 $CS --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee -a $LOG
rm -f $OUT/.san.tmp
echo "sanitize overall: $([ $fail -eq 0 ] && echo PASS || echo FAIL)" | tee -a $LOG
exit $fail
