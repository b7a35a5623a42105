#!/bin/bash
# does programmatic dependent launch survive graph capture?  graph replay vs eager stream launches, PDL on / off
for nopdl in 0 1; do
  if [ $nopdl = 1 ]; then export EKV_NO_PDL=1; else unset EKV_NO_PDL; fi
  timeout 300 python tools/pdl_probe.py c2_b1 c2_b8 c5 2>&1 | grep workload
done
