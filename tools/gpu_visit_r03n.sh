#!/bin/bash
for g in 1 2 4 8; do for n in 260 333 2304; do
  H=$((2*g))
  r=$(timeout 120 python tools/stream_probe.py 2 $H 2 $n float16 3 1 2>&1 | grep -v "^$" | tail -1 | cut -c1-60)
  echo "G=$g n=$n C=1: $r"
done; done
for g in 2 4; do for n in 333 2304; do H=$((2*g)); r=$(timeout 120 python tools/stream_probe.py 2 $H 2 $n float16 3 -1 2>&1 | grep -v "^$" | tail -1 | cut -c1-60); echo "persistent G=$g n=$n: $r"; done; done
