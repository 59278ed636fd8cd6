#!/bin/bash
# small-batch forward: 32-row tiles (default below ~4.7 K columns) against the 256-row tiles, and the scout-lane poll pause at C1
for B in 24 256 1024 4096; do
  for tn in 256 32; do
    echo "== B=$B DLADMM_PF_TN=$tn"; DLADMM_PF_TN=$tn python tools/host_overhead.py $B 2>&1 | grep "B=" 
  done
done
for sl in 0 200 1000 0 200; do
  echo "== C1 scout sleep $sl"; DLADMM_PF_SCOUT_SLEEP=$sl python - <<P
import sys; sys.path.insert(0, "tools")
import mix_check
mix_check.timing("tf32_bf16x2")
P
done
