#!/bin/bash
# runs every tools/probe_* variant: correctness at B=1000, timing at the two C1 product shapes
for p in tools/probe_*; do
  [ -x "$p" ] || continue
  echo "=== $p"
  timeout 60 $p 500 250 1000 0 | grep -v "^probe"
  timeout 60 $p 250 500 65536 20 | grep "NPASS=3"
  timeout 60 $p 500 250 65536 20 | grep "NPASS=3"
done
