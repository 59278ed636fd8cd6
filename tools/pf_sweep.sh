#!/bin/bash
# A/B sweep of the persistent forward's stream skew (DLADMM_PF_SKEW_US; 0 = one stream) and of the per-layer schedule
# (DLADMM_NO_PERSISTENT=1) on the C1 workload:  bash tools/pf_sweep.sh > gpurun_out/pf_sweep.log
for skew in 0 40 80 100 120 160 200 -1; do
  if [ "$skew" = "-1" ]; then unset DLADMM_PF_SKEW_US; else export DLADMM_PF_SKEW_US=$skew; fi
  for prec in tf32x3 tf32; do
    timeout 120 python bench.py --quick --train-columns 0 --c5-columns 0 --no-cpu-baseline --steps 10 --warmup 3 --precision $prec 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('skew', '$skew', '$prec', 'fwd_ms %.4f' % j['ms_per_step'], 'e2e_ms %.4f' % j['e2e']['ms_per_step'])"
  done
done
unset DLADMM_PF_SKEW_US
DLADMM_NO_PERSISTENT=1 timeout 120 python bench.py --quick --train-columns 0 --c5-columns 0 --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('per-layer tf32x3', 'fwd_ms %.4f' % j['ms_per_step'], 'e2e_ms %.4f' % j['e2e']['ms_per_step'])"
