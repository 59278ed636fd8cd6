#!/bin/bash
# A/B of persistent-forward switches on the C1 workload: bash tools/pf_ab.sh "VAR=val VAR2=val" ...   (one run per argument)
run() {
  env $1 timeout 120 python bench.py --quick --train-columns 0 --c5-columns 0 --no-cpu-baseline --steps 10 --warmup 3 --precision $2 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$1', '$2', 'fwd_ms %.4f' % j['ms_per_step'], 'e2e_ms %.4f' % j['e2e']['ms_per_step'])"
}
for cfg in "$@"; do
  for prec in tf32x3 tf32; do run "$cfg" $prec; done
done
