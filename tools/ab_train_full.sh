#!/bin/bash
# before changing a kernel: mkdir -p tools/ab && cp d-ladmm_b200/csrc/libdladmm.so tools/ab/libdladmm_prev.so  (git-ignored, travels with gpurun)
# same-box A/B of two builds (tools/ab/libdladmm_prev.so through DLADMM_LIB_PATH against the in-tree one): the three training legs
for i in 1 2; do
  for lib in prev new; do
    if [ $lib = prev ]; then export DLADMM_LIB_PATH=$PWD/tools/ab/libdladmm_prev.so; else unset DLADMM_LIB_PATH; fi
    python bench.py --no-cpu-baseline --c5-columns 0 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read())
print('$lib', ' | '.join('%s %.3f ms (dV %.4f)' % (k, l[k]['ms_per_step'], l[k]['roofline']['per_kernel']['bwd_gemm_dv']['avg_launch_ms']) for k in ('train','train_full','train_lasso')))"
  done
done
