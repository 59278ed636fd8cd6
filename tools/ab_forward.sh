#!/bin/bash
# before changing a kernel: mkdir -p tools/ab && cp d-ladmm_b200/csrc/libdladmm.so tools/ab/libdladmm_prev.so  (git-ignored, travels with gpurun)
# same-box A/B of two builds of the library: tools/ab/libdladmm_prev.so (DLADMM_LIB_PATH) against the in-tree one, C1 forward + training
for i in 1 2 3; do
  for lib in prev new; do
    if [ $lib = prev ]; then export DLADMM_LIB_PATH=$PWD/tools/ab/libdladmm_prev.so; else unset DLADMM_LIB_PATH; fi
    python bench.py --quick --no-cpu-baseline --c5-columns 0 --steps 20 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('$lib', 'fwd %.4f ms' % l['ms_per_step'], 'e2e %.4f' % l['e2e']['ms_per_step'], 'train %.4f' % l['train']['ms_per_step'], 'elem %.4f' % l['train']['library_kernel_ms_per_step']['bwd_elem'], 'clk', l['clocks']['sm_mhz'])"
  done
done
