#!/bin/bash
# gpurun --gpus N -- bash tools/c5_sync_sweep.sh N : sweep of the C5 training step's gradient-sync schedules (tools/c5_train_sync_ab.py)
N=${1:-2}
out=gpurun_out/r02_c5_sync_sweep_n$N.log
run() { env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) tools/c5_train_sync_ab.py 2>/dev/null | grep "N=" >> $out; }
for cfg in "${@:2}"; do run $cfg; done
tail -n $(( $# - 1 )) $out
