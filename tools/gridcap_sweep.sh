#!/bin/bash
# per-SM vs chip-wide bound of the C1 forward: the same forward on fewer SMs (DLADMM_GRID_CAP)
for cap in 148 132 116 100 74; do
  for mode in tf32x3 tf32_bf16x2 tf32; do
    DLADMM_GRID_CAP=$cap python - <<P
import sys; sys.path.insert(0, "tools")
import mix_check
mix_check.timing.__defaults__ = (65536, 250, 500, 15, 10)
import torch, os
import dladmm_b200 as dl
mode = "$mode"
B, m, d, K = 65536, 250, 500, 15
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
z = lambda r: torch.zeros(r, B, device="cuda")
model = dl.DLADMMNetScalar(m, 1, d, B, data.A, z(d), z(m), z(m), K, precision=mode)
with torch.no_grad():
    for _ in range(3): out = model(data.X)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out = model(data.X)
    e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print("cap $cap %-12s %.3f ms   ms*SMs/148 = %.3f" % (mode, t, t * $cap / 148))
P
  done
done
