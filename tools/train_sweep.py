"""Per-kernel time of one training step (forward + fused loss + backward) over batch sizes (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dladmm_b200 as dl
from dladmm_b200 import _lib
m, d, K = 250, 500, 15
for B in (4096, 16384, 32768, 65536):
    data = dl.gen_syn_data(B, m=m, d=d, seed=1)
    Z0 = torch.rand(d, B, device="cuda") / d
    z = lambda r: torch.zeros(r, B, device="cuda")
    model = dl.DLADMMNetScalar(m, 1, d, B, data.A, Z0, z(m), z(m), K)
    w = [0.2] * (K - 1) + [1.0]
    def step():
        model.zero_grad(set_to_none=True)
        loss, _ = model.l1l1_loss(data.X, 0.001, w)
        loss.backward()
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); step(); e1.record(); torch.cuda.synchronize()
    _lib.profile_start(); step(); prof = _lib.profile_stop()
    print("B=%6d step %.2f ms | " % (B, e0.elapsed_time(e1) / 2) + " ".join("%s=%.2f" % (k, v[0]) for k, v in prof.items() if v[1]), flush=True)
    del model, data
    torch.cuda.empty_cache()
