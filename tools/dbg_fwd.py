"""Debug helper: forward at a given batch size with per-launch synchronisation (DLADMM_DEBUG_SYNC=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
B = int(sys.argv[1]); K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m, d = 250, 500
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
z = lambda r: torch.zeros(r, B, device="cuda")
model = dl.DLADMMNetScalar(m, 1, d, B, data.A, torch.rand(d, B, device="cuda") / d, z(m), z(m), K)
with torch.no_grad():
    out = model(data.X)
torch.cuda.synchronize()
print("ok", B, out[0][-1].abs().mean().item())
