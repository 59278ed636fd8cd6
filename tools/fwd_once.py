"""A few C1 forwards (scalar, m 250, d 500, K 15, 65 536 columns, all iterates): profiling target for the forward kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m, d, K = 250, 500, 15
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
Z0 = torch.rand(d, B, device="cuda") / d
z = lambda r: torch.zeros(r, B, device="cuda")
torch.manual_seed(1126)
model = dl.DLADMMNetScalar(m, 1, d, B, data.A, Z0, z(m), z(m), K)
with torch.no_grad():
    for _ in range(n):
        out = model(data.X)
torch.cuda.synchronize()
print("ok", out[0][-1].abs().sum().item())
