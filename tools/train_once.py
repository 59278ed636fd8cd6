"""One forward+backward at C1 size (profiling target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m, d, K = 250, 500, 3
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
Z0 = torch.rand(d, B, device="cuda") / d
z = lambda r: torch.zeros(r, B, device="cuda")
model = dl.DLADMMNetScalar(m, 1, d, B, data.A, Z0, z(m), z(m), K)
for _ in range(2):
    model.zero_grad(set_to_none=True)
    loss, _ = model.l1l1_loss(data.X, 0.001, [0.5] * (K - 1) + [1.0])
    loss.backward()
torch.cuda.synchronize()
print("ok", loss.item())
