"""C1 (scalar, m 250, d 500, K 15, 65 536 columns): 3 inference forwards, then 3 training steps (forward + fused loss + backward).
Profiling target for the ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m, d, K = 250, 500, 15
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
Z0 = torch.rand(d, B, device="cuda") / d
z = lambda r: torch.zeros(r, B, device="cuda")
torch.manual_seed(1126)
model = dl.DLADMMNetScalar(m, 1, d, B, data.A, Z0, z(m), z(m), K)
with torch.no_grad():
    for _ in range(3):
        out = model(data.X)
del out
for _ in range(3):
    model.zero_grad(set_to_none=True)
    loss, _ = model.l1l1_loss(data.X, 0.001, [0.5] * (K - 1) + [1.0])
    loss.backward()
torch.cuda.synchronize()
print("ok", loss.item())
