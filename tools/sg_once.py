"""Two safeguarded evaluations (scalar K=20, 4096 columns): target of an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
B, K, m, d = 4096, 20, 250, 500
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
z = lambda r: torch.zeros(r, B, device="cuda")
torch.manual_seed(1126)
mdl = dl.DLADMMNetScalar(m, 1, d, B, data.A, torch.rand(d, B, device="cuda") / d, z(m), z(m), K)
for _ in range(2):
    out = mdl.forward_safeguarded(data.X, True, True, delta=0.0, mu_k_method="EMA", mu_k_param=0.5)
torch.cuda.synchronize()
print("ok", sum(out[4]))
