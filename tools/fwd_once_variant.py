"""A few forwards of one variant (profiling target): python tools/fwd_once_variant.py full 65536 3"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
variant = sys.argv[1] if len(sys.argv) > 1 else "full"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
K = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m, d = 250, 500
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
Z0 = torch.rand(d, B, device="cuda") / d
z = lambda r: torch.zeros(r, B, device="cuda")
torch.manual_seed(1126)
model = dl.VARIANT_CLASSES[variant](m, 1, d, B, data.A, Z0, z(m), z(m), K)
with torch.no_grad():
    for _ in range(4):
        out = model(data.X)
torch.cuda.synchronize()
print("ok", out[0][-1].abs().sum().item())
