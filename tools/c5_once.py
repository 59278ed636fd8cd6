"""A few forwards at the large-scale shape (scalar, m 1000, d 2000, all iterates): profiling target for the CTA-pair product kernel
(umma_gemm_pair_kernel, tcgen05.mma.cta_group::2).  argv: columns (default 32768), layers (default 4), calls (default 3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m, d = 1000, 2000
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
Z0 = torch.rand(d, B, device="cuda") / d
z = lambda r: torch.zeros(r, B, device="cuda")
torch.manual_seed(1126)
model = dl.DLADMMNetScalar(m, 1, d, B, data.A, Z0, z(m), z(m), K)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for i in range(n):
        if i == n - 1:
            t0.record()
        out = model(data.X)
t1.record()
torch.cuda.synchronize()
print("ok %.3f ms per %d-layer forward, checksum %.6e" % (t0.elapsed_time(t1), K, out[0][-1].abs().sum().item()))
