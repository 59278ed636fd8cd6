"""Small forward+backward of every variant, a quick run of every variant and precision (compute-sanitizer is closed on this pool, so it is only a smoke run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
torch.manual_seed(0)
m, d, B, K = 40, 72, 36, 2
for prec in ("tf32x3", "fp32"):
    for variant, cls in dl.VARIANT_CLASSES.items():
        if variant.endswith("newS"):
            continue          # forward(x, K) signature, covered by tests/test_newS.py
        data = dl.gen_syn_data(B, m=m, d=d, seed=3)
        z = lambda r: torch.zeros(r, B, device="cuda")
        bs = B
        model = cls(m, 1, d, bs, data.A, torch.rand(d, B, device="cuda") / d, z(m), z(m), K, precision=prec)
        out = model(data.X)
        loss = sum(t.abs().sum() for t in out[0]) + sum(t.abs().sum() for t in out[1])
        loss.backward()
        if variant in ("scalar", "lasso"):
            model.zero_grad()
            l2, _ = model.l1l1_loss(data.X, 0.01)
            l2.backward()
        torch.cuda.synchronize()
        print(prec, variant, "ok", float(loss))
