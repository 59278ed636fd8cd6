"""Forward / training-step latency sweep over batch sizes (development tool).
   python tools/sweep.py [variant]          env DLADMM_FUSED=0/1 selects the all-layer fused forward."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dladmm_b200 as dl

variant = sys.argv[1] if len(sys.argv) > 1 else "scalar"
m, d, K = 250, 500, 15
out = {}
for prec in ("tf32x3", "tf32", "fp32"):
    for B in (32, 256, 2048, 16384, 65536):
        if prec == "fp32" and B > 16384:
            continue
        data = dl.gen_syn_data(B, m=m, d=d, seed=1)
        Z0 = torch.rand(d, B, device="cuda") / d
        z = lambda r: torch.zeros(r, B, device="cuda")
        model = dl.VARIANT_CLASSES[variant](m, 1, d, B, data.A, Z0, z(m), z(m), K, precision=prec)
        def fwd():
            with torch.no_grad():
                model(data.X)
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20 if B <= 2048 else 5
        e0.record()
        for _ in range(n):
            fwd()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out["%s/B%d" % (prec, B)] = round(ms, 4)
        print("%-7s B=%6d  fwd %.3f ms  %.2f M inst/s" % (prec, B, ms, B / ms / 1e3), flush=True)
print(json.dumps(out))
