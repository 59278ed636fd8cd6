"""Measured maxima behind the test tolerances: every golden fixture and the config-size cases, per tensor-core precision."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dladmm_b200 as dl, dladmm_oracle as orc
from _util import GOLDEN_NAMES, Golden, build_model, rel_l2
c = lambda t: t.double()
for prec in ("tf32_bf16x2", "tf32x3"):
    worst = {"Z": 0, "E": 0, "L": 0, "T": 0, "guardZ": 0, "guardE": 0}
    for name in GOLDEN_NAMES:
        g = Golden(name)
        sd = {k: c(v) for k, v in g.sd.items()}
        Zo, Eo, Lo, To = orc.forward(g.variant, sd, c(g.A), c(g.X), c(g.Z0), c(g.E0), c(g.L0), g.K)
        model = build_model(g, "cuda", prec)
        with torch.no_grad():
            out = model(g.X.cuda())
        xn = g.X.norm().item()
        w = {"Z": 0, "E": 0, "L": 0, "T": 0, "guardZ": 0, "guardE": 0}
        for k in range(g.K):
            z, e, l = out[0][k].cpu(), out[1][k].cpu(), out[2][k].cpu()
            w["Z"] = max(w["Z"], rel_l2(z, Zo[k], floor=1e-3)); w["E"] = max(w["E"], rel_l2(e, Eo[k], floor=1e-3 * xn)); w["L"] = max(w["L"], rel_l2(l, Lo[k], floor=1e-3 * xn))
            for key, a, b in (("guardZ", z, Zo[k]), ("guardE", e, Eo[k])):
                mism = (a != 0) != (b != 0)
                if mism.any():
                    w[key] = max(w[key], float(torch.maximum(a.double().abs(), b.abs())[mism].max()))
            if g.returns_T:
                w["T"] = max(w["T"], rel_l2(out[3][k + 1].cpu(), To[k + 1], floor=1e-2 * xn))
        print("%-12s %-18s " % (prec, name) + " ".join("%s %.1e" % kv for kv in w.items()))
        for k in w: worst[k] = max(worst[k], w[k])
    print("%-12s WORST over fixtures: " % prec + " ".join("%s %.1e" % kv for kv in worst.items()))
# config-size cases (tests/test_gpu_configs.py)
for prec in ("tf32_bf16x2", "tf32x3"):
    for variant, K, B in (("full", 20, 10240), ("lasso", 15, 10240), ("tied", 15, 10112), ("scalar", 15, 20480)):
        m, d = 250, 500
        torch.manual_seed(3)
        data = dl.gen_syn_data(B, m=m, d=d, seed=1129, dense_noise_sigma=(1.0 / m ** 0.5 if variant == "lasso" else None))
        Z0 = torch.rand(d, B, device="cuda") / d
        E0 = torch.zeros(m, B, device="cuda"); L0 = torch.zeros(m, B, device="cuda")
        model = dl.VARIANT_CLASSES[variant](m, 1, d, B, data.A, Z0, E0, L0, K, precision=prec)
        with torch.no_grad():
            outs = model(data.X)
        d64 = lambda t: t.detach().double().cpu()
        sd = {k: d64(v) for k, v in model.state_dict().items()}
        ref = orc.forward(variant, sd, d64(data.A), d64(data.X), d64(Z0), d64(E0), d64(L0), K)
        e3 = max(rel_l2(outs[i][k].cpu(), ref[i][k], floor=1e-2 * (B ** 0.5)) for i in range(3) for k in range(3))
        eK = max(rel_l2(outs[i][k].cpu(), ref[i][k], floor=1e-2 * (B ** 0.5)) for i in range(3) for k in range(K))
        print("%-12s config %-7s K=%d B=%d: max err k<3 %.1e, all k %.1e" % (prec, variant, K, B, e3, eK))
