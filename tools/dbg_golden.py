import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import dladmm_b200 as dl, dladmm_oracle as orc
from _util import Golden, build_model, rel_l2
name = sys.argv[1] if len(sys.argv) > 1 else "ltheta_small"
g = Golden(name)
print(name, g.variant, "m", g.m, "d", g.d, "K", g.K, "bs", g.bs, "X", tuple(g.X.shape), "|X|", g.X.norm().item())
for k, v in g.sd.items():
    if not k.startswith("fc"): print("  ", k, tuple(v.shape), v.flatten()[:4].tolist())
c = lambda t: t.double()
sd = {k: c(v) for k, v in g.sd.items()}
Zo, Eo, Lo, To = orc.forward(g.variant, sd, c(g.A), c(g.X), c(g.Z0), c(g.E0), c(g.L0), g.K)
for prec in ("fp32", "tf32x3", "tf32_bf16x2"):
    for sched in (0, 1):
        if sched: os.environ["DLADMM_NO_PERSISTENT"] = "1"
        else: os.environ.pop("DLADMM_NO_PERSISTENT", None)
        model = build_model(g, "cuda", prec)
        with torch.no_grad():
            out = model(g.X.cuda())
        for k in range(g.K):
            print(prec, "per-layer" if sched else "default  ", "k", k, "Z %.2e E %.2e L %.2e" % (rel_l2(out[0][k].cpu(), Zo[k]), rel_l2(out[1][k].cpu(), Eo[k]), rel_l2(out[2][k].cpu(), Lo[k])),
                  "|Z| %.3g |E| %.3g |L| %.3g" % (Zo[k].norm(), Eo[k].norm(), Lo[k].norm()))
