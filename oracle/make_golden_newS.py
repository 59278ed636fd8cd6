"""Generate tests/golden/newS_*.npz by running the UNMODIFIED reference classes with the E -> L -> Z ordering and prefix
execution forward(x, K): main_syn_scalar_newS_layerwise.py, main_syn_scalar_tied_newS_layerwise.py,
main_syn_scalar_ptied_newS_layerwise.py.  Test infrastructure; build container only (needs /root/reference):

    python oracle/make_golden_newS.py

Each fixture: inputs, the (perturbed) reference state_dict in its registration order, the reference outputs Z, E, L of
forward(x, K_run), fixed cotangents on every returned iterate and the reference's autograd gradient of every parameter
(`nograd` lists the parameters the reference leaves without a gradient: layers beyond the prefix, and the E/L-step
parameters of the last executed layer)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import load_reference as lr  # noqa: E402
from make_golden import syn, OUT  # noqa: E402

CASES = [
    # name, variant, m, d, bs, layers, K_run, interval, seed
    ("newS_small", "newS", 24, 40, 20, 5, 5, None, 31),
    ("newS_prefix", "newS", 24, 40, 20, 6, 4, None, 32),          # forward(x, 4) of a 6-layer model
    ("newS_overrun", "newS", 17, 33, 8, 3, 7, None, 33),          # K > layers is clipped (:88)
    ("tied_newS_small", "tied_newS", 24, 40, 20, 4, 4, None, 34),
    ("ptied_newS_small", "ptied_newS", 24, 40, 20, 6, 6, 3, 35),
    ("ptied_newS_prefix", "ptied_newS", 24, 40, 12, 6, 5, 2, 36),
]


def make(name, variant, m, d, bs, layers, K_run, interval, seed):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    A, X = syn(m, d, bs, g)
    Z0 = torch.rand(d, bs, generator=g) / d
    E0 = 0.05 * torch.randn(m, bs, generator=g)
    L0 = 0.05 * torch.randn(m, bs, generator=g)
    extra = {"interval": interval} if interval else {}
    ref = lr.build(variant, m, 10000, d, bs, A, Z0, E0, L0, layers, **extra)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    for k in sd:
        if not k.startswith("fc"):
            sd[k] = sd[k] * (1 + 0.2 * torch.randn(sd[k].shape, generator=g))
        if k.startswith("active_para"):
            sd[k] = sd[k].abs() * 3 + 0.02          # the 0.01 defaults barely threshold anything at this scale
    ref.load_state_dict(sd)
    with lr.cuda_is_identity():
        Z, E, L = ref(X, K_run)
    n = len(Z)
    assert n == len(E) == len(L) == min(K_run, layers)
    cz = [torch.randn(d, bs, generator=g) for _ in range(n)]
    ce = [torch.randn(m, bs, generator=g) for _ in range(n)]
    cl = [torch.randn(m, bs, generator=g) for _ in range(n)]
    loss = sum((Z[k] * cz[k]).sum() + (E[k] * ce[k]).sum() + (L[k] * cl[k]).sum() for k in range(n))
    loss.backward()
    blob = dict(variant=np.array(variant), layers=np.array(layers), K_run=np.array(K_run), bs=np.array(bs),
                interval=np.array(interval or 0), name=np.array(ref.name()),
                A=A.numpy(), X=X.numpy(), Z0=Z0.numpy(), E0=E0.numpy(), L0=L0.numpy(),
                Z=torch.stack(Z).detach().numpy(), E=torch.stack(E).detach().numpy(), L=torch.stack(L).detach().numpy(),
                cz=torch.stack(cz).numpy(), ce=torch.stack(ce).numpy(), cl=torch.stack(cl).numpy(),
                loss=loss.detach().numpy(), keys=np.array(list(sd.keys())))
    nograd = []
    for k, v in sd.items():
        blob["sd/" + k] = v.numpy()
    for nme, p in ref.named_parameters():
        if p.grad is None:
            nograd.append(nme)
            blob["grad/" + nme] = np.zeros(tuple(p.shape), dtype=np.float32)
        else:
            blob["grad/" + nme] = p.grad.numpy()
    blob["nograd"] = np.array(nograd if nograd else [""])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    print("wrote", name, ref.name(), "n =", n, "loss", float(loss), "no grad:", len(nograd))


if __name__ == "__main__":
    if not lr.reference_available():
        sys.exit("reference not found at %s" % lr.REFERENCE_ROOT)
    for case in CASES:
        make(*case)
