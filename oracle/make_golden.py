"""Generate tests/golden/*.npz by running the UNMODIFIED reference DLADMMNet classes.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Each fixture holds: the inputs (A, X, Z0, E0, L0), the reference state_dict (default init with
perturbed step/threshold parameters so every parameter matters), the reference's forward outputs
for every layer, fixed cotangents on every returned iterate, and the reference's autograd
gradients for every parameter.  The reference has no golden vectors of its own and its data is
unseeded, so these are the pins for both the oracle and the CUDA path.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import load_reference as lr  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, variant, m, d, bs, K, seed, negative_threshold_layer
    ("lena_small", "lena", 24, 40, 20, 4, 11, None),
    ("ltheta_small", "ltheta", 24, 40, 20, 4, 12, None),
    ("scalar_small", "scalar", 24, 40, 20, 4, 13, None),
    ("full_small", "full", 24, 40, 20, 4, 14, None),
    ("tied_small", "tied", 24, 40, 20, 4, 15, None),
    ("lasso_small", "lasso", 24, 40, 20, 4, 16, None),
    ("scalar_negtheta", "scalar", 24, 40, 12, 3, 17, 1),
    ("full_negtheta", "full", 17, 33, 8, 3, 18, 2),
    ("scalar_c1shape", "scalar", 250, 500, 24, 2, 19, None),   # BASELINE config-1 shape, few columns
    ("lena_c4shape", "lena", 256, 512, 20, 2, 20, None),       # BASELINE config-4 shape
]


def syn(m, d, B, g, p=0.1, sigma=1.0):
    """gen_syn_data.py:12-47 semantics on a torch generator."""
    A = torch.randn(m, d, generator=g)
    A = A / A.pow(2).sum(dim=0, keepdim=True).sqrt()
    Z = (torch.rand(d, B, generator=g) < p).float() * torch.randn(d, B, generator=g) * sigma
    E = (torch.rand(m, B, generator=g) < p).float() * torch.randn(m, B, generator=g) * sigma
    return A, A.mm(Z) + E


def make(name, variant, m, d, bs, K, seed, neg_layer):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)                     # the reference draws W noise from the global RNG
    A, X = syn(m, d, bs, g)
    Z0 = torch.rand(d, bs, generator=g) / d     # main_syn_l1l1_scalar.py:226
    E0 = 0.05 * torch.randn(m, bs, generator=g)  # scripts use zeros; non-zero exercises every term
    L0 = 0.05 * torch.randn(m, bs, generator=g)
    ref = lr.build(variant, m, 10000, d, bs, A, Z0, E0, L0, K)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    for k in sd:
        if not k.startswith("fc"):
            sd[k] = sd[k] * (1 + 0.2 * torch.randn(sd[k].shape, generator=g))
    if neg_layer is not None:
        for nm in ("active_para", "active_para1"):
            key = "%s.%d" % (nm, neg_layer)
            t = sd[key].clone()
            t.view(-1)[:: 2] = -0.05             # negative thresholds: act() is no longer a shrinkage
            sd[key] = t
    ref.load_state_dict(sd)
    with lr.cuda_is_identity():
        out = ref(X)
    Z, E, L = out[0], out[1], out[2]
    T = out[3] if len(out) == 4 else None
    cz = [torch.randn(d, bs, generator=g) for _ in range(K)]
    ce = [torch.randn(m, bs, generator=g) for _ in range(K)]
    cl = [torch.randn(m, bs, generator=g) for _ in range(K)]
    ct = [torch.randn(m, bs, generator=g) for _ in range(K + 1)]
    loss = sum((Z[k] * cz[k]).sum() + (E[k] * ce[k]).sum() + (L[k] * cl[k]).sum() for k in range(K))
    if T is not None:
        loss = loss + sum((T[k] * ct[k]).sum() for k in range(K + 1))
    else:
        ct = [torch.zeros_like(c) for c in ct]
    loss.backward()
    blob = dict(variant=np.array(variant), K=np.array(K), bs=np.array(bs), returns_T=np.array(T is not None),
                A=A.numpy(), X=X.numpy(), Z0=Z0.numpy(), E0=E0.numpy(), L0=L0.numpy(),
                Z=torch.stack(Z).detach().numpy(), E=torch.stack(E).detach().numpy(),
                L=torch.stack(L).detach().numpy(),
                cz=torch.stack(cz).numpy(), ce=torch.stack(ce).numpy(), cl=torch.stack(cl).numpy(),
                ct=torch.stack(ct).numpy(), loss=loss.detach().numpy(),
                keys=np.array(list(sd.keys())))
    if T is not None:
        blob["T"] = torch.stack(T).detach().numpy()
    for k, v in sd.items():
        blob["sd/" + k] = v.numpy()
    for n, p in ref.named_parameters():
        blob["grad/" + n] = p.grad.numpy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    print("wrote", name, "loss", float(loss))


if __name__ == "__main__":
    if not lr.reference_available():
        sys.exit("reference not found at %s" % lr.REFERENCE_ROOT)
    for case in CASES:
        make(*case)
