"""Generate tests/golden/sg_*.npz from the UNMODIFIED evaluation class of the reference (test_syn_l1l1_scalar.py).

Build container only (needs /root/reference):   python oracle/make_golden_safeguard.py

Each fixture holds the inputs, a state_dict for the script's scalar class (a deliberately imperfect "learned" model
so that the safeguard fires on part of the columns), the flags, and the reference's outputs: the Z/E/L/T lists and
sg_count.  The per-layer (||S||, threshold) pairs come from the oracle restatement run on the same inputs (the
reference does not return them); they let the parity tests skip columns whose decision is within rounding.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dladmm_oracle as orc  # noqa: E402
import load_reference as lr  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, m, d, B, layers, seed, use_learned, use_safeguard, continued, num_iter, delta, method, param
    ("sg_ema", 24, 40, 64, 6, 41, True, True, False, 0, 0.3, "EMA", 0.5),
    ("sg_gs", 24, 40, 64, 6, 42, True, True, False, 0, 0.0, "GS", 0.3),
    ("sg_rt", 24, 40, 64, 6, 43, True, True, False, 0, 0.1, "RT", 0.0),
    ("sg_none", 24, 40, 64, 6, 44, True, True, False, 0, -99.0, "None", 0.0),
    ("sg_learned_only", 24, 40, 64, 6, 45, True, False, False, 0, -99.0, "None", 0.0),
    ("sg_km_only", 24, 40, 64, 6, 46, False, False, False, 12, -99.0, "None", 0.0),
    ("sg_continued", 24, 40, 64, 4, 47, True, True, True, 9, 0.1, "RT", 0.0),
    ("sg_c1shape", 250, 500, 32, 3, 48, True, True, False, 0, 0.1, "RT", 0.0),
]


def make(name, m, d, B, layers, seed, use_learned, use_safeguard, continued, num_iter, delta, method, param, alpha=0.01):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    A = torch.randn(m, d, generator=g)
    A = A / A.pow(2).sum(dim=0, keepdim=True).sqrt()
    Zs = (torch.rand(d, B, generator=g) < 0.2).float() * torch.randn(d, B, generator=g) * 2.0
    Es = (torch.rand(m, B, generator=g) < 0.2).float() * torch.randn(m, B, generator=g) * 2.0
    X = A.mm(Zs) + Es
    Z0, E0, L0 = torch.zeros(d, B), torch.zeros(m, B), torch.zeros(m, B)      # test_syn_l1l1_scalar.py:440-443
    cls = lr.load_eval_class(layers, alpha=alpha, delta=delta, mu_k_method=method, mu_k_param=param, continued=continued,
                             num_iter=num_iter, use_learned=use_learned, use_safeguard=use_safeguard)
    with lr.cuda_is_identity():
        ref = cls(m=m, n=B, d=d, batch_size=B, A=A, Z0=Z0, E0=E0, L0=L0, layers=layers)
        sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        lip = float(ref.L)
        for k in range(layers):
            sd["fc.%d.weight" % k] = (0.9 / lip) * (A.t() + 0.05 * torch.randn(d, m, generator=g))
            sd["active_para.%d" % k] = torch.full((1, 1), 0.02 + 0.01 * k)
            sd["active_para1.%d" % k] = torch.full((1, 1), 0.25 + 0.1 * k)
            sd["ss2.%d" % k] = torch.full((1, 1), 0.3 + 0.12 * k)
            sd["beta1.%d" % k] = torch.full((1, 1), 1.0 + 0.1 * k)
        ref.load_state_dict(sd)
        with torch.no_grad():
            out = ref(X, use_learned, use_safeguard, continued)
    Z, E, L, T = out[:4]
    cnt = np.asarray(out[4], dtype=np.float64) if len(out) == 5 else np.zeros(layers)
    o = orc.safeguarded_forward(sd, A, X, Z0, E0, L0, layers, use_learned, use_safeguard, continued=continued,
                                num_iter=num_iter, delta=delta, mu_method=method, mu_param=param, alpha=alpha, lip=lip)
    assert all(torch.equal(a, b) for a, b in zip(o[0], Z)), "oracle restatement differs from the reference"
    tests = o[5]
    blob = dict(A=A.numpy(), X=X.numpy(), lip=np.array(lip), layers=np.array(layers), alpha=np.array(alpha),
                use_learned=np.array(use_learned), use_safeguard=np.array(use_safeguard), continued=np.array(continued),
                num_iter=np.array(num_iter), delta=np.array(delta), method=np.array(method), param=np.array(param),
                Z=torch.stack(Z).numpy(), E=torch.stack(E).numpy(), L=torch.stack(L).numpy(), T=torch.stack(T).numpy(),
                sg_count=cnt, keys=np.array(list(sd.keys())))
    if tests:
        blob["s_norm"] = torch.stack([t[0] for t in tests]).numpy()
        blob["thr"] = torch.stack([t[1] for t in tests]).numpy()
    for k, v in sd.items():
        blob["sd/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    margin = min(((s - t).abs() / t.abs().clamp_min(1e-9)).min().item() for s, t in tests) if tests else float("nan")
    print("wrote", name, "iterations", len(Z), "sg_count", cnt.tolist(), "min margin %.2e" % margin)


if __name__ == "__main__":
    if not lr.reference_available():
        sys.exit("reference not found at %s" % lr.REFERENCE_ROOT)
    for case in CASES:
        make(*case)
