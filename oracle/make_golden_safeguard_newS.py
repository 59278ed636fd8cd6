"""Generate tests/golden/elz_*.npz from the UNMODIFIED evaluation class of test_syn_l1l1_newS_Acols.py (the safeguarded
forward of the E -> L -> Z ordering: KM_ZEL / KM_ELZ / Snorm_ELZ, :136-277).  Test infrastructure; build container only:

    python oracle/make_golden_safeguard_newS.py

The product path for this evaluation is not built yet; these fixtures pin the oracle restatement
(`dladmm_oracle.safeguarded_forward_newS`) on boxes without the reference tree and are the targets of the future GPU
parity tests.  Layout as tests/golden/sg_*.npz, except that Z/E/L hold K+1 entries (initial variables first) and there
is no T list."""
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
    ("elz_ema", 24, 40, 64, 6, 51, True, True, False, 0, 0.0, "EMA", 0.5),
    ("elz_gs_continued", 24, 40, 64, 4, 52, True, True, True, 9, 0.1, "GS", 0.9),
    ("elz_learned_only", 24, 40, 64, 6, 53, True, False, False, 0, -99.0, "None", 0.0),
    ("elz_km_only", 24, 40, 64, 6, 54, False, False, False, 12, -99.0, "None", 0.0),
]


def make(name, m, d, B, layers, seed, use_learned, use_safeguard, continued, num_iter, delta, method, param, alpha=0.01):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    A = torch.randn(m, d, generator=g)
    A = A / A.pow(2).sum(dim=0, keepdim=True).sqrt()
    Zs = (torch.rand(d, B, generator=g) < 0.2).float() * torch.randn(d, B, generator=g) * 2.0
    Es = (torch.rand(m, B, generator=g) < 0.2).float() * torch.randn(m, B, generator=g) * 2.0
    X = A.mm(Zs) + Es
    Z0, E0, L0 = torch.zeros(d, B), torch.zeros(m, B), torch.zeros(m, B)
    cls = lr.load_eval_class(layers, alpha=alpha, delta=delta, mu_k_method=method, mu_k_param=param, continued=continued,
                             num_iter=num_iter, use_learned=use_learned, use_safeguard=use_safeguard,
                             script="test_syn_l1l1_newS_Acols.py")
    K = layers if (not continued and (use_learned or use_safeguard)) else num_iter
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
            out = ref(X, use_learned, use_safeguard, continued, K)
    Z, E, L = out[:3]
    cnt = np.asarray(out[3], dtype=np.float64) if len(out) == 4 else np.zeros(layers)
    o = orc.safeguarded_forward_newS(sd, A, X, Z0, E0, L0, layers, use_learned, use_safeguard, continued, None, num_iter,
                                     delta, method, param, alpha, lip=lip)
    assert all(torch.equal(a, b) for a, b in zip(o[0] + o[1] + o[2], list(Z) + list(E) + list(L))), "oracle differs from the reference"
    tests = o[4]
    blob = dict(A=A.numpy(), X=X.numpy(), lip=np.array(lip), layers=np.array(layers), alpha=np.array(alpha),
                use_learned=np.array(use_learned), use_safeguard=np.array(use_safeguard), continued=np.array(continued),
                num_iter=np.array(num_iter), delta=np.array(delta), method=np.array(method), param=np.array(param),
                Z=torch.stack(list(Z)).numpy(), E=torch.stack(list(E)).numpy(), L=torch.stack(list(L)).numpy(),
                sg_count=cnt, keys=np.array(list(sd.keys())))
    if tests:
        blob["s_norm"] = torch.stack([t[0] for t in tests]).numpy()
        blob["thr"] = torch.stack([t[1] for t in tests]).numpy()
    for k, v in sd.items():
        blob["sd/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    print("wrote", name, "entries", len(Z), "sg_count", cnt.tolist())


if __name__ == "__main__":
    if not lr.reference_available():
        sys.exit("reference not found at %s" % lr.REFERENCE_ROOT)
    for case in CASES:
        make(*case)
