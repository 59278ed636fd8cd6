"""classical KM step, safeguard operator S, per-column safeguarded evaluation forward.

CPU part: the oracle restatement against the committed reference outputs (tests/golden/sg_*.npz, written by
oracle/make_golden_safeguard.py from the unmodified test_syn_l1l1_scalar.py class) and, when /root/reference is
present, live against that class; the product's mu_k updaters against the reference's mu_updater.py.
GPU part: DLADMMNetScalar.KM / S / forward_safeguarded (CUDA through the C ABI) against the same fixtures.
"""
import numpy as np
import pytest
import torch

import dladmm_oracle as orc
import load_reference as lr
from _util import SG_GOLDEN_NAMES, SgGolden, rel_l2

# A column whose ||S|| is within this relative distance of its threshold at any layer may legitimately take the
# other branch under fp32 rounding differences; such columns are excluded (and must stay a small minority).
MARGIN = 2e-4
TOL = {"fp32": 2e-5, "tf32x3": 1e-4}


def _oracle(g):
    return orc.safeguarded_forward(g.sd, g.A, g.X, g.Z0, g.E0, g.L0, g.layers, g.use_learned, g.use_safeguard,
                                   continued=g.continued, num_iter=g.num_iter, delta=g.delta, mu_method=g.method,
                                   mu_param=g.param, alpha=g.alpha, lip=g.lip)


@pytest.mark.parametrize("name", SG_GOLDEN_NAMES)
def test_oracle_matches_reference_evaluation_outputs(name):
    g = SgGolden(name)
    Z, E, L, T, cnt, _ = _oracle(g)
    cols = g.robust_columns(MARGIN)
    assert cols.float().mean() > 0.9
    assert len(Z) == g.Z.shape[0] and len(T) == g.T.shape[0]
    for got, want in ((Z, g.Z), (E, g.E), (L, g.L), (T, g.T)):
        assert rel_l2(torch.stack(got)[:, :, cols], want[:, :, cols]) < 1e-5
    if cols.all():
        assert cnt == g.sg_count


def test_fixture_set_exercises_both_branches():
    counts = {n: SgGolden(n).sg_count for n in SG_GOLDEN_NAMES}
    assert any(0 < c < SgGolden(n).B for n, cs in counts.items() for c in cs)
    assert counts["sg_none"] == [0.0] * 6          # BlankUpdater: mu_k = 1e10, the learned step is always kept


@pytest.mark.skipif(not lr.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("method,param,delta", [("EMA", 0.5, 0.3), ("GS", 0.3, 0.0), ("RT", 0.0, 0.1), ("None", 0.0, -99.0)])
def test_oracle_bit_exact_with_live_reference_class(method, param, delta):
    g = SgGolden("sg_rt")
    cls = lr.load_eval_class(g.layers, alpha=g.alpha, delta=delta, mu_k_method=method, mu_k_param=param)
    with lr.cuda_is_identity():
        ref = cls(m=g.m, n=g.B, d=g.d, batch_size=g.B, A=g.A, Z0=g.Z0, E0=g.E0, L0=g.L0, layers=g.layers)
        ref.load_state_dict(g.sd)
        with torch.no_grad():
            Zr, Er, Lr, Tr, cnt = ref(g.X, True, True, False)
    Z, E, L, T, c2, _ = orc.safeguarded_forward(g.sd, g.A, g.X, g.Z0, g.E0, g.L0, g.layers, True, True, delta=delta,
                                                mu_method=method, mu_param=param, alpha=g.alpha, lip=float(ref.L))
    for a, b in zip(Z + E + L + T, list(Zr) + list(Er) + list(Lr) + list(Tr)):
        assert torch.equal(a, b)
    assert list(cnt) == c2


@pytest.mark.skipif(not lr.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("method,param", [("EMA", 0.4), ("GS", 0.2), ("RT", 0.0), ("None", 0.0)])
def test_mu_updaters_match_reference(method, param):
    import dladmm_b200 as dl
    ref_dict = lr.load_mu_updater_dict()
    gen = torch.Generator().manual_seed(3)
    mu0 = torch.rand(50, generator=gen) + 0.5
    ours, ref = dl.mu_updater_dict[method](mu0.clone(), param), ref_dict[method](mu0.clone(), param)
    for _ in range(6):
        s = torch.rand(50, generator=gen) + 0.3
        b = (torch.rand(50, generator=gen) < 0.5).float()
        a, r = ours.step(s, b), ref.step(s, b)
        r = r if torch.is_tensor(r) else torch.full_like(s, r)
        assert torch.equal(a, r)


def test_recent_max_updater_keeps_a_window():
    import dladmm_b200 as dl
    up = dl.mu_updater_dict["RM"](torch.tensor([1.0, 5.0]), 2)
    assert up.step(torch.tensor([3.0, 1.0])).tolist() == [3.0, 5.0]
    assert up.step(torch.tensor([2.0, 2.0])).tolist() == [3.0, 2.0]


# ---- GPU --------------------------------------------------------------------------------------------------------------
def _model(g, precision):
    import dladmm_b200 as dl
    dev = torch.device("cuda:0")
    model = dl.DLADMMNetScalar(m=g.m, n=g.B, d=g.d, batch_size=g.B, A=g.A, Z0=g.Z0, E0=g.E0, L0=g.L0, layers=g.layers,
                               precision=precision, device=dev)
    model.load_state_dict(g.sd)       # same keys as the evaluation script's class
    return model, dev


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name", SG_GOLDEN_NAMES)
def test_safeguarded_forward_matches_reference(name, precision):
    g = SgGolden(name)
    model, dev = _model(g, precision)
    out = model.forward_safeguarded(g.X.to(dev), g.use_learned, g.use_safeguard, **g.kwargs())
    Z, E, L, T = out[:4]
    cols = g.robust_columns(50 * TOL[precision]).to(dev)
    assert cols.float().mean() > 0.8
    assert len(Z) == g.Z.shape[0] and len(T) == g.T.shape[0]
    for got, want in ((Z, g.Z), (E, g.E), (L, g.L), (T, g.T)):
        assert rel_l2(torch.stack(got)[:, :, cols].cpu(), want[:, :, cols.cpu()]) < TOL[precision]
    if g.use_learned and g.use_safeguard:
        assert len(out) == 5
        flips = sum(abs(a - b) for a, b in zip(out[4], g.sg_count))
        assert flips <= (~cols).sum().item() * g.layers


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_km_and_s_operator_match_oracle(precision):
    g = SgGolden("sg_rt")
    model, dev = _model(g, precision)
    gen = torch.Generator().manual_seed(9)
    Zk, Ek, Lk, Ep = (torch.randn(s, g.B, generator=gen) for s in (g.d, g.m, g.m, g.m))
    Tk = g.A.mm(Zk) + Ek - g.X
    lip = torch.tensor(g.lip)
    Zn, En, Ln, Tn = orc.km_step(g.A, g.X, Zk, Ek, Lk, Tk, g.alpha, 1.0, 0.999 / lip, 0.3)
    c = (0.3 / 0.7) ** 0.5
    S_want = torch.cat([Tn, c * (En - 2 * Ek + Ep)])
    a = [t.to(dev) for t in (Zk, Ek, Lk, Tk, g.X)]
    Var, Zg, Eg, Tg, Lg = model.KM(*a, alpha=g.alpha)
    for got, want in ((Var, Lk + Tk), (Zg, Zn), (Eg, En), (Tg, Tn), (Lg, Ln)):
        assert rel_l2(got.cpu(), want) < TOL[precision]
    S_got = model.S(*a, Ep.to(dev), alpha=g.alpha)
    assert S_got.shape == (2 * g.m, g.B)
    assert rel_l2(S_got.cpu(), S_want) < TOL[precision]
    n_got = model._s_norm(*a, Ep.to(dev), g.alpha)
    assert rel_l2(n_got.cpu(), model.two_norm(S_want)) < TOL[precision]


@pytest.mark.gpu
def test_select_kernel_is_an_exact_per_column_choice():
    import ctypes as C
    from dladmm_b200 import _lib
    dev = torch.device("cuda:0")
    B = 1000
    gen = torch.Generator().manual_seed(2)
    a, b = torch.randn(7, B, generator=gen).to(dev), torch.randn(7, B, generator=gen).to(dev)
    s, mu = torch.rand(B, generator=gen).to(dev), torch.rand(B, generator=gen).to(dev)
    out, keep = torch.empty_like(a), torch.empty(B, device=dev)
    pairs = (_lib.SgPair * 1)()
    pairs[0].a, pairs[0].b, pairs[0].out, pairs[0].rows = a.data_ptr(), b.data_ptr(), out.data_ptr(), 7
    _lib.check(_lib.load().dladmm_sg_select(1, pairs, B, s.data_ptr(), mu.data_ptr(), 0.9, keep.data_ptr(),
                                            torch.cuda.current_stream(dev).cuda_stream))
    want_keep = (s < 0.9 * mu).float()
    assert torch.equal(keep, want_keep)
    assert torch.equal(out, torch.where(want_keep.bool().unsqueeze(0), a, b))


@pytest.mark.gpu
def test_safeguard_needs_family_b():
    import dladmm_b200 as dl
    dev = torch.device("cuda:0")
    A = torch.randn(8, 12)
    net = dl.DLADMMNetLasso(m=8, n=4, d=12, batch_size=4, A=A, Z0=torch.zeros(12, 4), E0=torch.zeros(8, 4),
                            L0=torch.zeros(8, 4), layers=2, device=dev)
    with pytest.raises(NotImplementedError):
        net.forward_safeguarded(torch.zeros(8, 4, device=dev), True, True)


@pytest.mark.skipif(not lr.reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("ul,us,cont,method,param,delta", [(True, True, False, "EMA", 0.5, 0.0), (True, True, True, "GS", 0.9, 0.1),
                                                           (True, False, False, "None", 0.0, -99.0),
                                                           (False, False, False, "None", 0.0, -99.0)])
def test_newS_safeguard_oracle_bit_exact_with_live_reference_class(ul, us, cont, method, param, delta):
    """Groundwork for the SURVEY 8(f)-2 remainder (product path not built yet): the oracle's restatement of the
    E -> L -> Z safeguarded evaluation (KM_ZEL / KM_ELZ / Snorm_ELZ, test_syn_l1l1_newS_Acols.py:136-277) against the
    unmodified reference class, learned / classical / safeguarded / continued modes, with fallbacks actually happening."""
    m, d, B, layers, num_iter = 24, 40, 32, 5, 12
    g = torch.Generator().manual_seed(7)
    A = torch.randn(m, d, generator=g); A = A / A.pow(2).sum(0, keepdim=True).sqrt()
    Zs = (torch.rand(d, B, generator=g) < 0.1).float() * torch.randn(d, B, generator=g)
    Es = (torch.rand(m, B, generator=g) < 0.1).float() * torch.randn(m, B, generator=g)
    X = A.mm(Zs) + Es
    z = lambda r: torch.zeros(r, B)
    cls = lr.load_eval_class(layers, alpha=0.01, delta=delta, mu_k_method=method, mu_k_param=param, continued=cont,
                             num_iter=num_iter, use_learned=ul, use_safeguard=us, script="test_syn_l1l1_newS_Acols.py")
    torch.manual_seed(3)
    with lr.cuda_is_identity():
        ref = cls(m=m, n=1, d=d, batch_size=B, A=A, Z0=z(d), E0=z(m), L0=z(m), layers=layers)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    for k in sd:
        sd[k] = sd[k] * 0.3 if k.startswith("fc") else sd[k] * (1 + 0.2 * torch.randn(sd[k].shape, generator=g))
    ref.load_state_dict(sd)
    K = layers if (not cont and (ul or us)) else num_iter
    with torch.no_grad(), lr.cuda_is_identity():
        out = ref(X, ul, us, cont, K)
    Zo, Eo, Lo, cnt, _ = orc.safeguarded_forward_newS(sd, A, X, z(d), z(m), z(m), layers, ul, us, cont, None, num_iter, delta,
                                                      method, param, 0.01, lip=float(ref.L))
    assert len(out[0]) == len(Zo) == K + 1
    for a, b in zip(list(out[0]) + list(out[1]) + list(out[2]), Zo + Eo + Lo):
        assert torch.equal(a, b)
    if ul and us:
        assert [float(c) for c in out[3]] == cnt and 0 < sum(cnt) < layers * B


def test_newS_safeguard_oracle_matches_committed_reference_outputs():
    """tests/golden/elz_*.npz (oracle/make_golden_safeguard_newS.py, unmodified test_syn_l1l1_newS_Acols.py class): the
    oracle restatement reproduces every returned iterate and the fallback counts on any box."""
    import os
    from _util import ELZ_GOLDEN_NAMES, GOLDEN_DIR
    assert len(ELZ_GOLDEN_NAMES) >= 4
    fired = 0
    for name in ELZ_GOLDEN_NAMES:
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        t = lambda k: torch.from_numpy(z[k].copy())
        sd = {str(k): t("sd/" + str(k)) for k in z["keys"]}
        A, X = t("A"), t("X")
        m, d = A.shape
        B = X.shape[1]
        zz = lambda r: torch.zeros(r, B)
        Z, E, L, cnt, _ = orc.safeguarded_forward_newS(sd, A, X, zz(d), zz(m), zz(m), int(z["layers"]), bool(z["use_learned"]),
                                                       bool(z["use_safeguard"]), bool(z["continued"]), None, int(z["num_iter"]),
                                                       float(z["delta"]), str(z["method"]), float(z["param"]), float(z["alpha"]),
                                                       lip=float(z["lip"]))
        assert len(Z) == z["Z"].shape[0]
        for k in range(len(Z)):
            assert rel_l2(Z[k], t("Z")[k], floor=1e-6) < 2e-6 and rel_l2(E[k], t("E")[k], floor=1e-6) < 2e-6, (name, k)
            assert rel_l2(L[k], t("L")[k], floor=1e-6) < 2e-6, (name, k)
        if bool(z["use_learned"]) and bool(z["use_safeguard"]):
            assert cnt == z["sg_count"].tolist()
            fired += int(0 < sum(cnt) < int(z["layers"]) * B)
    assert fired >= 1
