"""SURVEY.md section 8(f)-2 remainder, GPU: the safeguarded evaluation of the E -> L -> Z ordering
(KM_ZEL / KM_ELZ / Snorm_ELZ / forward, test_syn_l1l1_newS_Acols.py:136-277) through the library's half-layer entry
points (dladmm_problem.start_half / stop_half), against the reference-generated fixtures tests/golden/elz_*.npz and the
oracle restatement (which the CPU suite pins bit-exactly to the live reference class)."""
import ctypes as C

import pytest
import torch

import dladmm_oracle as orc
from _util import ELZ_GOLDEN_NAMES, ElzGolden, SgGolden, rel_l2

pytestmark = pytest.mark.gpu
TOL = {"fp32": 2e-5, "tf32x3": 1e-4}


def _model(g, precision, cls="DLADMMNetNewS"):
    import dladmm_b200 as dl
    dev = torch.device("cuda:0")
    model = getattr(dl, cls)(m=g.m, n=g.B, d=g.d, batch_size=g.B, A=g.A, Z0=g.Z0, E0=g.E0, L0=g.L0, layers=g.layers,
                             precision=precision, device=dev)
    model.load_state_dict(g.sd)
    return model, dev


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name", ELZ_GOLDEN_NAMES)
def test_newS_safeguarded_forward_matches_reference(name, precision):
    g = ElzGolden(name)
    model, dev = _model(g, precision)
    out = model.forward_safeguarded(g.X.to(dev), g.use_learned, g.use_safeguard, **g.kwargs())
    Z, E, L = out[:3]
    cols = g.robust_columns(50 * TOL[precision]).to(dev)
    assert cols.float().mean() > 0.8
    assert len(Z) == g.Z.shape[0] == len(E) == len(L)
    for name_, got, want in (("Z", Z, g.Z), ("E", E, g.E), ("L", L, g.L)):
        assert rel_l2(torch.stack(got)[:, :, cols].cpu(), want[:, :, cols.cpu()]) < TOL[precision], name_
    if g.use_learned and g.use_safeguard:
        assert len(out) == 4
        flips = sum(abs(a - b) for a, b in zip(out[3], g.sg_count))
        assert flips <= (~cols).sum().item() * g.layers
        assert 0 < sum(out[3]) < g.layers * g.B          # both branches of the selection ran
    else:
        assert len(out) == 3


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_km_zel_km_elz_snorm_match_oracle(precision):
    g = ElzGolden("elz_ema")
    model, dev = _model(g, precision)
    gen = torch.Generator().manual_seed(11)
    Zk, Ek, Lk = (torch.randn(s, g.B, generator=gen) for s in (g.d, g.m, g.m))
    Tk = g.A.mm(Zk) + Ek - g.X
    lip = torch.tensor(g.lip)
    want = orc.km_zel(g.A, g.X, Zk, Ek, Lk, Tk, g.alpha, lip)
    got = model.KM_ZEL(*(t.to(dev) for t in (Zk, Ek, Lk, Tk, g.X)), alpha=g.alpha)
    for a, b in zip(got, want):
        assert rel_l2(a.cpu(), b) < TOL[precision]
    want = orc.km_elz(g.A, g.X, Ek, Lk, Zk, g.alpha, lip)
    got = model.KM_ELZ(*(t.to(dev) for t in (Ek, Lk, Zk, g.X)), alpha=g.alpha)
    for a, b in zip(got, want):
        assert rel_l2(a.cpu(), b) < TOL[precision]
    want = orc.snorm_elz(g.A, g.X, Ek, Lk, Zk, g.alpha, lip)
    got = model.Snorm_ELZ(*(t.to(dev) for t in (Ek, Lk, Zk, g.X)), alpha=g.alpha)
    assert rel_l2(got.cpu(), want) < 5 * TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("B", [26, 64])
def test_safeguarded_forwards_at_ragged_batch_against_oracle(B, precision):
    """Both orderings at a batch that is not a multiple of 4 (the scripts' own -bs 25 regime): the product runs at the
    padded pitch and returns narrowed views; compared with the oracle on the same columns."""
    import dladmm_b200 as dl
    dev = torch.device("cuda:0")
    m, d, layers = 24, 40, 5
    gen = torch.Generator().manual_seed(77)
    A = torch.randn(m, d, generator=gen); A = A / A.pow(2).sum(0, keepdim=True).sqrt()
    Zs = (torch.rand(d, B, generator=gen) < 0.2).float() * torch.randn(d, B, generator=gen) * 2.0
    Es = (torch.rand(m, B, generator=gen) < 0.2).float() * torch.randn(m, B, generator=gen) * 2.0
    X = A.mm(Zs) + Es
    z = lambda r: torch.zeros(r, B)
    lip = float(torch.linalg.matrix_norm(A.t().double() @ A.double(), ord=2))
    for cls, fn in (("DLADMMNetScalar", orc.safeguarded_forward), ("DLADMMNetNewS", orc.safeguarded_forward_newS)):
        torch.manual_seed(5)
        model = getattr(dl, cls)(m=m, n=B, d=d, batch_size=B, A=A, Z0=z(d), E0=z(m), L0=z(m), layers=layers,
                                 precision=precision, device=dev)
        with torch.no_grad():
            for k in range(layers):
                model.fc[k].weight.mul_(0.9 / (0.4 * lip))
                model.active_para[k].fill_(0.02 + 0.01 * k)
                model.active_para1[k].fill_(0.25 + 0.1 * k)
                model.ss2[k].fill_(0.3 + 0.12 * k)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        ref = fn(sd, A, X, z(d), z(m), z(m), layers, True, True, False, None, 0, 0.05, "EMA", 0.5, 0.01, lip=lip)
        out = model.forward_safeguarded(X.to(dev), True, True, delta=0.05, mu_k_method="EMA", mu_k_param=0.5, alpha=0.01)
        nlists = 4 if cls == "DLADMMNetScalar" else 3
        tests = ref[-1]
        s, thr = torch.stack([t[0] for t in tests]), torch.stack([t[1] for t in tests])
        cols = ((s - thr).abs() / thr.abs().clamp_min(1e-9)).min(dim=0).values > 50 * TOL[precision]
        assert cols.float().mean() > 0.7
        for i in range(nlists):
            assert out[i][0].shape[1] == B
            got, want = torch.stack(list(out[i])).cpu(), torch.stack(list(ref[i]))
            assert rel_l2(got[:, :, cols], want[:, :, cols]) < TOL[precision], (cls, i)
        flips = sum(abs(a - b) for a, b in zip(out[nlists], ref[nlists]))
        assert flips <= (~cols).sum().item() * layers


def test_half_layer_calls_compose_to_a_full_forward():
    """dladmm_problem.start_half / stop_half: [Z-step of layer 0] + [E/L of 0, Z of 1] + [E/L of 1, Z of 2] + ... reproduces the
    iterates of one K-layer call bit for bit (same kernels, same operands)."""
    import dladmm_b200 as dl
    from dladmm_b200.function import LayerSpec, run_forward
    dev = torch.device("cuda:0")
    m, d, B, K = 40, 72, 132, 4
    torch.manual_seed(2)
    data = dl.gen_syn_data(B, m=m, d=d, seed=5)
    Z0 = torch.rand(d, B, device=dev) / d
    E0 = torch.zeros(m, B, device=dev); L0 = torch.zeros(m, B, device=dev)
    for precision in ("fp32", "tf32x3"):
        model = dl.DLADMMNetFull(m, 1, d, B, data.A, Z0, E0, L0, K, precision=precision)
        with torch.no_grad():
            for k in range(K):
                model.beta2[k].add_(0.1 * torch.randn_like(model.beta2[k]))
                model.active_para[k].add_(0.05 * torch.randn_like(model.active_para[k]))
        spec, params = model._spec_and_params()
        params = [p.detach() for p in params]
        with torch.no_grad():
            Zf, Ef, Lf, Tf, _, _ = run_forward(spec, model.A, data.X, Z0, E0, L0, params, want_masks=False)
            sub = lambda ks: LayerSpec(spec.family, m, d, len(ks), spec.precision, [spec.slots[k] for k in ks],
                                       [spec.weights[k] for k in ks], spec.fixed)
            Z, E, L, T, _, _ = run_forward(sub([0]), model.A, data.X, Z0, E0, L0, params, want_masks=False, stop_half=True)
            assert torch.equal(Z[0], Zf[0]) and torch.equal(T[0], Tf[0])
            Zc, Ec, Lc = Z[0], E0, L0
            for k in range(1, K):
                Z, E, L, T, _, _ = run_forward(sub([k - 1, k]), model.A, data.X, Zc, Ec, Lc, params, want_masks=False,
                                               start_half=True, stop_half=True)
                assert torch.equal(E[0], Ef[k - 1]) and torch.equal(L[0], Lf[k - 1]) and torch.equal(T[1], Tf[k]), (precision, k)
                assert torch.equal(Z[1], Zf[k]), (precision, k)
                Zc, Ec, Lc = Z[1], E[0], L[0]
            # K = 0: only T_0
            Z, E, L, T, _, _ = run_forward(sub([]), model.A, data.X, Z0, E0, L0, params, want_masks=False)
            assert T.shape[0] == 1 and torch.equal(T[0], Tf[0])


def test_select_update_kernel_matches_host_updaters():
    import dladmm_b200 as dl
    from dladmm_b200 import _lib
    from dladmm_b200.mu_updater import METHOD_ID
    dev = torch.device("cuda:0")
    B = 1003
    gen = torch.Generator().manual_seed(4)
    for method, param in (("EMA", 0.4), ("GS", 0.2), ("RT", 0.0), ("None", 0.0)):
        a, b = torch.randn(5, B, generator=gen).to(dev), torch.randn(5, B, generator=gen).to(dev)
        s, mu = (torch.rand(B, generator=gen) + 0.1).to(dev), (torch.rand(B, generator=gen) + 0.1).to(dev)
        host = dl.mu_updater_dict[method](mu.clone(), param)
        out, keep, fb = torch.empty_like(a), torch.empty(B, device=dev), torch.zeros(1, device=dev)
        pairs = (_lib.SgPair * 1)()
        pairs[0].a, pairs[0].b, pairs[0].out, pairs[0].rows = a.data_ptr(), b.data_ptr(), out.data_ptr(), 5
        mu_dev = mu.clone()
        _lib.check(_lib.load().dladmm_sg_select_update(1, pairs, B, s.data_ptr(), mu_dev.data_ptr(), 0.9, METHOD_ID[method], param,
                                                       keep.data_ptr(), fb.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        want_keep = (s < 0.9 * mu).float()
        assert torch.equal(keep, want_keep)
        assert torch.equal(out, torch.where(want_keep.bool().unsqueeze(0), a, b))
        assert fb.item() == float(B) - want_keep.sum().item()
        want_mu = host.step(s, want_keep)
        assert torch.allclose(mu_dev, want_mu, rtol=1e-6, atol=0), method
