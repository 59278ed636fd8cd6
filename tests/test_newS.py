"""the E -> L -> Z ("newS") ordering with prefix execution forward(x, K) and its tied / partially tied
(fc[k // interval]) twins -- main_syn_scalar_newS_layerwise.py:82-112, main_syn_scalar_tied_newS_layerwise.py,
main_syn_scalar_ptied_newS_layerwise.py.  Fixtures are outputs and autograd gradients of the unmodified reference classes
(oracle/make_golden_newS.py)."""
import pytest
import torch

import dladmm_oracle as orc
import dladmm_b200 as dl
from _util import NEWS_GOLDEN_NAMES, NewSGolden, rel_l2

FWD_TOL = {"fp32": 5e-6, "tf32x3": 2e-5}
GRAD_TOL = {"fp32": 2e-4, "tf32x3": 5e-4}


@pytest.mark.parametrize("name", NEWS_GOLDEN_NAMES)
def test_oracle_newS_matches_reference_outputs_and_gradients(name):
    g = NewSGolden(name)
    Z, E, L = orc.forward_newS(g.variant, g.sd, g.A, g.X, g.Z0, g.E0, g.L0, g.K_run, g.layers, g.interval or None)
    n = min(g.K_run, g.layers)
    assert len(Z) == len(E) == len(L) == n
    for k in range(n):
        assert rel_l2(Z[k], g.Z[k]) < 2e-6 and rel_l2(E[k], g.E[k]) < 2e-6 and rel_l2(L[k], g.L[k]) < 2e-6, (name, k)
    assert torch.equal(E[0], g.E0) and torch.equal(L[0], g.L0)
    sd = {k: v.double().requires_grad_(True) for k, v in g.sd.items()}
    d64 = lambda t: t.double()
    Z, E, L = orc.forward_newS(g.variant, sd, d64(g.A), d64(g.X), d64(g.Z0), d64(g.E0), d64(g.L0), g.K_run, g.layers,
                               g.interval or None)
    loss = sum((Z[k] * g.cz[k]).sum() + (E[k] * g.ce[k]).sum() + (L[k] * g.cl[k]).sum() for k in range(n))
    loss.backward()
    assert abs(loss.item() - g.loss) < 1e-4 * max(1.0, abs(g.loss))
    for k in g.keys:
        if k in g.nograd:
            assert sd[k].grad is None or sd[k].grad.abs().max() == 0, k
        else:
            assert rel_l2(sd[k].grad, g.grads[k], floor=1e-6) < 5e-4, (name, k)


@pytest.mark.parametrize("name", NEWS_GOLDEN_NAMES)
def test_newS_module_mirrors_reference_state_dict_and_name(name):
    g = NewSGolden(name)
    model = g.build("cpu")
    assert list(model.state_dict().keys()) == g.keys                      # names AND registration order
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(g.sd[k].shape), k
    assert model.name() == g.ref_name
    fresh = dl.VARIANT_CLASSES[g.variant](g.m, 1, g.d, g.bs, g.A, g.Z0, g.E0, g.L0, g.layers, device="cpu",
                                          **({"interval": g.interval} if g.interval else {}))
    th = 0.1 if g.variant == "newS" else 0.01                             # newS.py:65-66, tied_newS.py:66-67
    assert all(abs(p.item() - th) < 1e-7 for p in list(fresh.active_para) + list(fresh.active_para1))
    w = fresh._weight_module(0).weight
    assert rel_l2(w, 0.4 * g.A.t()) < 2e-2                                # (A^T + 1e-3 randn) * 0.4, newS.py:75
    with pytest.raises(RuntimeError, match="CUDA"):
        model(g.X, g.K_run)


def test_ptied_needs_interval_and_others_refuse_it():
    g = NewSGolden("ptied_newS_small")
    with pytest.raises(ValueError):
        dl.DLADMMNetPtiedNewS(g.m, 1, g.d, g.bs, g.A, g.Z0, g.E0, g.L0, g.layers, device="cpu")
    with pytest.raises(ValueError):
        dl.DLADMMNetNewS(g.m, 1, g.d, g.bs, g.A, g.Z0, g.E0, g.L0, g.layers, interval=2, device="cpu")
    model = g.build("cpu")
    assert len(model.fc) == g.layers // g.interval and model._weight_module(4) is model.fc[4 // g.interval]


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name", NEWS_GOLDEN_NAMES)
def test_newS_forward_and_gradients_match_reference_on_gpu(name, precision):
    g = NewSGolden(name)
    model = g.build("cuda", precision)
    Z, E, L = model(g.X.cuda(), g.K_run)
    n = min(g.K_run, g.layers)
    assert len(Z) == len(E) == len(L) == n
    tol = FWD_TOL[precision]
    for k in range(n):
        assert rel_l2(Z[k].cpu(), g.Z[k], floor=1e-3) < tol and rel_l2(E[k].cpu(), g.E[k], floor=1e-3) < tol, (name, k)
        assert rel_l2(L[k].cpu(), g.L[k], floor=1e-3) < 5 * tol, (name, k)
    loss = sum((Z[k] * g.cz[k].cuda()).sum() + (E[k] * g.ce[k].cuda()).sum() + (L[k] * g.cl[k].cuda()).sum() for k in range(n))
    loss.backward()
    assert abs(loss.item() - g.loss) < 1e-3 * max(1.0, abs(g.loss))
    for nme, p in model.named_parameters():
        if nme in g.nograd:
            assert p.grad is None or p.grad.abs().max().item() == 0, nme   # the reference leaves these without a gradient
        else:
            assert rel_l2(p.grad.cpu(), g.grads[nme], floor=1e-5) < GRAD_TOL[precision], (name, nme)


@pytest.mark.gpu
def test_newS_objective_and_loss_follow_the_variant_conventions():
    """forward_objective / l1l1_loss on a newS module: the objective depends on Z_k only (identical lists), the returned
    E / L lists use the variant's own shifted convention, and the Z -> E -> L safeguard entry points refuse the variant."""
    g = NewSGolden("newS_small")
    model = g.build("cuda")
    x = g.X.cuda()
    with torch.no_grad():
        Z, E, L = model(x, g.layers)
    obj, (Zo, Eo, Lo) = model.forward_objective(x, 0.01)
    assert all(torch.equal(a, b) for a, b in zip(Z, Zo)) and all(torch.equal(a, b) for a, b in zip(E, Eo))
    assert all(torch.equal(a, b) for a, b in zip(L, Lo)) and torch.equal(Eo[0], model.E0)
    A = g.A.cuda()
    exp = torch.stack([0.01 * Z[k].abs().sum() + (x - A @ Z[k]).abs().sum() for k in range(g.layers)])
    assert rel_l2(obj, exp) < 1e-4
    loss, outs = model.l1l1_loss(x, 0.01)
    assert len(outs) == 3 and abs(loss.item() - exp.sum().item() / x.shape[1]) < 1e-4 * abs(loss.item())
    with pytest.raises(NotImplementedError):        # the Z -> E -> L entry points refuse the E -> L -> Z variants (and vice versa)
        model.KM(model.Z0, model.E0, model.L0, model.E0, x)
