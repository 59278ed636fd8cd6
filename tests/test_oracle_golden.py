"""CPU: the oracle restatement against the committed reference outputs (tests/golden, produced by
oracle/make_golden.py from the unmodified reference classes), the hand-derived backward against the
reference's autograd gradients, and the known-answer invariants of SURVEY.md section 4."""
import pytest
import torch

import dladmm_oracle as orc
from _util import GOLDEN_NAMES, Golden, rel_l2, syn

TOL = 2e-6   # fp32 CPU matrix products may differ in accumulation order between hosts


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_forward_matches_reference_outputs(name):
    g = Golden(name)
    Z, E, L, T = orc.forward(g.variant, g.sd, g.A, g.X, g.Z0, g.E0, g.L0, g.K)
    for k in range(g.K):
        assert rel_l2(Z[k], g.Z[k]) < TOL and rel_l2(E[k], g.E[k]) < TOL and rel_l2(L[k], g.L[k]) < TOL, (name, k)
        if g.T is not None:
            assert rel_l2(T[k + 1], g.T[k + 1], floor=1e-3 * g.X.norm().item()) < 10 * TOL
    assert len(Z) == len(E) == len(L) == g.K and len(T) == g.K + 1          # a10


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_autograd_of_oracle_matches_reference_gradients(name):
    g = Golden(name)
    loss_fn = lambda Z, E, L, T: orc.all_iterates_loss(Z, E, L, T, g.cz, g.ce, g.cl, g.ct)
    loss, grads = orc.autograd_grads(g.variant, g.sd, g.A, g.X, g.Z0, g.E0, g.L0, g.K, loss_fn)
    assert abs(loss.item() - g.loss) <= 1e-4 * max(1.0, abs(g.loss))
    for k in g.keys:
        assert rel_l2(grads[k], g.grads[k], floor=1e-6) < 5e-4, (name, k)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_manual_backward_matches_reference_gradients(name):
    """The recurrences the CUDA backward implements (fp64 here) reproduce the reference's autograd."""
    g = Golden(name)
    d64 = lambda t: t.double()
    sd = {k: d64(v) for k, v in g.sd.items()}
    grads = orc.manual_backward(g.variant, sd, d64(g.A), d64(g.X), d64(g.Z0), d64(g.E0), d64(g.L0), g.K,
                                list(d64(g.cz)), list(d64(g.ce)), list(d64(g.cl)), list(d64(g.ct)))
    for k in g.keys:
        assert rel_l2(grads[k], g.grads[k], floor=1e-6) < 5e-4, (name, k)


@pytest.mark.parametrize("variant", ["scalar", "full", "tied"])
def test_learned_forward_equals_classical_ladmm_under_km_parameters(variant):
    """W = ss1*A^T, betas = 1, theta1 = ss1*alpha, theta2 = ss2 makes the learned
    family-B layer the classical KM iteration (main_syn_l1l1_scalar.py:134-160)."""
    m, d, B, K, alpha = 50, 100, 16, 15, 0.01
    A, X = syn(m, d, B, seed=3)
    Z0 = torch.zeros(d, B); E0 = torch.zeros(m, B); L0 = torch.zeros(m, B)
    sd, ss1 = orc.km_state_dict(variant, A, K, alpha)
    Z, E, L, T = orc.forward(variant, sd, A, X, Z0, E0, L0, K)
    Zk, Ek, Lk, Tk = orc.km_iterations(A, X, Z0, E0, L0, K, alpha, 1.0, ss1, 0.3)
    for k in range(K):
        assert (Z[k] - Zk[k]).abs().max() < 5e-6 and (E[k] - Ek[k]).abs().max() < 5e-6
        assert (L[k] - Lk[k]).abs().max() < 5e-6


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_residual_identity(name):
    """X - A Z_k == E_k - T_{k+1}."""
    g = Golden(name)
    Z, E, L, T = orc.forward(g.variant, g.sd, g.A, g.X, g.Z0, g.E0, g.L0, g.K)
    for k in range(g.K):
        assert ((g.X - g.A.mm(Z[k])) - (E[k] - T[k + 1])).abs().max() < 2e-5


def test_km_objective_decreases():
    """the classical LADMM L1-L1 objective decreases after the first iterations."""
    m, d, B, K, alpha = 50, 100, 32, 60, 0.01
    A, X = syn(m, d, B, seed=5)
    Z0 = torch.zeros(d, B); E0 = torch.zeros(m, B); L0 = torch.zeros(m, B)
    Lip = torch.linalg.matrix_norm(A.t().double() @ A.double(), ord=2).item()
    Z, E, L, T = orc.km_iterations(A, X, Z0, E0, L0, K, alpha, 1.0, 0.5 / Lip, 0.3)
    obj = [(alpha * Z[k].abs().sum() + (X - A.mm(Z[k])).abs().sum()).item() / B for k in range(K)]
    assert obj[-1] < obj[5] < obj[0]
