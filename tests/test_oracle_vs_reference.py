"""CPU, build container only: the oracle and the module's state_dict layout against the reference's own
classes imported live from /root/reference (skipped where the reference is absent, e.g. the GPU box)."""
import pytest
import torch

import dladmm_oracle as orc
import load_reference as lr
from _util import syn

pytestmark = pytest.mark.skipif(not lr.reference_available(), reason="reference tree not present")

VARIANTS = list(orc.VARIANTS)


def _setup(variant, m=30, d=52, bs=10, K=5, seed=0):
    torch.manual_seed(seed)
    A, X = syn(m, d, bs, seed)
    Z0 = torch.rand(d, bs) / d
    E0 = torch.zeros(m, bs); L0 = torch.zeros(m, bs)
    ref = lr.build(variant, m, 10000, d, bs, A, Z0, E0, L0, K)
    return ref, A, X, Z0, E0, L0, K


@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_forward_bit_exact_with_reference_on_this_host(variant):
    ref, A, X, Z0, E0, L0, K = _setup(variant)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    Zr, Er, Lr, Tr = lr.run(ref, X)
    Z, E, L, T = orc.forward(variant, sd, A, X, Z0, E0, L0, K)
    for k in range(K):
        assert torch.equal(Z[k], Zr[k]) and torch.equal(E[k], Er[k]) and torch.equal(L[k], Lr[k])
    if Tr is not None:
        assert len(Tr) == K + 1 and all(torch.equal(a, b) for a, b in zip(T, Tr))
    assert orc.VARIANTS[variant]["returns_T"] == (Tr is not None)


@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_default_init_matches_reference(variant):
    ref, A, X, Z0, E0, L0, K = _setup(variant)
    sd = ref.state_dict()
    mine = orc.default_state_dict(variant, A, K, X.shape[1])
    assert list(mine.keys()) != [] and set(mine) == set(sd)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(mine[k].shape), k
        if not k.startswith("fc"):
            assert torch.equal(sd[k], mine[k]), k
        else:   # W = (A^T + 1e-3 randn) * scale: same up to the noise draw
            assert (sd[k] - mine[k]).abs().max() < 1e-2


@pytest.mark.parametrize("variant", VARIANTS)
def test_module_state_dict_is_interchangeable_with_reference(variant):
    """8(b): keys, shapes and registration order are the compatibility contract (checkpoints)."""
    import dladmm_b200 as dl
    ref, A, X, Z0, E0, L0, K = _setup(variant)
    model = dl.VARIANT_CLASSES[variant](m=A.shape[0], n=10000, d=A.shape[1], batch_size=X.shape[1], A=A, Z0=Z0,
                                        E0=E0, L0=L0, layers=K, device="cpu")
    sd_ref, sd_new = ref.state_dict(), model.state_dict()
    assert list(sd_ref.keys()) == list(sd_new.keys())
    for k in sd_ref:
        assert tuple(sd_ref[k].shape) == tuple(sd_new[k].shape) and sd_ref[k].dtype == sd_new[k].dtype, k
    model.load_state_dict(sd_ref)                    # reference checkpoint -> new module
    ref.load_state_dict(model.state_dict())          # and back
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in model.named_parameters()]
    assert model.name() == ref.name()
    # default initialisation of the non-weight parameters is the reference's
    fresh = dl.VARIANT_CLASSES[variant](m=A.shape[0], n=1, d=A.shape[1], batch_size=X.shape[1], A=A, Z0=Z0, E0=E0,
                                        L0=L0, layers=K, device="cpu")
    ref2 = lr.build(variant, A.shape[0], 1, A.shape[1], X.shape[1], A, Z0, E0, L0, K)
    for (n1, p1), (n2, p2) in zip(fresh.named_parameters(), ref2.named_parameters()):
        if not n1.startswith("fc"):
            assert torch.equal(p1, p2), n1
        else:
            scale = 1.0 if variant in ("lena", "ltheta") else 0.4
            assert (p1 - A.t() * scale).abs().max() < 1e-2


def test_golden_fixtures_are_reference_outputs():
    """Regenerate one fixture from the live reference and compare with the committed file."""
    from _util import Golden
    g = Golden("scalar_small")
    ref = lr.build(g.variant, g.m, 10000, g.d, g.bs, g.A, g.Z0, g.E0, g.L0, g.K)
    ref.load_state_dict(g.sd)
    Z, E, L, T = lr.run(ref, g.X)
    for k in range(g.K):
        assert torch.allclose(Z[k], g.Z[k], rtol=0, atol=1e-6) and torch.allclose(T[k + 1], g.T[k + 1], rtol=0, atol=1e-6)
