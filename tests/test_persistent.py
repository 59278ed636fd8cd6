"""GPU: the all-layer persistent forward kernel (csrc/umma_persist.cuh: one launch, cross-CTA readiness counters per
(stage, batch tile)) against the per-layer launches (DLADMM_NO_PERSISTENT=1) -- same tiles, same MMA order, same epilogue
arithmetic, so every iterate and prox mask must be BIT-IDENTICAL -- for every family, scalar and per-row parameters, ragged
batches, inference (last_only ping-pong slabs) and training mode (masks + kept V_k), and through a backward."""
import os

import pytest
import torch

import dladmm_b200 as dl
from dladmm_b200 import _lib
from dladmm_b200.function import run_forward

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_persistent():
    """The library only takes the one-launch schedule for HBM-bound shapes (m*d <= 512K); force it for every case here."""
    os.environ["DLADMM_PERSISTENT"] = "1"
    yield
    os.environ.pop("DLADMM_PERSISTENT", None)


class per_layer(object):
    def __enter__(self):
        os.environ["DLADMM_NO_PERSISTENT"] = "1"

    def __exit__(self, *a):
        os.environ.pop("DLADMM_NO_PERSISTENT", None)


def _model(variant, m, d, B, K, precision, seed=3):
    torch.manual_seed(seed)
    data = dl.gen_syn_data(B, m=m, d=d, seed=40 + seed, dense_noise_sigma=(1.0 / m ** 0.5 if variant == "lasso" else None))
    bs = 20 if variant == "lena" else B
    Z0 = torch.rand(d, bs, device="cuda") / d
    E0 = torch.zeros(m, bs, device="cuda"); L0 = torch.zeros(m, bs, device="cuda")
    model = dl.VARIANT_CLASSES[variant](m, 1, d, bs, data.A, Z0, E0, L0, K, precision=precision)
    if variant == "lena":
        rep = B // 20
        model.Z0, model.E0, model.L0 = (t.repeat(1, rep).contiguous() for t in (model.Z0, model.E0, model.L0))
    with torch.no_grad():                       # move every parameter off its constant default so that rows / layers differ
        for n, p in model.named_parameters():
            if not n.startswith("fc"):
                p.mul_(1.0 + 0.2 * torch.rand_like(p))
    return model, data


CASES = [("scalar", 250, 500, 20000, 5, "tf32_bf16x2"), ("full", 250, 500, 1028, 4, "tf32_bf16x2"), ("lasso", 120, 300, 516, 3, "tf32_bf16x2"),
         ("scalar", 250, 500, 20000, 5, "tf32x3"), ("scalar", 250, 500, 132, 4, "tf32"), ("full", 250, 500, 1028, 4, "tf32x3"),
         ("tied", 96, 200, 4100, 6, "tf32x3"), ("lasso", 250, 500, 2048, 3, "tf32x3"), ("ltheta", 64, 100, 512, 3, "tf32x3"),
         ("lena", 256, 512, 20 * 60, 3, "tf32x3"), ("scalar", 40, 72, 24, 2, "tf32x3"), ("full", 1000, 2000, 256, 2, "tf32x3")]


@pytest.mark.parametrize("variant,m,d,B,K,precision", CASES)
def test_persistent_forward_is_bit_identical_to_per_layer_launches(variant, m, d, B, K, precision):
    model, data = _model(variant, m, d, B, K, precision)
    spec, params = model._spec_and_params()
    params = [p.detach() for p in params]
    for kw in (dict(want_masks=False), dict(want_masks=False, last_only=True), dict(want_masks=True, extras={})):
        for objective in (None, 0.01):
            if objective is not None and "extras" not in kw:
                kw = dict(kw, extras={})
            n0 = _lib.launch_count()
            a = run_forward(spec, model.A, data.X, model.Z0, model.E0, model.L0, params, objective_alpha=objective, **kw)
            n_pers = _lib.launch_count() - n0
            ea = dict(kw.get("extras") or {})
            with per_layer():
                n0 = _lib.launch_count()
                b = run_forward(spec, model.A, data.X, model.Z0, model.E0, model.L0, params, objective_alpha=objective, **kw)
                n_layer = _lib.launch_count() - n0
            eb = dict(kw.get("extras") or {})
            assert n_pers < n_layer and n_layer >= 2 * K + 1, (n_pers, n_layer)      # (the persistent schedule really ran)
            for name, x, y in zip(("Z", "E", "L", "T", "maskZ", "maskE"), a, b):
                assert (x is None) == (y is None)
                if x is not None:
                    assert torch.equal(x, y), (variant, name, kw.keys(), (x.float() - y.float()).abs().max().item())
            if objective is not None:
                assert torch.allclose(ea["objective"], eb["objective"], rtol=2e-5), (ea["objective"], eb["objective"])
            if kw.get("want_masks"):
                assert torch.equal(ea["Vsave"], eb["Vsave"])


def test_training_step_through_the_persistent_forward():
    model, data = _model("scalar", 250, 500, 4096, 5, "tf32x3")
    loss, _ = model.l1l1_loss(data.X, 0.01)
    loss.backward()
    g1 = [p.grad.clone() for p in model.parameters()]
    l1 = loss.item()
    model.zero_grad(set_to_none=True)
    with per_layer():
        loss, _ = model.l1l1_loss(data.X, 0.01)
        loss.backward()
    assert abs(loss.item() - l1) < 1e-5 * abs(l1)
    for a, p in zip(g1, model.parameters()):
        assert torch.allclose(a, p.grad, rtol=1e-4, atol=1e-7)


def test_many_calls_back_to_back_reuse_the_counters():
    """The readiness counters live in the per-call workspace and are zeroed on the stream: 20 forwards in a row, different
    batch sizes interleaved, all equal to the per-layer result."""
    model, data = _model("scalar", 250, 500, 8192, 6, "tf32x3")
    with per_layer(), torch.no_grad():
        ref = model(data.X)
    with torch.no_grad():
        for i in range(20):
            out = model(data.X)
            assert all(torch.equal(a, b) for a, b in zip(out[0] + out[3], ref[0] + ref[3])), i


def test_host_drain_returns_every_result_in_order():
    dev = torch.device("cuda:0")
    drain = dl.HostDrain(dev)
    want, got = [], []
    for i in range(7):
        t = torch.full((33, 1000), float(i), device=dev) + torch.arange(1000, device=dev)
        want.append(t.cpu())
        h = drain.push(t)
        del t
        if h is not None:
            got.append(h.clone()); drain.recycle(h)
    got += [h.clone() for h in drain.flush()]
    assert len(got) == 7 and all(torch.equal(a, b) for a, b in zip(got, want))
    assert drain.bytes_copied == 7 * 33 * 1000 * 4


def test_host_drain_keeps_its_pinned_buffers_across_loops():
    """A drain that outlives one loop recycles its page-locked buffers: after the first loop no new host buffer appears
    (page-locking 131 MB per step was 28 of the 31 ms of an earlier bench line) and every buffer is pinned."""
    dev = torch.device("cuda:0")
    drain = dl.HostDrain(dev)
    seen = []
    for loop in range(3):
        ptrs = set()
        for i in range(6):
            t = torch.full((64, 512), float(10 * loop + i), device=dev)
            h = drain.push(t)
            if h is not None:
                assert h.is_pinned() and float(h[0, 0]) == 10 * loop + i - 2
                ptrs.add(h.data_ptr()); drain.recycle(h)
        for h in drain.flush():
            ptrs.add(h.data_ptr()); drain.recycle(h)
        seen.append(ptrs)
    assert len(seen[0]) == 3                              # depth 2 + the one being handed to the caller
    assert seen[1] <= seen[0] and seen[2] <= seen[0]


class env(object):
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update(self.kw)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("variant,m,d,B,K", [("scalar", 250, 500, 2048, 3), ("full", 250, 500, 1100, 2), ("lasso", 300, 1000, 644, 2),
                                             ("ltheta", 64, 100, 516, 3), ("tied", 1000, 2000, 384, 1)])
def test_cta_pair_kernels_are_bit_identical_to_the_single_cta_kernels(variant, m, d, B, K):
    """tcgen05.mma.cta_group::2 (umma_pair.cuh): two CTAs share the weight tile of a stage.  Per element the MMAs are issued in the same
    order with the same operands as in the single-CTA kernel, so every iterate and prox mask must be bit-identical, and the fused objective and -- through the
    backward products, which take the same path -- every gradient must agree to summation order; odd numbers of batch tiles, ragged
    batches, per-row parameters, several feature tiles.  (Default use: reduction lengths >= 768; DLADMM_PAIR=1 forces it here.)"""
    model, data = _model(variant, m, d, B, K, "tf32_bf16x2")
    spec, params = model._spec_and_params()
    det = [p.detach() for p in params]
    res = {}
    for pair in ("0", "1"):
        with env(DLADMM_PAIR=pair, DLADMM_NO_PERSISTENT="1"):
            ex = {}
            out = run_forward(spec, model.A, data.X, model.Z0, model.E0, model.L0, det, want_masks=True, objective_alpha=0.01, extras=ex)
            model.zero_grad(set_to_none=True)
            if variant == "lasso":
                loss, _ = model.lasso_loss(data.X, 0.01)
            elif variant == "ltheta":
                Z, E, L = model(data.X)
                loss = sum(z.abs().sum() for z in Z) + sum(e.abs().sum() for e in E)
            else:
                loss, _ = model.l1l1_loss(data.X, 0.01)
            loss.backward()
            res[pair] = (out, ex["objective"].clone(), loss.item(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    (o0, j0, l0, g0), (o1, j1, l1, g1) = res["0"], res["1"]
    for name, x, y in zip(("Z", "E", "L", "T", "maskZ", "maskE"), o0, o1):
        assert (x is None) == (y is None)
        if x is not None:
            assert torch.equal(x, y), (variant, name)
    # (the fused objective is a sum of per-CTA partial sums: another tile -> CTA assignment, another summation order)
    assert torch.allclose(j0, j1, rtol=2e-6) and abs(l0 - l1) <= 2e-6 * abs(l0)
    for n in g0:
        if n.startswith("fc"):      # dW goes through fp32 reductions in L2 in whatever order the CTAs arrive
            assert torch.allclose(g0[n], g1[n], rtol=1e-4, atol=1e-6 * float(g0[n].abs().max())), n
        else:
            assert torch.allclose(g0[n], g1[n], rtol=2e-5, atol=1e-7 * max(1.0, float(g0[n].abs().max()))), n
