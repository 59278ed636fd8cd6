"""CPU: host-side logic -- module construction, state_dict layout against the fixtures, the no-fallback
rule, column sharding, and the gradient allreduce over gloo with world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dladmm_b200 as dl
from _util import GOLDEN_NAMES, Golden, build_model


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_state_dict_layout_matches_reference_fixture(name):
    g = Golden(name)
    model = dl.VARIANT_CLASSES[g.variant](m=g.m, n=1, d=g.d, batch_size=g.bs, A=g.A, Z0=g.Z0, E0=g.E0, L0=g.L0,
                                          layers=g.K, device="cpu")
    sd = model.state_dict()
    assert list(sd.keys()) == g.keys                       # same names, same registration order
    for k in g.keys:
        assert tuple(sd[k].shape) == tuple(g.sd[k].shape), k
    model.load_state_dict(g.sd)
    assert all(torch.equal(model.state_dict()[k], g.sd[k]) for k in g.keys)
    # A, Z0, E0, L0 are plain attributes, not in the state_dict (reference: main_syn_l1l1_scalar.py:40-44)
    assert not any(k.startswith(("A", "Z0", "E0", "L0")) for k in sd)
    assert torch.equal(model.A, g.A) and torch.equal(model.E0, g.E0)


def test_forward_refuses_cpu_tensors():
    g = Golden("scalar_small")
    model = dl.DLADMMNetScalar(g.m, 1, g.d, g.bs, g.A, g.Z0, g.E0, g.L0, g.K, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(g.X)


def test_product_never_imports_the_oracle():
    """No import / include / path reference to oracle/ anywhere in the product package (comments may cite it)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "d-ladmm_b200")
    bad = re.compile(r"^\s*(import|from)\s+\S*(dladmm_oracle|load_reference|make_golden)|#\s*include[^\n]*oracle|[\"']oracle[/\"']|sys\.path[^\n]*oracle",
                     re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), os.path.join(dirpath, f)


def test_lipschitz_constant_is_lazy_and_correct():
    g = Golden("scalar_small")
    model = dl.DLADMMNetScalar(g.m, 1, g.d, g.bs, g.A, g.Z0, g.E0, g.L0, g.K, device="cpu")
    assert model._lipschitz is None
    ref = torch.linalg.matrix_norm(g.A.t() @ g.A, ord=2)
    assert tuple(model.L.shape) == (1, 1) and abs(model.L.item() - ref.item()) < 1e-4


@pytest.mark.parametrize("B,world", [(10, 3), (65536, 8), (7, 8), (0, 2), (1 << 20, 8)])
def test_column_shard_tiles_the_batch(B, world):
    spans = [dl.column_shard(B, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == B
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dladmm_b200 as dl2
        g = Golden("full_small")
        model = dl2.DLADMMNetFull(g.m, 1, g.d, g.bs, g.A, g.Z0, g.E0, g.L0, g.K, device="cpu")
        model.load_state_dict(g.sd)
        # each rank holds the gradient contribution of its own column shard: emulate with a known split
        gen = torch.Generator().manual_seed(100 + rank)
        for p in model.parameters():
            p.grad = torch.randn(p.shape, generator=gen)
        local = [p.grad.clone() for p in model.parameters()]
        calls = dl2.allreduce_gradients(list(model.parameters()), bucket_bytes=4096)
        assert calls >= 2              # bucketing splits the ~30 KB of gradients
        # expected: sum over ranks of the per-rank draws
        exp = []
        for i, p in enumerate(model.parameters()):
            tot = torch.zeros_like(p)
            for r in range(world):
                gr = torch.Generator().manual_seed(100 + r)
                draws = [torch.randn(q.shape, generator=gr) for q in model.parameters()]
                tot += draws[i]
            exp.append(tot)
        ok = all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(model.parameters(), exp))
        start, stop = dl2.column_shard(g.bs, rank, world)
        ret[rank] = (ok, start, stop, float(local[0].sum()))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_gloo_world_size_2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world))
    assert ret[0][1] == 0 and ret[0][2] == ret[1][1]


def test_mat_file_round_trip_in_reference_layout(tmp_path):
    """save_mat writes gen_syn_data.py:49-53's layout (sample-major, keys A/train_*/test_*); load_mat reads it back into
    the model's (features x B) layout; replace_A_columns follows gen_syn_unseen_data_Acols.py:14-26."""
    import scipy.io as sio
    import dladmm_b200 as dl
    g = torch.Generator().manual_seed(3)
    m, d = 12, 20
    A = torch.randn(m, d, generator=g); A = A / A.pow(2).sum(0, keepdim=True).sqrt()
    mk = lambda B: dl.SynData(A, torch.randn(m, B, generator=g), torch.randn(d, B, generator=g), torch.randn(m, B, generator=g))
    train, test = mk(30), mk(7)
    path = str(tmp_path / "syn_data.mat")
    dl.save_mat(path, train, test)
    raw = sio.loadmat(path)
    assert raw["train_x"].shape == (30, m) and raw["test_z"].shape == (7, d) and raw["A"].shape == (m, d)   # sample-major
    tr2, te2 = dl.load_mat(path)
    for a, b in ((tr2.X, train.X), (tr2.Z, train.Z), (tr2.E, train.E), (te2.X, test.X), (te2.Z, test.Z), (te2.E, test.E)):
        assert torch.equal(a, b)
    A2, cols = dl.replace_A_columns(A, 5, seed=1)
    keep = torch.ones(d, dtype=torch.bool); keep[cols] = False
    assert torch.equal(A2[:, keep], A[:, keep]) and cols.numel() == 5
    assert (A2.pow(2).sum(0).sqrt() - 1).abs().max() < 1e-6


def test_call_description_cache_tracks_replaced_parameters():
    """_spec_and_params caches the static call description; replacing a Parameter object (as the constructor itself does
    for the fc weights) or moving the module must not leave a stale entry behind."""
    import dladmm_b200 as dl
    g = Golden("scalar_small")
    model = build_model(g, "cpu")
    spec1, params1 = model._spec_and_params()
    spec2, params2 = model._spec_and_params()
    assert spec1 is spec2 and all(a is b for a, b in zip(params1, params2))          # cache hit
    model.fc[1].weight = torch.nn.Parameter(torch.zeros_like(model.fc[1].weight))
    spec3, params3 = model._spec_and_params()
    assert any(p is model.fc[1].weight for p in params3) and spec3 is not spec1
    model.float()                                                                    # _apply drops the cache
    assert "_spec_cache" not in model.__dict__
    news = dl.DLADMMNetNewS(g.m, 1, g.d, g.bs, g.A, g.Z0, g.E0, g.L0, 4, device="cpu")
    _, pa = news._spec_and_params(3, drop_last_estep=True)
    _, pb = news._spec_and_params(3, drop_last_estep=True)
    nograd = [p for p in pb if not p.requires_grad]
    assert len(nograd) == 4 and all(any(q.data_ptr() == p.data_ptr() for q in news.parameters()) for p in nograd)


def test_auto_precision_resolves_by_shape(monkeypatch):
    """precision="auto" (the default): the split-precision mode from min(m, d) >= 192, three-pass tf32x3 below
    (net.py::resolve_precision; error models in DESIGN.md 3.2); DLADMM_PRECISION overrides the default, an explicit
    precision= overrides both; unknown names are refused."""
    import dladmm_b200 as dl
    monkeypatch.delenv("DLADMM_PRECISION", raising=False)
    assert dl.default_precision() == "auto"
    assert dl.resolve_precision("auto", 250, 500) == "tf32_bf16x2" and dl.resolve_precision("auto", 1000, 2000) == "tf32_bf16x2"
    assert dl.resolve_precision("auto", 24, 40) == "tf32x3" and dl.resolve_precision("auto", 191, 4000) == "tf32x3"
    assert dl.resolve_precision("bf16", 24, 40) == "bf16"
    g = Golden("scalar_small")
    assert build_model(g, "cpu").precision == "tf32x3"                         # m = 20..64 fixtures
    big = dl.DLADMMNetScalar(250, 1, 500, 4, torch.randn(250, 500), torch.zeros(500, 4), torch.zeros(250, 4), torch.zeros(250, 4), 2,
                             device="cpu")
    assert big.precision == "tf32_bf16x2"
    monkeypatch.setenv("DLADMM_PRECISION", "tf32")
    assert build_model(g, "cpu").precision == "tf32" and build_model(g, "cpu", "fp32").precision == "fp32"
    with pytest.raises(ValueError):
        build_model(g, "cpu", "fp16")


def _cpu_model(cls, m, d, K, **kw):
    A = torch.randn(m, d)
    z = lambda r: torch.zeros(r, 4)
    return cls(m, 1, d, 4, A, z(d), z(m), z(m), K, device="cpu", **kw)


def test_gradient_bucket_plan_layout():
    """Bucketed in-backward allreduce: weights first in the order the backward finishes them (layer K-1 first),
    contiguous buckets of at least the bucket size that tile the weight part, each released by its LOWEST layer's event;
    below two buckets there is no plan (one collective after the backward)."""
    from dladmm_b200.function import plan_gradient_buckets, _flat_zero_grads
    m, d, K = 40, 64, 10
    model = _cpu_model(dl.DLADMMNetScalar, m, d, K)
    spec, params = model._spec_and_params()
    needs = [True] * len(params)
    wbytes = 4 * m * d
    assert plan_gradient_buckets(spec, params, needs) is None                     # opt-in: no bucket size, no plan
    assert plan_gradient_buckets(spec, params, needs, bucket_bytes=K * wbytes) is None
    plan = plan_gradient_buckets(spec, params, needs, bucket_bytes=3 * wbytes)
    # 10 weights, 3 per bucket, the undersized tail merged into the last bucket: layers 9-7, 6-4, 3-0
    assert [b[0] for b in plan.buckets] == [7, 4, 0]
    assert plan.buckets[0][1] == 0 and plan.buckets[-1][2] == plan.rest == K * m * d
    assert all(a[2] == b[1] for a, b in zip(plan.buckets, plan.buckets[1:]))
    assert [spec.weights.index(i) for i in plan.order[:K]] == list(range(K - 1, -1, -1))
    assert sorted(plan.order) == list(range(len(params)))
    grads, flat = _flat_zero_grads(params, needs, plan.order)
    for k in range(K):                                   # layer k's weight gradient sits in the bucket its event releases
        g = grads[spec.weights[k]]
        off = (g.data_ptr() - flat.data_ptr()) // 4
        layer, lo, hi = next(b for b in plan.buckets if b[0] <= k)
        assert lo <= off and off + g.numel() <= hi, (k, off, lo, hi)
    assert all(g.shape == p.shape for g, p in zip(grads, params))
    small = [g for i, g in enumerate(grads) if i not in set(spec.weights)]
    assert all((g.data_ptr() - flat.data_ptr()) // 4 >= plan.rest for g in small)
    # a weight without a gradient is left out of the plan and of the buffer
    needs2 = list(needs); needs2[spec.weights[K - 1]] = False
    plan2 = plan_gradient_buckets(spec, params, needs2, bucket_bytes=3 * wbytes)
    assert plan2.rest == (K - 1) * m * d and spec.weights[K - 1] not in plan2.order[:K - 1]


def test_gradient_bucket_plan_tied_weight_waits_for_layer_zero():
    from dladmm_b200.function import plan_gradient_buckets
    model = _cpu_model(dl.DLADMMNetTied, 40, 64, 6)
    spec, params = model._spec_and_params()
    assert len(set(spec.weights)) == 1
    plan = plan_gradient_buckets(spec, params, [True] * len(params), bucket_bytes=1)
    assert plan is None or [b[0] for b in plan.buckets] == [0]     # the one shared weight is final after layer 0 only
