"""GPU: size-independent properties at BASELINE sizes, edge cases (empty / ragged batches), last_only mode,
the on-device generator's statistics, the fused objective, and the no-fallback rule."""
import math

import pytest
import torch

import dladmm_oracle as orc
import dladmm_b200 as dl
from _util import Golden, build_model, rel_l2, syn

pytestmark = pytest.mark.gpu


def _model(variant, m, d, B, K, seed=0, precision=None, Z0=None):
    torch.manual_seed(seed)
    data = dl.gen_syn_data(B, m=m, d=d, seed=1126 + seed)
    if Z0 is None:
        Z0 = torch.rand(d, B, device="cuda") / d
    E0 = torch.zeros(m, B, device="cuda"); L0 = torch.zeros(m, B, device="cuda")
    model = dl.VARIANT_CLASSES[variant](m, 1, d, B, data.A, Z0, E0, L0, K, precision=precision)
    return model, data


@pytest.mark.parametrize("variant", ["scalar", "full", "tied", "lasso", "ltheta"])
def test_residual_identity_at_config1_size(variant):
    """X - A Z_k == E_k - T_{k+1} for every layer at m=250, d=500, K=15, B=16384."""
    model, data = _model(variant, 250, 500, 16384, 15)
    with torch.no_grad():
        out = model(data.X)
        Z, E = out[0], out[1]
        # recompute T from L for variants that do not return it: L_k = L_{k-1} + bL*T_{k+1}
        for k in (0, 7, 14):
            R = data.A @ Z[k]
            if len(out) == 4:
                T = out[3]
                lhs = data.X - R
                rhs = E[k] - T[k + 1]
                assert (lhs - rhs).abs().max().item() < 5e-5
            assert torch.isfinite(Z[k]).all() and torch.isfinite(E[k]).all()


def test_columns_are_independent_and_shards_tile_the_batch():
    """Running column shards separately reproduces the full-batch run bit for bit (8(e): no collective on
    the data path), for a ragged split."""
    m, d, B, K = 250, 500, 1200, 6           # shards of 600 / 400 columns keep the same (tensor-core) kernels
    model, data = _model("scalar", m, d, B, K)
    with torch.no_grad():
        Zf, Ef, Lf, Tf = model(data.X)
    for world in (2, 3):
        for r in range(world):
            a, b = dl.column_shard(B, r, world)
            sub = dl.DLADMMNetScalar(m, 1, d, b - a, data.A, model.Z0[:, a:b].contiguous(),
                                     model.E0[:, a:b].contiguous(), model.L0[:, a:b].contiguous(), K)
            sub.load_state_dict(model.state_dict())
            with torch.no_grad():
                Zs, Es, Ls, Ts = sub(data.X[:, a:b].contiguous())
            assert torch.equal(Zs[-1], Zf[-1][:, a:b]) and torch.equal(Ls[-1], Lf[-1][:, a:b])
            assert torch.equal(Ts[K], Tf[K][:, a:b])


@pytest.mark.parametrize("B", [1, 3, 20, 25, 255, 257, 1030])
def test_ragged_batch_sizes(B):
    m, d, K = 250, 500, 3
    A, X = syn(m, d, B, seed=B)
    Z0 = torch.rand(d, B) / d; E0 = torch.zeros(m, B); L0 = torch.zeros(m, B)
    model = dl.DLADMMNetScalar(m, 1, d, B, A, Z0, E0, L0, K)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        Z, E, L, T = model(X.cuda())
    Zo, Eo, Lo, To = orc.forward("scalar", sd, A, X, Z0, E0, L0, K)
    for k in range(K):
        assert rel_l2(Z[k].cpu(), Zo[k], floor=1e-3) < 2e-5 and rel_l2(L[k].cpu(), Lo[k], floor=1e-2) < 2e-5


def test_empty_batch():
    m, d, K = 24, 40, 2
    A = torch.randn(m, d)
    z = lambda r: torch.zeros(r, 0)
    model = dl.DLADMMNetScalar(m, 1, d, 0, A, z(d), z(m), z(m), K)
    with torch.no_grad():
        Z, E, L, T = model(torch.zeros(m, 0, device="cuda"))
    assert len(Z) == K and Z[0].shape == (d, 0) and len(T) == K + 1


def test_last_only_equals_full_run():
    model, data = _model("scalar", 250, 500, 2048, 15)
    with torch.no_grad():
        Z, E, L, T = model(data.X)
        Zl, El, Ll, Tl = model(data.X, last_only=True)
    assert torch.equal(Zl[0], Z[-1]) and torch.equal(El[0], E[-1]) and torch.equal(Ll[0], L[-1])
    assert torch.equal(Tl[0], T[-1])
    with pytest.raises(RuntimeError, match="inference-only"):
        model(data.X, last_only=True)


def test_lena_slot_parameters_repeat_over_larger_batches():
    """(m x bs) betas are indexed by batch slot; column c uses slot c mod bs."""
    g = Golden("lena_small")
    reps = 3
    B = g.bs * reps
    model = dl.DLADMMNetLena(g.m, 1, g.d, g.bs, g.A, g.Z0.repeat(1, reps), g.E0.repeat(1, reps),
                             g.L0.repeat(1, reps), g.K)
    model.load_state_dict(g.sd)
    out = model(g.X.repeat(1, reps).cuda())
    for k in range(g.K):
        for r in range(reps):
            assert rel_l2(out[0][k][:, r * g.bs:(r + 1) * g.bs].cpu(), g.Z[k], floor=1e-3) < 3e-5
    # gradients of the per-slot parameters accumulate over the repeats
    loss = sum((out[0][k] * g.cz[k].repeat(1, reps).cuda()).sum() + (out[1][k] * g.ce[k].repeat(1, reps).cuda()).sum()
               + (out[2][k] * g.cl[k].repeat(1, reps).cuda()).sum() for k in range(g.K))
    loss.backward()
    for n, p in model.named_parameters():
        assert rel_l2(p.grad.cpu(), reps * g.grads[n], floor=1e-5) < 1e-3, n


def test_generator_statistics_and_shard_reproducibility():
    """gen_syn_data.py semantics: unit-norm columns, Bernoulli(p) support, N(0, sigma) amplitudes, X = AZ+E."""
    m, d, B, p, sigma = 250, 500, 8192, 0.1, 2.0
    data = dl.gen_syn_data(B, m=m, d=d, p=p, sigma=sigma, seed=7)
    A, X, Z, E = data
    assert (A.pow(2).sum(dim=0).sqrt() - 1).abs().max().item() < 1e-5
    # A entries before normalisation are N(0,1): after normalisation mean ~ 0, var ~ 1/m
    assert abs(A.mean().item()) < 3e-3 and abs(A.var().item() * m - 1) < 0.05
    for F in (Z, E):
        n = F.numel()
        frac = (F != 0).float().mean().item()
        assert abs(frac - p) < 4 * math.sqrt(p * (1 - p) / n)
        nz = F[F != 0]
        assert abs(nz.mean().item()) < 5 * sigma / math.sqrt(nz.numel())
        assert abs(nz.std().item() / sigma - 1) < 0.02
        assert abs((nz.abs() < 0.6745 * sigma).float().mean().item() - 0.5) < 0.01   # Gaussian median
    assert rel_l2(X, A @ Z + E) < 1e-5
    # shards reproduce the global stream
    part = dl.gen_syn_data(1000, m=m, d=d, p=p, sigma=sigma, seed=7, A=A, col_offset=3001)
    assert torch.equal(part.Z, Z[:, 3001:4001]) and torch.equal(part.E, E[:, 3001:4001])
    other = dl.gen_syn_data(64, m=m, d=d, p=p, sigma=sigma, seed=8, A=A)
    assert not torch.equal(other.Z, Z[:, :64])
    dense = dl.gen_syn_data(4096, m=m, d=d, seed=7, A=A, dense_noise_sigma=1.0 / math.sqrt(m))
    assert (dense.E != 0).float().mean().item() > 0.999
    assert abs(dense.E.std().item() * math.sqrt(m) - 1) < 0.02


def test_fused_objective_matches_script_side_loop():
    model, data = _model("scalar", 250, 500, 4096, 15)
    alpha = 0.001
    with torch.no_grad():
        Z, E, L, T = model(data.X)
        got = dl.l1l1_objective(Z, E, T, alpha)
        exp = torch.stack([alpha * Z[k].abs().sum() + (data.X - data.A @ Z[k]).abs().sum() for k in range(15)])
    assert rel_l2(got, exp) < 1e-4


@pytest.mark.parametrize("variant,B,precision", [("scalar", 4096, "tf32x3"), ("full", 1000, "tf32x3"), ("lasso", 516, "tf32"),
                                               ("ltheta", 2048, "tf32x3"), ("tied", 1001, "tf32x3"), ("scalar", 300, "fp32")])
def test_epilogue_fused_objective_matches_script_side_loop(variant, B, precision):
    """forward_objective (dladmm_problem.objective: accumulated inside the product epilogues) against the script-side
    loop of main_syn_l1l1_scalar.py:333-334, at tile-multiple and ragged batch sizes (B % 4 != 0 runs the FFMA kernels
    followed by the reduction pass), and the same numbers with last_only=True where no iterate is kept."""
    K, alpha = 6, 0.01
    model, data = _model(variant, 250, 500, B, K, precision=precision)
    obj, outs = model.forward_objective(data.X, alpha)
    Z = outs[0]
    exp = torch.stack([alpha * Z[k].abs().sum() + (data.X - data.A @ Z[k]).abs().sum() for k in range(K)])
    # single-pass tf32 carries the product's 2^-11 operand rounding into E_k - T_{k+1}; the other precisions are fp32-level
    assert rel_l2(obj, exp) < (2e-3 if precision == "tf32" else 1e-4)
    plain = model(data.X)
    assert torch.equal(plain[0][K - 1], Z[K - 1])                 # same iterates as the plain forward
    if precision != "fp32" and B % 4 == 0:
        obj2, last = model.forward_objective(data.X, alpha, last_only=True)
        assert rel_l2(obj2, obj) < 1e-6
        assert torch.equal(last[0][0], Z[K - 1])


def test_objective_improves_with_depth_under_km_parameters():
    m, d, B, K, alpha = 250, 500, 2048, 40, 0.01
    data = dl.gen_syn_data(B, m=m, d=d, seed=3)
    z = lambda r: torch.zeros(r, B, device="cuda")
    sd, ss1 = orc.km_state_dict("scalar", data.A.cpu(), K, alpha)
    model = dl.DLADMMNetScalar(m, 1, d, B, data.A, z(d), z(m), z(m), K)
    model.load_state_dict(sd)
    with torch.no_grad():
        Z, E, L, T = model(data.X)
        obj = dl.l1l1_objective(Z, E, T, alpha) / B
    assert obj[-1] < obj[10] < obj[0]


def test_wrong_dtype_and_shape_raise():
    g = Golden("scalar_small")
    model = build_model(g, "cuda")
    with pytest.raises(RuntimeError, match="float32"):
        model(g.X.double().cuda())
    with pytest.raises(RuntimeError, match="batch"):
        model(g.X[:, :5].contiguous().cuda())


def test_native_library_is_the_code_that_runs():
    """The loaded shared object is the in-tree libdladmm.so (no torch-eager fallback exists)."""
    g = Golden("scalar_small")
    model = build_model(g, "cuda")
    with torch.no_grad():
        model(g.X.cuda())
    maps = open("/proc/self/maps").read()
    assert "libdladmm.so" in maps
    caps = dl.query_device(0)
    assert caps["supported"] == 1 and caps["cc_major"] == 10 and caps["sm_count"] >= 100


@pytest.mark.parametrize("precision", ["tf32_bf16x2", "tf32x3", "fp32"])
def test_config5_shape_multiple_feature_tiles(precision):
    """BASELINE config 5 shape (m=1000, d=2000): several 256-row feature tiles per product, K-loop over 2000;
    forward against the fp32 oracle and training gradients against fp64 autograd, at a batch the CPU finishes fast."""
    torch.manual_seed(7)
    m, d, B, K = 1000, 2000, 256, 3
    A, X = syn(m, d, B, seed=77)
    Z0 = torch.zeros(d, B); E0 = torch.zeros(m, B); L0 = torch.zeros(m, B)
    model = dl.DLADMMNetScalar(m, 1, d, B, A, Z0, E0, L0, K, precision=precision)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    loss, outs = model.l1l1_loss(X.cuda(), 0.001, [0.5, 0.5, 1.0])
    loss.backward()
    Zo, Eo, Lo, To = orc.forward("scalar", sd, A, X, Z0, E0, L0, K)
    # (K = 2000: 500 .. 750 accumulating MMAs per element; 7e-6 per product once their round-toward-zero bias is compensated)
    tol = 2.5e-5 if precision != "fp32" else 1e-5
    for k in range(K):
        assert rel_l2(outs[0][k].cpu(), Zo[k], floor=1e-3) < tol, (k, rel_l2(outs[0][k].cpu(), Zo[k]))
        assert rel_l2(outs[2][k].cpu(), Lo[k], floor=1e-2) < tol
    d64 = lambda t: t.double()
    lref, gref = orc.autograd_grads("scalar", {k: d64(v) for k, v in sd.items()}, d64(A), d64(X), d64(Z0), d64(E0), d64(L0), K,
                                    lambda Z, E, L, T: orc.l1l1_loss(Z, E, L, T, d64(A), d64(X), alpha=0.001, decay=0.5))
    assert abs(loss.item() - lref.item()) < 1e-4 * abs(lref.item())
    for n, p in model.named_parameters():
        floor = 2e-3 if p.numel() == 1 else 1e-4 * max(1.0, float(gref[n].abs().max()) * p.numel() ** 0.5)
        # the L1 objective is non-smooth: sign(E_k - T_{k+1}) and the prox masks flip for entries within rounding of
        # zero, and with only 256 columns a handful of flips moves dW by ~0.5 % in 3xTF32
        assert rel_l2(p.grad.cpu(), gref[n], floor=floor) < (2e-2 if precision != "fp32" else 2e-3), n


@pytest.mark.parametrize("variant,B", [("scalar", 21504), ("full", 21444), ("lasso", 21504)])
def test_half_tile_tail_round_matches_ffma_path(variant, B):
    """168 batch tiles on 148 SMs: the last round of the persistent tcgen05 kernels is cut into 128-row half tiles
    (umma::plan_tiles).  Forward iterates, loss and every parameter gradient against the FFMA (fp32) kernels, which the
    parity tests pin to the oracle; `full` exercises the per-row parameter-gradient reductions, 21444 a ragged last tile."""
    m, d, K = 250, 500, 3
    w = [0.5, 0.5, 1.0]
    res = {}
    for precision in ("fp32", "tf32x3"):
        model, data = _model(variant, m, d, B, K, seed=5, precision=precision)
        loss, outs = model.l1l1_loss(data.X, 0.01, w)
        loss.backward()
        res[precision] = (loss.item(), [t.clone() for t in outs[0]], [t.clone() for t in outs[1]],
                          {n: p.grad.clone() for n, p in model.named_parameters()})
    l0, Z0_, E0_, g0 = res["fp32"]
    l1, Z1_, E1_, g1 = res["tf32x3"]
    assert abs(l0 - l1) < 1e-5 * abs(l0)
    for k in range(K):
        assert rel_l2(Z1_[k], Z0_[k]) < 2e-5 and rel_l2(E1_[k], E0_[k]) < 2e-5
    # over 21 K columns a handful of prox masks sit within rounding of their threshold and differ between the two
    # arithmetics; that moves a gradient by ~1e-3 relative (a half-tile bug would be O(1))
    for n in g0:
        assert rel_l2(g1[n], g0[n], floor=1e-5) < 5e-3, n


def test_unseen_distribution_generators():
    """gen_syn_unseen_data_cosine.py / _logistic.py: the non-zero amplitudes are cos(g) / 1/(1+exp(g)) of the SAME
    Gaussian draws and Bernoulli supports as the plain generator with that seed; _Acols.py: c columns of A replaced."""
    m, d, B, p, sigma = 250, 500, 4096, 0.1, 1.5
    base = dl.gen_syn_data(B, m=m, d=d, p=p, sigma=sigma, mu=0.3, seed=11)
    for name, fn in (("cosine", torch.cos), ("logistic", lambda g: 1.0 / (1.0 + torch.exp(g)))):
        data = dl.gen_syn_data(B, m=m, d=d, p=p, sigma=sigma, mu=0.3, seed=11, A=base.A, amplitude=name)
        for F, G in ((data.Z, base.Z), (data.E, base.E)):
            nz = G != 0
            assert torch.equal(F != 0, nz) or ((F != 0) ^ nz).float().mean().item() < 1e-5   # cos(g) == 0 has measure zero
            assert (F[nz] - fn(G[nz])).abs().max().item() < 2e-6
        assert rel_l2(data.X, data.A @ data.Z + data.E) < 1e-5
    A2, cols = dl.replace_A_columns(base.A, 37, seed=5)
    assert len(set(cols.tolist())) == 37
    same = torch.ones(d, dtype=torch.bool); same[cols] = False
    assert torch.equal(A2[:, same], base.A[:, same]) and not torch.equal(A2[:, cols], base.A[:, cols])
    assert (A2.pow(2).sum(dim=0).sqrt() - 1).abs().max().item() < 1e-5
    with pytest.raises(ValueError):
        dl.gen_syn_data(8, amplitude="sine")


def test_host_feed_delivers_every_batch_in_order_while_overlapping_uploads():
    """HostFeed: batches come out in order, bit-identical to the pinned host tensors, buffers are recycled only after the
    consumer's work on them was enqueued, and the forward over a fed batch equals the forward over a resident copy."""
    model, data = _model("scalar", 60, 100, 256, 3)
    g = torch.Generator().manual_seed(0)
    host = [(data.X.cpu() * (1.0 + 0.1 * i) + 0.01 * torch.randn(60, 256, generator=g)).pin_memory() for i in range(5)]
    feed = dl.HostFeed(iter(host), "cuda")
    seen = 0
    for i, x in enumerate(feed):
        assert x.is_cuda and torch.equal(x.cpu(), host[i])
        with torch.no_grad():
            got = model(x)[0][-1].clone()
            ref = model(host[i].cuda())[0][-1]
        assert torch.equal(got, ref)
        seen += 1
    assert seen == 5 and feed.bytes_copied == 5 * 60 * 256 * 4
    with pytest.raises(RuntimeError):
        dl.HostFeed(iter(host), "cpu")
    with pytest.raises(ValueError):
        dl.HostFeed(iter(host), "cuda", depth=1)


@pytest.mark.parametrize("variant", ["lasso", "tied"])
def test_config3_k15_forward_backward_both_arithmetics_agree(variant):
    """BASELINE configs[2] (main_syn_lasso_scalar.py, tied/untied layers, K=15 forward + backward) at m=250, d=500: loss and
    every parameter gradient of the tcgen05 path against the FFMA (fp32) path, which the fixtures pin to the reference; the
    LASSO training objective of the reference driver is alpha*|Z_k|_1 + 0.5*|X - A Z_k|_2^2 per layer (squared residual)."""
    m, d, B, K = 250, 500, 4096, 15
    res = {}
    for precision in ("fp32", "tf32x3"):
        model, data = _model(variant, m, d, B, K, seed=9, precision=precision)
        if variant == "lasso":
            Z, E, L, T = model(data.X)
            loss = sum((0.6 ** (K - 1 - k)) * (0.01 * Z[k].abs().sum() + 0.5 * (E[k] - T[k + 1]).pow(2).sum()) for k in range(K)) / B
        else:
            loss, _ = model.l1l1_loss(data.X, 0.01, [0.6 ** (K - 1 - k) for k in range(K)])
        loss.backward()
        res[precision] = (loss.item(), {n: p.grad.clone() for n, p in model.named_parameters()})
    (l0, g0), (l1, g1) = res["fp32"], res["tf32x3"]
    assert abs(l0 - l1) < 2e-5 * abs(l0)
    # some scalar gradients are sums that cancel to ~1e-5 of the others: errors are judged against the largest scalar gradient too
    G = max(v.norm().item() for n, v in g0.items() if v.numel() == 1)
    for n in g0:
        assert torch.isfinite(g1[n]).all(), n
        assert rel_l2(g1[n], g0[n], floor=1e-2 * G if g0[n].numel() == 1 else 1e-5) < 1e-2, n


def test_sync_gradients_is_inert_without_a_process_group():
    """model.sync_gradients() allreduces inside backward only when torch.distributed is initialised with more than one
    rank (tools/check_grad_sync.py covers world size 2 under torchrun); alone it must not change anything."""
    model, data = _model("scalar", 60, 100, 64, 3)
    loss, _ = model.l1l1_loss(data.X, 0.01)
    loss.backward()
    ref = [p.grad.clone() for p in model.parameters()]
    model.zero_grad(set_to_none=True)
    model.sync_gradients(True)
    loss, _ = model.l1l1_loss(data.X, 0.01)
    loss.backward()
    assert all(rel_l2(p.grad, r, floor=1e-7) < 1e-5 for p, r in zip(model.parameters(), ref))    # dW atomics: order only


def test_mask_sign_bits_and_layer_table_cache():
    """ABI v6: bits 2 / 3 of the maskZ bytes are [Z_k > 0] / [Z_k < 0] (the fused-loss backward reads sign(Z_k) from them), next to
    the two prox indicators in bits 0 / 1 -- also with negative thresholds, where bits 0 and 1 can both be set.  And the host-side
    layer table is reused between calls only while every parameter sits at the same address."""
    from dladmm_b200.function import run_forward
    torch.manual_seed(5)
    model, data = _model("full", 96, 200, 516, 3, seed=2, precision="tf32x3")
    with torch.no_grad():
        model.active_para[1].mul_(-0.5)                                          # negative thresholds in layer 1
    spec, params = model._spec_and_params()
    params = [p.detach() for p in params]
    Z, E, L, T, maskZ, maskE = run_forward(spec, model.A, data.X, model.Z0, model.E0, model.L0, params, want_masks=True, extras={})
    assert torch.equal((maskZ & 4) != 0, Z > 0) and torch.equal((maskZ & 8) != 0, Z < 0)
    assert ((maskZ[1] & 3) == 3).any()                                           # both indicators set somewhere under theta < 0
    assert int(maskZ.max()) < 16 and int(maskE.max()) < 4
    table = spec._layer_cache[1]
    run_forward(spec, model.A, data.X, model.Z0, model.E0, model.L0, params, want_masks=False)
    assert spec._layer_cache[1] is table                                         # same addresses: reused
    moved = list(params); moved[0] = params[0].clone()
    Z2, _, _, _, _, _ = run_forward(spec, model.A, data.X, model.Z0, model.E0, model.L0, moved, want_masks=False)
    assert spec._layer_cache[1] is not table and torch.equal(Z2, Z)              # a parameter moved: rebuilt, same result


def test_accumulator_compensation_removes_the_round_toward_zero_bias():
    """The tensor core accumulates with round-toward-zero: one product shrinks by ~1.8e-8 per MMA (mean SIGNED error vs fp64).  The
    epilogues' first-order compensation (umma_gemm.cuh, ACC_RZ_BIAS_PER_MMA) must remove that mean in both fp32-class modes and
    bring the product to <= 2.5e-6 at K = 500 (measured 1.6e-6 / 1.8e-6; uncompensated 2.8e-6 / 3.8e-6 with a mean of -2.3e-6 /
    -3.3e-6: profiles/r02_precision_table*.md)."""
    torch.manual_seed(3)
    m, d, B = 250, 500, 4096
    A = torch.randn(m, d, device="cuda"); A = A / A.norm(dim=0, keepdim=True)
    Zd = torch.randn(d, B, device="cuda")
    z = torch.zeros(m, B, device="cuda")
    want = A.double() @ Zd.double()
    for mode in ("tf32_bf16x2", "tf32x3"):
        got = dl.DLADMMNetScalar(m, 1, d, B, A, Zd, z, z, 1, precision=mode)._t0(z).double()
        err = ((got - want).norm() / want.norm()).item()
        bias = (((got - want) * want.sign()).sum() / want.abs().sum()).item()
        assert err < 2.5e-6 and abs(bias) < 4e-7, (mode, err, bias)


@pytest.mark.parametrize("m,d,B", [(96, 200, 1000), (130, 260, 516), (300, 700, 2048), (64, 128, 40)])
@pytest.mark.parametrize("precision", ["tf32_bf16x2", "tf32x3"])
def test_weight_gradient_tiles_at_odd_shapes(m, d, B, precision):
    """The dW kernel works on 256 x 256 tiles of (d x m): one or two M = 128 halves per CTA, several tiles in both directions,
    ragged last tiles, batch slices that do not divide evenly.  Every parameter gradient (dW above all) of both fp32-class modes
    against the FFMA (fp32) path, which the fixtures pin to the reference."""
    K = 2
    res = {}
    for prec in ("fp32", precision):
        model, data = _model("scalar", m, d, B, K, seed=4, precision=prec)
        loss, _ = model.l1l1_loss(data.X, 0.01, [0.5, 1.0])
        loss.backward()
        res[prec] = (loss.item(), {n: p.grad.clone() for n, p in model.named_parameters()})
    (l0, g0), (l1, g1) = res["fp32"], res[precision]
    assert abs(l0 - l1) < 2e-5 * abs(l0)
    G = max(v.norm().item() for n, v in g0.items() if v.numel() == 1)
    for n in g0:
        assert torch.isfinite(g1[n]).all(), n
        assert rel_l2(g1[n], g0[n], floor=1e-2 * G if g0[n].numel() == 1 else 1e-5) < 1e-2, (n, rel_l2(g1[n], g0[n]))
