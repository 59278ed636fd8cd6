"""autograd.Function over the C ABI: the K-layer unrolled forward and its hand-written backward.

Replaces the ATen op chain + autograd graph the reference builds per call
(main_syn_l1l1_scalar.py:80-127 and `total_loss.backward()` at :301) with two library calls.
PyTorch here is plumbing only: it owns device memory and the stream.
"""
import ctypes as C

import torch

from . import _lib

# slot order of per-layer broadcast parameters in dladmm_layer
_SLOTS = ("beta1", "beta2", "beta3", "ss1", "ss2", "ss2_2", "theta1", "theta2")


class LayerSpec(object):
    """Static description of one call: family, sizes, and where each reference parameter sits in the
    flat tensor list handed to autograd.  `slots[k][slot]` is an index into that list or None."""

    def __init__(self, family, m, d, K, precision, slots, weights, fixed=None):
        self.family = family
        self.m, self.d, self.K = m, d, K
        self.precision = precision
        self.slots = slots          # list (K) of dict slot -> param index
        self.weights = weights      # list (K) of param index of fc weight
        self.fixed = fixed or {}    # slot -> non-learnable 0-dim tensor (lena thresholds)
        self.grad_sync = None       # (process group or None,) : sum-allreduce the step's gradient buffer inside backward


def _require_cuda_f32(name, t):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: d-ladmm_b200 has no CPU path" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32, got %s" % (name, t.dtype))


def _bparam(t, grad):
    bp = _lib.BParam()
    if t is None:
        return bp
    bp.ptr = t.data_ptr()
    bp.grad = grad.data_ptr() if grad is not None else None
    if t.numel() == 1:
        bp.row_stride, bp.col_period = 0, 0
    elif t.dim() == 2 and t.shape[1] == 1:
        bp.row_stride, bp.col_period = 1, 0
    elif t.dim() == 2:
        bp.row_stride, bp.col_period = t.shape[1], t.shape[1]
    else:
        raise RuntimeError("unsupported parameter shape %s" % (tuple(t.shape),))
    return bp


def _build_layers(spec, params, grads):
    arr = (_lib.Layer * spec.K)()
    for k in range(spec.K):
        lay = arr[k]
        for slot in _SLOTS:
            idx = spec.slots[k].get(slot)
            if idx is not None:
                t = params[idx]
                g = grads[idx] if grads is not None else None
            else:
                t, g = spec.fixed.get(slot), None
            setattr(lay, slot, _bparam(t, g))
        wi = spec.weights[k]
        lay.W = params[wi].data_ptr()
        lay.gW = grads[wi].data_ptr() if (grads is not None and grads[wi] is not None) else None
    return arr


def _problem(spec, A, X, Z0, E0, L0, layers, Z, E, L, T, maskZ, maskE, last_only, for_backward, T_init=None, Vsave=None,
             objective=None, objective_alpha=0.0, objective_kind=0, start_half=False, stop_half=False, metrics=None):
    lib = _lib.load()
    p = _lib.Problem()
    p.abi_version = _lib.ABI_VERSION
    p.family = spec.family
    p.precision = spec.precision
    p.m, p.d, p.K = spec.m, spec.d, spec.K
    p.B = X.shape[1]
    p.last_only = 1 if last_only else 0
    p.A, p.X, p.Z0, p.E0, p.L0 = A.data_ptr(), X.data_ptr(), Z0.data_ptr(), E0.data_ptr(), L0.data_ptr()
    p.layers = C.cast(layers, C.POINTER(_lib.Layer))
    p.Z, p.E, p.L, p.T = Z.data_ptr(), E.data_ptr(), L.data_ptr(), T.data_ptr()
    p.maskZ = maskZ.data_ptr() if maskZ is not None else None
    p.maskE = maskE.data_ptr() if maskE is not None else None
    p.T_init = T_init.data_ptr() if T_init is not None else None
    p.Vsave = Vsave.data_ptr() if Vsave is not None else None
    p.objective = objective.data_ptr() if objective is not None else None
    p.objective_alpha = float(objective_alpha)
    p.objective_kind = int(objective_kind)
    p.start_half, p.stop_half = (1 if start_half else 0), (1 if stop_half else 0)
    p.metrics = C.addressof(metrics) if metrics is not None else None
    nbytes = lib.dladmm_workspace_bytes(C.byref(p), 1 if for_backward else 0)
    ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=X.device)
    p.workspace = ws.data_ptr()
    p.workspace_bytes = nbytes
    return p, ws


def _check_inputs(spec, A, X, Z0, E0, L0, params):
    for name, t in (("A", A), ("x", X), ("Z0", Z0), ("E0", E0), ("L0", L0)):
        _require_cuda_f32(name, t)
    for i, t in enumerate(params):
        _require_cuda_f32("parameter %d" % i, t)
    m, d = spec.m, spec.d
    if tuple(A.shape) != (m, d):
        raise RuntimeError("A must be (%d,%d), got %s" % (m, d, tuple(A.shape)))
    if X.dim() != 2 or X.shape[0] != m:
        raise RuntimeError("x must be (m=%d, B), got %s" % (m, tuple(X.shape)))
    B = X.shape[1]
    if tuple(Z0.shape) != (d, B) or tuple(E0.shape) != (m, B) or tuple(L0.shape) != (m, B):
        raise RuntimeError("batch of x (%d) must match Z0/E0/L0 (%s, %s, %s)" %
                           (B, tuple(Z0.shape), tuple(E0.shape), tuple(L0.shape)))


def _pad_cols(t, Bp):
    """(.., B) -> contiguous (.., Bp) with zero columns appended (no copy when nothing is to be padded)."""
    if t.shape[-1] == Bp:
        return t.contiguous()
    out = t.new_zeros(t.shape[:-1] + (Bp,))
    out[..., :t.shape[-1]] = t
    return out


def padded_batch(spec, B):
    """The tensor-core kernels move every (rows x B) array with TMA, which needs a 16-byte multiple pitch: batches with
    B % 4 != 0 (the reference scripts' default -bs 25, main_syn_l1l1_scalar.py:20) are stored with a pitch of B rounded up
    to 4 -- zero observation columns, which stay exactly zero through every layer and contribute nothing to any
    parameter gradient -- and the caller gets views narrowed to B columns."""
    if spec.precision == _lib.PRECISIONS["fp32"] or B % 4 == 0:
        return B
    return (B + 3) // 4 * 4


def _metrics_desc(metrics, m, d, B_user, Bp, K, dev):
    """dict(want=[names of _lib.METRICS], Z_label=, E_label=, X_clean=, dual_alpha=) -> (dladmm_metrics, out (K, MET_COUNT), keepalive)"""
    mt = _lib.Metrics()
    keep = []
    for name in metrics.get("want", ()):
        if name not in _lib.METRICS:
            raise ValueError("unknown metric %r (known: %s)" % (name, ", ".join(_lib.METRICS)))
        mt.want |= 1 << _lib.METRICS.index(name)
    mt.dual_alpha = float(metrics.get("dual_alpha", 0.0))
    for field, rows in (("Z_label", d), ("E_label", m), ("X_clean", m)):
        t = metrics.get(field)
        if t is None:
            continue
        _require_cuda_f32(field, t)
        if tuple(t.shape) != (rows, B_user):
            raise RuntimeError("%s must be (%d, %d), got %s" % (field, rows, B_user, tuple(t.shape)))
        t = _pad_cols(t, Bp)
        keep.append(t)
        setattr(mt, field, t.data_ptr())
    out = torch.zeros((K, _lib.MET_COUNT), dtype=torch.float32, device=dev)
    mt.out = out.data_ptr()
    return mt, out, keep


def run_forward(spec, A, X, Z0, E0, L0, params, want_masks, last_only=False, T_init=None, objective_alpha=None, extras=None,
                objective_kind=0, start_half=False, stop_half=False, metrics=None):
    """Launch the K-layer forward.  Returns stacked (Z, E, L, T, maskZ, maskE).  `T_init` (m,B): T_0 supplied by the
    caller instead of A Z0 + E0 - X (single-layer steps from an arbitrary state).  `objective_alpha`: also compute the
    per-layer L1-L1 objective inside the product epilogues; `extras` (a dict) receives "objective" (K floats) and, in
    training mode (`want_masks`), "Vsave" (the kept W V operands the backward reuses) and "padded" (the inputs and
    outputs at the padded pitch, see padded_batch; the returned Z, E, L, T are then views narrowed to B columns)."""
    lib = _lib.load()
    # The layer table (K x 8 broadcast parameters + weight pointers as ctypes structs) only depends on where the parameters live:
    # it is kept on the spec and reused while every parameter still sits at the same address (in-place optimizer steps keep it;
    # at the reference scripts' batch sizes building it was half of the 0.3 ms host time of a call).
    ptrs = tuple([t.data_ptr() for t in params])
    hit = getattr(spec, "_layer_cache", None)
    if hit is not None and hit[0] == ptrs:
        layers = hit[1]
        _check_inputs(spec, A, X, Z0, E0, L0, ())
    else:
        _check_inputs(spec, A, X, Z0, E0, L0, params)
        layers = None
    B_user = X.shape[1]
    Bp = padded_batch(spec, B_user)
    X, Z0, E0, L0 = _pad_cols(X, Bp), _pad_cols(Z0, Bp), _pad_cols(E0, Bp), _pad_cols(L0, Bp)
    A = A.contiguous()
    if layers is None:
        contig = all(t.is_contiguous() for t in params)
        params = [t.contiguous() for t in params]
        layers = _build_layers(spec, params, None)
        spec._layer_cache = (ptrs, layers) if contig else None      # (a non-contiguous parameter is copied per call: no reuse)
    m, d, K, B = spec.m, spec.d, spec.K, Bp
    dev = X.device
    depth = 2 if last_only else K
    Z = torch.empty((depth, d, B), dtype=torch.float32, device=dev)
    E = torch.empty((depth, m, B), dtype=torch.float32, device=dev)
    L = torch.empty((depth, m, B), dtype=torch.float32, device=dev)
    T = torch.empty((2 if last_only else K + 1, m, B), dtype=torch.float32, device=dev)
    maskZ = maskE = None
    if want_masks:
        maskZ = torch.empty((K, d, B), dtype=torch.uint8, device=dev)
        if spec.family != _lib.FAMILY_C:
            maskE = torch.empty((K, m, B), dtype=torch.uint8, device=dev)
    if T_init is not None:
        _require_cuda_f32("T_init", T_init)
        if tuple(T_init.shape) != (m, B_user):
            raise RuntimeError("T_init must be (m, B)")
        T_init = _pad_cols(T_init, Bp)
    # the tensor-core path keeps every layer's W V operand for the backward (the FFMA path recomputes it in its loader)
    keep_v = want_masks and extras is not None and spec.precision != _lib.PRECISIONS["fp32"] and B % 4 == 0
    Vsave = torch.empty((K, m, B), dtype=torch.float32, device=dev) if keep_v else None
    obj = torch.empty(K, dtype=torch.float32, device=dev) if objective_alpha is not None else None
    mt = mt_out = mt_keep = None
    if metrics is not None:
        mt, mt_out, mt_keep = _metrics_desc(metrics, m, d, B_user, Bp, K, dev)
    with torch.cuda.device(dev):
        p, ws = _problem(spec, A, X, Z0, E0, L0, layers, Z, E, L, T, maskZ, maskE, last_only, False, T_init, Vsave, obj,
                         objective_alpha or 0.0, objective_kind, start_half, stop_half, mt)
        _lib.check(lib.dladmm_forward(C.byref(p), torch.cuda.current_stream(dev).cuda_stream))
        ws.record_stream(torch.cuda.current_stream(dev))
    if extras is not None:
        extras["Vsave"], extras["objective"] = Vsave, obj
        extras["metrics"] = mt_out
        extras["padded"] = (X, Z0, E0, L0, Z, E, L, T)
    if Bp != B_user:
        Z, E, L, T = Z[..., :B_user], E[..., :B_user], L[..., :B_user], T[..., :B_user]
    return Z, E, L, T, maskZ, maskE


def _flat_zero_grads(params, needs, order=None):
    """Zero-initialised gradient buffers for the parameters that need one, carved out of ONE flat allocation (one fill
    instead of one per parameter -- 106 launches at K=15; 16-byte aligned pieces).  autograd copies them into `.grad`.
    `order`: the sequence in which the parameters are laid out (default: list order); the bucketed gradient sync puts
    the weights first, in the order in which the backward finishes them."""
    sizes = [(-(-t.numel() // 4) * 4) if needs[i] else 0 for i, t in enumerate(params)]
    total = sum(sizes)
    if total == 0:
        return [None] * len(params), None
    flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
    grads, off = [None] * len(params), 0
    for i in (range(len(params)) if order is None else order):
        if not needs[i]:
            continue
        t = params[i]
        grads[i] = flat[off:off + t.numel()].view(t.shape)
        off += sizes[i]
    return grads, flat


# Opt-in (model.sync_gradients(bucket_mb=...)): weight gradients are allreduced in buckets of at least this many bytes while the
# layers below are still running.  The default stays ONE collective after the backward: measured at the large-scale shape (320 MB
# of weight gradients, 2 and 8 B200s, profiles/r02_c5_sync_sweep.md) the overlapped collective hides 0.6-0.9 ms of allreduce and
# costs the power-capped product kernels the same 0.6-1.1 ms.
GRAD_BUCKET_BYTES = None


class SyncPlan(object):
    """Layout of the step's flat gradient buffer for the bucketed in-backward allreduce: `order` = parameter
    indices, weights first by descending layer of completion (the backward runs K-1 ... 0; a weight shared by several
    layers is final after the lowest of them), then everything else; `buckets` = [(layer whose event releases the bucket,
    first element, one past the last element)] over the weight part; `rest` = first element of the remainder, which is
    reduced after the last kernel of the backward."""

    def __init__(self, order, buckets, rest):
        self.order, self.buckets, self.rest = order, buckets, rest


def plan_gradient_buckets(spec, params, needs, bucket_bytes=None):
    """None when one collective after the backward is the right schedule (weight gradients below two buckets)."""
    bucket_bytes = GRAD_BUCKET_BYTES if bucket_bytes is None else int(bucket_bytes)
    if bucket_bytes is None:
        return None
    ready = {}
    for k in range(spec.K):
        wi = spec.weights[k]
        if needs[wi]:
            ready[wi] = min(ready.get(wi, k), k)
    pad = lambda n: -(-n // 4) * 4
    wbytes = 4 * sum(pad(params[wi].numel()) for wi in ready)
    if not ready or wbytes < 2 * bucket_bytes:
        return None
    worder = sorted(ready, key=lambda wi: (-ready[wi], wi))
    buckets, lo, off = [], 0, 0
    for n, wi in enumerate(worder):
        off += pad(params[wi].numel())
        left = 4 * sum(pad(params[j].numel()) for j in worder[n + 1:])
        # close a bucket once it is large enough, unless what is left would make an undersized last one
        if 4 * (off - lo) >= bucket_bytes and (left == 0 or left >= bucket_bytes):
            buckets.append((ready[wi], lo, off))
            lo = off
    if lo < off:
        buckets.append((ready[worder[-1]], lo, off))
    inw = set(worder)
    order = worder + [i for i in range(len(params)) if i not in inw]
    return SyncPlan(order, buckets, off)


def _sync_world(spec):
    if spec.grad_sync is None:
        return 1
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(spec.grad_sync[0])
    return 1


def _sync_resources(spec, dev, nbuckets):
    """Side stream and one event per bucket, kept on the spec (events are re-recorded every step)."""
    res = getattr(spec, "_sync_res", None)
    if res is None or res[0] != dev or len(res[2]) < nbuckets:
        side = torch.cuda.Stream(device=dev)
        events = [torch.cuda.Event() for _ in range(nbuckets)]
        for ev in events:
            ev.record(torch.cuda.current_stream(dev))       # creates the CUDA event behind the lazy torch object
        res = spec._sync_res = (dev, side, events)
    return res[1], res[2][:nbuckets]


def _sync_prepare(spec, params, needs, cot, dev):
    """Before the backward call: the bucket plan (or None) and the per-layer event table handed to the library."""
    if _sync_world(spec) <= 1:
        return None
    plan = plan_gradient_buckets(spec, params, needs, getattr(spec, "grad_bucket_bytes", None))
    if plan is None:
        return None
    side, events = _sync_resources(spec, dev, len(plan.buckets))
    table = (C.c_void_p * spec.K)()
    for (layer, _, _), ev in zip(plan.buckets, events):
        table[layer] = ev.cuda_event
    cot.layer_events = C.cast(table, C.POINTER(C.c_void_p))
    plan.table, plan.side, plan.events = table, side, events
    return plan


def _sync_gradients(spec, flat, plan=None):
    """Data-parallel training over column shards: the parameter gradients of this rank's columns are one
    contiguous buffer, sum-allreduced in place (NCCL, stream-ordered after the backward kernels) before autograd hands
    them to `.grad` -- no staging copy.  One collective per backward, or with a bucket plan one per bucket of finished
    weight gradients, issued from a side stream as soon as the library's per-layer event fires (the layers below keep
    running on the caller's stream), plus one for the remainder after the last kernel."""
    if flat is None or _sync_world(spec) <= 1:
        return
    import torch.distributed as dist
    group = spec.grad_sync[0]
    if plan is None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return
    works = []
    for (_, lo, hi), ev in zip(plan.buckets, plan.events):
        plan.side.wait_event(ev)
        with torch.cuda.stream(plan.side):
            works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True))
    if plan.rest < flat.numel():
        dist.all_reduce(flat[plan.rest:], op=dist.ReduceOp.SUM, group=group)
    for w in works:
        w.wait()                                   # the caller's stream waits for the collectives; the host does not
    flat.record_stream(plan.side)


class UnrolledLADMM(torch.autograd.Function):
    """forward(spec, A, X, Z0, E0, L0, *params) -> (Z, E, L, T) stacks; gradients flow to `params` only
    (A, Z0, E0, L0 are plain tensors in the reference, main_syn_l1l1_scalar.py:40-44)."""

    @staticmethod
    def forward(ctx, spec, A, X, Z0, E0, L0, *params):
        extras = {}
        Z, E, L, T, maskZ, maskE = run_forward(spec, A, X, Z0, E0, L0, list(params), want_masks=True, extras=extras)
        ctx.spec = spec
        ctx.nparams = len(params)
        ctx.has_maskE = maskE is not None
        ctx.B_user = X.shape[1]
        Xp, Z0p, E0p, L0p, Zp, Ep, Lp, Tp = extras["padded"]       # the same tensors unless B % 4 != 0 (padded pitch)
        saved = [A, Xp, Z0p, E0p, L0p, Zp, Ep, Lp, Tp, maskZ, extras["Vsave"]] + ([maskE] if maskE is not None else []) + list(params)
        ctx.save_for_backward(*saved)
        ctx.set_materialize_grads(False)
        return Z, E, L, T

    @staticmethod
    def backward(ctx, gZ, gE, gL, gT):
        lib = _lib.load()
        spec = ctx.spec
        saved = ctx.saved_tensors
        A, X, Z0, E0, L0, Z, E, L, T, maskZ, Vsave = saved[:11]
        off = 11
        maskE = None
        if ctx.has_maskE:
            maskE = saved[11]
            off = 12
        params = [t.contiguous() for t in saved[off:]]
        needs = ctx.needs_input_grad[6:]
        cot = _lib.Cotangents()
        plan = _sync_prepare(spec, params, needs, cot, X.device)
        grads, flat = _flat_zero_grads(params, needs, plan.order if plan else None)
        keep = []
        for name, g in (("gZ", gZ), ("gE", gE), ("gL", gL), ("gT", gT)):
            if g is not None:
                _require_cuda_f32(name, g)
                g = _pad_cols(g, X.shape[1])
                keep.append(g)
                setattr(cot, name, g.data_ptr())
        layers = _build_layers(spec, params, grads)
        dev = X.device
        with torch.cuda.device(dev):
            p, ws = _problem(spec, A, X, Z0, E0, L0, layers, Z, E, L, T, maskZ, maskE, False, True, Vsave=Vsave)
            _lib.check(lib.dladmm_backward(C.byref(p), C.byref(cot), torch.cuda.current_stream(dev).cuda_stream))
            ws.record_stream(torch.cuda.current_stream(dev))
            _sync_gradients(spec, flat, plan)
        return (None, None, None, None, None, None) + tuple(grads)


_WEIGHT_CACHE = {}


def _layer_weight_tensor(weights, device):
    """Per-layer loss weights on the device, cached: building them per step would be a synchronous pageable H2D copy
    that makes the host wait for the whole forward before it can enqueue the backward."""
    key = (tuple(float(v) for v in weights), str(device))
    t = _WEIGHT_CACHE.get(key)
    if t is None:
        if len(_WEIGHT_CACHE) > 64:
            _WEIGHT_CACHE.clear()
        t = _WEIGHT_CACHE[key] = torch.tensor(key[0], dtype=torch.float32, device=device)
    return t


class UnrolledLADMML1L1(torch.autograd.Function):
    """forward(spec, kind, alpha, weights, A, X, Z0, E0, L0, *params) -> (loss, Z, E, L, T) with

        loss = (1/B) * sum_k weights[k] * sum_b ( alpha*||Z_k[:,b]||_1 + r(x[:,b] - A Z_k[:,b]) )

    kind 1: r = ||.||_1, the reference's L1-L1 training objective (main_syn_l1l1_scalar.py:289-299);
    kind 2: r = 0.5*||.||_2^2, the LASSO training objective (main_syn_lasso_scalar.py:276-281).  The loss cotangents are generated
    inside the backward kernels (dladmm_cotangents.loss_kind = 1): no per-layer A@Z products, no (K,d,B)/(K,m,B)
    gradient stacks, no elementwise autograd graph.  The returned iterates are not differentiable."""

    @staticmethod
    def forward(ctx, spec, kind, alpha, weights, A, X, Z0, E0, L0, *params):
        extras = {}
        Z, E, L, T, maskZ, maskE = run_forward(spec, A, X, Z0, E0, L0, list(params), want_masks=True,
                                               objective_alpha=float(alpha), extras=extras, objective_kind=int(kind))
        ctx.kind = int(kind)
        B = X.shape[1]
        obj = extras["objective"]
        w = _layer_weight_tensor(weights, X.device)
        loss = (obj * w).sum() / float(max(B, 1))
        ctx.spec, ctx.alpha, ctx.weights, ctx.B = spec, float(alpha), [float(v) for v in weights], B
        ctx.has_maskE = maskE is not None
        saved = [A, X, Z0, E0, L0, Z, E, L, T, maskZ, extras["Vsave"]] + ([maskE] if maskE is not None else []) + list(params)
        ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable(Z, E, L, T)
        # without this the engine hands backward zero-filled "gradients" of the four non-differentiable iterate stacks: 5 GB of
        # fills per C1 step (ncu launch list: four FillFunctor launches, 0.65 ms, in front of every backward)
        ctx.set_materialize_grads(False)
        return loss, Z, E, L, T

    @staticmethod
    def backward(ctx, gloss, *unused):
        lib = _lib.load()
        spec = ctx.spec
        if gloss is None:
            return (None,) * (9 + len(ctx.needs_input_grad[9:]))
        saved = ctx.saved_tensors
        A, X, Z0, E0, L0, Z, E, L, T, maskZ, Vsave = saved[:11]
        off = 11
        maskE = None
        if ctx.has_maskE:
            maskE = saved[11]
            off = 12
        params = [t.contiguous() for t in saved[off:]]
        needs = ctx.needs_input_grad[9:]
        cot = _lib.Cotangents()
        plan = _sync_prepare(spec, params, needs, cot, X.device)
        grads, flat = _flat_zero_grads(params, needs, plan.order if plan else None)
        scale = (gloss.detach().to(torch.float32) / float(max(ctx.B, 1))).reshape(1).contiguous()
        cot.loss_kind = ctx.kind
        cot.loss_alpha = ctx.alpha
        warr = (C.c_float * spec.K)(*ctx.weights)
        cot.loss_layer_weight = C.cast(warr, C.POINTER(C.c_float))
        cot.loss_scale = scale.data_ptr()
        layers = _build_layers(spec, params, grads)
        dev = X.device
        with torch.cuda.device(dev):
            p, ws = _problem(spec, A, X.contiguous(), Z0, E0, L0, layers, Z, E, L, T, maskZ, maskE, False, True, Vsave=Vsave)
            _lib.check(lib.dladmm_backward(C.byref(p), C.byref(cot), torch.cuda.current_stream(dev).cuda_stream))
            ws.record_stream(torch.cuda.current_stream(dev))
            scale.record_stream(torch.cuda.current_stream(dev))
            _sync_gradients(spec, flat, plan)
        return (None,) * 9 + tuple(grads)

