"""Host-side mirror of the reference's ``DLADMMNet`` nn.Module for every in-scope variant.

Same constructor arguments, parameter names, shapes and registration order (so ``state_dict`` /
``load_state_dict`` interchange with the reference), same ``forward(x)`` return structure, same
``name()``.  The arithmetic runs in libdladmm.so (hand-written sm_100a CUDA); there is no CPU or
PyTorch-eager fallback -- calling ``forward`` without the library or without a CUDA tensor raises.

    variant   reference class                                   forward returns
    scalar    main_syn_l1l1_scalar.py:34-131                    Z, E, L, T
    full      main_syn_l1l1_full.py:16-110                      Z, E, L
    tied      main_syn_l1l1_scalar_tied.py:34-134               Z, E, L, T
    lasso     main_syn_lasso_scalar.py:17-118                   Z, E, L, T
    lena      main_lena.py:16-102                               Z, E, L
    ltheta    main_syn_l1l1_ltheta.py:16-108                    Z, E, L
    newS        main_syn_scalar_newS_layerwise.py:34-116        forward(x, K) -> Z, E, L   (E -> L -> Z ordering)
    tied_newS   main_syn_scalar_tied_newS_layerwise.py:35-117   forward(x, K) -> Z, E, L   (one fc, ss1 per layer)
    ptied_newS  main_syn_scalar_ptied_newS_layerwise.py:37-124  forward(x, K) -> Z, E, L   (fc[k // interval])
"""
import ctypes as C
import os
from math import sqrt

import torch
import torch.nn as nn

from . import _lib
from .function import LayerSpec, UnrolledLADMM, UnrolledLADMML1L1, run_forward, padded_batch, _pad_cols
from .mu_updater import mu_updater_dict, METHOD_ID

_FAMILY = {"lena": _lib.FAMILY_A, "ltheta": _lib.FAMILY_A, "scalar": _lib.FAMILY_B, "full": _lib.FAMILY_B,
           "tied": _lib.FAMILY_B, "lasso": _lib.FAMILY_C, "newS": _lib.FAMILY_B, "tied_newS": _lib.FAMILY_B,
           "ptied_newS": _lib.FAMILY_B}
_RETURNS_T = {"lena": False, "ltheta": False, "scalar": True, "full": False, "tied": True, "lasso": True,
              "newS": False, "tied_newS": False, "ptied_newS": False}
_NAME = {"scalar": "DLADMMNet_scalar", "tied": "DLADMMNet_scalar_tied", "newS": "DLADMMNet_scalar_newS_layerwise",
         "tied_newS": "DLADMMNet_scalar_tied_newS_layerwise"}
# which fc a layer uses: one per layer, one for all layers, or one per `interval` consecutive layers
_TIE = {"tied": "all", "tied_newS": "all", "ptied_newS": "interval"}
_NEWS = ("newS", "tied_newS", "ptied_newS")

# reference parameter name -> slot of dladmm_layer
_SLOT_OF = {
    "A": {"beta1": "beta1", "beta2": "beta2", "active_para": "theta1", "active_para1": "theta2"},
    "B": {"beta1": "beta1", "beta2": "beta2", "beta3": "beta3", "ss1": "ss1", "ss2": "ss2",
          "active_para": "theta1", "active_para1": "theta2"},
    "C": {"beta1": "beta1", "beta3": "beta3", "ss2_1": "ss2", "ss2_2": "ss2_2", "active_para": "theta1"},
}
_FAMILY_KEY = {_lib.FAMILY_A: "A", _lib.FAMILY_B: "B", _lib.FAMILY_C: "C"}


# Shapes from which the split-precision mode (one kind::tf32 pass + two bf16 correction passes) is both faster and more accurate
# than three kind::tf32 passes.  Measured error models per product with reduction length K (profiles/r02_precision_table.md):
#   tf32x3       7.5e-9 * K   -- the tensor core adds into its fp32 accumulator with round-toward-zero, once per MMA (3 K / 8 MMAs)
#   tf32_bf16x2  5e-9 * K (K / 4 MMAs) plus ~1e-6 of unbiased bf16 rounding of the correction operands, independent of K
# They cross at K ~ 180.
AUTO_MIXED_MIN_DIM = 192


def default_precision():
    """DLADMM_PRECISION, or "auto": tf32_bf16x2 when min(m, d) >= AUTO_MIXED_MIN_DIM, else tf32x3 (the more accurate fp32-class mode
    at that shape; below ~200 rows the products are latency-bound and the pass count does not matter)."""
    return os.environ.get("DLADMM_PRECISION", "auto")


def resolve_precision(precision, m, d):
    if precision == "auto":
        return "tf32_bf16x2" if min(m, d) >= AUTO_MIXED_MIN_DIM else "tf32x3"
    return precision


def _param_table(variant, m, d, bs):
    """(name, shape, init value) in the reference's registration order."""
    one = lambda *s: (s if s else (1, 1))
    if variant == "lena":        # main_lena.py:30-37
        return [("beta1", (m, bs), 1.0), ("beta2", (m, bs), 1.0)]
    if variant == "ltheta":      # main_syn_l1l1_ltheta.py:30-43
        return [("beta1", (m, 1), 1.0), ("beta2", (m, 1), 1.0), ("active_para", (d, 1), 0.025),
                ("active_para1", (m, 1), 0.06)]
    if variant == "scalar":      # main_syn_l1l1_scalar.py:50-65
        return [("beta1", one(), 1.0), ("beta2", one(), 1.0), ("beta3", one(), 1.0), ("ss2", one(), 1.0),
                ("active_para", one(), 0.2), ("active_para1", one(), 0.8)]
    if variant == "full":        # main_syn_l1l1_full.py:29-44
        return [("beta1", (m, 1), 1.0), ("beta2", (m, 1), 1.0), ("beta3", (m, 1), 1.0), ("ss2", (m, 1), 1.0),
                ("active_para", (d, 1), 0.2), ("active_para1", (m, 1), 0.8)]
    if variant == "tied":        # main_syn_l1l1_scalar_tied.py:50-67
        return [("beta1", one(), 1.0), ("beta2", one(), 1.0), ("beta3", one(), 1.0), ("ss1", one(), 1.0),
                ("ss2", one(), 1.0), ("active_para", one(), 1e-4), ("active_para1", one(), 1e-2)]
    if variant == "lasso":       # main_syn_lasso_scalar.py:33-50
        return [("beta1", one(), 1.0), ("beta3", one(), 1.0), ("ss2_1", one(), 0.5), ("ss2_2", one(), 0.5),
                ("active_para", one(), 0.2)]
    if variant == "newS":        # main_syn_scalar_newS_layerwise.py:50-68
        return [("beta1", one(), 1.0), ("beta2", one(), 1.0), ("beta3", one(), 1.0), ("ss2", one(), 1.0),
                ("active_para", one(), 0.1), ("active_para1", one(), 0.1)]
    if variant in ("tied_newS", "ptied_newS"):   # main_syn_scalar_tied_newS_layerwise.py:51-68, ptied: :53-74
        return [("beta1", one(), 1.0), ("beta2", one(), 1.0), ("beta3", one(), 1.0), ("ss1", one(), 1.0),
                ("ss2", one(), 1.0), ("active_para", one(), 0.01), ("active_para1", one(), 0.01)]
    raise ValueError("unknown variant %r" % (variant,))


class DLADMMNet(nn.Module):
    """Drop-in for the reference ``DLADMMNet(m, n, d, batch_size, A, Z0, E0, L0, layers)``.

    ``variant`` selects which of the reference's re-declared classes is mirrored (default ``scalar``);
    the per-variant subclasses below fix it so a script can swap its inline class for an import.
    ``precision``: "auto" (default: "tf32_bf16x2" from min(m, d) >= 192, "tf32x3" below), "tf32_bf16x2" (tcgen05, one kind::tf32
    pass + two bf16 correction passes, fp32-level products), "tf32x3" (three kind::tf32 passes), "tf32" / "bf16" (single pass,
    stated-tolerance options), "fp32" (CUDA-core FFMA).
    """
    variant = "scalar"

    def __init__(self, m, n, d, batch_size, A, Z0, E0, L0, layers, interval=None, variant=None, precision=None, device=None):
        super(DLADMMNet, self).__init__()
        if variant is not None:
            self.variant = variant
        if self.variant not in _FAMILY:
            raise ValueError("unknown variant %r" % (self.variant,))
        self._tie = _TIE.get(self.variant)
        if self._tie == "interval":      # main_syn_scalar_ptied_newS_layerwise.py:38,49,63-64
            if interval is None or int(interval) <= 0:
                raise ValueError("the ptied variant needs a positive `interval`")
            self.interval = int(interval)
            if layers // self.interval * self.interval < layers:
                raise ValueError("layers (%d) must be covered by layers // interval (%d) weights" % (layers, layers // self.interval))
        elif interval is not None:
            raise ValueError("`interval` is only meaningful for the ptied variant")
        self.m, self.n, self.d = m, n, d
        self.batch_size = batch_size
        self.layers = layers
        self.precision = resolve_precision(precision or default_precision(), int(m), int(d))
        if self.precision not in _lib.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_lib.PRECISIONS))
        if device is None:
            # the reference calls .cuda() unconditionally (main_syn_l1l1_scalar.py:40-44); on a box
            # without a GPU the module can still be built (state_dict work) but forward() raises.
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self._device = torch.device(device)
        # plain attributes, not buffers: absent from state_dict exactly like the reference
        self.A = A.detach().to(self._device, torch.float32)
        self.At = self.A.t()
        self.Z0 = Z0.detach().to(self._device, torch.float32)
        self.E0 = E0.detach().to(self._device, torch.float32)
        self.L0 = L0.detach().to(self._device, torch.float32)
        self._lipschitz = None

        table = _param_table(self.variant, m, d, batch_size)
        for name, _, _ in table:
            setattr(self, name, nn.ParameterList())
        self.fc = nn.Linear(m, d, bias=False) if self._tie == "all" else nn.ModuleList()
        if self._tie == "interval":
            for _ in range(layers // self.interval):
                self.fc.append(nn.Linear(m, d, bias=False))
        for k in range(layers):
            for name, shape, val in table:
                getattr(self, name).append(nn.Parameter(val * torch.ones(shape, dtype=torch.float32)))
            if self._tie is None:
                self.fc.append(nn.Linear(m, d, bias=False))
        if self.variant == "lena":   # fixed, non-learnable thresholds (main_lena.py:40-41)
            self.active_para = torch.tensor(0.025, dtype=torch.float32, device=self._device)
            self.active_para1 = torch.tensor(0.06, dtype=torch.float32, device=self._device)
        # W init: A^T + 1e-3*randn, times 0.4 outside family A (main_lena.py:49, main_syn_l1l1_scalar.py:72)
        scale = 1.0 if _FAMILY[self.variant] == _lib.FAMILY_A else 0.4
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                At = self.A.t()
                mod.weight = nn.Parameter(((At + 1e-3 * torch.randn_like(At)) * scale).contiguous())
        self.to(self._device)

    # ---- reference surface ---------------------------------------------------------------------
    def name(self):
        if self.variant == "ptied_newS":     # main_syn_scalar_ptied_newS_layerwise.py:123
            return "DLADMMNet_scalar_ptied{}_newS_layerwise".format(self.interval)
        return _NAME.get(self.variant, "DLADMMNet")

    def self_active(self, x, thershold):
        """main_syn_l1l1_scalar.py:76-77 (kept for scripts that call it on returned tensors)."""
        return torch.relu(x - thershold) - torch.relu(-1.0 * x - thershold)

    @property
    def L(self):
        """||A^T A||_2 as a (1,1) tensor (main_syn_l1l1_scalar.py:47-48), computed lazily."""
        if self._lipschitz is None:
            A64 = self.A.detach().double().cpu()
            val = torch.linalg.matrix_norm(A64.t() @ A64, ord=2).item()
            self._lipschitz = (val * torch.ones(1, 1)).float().to(self.A.device)
        return self._lipschitz

    def _apply(self, fn, *args, **kwargs):
        # keep the plain-tensor attributes with the parameters when the user calls .cuda()/.to()
        out = super(DLADMMNet, self)._apply(fn, *args, **kwargs)
        for attr in ("A", "Z0", "E0", "L0"):
            t = getattr(self, attr, None)
            if isinstance(t, torch.Tensor):
                setattr(self, attr, fn(t))
        if isinstance(getattr(self, "A", None), torch.Tensor):
            self.At = self.A.t()
            if self.variant == "lena":
                self.active_para = fn(self.active_para)
                self.active_para1 = fn(self.active_para1)
        self._lipschitz = None
        self.__dict__.pop("_spec_cache", None)       # (the lena thresholds above are new tensors)
        return out

    # ---- call description for the library ---------------------------------------------------------
    def sync_gradients(self, enable=True, group=None, bucket_mb=None):
        """Data-parallel training over column shards: with this on, every backward through the module sum-allreduces the
        step's parameter gradients across the ranks of `group` (default: the world) in place, inside the backward, so that
        `.grad` already holds the global sum (normalise the loss by the GLOBAL batch).  Replaces a separate
        `allreduce_gradients(model.parameters())` call and its staging copies.
        Default: ONE collective, stream-ordered after the last backward kernel.  `bucket_mb` (opt-in): when the weight
        gradients of a step exceed two buckets of that size, the weights are reduced bucket by bucket in reverse layer order
        from a side stream while the layers below are still running (function.plan_gradient_buckets; the library records one
        event per bucket, dladmm_cotangents.layer_events); the result is the same sum.  Measured on 2 and 8 B200s at 320 MB of
        weight gradients the two schedules take the same time, hence the simpler default."""
        self._grad_sync = (group,) if enable else None
        self._grad_bucket_bytes = None if (bucket_mb is None or bucket_mb <= 0) else int(bucket_mb * (1 << 20))
        self.__dict__.pop("_spec_cache", None)
        return self

    def _weight_module(self, k):
        if self._tie == "all":
            return self.fc
        if self._tie == "interval":
            return self.fc[k // self.interval]
        return self.fc[k]

    def _spec_and_params(self, nlayers=None, drop_last_estep=False):
        """Flat parameter list + per-layer slot map for the first `nlayers` layers.  Layers that share an fc (tied / ptied)
        point at the same list entry, so the library sees one weight pointer and accumulates one gradient.
        `drop_last_estep`: the E/L-step parameters of the last layer do not reach any output (newS ordering); they are
        passed detached so that, as in the reference, they get no gradient.
        The description is static for a module, so it is cached and revalidated by parameter identity (walking
        nn.ParameterList per call cost ~0.25 ms per forward at K=15)."""
        K = self.layers if nlayers is None else int(nlayers)
        key = (K, bool(drop_last_estep), self.precision)
        cache = self.__dict__.setdefault("_spec_cache", {})
        hit = cache.get(key)
        if hit is not None:
            spec, entries, owners = hit
            cur = self._registered_parameters()
            if len(cur) == len(owners) and all(o is q for o, q in zip(owners, cur)):
                return spec, [p.detach() if det else p for p, det in entries]
        spec, entries = self._build_spec_and_params(K, drop_last_estep)
        cache.clear()
        cache[key] = (spec, entries, self._registered_parameters())
        return spec, [p.detach() if det else p for p, det in entries]

    def _registered_parameters(self):
        """Every registered parameter slot of the module tree, in registration order (what nn.Module.parameters() walks, without its
        name formatting and de-duplication set: 10 us instead of 70 at K = 15).  Only compared by identity against an earlier walk."""
        out = []
        for mod in self.modules():
            out.extend(mod._parameters.values())
        return out

    def _build_spec_and_params(self, K, drop_last_estep):
        fam = _FAMILY[self.variant]
        names = _SLOT_OF[_FAMILY_KEY[fam]]
        params, slots, weights, windex = [], [], [], {}
        table = _param_table(self.variant, self.m, self.d, self.batch_size)
        for k in range(K):
            s = {}
            for name, _, _ in table:
                det = drop_last_estep and k == K - 1 and name in ("beta2", "beta3", "ss2", "active_para1")
                s[names[name]] = len(params)
                params.append((getattr(self, name)[k], det))
            slots.append(s)
            wm = self._weight_module(k)
            if id(wm) not in windex:
                windex[id(wm)] = len(params)
                params.append((wm.weight, False))
            weights.append(windex[id(wm)])
        fixed = {}
        if self.variant == "lena":
            fixed = {"theta1": self.active_para, "theta2": self.active_para1}
        spec = LayerSpec(fam, self.m, self.d, K, _lib.PRECISIONS[self.precision], slots, weights, fixed)
        spec.grad_sync = getattr(self, "_grad_sync", None)
        spec.grad_bucket_bytes = getattr(self, "_grad_bucket_bytes", None)
        return spec, params

    def _forward_newS(self, x, K):
        """forward(x, K) of the E -> L -> Z ordering (main_syn_scalar_newS_layerwise.py:82-112): the first min(K, layers)
        layers of the family-B recursion, returned as Z[k] = Z_{k+1}, E = [E0, E_1, ..], L = [L0, L_1, ..] -- i.e. the
        E/L lists shifted by one and the last layer's E/L-step not part of the result."""
        Kp = min(self.layers if K is None else int(K), self.layers)
        if Kp <= 0:
            return [], [], []
        spec, params = self._spec_and_params(Kp, drop_last_estep=True)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            Z, E, L, _ = UnrolledLADMM.apply(spec, self.A, x, self.Z0, self.E0, self.L0, *params)
        else:
            Z, E, L, _, _, _ = run_forward(spec, self.A, x, self.Z0, self.E0, self.L0, [p.detach() for p in params],
                                           want_masks=False)
        return list(Z.unbind(0)), [self.E0] + list(E.unbind(0))[:Kp - 1], [self.L0] + list(L.unbind(0))[:Kp - 1]

    def forward(self, x, K=None, last_only=False):
        """x: (m, B) float32 CUDA, B equal to the batch of Z0/E0/L0.  Returns lists (Z, E, L[, T]) of the
        per-layer iterates like the reference (a10).  ``last_only=True`` (inference, opt-in) keeps only
        the final iterate in each list and never materialises the other K-1.  ``K`` (newS variants only, as in
        their reference signature forward(x, K)): run the first min(K, layers) layers."""
        if not x.is_cuda:
            raise RuntimeError("DLADMMNet.forward needs a CUDA tensor: d-ladmm_b200 has no CPU path")
        if self.variant in _NEWS:
            if last_only:
                raise RuntimeError("last_only is not offered for the newS variants")
            return self._forward_newS(x, K)
        if K is not None:
            raise TypeError("forward(x, K) exists only in the newS variants")
        spec, params = self._spec_and_params()
        K = self.layers
        train = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if last_only:
            if train:
                raise RuntimeError("last_only=True is inference-only; wrap the call in torch.no_grad()")
            Z, E, L, T, _, _ = run_forward(spec, self.A, x, self.Z0, self.E0, self.L0, params, want_masks=False, last_only=True)
            Zl, El, Ll, Tl = [Z[(K - 1) % 2]], [E[(K - 1) % 2]], [L[(K - 1) % 2]], [T[K % 2]]
        else:
            if train:
                Z, E, L, T = UnrolledLADMM.apply(spec, self.A, x, self.Z0, self.E0, self.L0, *params)
            else:
                with torch.no_grad():            # (run_forward only takes addresses: no per-parameter detach needed)
                    Z, E, L, T, _, _ = run_forward(spec, self.A, x, self.Z0, self.E0, self.L0, params, want_masks=False)
            Zl, El, Ll, Tl = list(Z.unbind(0)), list(E.unbind(0)), list(L.unbind(0)), list(T.unbind(0))
        if _RETURNS_T[self.variant]:
            return Zl, El, Ll, Tl
        return Zl, El, Ll


    def forward_objective(self, x, alpha, last_only=False):
        """Inference: forward(x) and the per-layer L1-L1 objective of the reference's evaluation loop
        (main_syn_l1l1_scalar.py:333-334), obj[k] = sum_b ( alpha*||Z_k[:,b]||_1 + ||x[:,b] - A Z_k[:,b]||_1 ), accumulated
        inside the product epilogues (no second pass over the iterates, no A@Z_k products).  Returns (obj (K,), outputs)
        with `outputs` what forward(x, last_only) returns."""
        if not x.is_cuda:
            raise RuntimeError("DLADMMNet.forward_objective needs a CUDA tensor: d-ladmm_b200 has no CPU path")
        spec, params = self._spec_and_params()
        K = self.layers
        extras = {}
        with torch.no_grad():
            Z, E, L, T, _, _ = run_forward(spec, self.A, x, self.Z0, self.E0, self.L0, [p.detach() for p in params],
                                           want_masks=False, last_only=last_only, objective_alpha=float(alpha), extras=extras)
        if last_only:
            if self.variant in _NEWS:
                raise RuntimeError("last_only is not offered for the newS variants")
            Zl, El, Ll, Tl = [Z[(K - 1) % 2]], [E[(K - 1) % 2]], [L[(K - 1) % 2]], [T[K % 2]]
            outs = (Zl, El, Ll, Tl) if _RETURNS_T[self.variant] else (Zl, El, Ll)
        else:
            outs = self._as_lists(Z, E, L, T)
        return extras["objective"], outs

    def _as_lists(self, Z, E, L, T):
        Zl, El, Ll, Tl = list(Z.unbind(0)), list(E.unbind(0)), list(L.unbind(0)), list(T.unbind(0))
        if self.variant in _NEWS:        # the variant's own return convention: E = [E0, E_1, ..], L = [L0, L_1, ..]
            return Zl, [self.E0] + El[:-1], [self.L0] + Ll[:-1]
        return (Zl, El, Ll, Tl) if _RETURNS_T[self.variant] else (Zl, El, Ll)

    def _fused_loss(self, kind, x, alpha, layer_weights, global_batch):
        if not x.is_cuda:
            raise RuntimeError("the fused training losses need a CUDA tensor: d-ladmm_b200 has no CPU path")
        spec, params = self._spec_and_params()
        w = [1.0] * self.layers if layer_weights is None else [float(v) for v in layer_weights]
        if len(w) != self.layers:
            raise ValueError("layer_weights must have one entry per layer")
        B = x.shape[1]
        if global_batch is None:
            global_batch = B
            if spec.grad_sync is not None:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    global_batch = B * dist.get_world_size(spec.grad_sync[0])
        loss, Z, E, L, T = UnrolledLADMML1L1.apply(spec, kind, float(alpha), w, self.A, x, self.Z0, self.E0, self.L0, *params)
        if global_batch != B:
            loss = loss * (float(B) / float(global_batch))
        return loss, self._as_lists(Z, E, L, T)

    def l1l1_loss(self, x, alpha, layer_weights=None, global_batch=None):
        """Fused training objective of the reference drivers (main_syn_l1l1_scalar.py:289-299):

            loss = sum_k w_k * ( alpha * mean_b ||Z_k[:,b]||_1 + mean_b ||x[:,b] - A Z_k[:,b]||_1 )

        with w_k = layer_weights[k] (the scripts use 0.6**epoch for k < K-1 and 1 for the last layer).
        Returns (loss, outputs) where `outputs` is what forward(x) returns, detached.  loss.backward() runs
        the library's backward with the loss cotangents generated inside the kernels.

        Normalisation under data parallelism: the mean is over `global_batch` columns -- by default this rank's B, or
        B x world size when `sync_gradients()` is on (equal column shards), so that the SUM-allreduced `.grad` is the
        gradient of the global-mean loss and each rank's returned loss is its share of it (they add up over ranks)."""
        return self._fused_loss(1, x, alpha, layer_weights, global_batch)

    def lasso_loss(self, x, alpha, layer_weights=None, global_batch=None):
        """Fused LASSO training objective (main_syn_lasso_scalar.py:276-281):

            loss = sum_k w_k * ( alpha * mean_b ||Z_k[:,b]||_1 + 0.5 * mean_b ||x[:,b] - A Z_k[:,b]||_2^2 )

        Same mechanics and normalisation as l1l1_loss (objective_kind / loss_kind = 2 of the C ABI)."""
        return self._fused_loss(2, x, alpha, layer_weights, global_batch)

    def forward_metrics(self, x, want, Z_label=None, E_label=None, X_clean=None, dual_alpha=0.0, last_only=False):
        """Inference: forward(x) with the per-layer sums the reference's drivers compute script-side accumulated inside
        the product epilogues (no second pass over the iterates, no extra A@Z_k products).  `want`: names out of
        _lib.METRICS.  Returns (dict name -> (K,) tensor of sums over the batch, outputs):

            l1_z sum|Z_k|; sqerr_z sum(Z_label-Z_k)^2; l1_res sum|x-A Z_k|; sq_res sum(x-A Z_k)^2; sqerr_e sum(E_label-E_k)^2;
            sqerr_az sum(X_clean-A Z_k)^2; l1_e sum|E_k|; dot_lx sum L_k*x; dgap_l sum dual_gap(L_k,1);
            dgap_atl sum dual_gap(A^T L_k, dual_alpha)  (one extra product per layer, only when asked for)

        See nmse_db / psnr_db / dual_gap_loss below for the reference's formulas on top of these sums."""
        if not x.is_cuda:
            raise RuntimeError("DLADMMNet.forward_metrics needs a CUDA tensor: d-ladmm_b200 has no CPU path")
        if self.variant in _NEWS and last_only:
            raise RuntimeError("last_only is not offered for the newS variants")
        spec, params = self._spec_and_params()
        K = self.layers
        extras = {}
        desc = dict(want=list(want), Z_label=Z_label, E_label=E_label, X_clean=X_clean, dual_alpha=dual_alpha)
        with torch.no_grad():
            Z, E, L, T, _, _ = run_forward(spec, self.A, x, self.Z0, self.E0, self.L0, [p.detach() for p in params],
                                           want_masks=False, last_only=last_only, extras=extras, metrics=desc)
        if last_only:
            Zl, El, Ll, Tl = [Z[(K - 1) % 2]], [E[(K - 1) % 2]], [L[(K - 1) % 2]], [T[K % 2]]
            outs = (Zl, El, Ll, Tl) if _RETURNS_T[self.variant] else (Zl, El, Ll)
        else:
            outs = self._as_lists(Z, E, L, T)
        out = extras["metrics"]
        return {name: out[:, _lib.METRICS.index(name)] for name in want}, outs

    @staticmethod
    def nmse_db(sums, Z_label, E_label):
        """NMSE per layer in dB as test_syn_l1l1_scalar.py:537-541 forms it from the accumulated squared errors:
        10 log10( sum(Z_label-Z_k)^2 / sum Z_label^2 + sum(E_label-E_k)^2 / sum E_label^2 )  (the 1/n_test factors cancel)."""
        return 10.0 * torch.log10(sums["sqerr_z"] / Z_label.pow(2).sum() + sums["sqerr_e"] / E_label.pow(2).sum())

    @staticmethod
    def psnr_db(sums, numel):
        """PSNR per layer as main_lena.py:138-143 / 262-267: mse = mean((255*gt - 255*A Z_k)^2) over `numel` elements,
        psnr = -10 log10(mse) + 48.131."""
        return -10.0 * torch.log10(sums["sqerr_az"] * (255.0 * 255.0 / float(numel))) + 48.131

    def dual_gap_loss(self, sums, alpha, B):
        """Per-layer training loss of main_lena.py:221-228 from the fused sums (values only; training on it goes through
        the generic autograd path): alpha*mean|Z_k| + mean|E_k| + mean dual_gap(A^T L_k, alpha) + mean dual_gap(L_k, 1)
        + mean(L_k * x)."""
        nz, nm = float(self.d * B), float(self.m * B)
        return (alpha * sums["l1_z"] / nz + sums["l1_e"] / nm + sums["dgap_atl"] / nz + sums["dgap_l"] / nm + sums["dot_lx"] / nm)


    # ---- classical LADMM step, safeguard operator and safeguarded evaluation ------------------------------------------
    #   Z -> E -> L ordering (scalar / full / tied):  KM, S, forward_safeguarded   test_syn_l1l1_scalar.py:126-317
    #   E -> L -> Z ordering (newS variants):        KM_ZEL, KM_ELZ, Snorm_ELZ, forward_safeguarded
    #                                                                              test_syn_l1l1_newS_Acols.py:136-277
    # Every step is a K = 1 or K = 2 call of the library from an arbitrary state (dladmm_problem.T_init / start_half /
    # stop_half); norms, the per-column selection and the mu_k update are csrc/safeguard.cu.
    def two_norm(self, z, dim=0):
        """test_syn_l1l1_scalar.py:126-128"""
        return (z ** 2).sum(dim=dim).sqrt()

    def _lip(self):
        """||A^T A||_2 as a host float (the tensor `self.L` stays for scripts): no device sync per KM call."""
        if getattr(self, "_lip_float", None) is None or self._lipschitz is None:
            self._lip_float = float(self.L.item())
        return self._lip_float

    def _steps(self, spec, params, Zk, Ek, Lk, X, T_init=None, start_half=False, stop_half=False):
        Z, E, L, T, _, _ = run_forward(spec, self.A, X, Zk, Ek, Lk, params, want_masks=False, T_init=T_init,
                                       start_half=start_half, stop_half=stop_half)
        return Z, E, L, T

    def _t0(self, X, Z0=None, E0=None):
        """T_0 = A Z0 + E0 - X through the library (a K = 0 call: the first fused product alone)."""
        spec0 = LayerSpec(_FAMILY[self.variant], self.m, self.d, 0, _lib.PRECISIONS[self.precision], [], [])
        z = self.Z0 if Z0 is None else Z0
        e = self.E0 if E0 is None else E0
        return self._steps(spec0, [], z, e, e, X)[3][0]

    def _one_layer(self, spec1, params, Zk, Ek, Lk, Tk, X):
        Z, E, L, T = self._steps(spec1, params, Zk, Ek, Lk, X, T_init=Tk)
        return Z[0], E[0], L[0], T[1]

    def _scalar(self, v):
        return torch.full((1, 1), float(v), dtype=torch.float32, device=self.A.device)

    def _sub_spec(self, spec, params, ks):
        """The call description of layers `ks` alone, with ONLY their parameters (re-indexed), kept on the parent spec: the
        single-layer calls of the safeguarded evaluation then validate ~8 tensors instead of the module's 7K and reuse their
        layer table (run_forward's cache is per spec object)."""
        if os.environ.get("DLADMM_SG_NOSUB"):          # A/B switch: a fresh description over the full parameter list per call
            return LayerSpec(spec.family, self.m, self.d, len(ks), spec.precision, [spec.slots[k] for k in ks],
                             [spec.weights[k] for k in ks], spec.fixed), params
        cache = spec.__dict__.setdefault("_sub", {})
        key = tuple(ks)
        hit = cache.get(key)
        if hit is None:
            idx, remap = [], {}

            def r(i):
                if i not in remap:
                    remap[i] = len(idx)
                    idx.append(i)
                return remap[i]
            slots = [{name: r(i) for name, i in spec.slots[k].items() if i is not None} for k in ks]
            weights = [r(spec.weights[k]) for k in ks]
            hit = cache[key] = (LayerSpec(spec.family, self.m, self.d, len(ks), spec.precision, slots, weights, spec.fixed), idx)
        sub, idx = hit
        return sub, [params[i] for i in idx]

    def _km_spec(self, beta, ss1, ss2, alpha):
        """Classical step of the Z -> E -> L ordering as ONE family-B layer: W = A^T with ss1 in the tied slot, all betas =
        beta, thresholds ss1*alpha and ss2 (test_syn_l1l1_scalar.py:131-160)."""
        key = ("zel_b", float(beta), float(ss1), float(ss2), float(alpha))
        cache = self.__dict__.setdefault("_km_cache", {})
        if key not in cache:
            s = self._scalar
            params = [self.A.t().contiguous(), s(beta), s(ss1), s(ss2), s(float(ss1) * float(alpha)), s(ss2)]
            slots = {"beta1": 1, "beta2": 1, "beta3": 1, "ss1": 2, "ss2": 3, "theta1": 4, "theta2": 5}
            spec = LayerSpec(_lib.FAMILY_B, self.m, self.d, 1, _lib.PRECISIONS[self.precision], [slots], [0])
            if len(cache) > 8:
                cache.clear()
            cache[key] = (spec, params)
        return cache[key]

    def _km_spec_elz(self, beta, ss1, alpha, K):
        """Classical step with ss2 = 1/beta (KM_ZEL / KM_ELZ, test_syn_l1l1_newS_Acols.py:136-171): K identical family-A
        layers  Z = act(Z - ss1 A^T (L + beta T), ss1 alpha),  E = act(X - A Z - L/beta, 1/beta),  L += beta T."""
        key = ("elz_a", float(beta), float(ss1), float(alpha), int(K))
        cache = self.__dict__.setdefault("_km_cache", {})
        if key not in cache:
            s = self._scalar
            params = [self.A.t().contiguous(), s(beta), s(1.0 / beta), s(ss1), s(float(ss1) * float(alpha)), s(1.0 / beta)]
            slots = {"beta1": 1, "beta2": 2, "ss1": 3, "theta1": 4, "theta2": 5}
            spec = LayerSpec(_lib.FAMILY_A, self.m, self.d, int(K), _lib.PRECISIONS[self.precision], [slots] * int(K), [0] * int(K))
            if len(cache) > 8:
                cache.clear()
            cache[key] = (spec, params)
        return cache[key]

    def _need(self, newS, what):
        ok = _FAMILY[self.variant] == _lib.FAMILY_B and ((self.variant in _NEWS) == newS)
        if not ok:
            raise NotImplementedError("%s exists for the %s variants only (as in the reference's evaluation scripts)" %
                                      (what, "newS (E -> L -> Z)" if newS else "scalar / full / tied (Z -> E -> L)"))

    def KM(self, Zk, Ek, Lk, Tk, X, **kwargs):
        """One classical LADMM iteration (test_syn_l1l1_scalar.py:131-160).  Returns (Varn, Zn, En, Tn, Ln)."""
        self._need(False, "KM")
        beta = float(kwargs.get("beta", 1.0))
        ss1 = kwargs.get("ss1", None)
        ss1 = 0.999 / self._lip() if ss1 is None else float(ss1)
        ss2 = float(kwargs.get("ss2", 0.3))
        alpha = float(kwargs.get("alpha", 0.01))
        spec1, params = self._km_spec(beta, ss1, ss2, alpha)
        Zn, En, Ln, Tn = self._one_layer(spec1, params, Zk, Ek, Lk, Tk, X)
        return Lk + beta * Tk, Zn, En, Tn, Ln

    def S(self, Zk, Ek, Lk, Tk, X, Ep, **kwargs):
        """Safeguard operator (test_syn_l1l1_scalar.py:163-176): [beta*T^{k+1} ; c*(E^{k+1} - 2 E^k + E^{k-1})]."""
        beta, ss2 = 1.0, 0.3
        _, _, En, Tn, _ = self.KM(Zk, Ek, Lk, Tk, X, beta=beta, ss2=ss2, alpha=kwargs.get("alpha", 0.01))
        c1 = beta * ss2
        c = sqrt(c1 / (1 - c1))
        return torch.cat([beta * Tn, c * (En - 2 * Ek + Ep)])

    def _s_norm(self, Zk, Ek, Lk, Tk, X, Ep, alpha):
        self._need(False, "S")
        beta, ss2 = 1.0, 0.3
        spec1, params = self._km_spec(beta, 0.999 / self._lip(), ss2, alpha)
        _, En, _, Tn = self._one_layer(spec1, params, Zk, Ek, Lk, Tk, X)
        c1 = beta * ss2
        out = torch.empty(X.shape[1], dtype=torch.float32, device=X.device)
        stream = torch.cuda.current_stream(X.device).cuda_stream
        _lib.check(_lib.load().dladmm_sg_norm(self.m, X.shape[1], beta, sqrt(c1 / (1 - c1)), Tn.data_ptr(), En.data_ptr(),
                                              Ek.contiguous().data_ptr(), Ep.contiguous().data_ptr(), out.data_ptr(), stream))
        return out

    # E -> L -> Z ordering ------------------------------------------------------------------------------------------
    def KM_ZEL(self, Zk, Ek, Lk, Tk, X, **kwargs):
        """test_syn_l1l1_newS_Acols.py:136-152: classical step in Z -> E -> L order with ss2 = 1/beta.
        Returns (Varn, Zn, En, Tn, Ln)."""
        self._need(True, "KM_ZEL")
        beta = float(kwargs.get("beta", 1.0))
        ss1 = float(kwargs.get("ss1", 0.999 / self._lip()))
        spec1, params = self._km_spec_elz(beta, ss1, float(kwargs.get("alpha", 0.01)), 1)
        Zn, En, Ln, Tn = self._one_layer(spec1, params, Zk, Ek, Lk, Tk, X)
        return Lk + beta * Tk, Zn, En, Tn, Ln

    def _km_elz(self, Ek, Lk, Zn, X, beta, ss1, alpha, with_next_estep=False):
        """E- and L-step from the given Z, then the next Z-step: one K = 2 call that starts with an E/T/L half layer
        (start_half) and, unless the safeguard norm needs A Znn, stops after the Z half of the second (stop_half)."""
        spec2, params = self._km_spec_elz(beta, ss1, alpha, 2)
        return self._steps(spec2, params, Zn, Ek, Lk, X, start_half=True, stop_half=not with_next_estep)

    def KM_ELZ(self, Ek, Lk, Zn, X, **kwargs):
        """test_syn_l1l1_newS_Acols.py:155-171.  Returns (En, Tn, Ln, Varnn, Znn)."""
        self._need(True, "KM_ELZ")
        beta = float(kwargs.get("beta", 1.0))
        ss1 = float(kwargs.get("ss1", 0.999 / self._lip()))
        Z, E, L, T = self._km_elz(Ek, Lk, Zn, X, beta, ss1, float(kwargs.get("alpha", 0.01)))
        return E[0], T[1], L[0], L[0] + beta * T[1], Z[1]

    def Snorm_ELZ(self, Ek, Lk, Zn, X, **kwargs):
        """test_syn_l1l1_newS_Acols.py:174-192: sqrt(||A Znn + En - X||^2 + (Znn - Zn)^T P2 (Znn - Zn)) per column, with
        P2 = I/(beta ss1) - A^T A.  Computed without the d x d operator: the quadratic form is
        ||Znn - Zn||^2/(beta ss1) - ||A (Znn - Zn)||^2 and A (Znn - Zn) = Tnn - Tn comes out of the fused products."""
        self._need(True, "Snorm_ELZ")
        beta, ss1 = 1.0, 0.999 / self._lip()
        Zn = Zn.contiguous()
        Z, E, L, T = self._km_elz(Ek, Lk, Zn, X, beta, ss1, float(kwargs.get("alpha", 0.01)), with_next_estep=True)
        B = X.shape[1]
        out = torch.empty(B, dtype=torch.float32, device=X.device)
        if Z.shape[-1] != B or not Z.is_contiguous():      # (narrowed views of a padded call: the kernels take pitch B)
            Z, E, T = Z.contiguous(), E.contiguous(), T.contiguous()
        stream = torch.cuda.current_stream(X.device).cuda_stream
        # T[2] = A Znn + E' - X with E' = E[1] (the E-step after Znn); swap E' for En = E[0]
        _lib.check(_lib.load().dladmm_sg_norm_elz(self.m, self.d, B, 1.0 / (beta * ss1), T[2].data_ptr(), T[1].data_ptr(),
                                                  Z[1].data_ptr(), Zn.data_ptr(), E[1].data_ptr(), E[0].data_ptr(),
                                                  out.data_ptr(), stream))
        return out

    def _select(self, pairs_in, s_norm, mu_state, delta, keep_row):
        """Per-column choice between the learned and the classical candidate + the mu_k update, on the device.
        mu_state = [mu tensor (B,), method name, param, host updater or None]."""
        lib = _lib.load()
        B = s_norm.shape[0]
        outs = [torch.empty_like(a) for a, _ in pairs_in]
        pairs = (_lib.SgPair * len(pairs_in))()
        for i, ((a, b_), o) in enumerate(zip(pairs_in, outs)):
            pairs[i].a, pairs[i].b, pairs[i].out, pairs[i].rows = a.data_ptr(), b_.data_ptr(), o.data_ptr(), a.shape[0]
        stream = torch.cuda.current_stream(s_norm.device).cuda_stream
        mu, method, param, host_updater = mu_state
        if host_updater is None:
            _lib.check(lib.dladmm_sg_select_update(len(pairs_in), pairs, B, s_norm.data_ptr(), mu.data_ptr(), float(1.0 - delta),
                                                   METHOD_ID[method], float(param), keep_row.data_ptr(), None, stream))
        else:                       # "RM": a window of recent norms, kept by the host-side updater object
            _lib.check(lib.dladmm_sg_select(len(pairs_in), pairs, B, s_norm.data_ptr(), mu.data_ptr(), float(1.0 - delta),
                                            keep_row.data_ptr(), stream))
            mu_state[0] = host_updater.step(s_norm, keep_row).contiguous()
        return outs

    def _mu_state(self, mu0, method, param):
        if method not in mu_updater_dict:
            raise ValueError("unknown mu_k updater %r" % (method,))
        host = None if method in METHOD_ID else mu_updater_dict[method](mu0, param)
        mu = mu0.clone() if host is None else mu0          # the fused kernel updates mu in place
        return [mu.contiguous(), method, param, host]

    def forward_safeguarded(self, x, use_learned, use_safeguard, continued=False, K=None, num_iter=200, delta=-99.0,
                            mu_k_method="None", mu_k_param=0.0, alpha=0.01):
        """The evaluation forward of the reference's test scripts: per layer the classical step, the learned step, the
        safeguard test ||S(u_L2O)|| < (1-delta)*mu_k per column, the per-column selection and the mu_k update.

        scalar / full / tied (test_syn_l1l1_scalar.py:179-317): returns (Z, E, L, T) lists;
        newS variants (test_syn_l1l1_newS_Acols.py:195-277):    returns (Z, E, L) with the initial variables first;
        plus sg_count (columns that fell back to the classical step, per layer) when both flags are set."""
        if _FAMILY[self.variant] != _lib.FAMILY_B:
            raise NotImplementedError("forward_safeguarded exists for the family-B variants (the reference only has it there)")
        layers = self.layers
        if K is None:
            K = layers if (not continued and (use_learned or use_safeguard)) else num_iter
        if use_safeguard and not use_learned:
            raise AssertionError("use_safeguard needs use_learned (test_syn_l1l1_scalar.py:240)")
        spec, params = self._spec_and_params()
        params = [p.detach() for p in params]
        # run at the padded pitch throughout (zero columns stay zero and are dropped at the end): the norm / selection
        # kernels address (rows x B) arrays with pitch B
        B_user = x.shape[1]
        Bp = padded_batch(spec, B_user)
        pad = lambda t: _pad_cols(t, Bp)
        X, Z0, E0, L0 = pad(x), pad(self.Z0), pad(self.E0), pad(self.L0)
        with torch.no_grad():
            if self.variant in _NEWS:
                out = self._safeguarded_newS(spec, params, X, Z0, E0, L0, use_learned, use_safeguard, continued, K, delta,
                                             mu_k_method, mu_k_param, alpha)
            else:
                out = self._safeguarded_zel(spec, params, X, Z0, E0, L0, use_learned, use_safeguard, continued, K, delta,
                                            mu_k_method, mu_k_param, alpha)
        lists, keep_all = out
        if Bp != B_user:
            lists = [[t[:, :B_user] for t in lst] for lst in lists]
        if keep_all is not None:
            cnt = (float(B_user) - keep_all[:, :B_user].sum(dim=1)).tolist()       # the one host read of the evaluation
            sg_count = [0.0] * layers
            for k in range(min(layers, len(cnt))):
                sg_count[k] = cnt[k]
            return tuple(lists) + (sg_count,)
        return tuple(lists)

    def _safeguarded_zel(self, spec, params, X, Z0, E0, L0, use_learned, use_safeguard, continued, K, delta, method, param, alpha):
        layers, B, dev = self.layers, X.shape[1], X.device
        km_spec, km_params = self._km_spec(1.0, 0.999 / self._lip(), 0.3, alpha)
        T = [self._t0(X, Z0, E0)]
        Z, E, L = [], [], []
        both = use_learned and use_safeguard
        keep_all = torch.ones((min(K, layers), B), dtype=torch.float32, device=dev) if both else None
        if both:
            mu_state = self._mu_state(self._s_norm(Z0, E0, L0, T[-1], X, E0, alpha), method, param)
        for k in range(K):
            if continued and k == layers:
                use_learned = use_safeguard = False
            Zp, Ep_, Lp = (Z0, E0, L0) if k == 0 else (Z[-1], E[-1], L[-1])
            Zn_KM, En_KM, Ln_KM, Tn_KM = self._one_layer(km_spec, km_params, Zp, Ep_, Lp, T[-1], X)
            if use_learned:
                spec1, params1 = self._sub_spec(spec, params, (k,))
                Zn, En, Ln, Tn = self._one_layer(spec1, params1, Zp, Ep_, Lp, T[-1], X)
            if use_safeguard:
                s_norm = self._s_norm(Zn, En, Ln, Tn, X, Ep_, alpha)
                outs = self._select([(Zn, Zn_KM), (En, En_KM), (Tn, Tn_KM), (Ln, Ln_KM)], s_norm, mu_state, delta, keep_all[k])
                Z.append(outs[0]); E.append(outs[1]); T.append(outs[2]); L.append(outs[3])
            elif use_learned:
                Z.append(Zn); E.append(En); T.append(Tn); L.append(Ln)
            else:
                Z.append(Zn_KM); E.append(En_KM); T.append(Tn_KM); L.append(Ln_KM)
        return [Z, E, L, T], keep_all

    def _safeguarded_newS(self, spec, params, X, Z0, E0, L0, use_learned, use_safeguard, continued, K, delta, method, param, alpha):
        layers, B, dev = self.layers, X.shape[1], X.device
        beta, ss1 = 1.0, 0.999 / self._lip()
        Z, E, L = [], [], []
        both = use_learned and use_safeguard
        keep_all = torch.ones((min(K, layers), B), dtype=torch.float32, device=dev) if both else None
        if both:
            mu_state = self._mu_state(self.Snorm_ELZ(E0, L0, Z0, X, alpha=alpha), method, param)
        for k in range(K):
            if continued and k == layers:
                use_learned = use_safeguard = False
            if k == 0:
                # both candidates are a Z-step from (Z0, E0, L0) with T_0 = A Z0 + E0 - X (computed by the call itself)
                km1, kmp = self._km_spec_elz(beta, ss1, alpha, 1)
                En_KM, Ln_KM = E0, L0
                Zn_KM = self._steps(km1, kmp, Z0, E0, L0, X, stop_half=True)[0][0]
                if use_learned:
                    spec1, params1 = self._sub_spec(spec, params, (0,))
                    En_L, Ln_L = E0, L0
                    Zn_L = self._steps(spec1, params1, Z0, E0, L0, X, stop_half=True)[0][0]
            else:
                Zk, Ek, Lk, _ = self._km_elz(E[-1], L[-1], Z[-1], X, beta, ss1, alpha)
                En_KM, Ln_KM, Zn_KM = Ek[0], Lk[0], Zk[1]
                if use_learned:
                    # E- and L-step with layer k-1's parameters, Z-step with layer k's (:234-242)
                    spec2, params2 = self._sub_spec(spec, params, (k - 1, k))
                    Zl, El, Ll, _ = self._steps(spec2, params2, Z[-1], E[-1], L[-1], X, start_half=True, stop_half=True)
                    En_L, Ln_L, Zn_L = El[0], Ll[0], Zl[1]
            if use_safeguard:
                s_norm = self.Snorm_ELZ(En_L, Ln_L, Zn_L, X, alpha=alpha)
                outs = self._select([(En_L, En_KM), (Ln_L, Ln_KM), (Zn_L, Zn_KM)], s_norm, mu_state, delta, keep_all[k])
                E.append(outs[0]); L.append(outs[1]); Z.append(outs[2])
            elif use_learned:
                E.append(En_L); L.append(Ln_L); Z.append(Zn_L)
            else:
                E.append(En_KM); L.append(Ln_KM); Z.append(Zn_KM)
        return [[Z0] + Z, [E0] + E, [L0] + L], keep_all


class DLADMMNetScalar(DLADMMNet):
    variant = "scalar"


class DLADMMNetFull(DLADMMNet):
    variant = "full"


class DLADMMNetTied(DLADMMNet):
    variant = "tied"


class DLADMMNetLasso(DLADMMNet):
    variant = "lasso"


class DLADMMNetLena(DLADMMNet):
    variant = "lena"


class DLADMMNetLtheta(DLADMMNet):
    variant = "ltheta"


class DLADMMNetNewS(DLADMMNet):
    variant = "newS"


class DLADMMNetTiedNewS(DLADMMNet):
    variant = "tied_newS"


class DLADMMNetPtiedNewS(DLADMMNet):
    variant = "ptied_newS"


VARIANT_CLASSES = {"scalar": DLADMMNetScalar, "full": DLADMMNetFull, "tied": DLADMMNetTied,
                   "lasso": DLADMMNetLasso, "lena": DLADMMNetLena, "ltheta": DLADMMNetLtheta,
                   "newS": DLADMMNetNewS, "tied_newS": DLADMMNetTiedNewS, "ptied_newS": DLADMMNetPtiedNewS}
