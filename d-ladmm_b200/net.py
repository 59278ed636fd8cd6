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
from .function import LayerSpec, UnrolledLADMM, UnrolledLADMML1L1, run_forward
from .mu_updater import mu_updater_dict

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


def default_precision():
    return os.environ.get("DLADMM_PRECISION", "tf32x3")


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
    ``precision``: "fp32" (CUDA-core FFMA), "tf32x3" / "tf32" (tcgen05).
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
        self.precision = precision or default_precision()
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
    def sync_gradients(self, enable=True, group=None):
        """Data-parallel training over column shards: with this on, every backward through the module sum-allreduces the
        step's parameter gradients across the ranks of `group` (default: the world) in place, inside the backward, so that
        `.grad` already holds the global sum (normalise the loss by the GLOBAL batch).  Replaces a separate
        `allreduce_gradients(model.parameters())` call and its staging copies."""
        self._grad_sync = (group,) if enable else None
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
            cur = list(self.parameters())
            if len(cur) == len(owners) and all(o is q for o, q in zip(owners, cur)):
                return spec, [p.detach() if det else p for p, det in entries]
        spec, entries = self._build_spec_and_params(K, drop_last_estep)
        cache.clear()
        cache[key] = (spec, entries, list(self.parameters()))
        return spec, [p.detach() if det else p for p, det in entries]

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
            Z, E, L, T, _, _ = run_forward(spec, self.A, x, self.Z0, self.E0, self.L0,
                                           [p.detach() for p in params], want_masks=False, last_only=True)
            Zl, El, Ll, Tl = [Z[(K - 1) % 2]], [E[(K - 1) % 2]], [L[(K - 1) % 2]], [T[K % 2]]
        else:
            if train:
                Z, E, L, T = UnrolledLADMM.apply(spec, self.A, x, self.Z0, self.E0, self.L0, *params)
            else:
                Z, E, L, T, _, _ = run_forward(spec, self.A, x, self.Z0, self.E0, self.L0,
                                               [p.detach() for p in params], want_masks=False)
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

    def l1l1_loss(self, x, alpha, layer_weights=None):
        """Fused training objective of the reference drivers (main_syn_l1l1_scalar.py:289-299):

            loss = sum_k w_k * ( alpha * mean_b ||Z_k[:,b]||_1 + mean_b ||x[:,b] - A Z_k[:,b]||_1 )

        with w_k = layer_weights[k] (the scripts use 0.6**epoch for k < K-1 and 1 for the last layer).
        Returns (loss, outputs) where `outputs` is what forward(x) returns, detached.  loss.backward() runs
        the library's backward with the loss cotangents generated inside the kernels."""
        if not x.is_cuda:
            raise RuntimeError("DLADMMNet.l1l1_loss needs a CUDA tensor: d-ladmm_b200 has no CPU path")
        spec, params = self._spec_and_params()
        w = [1.0] * self.layers if layer_weights is None else [float(v) for v in layer_weights]
        if len(w) != self.layers:
            raise ValueError("layer_weights must have one entry per layer")
        loss, Z, E, L, T = UnrolledLADMML1L1.apply(spec, float(alpha), w, self.A, x, self.Z0, self.E0, self.L0, *params)
        return loss, self._as_lists(Z, E, L, T)


    # ---- classical LADMM step, safeguard operator and safeguarded evaluation (family B) -------------------------------
    def two_norm(self, z, dim=0):
        """test_syn_l1l1_scalar.py:126-128"""
        return (z ** 2).sum(dim=dim).sqrt()

    def _one_layer(self, spec1, params, Zk, Ek, Lk, Tk, X):
        Z, E, L, T, _, _ = run_forward(spec1, self.A, X, Zk, Ek, Lk, [p.detach() for p in params], want_masks=False, T_init=Tk)
        return Z[0], E[0], L[0], T[1]

    def _km_spec(self, beta, ss1, ss2, alpha):
        key = (float(beta), float(ss1), float(ss2), float(alpha))
        cache = self.__dict__.setdefault("_km_cache", {})
        if key not in cache:
            dev = self.A.device
            s = lambda v: torch.full((1, 1), float(v), dtype=torch.float32, device=dev)
            params = [self.A.t().contiguous(), s(beta), s(ss1), s(ss2), s(float(ss1) * float(alpha)), s(ss2)]
            slots = {"beta1": 1, "beta2": 1, "beta3": 1, "ss1": 2, "ss2": 3, "theta1": 4, "theta2": 5}
            spec = LayerSpec(_lib.FAMILY_B, self.m, self.d, 1, _lib.PRECISIONS[self.precision], [slots], [0])
            cache.clear()
            cache[key] = (spec, params)
        return cache[key]

    def KM(self, Zk, Ek, Lk, Tk, X, **kwargs):
        """One classical LADMM iteration (test_syn_l1l1_scalar.py:131-160): the learned family-B layer with W = A^T scaled
        by ss1, all betas = beta, thresholds ss1*alpha and ss2.  Returns (Varn, Zn, En, Tn, Ln)."""
        if _FAMILY[self.variant] != _lib.FAMILY_B or self.variant in _NEWS:
            raise NotImplementedError("KM/S/safeguard are built for the family-B variants (scalar, full, tied)")
        beta = float(kwargs.get("beta", 1.0))
        ss1 = kwargs.get("ss1", None)
        ss1 = 0.999 / self.L.item() if ss1 is None else float(ss1)
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
        beta, ss2 = 1.0, 0.3
        _, _, En, Tn, _ = self.KM(Zk, Ek, Lk, Tk, X, beta=beta, ss2=ss2, alpha=alpha)
        c1 = beta * ss2
        out = torch.empty(X.shape[1], dtype=torch.float32, device=X.device)
        stream = torch.cuda.current_stream(X.device).cuda_stream
        _lib.check(_lib.load().dladmm_sg_norm(self.m, X.shape[1], beta, sqrt(c1 / (1 - c1)), Tn.data_ptr(), En.data_ptr(),
                                              Ek.contiguous().data_ptr(), Ep.contiguous().data_ptr(), out.data_ptr(), stream))
        return out

    def forward_safeguarded(self, x, use_learned, use_safeguard, continued=False, K=None, num_iter=200, delta=-99.0,
                            mu_k_method="None", mu_k_param=0.0, alpha=0.01):
        """The evaluation forward of test_syn_l1l1_scalar.py:179-317: per layer the classical KM step, the learned step,
        the safeguard test ||S(u_L2O)|| < (1-delta)*mu_k per column, the per-column selection and the mu_k update.
        Returns (Z, E, L, T) lists, plus sg_count (columns that fell back to KM, per layer) when both flags are set."""
        if _FAMILY[self.variant] != _lib.FAMILY_B or self.variant in _NEWS:
            raise NotImplementedError("forward_safeguarded is built for the family-B variants (scalar, full, tied); the newS "
                                      "ordering has its own safeguard (KM_ELZ / Snorm_ELZ), not built")
        layers = self.layers
        if K is None:
            K = layers if (not continued and (use_learned or use_safeguard)) else num_iter
        X = x.contiguous()
        B = X.shape[1]
        lib = _lib.load()
        spec, params = self._spec_and_params()
        with torch.no_grad():
            T = [self.A.mm(self.Z0) + self.E0 - X]
            Z, E, L = [], [], []
            return_cnt = use_learned and use_safeguard
            if return_cnt:
                mu_k = self._s_norm(self.Z0, self.E0, self.L0, T[-1], X, self.E0, alpha)
                updater = mu_updater_dict[mu_k_method](mu_k, mu_k_param)
                sg_count = [0.0] * layers
            for k in range(K):
                if continued and k == layers:
                    use_learned = use_safeguard = False
                Zp, Ep_, Lp = (self.Z0, self.E0, self.L0) if k == 0 else (Z[-1], E[-1], L[-1])
                _, Zn_KM, En_KM, Tn_KM, Ln_KM = self.KM(Zp, Ep_, Lp, T[-1], X, alpha=alpha)
                if use_learned:
                    spec1 = LayerSpec(spec.family, self.m, self.d, 1, spec.precision, [spec.slots[k]], [spec.weights[k]], spec.fixed)
                    Zn, En, Ln, Tn = self._one_layer(spec1, params, Zp, Ep_, Lp, T[-1], X)
                if use_safeguard:
                    assert use_learned
                    s_norm = self._s_norm(Zn, En, Ln, Tn, X, Ep_, alpha)
                    keep = torch.empty(B, dtype=torch.float32, device=X.device)
                    outs = [torch.empty_like(t) for t in (Zn, En, Tn, Ln)]
                    pairs = (_lib.SgPair * 4)()
                    for i, (a, b_, o) in enumerate(zip((Zn, En, Tn, Ln), (Zn_KM, En_KM, Tn_KM, Ln_KM), outs)):
                        pairs[i].a, pairs[i].b, pairs[i].out, pairs[i].rows = a.data_ptr(), b_.data_ptr(), o.data_ptr(), a.shape[0]
                    mu_vec = mu_k if isinstance(mu_k, torch.Tensor) else torch.full((B,), float(mu_k), device=X.device)
                    _lib.check(lib.dladmm_sg_select(4, pairs, B, s_norm.data_ptr(), mu_vec.contiguous().data_ptr(),
                                                    float(1.0 - delta), keep.data_ptr(),
                                                    torch.cuda.current_stream(X.device).cuda_stream))
                    mu_k = updater.step(s_norm, keep)
                    Z.append(outs[0]); E.append(outs[1]); T.append(outs[2]); L.append(outs[3])
                    if k < layers:
                        sg_count[k] = float(B - keep.sum().item())
                elif use_learned:
                    Z.append(Zn); E.append(En); T.append(Tn); L.append(Ln)
                else:
                    Z.append(Zn_KM); E.append(En_KM); T.append(Tn_KM); L.append(Ln_KM)
        if return_cnt:
            return Z, E, L, T, sg_count
        return Z, E, L, T


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
