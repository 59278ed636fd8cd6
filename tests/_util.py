"""Shared helpers for the tests: golden fixtures, model construction, error metrics."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_ALL_NPZ = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
NEWS_GOLDEN_NAMES = [n for n in _ALL_NPZ if "newS" in n]             # E -> L -> Z ordering fixtures (tests/test_newS.py)
ELZ_GOLDEN_NAMES = [n for n in _ALL_NPZ if n.startswith("elz_")]     # newS-ordering safeguarded evaluation (oracle only, so far)
GOLDEN_NAMES = [n for n in _ALL_NPZ if not n.startswith(("sg_", "elz_")) and n != "lena_psnr" and "newS" not in n]   # lena_psnr: tests/test_lena_psnr.py
SG_GOLDEN_NAMES = [n for n in _ALL_NPZ if n.startswith("sg_")]      # safeguarded evaluation fixtures
SMALL_GOLDEN = [n for n in GOLDEN_NAMES if "shape" not in n]


class Golden(object):
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.variant = str(z["variant"])
        self.K = int(z["K"])
        self.bs = int(z["bs"])
        self.returns_T = bool(z["returns_T"])
        t = lambda k: torch.from_numpy(z[k].copy())
        self.A, self.X, self.Z0, self.E0, self.L0 = t("A"), t("X"), t("Z0"), t("E0"), t("L0")
        self.Z, self.E, self.L = t("Z"), t("E"), t("L")
        self.T = t("T") if "T" in z.files else None
        self.cz, self.ce, self.cl, self.ct = t("cz"), t("ce"), t("cl"), t("ct")
        self.loss = float(z["loss"])
        self.keys = [str(k) for k in z["keys"]]
        self.sd = {k: t("sd/" + k) for k in self.keys}
        self.grads = {k: t("grad/" + k) for k in self.keys}
        self.m, self.d = self.A.shape


def rel_l2(a, b, floor=0.0):
    a, b = a.double(), b.double()
    return ((a - b).norm() / max(b.norm().item(), floor, 1e-300)).item()


def build_model(g, device, precision=None):
    import dladmm_b200 as dl
    cls = dl.VARIANT_CLASSES[g.variant]
    model = cls(m=g.m, n=10000, d=g.d, batch_size=g.bs, A=g.A, Z0=g.Z0, E0=g.E0, L0=g.L0, layers=g.K,
                precision=precision, device=device)
    model.load_state_dict(g.sd)
    return model


def syn(m, d, B, seed, p=0.1, sigma=1.0):
    gen = torch.Generator().manual_seed(seed)
    A = torch.randn(m, d, generator=gen)
    A = A / A.pow(2).sum(dim=0, keepdim=True).sqrt()
    Z = (torch.rand(d, B, generator=gen) < p).float() * torch.randn(d, B, generator=gen) * sigma
    E = (torch.rand(m, B, generator=gen) < p).float() * torch.randn(m, B, generator=gen) * sigma
    return A, A.mm(Z) + E


class NewSGolden(object):
    """A fixture written by oracle/make_golden_newS.py (reference newS / tied_newS / ptied_newS outputs)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.variant, self.ref_name = str(z["variant"]), str(z["name"])
        self.layers, self.K_run, self.bs, self.interval = int(z["layers"]), int(z["K_run"]), int(z["bs"]), int(z["interval"])
        t = lambda k: torch.from_numpy(z[k].copy())
        self.A, self.X, self.Z0, self.E0, self.L0 = t("A"), t("X"), t("Z0"), t("E0"), t("L0")
        self.Z, self.E, self.L = t("Z"), t("E"), t("L")
        self.cz, self.ce, self.cl = t("cz"), t("ce"), t("cl")
        self.loss = float(z["loss"])
        self.keys = [str(k) for k in z["keys"]]
        self.sd = {k: t("sd/" + k) for k in self.keys}
        self.grads = {k: t("grad/" + k) for k in self.keys}
        self.nograd = [str(k) for k in z["nograd"] if str(k)]
        self.m, self.d = self.A.shape

    def build(self, device, precision=None):
        import dladmm_b200 as dl
        extra = {"interval": self.interval} if self.interval else {}
        model = dl.VARIANT_CLASSES[self.variant](m=self.m, n=10000, d=self.d, batch_size=self.bs, A=self.A, Z0=self.Z0,
                                                 E0=self.E0, L0=self.L0, layers=self.layers, precision=precision,
                                                 device=device, **extra)
        model.load_state_dict(self.sd)
        return model


class SgGolden(object):
    """A fixture written by oracle/make_golden_safeguard.py (reference test_syn_l1l1_scalar.py outputs)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        t = lambda k: torch.from_numpy(z[k].copy())
        self.A, self.X = t("A"), t("X")
        self.m, self.d = self.A.shape
        self.B = self.X.shape[1]
        self.lip, self.layers, self.alpha = float(z["lip"]), int(z["layers"]), float(z["alpha"])
        self.use_learned, self.use_safeguard = bool(z["use_learned"]), bool(z["use_safeguard"])
        self.continued, self.num_iter = bool(z["continued"]), int(z["num_iter"])
        self.delta, self.method, self.param = float(z["delta"]), str(z["method"]), float(z["param"])
        self.Z, self.E, self.L, self.T = t("Z"), t("E"), t("L"), t("T")
        self.sg_count = z["sg_count"].tolist()
        self.keys = [str(k) for k in z["keys"]]
        self.sd = {k: t("sd/" + k) for k in self.keys}
        self.Z0, self.E0, self.L0 = torch.zeros(self.d, self.B), torch.zeros(self.m, self.B), torch.zeros(self.m, self.B)
        if "s_norm" in z.files:
            s, thr = t("s_norm"), t("thr")
            margin = ((s - thr).abs() / thr.abs().clamp_min(1e-9)).min(dim=0).values
        else:
            margin = torch.full((self.B,), float("inf"))
        self.margin = margin        # per column: smallest relative distance of ||S|| from the threshold over the layers

    def robust_columns(self, tol):
        return self.margin > tol

    def kwargs(self):
        return dict(continued=self.continued, num_iter=self.num_iter, delta=self.delta, mu_k_method=self.method,
                    mu_k_param=self.param, alpha=self.alpha)


class ElzGolden(object):
    """A fixture written by oracle/make_golden_safeguard_newS.py (reference test_syn_l1l1_newS_Acols.py outputs):
    Z/E/L hold K+1 entries (initial variables first), there is no T list."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        t = lambda k: torch.from_numpy(z[k].copy())
        self.A, self.X = t("A"), t("X")
        self.m, self.d = self.A.shape
        self.B = self.X.shape[1]
        self.lip, self.layers, self.alpha = float(z["lip"]), int(z["layers"]), float(z["alpha"])
        self.use_learned, self.use_safeguard = bool(z["use_learned"]), bool(z["use_safeguard"])
        self.continued, self.num_iter = bool(z["continued"]), int(z["num_iter"])
        self.delta, self.method, self.param = float(z["delta"]), str(z["method"]), float(z["param"])
        self.Z, self.E, self.L = t("Z"), t("E"), t("L")
        self.sg_count = z["sg_count"].tolist()
        self.keys = [str(k) for k in z["keys"]]
        self.sd = {k: t("sd/" + k) for k in self.keys}
        self.Z0, self.E0, self.L0 = torch.zeros(self.d, self.B), torch.zeros(self.m, self.B), torch.zeros(self.m, self.B)
        if "s_norm" in z.files:
            s, thr = t("s_norm"), t("thr")
            self.margin = ((s - thr).abs() / thr.abs().clamp_min(1e-9)).min(dim=0).values
        else:
            self.margin = torch.full((self.B,), float("inf"))

    def robust_columns(self, tol):
        return self.margin > tol

    def kwargs(self):
        return dict(continued=self.continued, num_iter=self.num_iter, delta=self.delta, mu_k_method=self.method,
                    mu_k_param=self.param, alpha=self.alpha)
