"""Load the reference's own ``DLADMMNet`` classes from /root/reference (build container only).

TEST INFRASTRUCTURE -- not product code.  Used by ``oracle/make_golden.py``, by the CPU
tests that pin ``oracle/dladmm_oracle.py`` to the reference, and by ``bench.py``'s reference legs.
``/root/reference`` does not exist on the GPU box: there the loader finds the byte copies that
``oracle/fetch_ref.py`` staged under ``oracle/_ref/`` (git-ignored), if any.

The reference scripts hard-code ``.cuda()`` in ``DLADMMNet.__init__`` (e.g.
main_syn_l1l1_scalar.py:40-44) and most of them run training at import time, so the class is
extracted by AST (``ClassDef DLADMMNet``) and executed in a namespace holding only the names
the class body needs.  ``Tensor.cuda`` / ``Module.cuda`` are patched to identity while the
class is constructed so that it runs on CPU, unmodified.
"""
import ast
import contextlib
import os
from math import sqrt

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

def _find_root():
    """$DLADMM_REFERENCE_ROOT, else /root/reference (build container), else oracle/_ref (the files oracle/fetch_ref.py staged,
    which travel to the GPU box)."""
    env = os.environ.get("DLADMM_REFERENCE_ROOT")
    if env:
        return env
    staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if not os.path.isfile("/root/reference/main_syn_l1l1_scalar.py") and os.path.isfile(os.path.join(staged, "main_syn_l1l1_scalar.py")):
        return staged
    return "/root/reference"


REFERENCE_ROOT = _find_root()

# variant name -> reference script that declares the class
VARIANT_FILES = {
    "lena": "main_lena.py",                       # family A, (m x bs) betas, fixed thresholds
    "ltheta": "main_syn_l1l1_ltheta.py",          # family A, (m,1) betas, learnable thresholds
    "scalar": "main_syn_l1l1_scalar.py",          # family B, (1,1) params
    "full": "main_syn_l1l1_full.py",              # family B, per-row params
    "tied": "main_syn_l1l1_scalar_tied.py",       # family B, one shared fc + ss1
    "lasso": "main_syn_lasso_scalar.py",          # family C, linear E-step
    # E -> L -> Z ordering with prefix execution forward(x, K)
    "newS": "main_syn_scalar_newS_layerwise.py",
    "tied_newS": "main_syn_scalar_tied_newS_layerwise.py",
    "ptied_newS": "main_syn_scalar_ptied_newS_layerwise.py",      # constructor takes `interval`
}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, VARIANT_FILES["scalar"]))


@contextlib.contextmanager
def cuda_is_identity():
    """Patch ``.cuda()`` to a no-op so the reference's constructors run on CPU."""
    t_cuda, m_cuda = torch.Tensor.cuda, nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, nn.Module.cuda = t_cuda, m_cuda


def load_class(variant, alpha=0.001, script=None):
    """Return the reference ``DLADMMNet`` class object for ``variant`` (unmodified source)."""
    path = os.path.join(REFERENCE_ROOT, script or VARIANT_FILES[variant])
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    nodes = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DLADMMNet"]
    if not nodes:
        raise RuntimeError("no DLADMMNet class in %s" % path)
    module = ast.Module(body=nodes, type_ignores=[])
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "sqrt": sqrt, "alpha": alpha,
          "__name__": "reference_" + variant}
    exec(compile(module, path, "exec"), ns)
    return ns["DLADMMNet"]


def build(variant, m, n, d, batch_size, A, Z0, E0, L0, layers, alpha=0.001, **extra):
    """Instantiate the reference model on CPU with its own default initialisation (`extra`: e.g. interval=...)."""
    cls = load_class(variant, alpha=alpha)
    with cuda_is_identity():
        model = cls(m=m, n=n, d=d, batch_size=batch_size, A=A, Z0=Z0, E0=E0, L0=L0, layers=layers, **extra)
    return model


def run(model, x):
    """Forward through the reference; always returns (Z, E, L, T-or-None) lists."""
    with cuda_is_identity():
        out = model(x)
    if len(out) == 4:
        return out
    Z, E, L = out
    return Z, E, L, None


def load_mu_updater_dict():
    """The reference's own mu_updater.py (numpy-only import), loaded as a module object."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_mu_updater", os.path.join(REFERENCE_ROOT, "mu_updater.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.mu_updater_dict


def load_eval_class(layers, alpha=0.01, delta=-99.0, mu_k_method="None", mu_k_param=0.0, continued=False, num_iter=200,
                    use_learned=True, use_safeguard=True, script="test_syn_l1l1_scalar.py"):
    """The evaluation ``DLADMMNet`` of test_syn_l1l1_scalar.py:75-317 (KM, S, safeguarded forward), unmodified.

    The class reads script-level globals (test_syn_l1l1_scalar.py:39-55: alpha, delta, mu_k_method, mu_k_param,
    layers, K, args.continued, mu_updater_dict); they are supplied through the exec namespace."""
    import types
    path = os.path.join(REFERENCE_ROOT, script)       # e.g. test_syn_l1l1_newS_Acols.py: the E -> L -> Z evaluation class
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    nodes = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DLADMMNet"]
    K = layers if (not continued and (use_learned or use_safeguard)) else num_iter
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "sqrt": sqrt, "alpha": alpha, "delta": delta,
          "mu_k_method": mu_k_method, "mu_k_param": mu_k_param, "layers": layers, "K": K,
          "args": types.SimpleNamespace(continued=continued), "mu_updater_dict": load_mu_updater_dict(),
          "__name__": "reference_eval_scalar"}
    exec(compile(ast.Module(body=nodes, type_ignores=[]), path, "exec"), ns)
    return ns["DLADMMNet"]
