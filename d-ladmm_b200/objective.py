"""Per-layer L1-L1 objective from the iterates a forward left on the device (no extra A*Z product).

Replaces the script-side loop main_syn_l1l1_scalar.py:333-334
    l1l1_values[jj] += alpha*sum|Z[jj]| + sum|x - A Z[jj]|
using the identity x - A Z_k = E_k - T_{k+1}."""
import ctypes as C

import torch

from . import _lib


def _as_stack(lst):
    if isinstance(lst, torch.Tensor):
        return lst.contiguous()
    base = getattr(lst[0], "_base", None)
    if base is not None and base.dim() == 3 and base.shape[0] == len(lst) and base.is_contiguous() and all(
            getattr(t, "_base", None) is base for t in lst) and lst[0].data_ptr() == base.data_ptr():
        return base
    return torch.stack(list(lst))


def l1l1_objective(Z, E, T, alpha):
    """Z: list/stack (K,d,B); E: (K,m,B); T: (K+1,m,B) as returned by DLADMMNet.forward.  Returns a (K,)
    tensor: sum over the batch of alpha*||Z_k||_1 + ||x - A Z_k||_1."""
    lib = _lib.load()
    Zs, Es, Ts = _as_stack(Z), _as_stack(E), _as_stack(T)
    K, d, B = Zs.shape
    m = Es.shape[1]
    if Ts.shape[0] != K + 1:
        raise RuntimeError("T must hold K+1 slabs")
    for t in (Zs, Es, Ts):
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError("l1l1_objective needs float32 CUDA tensors")
    p = _lib.Problem()
    p.abi_version = _lib.ABI_VERSION
    p.m, p.d, p.K, p.B = m, d, K, B
    p.Z, p.E, p.T = Zs.data_ptr(), Es.data_ptr(), Ts.data_ptr()
    out = torch.empty(K, dtype=torch.float32, device=Zs.device)
    with torch.cuda.device(Zs.device):
        _lib.check(lib.dladmm_objective(C.byref(p), float(alpha), out.data_ptr(),
                                        torch.cuda.current_stream(Zs.device).cuda_stream))
    return out
