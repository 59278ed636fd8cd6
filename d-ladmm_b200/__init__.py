"""d-ladmm_b200: B200-native (sm_100a) implementation of the D-LADMM unrolled forward/backward hot path
behind the reference's own ``DLADMMNet`` nn.Module interface (xhchrn/D-LADMM).

Import as ``dladmm_b200`` (the repo root carries a one-line shim because the directory name contains a
hyphen).  Everything that computes goes through libdladmm.so (include/dladmm.h); nothing here falls back
to PyTorch eager or to the CPU.
"""
from . import _lib
from .net import (DLADMMNet, DLADMMNetFull, DLADMMNetLasso, DLADMMNetLena, DLADMMNetLtheta, DLADMMNetScalar,
                  DLADMMNetTied, DLADMMNetNewS, DLADMMNetTiedNewS, DLADMMNetPtiedNewS, VARIANT_CLASSES, default_precision, resolve_precision)
from .function import UnrolledLADMM, LayerSpec, run_forward
from .gen_syn import gen_syn_data, SynData, replace_A_columns, load_mat, save_mat
from .objective import l1l1_objective
from .mu_updater import mu_updater_dict
from .sharding import column_shard, allreduce_gradients, ShardedTrainer
from .host_feed import HostFeed, HostDrain

__all__ = ["DLADMMNet", "DLADMMNetScalar", "DLADMMNetFull", "DLADMMNetTied", "DLADMMNetLasso", "DLADMMNetLena",
           "DLADMMNetLtheta", "DLADMMNetNewS", "DLADMMNetTiedNewS", "DLADMMNetPtiedNewS", "VARIANT_CLASSES", "UnrolledLADMM", "LayerSpec", "run_forward", "gen_syn_data",
           "SynData", "l1l1_objective", "column_shard", "allreduce_gradients", "ShardedTrainer",
           "default_precision", "resolve_precision", "library_path", "query_device", "mu_updater_dict", "HostFeed", "HostDrain", "replace_A_columns", "load_mat", "save_mat"]


def library_path():
    return _lib.LIB_PATH


def query_device(device=0):
    """dladmm_query: compute capability, SM count and whether this build has the tcgen05 kernels."""
    import ctypes as C
    caps = _lib.Caps()
    _lib.check(_lib.load().dladmm_query(int(device), C.byref(caps)))
    return {f[0]: getattr(caps, f[0]) for f in _lib.Caps._fields_}
