"""ctypes binding of libdladmm.so (include/dladmm.h).  No torch types cross this boundary: only raw
device pointers, sizes and a cudaStream_t handle."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# (DLADMM_LIB_PATH: an alternative build of the same library, for A/B measurements of compile-time switches)
LIB_PATH = os.environ.get("DLADMM_LIB_PATH") or os.path.join(HERE, "csrc", "libdladmm.so")

ABI_VERSION = 7
FAMILY_A, FAMILY_B, FAMILY_C = 0, 1, 2
PREC_FP32, PREC_TF32X3, PREC_TF32, PREC_BF16, PREC_TF32_BF16X2 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": PREC_FP32, "tf32x3": PREC_TF32X3, "tf32": PREC_TF32, "bf16": PREC_BF16, "tf32_bf16x2": PREC_TF32_BF16X2}
# dladmm_metric: index of each per-layer metric in the (K, MET_COUNT) output of a forward with `metrics`
METRICS = ("l1_z", "sqerr_z", "l1_res", "sq_res", "sqerr_e", "sqerr_az", "l1_e", "dot_lx", "dgap_l", "dgap_atl")
MET_COUNT = len(METRICS)

c_float_p = C.c_void_p   # device pointers are opaque to the host


class BParam(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("grad", C.c_void_p), ("row_stride", C.c_int32), ("col_period", C.c_int32)]


class Layer(C.Structure):
    _fields_ = [("beta1", BParam), ("beta2", BParam), ("beta3", BParam), ("ss1", BParam), ("ss2", BParam),
                ("ss2_2", BParam), ("theta1", BParam), ("theta2", BParam), ("W", C.c_void_p), ("gW", C.c_void_p)]


class Problem(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("family", C.c_int32), ("precision", C.c_int32),
                ("m", C.c_int32), ("d", C.c_int32), ("K", C.c_int32), ("B", C.c_int64),
                ("last_only", C.c_int32), ("reserved", C.c_int32),
                ("A", C.c_void_p), ("X", C.c_void_p), ("Z0", C.c_void_p), ("E0", C.c_void_p), ("L0", C.c_void_p),
                ("layers", C.POINTER(Layer)),
                ("Z", C.c_void_p), ("E", C.c_void_p), ("L", C.c_void_p), ("T", C.c_void_p),
                ("maskZ", C.c_void_p), ("maskE", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("T_init", C.c_void_p),
                ("Vsave", C.c_void_p), ("objective", C.c_void_p), ("objective_alpha", C.c_float), ("objective_kind", C.c_int32),
                ("start_half", C.c_int32), ("stop_half", C.c_int32), ("metrics", C.c_void_p)]


class Metrics(C.Structure):
    _fields_ = [("want", C.c_uint32), ("dual_alpha", C.c_float), ("Z_label", C.c_void_p), ("E_label", C.c_void_p),
                ("X_clean", C.c_void_p), ("out", C.c_void_p)]


class Cotangents(C.Structure):
    _fields_ = [("gZ", C.c_void_p), ("gE", C.c_void_p), ("gL", C.c_void_p), ("gT", C.c_void_p),
                ("loss_kind", C.c_int32), ("loss_alpha", C.c_float), ("loss_layer_weight", C.POINTER(C.c_float)),
                ("loss_scale", C.c_void_p), ("layer_events", C.POINTER(C.c_void_p))]


class Caps(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("sm_count", C.c_int32), ("supported", C.c_int32), ("has_tcgen05", C.c_int32),
                ("total_mem", C.c_int64)]


class GenDesc(C.Structure):
    _fields_ = [("m", C.c_int32), ("d", C.c_int32), ("B", C.c_int64), ("col_offset", C.c_int64),
                ("seed", C.c_uint64), ("p", C.c_float), ("mu", C.c_float), ("sigma", C.c_float),
                ("dense_noise", C.c_int32), ("sigma_e", C.c_float), ("generate_A", C.c_int32),
                ("A", C.c_void_p), ("Zs", C.c_void_p), ("Es", C.c_void_p), ("X", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("amplitude", C.c_int32), ("reserved", C.c_int32)]


class SgPair(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("out", C.c_void_p), ("rows", C.c_int32)]


EXPORTS = ["dladmm_sg_norm", "dladmm_sg_select", "dladmm_sg_norm_elz", "dladmm_sg_select_update", "dladmm_workspace_bytes", "dladmm_forward", "dladmm_backward", "dladmm_gen_workspace_bytes",
           "dladmm_gen_syn", "dladmm_objective", "dladmm_query", "dladmm_last_error", "dladmm_launch_count",
           "dladmm_profile_start", "dladmm_profile_stop", "dladmm_debug_trace"]

KIND_NAMES = ["prep", "gemm_t0", "gemm_z", "gemm_elt", "bwd_elem", "bwd_gemm_dz", "bwd_gemm_dw", "bwd_gemm_dv",
              "bwd_reduce", "gen", "objective", "metric_gemm", "safeguard", "fwd_persistent"]

_lib = None


def load():
    """Load the in-tree shared library; there is no fallback if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libdladmm.so not found at %s -- build it with `python d-ladmm_b200/build.py` "
            "(or __graft_entry__.build()); this package has no CPU/PyTorch fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.dladmm_workspace_bytes.restype = C.c_size_t
    lib.dladmm_workspace_bytes.argtypes = [C.POINTER(Problem), C.c_int]
    lib.dladmm_forward.restype = C.c_int
    lib.dladmm_forward.argtypes = [C.POINTER(Problem), C.c_void_p]
    lib.dladmm_backward.restype = C.c_int
    lib.dladmm_backward.argtypes = [C.POINTER(Problem), C.POINTER(Cotangents), C.c_void_p]
    lib.dladmm_gen_workspace_bytes.restype = C.c_size_t
    lib.dladmm_gen_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.dladmm_gen_syn.restype = C.c_int
    lib.dladmm_gen_syn.argtypes = [C.POINTER(GenDesc), C.c_void_p]
    lib.dladmm_objective.restype = C.c_int
    lib.dladmm_objective.argtypes = [C.POINTER(Problem), C.c_float, C.c_void_p, C.c_void_p]
    lib.dladmm_sg_norm.restype = C.c_int
    lib.dladmm_sg_norm.argtypes = [C.c_int32, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
    lib.dladmm_sg_select.restype = C.c_int
    lib.dladmm_sg_select.argtypes = [C.c_int32, C.POINTER(SgPair), C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                     C.c_void_p]
    lib.dladmm_sg_norm_elz.restype = C.c_int
    lib.dladmm_sg_norm_elz.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dladmm_sg_select_update.restype = C.c_int
    lib.dladmm_sg_select_update.argtypes = [C.c_int32, C.POINTER(SgPair), C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_int32,
                                            C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dladmm_query.restype = C.c_int
    lib.dladmm_query.argtypes = [C.c_int, C.POINTER(Caps)]
    lib.dladmm_launch_count.restype = C.c_int64
    lib.dladmm_launch_count.argtypes = []
    lib.dladmm_profile_start.restype = C.c_int
    lib.dladmm_profile_start.argtypes = []
    lib.dladmm_profile_stop.restype = C.c_int
    lib.dladmm_profile_stop.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.dladmm_debug_trace.restype = C.c_int64
    lib.dladmm_debug_trace.argtypes = [C.POINTER(C.c_int64), C.c_int64]
    lib.dladmm_last_error.restype = C.c_char_p
    lib.dladmm_last_error.argtypes = []
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().dladmm_last_error().decode("utf-8", "replace")
        raise RuntimeError("libdladmm error %d: %s" % (rc, msg))


def launch_count():
    return int(load().dladmm_launch_count())


def profile_start():
    check(load().dladmm_profile_start())


def profile_stop():
    """-> dict kind name -> (milliseconds summed over launches, launches)"""
    n = len(KIND_NAMES)
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    check(load().dladmm_profile_stop(ms, cnt))
    return {KIND_NAMES[i]: (ms[i], int(cnt[i])) for i in range(n)}
