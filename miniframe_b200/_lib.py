"""ctypes binding of libminiframe_b200.so (C ABI: include/miniframe_b200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry
point is called without a CUDA device, this raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64,
                    c_size_t, c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libminiframe_b200.so")

MF_OK = 0
COV_LOWER, COV_FULL, COV_LOWER_INTERLEAVED = 0, 1, 2
FRAMEWORK_BIGGP, FRAMEWORK_BIGGP_JITT, FRAMEWORK_SMALLGP = 0, 1, 2
KERNEL_NONE, KERNEL_SE, KERNEL_QP = -1, 0, 1
KERNEL_IDS = {"SquaredExponential": KERNEL_SE, "QuasiPeriodic": KERNEL_QP}
KERNEL_NPAR = {KERNEL_SE: 2, KERNEL_QP: 4}
MAX_SERIES = 8


class Problem(Structure):
    """mf_problem_t"""
    _fields_ = [("framework", c_int32), ("kernel", c_int32), ("extra_kernel", c_int32),
                ("n_series", c_int32), ("n_epochs", c_int32), ("hyper_len", c_int32),
                ("extra_len", c_int32), ("reserved", c_int32), ("nugget", c_double)]


# every symbol include/miniframe_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mf_version": (c_int, []),
    "mf_status_string": (c_char_p, [c_int]),
    "mf_create": (c_int, [POINTER(c_void_p), c_int]),
    "mf_destroy": (c_int, [c_void_p]),
    "mf_last_error": (c_char_p, [c_void_p]),
    "mf_device_info": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "mf_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "mf_ld": (c_int64, [c_int]),
    "mf_matrix_stride": (c_int64, [c_int]),
    "mf_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mf_build_cov": (c_int, [c_void_p, POINTER(Problem), c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p]),
    "mf_interleave_resid": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_void_p,
                                    c_int64, c_void_p]),
    "mf_potrf_loglike": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_int64, c_void_p,
                                 c_int64, c_void_p, c_void_p, c_void_p]),
    "mf_loglike": (c_int, [c_void_p, POINTER(Problem), c_int, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p, c_void_p,
                           c_void_p]),
    "mf_solve_logdet": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_int64, c_void_p,
                                c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mf_kepler_rv": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int,
                             c_void_p]),
    "mf_grad_ctas": (c_int, [c_int]),
    "mf_grad_contract": (c_int, [c_void_p, POINTER(Problem), c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "mf_probe_fp64": (c_int, [c_void_p, c_int, c_int, POINTER(c_float), POINTER(c_double),
                              c_void_p]),
}

_lib = None


def load():
    """dlopen the library and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "miniframe_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C miniframe_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class MiniframeError(RuntimeError):
    pass


def check(lib, handle, status, what):
    if status != MF_OK:
        detail = lib.mf_last_error(handle).decode() if handle else ""
        raise MiniframeError("%s failed: %s%s" % (what, lib.mf_status_string(status).decode(),
                                                  (" (" + detail + ")") if detail else ""))
