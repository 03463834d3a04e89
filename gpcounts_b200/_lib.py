"""ctypes binding of ``libgpcounts_b200.so`` (the C ABI declared in ``include/gpcounts_b200.h``).

There is no CPU fallback: if the shared library is missing or cannot be loaded the import of the
compute path fails loudly.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgpcounts_b200.so")

GPC_VGP, GPC_SVGP, GPC_GPR, GPC_SGPR = 0, 1, 2, 3
GPC_RBF, GPC_CONST, GPC_BRANCH = 0, 1, 2
GPC_NB, GPC_ZINB, GPC_POISSON, GPC_GAUSS = 0, 1, 2, 3
GPC_ST_OK, GPC_ST_CHOL, GPC_ST_OPT, GPC_ST_SKIPPED = 0, 1, 2, 3
GPC_NHYP, GPC_NSTAT, GPC_MAX_RESTART = 4, 16, 10
S_LL, S_STATUS, S_NIT, S_NFEV, S_RESTARTS, S_VAR, S_LS, S_ALPHA, S_KM, S_NFEV_TOTAL, S_GNORM, S_TASK = range(12)


class gpc_model_t(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("kernel", C.c_int32), ("lik", C.c_int32), ("N", C.c_int32), ("M", C.c_int32),
        ("d", C.c_int32), ("n_off", C.c_int32), ("train_mask", C.c_int32), ("n_zsets", C.c_int32),
        ("handoff_alpha", C.c_int32), ("xp", C.c_double), ("jitter", C.c_double),
        ("X", C.c_void_p), ("Z", C.c_void_p), ("Zgene", C.c_void_p), ("restart", C.c_void_p),
    ]


class gpc_opt_t(C.Structure):
    _fields_ = [("m", C.c_int32), ("maxiter", C.c_int32), ("maxfun", C.c_int32), ("maxls", C.c_int32),
                ("ftol", C.c_double), ("gtol", C.c_double),
                ("variant", C.c_int32), ("trace_cap", C.c_int32), ("trace_gene", C.c_int32), ("trace_model", C.c_int32),
                ("trace", C.c_void_p), ("attempt0", C.c_void_p)]


EXPORTS = ["gpc_default_opt", "gpc_param_count", "gpc_default_slots", "gpc_workspace_bytes", "gpc_elbo_grad",
           "gpc_fit", "gpc_predict", "gpc_predict_workspace_bytes", "gpc_dbg_linalg", "gpc_dbg_linalg_scratch_bytes", "gpc_launch_count", "gpc_last_error_string", "gpc_version"]

_lib = None


def load():
    """Load the library (once) and declare the signatures of every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run `python -c \"import __graft_entry__ as g; g.build()\"` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.gpc_default_opt.argtypes = [P(gpc_opt_t)]
    lib.gpc_default_opt.restype = None
    lib.gpc_param_count.argtypes = [P(gpc_model_t)]
    lib.gpc_param_count.restype = C.c_int64
    lib.gpc_default_slots.argtypes = [P(gpc_model_t), C.c_int32]
    lib.gpc_default_slots.restype = C.c_int32
    lib.gpc_workspace_bytes.argtypes = [P(gpc_model_t), C.c_int32, P(gpc_opt_t), C.c_int32]
    lib.gpc_workspace_bytes.restype = C.c_size_t
    lib.gpc_elbo_grad.argtypes = [P(gpc_model_t), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_int32, C.c_void_p]
    lib.gpc_elbo_grad.restype = C.c_int
    lib.gpc_fit.argtypes = [P(gpc_model_t), C.c_int32, C.c_int32, P(gpc_opt_t), C.c_int32, C.c_void_p, C.c_void_p,
                            C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t,
                            C.c_int32, C.c_void_p]
    lib.gpc_fit.restype = C.c_int
    lib.gpc_predict_workspace_bytes.argtypes = [P(gpc_model_t), C.c_int32, C.c_int32]
    lib.gpc_predict_workspace_bytes.restype = C.c_size_t
    lib.gpc_predict.argtypes = [P(gpc_model_t), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]
    lib.gpc_predict.restype = C.c_int
    lib.gpc_dbg_linalg.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]
    lib.gpc_dbg_linalg.restype = C.c_int
    lib.gpc_dbg_linalg_scratch_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.gpc_dbg_linalg_scratch_bytes.restype = C.c_size_t
    lib.gpc_launch_count.argtypes = []
    lib.gpc_launch_count.restype = C.c_int64
    lib.gpc_last_error_string.argtypes = []
    lib.gpc_last_error_string.restype = C.c_char_p
    lib.gpc_version.argtypes = []
    lib.gpc_version.restype = C.c_char_p
    _lib = lib
    return lib


class GpcError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise GpcError(f"gpcounts_b200 error {rc}: {load().gpc_last_error_string().decode()}")
