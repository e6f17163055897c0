"""ctypes binding of libipm_b200.so (include/ipm_b200.h).  No fallback: a missing library or a missing GPU
raises — the product path never routes through a CPU implementation."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libipm_b200.so")

IPM_OK = 0
ERRORS = {-1: "IPM_ERR_CUDA", -2: "IPM_ERR_ARG", -3: "IPM_ERR_SHAPE", -4: "IPM_ERR_STATE", -5: "IPM_ERR_NOMEM"}
STATUS = {0: "converged", 1: "max_iter", 2: "nan"}
BOPT_REFINE, BOPT_STRIP_TMA, BOPT_HANDOFF, BOPT_SYRK_RHS = 1, 2, 3, 4      # ipm_batched_set_option
REFRESH_DEFAULT = 12     # ipm_batched_set_variant: default period of the from-scratch residual check

# every symbol include/ipm_b200.h declares: name -> (restype, argtypes)
_dp = POINTER(c_double)
_ip = POINTER(c_int)
_i32p = POINTER(c_int32)
SYMBOLS = {
    "ipm_create": (c_int, [POINTER(c_void_p), c_int]),
    "ipm_destroy": (None, [c_void_p]),
    "ipm_last_error": (c_char_p, [c_void_p]),
    "ipm_version": (c_char_p, []),
    "ipm_launch_count": (c_int64, []),
    "ipm_set_pivot_threshold": (c_int, [c_void_p, c_double]),
    "ipm_load_csr": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ipm_load_csc": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ipm_set_ingest_mode": (c_int, [c_int, c_int]),
    "ipm_pattern_info": (c_int, [c_void_p, c_void_p]),
    "ipm_get_pattern": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "ipm_get_values": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ipm_pattern_cache_stats": (c_int, [c_void_p]),
    "ipm_pattern_cache_clear": (c_int, []),
    "ipm_load_dense": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "ipm_load_dense_d": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "ipm_init_state": (c_int, [c_void_p, c_int]),
    "ipm_set_state": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ipm_get_state": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "ipm_residual_norms": (c_int, [c_void_p, c_void_p]),
    "ipm_get_residuals": (c_int, [c_void_p, c_void_p, c_void_p]),
    "ipm_assemble_normal": (c_int, [c_void_p]),
    "ipm_get_M": (c_int, [c_void_p, c_void_p]),
    "ipm_factor": (c_int, [c_void_p, c_double, _ip]),
    "ipm_direction": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ipm_ratio_test": (c_int, [c_void_p, c_int, c_double, c_void_p]),
    "ipm_sigma": (c_int, [c_void_p, c_void_p]),
    "ipm_update": (c_int, [c_void_p, c_double, c_double]),
    "ipm_op_ratio_test": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p]),
    "ipm_op_step_size_bounded": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                                 c_void_p]),
    "ipm_op_sigma": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ipm_op_update": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                              c_double]),
    "ipm_solve_spd": (c_int, [c_int, c_int, c_void_p, c_void_p, c_double, c_void_p, _ip]),
    "ipm_start_mehrotra": (c_int, [c_void_p]),
    "ipm_detect_dependent_rows": (c_int, [c_void_p, c_double, _ip]),
    "ipm_set_refinement": (c_int, [c_void_p, c_double]),
    "ipm_solve": (c_int, [c_void_p, c_double, c_int, c_int, c_void_p, c_void_p, c_void_p, _dp, _ip, _ip, c_void_p]),
    "ipm_solve_batched_dense": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_double, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "ipm_solve_batched_dense_d": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_double, c_int,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, _ip]),
    "ipm_release_cached": (c_int, []),
    "ipm_batched_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "ipm_batched_set_variant": (c_int, [c_int, c_int]),
    "ipm_batched_set_option": (c_int, [c_int, c_int]),
    "ipm_batched_last_handoffs": (c_int, []),
    "ipm_set_kkt_cluster": (c_int, [c_int]),
    "ipm_kkt_last_profile": (c_int, [c_void_p]),
    "ipm_solve_dense_kkt": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_double, c_int, c_void_p, c_void_p,
                                    c_void_p, _dp, _ip, _ip]),
    "ipm_profile_enable": (c_int, [c_int]),
    "ipm_profile_read": (c_int, [c_void_p, c_void_p, POINTER(c_int64)]),
    "ipm_profile_last": (c_int, [c_void_p, c_void_p, c_int]),
    "ipm_measure_dmma_peak": (c_double, [c_int]),
    "ipm_set_syrk_stage_width": (c_int, [c_int]),
    "ipm_set_chol_fused_diag": (c_int, [c_int]),
    "ipm_set_small_lp_fused": (c_int, [c_int]),
    "ipm_set_syrk_consumers": (c_int, [c_int]),
    "ipm_syrk_d": (c_int, [c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int64]),
    "ipm_potrf_d": (c_int, [c_int, c_int, c_void_p, c_int64, c_double, _ip]),
    "ipm_syrk_batched_d": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64]),
    "ipm_potrf_batched_d": (c_int, [c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_double, _ip]),
}

_lib = None


class IpmError(RuntimeError):
    pass


def load():
    """Load the C-ABI library; raises ImportError with the build recipe when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libipm_b200.so is not built (%s). Build it with `python -m interiorpointmethod_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, handle=None, what: str = ""):
    if rc == IPM_OK:
        return
    lib = load()
    msg = lib.ipm_last_error(handle)
    raise IpmError("%s failed: %s (%d): %s" % (what or "ipm call", ERRORS.get(rc, "?"), rc,
                                                msg.decode() if msg else ""))
