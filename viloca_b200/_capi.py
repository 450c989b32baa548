"""ctypes binding of ``include/viloca_b200.h`` (the C ABI of the CUDA engine).

There is no CPU fallback: if the shared library is missing the import of this module raises,
and every product entry point that needs the engine fails with it.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ABI_VERSION = 2
OPT_SKIP_DEAD = 1
OPT_FUSED_CHUNKS = 3
N_BASES = 5
CODE_N = 255
MAX_CLUSTERS = 128
MODE_UQS, MODE_LEP = 0, 1
STATUS_RUNNING, STATUS_CONVERGED, STATUS_DECREASING, STATUS_NAN, STATUS_MAXITER = 0, 1, 2, 3, 4
INIT_STRIDE = 8
SCALAR_STRIDE = 16
(S_ELBO, S_GAMMA_A, S_GAMMA_B, S_MLG0, S_MLG1, S_THETA_C, S_THETA_D, S_MLT0, S_MLT1,
 S_DSUM_ALPHA, S_DSUM_AB, S_DSUM_CD, S_TERM_DATA, S_TERM_PI, S_TERM_CLUSTER, S_TERM_HAPLO) = range(16)

# exit_message strings of run_cavi (use_quality_scores/cavi.py:155-165)
STATUS_MESSAGE = {
    STATUS_RUNNING: "",
    STATUS_CONVERGED: "ELBO converged.",
    STATUS_DECREASING: "Error: ELBO is decreasing.",
    STATUS_NAN: "Error: ELBO is nan.",
    STATUS_MAXITER: "Maximum number of iterations reached.",
}

_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)


class WindowDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("n_reads", C.c_int32), ("length", C.c_int32),
        ("n_clusters", C.c_int32), ("n_starts", C.c_int32), ("max_iter", C.c_int32),
        ("alpha0", C.c_double), ("conv_thres", C.c_double),
        ("codes", _u8p), ("log_hit", _f64p), ("log_off", _f64p), ("weights", _f64p),
        ("ref_codes", _u8p), ("init_cluster", _f64p), ("init_scalars", _f64p),
    ]


class WindowResult(C.Structure):
    _fields_ = [
        ("elbo", _f64p), ("n_iterations", _i32p), ("n_updates", _i32p), ("status", _i32p),
        ("converged", _i32p), ("history_len", _i32p), ("history", _f64p),
        ("history_stride", C.c_int32), ("state_restart", C.c_int32), ("best_restart", C.c_int32),
        ("mean_haplo", _f64p), ("mean_cluster", _f64p), ("alpha", _f64p),
        ("mean_log_pi", _f64p), ("scalars", _f64p),
    ]


EXPORTS = (
    "vl_abi_version", "vl_last_error", "vl_device_count", "vl_batch_workspace_bytes",
    "vl_batch_create", "vl_batch_destroy", "vl_batch_upload", "vl_batch_run", "vl_batch_step",
    "vl_batch_fetch", "vl_batch_set_state", "vl_batch_merge_haplo", "vl_batch_profile_step",
    "vl_batch_last_device_ms", "vl_batch_set_option", "vl_batch_problem_stats",
    "vl_window_plan", "vl_window_plan_capacity", "vl_debug_counters", "vl_debug_exp", "vl_debug_occupancy", "vl_debug_fp64_peak",
    "vl_reads_parse", "vl_reads_fill", "vl_reads_free",
)


class EngineError(RuntimeError):
    """A hard error reported by the C ABI (bad argument or CUDA failure)."""


def library_path() -> str:
    return os.environ.get("VILOCA_B200_LIB", _build.LIB_PATH)


def _load():
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: the CUDA engine is not built (run `python -m viloca_b200.build`); "
            "there is no CPU fallback")
    lib = C.CDLL(path)
    lib.vl_abi_version.restype = C.c_int
    lib.vl_last_error.restype = C.c_char_p
    lib.vl_device_count.restype = C.c_int
    lib.vl_batch_workspace_bytes.argtypes = [C.POINTER(WindowDesc), C.c_int32, C.POINTER(C.c_size_t)]
    lib.vl_batch_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(WindowDesc), C.c_int32,
                                    C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vl_batch_destroy.argtypes = [C.c_void_p]
    lib.vl_batch_destroy.restype = None
    lib.vl_batch_upload.argtypes = [C.c_void_p, C.POINTER(WindowDesc), C.c_int32]
    lib.vl_batch_run.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    lib.vl_batch_step.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
    lib.vl_batch_fetch.argtypes = [C.c_void_p, C.POINTER(WindowResult), C.c_int32]
    lib.vl_batch_set_state.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _f64p, _f64p, _f64p, C.c_int32]
    lib.vl_batch_merge_haplo.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _i32p, C.c_int32, _f64p, _f64p]
    lib.vl_batch_profile_step.argtypes = [C.c_void_p, C.c_int32, _f64p, C.POINTER(C.c_int64)]
    lib.vl_window_plan_capacity.argtypes = [C.c_int32]
    lib.vl_window_plan_capacity.restype = C.c_int32
    lib.vl_window_plan.argtypes = [C.POINTER(WindowDesc), _i32p, C.POINTER(C.c_int64), _i32p]
    lib.vl_reads_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_int32, C.POINTER(C.c_void_p), _i32p, _i32p, _i32p]
    lib.vl_reads_fill.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int32, _u8p, _f64p, _f64p, _i32p,
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.vl_reads_free.argtypes = [C.c_void_p]
    lib.vl_reads_free.restype = None
    lib.vl_debug_occupancy.argtypes = [C.c_int32] * 4 + [_i32p] * 3
    lib.vl_debug_exp.argtypes = [C.c_int32, _f64p, _f64p, C.c_int32]
    if hasattr(lib, "vl_debug_fp64_peak"):      # (absent from older development builds selected with VILOCA_B200_LIB)
        lib.vl_debug_fp64_peak.argtypes = [C.c_int32, _f64p]
    lib.vl_debug_counters.argtypes = [C.POINTER(C.c_uint64), C.c_int32]
    lib.vl_batch_set_option.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.vl_batch_problem_stats.argtypes = [C.c_void_p, _i32p, C.c_int32]
    lib.vl_batch_last_device_ms.argtypes = [C.c_void_p]
    lib.vl_batch_last_device_ms.restype = C.c_double
    if lib.vl_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.vl_abi_version()} != {ABI_VERSION}; rebuild")
    return lib


lib = _load()


def last_error() -> str:
    return lib.vl_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise EngineError(lib.vl_last_error().decode("utf-8", "replace"))


def f64_ptr(a):
    return a.ctypes.data_as(_f64p) if a is not None else None


def u8_ptr(a):
    return a.ctypes.data_as(_u8p) if a is not None else None


def i32_ptr(a):
    return a.ctypes.data_as(_i32p) if a is not None else None
