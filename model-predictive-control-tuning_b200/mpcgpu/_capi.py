"""ctypes binding of include/mpcgpu.h (the same binding a MEX gateway would make)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPCGPU_LIB") or os.path.join(os.path.dirname(_PKG), "csrc", "libmpcgpu.so")   # MPCGPU_LIB: A/B builds

COST_RAW, COST_GAM, COST_VNS = 0, 1, 2
OPT_VNS_LEGALITY = 1
_MODES = {"raw": COST_RAW, "gam": COST_GAM, "vns": COST_VNS}


class ProblemStruct(C.Structure):
    """mpcgpu_problem (include/mpcgpu.h)."""
    _fields_ = [("ny", C.c_int32), ("nu", C.c_int32), ("nd", C.c_int32), ("nit", C.c_int32),
                ("pmax", C.c_int32), ("mmax", C.c_int32), ("inK", C.c_int32), ("reserved", C.c_int32),
                ("a", C.c_void_p), ("b0", C.c_void_p), ("b1", C.c_void_p), ("d", C.c_void_p),
                ("umin", C.c_void_p), ("umax", C.c_void_p), ("dumin", C.c_void_p), ("dumax", C.c_void_p),
                ("ymin", C.c_void_p), ("ymax", C.c_void_p), ("ecr_min", C.c_void_p), ("ecr_max", C.c_void_p),
                ("su", C.c_void_p), ("sy", C.c_void_p), ("rho_ecr", C.c_double),
                ("r", C.c_void_p), ("v", C.c_void_p), ("yref", C.c_void_p), ("dmin", C.c_void_p)]


class Counters(C.Structure):
    """mpcgpu_counters (include/mpcgpu.h)."""
    _fields_ = [("candidates", C.c_uint64), ("closed_loops", C.c_uint64), ("qp_solves", C.c_uint64),
                ("qp_constrained", C.c_uint64), ("as_iterations", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("last_build_ms", C.c_double), ("last_sim_ms", C.c_double), ("last_total_ms", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def make_problem_struct(prob, r=None, v=None, yref=None, nit=None):
    """LinearProblem -> (mpcgpu_problem, keepalive list of numpy buffers)."""
    nit = int(prob.nit if nit is None else nit)
    ch = prob.plant
    bufs = dict(
        a=_f64(ch.a), b0=_f64(ch.b0), b1=_f64(ch.b1), d=np.ascontiguousarray(ch.d, dtype=np.int32),
        umin=_f64(prob.umin), umax=_f64(prob.umax), dumin=_f64(prob.dumin), dumax=_f64(prob.dumax),
        ymin=_f64(prob.ymin), ymax=_f64(prob.ymax), ecr_min=_f64(prob.ecr_min), ecr_max=_f64(prob.ecr_max),
        su=_f64(prob.su), sy=_f64(prob.sy),
        r=_f64(prob.r if r is None else r).reshape(nit, prob.ny),
        v=_f64(prob.v if v is None else v).reshape(nit, prob.nd),
        yref=_f64(prob.yref if yref is None else yref).reshape(prob.ny, nit),
        dmin=np.ascontiguousarray(prob.dmin, dtype=np.int32))
    ps = ProblemStruct(int(prob.ny), int(prob.nu), int(prob.nd), nit, 2 ** int(prob.nbp) - 1, 2 ** int(prob.nbc) - 1,
                       int(prob.inK), 0)
    for k, arr in bufs.items():
        setattr(ps, k, arr.ctypes.data if arr.size else None)
    ps.rho_ecr = float(prob.rho_ecr)
    return ps, bufs


_lib = None


def load_library():
    """Load libmpcgpu.so.  There is no CPU fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  mpcgpu has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.mpcgpu_last_error.restype = C.c_char_p
    lib.mpcgpu_last_error.argtypes = [C.c_void_p]
    lib.mpcgpu_create.argtypes = [C.POINTER(ProblemStruct), C.c_int, C.POINTER(C.c_void_p)]
    lib.mpcgpu_destroy.argtypes = [C.c_void_p]
    lib.mpcgpu_destroy.restype = None
    lib.mpcgpu_set_signals.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mpcgpu_eval_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 6
    lib.mpcgpu_upload.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    lib.mpcgpu_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.mpcgpu_download.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
    lib.mpcgpu_cost_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    lib.mpcgpu_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    lib.mpcgpu_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.mpcgpu_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.mpcgpu_set_mismatch.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int]
    lib.mpcgpu_closedloop.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 7
    lib.mpcgpu_create_multi.argtypes = [C.POINTER(ProblemStruct), C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.mpcgpu_destroy_multi.argtypes = [C.c_void_p]
    lib.mpcgpu_destroy_multi.restype = None
    lib.mpcgpu_multi_device_count.argtypes = [C.c_void_p]
    lib.mpcgpu_multi_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.mpcgpu_multi_set_signals.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mpcgpu_multi_eval_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 2
    lib.mpcgpu_multi_get_counters.argtypes = [C.c_void_p, C.c_int, C.POINTER(Counters)]
    lib.mpcgpu_multi_last_error.restype = C.c_char_p
    lib.mpcgpu_multi_last_error.argtypes = [C.c_void_p]
    lib.mpcgpu_work_estimate.argtypes = [C.POINTER(ProblemStruct), C.c_int] + [C.c_void_p] * 5
    lib.mpcgpu_dtc_last_error.restype = C.c_char_p
    lib.mpcgpu_dtc_last_error.argtypes = [C.c_void_p]
    lib.mpcgpu_dtc_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.mpcgpu_dtc_destroy.argtypes = [C.c_void_p]
    lib.mpcgpu_dtc_destroy.restype = None
    lib.mpcgpu_dtc_eval_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11
    lib.mpcgpu_dtc_eval_batch_design.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 10
    lib.mpcgpu_dtc_host_tables.argtypes = [C.c_void_p] * 5
    lib.mpcgpu_dtc_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    lib.mpcgpu_nmpc_last_error.restype = C.c_char_p
    lib.mpcgpu_nmpc_last_error.argtypes = [C.c_void_p]
    lib.mpcgpu_nmpc_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.mpcgpu_nmpc_destroy.argtypes = [C.c_void_p]
    lib.mpcgpu_nmpc_destroy.restype = None
    lib.mpcgpu_nmpc_eval_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 7
    lib.mpcgpu_nmpc_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    lib.mpcgpu_ssnmpc_last_error.restype = C.c_char_p
    lib.mpcgpu_ssnmpc_last_error.argtypes = [C.c_void_p]
    lib.mpcgpu_ssnmpc_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.mpcgpu_ssnmpc_destroy.argtypes = [C.c_void_p]
    lib.mpcgpu_ssnmpc_destroy.restype = None
    lib.mpcgpu_ssnmpc_eval_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 10
    lib.mpcgpu_ssnmpc_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "mpcgpu_create", "mpcgpu_destroy", "mpcgpu_set_signals", "mpcgpu_eval_batch", "mpcgpu_upload", "mpcgpu_run",
    "mpcgpu_download", "mpcgpu_cost_device_ptr", "mpcgpu_get_counters", "mpcgpu_last_error", "mpcgpu_device_count",
    "mpcgpu_measure_fp64_peak", "mpcgpu_set_option", "mpcgpu_closedloop", "mpcgpu_set_mismatch",
    "mpcgpu_create_multi", "mpcgpu_destroy_multi", "mpcgpu_multi_device_count", "mpcgpu_multi_set_option",
    "mpcgpu_multi_set_signals", "mpcgpu_multi_eval_batch", "mpcgpu_multi_get_counters", "mpcgpu_multi_last_error",
    "mpcgpu_work_estimate",
    "mpcgpu_dtc_create", "mpcgpu_dtc_destroy", "mpcgpu_dtc_eval_batch", "mpcgpu_dtc_eval_batch_design", "mpcgpu_dtc_last_error",
    "mpcgpu_dtc_host_tables", "mpcgpu_dtc_get_counters",
    "mpcgpu_nmpc_create", "mpcgpu_nmpc_destroy", "mpcgpu_nmpc_eval_batch", "mpcgpu_nmpc_get_counters", "mpcgpu_nmpc_last_error",
    "mpcgpu_ssnmpc_create", "mpcgpu_ssnmpc_destroy", "mpcgpu_ssnmpc_eval_batch", "mpcgpu_ssnmpc_get_counters", "mpcgpu_ssnmpc_last_error",
]
