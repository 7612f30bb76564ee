"""ctypes front end of oracle/nmpc_port.cpp: the CPU baseline ("port") of the nonlinear path.  BENCH / TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "-s", "libnmpcport.so"])
        _LIB = C.CDLL(os.path.join(_HERE, "libnmpcport.so"))
    return _LIB


def eval_batch(prob, N, Nu, delta, lam, mode="gam", nthreads=0):
    """prob: an NmpcProblem-shaped object (attributes only).  Returns (cost, status)."""
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    N = np.ascontiguousarray(N, dtype=np.int32); Nu = np.ascontiguousarray(Nu, dtype=np.int32)
    n = len(N)
    m = {"gam": 1, "vns": 2}[mode]
    bufs = [f64(getattr(prob, k)) for k in ("x0", "u0", "umin", "umax", "xmin", "xmax", "su", "sy", "r", "yref")]
    dl = f64(delta).reshape(n, 2); lm = f64(lam).reshape(n, 2)
    cost = np.empty((n, 2)) if m == 1 else np.empty(n)
    status = np.zeros(n, dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().nmpc_port_eval_batch(C.c_int(int(prob.nit)), C.c_int(2 ** prob.nbp - 1), C.c_int(2 ** prob.nbc - 1), C.c_int(int(prob.inK)),
                               C.c_int(int(prob.nsub)), C.c_int(int(prob.max_sqp)), C.c_double(float(prob.Ts)),
                               *[P(b) for b in bufs], C.c_int(n), P(N), P(Nu), P(dl), P(lm), C.c_int(m), P(cost), P(status),
                               C.c_int(int(nthreads)))
    return cost, status
