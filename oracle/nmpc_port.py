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


def ssnmpc_eval_batch(prob, N, Nu, Q, W, noise=None, traj=False, nthreads=0, counters=None):
    """Host build of csrc/mpc_ssnmpc_core.h (the single-shooting formulation).  prob: an SsnmpcProblem-shaped object.
    Returns (cost n x 2, status[, y, u])."""
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    N = np.ascontiguousarray(np.atleast_1d(N), dtype=np.int32); n = len(N)
    Nu = np.ascontiguousarray(Nu, dtype=np.int32).reshape(n, 2)
    Qa = f64(Q).reshape(n, 2); Wa = f64(W).reshape(n, 2)
    x0, u0, lb, ub, r = f64(prob.x0), f64(prob.u0), f64(prob.lb), f64(prob.ub), f64(prob.r)
    xc = np.ascontiguousarray(prob.x_control, dtype=np.int32)
    nz = None if noise is None else f64(noise)
    nit = int(prob.nit)
    cost = np.empty((n, 2)); status = np.zeros(n, dtype=np.int32)
    y = np.empty((n, 2, nit)) if traj else None
    u = np.empty((n, 2, nit)) if traj else None
    cnt = np.zeros(2, dtype=np.uint64)
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    lib().ssnmpc_port_eval_batch(C.c_int(nit), C.c_int(int(prob.pmax)), C.c_int(int(prob.inK)), C.c_int(int(prob.nsub)),
                                 C.c_int(int(prob.max_sqp)), C.c_double(float(prob.Ts)), P(x0), P(u0), P(lb), P(ub), P(xc), P(r), P(nz),
                                 C.c_int(n), P(N), P(Nu), P(Qa), P(Wa), P(cost), P(y), P(u), P(status), C.c_int(int(nthreads)), P(cnt))
    if counters is not None:
        counters.update(controller_calls=int(cnt[0]), gn_iterations=int(cnt[1]))
    return (cost, status, y, u) if traj else (cost, status)


def ssnmpc_controller(prob, x, uprev, r, N, Nu, Q, W):
    """One NMPC_Controller call of the host build: returns (X, n_sqp, rc)."""
    f64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    Nu = np.ascontiguousarray(Nu, dtype=np.int32); xc = np.ascontiguousarray(prob.x_control, dtype=np.int32)
    X = np.zeros(int(Nu.sum())); ns = C.c_int(0)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    bufs = [f64(prob.lb), f64(prob.ub), xc, f64(x), f64(uprev), f64(r)]
    Qa, Wa = f64(Q), f64(W)
    rc = lib().ssnmpc_port_controller(C.c_int(int(prob.nsub)), C.c_int(int(prob.max_sqp)), C.c_double(float(prob.Ts)),
                                      *[P(b) for b in bufs], C.c_int(int(N)), P(Nu), P(Qa), P(Wa), P(X), C.byref(ns))
    return X, ns.value, rc
