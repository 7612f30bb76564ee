"""ctypes front end of the CPU oracle (oracle/mpc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It takes plain numpy arrays (a `LinearProblem`-shaped object is read by attribute) so
that it does not depend on the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Problem(C.Structure):
    _fields_ = [("ny", C.c_int), ("nu", C.c_int), ("nd", C.c_int), ("nit", C.c_int),
                ("a", C.c_void_p), ("b0", C.c_void_p), ("b1", C.c_void_p), ("d", C.c_void_p),
                ("umin", C.c_void_p), ("umax", C.c_void_p), ("dumin", C.c_void_p), ("dumax", C.c_void_p),
                ("ymin", C.c_void_p), ("ymax", C.c_void_p), ("ecr_min", C.c_void_p), ("ecr_max", C.c_void_p),
                ("su", C.c_void_p), ("sy", C.c_void_p), ("rho_ecr", C.c_double),
                ("r", C.c_void_p), ("v", C.c_void_p)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "mpc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


_ALT = False     # use_alt_build(True): route eval_batch through liboracle_alt.so (same source, different rounding)
_LIB_ALT = None


def use_alt_build(on: bool):
    global _ALT
    _ALT = bool(on)


def lib():
    global _LIB, _LIB_ALT
    if _ALT:
        if _LIB_ALT is None:
            subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle_alt.so"])
            _LIB_ALT = C.CDLL(os.path.join(_HERE, "liboracle_alt.so"))
            _LIB_ALT.orc_cost_vns.restype = C.c_double
        return _LIB_ALT
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_cost_vns.restype = C.c_double
    return _LIB


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class OracleProblem:
    """Keeps the numpy buffers alive next to the C struct."""

    def __init__(self, prob, r=None, v=None, nit=None):
        self.ny, self.nu, self.nd = int(prob.ny), int(prob.nu), int(prob.nd)
        self.nit = int(prob.nit if nit is None else nit)
        ch = prob.plant
        self._bufs = dict(
            a=_f64(ch.a), b0=_f64(ch.b0), b1=_f64(ch.b1), d=np.ascontiguousarray(ch.d, dtype=np.int32),
            umin=_f64(prob.umin), umax=_f64(prob.umax), dumin=_f64(prob.dumin), dumax=_f64(prob.dumax),
            ymin=_f64(prob.ymin), ymax=_f64(prob.ymax), ecr_min=_f64(prob.ecr_min), ecr_max=_f64(prob.ecr_max),
            su=_f64(prob.su), sy=_f64(prob.sy),
            r=_f64(prob.r if r is None else r).reshape(self.nit, self.ny),
            v=_f64(prob.v if v is None else v).reshape(self.nit, self.nd))
        self.c = _Problem(self.ny, self.nu, self.nd, self.nit)
        for k, arr in self._bufs.items():
            setattr(self.c, k, arr.ctypes.data)
        self.c.rho_ecr = float(prob.rho_ecr)
        self.yref = _f64(prob.yref)
        self.inK = int(prob.inK)


def closedloop(op: OracleProblem, N: int, Nu: int, delta, lam, open_loop: bool = True):
    """closedloop_toolbox.m: returns y,u,ys,uopt as signals x time plus (status, stats[3])."""
    ny, nu, nit = op.ny, op.nu, op.nit
    y = np.zeros((ny, nit)); u = np.zeros((nu, nit)); ys = np.zeros((ny, nit)); uo = np.zeros((nu, nit))
    stats = np.zeros(3, dtype=np.int64)
    dl, lm = _f64(delta), _f64(lam)
    rc = lib().orc_closedloop(C.byref(op.c), int(N), int(Nu), dl.ctypes.data_as(C.c_void_p), lm.ctypes.data_as(C.c_void_p),
                              y.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p),
                              ys.ctypes.data_as(C.c_void_p) if open_loop else None,
                              uo.ctypes.data_as(C.c_void_p) if open_loop else None,
                              stats.ctypes.data_as(C.c_void_p))
    return y, u, ys, uo, rc, stats


def closedloop_est(op: OracleProblem, plant, M, N: int, Nu: int, delta, lam):
    """Mismatch validation run (orc_closedloop_est): `plant` = the real process (Channels, scaled like the model), M = the
    estimator gain in the state order of mpc_oracle.c E1.  Returns y, u (signals x time), status, stats."""
    ny, nu, nit = op.ny, op.nu, op.nit
    y = np.zeros((ny, nit)); u = np.zeros((nu, nit))
    stats = np.zeros(3, dtype=np.int64)
    keep = [_f64(plant.a), _f64(plant.b0), _f64(plant.b1), np.ascontiguousarray(plant.d, dtype=np.int32), _f64(M), _f64(delta), _f64(lam)]
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().orc_closedloop_est(C.byref(op.c), P(keep[0]), P(keep[1]), P(keep[2]), P(keep[3]), P(keep[4]), int(N), int(Nu),
                                  P(keep[5]), P(keep[6]), P(y), P(u), P(stats))
    return y, u, rc, stats


def set_pivot_rule(rule: int):
    """0: most violated constraint first (default); 1: first violated in index order; 2: the previous sample's final active set
    first (resolution probes, oracle/parity.py)."""
    lib().orc_set_pivot_rule(int(rule))


def eval_batch(op: OracleProblem, N, Nu, delta, lam, mode: str = "gam", nthreads: int = 0):
    """mode 'gam' -> cost (n, ny); 'vns' -> cost (n,).  Returns cost, status, stats."""
    N = np.ascontiguousarray(N, dtype=np.int32); Nu = np.ascontiguousarray(Nu, dtype=np.int32)
    n = N.shape[0]
    dl, lm = _f64(delta).reshape(n, op.ny), _f64(lam).reshape(n, op.nu)
    m = {"gam": 0, "vns": 1}[mode]
    cost = np.zeros((n, op.ny) if m == 0 else (n,))
    status = np.zeros(n, dtype=np.int32)
    stats = np.zeros(3, dtype=np.int64)
    lib().orc_eval_batch(C.byref(op.c), n, N.ctypes.data_as(C.c_void_p), Nu.ctypes.data_as(C.c_void_p),
                         dl.ctypes.data_as(C.c_void_p), lm.ctypes.data_as(C.c_void_p), m,
                         op.yref.ctypes.data_as(C.c_void_p), op.inK, cost.ctypes.data_as(C.c_void_p),
                         status.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p), int(nthreads))
    return cost, status, stats


def single_qp(op: OracleProblem, N, Nu, delta, lam, xs, whist, hv, rk, uprev):
    """One mpcmove from an arbitrary state; returns z, H, f, G, yfree, iters (for QP-level checks)."""
    ny, nu, nd = op.ny, op.nu, op.nd
    nw = nu + nd
    whist = _f64(whist).reshape(nw, -1)
    hl = whist.shape[1]
    nzmax = nu * Nu + 1
    z = np.zeros(nzmax); H = np.zeros(nzmax * nzmax); f = np.zeros(nzmax)
    G = np.zeros(ny * N * nu * Nu); yf = np.zeros(ny * N)
    nz = C.c_int(0); it = C.c_int(0)
    P = lambda a: _f64(a).ctypes.data_as(C.c_void_p)
    keep = [_f64(delta), _f64(lam), _f64(xs), whist, _f64(hv), _f64(rk), _f64(uprev)]
    rc = lib().orc_single_qp(C.byref(op.c), int(N), int(Nu), *[k.ctypes.data_as(C.c_void_p) for k in keep[:4]],
                             hl, *[k.ctypes.data_as(C.c_void_p) for k in keep[4:]],
                             z.ctypes.data_as(C.c_void_p), H.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p),
                             G.ctypes.data_as(C.c_void_p), yf.ctypes.data_as(C.c_void_p), C.byref(nz), C.byref(it))
    n = nz.value
    return z[:n], H[:n * n].reshape(n, n), f[:n], G.reshape(ny * N, nu * Nu), yf, it.value, rc


def num_threads() -> int:
    return int(lib().orc_num_threads())
