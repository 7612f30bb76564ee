"""TEST-ONLY loader of tests/host_emulation/libmpcemu.so (lane-serialised host build of the kernel source)."""
import ctypes as C
import os
import subprocess

import numpy as np

from mpcgpu._capi import make_problem_struct, ProblemStruct

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emulation")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s", "libmpcemu.so"])
        _lib = C.CDLL(os.path.join(_HERE, "libmpcemu.so"))
    return _lib


def eval_batch(prob, N, Nu, delta, lam, mode="gam", traj=False, r=None, v=None, yref=None, nit=None):
    ps, keep = make_problem_struct(prob, r=r, v=v, yref=yref, nit=nit)
    nit = ps.nit
    N = np.ascontiguousarray(N, dtype=np.int32); Nu = np.ascontiguousarray(Nu, dtype=np.int32)
    n = len(N)
    dl = np.ascontiguousarray(delta, dtype=np.float64).reshape(n, prob.ny)
    lm = np.ascontiguousarray(lam, dtype=np.float64).reshape(n, prob.nu)
    m = {"raw": 0, "gam": 1, "vns": 2}[mode]
    cost = np.zeros((n, prob.ny) if m == 1 else (n,))
    status = np.zeros(n, dtype=np.int32)
    counters = np.zeros(2, dtype=np.uint64)
    tr = [np.zeros((n, prob.ny, nit)), np.zeros((n, prob.nu, nit)), np.zeros((n, prob.ny, nit)), np.zeros((n, prob.nu, nit))] if traj else [None] * 4
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    err = C.create_string_buffer(256)
    rc = lib().emu_eval_batch(C.byref(ps), n, P(N), P(Nu), P(dl), P(lm), m, P(cost), *[P(t) for t in tr], P(status),
                              P(counters), err, 256)
    if rc:
        raise RuntimeError(err.value.decode())
    return cost, status, counters, tr


def eval_est(prob, plant, gain, hl, N, Nu, delta, lam):
    """Host emulation of the validation run against a mismatched plant (k_soft<NU,16,true>), one candidate: y, u, cost, status."""
    ps, keep = make_problem_struct(prob)
    a, b0, b1 = (np.ascontiguousarray(x, dtype=np.float64) for x in (plant.a, plant.b0, plant.b1))
    d = np.ascontiguousarray(plant.d, dtype=np.int32)
    g = np.ascontiguousarray(gain, dtype=np.float64)
    dl = np.ascontiguousarray(delta, dtype=np.float64); lm = np.ascontiguousarray(lam, dtype=np.float64)
    y = np.zeros((prob.ny, ps.nit)); u = np.zeros((prob.nu, ps.nit)); cost = np.zeros(prob.ny)
    P = lambda x: x.ctypes.data_as(C.c_void_p)
    err = C.create_string_buffer(256)
    rc = lib().emu_eval_est(C.byref(ps), P(a), P(b0), P(b1), P(d), P(g), int(hl), int(N), int(Nu), P(dl), P(lm), P(cost), P(y), P(u), err, 256)
    if rc < 0:
        raise RuntimeError(err.value.decode())
    return y, u, cost, rc


def nmpc_eval(prob, N, Nu, delta, lam, mode="gam", traj=False):
    """Host emulation of the NMPC group kernel (nmg_run<32>), one candidate: cost, status, (y, u), counters."""
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    bufs = [f64(getattr(prob, k)) for k in ("x0", "u0", "umin", "umax", "xmin", "xmax", "su", "sy", "r", "yref")]
    dl = f64(delta); lm = f64(lam)
    m = {"raw": 0, "gam": 1, "vns": 2}[mode]
    cost = np.zeros(2 if m == 1 else 1)
    y = np.zeros((2, prob.nit)) if traj else None
    u = np.zeros((2, prob.nit)) if traj else None
    cnt = np.zeros(2, dtype=np.uint64)
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    st = lib().emu_nmpc_eval(C.c_int(int(prob.nit)), C.c_int(2 ** prob.nbp - 1), C.c_int(2 ** prob.nbc - 1), C.c_int(int(prob.inK)),
                             C.c_int(int(prob.nsub)), C.c_int(int(prob.max_sqp)), C.c_double(float(prob.Ts)), *[P(b) for b in bufs],
                             C.c_int(int(N)), C.c_int(int(Nu)), P(dl), P(lm), C.c_int(m), P(cost), P(y), P(u), P(cnt))
    return cost, st, (y, u), cnt


def nmpc_eval_group(prob, N, Nu, delta, lam, G=16):
    """Host emulation of the NMPC group kernel with G lanes per run: a population of candidates, 32 / G runs per warp (the
    groups of a warp diverge where their horizons and iteration counts differ).  GAM costs and status."""
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    bufs = [f64(getattr(prob, k)) for k in ("x0", "u0", "umin", "umax", "xmin", "xmax", "su", "sy", "r", "yref")]
    N = np.ascontiguousarray(N, dtype=np.int32); Nu = np.ascontiguousarray(Nu, dtype=np.int32); n = len(N)
    dl = f64(delta).reshape(n, 2); lm = f64(lam).reshape(n, 2)
    cost = np.zeros((n, 2)); status = np.zeros(n, dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().emu_nmpc_eval_group(C.c_int(int(G)), C.c_int(int(prob.nit)), C.c_int(2 ** prob.nbp - 1), C.c_int(2 ** prob.nbc - 1), C.c_int(int(prob.inK)),
                              C.c_int(int(prob.nsub)), C.c_int(int(prob.max_sqp)), C.c_double(float(prob.Ts)), *[P(b) for b in bufs],
                              C.c_int(n), P(N), P(Nu), P(dl), P(lm), P(cost), P(status))
    return cost, status


def dtc_eval(prob, p, m, delta, lam, alfa, raio, traj=True):
    """Host emulation of the DTC-GPC kernels (k_dtc_filter + k_dtc): ise (n x ny), status, y, u."""
    from mpcgpu.dtcgpc import DtcProblemStruct
    ny, nu = prob.pnz.a.shape
    nq = prob.pq.a.shape[1]
    nit = int(prob.nit)
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32))
    keep = dict(ma=f64(prob.pnz.a), mb0=f64(prob.pnz.b0), mb1=f64(prob.pnz.b1), md=i32(prob.pnz.d),
                pa=f64(prob.preal.a), pb0=f64(prob.preal.b0), pb1=f64(prob.preal.b1), pd=i32(prob.preal.d),
                qa=f64(prob.pq.a), qb0=f64(prob.pq.b0), qb1=f64(prob.pq.b1), qd=i32(prob.pq.d),
                L=f64(prob.L), R=f64(prob.R), r=f64(prob.r).reshape(ny, nit), q=f64(prob.q).reshape(nq, nit))
    ps = DtcProblemStruct(ny, nu, nq, nit, int(prob.pmax), int(prob.mmax), int(prob.k_start), 0)
    for k, arr in keep.items():
        setattr(ps, k, arr.ctypes.data if arr.size else None)
    p = i32(np.atleast_2d(p)); n = p.shape[0]
    m = i32(np.atleast_2d(m)); dl = f64(delta).reshape(n, ny); lm = f64(lam).reshape(n, nu)
    al = f64(np.broadcast_to(alfa, (n,))); ra = f64(np.broadcast_to(raio, (n,)))
    ise = np.zeros((n, ny)); status = np.full(n, -1, dtype=np.int32)
    y = np.zeros((n, ny, nit)) if traj else None
    u = np.zeros((n, nu, nit)) if traj else None
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    err = C.create_string_buffer(256)
    rc = lib().emu_dtc_eval(C.byref(ps), n, P(p), P(m), P(dl), P(lm), P(al), P(ra), P(ise), P(y), P(u), P(status), err, 256)
    if rc:
        raise RuntimeError(err.value.decode())
    return ise, status, y, u


def build_matrices(prob, p, m, P, delta, lam, threads=0):
    """The builder's output (M: nst x R as [col][row], W: 2R x R) for one candidate: threads=0 the serial host build of
    mpc_core.cuh, threads=128 the same source executed by that many host threads with real block barriers."""
    ps, keep = make_problem_struct(prob)
    R = prob.nu * P
    dl = np.ascontiguousarray(delta, dtype=np.float64); lm = np.ascontiguousarray(lam, dtype=np.float64)
    Mg = np.zeros(256 * R); Wg = np.zeros(2 * R * R)
    err = C.create_string_buffer(256)
    Pp = lambda a: a.ctypes.data_as(C.c_void_p)
    if threads:
        rc = lib().emu_build_threads(C.byref(ps), int(p), int(m), int(P), Pp(dl), Pp(lm), Pp(Mg), Pp(Wg), int(threads), err, 256)
    else:
        rc = lib().emu_build_serial(C.byref(ps), int(p), int(m), int(P), Pp(dl), Pp(lm), Pp(Mg), Pp(Wg), err, 256)
    if rc < 0:
        raise RuntimeError(err.value.decode())
    return Mg, Wg, rc
