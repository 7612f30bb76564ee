"""The MEX gateway (mex/mpcgpu_mex.c) compiled against a stub mex.h (tests/mex_stub) and driven through its mexFunction.
CPU: it compiles, parses its arguments and raises MATLAB-style errors (no GPU here: create fails loudly, no CPU fallback).
GPU box: create -> eval -> closedloop (5 outputs, t included) -> nmpc_* -> dtc_* -> destroy, results identical to the
ctypes binding of the same C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mpcgpu
from mpcgpu import shell3x3, synthetic_population

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mex_stub")
DBL, I32, U64 = 6, 12, 15


class Handle:
    def __init__(self, p):
        self.p = p


class Mex:
    def __init__(self):
        subprocess.check_call(["make", "-C", HERE, "-s", "libmexdrv.so"])
        self.l = C.CDLL(os.path.join(HERE, "libmexdrv.so"))
        for f in ("stub_double", "stub_int32", "stub_string", "stub_struct"):
            getattr(self.l, f).restype = C.c_void_p
        self.l.stub_double.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p]
        self.l.stub_int32.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p]
        self.l.stub_string.argtypes = [C.c_char_p]
        self.l.stub_set_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        self.l.stub_call.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        for f in ("stub_m", "stub_n"):
            getattr(self.l, f).restype = C.c_size_t; getattr(self.l, f).argtypes = [C.c_void_p]
        self.l.stub_class.argtypes = [C.c_void_p]
        self.l.stub_data.restype = C.c_void_p; self.l.stub_data.argtypes = [C.c_void_p]
        self.l.stub_errid.restype = C.c_char_p; self.l.stub_errmsg.restype = C.c_char_p

    def mx(self, v):
        """Python value -> mxArray* (column-major, like MATLAB)."""
        if isinstance(v, Handle):                    # an opaque handle coming back from create: the mxArray itself
            return v.p
        if isinstance(v, str):
            return self.l.stub_string(v.encode())
        if isinstance(v, dict):
            s = self.l.stub_struct()
            for k, x in v.items():
                self.l.stub_set_field(s, k.encode(), self.mx(x))
            return s
        a = np.asarray(v)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        m = a.shape[0]; n = int(np.prod(a.shape[1:]))
        if a.dtype == np.int32:
            f = np.asfortranarray(a.reshape(m, n, order="F") if a.ndim > 2 else a.reshape(m, n)); return self.l.stub_int32(m, n, f.ctypes.data_as(C.c_void_p))
        f = np.asfortranarray(a.reshape(m, n, order="F") if a.ndim > 2 else a, dtype=np.float64)
        return self.l.stub_double(m, n, f.ctypes.data_as(C.c_void_p))

    def call(self, nlhs, *args):
        prhs = (C.c_void_p * len(args))(*[self.mx(a) for a in args])
        plhs = (C.c_void_p * max(nlhs, 1))()
        rc = self.l.stub_call(nlhs, plhs, len(args), prhs)
        if rc:
            raise RuntimeError(f"{self.l.stub_errid().decode()}: {self.l.stub_errmsg().decode()}")
        out = []
        for o in range(nlhs):
            p = plhs[o]
            m, n, cls = self.l.stub_m(p), self.l.stub_n(p), self.l.stub_class(p)
            if cls == U64:
                out.append(Handle(p))
            else:
                dt = np.float64 if cls == DBL else np.int32
                buf = (C.c_char * (m * n * np.dtype(dt).itemsize)).from_address(self.l.stub_data(p))
                out.append(np.frombuffer(buf, dtype=dt).reshape(m, n, order="F").copy())
        return out


def problem_struct(p):
    """The struct a MATLAB caller fills from Par / the scaled mpc object (fields of mpcgpu_problem, row-major matrices)."""
    ch = p.plant
    return dict(ny=float(p.ny), nu=float(p.nu), nd=float(p.nd), nit=float(p.nit), pmax=float(2 ** p.nbp - 1), mmax=float(2 ** p.nbc - 1),
                inK=float(p.inK), Ts=float(p.Ts), a=ch.a.ravel(), b0=ch.b0.ravel(), b1=ch.b1.ravel(), d=ch.d.ravel().astype(np.int32),
                umin=p.umin, umax=p.umax, dumin=p.dumin, dumax=p.dumax, ymin=p.ymin, ymax=p.ymax, ecr_min=p.ecr_min, ecr_max=p.ecr_max,
                su=p.su, sy=p.sy, rho_ecr=float(p.rho_ecr), r=np.ascontiguousarray(p.r).ravel(), v=np.ascontiguousarray(p.v).ravel(),
                yref=np.ascontiguousarray(p.yref).ravel(), dmin=np.asarray(p.dmin, dtype=np.int32))


@pytest.fixture(scope="module")
def mex():
    return Mex()


def test_gateway_compiles_and_reports_errors_like_matlab(mex):
    with pytest.raises(RuntimeError, match="mpcgpu:arg: unknown command"):
        mex.call(0, "frobnicate")
    with pytest.raises(RuntimeError, match="problem struct expected"):
        mex.call(1, "create", 3.0)
    with pytest.raises(RuntimeError, match="problem field a missing"):
        mex.call(1, "create", dict(ny=3.0, nu=3.0, nd=0.0, nit=10.0, pmax=127.0, mmax=15.0, inK=10.0))
    import torch
    if not torch.cuda.is_available():     # no GPU: the library refuses (no CPU fallback) and the gateway raises mpcgpu:create
        with pytest.raises(RuntimeError, match="mpcgpu:create: no CUDA device"):
            mex.call(1, "create", problem_struct(shell3x3(2)))


@pytest.mark.gpu
def test_linear_commands_match_the_ctypes_binding(mex):
    p = shell3x3(2)
    (h,) = mex.call(1, "create", problem_struct(p))
    ev = mpcgpu.Evaluator(p, device=0)
    N, Nu, dl, lm = synthetic_population(p, 200, seed=21)
    for mode in ("gam", "vns"):
        cost, st = mex.call(2, "eval", h, N.astype(np.float64), Nu, dl, lm, mode)     # N as double, Nu as int32: both accepted
        ref = ev.eval_batch(N, Nu, dl, lm, mode=mode)
        assert np.array_equal(st[:, 0], ref["status"])
        assert np.array_equal(cost if mode == "gam" else cost[:, 0], ref["cost"])
    # closedloop_toolbox.m:1 -- five outputs, signals x time, t = k*Ts; r given signals x time, N / Nu as vectors
    nit = 54
    r_ma = (np.ones((3, nit)) * p.L[:, None]) * np.array([[1.0], [0.0], [0.0]])
    y, u, t, ys, uopt = mex.call(5, "closedloop", h, r_ma, np.zeros((nit, 0)), np.array([24.0, 20.0, 7.0]), np.array([6.0, 2.0, 2.0]), dl[0], lm[0], float(nit))
    y0, u0, t0, ys0, uo0 = mpcgpu.closedloop_toolbox(ev, r_ma, np.zeros((nit, 0)), [24, 20, 7], [6, 2, 2], dl[0], lm[0], nit)
    assert y.shape == (3, nit) and u.shape == (3, nit) and t.shape == (1, nit) and ys.shape == (3, nit) and uopt.shape == (3, nit)
    for a, b in ((y, y0), (u, u0), (t, t0), (ys, ys0), (uopt, uo0)):
        assert np.array_equal(a, b)
    # ... and the handle's own signals are untouched: same GAM cost afterwards
    cost2, _ = mex.call(2, "eval", h, N, Nu, dl, lm, "gam")
    assert np.array_equal(cost2, ev.eval_batch(N, Nu, dl, lm, mode="gam")["cost"])
    # one output: a failed candidate raises, which is what the reference's try/catch expects
    with pytest.raises(RuntimeError, match="mpcgpu:candidate"):
        mex.call(1, "eval", h, np.array([3.0]), np.array([9.0]), dl[:1], lm[:1], "gam")
    mex.call(0, "option", h, "vns_legality", 1.0)
    _, st = mex.call(2, "eval", h, np.array([30.0, 30.0]), np.array([1.0, 2.0]), dl[:2], lm[:2], "vns")
    assert list(st[:, 0]) == [4, 0]
    # all GPUs of the box through one handle
    (hm,) = mex.call(1, "create_multi", problem_struct(p), np.zeros((0, 1), dtype=np.int32))
    costm, stm = mex.call(2, "eval_multi", hm, N, Nu, dl, lm, "gam")
    assert np.array_equal(costm, ev.eval_batch(N, Nu, dl, lm, mode="gam")["cost"])
    mex.call(0, "destroy_multi", hm)
    mex.call(0, "destroy", h)
    ev.close()


@pytest.mark.gpu
def test_nmpc_and_dtc_commands(mex):
    pn = mpcgpu.vandevusse()
    Pn = dict(nit=float(pn.nit), pmax=float(2 ** pn.nbp - 1), mmax=float(2 ** pn.nbc - 1), inK=float(pn.inK), nsub=float(pn.nsub),
              max_sqp=float(pn.max_sqp), Ts=float(pn.Ts), x0=pn.x0, u0=pn.u0, umin=pn.umin, umax=pn.umax, xmin=pn.xmin, xmax=pn.xmax,
              su=pn.su, sy=pn.sy, r=np.ascontiguousarray(pn.r).ravel(), yref=np.ascontiguousarray(pn.yref).ravel())
    (hn,) = mex.call(1, "nmpc_create", Pn)
    en = mpcgpu.NmpcEvaluator(pn, device=0)
    N, Nu, dl, lm = mpcgpu.synthetic_nmpc_population(pn, 24, seed=3)
    cost, st = mex.call(2, "nmpc_eval", hn, N, Nu, dl, lm, "gam")
    ref = en.eval_batch(N, Nu, dl, lm, mode="gam")
    assert np.array_equal(cost, ref["cost"]) and np.array_equal(st[:, 0], ref["status"])
    y, u, yopt, uopt = mex.call(4, "nmpc_closedloop", hn, pn.r, 10.0, np.array([2.0, 2.0]), np.array([1.0, 1.0]), np.array([0.1, 0.1]), float(pn.nit))
    y0, u0, yo0, uo0 = mpcgpu.closedloop_toolbox_nmpc(en, None, None, pn.r, 10, [2, 2], [1, 1], [0.1, 0.1], pn.nit)
    for a, b in ((y, y0), (u, u0), (yopt, yo0), (uopt, uo0)):
        assert a.shape == (2, pn.nit) and np.array_equal(a, b)
    mex.call(0, "nmpc_destroy", hn)
    en.close()
    # DTC-GPC sweep
    pd = mpcgpu.woodberry_dtc()
    ed = mpcgpu.DtcEvaluator(pd, device=0)
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64)).ravel()
    Pd = dict(ny=2.0, nu=2.0, nq=float(pd.pq.a.shape[1]), nit=float(pd.nit), pmax=float(pd.pmax), mmax=float(pd.mmax), k_start=4.0,
              ma=f64(pd.pnz.a), mb0=f64(pd.pnz.b0), mb1=f64(pd.pnz.b1), md=pd.pnz.d.ravel().astype(np.int32),
              pa=f64(pd.preal.a), pb0=f64(pd.preal.b0), pb1=f64(pd.preal.b1), pd=pd.preal.d.ravel().astype(np.int32),
              qa=f64(pd.pq.a), qb0=f64(pd.pq.b0), qb1=f64(pd.pq.b1), qd=pd.pq.d.ravel().astype(np.int32),
              L=f64(pd.L), R=f64(pd.R), r=f64(pd.r), q=f64(pd.q))
    (hd,) = mex.call(1, "dtc_create", Pd)
    p_, m_, dl, lm, alfa, raio = mpcgpu.synthetic_dtc_population(pd, 32, seed=5)
    filters = [mpcgpu.mimo_filter(pd.pnz, float(a), float(r_)) for a, r_ in zip(alfa, raio)]
    num, den, ln = ed.pack_filters(filters)
    ise, st = mex.call(2, "dtc_eval", hd, p_, m_, dl, lm, num, den, ln)
    ref = ed.eval_batch(p_, m_, dl, lm, filters=filters)
    assert np.array_equal(st[:, 0], ref["status"]) and np.array_equal(ise, ref["ise"])
    mex.call(0, "dtc_destroy", hd)
    ed.close()
