"""Host-side mirror of the reference's evaluator interface on top of libmpcgpu.so.

Reference interface -> here (same names, argument meaning, orientation and error behaviour):
  closedloop_toolbox(mpc_toolbox,r,v,N,Nu,delta,lambda,nit)   MPC_Tuning/closedloop_toolbox.m:1
  GAM_fun(X,Par)                                              MPC_Tuning/GAM_fun.m:1
  VNS trial cost (inline block)                               MPC_Tuning/VNS2.m:147-195
plus the batched forms the reference cannot express (`Evaluator.eval_batch`), which is what the GPU is for.
There is no CPU fallback: everything goes through the C ABI and fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .problems import LinearProblem


class MpcGpuError(RuntimeError):
    pass


def row2col(v):
    """MPC_Tuning/row2col.m:3-8: transpose when there are fewer rows than columns."""
    v = np.atleast_2d(np.asarray(v, dtype=np.float64))
    return v.T if v.shape[0] < v.shape[1] else v


def col2row(v):
    """MPC_Tuning/col2row.m:3-8."""
    v = np.atleast_2d(np.asarray(v, dtype=np.float64))
    return v.T if v.shape[0] > v.shape[1] else v


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Evaluator:
    """One `mpcgpu_handle`: a problem (plant, limits, signals) resident on one B200."""

    def __init__(self, prob: LinearProblem, device: int = -1, r=None, v=None, yref=None, nit=None):
        self.lib = _capi.load_library()
        self.prob = prob
        ps, self._keep = _capi.make_problem_struct(prob, r=r, v=v, yref=yref, nit=nit)
        self.ny, self.nu, self.nd, self.nit = prob.ny, prob.nu, prob.nd, ps.nit
        h = C.c_void_p()
        rc = self.lib.mpcgpu_create(C.byref(ps), int(device), C.byref(h))
        if rc != 0:
            raise MpcGpuError(f"mpcgpu_create failed ({rc}): {self.lib.mpcgpu_last_error(None).decode()}")
        self.h = h
        self.n = 0
        # current horizons of the tuner state `Par` (MPCTuning.m:283-289: N = 2^nbp-1, Nu = 2)
        self.N = 2 ** int(prob.nbp) - 1
        self.Nu = 2

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.mpcgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise MpcGpuError(f"{what} failed ({rc}): {self.lib.mpcgpu_last_error(self.h).decode()}")

    # -- signals -------------------------------------------------------------------------------
    def set_signals(self, r, v=None, yref=None, nit=None):
        """r: nit x ny (either orientation, row2col.m), v: nit x nd, yref: ny x nit."""
        r = np.ascontiguousarray(row2col(r))
        nit = r.shape[0] if nit is None else int(nit)
        r = np.ascontiguousarray(r[:nit])
        if self.nd:
            v = np.ascontiguousarray(row2col(v)[:nit])
        else:
            v = None
        yr = None if yref is None else np.ascontiguousarray(col2row(yref)[:, :nit])
        self._check(self.lib.mpcgpu_set_signals(self.h, nit, _ptr(r), _ptr(v), _ptr(yr)), "mpcgpu_set_signals")
        self.nit = nit

    # -- population ----------------------------------------------------------------------------
    def _pop(self, N, Nu, delta, lam):
        N = np.ascontiguousarray(np.atleast_1d(N), dtype=np.int32)
        Nu = np.ascontiguousarray(np.atleast_1d(Nu), dtype=np.int32)
        n = N.shape[0]
        delta = np.ascontiguousarray(delta, dtype=np.float64).reshape(n, self.ny)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(n, self.nu)
        return n, N, Nu, delta, lam

    def eval_batch(self, N, Nu, delta, lam, mode="gam", traj=False):
        """Host arrays in, host arrays out (the reference-facing call).
        Returns dict(cost, status[, y, u, ys, uopt])."""
        n, N, Nu, delta, lam = self._pop(N, Nu, delta, lam)
        m = _capi._MODES[mode]
        cost = np.empty((n, self.ny)) if m == _capi.COST_GAM else (np.empty(n) if m == _capi.COST_VNS else None)
        status = np.zeros(n, dtype=np.int32)
        tr = [None] * 4
        if traj or m == _capi.COST_RAW:
            tr = [np.empty((n, self.ny, self.nit)), np.empty((n, self.nu, self.nit)),
                  np.empty((n, self.ny, self.nit)), np.empty((n, self.nu, self.nit))]
        rc = self.lib.mpcgpu_eval_batch(self.h, n, _ptr(N), _ptr(Nu), _ptr(delta), _ptr(lam), m, _ptr(cost),
                                        _ptr(tr[0]), _ptr(tr[1]), _ptr(tr[2]), _ptr(tr[3]), _ptr(status))
        self._check(rc, "mpcgpu_eval_batch")
        self.n = n
        out = {"cost": cost, "status": status}
        if tr[0] is not None:
            out.update(y=tr[0], u=tr[1], ys=tr[2], uopt=tr[3])
        return out

    def upload(self, N, Nu, delta, lam):
        n, N, Nu, delta, lam = self._pop(N, Nu, delta, lam)
        self._check(self.lib.mpcgpu_upload(self.h, n, _ptr(N), _ptr(Nu), _ptr(delta), _ptr(lam)), "mpcgpu_upload")
        self.n = n

    def run(self, mode="gam", traj=False, stream=None):
        """Kernels only, asynchronous on `stream` (a cudaStream_t as int, e.g. torch's current stream)."""
        self._check(self.lib.mpcgpu_run(self.h, _capi._MODES[mode], int(bool(traj)), C.c_void_p(stream) if stream else None),
                    "mpcgpu_run")

    def download(self, mode="gam", traj=False):
        m = _capi._MODES[mode]
        n = self.n
        cost = np.empty((n, self.ny)) if m == _capi.COST_GAM else (np.empty(n) if m == _capi.COST_VNS else None)
        status = np.zeros(n, dtype=np.int32)
        tr = [None] * 4
        if traj:
            tr = [np.empty((n, self.ny, self.nit)), np.empty((n, self.nu, self.nit)),
                  np.empty((n, self.ny, self.nit)), np.empty((n, self.nu, self.nit))]
        self._check(self.lib.mpcgpu_download(self.h, m, _ptr(cost), _ptr(tr[0]), _ptr(tr[1]), _ptr(tr[2]), _ptr(tr[3]),
                                             _ptr(status)), "mpcgpu_download")
        out = {"cost": cost, "status": status}
        if traj:
            out.update(y=tr[0], u=tr[1], ys=tr[2], uopt=tr[3])
        return out

    def cost_device_ptr(self, mode="gam"):
        p = C.c_void_p()
        cnt = C.c_int()
        self._check(self.lib.mpcgpu_cost_device_ptr(self.h, _capi._MODES[mode], C.byref(p), C.byref(cnt)),
                    "mpcgpu_cost_device_ptr")
        return p.value, cnt.value

    def counters(self) -> dict:
        c = _capi.Counters()
        self._check(self.lib.mpcgpu_get_counters(self.h, C.byref(c)), "mpcgpu_get_counters")
        return c.asdict()

    def set_option(self, option: int, value: int):
        """e.g. set_option(OPT_VNS_LEGALITY, 1): VNS2.m:135 legality (N <= dmin, Nu <= 1 -> status 4) inside the library."""
        self._check(self.lib.mpcgpu_set_option(self.h, int(option), int(value)), "mpcgpu_set_option")

    def set_mismatch(self, plant=None, gain=None, hl=None):
        """Validation run against a real plant that differs from the model (Shell3x3.m:271-286): `plant` a Channels object
        scaled like the model (mpcgpu.estimator.*_real_plant), `gain` the estimator gain in the state order of
        mpcgpu/estimator.py E1 (default: the restated Toolbox default, default_estimator_gain).  plant=None: nominal again."""
        if plant is None:
            self._check(self.lib.mpcgpu_set_mismatch(self.h, None, None, None, None, None, 0), "mpcgpu_set_mismatch")
            return
        from . import estimator
        hl = estimator.history_length(self.prob, plant) if hl is None else int(hl)
        gain = estimator.default_estimator_gain(self.prob, hl) if gain is None else gain
        keep = [np.ascontiguousarray(x, dtype=np.float64) for x in (plant.a, plant.b0, plant.b1)]
        d = np.ascontiguousarray(plant.d, dtype=np.int32)
        g = np.ascontiguousarray(gain, dtype=np.float64)
        assert g.shape == (self.ny * (self.nu + self.nd) + self.nu * hl + self.ny, self.ny), g.shape
        self._check(self.lib.mpcgpu_set_mismatch(self.h, _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), _ptr(d), _ptr(g), hl), "mpcgpu_set_mismatch")

    def closedloop(self, r, v, N, Nu, delta, lam, nit):
        """mpcgpu_closedloop: one closedloop_toolbox call with per-call signals; the handle's own signals are untouched."""
        nit = int(nit)
        r = np.ascontiguousarray(row2col(r)[:nit])
        vv = np.ascontiguousarray(row2col(v)[:nit]) if self.nd else None
        delta = np.ascontiguousarray(delta, dtype=np.float64).reshape(self.ny)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(self.nu)
        y, ys = np.empty((self.ny, nit)), np.empty((self.ny, nit))
        u, uopt = np.empty((self.nu, nit)), np.empty((self.nu, nit))
        st = np.zeros(1, dtype=np.int32)
        self._check(self.lib.mpcgpu_closedloop(self.h, nit, _ptr(r), _ptr(vv), int(N), int(Nu), _ptr(delta), _ptr(lam),
                                               _ptr(y), _ptr(u), _ptr(ys), _ptr(uopt), _ptr(st)), "mpcgpu_closedloop")
        return y, u, ys, uopt, int(st[0])


class MultiEvaluator:
    """`mpcgpu_multi`: the same problem on several B200s of one box, driven from ONE process (what the MEX gateway uses).
    `eval_batch` deals the population by estimated work, runs every device concurrently and returns the costs in
    population order -- bit-identical to a single-device `Evaluator.eval_batch`."""

    def __init__(self, prob: LinearProblem, devices=None, r=None, v=None, yref=None, nit=None):
        self.lib = _capi.load_library()
        self.prob = prob
        ps, self._keep = _capi.make_problem_struct(prob, r=r, v=v, yref=yref, nit=nit)
        self.ny, self.nu, self.nd, self.nit = prob.ny, prob.nu, prob.nd, ps.nit
        if devices is None:
            devices = list(range(self.lib.mpcgpu_device_count()))
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        rc = self.lib.mpcgpu_create_multi(C.byref(ps), _ptr(dv), len(dv), C.byref(h))
        if rc != 0:
            raise MpcGpuError(f"mpcgpu_create_multi failed ({rc}): {self.lib.mpcgpu_multi_last_error(None).decode()}")
        self.h = h
        self.devices = list(map(int, dv))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpcgpu_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, option: int, value: int):
        rc = self.lib.mpcgpu_multi_set_option(self.h, int(option), int(value))
        if rc != 0:
            raise MpcGpuError(f"mpcgpu_multi_set_option failed ({rc}): {self.lib.mpcgpu_multi_last_error(self.h).decode()}")

    def eval_batch(self, N, Nu, delta, lam, mode="gam"):
        N = np.ascontiguousarray(np.atleast_1d(N), dtype=np.int32)
        Nu = np.ascontiguousarray(np.atleast_1d(Nu), dtype=np.int32)
        n = N.shape[0]
        delta = np.ascontiguousarray(delta, dtype=np.float64).reshape(n, self.ny)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(n, self.nu)
        m = _capi._MODES[mode]
        cost = np.empty((n, self.ny)) if m == _capi.COST_GAM else np.empty(n)
        status = np.zeros(n, dtype=np.int32)
        rc = self.lib.mpcgpu_multi_eval_batch(self.h, n, _ptr(N), _ptr(Nu), _ptr(delta), _ptr(lam), m, _ptr(cost), _ptr(status))
        if rc != 0:
            raise MpcGpuError(f"mpcgpu_multi_eval_batch failed ({rc}): {self.lib.mpcgpu_multi_last_error(self.h).decode()}")
        return {"cost": cost, "status": status}

    def counters(self, device_index: int = 0) -> dict:
        c = _capi.Counters()
        self.lib.mpcgpu_multi_get_counters(self.h, int(device_index), C.byref(c))
        return c.asdict()


def measure_fp64_peak(device: int = -1) -> float:
    lib = _capi.load_library()
    tf = C.c_double()
    rc = lib.mpcgpu_measure_fp64_peak(int(device), C.byref(tf))
    if rc != 0:
        raise MpcGpuError(f"mpcgpu_measure_fp64_peak failed ({rc})")
    return tf.value


# ---------------------------------------------------------------------------------------------
# Reference-shaped functions
# ---------------------------------------------------------------------------------------------
def closedloop_toolbox(mpc_toolbox, r, v, N, Nu, delta, lam, nit):
    """[y,u,t,ys,uopt] = closedloop_toolbox(mpc_toolbox,r,v,N,Nu,delta,lambda,nit)
    (MPC_Tuning/closedloop_toolbox.m:1).  `mpc_toolbox` is an `Evaluator` (the stand-in for the scaled
    `mpc` object) or a `LinearProblem`.  N and Nu may be vectors; their max is used (:38-40).
    Outputs are signals x time (:103-107).  Raises MpcGpuError where the Toolbox would throw."""
    ev = mpc_toolbox if isinstance(mpc_toolbox, Evaluator) else Evaluator(mpc_toolbox)
    nit = int(nit)
    r = row2col(r)
    vv = row2col(v) if ev.nd else None
    if r.shape[0] < nit or r.shape[1] != ev.ny:
        raise MpcGpuError(f"set-point must be {nit} x {ev.ny} (either orientation)")
    # one ABI call with per-call signals: the evaluator's own Xsp / mdv / Yref / nit (the tuner state `Par`) stay as they are
    y, u, ys, uopt, st = ev.closedloop(r, vv, int(np.max(N)), int(np.max(Nu)), delta, lam, nit)
    if st != 0:
        raise MpcGpuError(f"closed-loop simulation failed with status {st}")
    t = np.arange(nit)[None, :] * ev.prob.Ts
    return y, u, t, ys, uopt


def gam_fun(X, par: Evaluator):
    """[g,h] = GAM_fun(X,Par) (MPC_Tuning/GAM_fun.m:1): X = [delta lambda]; |.| is taken (:56-57), band
    outputs keep delta = 0 (:62-66).  X may be 2-D (n x (ny+nu)) for a batched finite-difference step.
    Returns (g, h): g is ny (or n x ny), h is empty.  A failed simulation yields NaN like a stale-data
    `catch` would not; callers that need the reference's print-and-continue can test isnan."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    ny, nu = par.ny, par.nu
    delta = np.abs(X[:, :ny]).copy()
    lam = np.abs(X[:, ny:ny + nu])
    delta[:, par.prob.band_mask] = 0.0
    n = X.shape[0]
    N = np.full(n, int(np.max(par.N)), dtype=np.int32)
    Nu = np.full(n, int(np.max(par.Nu)), dtype=np.int32)
    out = par.eval_batch(N, Nu, delta, lam, mode="gam")
    g = out["cost"]
    return (g[0] if n == 1 else g), np.zeros(0)


def vns_cost(par: Evaluator, N, Nu, delta, lam):
    """The objective evaluated inside VNS2.m:147-195 for horizon candidates (N, Nu) at fixed weights;
    batched over candidates.  Illegal horizons (VNS2.m:135, PreCon.m:23) return +inf, which is how the
    search's 'reject and un-flip' branch reads."""
    N = np.atleast_1d(np.asarray(N, dtype=np.int32))
    Nu = np.atleast_1d(np.asarray(Nu, dtype=np.int32))
    n = N.shape[0]
    delta = np.broadcast_to(np.asarray(delta, float), (n, par.ny))
    lam = np.broadcast_to(np.asarray(lam, float), (n, par.nu))
    legal = np.array([par.prob.valid(int(a), int(b)) for a, b in zip(N, Nu)])
    F = np.full(n, np.inf)
    if legal.any():
        out = par.eval_batch(N[legal], Nu[legal], delta[legal], lam[legal], mode="vns")
        F[legal] = out["cost"]
    return F
