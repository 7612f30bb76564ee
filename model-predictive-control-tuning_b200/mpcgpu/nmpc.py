"""Nonlinear case (BASELINE.json configs[4]): the Van de Vusse CSTR NMPC tuning problem of
MPC-Tuning/VanDeVusse_NMPC.m as data, and the host mirror of closedloop_toolbox_nmpc.m on libmpcgpu.so."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np


@dataclass
class NmpcProblem:
    name: str
    Ts: float
    nit: int
    x0: np.ndarray          # steady state at u0 (VanDeVusse_NMPC.m:64-79)
    u0: np.ndarray
    umin: np.ndarray; umax: np.ndarray        # MV bounds (:45-56)
    xmin: np.ndarray; xmax: np.ndarray        # state bounds; outputs are states 2..3 (closedloop_toolbox_nmpc.m:73)
    su: np.ndarray; sy: np.ndarray            # ScaleFactors = ranges (:139-151)
    r: np.ndarray           # ny x nit set-point (:89-90)
    yref: np.ndarray        # ny x nit reference trajectory (:160-176)
    nbp: int = 5; nbc: int = 4                # MPCTuning(...,5,4,...) (:204): N <= 31, Nu <= 15
    inK: int = 10                             # VNS2.m:43
    nsub: int = 4                             # RK4 sub-steps per sample
    max_sqp: int = 30
    ny: int = 2; nu: int = 2; nx: int = 3
    band_mask: np.ndarray = field(default_factory=lambda: np.zeros(2, dtype=bool))

    def valid(self, N: int, Nu: int) -> bool:
        """VNS2.m:135 / PreCon.m:23 with zero dead times."""
        return 0 < Nu < N <= 2 ** self.nbp - 1 and Nu <= 2 ** self.nbc - 1 and Nu > 1


def _vdv_rhs(x, u):
    # vandevusse_model.m:39-77
    k10, k20, k30 = 1.287e12, 1.287e12, 9.043e9
    e1, e2, e3 = -9758.3, -9758.3, -8560.0
    dab, dbc, dad = -4.20, 11.00, 41.85
    rho, cp, kw, ar, vol, t0, ca0 = 0.9342, 3.01, 4032.0, 0.215, 10.0, 130.0, 5.10
    fov, tk = u
    ca, cb, t = x
    k1 = k10 * np.exp(e1 / (t + 273.15)); k2 = k20 * np.exp(e2 / (t + 273.15)); k3 = k30 * np.exp(e3 / (t + 273.15))
    return np.array([fov * (ca0 - ca) - k1 * ca - k3 * ca * ca, -fov * cb + k1 * ca - k2 * cb,
                     (k1 * ca * dab + k2 * cb * dbc + k3 * ca * ca * dad) / (rho * cp) + fov * (t0 - t) + kw * ar / (rho * cp * vol) * (tk - t)])


def steady_state(u0, guess):
    """`fsolve(@(x) model(ts,x,u0), X0)` (VanDeVusse_NMPC.m:79): Newton with a finite-difference Jacobian."""
    x = np.array(guess, float)
    for _ in range(50):
        f = _vdv_rhs(x, u0)
        J = np.zeros((3, 3))
        for k in range(3):
            h = 1e-7 * max(1.0, abs(x[k]))
            xp = x.copy(); xp[k] += h
            J[:, k] = (_vdv_rhs(xp, u0) - f) / h
        dx = np.linalg.solve(J, -f)
        x = x + dx
        if np.abs(dx).max() < 1e-13 * max(1.0, np.abs(x).max()):
            break
    return x


def vandevusse() -> NmpcProblem:
    Ts, nit = 0.05, 60                                                                   # :35-36
    umin = np.array([0.0, 40.0]); umax = np.array([150.0, 150.0])                        # :49-56
    xmin = np.array([0.0, 0.0, 40.0]); xmax = np.array([6.0, 1.2, 150.0])                # :45-48,58-59
    u0 = np.array([20.0, 130.0])                                                         # :71
    x0 = steady_state(u0, [5.1, 1.1163, 130.0])                                          # :65-79
    r = np.zeros((2, nit))
    r[0, :] = x0[1]; r[0, 9:] = 1.0                                                      # :89 (1-based 10:nit)
    r[1, :] = x0[2]; r[1, 40:] = 130.0                                                   # :90 (41:nit)
    tau = np.array([0.05, 0.0875])                                                       # :159 Pref (fast)
    a = np.exp(-Ts / tau)
    xsp = r - x0[1:3, None]                                                              # :171
    yr = np.zeros((2, nit))
    for k in range(1, nit):                                                              # lsim(Pref, Xspref, t, 'zoh') :176
        yr[:, k] = a * yr[:, k - 1] + (1 - a) * xsp[:, k - 1]
    yref = yr + x0[1:3, None]
    return NmpcProblem("VanDeVusse-NMPC", Ts, nit, x0, u0, umin, umax, xmin, xmax, umax - umin, (xmax - xmin)[1:3], r, yref)


def synthetic_nmpc_population(prob: NmpcProblem, n: int, seed: int = 0, wlo: float = 1e-3, whi: float = 10.0):
    """N ~ U{3..31}, Nu ~ U{2..min(15, N-1)}, delta, lambda log-uniform [1e-3, 10]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = rng.integers(3, 2 ** prob.nbp, size=n).astype(np.int32)
    Nu = np.array([rng.integers(2, min(2 ** prob.nbc - 1, int(a) - 1) + 1) for a in N], dtype=np.int32)
    delta = np.exp(rng.uniform(np.log(wlo), np.log(whi), size=(n, prob.ny)))
    lam = np.exp(rng.uniform(np.log(wlo), np.log(whi), size=(n, prob.nu)))
    return N, Nu, delta, lam


class NmpcProblemStruct(C.Structure):
    """mpcgpu_nmpc_problem (include/mpcgpu.h)."""
    _fields_ = [("nit", C.c_int32), ("pmax", C.c_int32), ("mmax", C.c_int32), ("inK", C.c_int32),
                ("nsub", C.c_int32), ("max_sqp", C.c_int32), ("model", C.c_int32), ("reserved", C.c_int32),
                ("Ts", C.c_double)] + [(k, C.c_void_p) for k in ("x0", "u0", "umin", "umax", "xmin", "xmax", "su", "sy", "r", "yref")]


class NmpcEvaluator:
    """One `mpcgpu_nmpc_handle`: batched closedloop_toolbox_nmpc + GAM / VNS costs on one B200 (no CPU fallback)."""

    def __init__(self, prob: NmpcProblem, device: int = -1):
        from . import _capi
        from .api import MpcGpuError
        self._err = MpcGpuError
        self.lib = _capi.load_library()
        self.prob = prob
        self.ny, self.nu, self.nit = prob.ny, prob.nu, int(prob.nit)
        self.N, self.Nu = 10, 2                                                   # VanDeVusse_NMPC.m:179-180
        f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        self._keep = dict(x0=f64(prob.x0), u0=f64(prob.u0), umin=f64(prob.umin), umax=f64(prob.umax), xmin=f64(prob.xmin),
                          xmax=f64(prob.xmax), su=f64(prob.su), sy=f64(prob.sy), r=f64(prob.r), yref=f64(prob.yref))
        ps = NmpcProblemStruct(self.nit, 2 ** prob.nbp - 1, 2 ** prob.nbc - 1, prob.inK, prob.nsub, prob.max_sqp, 0, 0, prob.Ts)
        for k, a in self._keep.items():
            setattr(ps, k, a.ctypes.data)
        h = C.c_void_p()
        rc = self.lib.mpcgpu_nmpc_create(C.byref(ps), int(device), C.byref(h))
        if rc != 0:
            raise MpcGpuError(f"mpcgpu_nmpc_create failed ({rc}): {self.lib.mpcgpu_nmpc_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpcgpu_nmpc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    reject_bound_crossing = True   # eval_batch: cost = inf for candidates whose closed loop crossed an OV / state bound (status 5)

    def counters(self) -> dict:
        from . import _capi
        c = _capi.Counters()
        self.lib.mpcgpu_nmpc_get_counters(self.h, C.byref(c))
        return c.asdict()

    def eval_batch(self, N, Nu, delta, lam, mode="gam", traj=False, r=None):
        """Host arrays in/out.  mode 'raw' | 'gam' | 'vns'; r (ny x nit) overrides the problem's set-point for this call."""
        N = np.ascontiguousarray(np.atleast_1d(N), dtype=np.int32); n = N.shape[0]
        Nu = np.ascontiguousarray(np.atleast_1d(Nu), dtype=np.int32)
        delta = np.ascontiguousarray(delta, dtype=np.float64).reshape(n, self.ny)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(n, self.nu)
        m = {"raw": 0, "gam": 1, "vns": 2}[mode]
        cost = np.empty((n, self.ny)) if m == 1 else (np.empty(n) if m == 2 else None)
        status = np.zeros(n, dtype=np.int32)
        tr = [None] * 4
        if traj or m == 0:
            tr = [np.empty((n, self.ny, self.nit)), np.empty((n, self.nu, self.nit)), np.empty((n, self.ny, self.nit)),
                  np.empty((n, self.nu, self.nit))]
        rr = None if r is None else np.ascontiguousarray(r, dtype=np.float64).reshape(self.ny, self.nit)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = self.lib.mpcgpu_nmpc_eval_batch(self.h, n, ptr(N), ptr(Nu), ptr(delta), ptr(lam), m, ptr(rr), ptr(cost),
                                             ptr(tr[0]), ptr(tr[1]), ptr(tr[2]), ptr(tr[3]), ptr(status))
        if rc != 0:
            raise self._err(f"mpcgpu_nmpc_eval_batch failed ({rc}): {self.lib.mpcgpu_nmpc_last_error(self.h).decode()}")
        if cost is not None and self.reject_bound_crossing:
            # status 5: the closed loop crossed an OV / state bound of VanDeVusse_NMPC.m:140-145, which nlmpc would have kept
            # (OV bounds soft with Weights.ECR, state bounds hard) and this solver does not enforce -- such a candidate must not
            # compete with costs from trajectories the Toolbox would have constrained: rejected (inf), like an illegal one
            cost[status == 5] = np.inf
        out = {"cost": cost, "status": status}
        if tr[0] is not None:
            out.update(y=tr[0], u=tr[1], yopt=tr[2], uopt=tr[3])
        return out


def closedloop_toolbox_nmpc(nmpcobj, model, init, r, N, Nu, delta, lam, nit):
    """[y,u,yopt,uopt] = closedloop_toolbox_nmpc(nmpcobj,model,init,r,N,Nu,delta,lambda,nit)
    (MPC_Tuning/closedloop_toolbox_nmpc.m:1).  `nmpcobj` is an NmpcEvaluator (stand-in for the nlmpc object); `model`
    and `init` are accepted for signature compatibility (the kernel carries the Van de Vusse model and the
    problem's x0/u0).  N, Nu may be vectors: their max is used (:48-51).  Outputs are signals x time."""
    ev = nmpcobj
    r = np.atleast_2d(np.asarray(r, float))
    if r.shape[0] > r.shape[1]:
        r = r.T
    if r.shape != (ev.ny, int(nit)) or int(nit) != ev.nit:
        raise ev._err(f"set-point must be {ev.ny} x {ev.nit}")
    out = ev.eval_batch([int(np.max(N))], [int(np.max(Nu))], np.asarray(delta, float).reshape(1, -1),
                        np.asarray(lam, float).reshape(1, -1), mode="raw", r=r)
    if out["status"][0] not in (0, 5):
        raise ev._err(f"closed-loop simulation failed with status {int(out['status'][0])}")
    return out["y"][0], out["u"][0], out["yopt"][0], out["uopt"][0]
