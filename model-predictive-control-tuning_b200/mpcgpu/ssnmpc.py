"""Single-shooting NMPC of the reference's `Explicit NMPC/` demo: the problem of main.m as
data and the host mirror of ClosedLoopNMPC.m on libmpcgpu.so (kernel k_ssnmpc, csrc/mpc_ssnmpc_core.h: S1-S5)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .nmpc import steady_state


@dataclass
class SsnmpcProblem:
    name: str
    Ts: float
    nit: int
    x0: np.ndarray            # steady state at u0 (main.m:24-39)
    u0: np.ndarray
    lb: np.ndarray; ub: np.ndarray            # MV bounds (main.m:44-51)
    x_control: np.ndarray     # 0-based indices of the controlled states (main.m:75: [2 3])
    r: np.ndarray             # ny x nit set-point (main.m:55-56)
    inK: int = 4              # main.m:52
    pmax: int = 31
    nsub: int = 4
    max_sqp: int = 400        # Gauss-Newton converges linearly after a set-point jump (large residual, W ~ 1e-4): up to ~160 iterations
    ny: int = 2; nu: int = 2; nx: int = 3


def explicit_nmpc() -> SsnmpcProblem:
    Ts, nit = 0.05, 150                                                                   # main.m:20-21
    u0 = np.array([20.0, 130.0])                                                          # :31
    x0 = steady_state(u0, [5.1, 1.1163, 130.0])                                           # :24-39 (fsolve)
    r = np.zeros((2, nit))
    r[0, :10] = x0[1]; r[0, 9:] = 1.2; r[0, 49:] = 1.0                                    # :55
    r[1, :] = x0[2]; r[1, 80:] = 130.0; r[1, 110:] = 120.0                                # :56
    return SsnmpcProblem("Explicit-NMPC", Ts, nit, x0, u0, np.array([0.0, 40.0]), np.array([150.0, 150.0]),
                         np.array([1, 2], dtype=np.int32), r)


# main.m:59-62
BASE_N, BASE_NU, BASE_Q, BASE_W = 5, (2, 2), (1.0214, 0.9999), (1.0e-4, 1.0e-4)


def synthetic_ssnmpc_population(prob: SsnmpcProblem, n: int, seed: int = 0, nmax: int = 12, numax: int = 4):
    """A sweep around main.m:59-62: N ~ U{2..nmax}, Nu_j ~ U{1..min(numax, N)} per input, Q log-uniform [0.1, 10],
    W log-uniform [1e-5, 1e-2]; candidate 0 is the reference's own setting."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = rng.integers(2, nmax + 1, size=n).astype(np.int32)
    Nu = np.stack([[rng.integers(1, min(numax, int(a)) + 1) for _ in range(prob.nu)] for a in N]).astype(np.int32)
    Q = np.exp(rng.uniform(np.log(0.1), np.log(10.0), size=(n, prob.ny)))
    W = np.exp(rng.uniform(np.log(1e-5), np.log(1e-2), size=(n, prob.nu)))
    N[0] = BASE_N; Nu[0] = BASE_NU; Q[0] = BASE_Q; W[0] = BASE_W
    return N, Nu, Q, W


class SsnmpcProblemStruct(C.Structure):
    """mpcgpu_ssnmpc_problem (include/mpcgpu.h)."""
    _fields_ = [("nit", C.c_int32), ("pmax", C.c_int32), ("inK", C.c_int32), ("nsub", C.c_int32), ("max_sqp", C.c_int32),
                ("model", C.c_int32), ("x_control", C.c_int32 * 2), ("Ts", C.c_double)] + \
               [(k, C.c_void_p) for k in ("x0", "u0", "lb", "ub", "r")]


class SsnmpcEvaluator:
    """One `mpcgpu_ssnmpc_handle`: batched ClosedLoopNMPC runs on one B200 (no CPU fallback)."""

    def __init__(self, prob: SsnmpcProblem, device: int = -1):
        from . import _capi
        from .api import MpcGpuError
        self._err = MpcGpuError
        self.lib = _capi.load_library()
        self.prob = prob
        self.ny, self.nu, self.nx, self.nit = prob.ny, prob.nu, prob.nx, int(prob.nit)
        f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        self._keep = dict(x0=f64(prob.x0), u0=f64(prob.u0), lb=f64(prob.lb), ub=f64(prob.ub), r=f64(prob.r))
        ps = SsnmpcProblemStruct(self.nit, int(prob.pmax), int(prob.inK), int(prob.nsub), int(prob.max_sqp), 0,
                                 (C.c_int32 * 2)(*[int(a) for a in prob.x_control]), float(prob.Ts))
        for k, a in self._keep.items():
            setattr(ps, k, a.ctypes.data)
        h = C.c_void_p()
        rc = self.lib.mpcgpu_ssnmpc_create(C.byref(ps), int(device), C.byref(h))
        if rc != 0:
            raise MpcGpuError(f"mpcgpu_ssnmpc_create failed ({rc}): {self.lib.mpcgpu_ssnmpc_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpcgpu_ssnmpc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counters(self) -> dict:
        from . import _capi
        c = _capi.Counters()
        self.lib.mpcgpu_ssnmpc_get_counters(self.h, C.byref(c))
        return c.asdict()

    def eval_batch(self, N, Nu, Q, W, traj=False, r=None, noise=None):
        """Host arrays in/out.  N: n; Nu: n x nu (per-input control horizons); Q: n x ny; W: n x nu.  r (ny x nit) overrides
        the problem's set-point, noise (nx x nit) is added to the plant state after each step (ClosedLoopNMPC.m:88-90).
        Returns cost (n x ny: squared tracking error over k = inK..nit), status and, with traj, y and u (n x 2 x nit)."""
        N = np.ascontiguousarray(np.atleast_1d(N), dtype=np.int32); n = N.shape[0]
        Nu = np.ascontiguousarray(Nu, dtype=np.int32).reshape(n, self.nu)
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(n, self.ny)
        W = np.ascontiguousarray(W, dtype=np.float64).reshape(n, self.nu)
        cost = np.empty((n, self.ny)); status = np.zeros(n, dtype=np.int32)
        y = np.empty((n, self.ny, self.nit)) if traj else None
        u = np.empty((n, self.nu, self.nit)) if traj else None
        rr = None if r is None else np.ascontiguousarray(r, dtype=np.float64).reshape(self.ny, self.nit)
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64).reshape(self.nx, self.nit)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = self.lib.mpcgpu_ssnmpc_eval_batch(self.h, n, ptr(N), ptr(Nu), ptr(Q), ptr(W), ptr(rr), ptr(nz), ptr(cost), ptr(y),
                                               ptr(u), ptr(status))
        if rc != 0:
            raise self._err(f"mpcgpu_ssnmpc_eval_batch failed ({rc}): {self.lib.mpcgpu_ssnmpc_last_error(self.h).decode()}")
        out = {"cost": cost, "status": status}
        if traj:
            out.update(y=y, u=u)
        return out


def ClosedLoopNMPC(ev, x_control, u0, r, N, Nu, Q, W, nit, ub1, lb1, inK, Ts, noise=None):
    """[y, u] = ClosedLoopNMPC(x0_model, x_control, u0, r, N, Nu, Q, W, nit, ub1, lb1, inK, Ts)  (ClosedLoopNMPC.m:1).
    `ev` (an SsnmpcEvaluator) stands in for x0_model: the handle carries the model, its steady state, bounds and timing,
    and the call checks that the remaining arguments agree with it.  x_control is 1-based as in MATLAB."""
    p = ev.prob
    same = (int(nit) == ev.nit and int(inK) == int(p.inK) and abs(float(Ts) - p.Ts) < 1e-15
            and np.array_equal(np.asarray(x_control, int) - 1, np.asarray(p.x_control, int))
            and np.allclose(u0, p.u0, rtol=0, atol=0) and np.allclose(ub1, p.ub, rtol=0, atol=0) and np.allclose(lb1, p.lb, rtol=0, atol=0))
    if not same:
        raise ev._err("ClosedLoopNMPC: nit, inK, Ts, x_control, u0, ub1, lb1 must equal the handle's problem")
    r = np.atleast_2d(np.asarray(r, float))
    if r.shape != (ev.ny, ev.nit):
        raise ev._err(f"set-point must be {ev.ny} x {ev.nit}")
    out = ev.eval_batch([int(np.atleast_1d(N)[0])], np.asarray(Nu, int).reshape(1, -1), np.asarray(Q, float).reshape(1, -1),
                        np.asarray(W, float).reshape(1, -1), traj=True, r=r, noise=noise)
    if out["status"][0] != 0:
        raise ev._err(f"closed-loop simulation failed with status {int(out['status'][0])}")
    return out["y"][0], out["u"][0]
