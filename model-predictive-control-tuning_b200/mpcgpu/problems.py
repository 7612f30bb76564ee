"""The reference's case studies as data: what MPCTuning.m hands to the closed-loop evaluator.

Each builder restates one case script up to the `MPCTuning(...)` call and then applies the
scaling block of MPC_Tuning/MPCTuning.m:154-200 (plant L*Pz*R, MV limits / R, OV limits * L,
Yref and Xsp * L, mdv / Rv, ScaleFactors only when != 1), so that a `LinearProblem` is exactly the
state the tuner's objective functions (GAM_fun.m, VNS2.m) see in `Par`.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .plant import Channels, c2d_fopdt, simulate

INF = float("inf")


@dataclass
class LinearProblem:
    name: str
    Ts: float
    nit: int
    ny: int
    nu: int
    nd: int
    plant: Channels              # scaled discrete plant, ny x (nu+nd)
    umin: np.ndarray             # MV(i).Min   (scaled)
    umax: np.ndarray
    dumin: np.ndarray            # MV(i).RateMin (scaled), -inf if none
    dumax: np.ndarray
    ymin: np.ndarray             # OV(i).Min (scaled), -inf if none
    ymax: np.ndarray
    ecr_min: np.ndarray          # OV(i).MinECR
    ecr_max: np.ndarray
    su: np.ndarray               # MV ScaleFactor
    sy: np.ndarray               # OV ScaleFactor
    rho_ecr: float               # Weights.ECR seen *during* tuning
    r: np.ndarray                # (nit, ny) scaled set-point, time-major
    v: np.ndarray                # (nit, nd) scaled measured disturbance
    yref: np.ndarray             # (ny, nit) scaled reference trajectory
    dmin: np.ndarray             # (ny,) int32
    band_mask: np.ndarray        # (ny,) bool: OV weight initially 0 => delta forced to 0 (GAM_fun.m:62-66)
    L: np.ndarray
    R: np.ndarray
    inK: int = 10                # VNS2.m:43
    w_pareto: np.ndarray = field(default_factory=lambda: np.zeros(0))
    nbp: int = 7
    nbc: int = 4

    @property
    def square(self) -> bool:
        return self.ny == self.nu

    def vns_setpoint(self) -> np.ndarray:
        """Set-point the VNS objective uses for linear plants: an *unscaled* unit step from
        sample inK (1-based) on every output (VNS2.m:58-61)."""
        r = np.zeros((self.nit, self.ny))
        r[self.inK - 1:, :] = 1.0
        return r

    def valid(self, N: int, Nu: int) -> bool:
        """VNS2.m:135 with PreCon.m:23 for scalar horizons."""
        return bool(N > Nu and N != 0 and Nu != 0 and np.all(N > self.dmin) and Nu > 1)


def _mset(row: np.ndarray, lo: int, hi: int, val: float) -> None:
    """MATLAB `row(lo:hi) = val` with 1-based inclusive indices."""
    row[lo - 1:hi] = val


def _scale_problem(name, Ts, nit, ch: Channels, nu, nd, L, R, umin, umax, dumin, dumax, ymin, ymax,
                   ecr_min, ecr_max, su, sy, rho_ecr, Xsp, mdv, Yref, ov_w0, w_pareto, nbp, nbc):
    """MPCTuning.m:154-200."""
    ny = ch.a.shape[0]
    L = np.asarray(L, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    Ru, Rv = R[:nu], R[nu:]
    pze = ch.scaled(L, R)
    su = np.asarray(su, dtype=np.float64).copy()
    sy = np.asarray(sy, dtype=np.float64).copy()
    su = np.where(su != 1.0, su / Ru, su)          # MPCTuning.m:175-177
    sy = np.where(sy != 1.0, sy * L, sy)           # MPCTuning.m:182-184
    r = (np.asarray(Xsp, dtype=np.float64) * L[:, None]).T.copy()      # (nit, ny)
    yref = np.asarray(Yref, dtype=np.float64) * L[:, None]
    v = (np.asarray(mdv, dtype=np.float64).reshape(nd, nit) / Rv[:, None]).T.copy() if nd else np.zeros((nit, 0))
    return LinearProblem(
        name=name, Ts=Ts, nit=nit, ny=ny, nu=nu, nd=nd, plant=pze,
        umin=np.asarray(umin, float) / Ru, umax=np.asarray(umax, float) / Ru,
        dumin=np.asarray(dumin, float) / Ru, dumax=np.asarray(dumax, float) / Ru,
        ymin=np.asarray(ymin, float) * L, ymax=np.asarray(ymax, float) * L,
        ecr_min=np.asarray(ecr_min, float), ecr_max=np.asarray(ecr_max, float),
        su=su, sy=sy, rho_ecr=float(rho_ecr), r=r, v=v, yref=yref, dmin=pze.dmin(),
        band_mask=(np.asarray(ov_w0, float) == 0.0), L=L, R=R,
        w_pareto=np.asarray(w_pareto, float), nbp=nbp, nbc=nbc)


# ---------------------------------------------------------------------------------------------
# Shell 3x3 heavy-oil fractionator  (MPC-Tuning/Shell3x3.m)
# ---------------------------------------------------------------------------------------------
SHELL3X3_L = np.array([0.43577812475231503, 0.4205588479390135, 0.5932860051199568])
SHELL3X3_R = np.array([0.661867070834956, 0.2756082654542081, 0.41172304878568067])


def shell3x3(caso: int = 2, L=None, R=None) -> LinearProblem:
    """Shell3x3.m:43-163.  L,R default to the diagonals stored in the reference's result files
    (CondMin is not reproducible)."""
    K = np.array([[4.05, 1.77, 5.88], [5.39, 5.72, 6.9], [4.38, 4.42, 7.2]])          # :52-54
    tau = np.array([[50, 60, 50], [50, 60, 40], [33, 44, 19]], dtype=float)
    theta = np.array([[27, 28, 27], [18, 14, 15], [20, 22, 0]], dtype=float)           # :57
    Ts, nit = 4.0, 500                                                                 # :61-62
    pz = c2d_fopdt(K, tau, theta, Ts)                                                  # :65
    tref = np.array([5.0, 9.0, 5.7]) if caso == 1 else np.array([30.0, 30.0, 30.0])    # :71-75
    dref = np.array([27.0, 14.0, 0.0])                                                 # :76
    Xsp = np.zeros((3, nit))                                                           # :89-92
    _mset(Xsp[0], 10, 80, 0.2); _mset(Xsp[0], 80, 200, 0.0); _mset(Xsp[0], 200, 400, 0.1); _mset(Xsp[0], 400, 500, 0.0)
    _mset(Xsp[1], 10, 80, 0.2); _mset(Xsp[1], 80, 200, 0.4); _mset(Xsp[1], 200, 400, 0.3); _mset(Xsp[1], 400, 500, 0.0)
    _mset(Xsp[2], 10, 80, 0.2); _mset(Xsp[2], 80, 200, 0.1); _mset(Xsp[2], 200, 400, 0.0); _mset(Xsp[2], 400, 500, 0.0)
    Yref = np.zeros((3, nit))                                                          # :98-99
    for i in range(3):
        chi = c2d_fopdt([[1.0]], [[tref[i]]], [[dref[i]]], Ts)
        Yref[i] = simulate(chi, Xsp[i][:, None])[:, 0]
    return _scale_problem(
        f"Shell3x3-caso{caso}", Ts, nit, pz, 3, 0,
        SHELL3X3_L if L is None else L, SHELL3X3_R if R is None else R,
        umin=[-1, -1, -1], umax=[0.5, 0.5, 0.5], dumin=[-0.05] * 3, dumax=[0.05] * 3,  # :120-123
        ymin=[-INF] * 3, ymax=[INF] * 3, ecr_min=[1.0] * 3, ecr_max=[1.0] * 3,
        su=[1.0] * 3, sy=[1.0] * 3, rho_ecr=1e5,          # Toolbox default during tuning (MPCTuning.m:354 is post-tuning)
        Xsp=Xsp, mdv=np.zeros((0, nit)), Yref=Yref, ov_w0=[1.0] * 3,
        w_pareto=[0.05, 0.40, 0.55], nbp=7, nbc=4)                                     # :161-163


# ---------------------------------------------------------------------------------------------
# Wood-Berry column with one measured disturbance  (MPC-Tuning/WoodBerry.m)
# ---------------------------------------------------------------------------------------------
# L,R: no result file exists for this case and CondMin on the 2x3 gain [G D] is degenerate (the
# minimiser zeroes a column).  These are plant.cond_min() of the 2x2 MV block (SLSQP, X0=0.1; reaches
# cond 5.867 from 7.481), Rv = 1, recorded so the problem is reproducible bit-for-bit.  They are
# *inputs* to the hot path, not results of it.
WOODBERRY_L = np.array([0.1669690850920105, 0.22950868976642672])
WOODBERRY_R = np.array([0.2988184790052876, 0.1434341550021489, 1.0])


def woodberry(caso: int = 1, L=None, R=None) -> LinearProblem:
    """WoodBerry.m:44-156."""
    K = np.array([[12.8, -18.9, 3.8], [6.6, -19.4, 4.9]])                               # :49-53
    tau = np.array([[16.7, 21.0, 14.9], [10.9, 14.4, 13.2]])
    theta = np.array([[1.0, 2.0, 8.1], [2.0, 1.0, 3.4]])
    Ts, nit = 1.0, 400                                                                  # :56-57
    pz = c2d_fopdt(K, tau, theta, Ts)
    tref = np.array([10.0, 7.0]) if caso == 1 else np.array([15.0, 12.0])               # :68-72
    dref = np.array([1.0, 1.0])                                                         # :74
    Xsp = np.zeros((2, nit))
    _mset(Xsp[0], 10, nit, 0.8)                                                         # :87-89
    _mset(Xsp[1], 200, nit, 0.5)
    mdv = np.zeros((1, nit))
    _mset(mdv[0], 300, nit, -0.25)                                                      # :93-94
    Yref = np.zeros((2, nit))
    for i in range(2):
        chi = c2d_fopdt([[1.0]], [[tref[i]]], [[dref[i]]], Ts)
        Yref[i] = simulate(chi, Xsp[i][:, None])[:, 0]
    return _scale_problem(
        f"WoodBerry-caso{caso}", Ts, nit, pz, 2, 1,
        WOODBERRY_L if L is None else L, WOODBERRY_R if R is None else R,
        umin=[-0.5, -0.5], umax=[0.5, 0.5], dumin=[-0.05, -0.05], dumax=[0.05, 0.05],   # :119-122
        ymin=[-INF] * 2, ymax=[INF] * 2, ecr_min=[1.0] * 2, ecr_max=[1.0] * 2,
        su=[1.0] * 2, sy=[1.0] * 2, rho_ecr=1e5,
        Xsp=Xsp, mdv=mdv, Yref=Yref, ov_w0=[1.0] * 2, w_pareto=[0.1, 0.5], nbp=7, nbc=4)  # :154-156


# ---------------------------------------------------------------------------------------------
# Shell 7x5 (3 MV + 2 MD, 7 outputs, band control)  (MPC-Tuning/Shell7x5.m)
# ---------------------------------------------------------------------------------------------
SHELL7X5_L = np.array([0.4400615063022943, 0.2319273262887009, 0.6265090010777253, 0.5431290766409146,
                       0.6006058918173808, 0.20692945405215463, 0.39416907820719865])
SHELL7X5_R = np.array([0.2639712478155768, 0.1350971290956903, 0.1156440799331315,
                       0.781865375367461, 0.4665315477471682])


def shell7x5(L=None, R=None) -> LinearProblem:
    """Shell7x5.m:46-204."""
    Kg = np.array([[4.05, 1.77, 5.88], [5.39, 5.72, 6.9], [3.66, 1.65, 5.53], [5.92, 2.54, 8.10],
                   [4.13, 2.38, 6.23], [4.06, 4.18, 6.53], [4.38, 4.42, 7.2]])                      # :69-75
    tg = np.array([[50, 60, 50], [50, 60, 40], [9, 30, 40], [12, 27, 20], [8, 19, 10], [13, 33, 9], [33, 44, 19]], float)
    dg = np.array([[27, 28, 27], [18, 14, 15], [2, 20, 2], [11, 12, 2], [5, 7, 2], [8, 4, 1], [20, 22, 0]], float)  # :76
    Kd = np.array([[1.20, 1.44], [1.52, 1.83], [1.16, 1.27], [1.73, 1.79], [1.31, 1.26], [1.19, 1.17], [1.14, 1.26]])  # :79-85
    td = np.array([[45, 40], [25, 20], [11, 6], [5, 19], [2, 22], [19, 24], [24, 32]], float)
    dd = np.array([[27, 27], [15, 15], [0, 0], [0, 0], [0, 0], [0, 0], [0, 0]], float)               # :86
    Ts, nit = 4.0, 200                                                                                  # :91-92
    pz = c2d_fopdt(np.hstack([Kg, Kd]), np.hstack([tg, td]), np.hstack([dg, dd]), Ts)
    Ymn = np.array([-0.005, -0.005, -0.5, -0.5, -0.5, -0.5, -0.5])                                     # :102-103
    Ymx = -Ymn
    Umx = np.array([0.5, 0.5, 0.5])                                                                     # :106-107
    dref = np.hstack([dg, dd]).min(axis=1)                                                              # :113
    Xsp = np.zeros((7, nit))                                                                            # :118
    mdv = np.zeros((2, nit))
    _mset(mdv[0], 20, nit, 0.5); _mset(mdv[1], 20, nit, 0.5)                                            # :121-123
    Xref = np.zeros((7, nit))
    for i in range(7):
        _mset(Xref[i], 20, 25, Ymx[i])                                                                  # :130-132
    Yref = np.zeros((7, nit))
    for i in range(7):
        chi = c2d_fopdt([[1.0]], [[50.0]], [[dref[i]]], Ts)                                             # :110-113
        Yref[i] = simulate(chi, Xref[i][:, None])[:, 0]
    ecr = np.array([0.1, 0.5, 1, 1, 1, 1, 1], float)                                                    # :155-165
    return _scale_problem(
        "Shell7x5", Ts, nit, pz, 3, 2,
        SHELL7X5_L if L is None else L, SHELL7X5_R if R is None else R,
        umin=-Umx, umax=Umx, dumin=[-INF] * 3, dumax=[INF] * 3,
        ymin=Ymn, ymax=Ymx, ecr_min=ecr, ecr_max=ecr,
        su=Umx - (-Umx), sy=Ymx - Ymn,                                                                  # :170-179
        rho_ecr=1e4,                                                                                    # :189
        Xsp=Xsp, mdv=mdv, Yref=Yref, ov_w0=[0.0] * 7,                                                   # :188
        w_pareto=[0.0001, 0.0001, 1, 0.5, 1, 0.5, 1], nbp=7, nbc=4)                                     # :195-198


CASES = {"shell3x3": shell3x3, "woodberry": woodberry, "shell7x5": shell7x5}


# ---------------------------------------------------------------------------------------------
# Synthetic candidate populations
# ---------------------------------------------------------------------------------------------
def synthetic_population(prob: LinearProblem, n: int, seed: int = 0, fixed=None,
                         wlo: float = 1e-4, whi: float = 10.0):
    """Seeded random tuning candidates, all legal under VNS2.m:135.
    Returns (N int32[n], Nu int32[n], delta f64[n,ny], lambda f64[n,nu]).
    `fixed=(p,m)` pins every candidate's horizons (peak-rate / tuned-point populations)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pmax, mmax = 2 ** prob.nbp - 1, 2 ** prob.nbc - 1
    pmin = int(max(prob.dmin.max() + 1, 3))
    if fixed is None:
        N = rng.integers(pmin, pmax + 1, size=n).astype(np.int32)
        Nu = np.array([rng.integers(2, min(mmax, int(p) - 1) + 1) for p in N], dtype=np.int32)
    else:
        N = np.full(n, fixed[0], dtype=np.int32)
        Nu = np.full(n, fixed[1], dtype=np.int32)
    delta = np.exp(rng.uniform(np.log(wlo), np.log(whi), size=(n, prob.ny)))
    lam = np.exp(rng.uniform(np.log(wlo), np.log(whi), size=(n, prob.nu)))
    delta[:, prob.band_mask] = 0.0
    return N, Nu, delta, lam
