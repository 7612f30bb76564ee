"""Plant ingest for the tuning hot path: FOPDT transfer matrices -> scaled discrete channels.

Every plant in the reference's case studies is a matrix of first-order-plus-dead-time
channels  K/(tau*s+1)*exp(-theta*s)  (MPC-Tuning/Shell3x3.m:52-58, WoodBerry.m:49-53,
Shell7x5.m:69-88).  `c2d(Ps,Ts,'zoh')` (Shell3x3.m:65) of such a channel has the closed form

    y(k) = a*y(k-1) + b0*u(k-d) + b1*u(k-d-1)
    a = exp(-Ts/tau),  d = ceil(theta/Ts),  f = d*Ts - theta  (0 <= f < Ts)
    b0 = K*(1 - exp(-f/tau)),   b1 = K*(exp(-f/tau) - a)

which is what MATLAB stores as num=[b0 b1], den=[1 -a], IODelay=d.  The closed form is checked
against the reference's own saved objects in tests/test_plant_kats.py (all 9 + 35 channels).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Channels:
    """ny x nw matrix of discrete first-order channels (nw = inputs: MVs then MDs)."""
    a: np.ndarray
    b0: np.ndarray
    b1: np.ndarray
    d: np.ndarray  # int32 sample delay

    @property
    def shape(self):
        return self.a.shape

    def scaled(self, L: np.ndarray, R: np.ndarray) -> "Channels":
        """Pze = L*Pz*R with diagonal L (ny) and R (nw)  (MPCTuning.m:165)."""
        g = np.outer(L, R)
        return Channels(self.a.copy(), self.b0 * g, self.b1 * g, self.d.copy())

    def dcgain(self) -> np.ndarray:
        return (self.b0 + self.b1) / (1.0 - self.a)

    def descomp_delay(self) -> np.ndarray:
        """Delay matrix as DTC-GPC/descompMPC.m:33-38 reports it: if the leading numerator
        coefficient is non-zero the delay is reduced by one (a zero is prepended)."""
        return np.where(self.b0 != 0.0, self.d - 1, self.d).astype(np.int32)

    def dmin(self) -> np.ndarray:
        """Per-output minimum dead time used by the VNS legality test (MPCTuning.m:257-262)."""
        return self.descomp_delay().min(axis=1).astype(np.int32)


def c2d_fopdt(K, tau, theta, Ts) -> Channels:
    K = np.asarray(K, dtype=np.float64)
    tau = np.asarray(tau, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    a = np.exp(-Ts / tau)
    # ceil with a guard so that theta an exact multiple of Ts stays on that multiple
    d = np.ceil(theta / Ts - 1e-12).astype(np.int32)
    f = d * Ts - theta
    f = np.where(np.abs(f) < 1e-12, 0.0, f)
    e = np.exp(-f / tau)
    b0 = K * (1.0 - e)
    b1 = K * (e - a)
    b0 = np.where(f == 0.0, 0.0, b0)
    return Channels(a, b0, b1, d)


def simulate(ch: Channels, w: np.ndarray) -> np.ndarray:
    """lsim of the discrete channel matrix: w is (nit, nw) -> y (nit, ny). Zero initial state,
    signals before k=0 are zero.  This is `lsim(Pz,[u v],t)` (closedloop_toolbox.m:100) and,
    with a diagonal reference model, `lsim(Pref,Xsp,t,'zoh')` (Shell3x3.m:99)."""
    w = np.asarray(w, dtype=np.float64)
    nit, nw = w.shape
    ny = ch.a.shape[0]
    assert ch.a.shape[1] == nw
    pad = int(ch.d.max()) + 1
    wp = np.vstack([np.zeros((pad, nw)), w])  # wp[pad + k] = w[k]
    x = np.zeros((ny, nw))
    y = np.zeros((nit, ny))
    jj = np.arange(nw)[None, :].repeat(ny, 0)
    for k in range(nit):
        u0 = wp[pad + k - ch.d, jj]
        u1 = wp[pad + k - ch.d - 1, jj]
        x = ch.a * x + ch.b0 * u0 + ch.b1 * u1
        y[k] = x.sum(axis=1)
    return y


def cond_min(Km: np.ndarray, seed_x0: float = 0.1):
    """Diagonal scaling that minimises cond(L*K*R) from X0=0.1 with bounds [0,1]
    (MPC_Tuning/CondMin.m:31-66).  The reference uses fmincon's default interior-point; the
    minimiser is a scale family, so this SLSQP result is *a* minimiser, not MATLAB's.
    """
    from scipy.optimize import minimize

    m, n = Km.shape

    def fobj(X):
        return np.linalg.cond(np.diag(X[:m]) @ Km @ np.diag(X[m:]))

    x0 = np.full(m + n, seed_x0)
    res = minimize(fobj, x0, method="SLSQP", bounds=[(1e-6, 1.0)] * (m + n),
                   options={"maxiter": 500, "ftol": 1e-12})
    return res.x[:m].copy(), res.x[m:].copy(), float(res.fun)
