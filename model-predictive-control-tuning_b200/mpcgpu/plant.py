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


# ---------------------------------------------------------------------------------------------
# General-order ingest: c2d(tf(num,den,'iodelay',theta),Ts,'zoh') and descompMPC
# ---------------------------------------------------------------------------------------------
def c2d_tf(num, den, theta: float, Ts: float):
    """ZOH discretisation of  num(s)/den(s) * exp(-theta s)  of ANY order with a fractional dead time, what
    `c2d(Ps,Ts,'zoh')` (Shell3x3.m:65, DTC_GPC_WW.m) returns for one channel: (bz, az, d) with
        y(k) = -az[1] y(k-1) - ... + bz[0] u(k-d) + bz[1] u(k-d-1) + ...      (az[0] = 1),
    d = ceil(theta/Ts) whole samples and the fraction f = d Ts - theta absorbed in the numerator (modified z-transform:
    over one sample the delayed input is u(k-1) for f_bar = Ts - f seconds ... then u(k)), which adds one numerator tap.
    For a first-order channel this is c2d_fopdt's closed form (tests/test_plant_kats.py)."""
    from scipy.linalg import expm
    from scipy.signal import tf2ss
    num = np.atleast_1d(np.asarray(num, float)); den = np.atleast_1d(np.asarray(den, float))
    A, B, Cc, D = tf2ss(num, den)
    n = A.shape[0]
    assert abs(D[0, 0]) == 0.0, "strictly proper channels only (every plant of the reference is)"
    d = int(np.ceil(theta / Ts - 1e-12))
    f = d * Ts - theta
    if abs(f) < 1e-12:
        f = 0.0

    def phi_gamma(t):   # [Phi(t), Gamma(t)] = expm([[A, B], [0, 0]] t)
        Mx = np.zeros((n + 1, n + 1)); Mx[:n, :n] = A; Mx[:n, n:] = B
        E = expm(Mx * t)
        return E[:n, :n], E[:n, n:]

    Phi, Gam = phi_gamma(Ts)
    # With v(t) = u(t - theta) the undelayed dynamics see u(k-d) during the first Ts - f seconds of sample k and u(k-d+1) during
    # the last f seconds:  x(k+1) = Phi x(k) + G0 u(k-d) + G1 u(k-d+1),  G1 = Gamma(f),  G0 = Gamma(Ts) - Gamma(f).
    _, G1 = phi_gamma(f)
    G0 = Gam - G1
    az = np.poly(Phi)                                 # 1, a1, ..., an

    def num_of(G):                                    # C adj(zI - Phi) G = det(zI - Phi + G C) - det(zI - Phi): n coefficients
        return (np.poly(Phi - G @ Cc) - az)[1:]
    n0 = num_of(G0)
    if f == 0.0:                                      # whole-sample delay: taps u(k-d-1) ... u(k-d-n), the plain c2d
        return np.concatenate([[0.0], n0]), az, d
    bz = np.zeros(n + 1)
    bz[:n] += num_of(G1)                              # taps u(k-d) ... u(k-d-n+1)
    bz[1:] += n0                                      # taps u(k-d-1) ... u(k-d-n)
    return bz, az, d


def descomp_mpc(bz, az, d: int):
    """DTC-GPC/descompMPC.m:19-43 for one channel: numerator B, denominator A, delay; when the leading numerator
    coefficient is non-zero the delay is reduced by one and a zero is prepended (:35-38)."""
    bz = np.asarray(bz, float).copy()
    if bz[0] != 0.0:
        return np.concatenate([[0.0], bz]), np.asarray(az, float), d - 1
    return bz, np.asarray(az, float), d
