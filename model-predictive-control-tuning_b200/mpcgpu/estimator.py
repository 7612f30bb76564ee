"""Plant-model mismatch validation run: the controller's state estimator and the "real" plants.

The reference validates a tuning by simulating the controller against a plant that differs from its prediction model
(`options = mpcsimopt(mpc); options.Model = plant; sim(mpc, nit, r, [], options)`: Shell3x3.m:271-286,
WoodBerry.m:263-278, Shell7x5.m:293-306).  The controller then runs the Toolbox's DEFAULT state estimator.  Restated from
the Toolbox documentation ("Controller State Estimation"; parity unpinned like the rest of the Toolbox arithmetic):

  E1  controller state x_c = [model channel states (ny*nw); MV delay-line states w_j(k-1-q), q = 0..hl-1; output-disturbance
      states (ny)]
  E2  output-disturbance model: one discrete integrator per measured output driven by unit-variance white noise;
      measurement noise: unit-variance white noise
  E3  unit-variance white noise added to every MV; measured disturbances are known
  E4  steady-state Kalman filter of that model, innovation form: x_c(k|k) = x_c(k|k-1) + M (y(k) - C x_c(k|k-1))

`default_estimator_gain` builds (A, C, G) from the problem's own discrete channels and returns M; the C ABI takes M as an
input (mpcgpu_set_mismatch), so a MATLAB caller may pass the gain of its own `getEstimator(mpcobj)` mapped to E1's order.
"""
from __future__ import annotations

import numpy as np

from .plant import Channels, c2d_fopdt
from .problems import LinearProblem, SHELL3X3_L, SHELL3X3_R, WOODBERRY_L, WOODBERRY_R, SHELL7X5_L, SHELL7X5_R


def history_length(prob: LinearProblem, plant: Channels | None = None) -> int:
    """hl = longest sample delay of model and plant + 2 (the layout of the library's input histories)."""
    d = int(prob.plant.d.max())
    if plant is not None:
        d = max(d, int(plant.d.max()))
    return d + 2


def estimator_model(prob: LinearProblem, hl: int):
    """(A, Bu, C, G) of E1-E3 in the library's state order."""
    ch = prob.plant
    ny, nu, nd = prob.ny, prob.nu, prob.nd
    nw = nu + nd
    nch = ny * nw
    n = nch + nu * hl + ny
    A = np.zeros((n, n)); Bu = np.zeros((n, nu)); Cm = np.zeros((ny, n))
    hidx = lambda j, q: nch + j * hl + q          # w_j(k-1-q)
    for i in range(ny):
        for j in range(nw):
            s = i * nw + j
            A[s, s] = ch.a[i, j]
            Cm[i, s] = 1.0
            if j >= nu:
                continue                           # measured disturbance: a known input, no state of the estimator
            d = int(ch.d[i, j])
            # x+ = a x + b0 w(k+1-d) + b1 w(k-d);  w(k) = u(k),  w(k-q') = history state q'-1
            for coef, lag in ((ch.b0[i, j], d - 1), (ch.b1[i, j], d)):
                if coef == 0.0:
                    continue
                if lag == 0:
                    Bu[s, j] += coef
                else:
                    A[s, hidx(j, lag - 1)] += coef
    for j in range(nu):
        Bu[hidx(j, 0), j] = 1.0
        for q in range(1, hl):
            A[hidx(j, q), hidx(j, q - 1)] = 1.0
    for i in range(ny):
        A[nch + nu * hl + i, nch + nu * hl + i] = 1.0
        Cm[i, nch + nu * hl + i] = 1.0
    G = np.zeros((n, nu + ny))
    G[:, :nu] = Bu                                 # E3
    for i in range(ny):
        G[nch + nu * hl + i, nu + i] = 1.0         # E2
    return A, Bu, Cm, G


def default_estimator_gain(prob: LinearProblem, hl: int | None = None) -> np.ndarray:
    """M of E4, (ny*nw + nu*hl + ny) x ny, row-major."""
    from scipy.linalg import solve_discrete_are
    hl = history_length(prob) if hl is None else int(hl)
    A, Bu, Cm, G = estimator_model(prob, hl)
    Q = G @ G.T
    Rn = np.eye(prob.ny)
    P = solve_discrete_are(A.T, Cm.T, Q, Rn)
    M = P @ Cm.T @ np.linalg.inv(Cm @ P @ Cm.T + Rn)
    return np.ascontiguousarray(M)


def _scaled(K, tau, theta, Ts, L, R) -> Channels:
    return c2d_fopdt(np.asarray(K, float), np.asarray(tau, float), np.asarray(theta, float), Ts).scaled(np.asarray(L), np.asarray(R))


def shell3x3_real_plant(e=(0.2, 0.2, 0.3), L=SHELL3X3_L, R=SHELL3X3_R) -> Channels:
    """Psr of Shell3x3.m:36-48 (gain errors e1, e2, e3 per input), scaled like the model (:271 real_plant = L*Psr*R)."""
    e1, e2, e3 = e
    K = [[4.05 + 2.11 * e1, 1.77 + 0.39 * e2, 5.88 + 0.59 * e3], [5.39 + 3.29 * e1, 5.72 + 0.57 * e2, 6.9 + 0.89 * e3],
         [4.38 + 3.11 * e1, 4.42 + 0.73 * e2, 7.2 + 1.33 * e3]]
    return _scaled(K, [[50, 60, 50], [50, 60, 40], [33, 44, 19]], [[27, 28, 27], [18, 14, 15], [20, 22, 0]], 4.0, L, R)


def woodberry_real_plant(deltak=0.2, deltaL=1.0, L=WOODBERRY_L, R=WOODBERRY_R) -> Channels:
    """Gsr of WoodBerry.m:33-47 (gain and dead-time errors on the MV channels; the disturbance channels are the model's)."""
    K = [[12.8 * (1 + deltak), -18.9 * (1 + deltak), 3.8], [6.6 * (1 + deltak), -19.4 * (1 + deltak), 4.9]]
    return _scaled(K, [[16.7, 21, 14.9], [10.9, 14.4, 13.2]], [[1 + deltaL, 2 + deltaL, 8.1], [2 + deltaL, 1 + deltaL, 3.4]], 1.0, L, R)


def shell7x5_real_plant(e=(0.2, 0.2, 0.3, 0.5, 0.5), L=SHELL7X5_L, R=SHELL7X5_R) -> Channels:
    """Gr / Gdr of Shell7x5.m:38-66 (gain errors e1..e5 per input), scaled like the model."""
    e1, e2, e3, e4, e5 = e
    Kg = np.array([[4.05 + 2.11 * e1, 1.77 + 0.39 * e2, 5.88 + 0.59 * e3], [5.39 + 3.29 * e1, 5.72 + 0.57 * e2, 6.9 + 0.89 * e3],
                   [3.66 + 2.29 * e1, 1.65 + 0.35 * e2, 5.53 + 0.67 * e3], [5.92 + 2.34 * e1, 2.54 + 0.24 * e2, 8.10 + 0.32 * e3],
                   [4.13 + 1.71 * e1, 2.38 + 0.93 * e2, 6.23 + 0.30 * e3], [4.06 + 2.39 * e1, 4.18 + 0.35 * e2, 6.53 + 0.72 * e3],
                   [4.38 + 3.11 * e1, 4.42 + 0.73 * e2, 7.2 + 1.33 * e3]])
    Kd = np.array([[1.20 + 0.12 * e4, 1.44 + 0.16 * e5], [1.52 + 0.13 * e4, 1.83 + 0.13 * e5], [1.16 + 0.08 * e4, 1.27 + 0.08 * e5],
                   [1.73 + 0.02 * e4, 1.79 + 0.04 * e5], [1.31 + 0.03 * e4, 1.26 + 0.02 * e5], [1.19 + 0.08 * e4, 1.17 + 0.01 * e5],
                   [1.14 + 0.18 * e4, 1.26 + 0.10 * e5]])
    tg = [[50, 60, 50], [50, 60, 40], [9, 30, 40], [12, 27, 20], [8, 19, 10], [13, 33, 9], [33, 44, 19]]
    dg = [[27, 28, 27], [18, 14, 15], [2, 20, 2], [11, 12, 2], [5, 7, 2], [8, 4, 1], [20, 22, 0]]
    td = [[45, 40], [25, 20], [11, 6], [5, 19], [2, 22], [19, 24], [24, 32]]
    dd = [[27, 27], [15, 15], [0, 0], [0, 0], [0, 0], [0, 0], [0, 0]]
    return _scaled(np.hstack([Kg, Kd]), np.hstack([tg, td]), np.hstack([dg, dd]), 4.0, L, R)
