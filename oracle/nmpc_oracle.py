"""CPU ORACLE for the nonlinear path (BASELINE.json configs[4], Van de Vusse CSTR).  TEST INFRASTRUCTURE ONLY:
may be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs.

What it restates
    closedloop_toolbox_nmpc        /root/reference/MPC-Tuning/MPC_Tuning/closedloop_toolbox_nmpc.m:36-97
    vandevusse_model (RHS)         /root/reference/MPC-Tuning/vandevusse_model.m:39-77  (= nmpc_vandevusse_state.m:44-82)
    GAM / VNS objectives           GAM_fun.m:110-115, VNS2.m:147-195 (nonlinear branch: `Xsp.*sel`, inK = 10)
    problem data                   VanDeVusse_NMPC.m:35-204

PARITY UNPINNED.  `nlmpcmove` (fmincon-SQP on a trapezoidal-collocation NLP) and the plant integrator `ode15s`
(RelTol 1e-3) are MathWorks code that is not in the reference tree; their iterates are path dependent, so not
even MATLAB reproduces itself to 1e-6 across releases.  What is restated is the *optimisation problem* of one
nlmpcmove call, from the Toolbox documentation ("Optimization Problem", nlmpc "Model.IsContinuousTime"):
    N1  decision: the MV levels u(k+c), c = 0..m-1, held afterwards (ControlHorizon = m, closedloop_toolbox_nmpc.m:51)
    N2  cost  sum_{i=1..p} sum_j (delta_j/sy_j (r_j - y_j(k+i)))^2 + sum_{c<m} sum_j (lambda_j/su_j du_j(k+c))^2,
        weights SQUARED, ScaleFactors divide, reference held over the horizon (r(:,i)', :69)
    N3  MV bounds hard; OV/state bounds are soft in the Toolbox (ECR 1e5) and inactive in the tuning scenario
        (cB <= 1.2, 40 <= T <= 150 around set-points 1.0 / 130): they are checked, not enforced (status 5 if crossed)
    N4  prediction = plant = RK4 with NSUB sub-steps per sample (the reference: implicit trapezoid inside
        nlmpc, ode15s for the plant; SURVEY.md section 7 explains why RK4 needs >= 4 sub-steps here)
and THIS file solves that problem with an independent method: scipy.optimize.least_squares (trust-region
reflective, bounds, finite-difference Jacobian) to 1e-13 -- no code shared with the CUDA kernel's Gauss-Newton /
active-set iteration.  Agreement is therefore at solver tolerance, stated in tests/ (1e-5 on trajectories).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import least_squares

# vandevusse_model.m:42-57
K10, K20, K30 = 1.287e12, 1.287e12, 9.043e9
E1, E2, E3 = -9758.3, -9758.3, -8560.0
DAB, DBC, DAD = -4.20, 11.00, 41.85
RHO, CP, KW, AR, VOL, T0, CA0 = 0.9342, 3.01, 4032.0, 0.215, 10.0, 130.00, 5.10


def vandevusse_model(x, u):
    """vandevusse_model.m:59-77."""
    fov, Tk = u
    ca, cb, T = x
    k1 = K10 * np.exp(E1 / (T + 273.15))
    k2 = K20 * np.exp(E2 / (T + 273.15))
    k3 = K30 * np.exp(E3 / (T + 273.15))
    return np.array([
        fov * (CA0 - ca) - k1 * ca - k3 * ca * ca,
        -fov * cb + k1 * ca - k2 * cb,
        (1 / (RHO * CP)) * (k1 * ca * DAB + k2 * cb * DBC + k3 * ca ** 2 * DAD) + fov * (T0 - T) + (KW * AR / (RHO * CP * VOL)) * (Tk - T)])


def rk4_sample(x, u, Ts, nsub):
    h = Ts / nsub
    for _ in range(nsub):
        k1 = vandevusse_model(x, u)
        k2 = vandevusse_model(x + 0.5 * h * k1, u)
        k3 = vandevusse_model(x + 0.5 * h * k2, u)
        k4 = vandevusse_model(x + h * k3, u)
        x = x + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    return x


def nlmpcmove(prob, x, uprev, r, p, m, delta, lam, v0=None):
    """One controller call (N1-N4): returns the optimal MV plan v (m x nu)."""
    nu = 2
    wy = np.asarray(delta, float) / prob.sy
    wdu = np.asarray(lam, float) / prob.su
    lo = np.tile(prob.umin, m); hi = np.tile(prob.umax, m)
    scale = np.tile(prob.su, m)

    def residuals(vs):
        v = (vs * scale).reshape(m, nu)
        xx = np.array(x, float)
        res = []
        for i in range(p):
            xx = rk4_sample(xx, v[min(i, m - 1)], prob.Ts, prob.nsub)
            res.extend(wy * (r - xx[1:3]))
        prev = np.asarray(uprev, float)
        for c in range(m):
            res.extend(wdu * (v[c] - prev))
            prev = v[c]
        return np.array(res)

    start = np.tile(np.clip(uprev, prob.umin, prob.umax), m) if v0 is None else np.clip(np.asarray(v0, float).ravel(), lo, hi)
    sol = least_squares(residuals, start / scale, bounds=(lo / scale, hi / scale), method="trf", jac="3-point",
                        xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=400)
    return (sol.x * scale).reshape(m, nu)


def closedloop_toolbox_nmpc(prob, r, N, Nu, delta, lam, nit=None):
    """closedloop_toolbox_nmpc.m:36-97: returns y, u, yopt, uopt (signals x time) and a status
    (0 ok, 5 an OV/state bound of N3 was crossed)."""
    nit = prob.nit if nit is None else nit
    p, m = int(np.max(N)), int(np.max(Nu))
    nx, ny, nu = 3, 2, 2
    X = np.zeros((nx, nit)); Y = np.zeros((ny, nit)); U = np.zeros((nu, nit))
    X[:, 0] = prob.x0; Y[:, 0] = prob.x0[1:3]; U[:, 0] = prob.u0                       # :61-63
    status = 0
    plan = None
    for i in range(1, nit):                                                            # :67-74
        plan = nlmpcmove(prob, X[:, i - 1], U[:, i - 1], r[:, i], p, m, delta, lam, plan)
        U[:, i] = plan[0]
        X[:, i] = rk4_sample(X[:, i - 1], U[:, i], prob.Ts, prob.nsub)
        Y[:, i] = X[1:, i]
        if (X[:, i] < prob.xmin - 1e-9).any() or (X[:, i] > prob.xmax + 1e-9).any():
            status = 5
    mv = nlmpcmove(prob, prob.x0, prob.u0, r[:, -1], p, m, delta, lam)                  # :79
    mvopt = np.vstack([mv[min(c, m - 1)] for c in range(p + 1)])                        # Info.MVopt: (p+1) x nu
    uopt = np.vstack([mvopt, np.tile(mvopt[-1], (max(nit - (p + 1), 0), 1))])[:nit].T   # :80
    Xop = np.zeros((nx, nit)); yopt = np.zeros((ny, nit))
    Xop[:, 0] = prob.x0; yopt[:, 0] = prob.x0[1:3]                                     # :87-88
    for i in range(1, nit):                                                            # :89-94
        Xop[:, i] = rk4_sample(Xop[:, i - 1], uopt[:, i], prob.Ts, prob.nsub)
        yopt[:, i] = Xop[1:, i]
    return Y, U, yopt, uopt, status


def gam_cost(prob, N, Nu, delta, lam):
    """GAM_fun.m:87,110-115."""
    y, u, yo, uo, st = closedloop_toolbox_nmpc(prob, prob.r, N, Nu, delta, lam)
    return ((y - prob.yref) ** 2).sum(axis=1), st


def vns_cost(prob, N, Nu, delta, lam, inK=10):
    """VNS2.m:147-195, nonlinear branch: one run per output with the other set-points ZEROED (`Xsp.*sel`)."""
    ny = 2
    Xy = np.zeros((ny, prob.nit)); Xyma = np.zeros_like(Xy); Xuma = np.zeros_like(Xy)
    st = 0
    for i in range(ny):
        sel = np.zeros((ny, 1)); sel[i] = 1.0
        y, u, yo, uo, s1 = closedloop_toolbox_nmpc(prob, prob.r * sel, N, Nu, delta, lam)
        Xy[i] = y[i]; Xyma[i] = yo[i]; Xuma[i] = uo[i]
        st = max(st, s1)
    k0 = inK - 1
    j21 = ((Xy[:, k0:] - Xyma[:, k0:]) ** 2).sum(axis=1)
    j22 = ((Xy[:, k0:] - prob.yref[:, k0:]) ** 2).sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        Xnu = np.abs(Xuma[:, :1]) / np.abs(np.diff(Xuma, axis=1))
    Xnu[~np.isfinite(Xnu)] = 0.0
    return float((j21 + j22).sum() + int(np.max(N)) + (Xnu ** 2).sum()), st
