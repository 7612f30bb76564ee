"""CPU ORACLE for the single-shooting NMPC of the reference's `Explicit NMPC/` demo.
TEST INFRASTRUCTURE ONLY: may be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs.

What it restates
    NMPC_Controller + objectiveFunction   /root/reference/Explicit NMPC/NMPC_Controller.m:1-141
    ClosedLoopNMPC                        /root/reference/Explicit NMPC/ClosedLoopNMPC.m:1-110
    plant_model                           /root/reference/Explicit NMPC/plant_model.m:1-56 (same constants as vandevusse_model.m)
    problem data                          /root/reference/Explicit NMPC/main.m:20-75

PARITY UNPINNED: `fmincon` (SQP, TolX 1e-6, TolFun 1e-7), `ode23t` / `ode45` are MathWorks code, the reference holds no
output of this demo, and its loop draws `0.01*randn` state noise per sample (ClosedLoopNMPC.m:88-90) -- here the draws are an
INPUT (`noise`, nx x nit; zeros = the noise-free run).  `objective` below is the reference's objective function line by line
(loops, the `uf` hold-over of :87-96, the model-deviation term of :106-123); the minimiser is found with an independent
method -- scipy.optimize.least_squares (trust-region reflective, bounds, 3-point finite-difference Jacobian) to 1e-14, then polished by projected Newton steps on a long-double
evaluation of the same objective (`_polish`: the cost-based stop of any double-precision minimiser leaves X open to ~1e-7 of
the MV range on this flat objective) -- with no code shared with the CUDA kernel's Gauss-Newton / active-set iteration.  The integrator is the product's (RK4, nsub
sub-steps); `integrator="ivp"` swaps in scipy's adaptive solver to measure what that choice costs (tests/, loose tolerance).
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import solve_ivp
from scipy.optimize import least_squares

from .nmpc_oracle import rk4_sample, vandevusse_model


def _step(x, u, Ts, nsub, integrator):
    if integrator == "rk4":
        return rk4_sample(np.asarray(x, float), np.asarray(u, float), Ts, nsub)
    sol = solve_ivp(lambda t, xx: vandevusse_model(xx, u), (0.0, Ts), np.asarray(x, float), method="Radau", rtol=1e-10, atol=1e-12)
    return sol.y[:, -1]


def objective_terms(X, Par):
    """objectiveFunction (NMPC_Controller.m:46-141): returns (error, deltaU) with fobj = error' Qd error + deltaU' Ql deltaU."""
    r, k, N, Nu, u, Ts = Par["r"], Par["k"], Par["N"], Par["Nu"], Par["u"], Par["Ts"]
    my, ny, x_control = Par["my"], Par["ny"], Par["x_control"]
    step = lambda x, uu: _step(x, uu, Ts, Par["nsub"], Par["integrator"])
    Xy = np.zeros(my * N); Xsp = np.zeros(my * N); Xyr = np.zeros(my * N)
    for i in range(my):                                                                  # :68-72
        Xsp[i * N:(i + 1) * N] = r[i, k]
    deltaU = np.asarray(X, float)                                                        # :74-78 (blocks in input order)
    off = np.concatenate([[0], np.cumsum(Nu)])
    xini = np.array(Par["x0plant"], float)                                               # :87
    uf = np.zeros(ny)
    for i in range(1, N + 1):                                                            # :88-108
        if i <= max(Nu):
            for j in range(ny):
                if i <= Nu[j]:
                    uf[j] = u[j, k - 1] + deltaU[off[j] + i - 1]
        xini = step(xini, uf.copy())
        for j in range(my):
            Xy[j * N + i - 1] = xini[x_control[j]]
    xk = step(np.array(Par["x0plant"], float), u[:, k - 1])                              # :116-120
    for i in range(my):                                                                  # :122-126
        Xyr[i * N:(i + 1) * N] = Par["x0plant"][x_control[i]] - xk[x_control[i]]
    error = Xsp - (Xy + Xyr)                                                             # :129-132
    return error, deltaU


def nmpc_controller(Par, polish=True):
    """NMPC_Controller.m:1-44: min fobj over lb - u(k-1) <= X <= ub - u(k-1), from X = 0."""
    Nu, u, k = Par["Nu"], Par["u"], Par["k"]
    aux = np.concatenate([np.full(Nu[i], u[i, k - 1]) for i in range(len(Nu))])           # :15-19
    li = Par["lb"] - aux; ls = Par["ub"] - aux                                           # :21-22
    sq = np.sqrt(np.concatenate([np.full(Par["N"], q) for q in Par["Q"]]))
    sw = np.sqrt(np.concatenate([np.full(Nu[i], Par["W"][i]) for i in range(len(Nu))]))
    scale = np.concatenate([np.full(Nu[i], Par["ub1"][i] - Par["lb1"][i]) for i in range(len(Nu))])

    def residuals(Xs):
        e, dU = objective_terms(Xs * scale, Par)
        return np.concatenate([sq * e, sw * dU])

    x0 = np.clip(np.zeros(sum(Nu)), li, ls)
    sol = least_squares(residuals, x0 / scale, bounds=(li / scale, ls / scale), method="trf", jac="3-point",
                        xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=400)
    X = sol.x * scale
    return _polish(X, Par, li, ls, scale) if polish else X


def objective_ld(X, Par):
    """The same objective in extended precision (numpy long double), vectorised -- a second statement of
    NMPC_Controller.m:46-141 used only to polish the minimiser (tests check it against `objective_terms`)."""
    LD = np.longdouble
    X = np.asarray(X, LD); Nu = Par["Nu"]; N = Par["N"]; k = Par["k"]; xc = Par["x_control"]
    off = np.concatenate([[0], np.cumsum(Nu)])
    up = np.asarray(Par["u"][:, k - 1], LD); x = np.asarray(Par["x0plant"], LD); Ts = LD(Par["Ts"])
    xk = rk4_sample(x, up, Ts, Par["nsub"])
    bias = x[xc] - xk[xc]
    r = np.asarray(Par["r"][:, k], LD); Q = np.asarray(Par["Q"], LD)
    J = LD(0)
    for i in range(N):
        uf = up + np.array([X[off[j] + min(i, Nu[j] - 1)] for j in range(len(Nu))], LD)
        x = rk4_sample(x, uf, Ts, Par["nsub"])
        e = r - (x[xc] + bias)
        J = J + (Q * e * e).sum()
    Wv = np.concatenate([np.full(Nu[j], LD(Par["W"][j])) for j in range(len(Nu))])
    return J + (Wv * X * X).sum()


def _polish(X, Par, li, ls, scale):
    """least_squares stops when the COST stops resolving a step (1e-16 relative), which leaves X open to ~1e-7 of the MV
    range along the flat directions of this objective (W ~ 1e-4) -- and the closed loop amplifies that.  Projected Newton
    steps on the long-double objective (central differences: gradient to ~1e-13, Hessian from gradient differences) close
    that gap: the result is the stationary point to ~1e-10 of the MV range."""
    LD = np.longdouble
    n = len(X)
    X0 = X
    X = np.asarray(X, LD); li = np.asarray(li, LD); ls = np.asarray(ls, LD); sc = np.asarray(scale, LD)

    def grad(Z, h=LD(1e-6)):
        g = np.zeros(n, LD)
        for i in range(n):
            e = np.zeros(n, LD); e[i] = h * sc[i]
            g[i] = (objective_ld(Z + e, Par) - objective_ld(Z - e, Par)) / (2 * e[i])
        return g

    H = None
    for _ in range(4):
        g = grad(X)
        at_lo = (X - li <= 1e-7 * sc) & (g > 0); at_hi = (ls - X <= 1e-7 * sc) & (g < 0)   # TRF stays just inside an active bound
        X = np.where(at_lo, li, np.where(at_hi, ls, X))
        free = ~(at_lo | at_hi)
        if not free.any():
            break
        if H is None:      # one Hessian per call (forward differences of the gradient), reused: chord Newton
            H = np.zeros((n, n), LD)
            for i in range(n):
                e = np.zeros(n, LD); e[i] = LD(1e-4) * sc[i]
                H[:, i] = (grad(X + e) - g) / e[i]
            H = (H + H.T) / 2
        idx = np.where(free)[0]
        d = np.linalg.solve(np.asarray(H[np.ix_(idx, idx)], float), -np.asarray(g[idx], float))
        Xn = X.copy(); Xn[idx] = np.clip(X[idx] + np.asarray(d, LD), li[idx], ls[idx])
        step = np.abs((Xn - X) / sc).max()
        if step > 1e-4:    # a polish moves X by ~1e-7 of the range; anything larger means the active set is not settled: keep X0
            return np.asarray(X0, float)
        X = Xn
        if step < 1e-11:
            break
    return np.asarray(X, float)


def closed_loop_nmpc(x0_model, x_control, u0, r, N, Nu, Q, W, nit, ub1, lb1, inK, Ts, noise=None, nsub=4, integrator="rk4", polish=True):
    """[y, u] = ClosedLoopNMPC(x0_model, x_control, u0, r, N, Nu, Q, W, nit, ub1, lb1, inK, Ts)  (ClosedLoopNMPC.m:1).
    x_control 0-based here.  inK as in MATLAB (1-based first simulated sample)."""
    Nu = [int(a) for a in np.atleast_1d(Nu)]
    my, ny = len(Q), len(W)
    N = int(np.atleast_1d(N)[0])                                                          # Q(i)*eye(N(1)) (:37)
    ub = np.concatenate([np.full(Nu[i], ub1[i]) for i in range(ny)])                      # :54-59
    lb = np.concatenate([np.full(Nu[i], lb1[i]) for i in range(ny)])
    x0plant = np.array(x0_model, float)
    u = np.tile(np.asarray(u0, float)[:, None], (1, nit))                                 # :62
    y = np.tile(x0plant[list(x_control)][:, None], (1, nit))                              # :64
    Par = dict(r=np.asarray(r, float), N=N, Nu=Nu, Ts=Ts, my=my, ny=ny, x_control=list(x_control), lb=lb, ub=ub, Q=Q, W=W,
               ub1=np.asarray(ub1, float), lb1=np.asarray(lb1, float), nsub=nsub, integrator=integrator, u=u)
    for k in range(inK - 1, nit):                                                         # :79 (k = inK:nit, 1-based)
        x0plant = _step(x0plant, u[:, k - 1], Ts, nsub, integrator)                       # :82-86
        if noise is not None:
            x0plant = x0plant + noise[:, k]                                               # :89
        Par["x0plant"] = x0plant; Par["k"] = k
        y[:, k] = x0plant[list(x_control)]                                                # :93
        duo = nmpc_controller(Par, polish=polish and integrator == "rk4")                      # :96
        off = np.concatenate([[0], np.cumsum(Nu)])
        u[:, k] = u[:, k - 1] + np.array([duo[off[j]] for j in range(ny)])                # :99-105
    return y, u


def sweep_cost(y, r, inK):
    """Sum of squared tracking errors over the simulated window k = inK..nit (the library's sweep objective)."""
    return ((y[:, inK - 1:] - np.asarray(r)[:, inK - 1:]) ** 2).sum(axis=1)
