"""Pins the CPU oracle (oracle/mpc_oracle.c) without MATLAB (SURVEY.md §8c "How the oracle earns trust"):
 (i)   every sampled QP optimum satisfies the KKT conditions of the stated QP and agrees with scipy;
 (ii)  the unconstrained limit equals the closed-form GPC gain K=(G'QG+W)\\G'Q (DTC_GPC_WW.m:98-100);
 (iii) the in-loop plant stepping equals an independent lsim of the recorded inputs;
 (iv)  invariants: limits respected, huge lambda => u == 0 and cost == sum(Yref^2).
The reference pins no per-candidate cost/trajectory, so parity versus the Toolbox is UNPINNED."""
import numpy as np
import pytest
from scipy.optimize import minimize

from mpcgpu import shell3x3, woodberry, shell7x5, simulate, synthetic_population
from oracle import oracle as orc


def dense_constraints(p, N, Nu, G, yfree, uprev):
    """Independent (numpy) statement of the constraint set: rows C z >= b."""
    nu, ny = p.nu, p.ny
    nzu = nu * Nu
    has_eps = np.isfinite(p.ymin).any() or np.isfinite(p.ymax).any()
    nz = nzu + int(has_eps)
    rows, rhs = [], []
    for c in range(Nu):
        for j in range(nu):
            e = np.zeros(nz); e[c * nu + j] = 1.0
            pre = np.zeros(nz); pre[[cc * nu + j for cc in range(c + 1)]] = 1.0
            if np.isfinite(p.dumin[j]): rows.append(e); rhs.append(p.dumin[j])
            if np.isfinite(p.dumax[j]): rows.append(-e); rhs.append(-p.dumax[j])
            if np.isfinite(p.umin[j]): rows.append(pre); rhs.append(p.umin[j] - uprev[j])
            if np.isfinite(p.umax[j]): rows.append(-pre); rhs.append(uprev[j] - p.umax[j])
    if has_eps:
        for t in range(N):
            for i in range(ny):
                row = t * ny + i
                if np.isfinite(p.ymax[i]):
                    a = np.zeros(nz); a[:nzu] = -G[row]; a[nzu] = p.ecr_max[i] * p.sy[i]
                    rows.append(a); rhs.append(yfree[row] - p.ymax[i])
                if np.isfinite(p.ymin[i]):
                    a = np.zeros(nz); a[:nzu] = G[row]; a[nzu] = p.ecr_min[i] * p.sy[i]
                    rows.append(a); rhs.append(p.ymin[i] - yfree[row])
        a = np.zeros(nz); a[nzu] = 1.0
        rows.append(a); rhs.append(0.0)
    return np.array(rows), np.array(rhs)


def kkt_check(H, f, C, b, z, tol=1e-8):
    s = C @ z - b
    assert s.min() > -tol, f"primal infeasible by {s.min()}"
    act = np.where(s < 1e-7)[0]
    g = H @ z + f
    if len(act) == 0:
        assert np.abs(g).max() < tol * (1 + np.abs(f).max())
        return 0
    # non-negative least squares certificate: g = C_A' mu, mu >= 0
    from scipy.optimize import nnls
    mu, res = nnls(C[act].T, g)
    assert res < 1e-7 * (1 + np.linalg.norm(g)), f"stationarity residual {res}"
    return len(act)


def random_state(p, rng, hl):
    nw = p.nu + p.nd
    xs = rng.standard_normal((p.ny, nw)) * 0.2
    lo = np.concatenate([p.umin, -np.ones(p.nd)]); hi = np.concatenate([p.umax, np.ones(p.nd)])
    lo = np.where(np.isfinite(lo), lo, -1.0); hi = np.where(np.isfinite(hi), hi, 1.0)
    wh = rng.uniform(lo[:, None], hi[:, None], size=(nw, hl))
    uprev = wh[:p.nu, 0].copy()
    hv = np.concatenate([uprev, rng.uniform(-1, 1, size=p.nd)])
    rk = rng.standard_normal(p.ny) * 0.5
    return xs, wh, hv, rk, uprev


@pytest.mark.parametrize("case", ["shell3x3", "woodberry", "shell7x5"])
def test_qp_optimum_kkt_and_scipy(case):
    p = {"shell3x3": lambda: shell3x3(2), "woodberry": woodberry, "shell7x5": shell7x5}[case]()
    op = orc.OracleProblem(p)
    rng = np.random.default_rng(7)
    hl = int(p.plant.d.max()) + 2
    nact_total = 0
    n_scipy_ok = 0
    trials = 12 if case != "shell7x5" else 6
    for trial in range(trials):
        N = int(rng.integers(8, 40)); Nu = int(rng.integers(2, 7))
        delta = np.exp(rng.uniform(np.log(1e-2), np.log(3), p.ny)); delta[p.band_mask] = 0
        lam = np.exp(rng.uniform(np.log(1e-3), np.log(3), p.nu))
        xs, wh, hv, rk, uprev = random_state(p, rng, hl)
        z, H, f, G, yfree, iters, rc = orc.single_qp(op, N, Nu, delta, lam, xs, wh, hv, rk, uprev)
        assert rc == 0
        C, b = dense_constraints(p, N, Nu, G, yfree, uprev)
        nact_total += kkt_check(H, f, C, b, z)
        # independent solver on the same QP
        res = minimize(lambda x: 0.5 * x @ H @ x + f @ x, np.zeros_like(z), jac=lambda x: H @ x + f,
                       constraints=[{"type": "ineq", "fun": lambda x: C @ x - b, "jac": lambda x: C}],
                       method="SLSQP", options={"ftol": 1e-15, "maxiter": 2000})
        j_or = 0.5 * z @ H @ z + f @ z
        # (feasible start z=0 is not guaranteed for soft rows, SLSQP handles it)
        if (C @ res.x - b).min() > -1e-9:   # SLSQP often exits with status 8 *at* the optimum; feasibility is what matters
            n_scipy_ok += 1
            assert j_or <= res.fun + 1e-9 * (1 + abs(res.fun)), (j_or, res.fun)
            assert abs(j_or - res.fun) <= 1e-8 * (1 + abs(res.fun))
            assert np.abs(res.x - z).max() < 1e-6
    assert nact_total > 0, "test never exercised an active constraint"
    if case != "shell7x5":
        assert n_scipy_ok >= trials // 2


def test_unconstrained_limit_is_gpc_gain():
    p = shell3x3(2)
    for a in ("umin", "dumin"): setattr(p, a, np.full(3, -np.inf))
    for a in ("umax", "dumax"): setattr(p, a, np.full(3, np.inf))
    op = orc.OracleProblem(p)
    rng = np.random.default_rng(3)
    hl = int(p.plant.d.max()) + 2
    N, Nu = 30, 5
    delta = np.array([0.7, 1.3, 0.4]); lam = np.array([0.2, 0.05, 0.6])
    xs, wh, hv, rk, uprev = random_state(shell3x3(2), rng, hl)
    z, H, f, G, yfree, iters, rc = orc.single_qp(op, N, Nu, delta, lam, xs, wh, hv, rk, uprev)
    Q = np.diag(np.tile(delta ** 2, N)); W = np.diag(np.tile(lam ** 2, Nu))
    K = np.linalg.solve(G.T @ Q @ G + W, G.T @ Q)
    z_ref = K @ (np.tile(rk, N) - yfree)
    np.testing.assert_allclose(z, z_ref, rtol=1e-9, atol=1e-12)
    assert iters == 0


@pytest.mark.parametrize("case", ["shell3x3", "woodberry"])
def test_closed_loop_plant_equals_lsim_and_limits(case):
    p = {"shell3x3": lambda: shell3x3(2), "woodberry": woodberry}[case]()
    op = orc.OracleProblem(p)
    N, Nu, delta, lam = synthetic_population(p, 6, seed=11)
    for c in range(6):
        y, u, ys, uo, rc, st = orc.closedloop(op, N[c], Nu[c], delta[c], lam[c])
        assert rc == 0
        np.testing.assert_allclose(y.T, simulate(p.plant, np.hstack([u.T, p.v])), atol=1e-11)
        np.testing.assert_allclose(ys.T, simulate(p.plant, np.hstack([uo.T, p.v])), atol=1e-11)
        assert (u >= p.umin[:, None] - 1e-9).all() and (u <= p.umax[:, None] + 1e-9).all()
        du = np.diff(np.hstack([np.zeros((p.nu, 1)), u]), axis=1)
        assert (du >= p.dumin[:, None] - 1e-9).all() and (du <= p.dumax[:, None] + 1e-9).all()
        assert st[0] == p.nit + 1
        # Info.Uopt rows m..p repeat row m-1, padded to nit (closedloop_toolbox.m:94-98)
        assert np.all(uo[:, Nu[c] - 1:] == uo[:, Nu[c] - 1:Nu[c]])


def test_huge_lambda_freezes_controller():
    p = shell3x3(2)
    op = orc.OracleProblem(p)
    cost, status, _ = orc.eval_batch(op, [30], [4], [[1e-3] * 3], [[1e9] * 3], "gam")
    np.testing.assert_allclose(cost[0], (p.yref ** 2).sum(axis=1), rtol=1e-6)


def test_survey_appendix_c_activity_counts():
    """SURVEY.md appendix C (an independent throwaway restatement made during the survey) counted the
    closed-loop QPs with >=1 active constraint for four Shell3x3 candidates: 20 / 11 / 38 / 99.
    Our counter also includes the open-loop QP (closedloop_toolbox.m:91)."""
    p = shell3x3(1)
    op = orc.OracleProblem(p)
    tuned = ([0.010659948215964849, 0.004019856475662751, 0.0007926546087416782],
             [9.247457388705409e-05, 0.0005523146971406108, 0.0015219790494510478])
    for (N, Nu, d, l, expect) in [(24, 6, tuned[0], tuned[1], 20), (127, 2, [1] * 3, [1] * 3, 11),
                                  (127, 15, [1] * 3, [0.1] * 3, 38), (40, 8, [1] * 3, [1e-3] * 3, 99)]:
        y, u, ys, uo, rc, st = orc.closedloop(op, N, Nu, d, l)
        assert rc == 0 and abs(int(st[2]) - expect) <= 3, (N, Nu, st)
    # tuned case 1 settles on the set-points (appendix C row 1)
    y, *_ = orc.closedloop(op, 24, 6, *tuned)
    np.testing.assert_allclose(y[:, 398] / p.L, [0.1, 0.3, 0.0], atol=1e-4)


def test_vns_and_gam_cost_definitions():
    """orc_eval_batch against a direct numpy transcription of GAM_fun.m:110-115 / VNS2.m:148-195."""
    p = shell3x3(2)
    op = orc.OracleProblem(p)
    N, Nu, delta, lam = synthetic_population(p, 3, seed=5)
    g, st, _ = orc.eval_batch(op, N, Nu, delta, lam, "gam")
    F, st2, _ = orc.eval_batch(op, N, Nu, delta, lam, "vns")
    assert (st == 0).all() and (st2 == 0).all()
    for c in range(3):
        y, *_ = orc.closedloop(op, N[c], Nu[c], delta[c], lam[c])
        np.testing.assert_allclose(g[c], np.diag((y - p.yref) @ (y - p.yref).T), rtol=1e-12)
        Xy = np.zeros((3, p.nit)); Xyma = np.zeros_like(Xy); Xuma = np.zeros_like(Xy)
        for i in range(3):
            r = p.vns_setpoint() * np.eye(3)[i]
            opi = orc.OracleProblem(p, r=r)
            yi, ui, ysi, uoi, rc, _ = orc.closedloop(opi, N[c], Nu[c], delta[c], lam[c])
            Xy[i], Xyma[i], Xuma[i] = yi[i], ysi[i], uoi[i]
        k0 = p.inK - 1
        j21 = np.diag((Xy[:, k0:] - Xyma[:, k0:]) @ (Xy[:, k0:] - Xyma[:, k0:]).T)
        j22 = np.diag((Xy[:, k0:] - p.yref[:, k0:]) @ (Xy[:, k0:] - p.yref[:, k0:]).T)
        with np.errstate(divide="ignore", invalid="ignore"):
            Xnu = np.abs(Xuma[:, :1]) / np.abs(np.diff(Xuma, axis=1))
        Xnu[~np.isfinite(Xnu)] = 0
        Fref = (j21 + j22).sum() + N[c] + np.diag(Xnu @ Xnu.T).sum()
        np.testing.assert_allclose(F[c], Fref, rtol=1e-12)
