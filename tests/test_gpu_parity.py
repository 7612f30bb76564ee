"""GPU parity tests proper: libmpcgpu.so (CUDA, sm_100a) through the C ABI versus the CPU oracle.
Tolerances (DESIGN.md "Tolerance"): 1e-6 relative on cost, 1e-5 absolute on trajectories
(BASELINE.json north_star), relaxed to 20*cond(H)*eps only where fp64 cannot resolve 1e-6."""
import json
import os

import numpy as np
import pytest

import mpcgpu
from mpcgpu import shell3x3, woodberry, synthetic_population
from oracle import oracle as orc
from parity_util import check_cost, oracle_sensitivity, TOL_COST, TOL_TRAJ, vns_well_posed

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ev3():
    e = mpcgpu.Evaluator(shell3x3(2), device=0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def evwb():
    e = mpcgpu.Evaluator(woodberry(), device=0)
    yield e
    e.close()


def test_extension_loaded_and_fp64_peak():
    tf = mpcgpu.measure_fp64_peak(0)
    assert 5.0 < tf < 80.0, tf     # B200 fp64 FMA: ~37 TFLOP/s nominal


@pytest.mark.parametrize("case,n", [("shell3x3", 512), ("woodberry", 256)])
def test_gam_cost_parity(case, n, ev3, evwb):
    ev = ev3 if case == "shell3x3" else evwb
    p = ev.prob
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, n, seed=2)
    g0, st0, stats = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    c0 = ev.counters()
    out = ev.eval_batch(N, Nu, dl, lm, mode="gam")
    c1 = ev.counters()
    assert (out["status"] == 0).all()
    rel, strict = check_cost(out["cost"], g0, oracle_sensitivity(op, N, Nu, dl, lm, "gam", g0), case, min_strict=0.75)
    assert np.median(rel) < 1e-10
    assert c1["as_iterations"] > c0["as_iterations"] and c1["qp_constrained"] > c0["qp_constrained"]
    assert c1["kernel_launches"] > c0["kernel_launches"]


@pytest.mark.parametrize("case", ["shell3x3", "woodberry"])
def test_trajectory_parity(case, ev3, evwb):
    ev = ev3 if case == "shell3x3" else evwb
    p = ev.prob
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, 48, seed=3, wlo=1e-3, whi=3.0)
    out = ev.eval_batch(N, Nu, dl, lm, mode="raw")
    assert (out["status"] == 0).all()
    g0, _, _ = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    sens = oracle_sensitivity(op, N, Nu, dl, lm, "gam", g0)
    assert (sens < 1e-8).sum() >= 32          # the trajectories of these must agree to 1e-5
    for c in np.where(sens < 1e-8)[0]:
        y, u, ys, uo, rc, _ = orc.closedloop(op, N[c], Nu[c], dl[c], lm[c])
        for k, b in zip(("y", "u", "ys", "uopt"), (y, u, ys, uo)):
            assert np.abs(out[k][c] - b).max() < TOL_TRAJ, (c, k)


@pytest.mark.parametrize("case", ["shell3x3", "woodberry"])
def test_vns_cost_parity(case, ev3, evwb):
    ev = ev3 if case == "shell3x3" else evwb
    p = ev.prob
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, 64, seed=4, wlo=1e-3, whi=3.0)
    F0, st0, _ = orc.eval_batch(op, N, Nu, dl, lm, "vns")
    out = ev.eval_batch(N, Nu, dl, lm, mode="vns")
    assert (out["status"] == 0).all()
    ok = vns_well_posed(p, lambda r: orc.OracleProblem(p, r=r), N, Nu, dl, lm)
    sens = oracle_sensitivity(op, N, Nu, dl, lm, "vns", F0)
    assert ok.sum() >= 32
    check_cost(out["cost"][ok], F0[ok], sens[ok], case + " vns", min_strict=0.5)


def test_golden_fixture(ev3):
    """Committed oracle outputs (tests/golden/make_oracle_golden.py) for fixed candidates incl. the
    reference's two tuned Shell3x3 results."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden_shell3x3.npz"))
    out = ev3.eval_batch(gold["N"], gold["Nu"], gold["delta"], gold["lam"], mode="gam", traj=True)
    assert (out["status"] == 0).all()
    rel = np.abs(out["cost"] - gold["gam"]) / np.abs(gold["gam"])
    assert rel.max() < 1e-6
    assert np.abs(out["y"][:, :, ::10] - gold["y_sub"]).max() < TOL_TRAJ
    assert np.abs(out["u"][:, :, ::10] - gold["u_sub"]).max() < TOL_TRAJ
    vns = ev3.eval_batch(gold["N"], gold["Nu"], gold["delta"], gold["lam"], mode="vns")
    ok = gold["vns_ok"].astype(bool)
    relv = np.abs(vns["cost"] - gold["vns"]) / np.abs(gold["vns"])
    assert relv[ok].max() < 1e-6


def test_closedloop_toolbox_signature(ev3):
    """Drop-in call as Shell3x3.m:231: r_ma.*sel in signals x time orientation, nit_open = N + 30."""
    p = ev3.prob
    N, Nu = 24, [6, 2, 2]
    delta = [0.010659948215964849, 0.004019856475662751, 0.0007926546087416782]
    lam = [9.247457388705409e-05, 0.0005523146971406108, 0.0015219790494510478]
    nit_open = N + 30
    r_ma = (np.ones((3, nit_open)) * p.L[:, None]) * np.array([[1.0], [0.0], [0.0]])
    y, u, t, ys, uopt = mpcgpu.closedloop_toolbox(ev3, r_ma, np.zeros((nit_open, 0)), N, Nu, delta, lam, nit_open)
    assert y.shape == (3, nit_open) and u.shape == (3, nit_open) and t.shape == (1, nit_open)
    op = orc.OracleProblem(p, r=r_ma.T, v=np.zeros((nit_open, 0)), nit=nit_open)
    y0, u0, ys0, uo0, rc, _ = orc.closedloop(op, N, max(Nu), delta, lam)
    for a, b in ((y, y0), (u, u0), (ys, ys0), (uopt, uo0)):
        assert np.abs(a - b).max() < TOL_TRAJ
    # restore the problem's own signals for the other tests
    ev3.set_signals(p.r, None, p.yref, p.nit)


def test_full_size_properties(ev3):
    """BASELINE.json's full population (4096) through size-independent properties."""
    p = ev3.prob
    N, Nu, dl, lm = synthetic_population(p, 4096, seed=0)
    a = ev3.eval_batch(N, Nu, dl, lm, mode="gam")
    assert (a["status"] == 0).all() and np.isfinite(a["cost"]).all() and (a["cost"] >= 0).all()
    b = ev3.eval_batch(N, Nu, dl, lm, mode="gam")
    assert np.array_equal(a["cost"], b["cost"]), "not deterministic run-to-run"
    perm = np.random.default_rng(0).permutation(4096)
    c = ev3.eval_batch(N[perm], Nu[perm], dl[perm], lm[perm], mode="gam")
    assert np.array_equal(a["cost"][perm], c["cost"]), "result depends on the position in the population"
    # spot-check 64 of them against the oracle
    idx = np.random.default_rng(1).choice(4096, 64, replace=False)
    op = orc.OracleProblem(p)
    g0, _, _ = orc.eval_batch(op, N[idx], Nu[idx], dl[idx], lm[idx], "gam")
    check_cost(a["cost"][idx], g0, oracle_sensitivity(op, N[idx], Nu[idx], dl[idx], lm[idx], "gam", g0), "full-size spot check")
    # limits hold on a trajectory subset
    t = ev3.eval_batch(N[:256], Nu[:256], dl[:256], lm[:256], mode="raw")
    u = t["u"]
    assert (u >= p.umin[None, :, None] - 1e-9).all() and (u <= p.umax[None, :, None] + 1e-9).all()
    du = np.diff(np.concatenate([np.zeros((256, 3, 1)), u], axis=2), axis=2)
    assert (np.abs(du) <= p.dumax[None, :, None] + 1e-9).all()


def test_edge_cases(ev3):
    p = ev3.prob
    # empty population
    out = ev3.eval_batch(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 3)), np.zeros((0, 3)), mode="gam")
    assert out["cost"].shape == (0, 3)
    # illegal horizons are flagged, legal neighbours unaffected
    N = np.array([5, 300, 10, 20, 127], dtype=np.int32); Nu = np.array([5, 3, 0, 4, 15], dtype=np.int32)
    out = ev3.eval_batch(N, Nu, np.ones((5, 3)), np.ones((5, 3)), mode="gam")
    assert list(out["status"]) == [4, 4, 4, 0, 0]
    assert np.isnan(out["cost"][:3]).all() and np.isfinite(out["cost"][3:]).all()
    # huge lambda freezes the controller: cost == sum(Yref^2)
    out = ev3.eval_batch([30], [4], [[1e-3] * 3], [[1e9] * 3], mode="gam")
    np.testing.assert_allclose(out["cost"][0], (p.yref ** 2).sum(axis=1), rtol=1e-6)
    # maximum sizes
    out = ev3.eval_batch([127], [15], [[1.0] * 3], [[0.1] * 3], mode="vns")
    assert out["status"][0] == 0 and np.isfinite(out["cost"][0])
    # gam_fun / vns_cost wrappers
    ev3.N, ev3.Nu = 24, 6
    g, h = mpcgpu.gam_fun([1, 1, 1, -0.1, 0.1, 0.1], ev3)
    assert g.shape == (3,) and h.size == 0
    F = mpcgpu.vns_cost(ev3, [24, 6, 24], [6, 2, 30], [1, 1, 1], [.1, .1, .1])
    assert np.isfinite(F[0]) and np.isinf(F[1]) and np.isinf(F[2])


# ---------------------------------------------------------------------------------------------
# DTC-GPC batched sweep (BASELINE.json configs[3])
# ---------------------------------------------------------------------------------------------
def _dtc_oracle(prob, p, m, dl, lm, alfa, raio):
    from oracle import dtc_gpc_oracle as dorc
    ys, us = [], []
    for c in range(len(p)):
        fr = dorc.mimofilter_Fr(prob.pnz, alfa[c], raio[c])
        y, u = dorc.dtc_gpc_closed_loop(prob, p[c], m[c], dl[c], lm[c], fr)
        ys.append(y); us.append(u)
    return np.array(ys), np.array(us)


@pytest.mark.parametrize("deltak,deltaL,min_tame", [(0.0, 0.0, 60), (0.1, 0.0, 48), (0.1, 0.3, 16)])
def test_dtc_gpc_parity(deltak, deltaL, min_tame):
    """GPU sweep vs the restated MATLAB on seeded candidates: ISE 1e-6 relative, y/u 1e-5 absolute.  Candidates
    whose closed loop is unstable (the sweep's ranges include such tunings) are compared while they are
    bounded: an exponentially growing loop amplifies rounding differences without limit."""
    from mpcgpu.dtcgpc import woodberry_dtc, synthetic_dtc_population, DtcEvaluator
    prob = woodberry_dtc(deltak=deltak, deltaL=deltaL)      # plant/model mismatch knobs of DTC_GPC_WW.m:18-19
    ev = DtcEvaluator(prob, device=0)
    p, m, dl, lm, alfa, raio = synthetic_dtc_population(prob, 64, seed=11)
    out = ev.eval_batch(p, m, dl, lm, alfa=alfa, raio=raio, traj=True)
    assert (out["status"] == 0).all()
    y0, u0 = _dtc_oracle(prob, p, m, dl, lm, alfa, raio)
    ise0 = ((y0 - prob.r[None]) ** 2).sum(axis=2)
    tame = np.abs(y0).max(axis=(1, 2)) < 1e3
    assert tame.sum() >= min_tame
    assert np.abs(out["y"][tame] - y0[tame]).max() < TOL_TRAJ
    assert np.abs(out["u"][tame] - u0[tame]).max() < TOL_TRAJ
    rel = np.abs(out["ise"][tame] - ise0[tame]) / np.abs(ise0[tame])
    assert rel.max() < 1e-6, rel.max()
    ev.close()


def test_dtc_gpc_golden_and_edges():
    from mpcgpu.dtcgpc import woodberry_dtc, DtcEvaluator, dtc_gpc_ww
    prob = woodberry_dtc()
    ev = DtcEvaluator(prob, device=0)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden_dtc.npz"))
    out = ev.eval_batch(gold["p"], gold["m"], gold["delta"], gold["lam"], alfa=gold["alfa"], raio=gold["raio"], traj=True)
    assert (out["status"] == 0).all()
    assert np.abs(out["y"][:, :, ::4] - gold["y_sub"]).max() < TOL_TRAJ
    assert np.abs(out["u"][:, :, ::4] - gold["u_sub"]).max() < TOL_TRAJ
    assert (np.abs(out["ise"] - gold["ise"]) / gold["ise"]).max() < 1e-6
    # the reference script's own run through the reference-shaped call
    y, u = dtc_gpc_ww(prob, ev=ev)
    assert np.abs(y[:, ::4] - gold["y_sub"][0]).max() < TOL_TRAJ
    # illegal horizons are flagged, neighbours unaffected; cost-only call equals the trajectory call
    p = np.array([[3, 3], [0, 3], [3, 3], [31, 3]], dtype=np.int32); m = np.array([[3, 3], [3, 3], [11, 3], [3, 3]], dtype=np.int32)
    o2 = ev.eval_batch(p, m, np.ones((4, 2)), np.ones((4, 2)), alfa=0.7, raio=0.8)
    assert list(o2["status"]) == [0, 4, 4, 4] and np.isnan(o2["ise"][1:]).all()
    np.testing.assert_array_equal(o2["ise"][0], out["ise"][0])
    # large population: deterministic and position-independent
    from mpcgpu.dtcgpc import synthetic_dtc_population
    P = synthetic_dtc_population(prob, 4096, seed=0)
    a = ev.eval_batch(*P[:4], alfa=P[4], raio=P[5])
    perm = np.random.default_rng(0).permutation(4096)
    b = ev.eval_batch(*(x[perm] for x in P[:4]), alfa=P[4][perm], raio=P[5][perm])
    assert np.array_equal(a["ise"][perm], b["ise"], equal_nan=True)
    ev.close()


# ---------------------------------------------------------------------------------------------
# Shell7x5: soft output constraints (band control), measured disturbances (BASELINE.json configs[2])
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ev75():
    e = mpcgpu.Evaluator(mpcgpu.shell7x5(), device=0)
    yield e
    e.close()


def test_shell7x5_soft_constraint_parity(ev75):
    p = ev75.prob
    op = orc.OracleProblem(p)
    # lambda in [1e-2, 10] brackets the reference's tuned Shell7x5 result (0.056, 0.0167, 1.61); the survey's full range
    # (down to 1e-4, cond(H) to 1e12) is test_shell7x5_full_weight_range.
    N, Nu, dl, lm = synthetic_population(p, 48, seed=5, wlo=1e-2)
    g0, st0, stats = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    out = ev75.eval_batch(N, Nu, dl, lm, mode="gam", traj=True)
    ok = (st0 == 0) & (out["status"] == 0)
    assert ok.sum() >= 40, (st0, out["status"])
    from oracle import parity
    # what fp64 resolves on a candidate: the oracle's own spread under 1e-13 perturbations, its three pivot rules and a no-FMA build
    sens = parity.sensitivity(op, N[ok], Nu[ok], dl[ok], lm[ok], "gam", g0[ok])
    rel, _ = check_cost(out["cost"][ok], g0[ok], sens, "shell7x5", min_strict=0.5)   # (the oracle's own pivot rules disagree on the rest)
    assert (rel <= TOL_COST).mean() >= 0.9, rel                                      # measured 46 of 48
    ntraj = 0
    for c in np.where(ok)[0]:
        if sens[list(np.where(ok)[0]).index(c)] >= 1e-8 or ntraj >= 16:
            continue
        y, u, ys, uo, rc, _ = orc.closedloop(op, N[c], Nu[c], dl[c], lm[c])
        for k, b in zip(("y", "u", "ys", "uopt"), (y, u, ys, uo)):
            assert np.abs(out[k][c] - b).max() < TOL_TRAJ, (c, k, np.abs(out[k][c] - b).max())
        ntraj += 1
    assert ntraj >= 12
    # the bands hold up to the slack the optimiser buys: MV limits are hard
    u = out["u"][ok]
    assert (u >= p.umin[None, :, None] - 1e-9).all() and (u <= p.umax[None, :, None] + 1e-9).all()
    c = ev75.counters()
    assert c["qp_constrained"] > 0 and c["as_iterations"] > 0


def test_shell7x5_full_weight_range(ev75):
    """SURVEY.md 8d population (lambda log-uniform on [1e-4, 10], delta = 0): every candidate is solved on both sides.
    Round 1 lost a third of them ("infeasible" / iteration cap) to a dependence test at 1e-15 |d|^2 -- with zero tracking
    weights the slack direction, all that separates two band rows, is ~1e-15 of |d|^2 at lambda = 1e-4 -- and to an
    iteration cap of 20 (n + 10); see DEP_TOL in oracle/mpc_oracle.c.  Where fp64 resolves the candidate (the oracle's own
    cost is stable under 1e-13 perturbations of the weights) the costs agree to 1e-6."""
    from oracle import parity
    p = ev75.prob
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, 256, seed=5)
    g0, st0, _ = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    out = ev75.eval_batch(N, Nu, dl, lm, mode="gam")
    assert (st0 == 0).all() and (out["status"] == 0).all(), (np.bincount(st0), np.bincount(out["status"]))
    assert np.isfinite(out["cost"]).all()
    summ = parity.summary(out["cost"], out["status"], g0, st0, parity.sensitivity(op, N, Nu, dl, lm, "gam", g0))
    print("shell7x5 full range:", summ)
    # the sensitivity probes are a sample of the oracle's own spread, not a bound: up to 2 % of the candidates move by more than
    # 10 x what the probes saw (measured 3 of 256), and the worst candidate the probes call well-posed is at 1.2e-6
    assert summ["n_out_of_tolerance"] <= summ["n"] // 50, summ
    assert summ["max_rel_well_posed"] <= 2e-6 and summ["frac_le_1e-6"] >= 0.5, summ


def test_shell7x5_vns_and_determinism(ev75):
    p = ev75.prob
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, 64, seed=6, wlo=1e-2)
    F0, st0, _ = orc.eval_batch(op, N, Nu, dl, lm, "vns")
    a = ev75.eval_batch(N, Nu, dl, lm, mode="vns")
    b = ev75.eval_batch(N, Nu, dl, lm, mode="vns")
    assert np.array_equal(a["cost"], b["cost"], equal_nan=True)
    ok = (st0 == 0) & (a["status"] == 0) & vns_well_posed(p, lambda r: orc.OracleProblem(p, r=r), N, Nu, dl, lm)
    assert ok.sum() >= 16
    # the Jnu term (|uopt(0)| / |diff(uopt)|)^2 amplifies the last digits of the degenerate open-loop QP
    rel = np.abs(a["cost"][ok] - F0[ok]) / np.abs(F0[ok])
    assert rel.max() < 1e-3 and (rel < 1e-6).mean() >= 0.8, (rel.max(), (rel < 1e-6).mean())


# ---------------------------------------------------------------------------------------------
# Van de Vusse NMPC (BASELINE.json configs[4])
# ---------------------------------------------------------------------------------------------
TOL_NMPC_TRAJ = 1e-5   # relative to the signal's scale factor: two different NLP solvers meet at their tolerances
TOL_NMPC_COST = 1e-5   # (the reference's own fmincon + ode15s path is only reproducible to ~1e-3)


def test_nmpc_golden_parity():
    """GPU (Gauss-Newton SQP + exact box QP, analytic sensitivities) vs the oracle's committed outputs (scipy
    trust-region least squares, finite differences) on the same restated nlmpcmove problem."""
    from mpcgpu.nmpc import vandevusse, NmpcEvaluator, closedloop_toolbox_nmpc
    prob = vandevusse()
    ev = NmpcEvaluator(prob, device=0)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden_nmpc.npz"))
    assert np.abs(gold["x0"] - prob.x0).max() < 1e-12
    out = ev.eval_batch(gold["N"], gold["Nu"], gold["delta"], gold["lam"], mode="gam", traj=True)
    assert (out["status"] == 0).all(), out["status"]
    sy, su = prob.sy[None, :, None], prob.su[None, :, None]
    for k, ref, sc in (("y", gold["y"], sy), ("u", gold["u"], su), ("yopt", gold["yopt"], sy), ("uopt", gold["uopt"], su)):
        err = (np.abs(out[k] - ref) / sc).max()
        assert err < TOL_NMPC_TRAJ, (k, err)
    rel = np.abs(out["cost"] - gold["gam"]) / np.abs(gold["gam"])
    assert rel.max() < TOL_NMPC_COST, rel
    v = ev.eval_batch(gold["N"][:2], gold["Nu"][:2], gold["delta"][:2], gold["lam"][:2], mode="vns")
    relv = np.abs(v["cost"] - gold["vns"]) / np.abs(gold["vns"])
    # Jnu = sum (|uopt(0)| / |diff(uopt)|)^2 (VNS2.m:183-191) amplifies the solvers' 1e-7 disagreement on the plan by
    # |uopt| / |diff| squared (measured 5e-4 .. 1e-3 on these two candidates)
    assert relv.max() < 5e-3, relv
    # reference-shaped call
    y, u, yo, uo = closedloop_toolbox_nmpc(ev, None, None, prob.r, 10, [2, 2], [1, 1], [0.1, 0.1], prob.nit)
    assert (np.abs(y - gold["y"][1]) / prob.sy[:, None]).max() < TOL_NMPC_TRAJ
    ev.close()


def test_nmpc_population_properties():
    from mpcgpu.nmpc import vandevusse, NmpcEvaluator, synthetic_nmpc_population
    prob = vandevusse()
    ev = NmpcEvaluator(prob, device=0)
    N, Nu, dl, lm = synthetic_nmpc_population(prob, 2048, seed=0)
    a = ev.eval_batch(N, Nu, dl, lm, mode="gam")
    ok = (a["status"] == 0) | (a["status"] == 5)
    assert ok.mean() > 0.98 and np.isfinite(a["cost"][a["status"] == 0]).all() and (a["cost"][ok] >= 0).all()
    assert np.isinf(a["cost"][a["status"] == 5]).all()     # crossed an OV / state bound the Toolbox would have kept: rejected by default
    b = ev.eval_batch(N, Nu, dl, lm, mode="gam")
    assert np.array_equal(a["cost"], b["cost"], equal_nan=True)
    perm = np.random.default_rng(0).permutation(len(N))
    c = ev.eval_batch(N[perm], Nu[perm], dl[perm], lm[perm], mode="gam")
    assert np.array_equal(a["cost"][perm], c["cost"], equal_nan=True)
    t = ev.eval_batch(N[:128], Nu[:128], dl[:128], lm[:128], mode="raw")
    u = t["u"]
    assert (u >= prob.umin[None, :, None] - 1e-9).all() and (u <= prob.umax[None, :, None] + 1e-9).all()
    # illegal horizons
    o = ev.eval_batch([5, 40, 6], [5, 3, 2], np.ones((3, 2)), np.ones((3, 2)), mode="gam")
    assert list(o["status"]) == [4, 4, 0] and np.isnan(o["cost"][:2]).all()
    # spot-check a few small candidates against the oracle itself (seconds each)
    from oracle import nmpc_oracle as no
    idx = np.argsort(N * Nu)[:3]
    for i in idx:
        g0, st0 = no.gam_cost(prob, N[i], Nu[i], dl[i], lm[i])
        assert (np.abs(a["cost"][i] - g0) / np.abs(g0)).max() < 1e-4, (i, a["cost"][i], g0)
    ev.close()


def test_batched_tuner_on_shell3x3(ev3):
    """the hybrid tuner driven in populations (mpcgpu/tuner.py).  Starting from the reference's
    start point (N = 127, Nu = 2, delta = lambda = 1, MPCTuning.m:283-302) it must end at a legal tuning whose GAM
    cost is well below the start point's, within a few thousand closed-loop evaluations per outer iteration."""
    from mpcgpu import tuner
    p = ev3.prob
    lines = []
    out = tuner.tune_linear(ev3, w=[0.05, 0.40, 0.55], log=lines.append, pop=256, iters=8, max_outer=3)   # w: Shell3x3.m:161
    assert p.valid(int(out["N"]), int(out["Nu"].max()))
    g0 = ev3.eval_batch([127], [2], np.ones((1, 3)), np.ones((1, 3)), mode="gam")["cost"][0]
    g1 = ev3.eval_batch([out["N"]], [int(out["Nu"].max())], out["delta"][None], out["lam"][None], mode="gam")["cost"][0]
    assert g1.sum() < 0.5 * g0.sum(), (g0, g1, out)      # a short search budget; the point is the batched mechanics
    assert out["evaluations"] > 2000 and any(l.startswith("Fvns=") for l in lines)


# ---------------------------------------------------------------------------------------------
# Other plant shapes: every template instantiation (nu = 1, 2, 4) of the closed-loop kernels
# ---------------------------------------------------------------------------------------------
def _synthetic_problem(ny, nu, nd, nit, seed, soft=False):
    """A random stable FOPDT plant with hard MV limits (and soft output bands if `soft`), in the shape MPCTuning.m
    hands to the evaluator; not a reference case -- it exists to run the nu = 1, 2, 4 kernels against the oracle."""
    from mpcgpu.plant import c2d_fopdt
    from mpcgpu.problems import LinearProblem
    rng = np.random.default_rng(seed)
    nw = nu + nd
    K = rng.uniform(0.5, 2.0, size=(ny, nw)) * rng.choice([-1.0, 1.0], size=(ny, nw))
    K[np.arange(min(ny, nu)), np.arange(min(ny, nu))] = np.abs(K[np.arange(min(ny, nu)), np.arange(min(ny, nu))]) + 1.5
    tau = rng.uniform(5.0, 30.0, size=(ny, nw)); theta = rng.uniform(0.0, 9.0, size=(ny, nw))
    ch = c2d_fopdt(K, tau, theta, 2.0)
    r = np.zeros((nit, ny))
    for i in range(ny):
        r[10 + 15 * i:, i] = 0.3 * (i + 1) * (-1) ** i
    v = np.zeros((nit, nd)); v[nit // 2:, :] = 0.4
    yref = r.T.copy()
    dmin = ch.d[:, :nu].min(axis=1).astype(np.int32)
    big = np.inf
    return LinearProblem(f"synthetic{ny}x{nu}", 2.0, nit, ny, nu, nd, ch, -0.6 * np.ones(nu), 0.8 * np.ones(nu), -0.07 * np.ones(nu),
                         0.09 * np.ones(nu), (-0.35 * np.ones(ny) if soft else -big * np.ones(ny)), (0.5 * np.ones(ny) if soft else big * np.ones(ny)),
                         np.ones(ny), np.ones(ny), np.ones(nu), np.ones(ny), 1e3, r, v, yref, dmin, np.zeros(ny, dtype=bool), np.ones(ny), np.ones(nw))


@pytest.mark.parametrize("ny,nu,nd,soft", [(1, 1, 0, False), (2, 2, 1, False), (3, 4, 0, False), (4, 4, 1, False), (2, 1, 1, True), (3, 2, 0, True)])
def test_other_plant_shapes(ny, nu, nd, soft):
    p = _synthetic_problem(ny, nu, nd, 160, seed=100 + 10 * ny + nu, soft=soft)
    ev = mpcgpu.Evaluator(p, device=0)
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, 40, seed=ny + nu, wlo=1e-2 if soft else 1e-3, whi=3.0)
    N[0], Nu[0] = 127, 15
    N[1], Nu[1] = max(int(p.dmin.max()) + 2, 12), 2
    g0, st0, stats = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    out = ev.eval_batch(N, Nu, dl, lm, mode="gam", traj=True)
    ok = (st0 == 0) & (out["status"] == 0)
    assert ok.sum() >= 36 and stats[2] > 0, (st0, out["status"])
    sens = oracle_sensitivity(op, N[ok], Nu[ok], dl[ok], lm[ok], "gam", g0[ok])
    check_cost(out["cost"][ok], g0[ok], sens, p.name, min_strict=0.7)
    idx = np.where(ok)[0]
    for c in idx[sens < 1e-8][:10]:
        y, u, ys, uo, rc, _ = orc.closedloop(op, N[c], Nu[c], dl[c], lm[c])
        for k, b in zip(("y", "u", "ys", "uopt"), (y, u, ys, uo)):
            assert np.abs(out[k][c] - b).max() < TOL_TRAJ, (c, k)
    if p.square and not soft:
        F0, s0, _ = orc.eval_batch(op, N[:12], Nu[:12], dl[:12], lm[:12], "vns")
        F = ev.eval_batch(N[:12], Nu[:12], dl[:12], lm[:12], mode="vns")
        okv = (s0 == 0) & (F["status"] == 0) & vns_well_posed(p, lambda r: orc.OracleProblem(p, r=r), N[:12], Nu[:12], dl[:12], lm[:12])
        rel = np.abs(F["cost"][okv] - F0[okv]) / np.abs(F0[okv])
        assert okv.sum() >= 4 and rel.max() < 1e-5, rel
    ev.close()


def test_nmpc_lane_mappings(monkeypatch):
    """The lane mappings of the NMPC kernel (a group of 16 lanes per run: the default; MPCGPU_NMPC_GROUP=8: eight; MPCGPU_NMPC_THREAD_PER_RUN=1: thread per run)
    compute the same thing."""
    from mpcgpu.nmpc import vandevusse, NmpcEvaluator
    prob = vandevusse()
    ev = NmpcEvaluator(prob, device=0)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden_nmpc.npz"))
    a = ev.eval_batch(gold["N"], gold["Nu"], gold["delta"], gold["lam"], mode="gam", traj=True)
    monkeypatch.setenv("MPCGPU_NMPC_GROUP", "8")
    g8 = ev.eval_batch(gold["N"], gold["Nu"], gold["delta"], gold["lam"], mode="gam", traj=True)
    assert (g8["status"] == 0).all() and (np.abs(a["u"] - g8["u"]) / prob.su[None, :, None]).max() < 1e-6
    monkeypatch.delenv("MPCGPU_NMPC_GROUP")
    monkeypatch.setenv("MPCGPU_NMPC_THREAD_PER_RUN", "1")
    b = ev.eval_batch(gold["N"], gold["Nu"], gold["delta"], gold["lam"], mode="gam", traj=True)
    assert (b["status"] == 0).all()
    assert (np.abs(b["y"] - gold["y"]) / prob.sy[None, :, None]).max() < TOL_NMPC_TRAJ
    assert (np.abs(a["u"] - b["u"]) / prob.su[None, :, None]).max() < 1e-6
    v = ev.eval_batch(gold["N"][:2], gold["Nu"][:2], gold["delta"][:2], gold["lam"][:2], mode="vns")
    assert (np.abs(v["cost"] - gold["vns"]) / np.abs(gold["vns"])).max() < 5e-3
    ev.close()


def test_two_phase_mode_on_limit_cycles(ev3, monkeypatch):
    """The short-horizon family (N = 7..9) of the 32768-candidate population limit-cycles with period 2; k_sim re-enters
    the parked factor of the other phase.  Same costs as the oracle and as the kernel with the
    mode switched off, and several times fewer active-set iterations."""
    p = ev3.prob
    op = orc.OracleProblem(p)
    Ng, Nug, dlg, lmg = synthetic_population(p, 32768, seed=0)
    idx = np.array([4169, 12612, 26241, 31292, 11416, 26303])
    N, Nu, dl, lm = Ng[idx], Nug[idx], dlg[idx], lmg[idx]
    assert (N <= 9).all()
    g0, st0, _ = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    its = {}
    cost = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("MPCGPU_TWO_PHASE", flag)
        c0 = ev3.counters()
        out = ev3.eval_batch(N, Nu, dl, lm, mode="gam")
        c1 = ev3.counters()
        assert (out["status"] == 0).all()
        its[flag] = c1["as_iterations"] - c0["as_iterations"]
        cost[flag] = out["cost"]
    assert its["1"] * 3 < its["0"], its
    sens = oracle_sensitivity(op, N, Nu, dl, lm, "gam", g0)
    check_cost(cost["1"], g0, sens, "shell3x3 limit cycles", min_strict=0.0)
    check_cost(cost["0"], g0, sens, "shell3x3 limit cycles, mode off", min_strict=0.0)
    check_cost(cost["1"], cost["0"], sens, "shell3x3 limit cycles, mode on vs off", min_strict=0.0)


def test_dtc_filter_design_on_device():
    """the robustness filter of every candidate designed on the device from (alfa, raio)
    (mimofilter.m:33-50, filtro_siso.m:26-96 as a batched op) gives the sweep the host design gives."""
    from mpcgpu.dtcgpc import woodberry_dtc, synthetic_dtc_population, DtcEvaluator
    prob = woodberry_dtc(deltak=0.1, deltaL=1.0)
    ev = DtcEvaluator(prob, device=0)
    p, m, dl, lm, alfa, raio = synthetic_dtc_population(prob, 512, seed=3)
    raio[:64] = 0.999      # no pole counts as slow: Fr = 1
    raio[64:128] = 0.5     # every pole does
    a = ev.eval_batch(p, m, dl, lm, alfa=alfa, raio=raio, design_on_device=False, traj=True)
    b = ev.eval_batch(p, m, dl, lm, alfa=alfa, raio=raio, design_on_device=True, traj=True)
    assert np.array_equal(a["status"], b["status"])
    ok = a["status"] == 0
    assert ok.sum() > 400
    stable = ok & np.isfinite(a["ise"]).all(axis=1) & (np.abs(a["y"]).max(axis=(1, 2)) < 1e3)
    rel = np.abs(a["ise"][stable] - b["ise"][stable]) / np.maximum(np.abs(a["ise"][stable]), 1e-300)
    assert rel.max() < 1e-9, rel.max()
    assert np.abs(a["y"][stable] - b["y"][stable]).max() < 1e-8
    ev.close()
