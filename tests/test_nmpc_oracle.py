"""Nonlinear path (BASELINE.json configs[4]) on the CPU: problem data KATs from the reference script, the oracle's
model and integrator against independent scipy implementations, and the committed oracle fixture."""
import os

import numpy as np
import pytest
from scipy.integrate import solve_ivp

from mpcgpu.nmpc import vandevusse, synthetic_nmpc_population
from oracle import nmpc_oracle as no

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def prob():
    return vandevusse()


def test_problem_data_matches_the_reference_script(prob):
    # VanDeVusse_NMPC.m:35-90,139-151 and the saved nlmpc object (ScaleFactors 150,110 / 1.2,110)
    assert (prob.Ts, prob.nit) == (0.05, 60)
    np.testing.assert_array_equal(prob.su, [150.0, 110.0]); np.testing.assert_array_equal(prob.sy, [1.2, 110.0])
    np.testing.assert_array_equal(prob.umin, [0.0, 40.0]); np.testing.assert_array_equal(prob.umax, [150.0, 150.0])
    assert np.abs(no.vandevusse_model(prob.x0, prob.u0)).max() < 1e-10          # fsolve steady state (:79)
    assert abs(prob.x0[1] - 0.9052) < 1e-3 and abs(prob.x0[2] - 134.95) < 1e-2
    assert prob.r[0, 8] == prob.x0[1] and prob.r[0, 9] == 1.0 and prob.r[1, 39] == prob.x0[2] and prob.r[1, 40] == 130.0
    # Yref = lsim(Pref, r - x0(xc)) + x0(xc): first-order lags of 0.05 h and 0.0875 h, zero-order hold
    assert prob.yref[0, 9] == prob.x0[1] and abs(prob.yref[0, 10] - (prob.x0[1] + (1 - np.exp(-1.0)) * (1.0 - prob.x0[1]))) < 1e-14
    assert abs(prob.yref[1, -1] - 130.0) < 1e-3


def test_rhs_is_the_reference_model():
    # hand-evaluated vandevusse_model.m:59-77 at a round state
    x = np.array([2.0, 1.0, 110.0]); u = np.array([30.0, 100.0])
    T = 110.0 + 273.15
    k1 = 1.287e12 * np.exp(-9758.3 / T); k3 = 9.043e9 * np.exp(-8560.0 / T)
    f = no.vandevusse_model(x, u)
    assert abs(f[0] - (30 * (5.1 - 2) - k1 * 2 - k3 * 4)) < 1e-12
    assert abs(f[1] - (-30 * 1 + k1 * 2 - k1 * 1)) < 1e-12
    beta = 4032.0 * 0.215 / (0.9342 * 3.01 * 10)
    assert abs(f[2] - ((k1 * 2 * -4.2 + k1 * 1 * 11.0 + k3 * 4 * 41.85) / (0.9342 * 3.01) + 30 * (130 - 110) + beta * (100 - 110))) < 1e-10


def test_rk4_with_four_substeps_resolves_the_plant(prob):
    """N4: against a tight adaptive integration of the same ODE over one sample, from the steady state with a
    large input step (the reference integrates the plant with ode15s at RelTol 1e-3)."""
    u = np.array([60.0, 100.0])
    ref = solve_ivp(lambda t, x: no.vandevusse_model(x, u), [0, prob.Ts], prob.x0, rtol=1e-12, atol=1e-14, method="LSODA").y[:, -1]
    got = no.rk4_sample(prob.x0, u, prob.Ts, prob.nsub)
    assert np.abs((got - ref) / ref).max() < 2e-3     # measured 1.0e-3: the level of the reference's own integrator tolerance
    assert np.abs((no.rk4_sample(prob.x0, u, prob.Ts, 1) - ref) / ref).max() > np.abs((got - ref) / ref).max()


def test_controller_call_is_a_local_optimum(prob):
    """The oracle's nlmpcmove solution satisfies first-order optimality of N2 under the MV bounds."""
    p, m = 8, 3
    delta, lam = np.array([1.0, 1.0]), np.array([0.1, 0.1])
    r = np.array([1.0, 130.0])
    v = no.nlmpcmove(prob, prob.x0, prob.u0, r, p, m, delta, lam)

    def cost(vv):
        vv = vv.reshape(m, 2); x = prob.x0.copy(); J = 0.0
        for i in range(p):
            x = no.rk4_sample(x, vv[min(i, m - 1)], prob.Ts, prob.nsub)
            J += (((delta / prob.sy) * (r - x[1:3])) ** 2).sum()
        prev = prob.u0
        for c in range(m):
            J += (((lam / prob.su) * (vv[c] - prev)) ** 2).sum(); prev = vv[c]
        return J
    g = np.zeros(2 * m); v0 = v.ravel()
    for i in range(2 * m):
        h = 1e-5 * prob.su[i % 2]
        e = np.zeros(2 * m); e[i] = h
        g[i] = (cost(v0 + e) - cost(v0 - e)) / (2 * h)
    lo = np.tile(prob.umin, m); hi = np.tile(prob.umax, m)
    free = (v0 > lo + 1e-9) & (v0 < hi - 1e-9)
    assert np.abs(g[free] * np.tile(prob.su, m)[free]).max() < 1e-6
    assert (g[v0 <= lo + 1e-9] >= -1e-9).all() and (g[v0 >= hi - 1e-9] <= 1e-9).all()


def test_golden_fixture_reproduces(prob):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden_nmpc.npz"))
    c = 0   # the reference's own tuned result for this case: N=3, Nu=2
    y, u, yo, uo, st = no.closedloop_toolbox_nmpc(prob, prob.r, gold["N"][c], gold["Nu"][c], gold["delta"][c], gold["lam"][c])
    assert st == 0
    assert np.abs(y - gold["y"][c]).max() < 1e-8 and np.abs(u - gold["u"][c]).max() < 1e-6
    # the tuned controller does its job: both outputs on their set-points at the end of the run
    assert abs(y[0, 35] - 1.0) < 1e-2 and abs(y[1, -1] - 130.0) < 0.5


def test_population_is_legal(prob):
    N, Nu, dl, lm = synthetic_nmpc_population(prob, 256, seed=0)
    assert all(prob.valid(int(a), int(b)) for a, b in zip(N, Nu))
    assert N.max() <= 31 and Nu.max() <= 15 and dl.min() >= 1e-3 and lm.max() <= 10


def test_unenforced_bounds_are_inactive_on_the_benchmark_population():
    """The restated NLP (N3) does not enforce the OV / state bounds of VanDeVusse_NMPC.m:140-145.  On the benchmark's own
    seeded population no controller call's optimal plan predicts a state outside them (tools/nmpc_bounds): the unconstrained
    optimum is feasible for nlmpc's constrained problem, so the two have the same solution in every call."""
    import os, runpy, sys, io, contextlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    buf = io.StringIO()
    argv = sys.argv
    sys.argv = ["run.py", "512"]
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(root, "tools", "nmpc_bounds", "run.py"), run_name="__main__")
    finally:
        sys.argv = argv
    out = buf.getvalue()
    assert "outside its bounds: 0 (0.00%)" in out, out
