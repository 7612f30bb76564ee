"""GPU parity of the single-shooting NMPC sweep (`Explicit NMPC/ClosedLoopNMPC.m`, NMPC_Controller.m):
k_ssnmpc through the C ABI against the committed oracle closed loops, against the host build of the same source on a seeded
population, edge cases, the Python mirror of the reference's signature and the MEX commands.  (Named to run after the other
GPU files: it was written after the round's last GPU session.)"""
import os

import numpy as np
import pytest

import mpcgpu
from mpcgpu import ssnmpc
from oracle import nmpc_port

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def prob():
    return mpcgpu.explicit_nmpc()


@pytest.fixture(scope="module")
def ev(prob):
    e = mpcgpu.SsnmpcEvaluator(prob, device=0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden_ssnmpc.npz"))


def test_golden_closed_loops(prob, ev, gold):
    """Tolerances: trajectories 1e-5 (y absolute in mol/l and degC; u relative to the MV range), cost 1e-6 relative."""
    g = ev.eval_batch(gold["N"], gold["Nu"], gold["Q"], gold["W"], traj=True)
    assert (g["status"] == 0).all()
    assert np.abs(g["y"] - gold["y"]).max() < 1e-5
    assert (np.abs(g["u"] - gold["u"]) / (prob.ub - prob.lb)[None, :, None]).max() < 1e-5
    assert (np.abs(g["cost"] - gold["cost"]) / gold["cost"]).max() < 1e-6
    n = ev.eval_batch(gold["N"][:1], gold["Nu"][:1], gold["Q"][:1], gold["W"][:1], traj=True, noise=gold["noise"])
    assert n["status"][0] == 0 and np.abs(n["y"][0] - gold["y_noise"]).max() < 1e-5
    assert (np.abs(n["u"][0] - gold["u_noise"]) / (prob.ub - prob.lb)[:, None]).max() < 1e-5


def test_population_vs_host_build(prob, ev):
    """512 seeded candidates: the device against the same source compiled for the host.  The two differ in FMA contraction
    only, but a closed loop amplifies a last-bit difference of one controller call; how much is measured per candidate on the
    host (largest trajectory change under four relative perturbations of Q, W of 1e-9 ... 1e-12: 1e-9 at the median, 1e-6 at
    the 99th percentile -- tests/test_ssnmpc_oracle.py).  Strict tolerance (trajectories 1e-5, cost 1e-6) where that spread is
    below 1e-8, elsewhere 1000 x the candidate's own spread is added (dry run of this test with a second host build --
    no FMA contraction, -O1 -- in place of the device: 250 x margin on every candidate but one).  One candidate in 512
    (seed 11: #16, N = 10, Nu = [1 3], W ~ 1e-5) is bistable: two builds of the same source follow different branches of the
    non-convex controller problem while the perturbation probes stay on one -- hence 99 %, not all."""
    N, Nu, Q, W = mpcgpu.synthetic_ssnmpc_population(prob, 512, seed=11)
    g = ev.eval_batch(N, Nu, Q, W, traj=True)
    c0, s0, y0, u0 = nmpc_port.ssnmpc_eval_batch(prob, N, Nu, Q, W, traj=True)
    assert np.array_equal(g["status"], s0) and (s0 == 0).all()
    spread = np.zeros(512); spread_c = np.zeros(512)
    for qs, ws in ((1 + 1e-9, 1), (1 - 1e-9, 1), (1, 1 + 1e-9), (1 + 1e-12, 1 - 1e-12)):
        c1, _, y1, _ = nmpc_port.ssnmpc_eval_batch(prob, N, Nu, Q * qs, W * ws, traj=True)
        spread = np.maximum(spread, np.abs(y1 - y0).max(axis=(1, 2)))
        spread_c = np.maximum(spread_c, (np.abs(c1 - c0) / np.abs(c0)).max(axis=1))
    dy = np.abs(g["y"] - y0).max(axis=(1, 2))
    du = (np.abs(g["u"] - u0) / (prob.ub - prob.lb)[None, :, None]).max(axis=(1, 2))
    rel = (np.abs(g["cost"] - c0) / np.abs(c0)).max(axis=1)
    well = spread < 1e-8
    assert well.mean() > 0.5, well.mean()
    strict = (dy < 1e-5) & (du < 1e-5) & (rel < 1e-6)
    assert strict[well].mean() >= 0.99, (strict[well].mean(), np.where(well & ~strict)[0])
    relaxed = (dy <= 1e-5 + 1e3 * spread) & (rel <= 1e-6 + 1e3 * spread_c)
    assert relaxed.mean() >= 0.99, (relaxed.mean(), np.where(~relaxed)[0])
    assert strict.mean() > 0.9, strict.mean()
    print(f"ssnmpc population: {well.mean():.3f} well-posed, strict on {strict.mean():.4f}, outside the spread-scaled tolerance: {np.where(~relaxed)[0]}")
    assert (g["u"] >= prob.lb[None, :, None] - 1e-12).all() and (g["u"] <= prob.ub[None, :, None] + 1e-12).all()
    # the launch sorts by horizon: results do not depend on the order of the population; cost-only == cost with trajectories
    perm = np.random.default_rng(0).permutation(512)
    h = ev.eval_batch(N[perm], Nu[perm], Q[perm], W[perm])
    assert np.array_equal(h["cost"], g["cost"][perm])
    c = ev.counters()
    assert c["kernel_launches"] >= 2 and c["qp_solves"] >= 2 * 512 * (prob.nit - prob.inK + 1)


def test_edge_cases(prob, ev):
    N = np.array([5, 0, 3, 40, 20], dtype=np.int32)
    Nu = np.array([[2, 2], [1, 1], [4, 1], [2, 2], [16, 15]], dtype=np.int32)
    g = ev.eval_batch(N, Nu, np.ones((5, 2)), np.full((5, 2), 1e-3))
    assert list(g["status"]) == [0, 4, 4, 4, 4] and np.isnan(g["cost"][1:]).all() and np.isfinite(g["cost"][0]).all()
    e = ev.eval_batch(np.zeros(0, dtype=np.int32), np.zeros((0, 2), dtype=np.int32), np.zeros((0, 2)), np.zeros((0, 2)))
    assert e["cost"].shape == (0, 2)
    # set-point override per call (ClosedLoopNMPC takes r per call); a constant set-point at the steady state keeps the loop there
    r = np.tile(prob.x0[1:3, None], (1, prob.nit))
    s = ev.eval_batch([5], [[2, 2]], [ssnmpc.BASE_Q], [ssnmpc.BASE_W], traj=True, r=r)
    assert s["cost"].max() < 1e-12 and np.abs(s["u"][0] - prob.u0[:, None]).max() < 1e-5


def test_reference_signature_and_mex(prob, ev, gold):
    y, u = mpcgpu.ClosedLoopNMPC(ev, [2, 3], prob.u0, prob.r, 5, [2, 2], ssnmpc.BASE_Q, ssnmpc.BASE_W, prob.nit, prob.ub, prob.lb, prob.inK, prob.Ts)
    assert y.shape == (2, prob.nit) and np.abs(y - gold["y"][0]).max() < 1e-5
    with pytest.raises(mpcgpu.MpcGpuError):
        mpcgpu.ClosedLoopNMPC(ev, [1, 3], prob.u0, prob.r, 5, [2, 2], ssnmpc.BASE_Q, ssnmpc.BASE_W, prob.nit, prob.ub, prob.lb, prob.inK, prob.Ts)
    from test_mex import Mex
    mex = Mex()
    Ps = dict(nit=float(prob.nit), pmax=float(prob.pmax), inK=float(prob.inK), Ts=float(prob.Ts), x_control=np.array([2.0, 3.0]),
              x0=prob.x0, u0=prob.u0, lb=prob.lb, ub=prob.ub, r=np.ascontiguousarray(prob.r).ravel())
    (hs,) = mex.call(1, "ssnmpc_create", Ps)
    ym, um = mex.call(2, "ssnmpc_closedloop", hs, prob.r, 5.0, np.array([2.0, 2.0]), np.array(ssnmpc.BASE_Q), np.array(ssnmpc.BASE_W))
    assert np.array_equal(ym, y) and np.array_equal(um, u)
    yn, un = mex.call(2, "ssnmpc_closedloop", hs, prob.r.T, 5.0, np.array([2.0, 2.0]), np.array(ssnmpc.BASE_Q), np.array(ssnmpc.BASE_W), gold["noise"].T)
    assert np.abs(yn - gold["y_noise"]).max() < 1e-5
    N, Nu, Q, W = mpcgpu.synthetic_ssnmpc_population(prob, 16, seed=2)
    cost, st = mex.call(2, "ssnmpc_eval", hs, N, Nu, Q, W)
    assert np.array_equal(cost, ev.eval_batch(N, Nu, Q, W)["cost"]) and (st == 0).all()
    mex.call(0, "ssnmpc_destroy", hs)
