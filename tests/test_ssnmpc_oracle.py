"""Single-shooting NMPC (`Explicit NMPC/`) without a GPU: the oracle's objective against a hand
evaluation, the kernel's per-run source (csrc/mpc_ssnmpc_core.h compiled for the host, oracle/nmpc_port) against the scipy
minimiser of the restated objective, and against the committed oracle output (tests/golden/oracle_golden_ssnmpc.npz)."""
import os

import numpy as np
import pytest

import mpcgpu
from mpcgpu import ssnmpc
from oracle import nmpc_port, ssnmpc_oracle as so
from oracle.nmpc_oracle import rk4_sample

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "oracle_golden_ssnmpc.npz")


@pytest.fixture(scope="module")
def prob():
    return mpcgpu.explicit_nmpc()


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _par(prob, x, uprev, k, N, Nu, Q, W):
    u = np.tile(np.asarray(uprev, float)[:, None], (1, prob.nit))
    return dict(r=prob.r, N=N, Nu=list(Nu), Ts=prob.Ts, my=2, ny=2, x_control=[1, 2], lb=np.repeat(prob.lb, Nu), ub=np.repeat(prob.ub, Nu),
                Q=Q, W=W, ub1=prob.ub, lb1=prob.lb, nsub=prob.nsub, integrator="rk4", u=u, x0plant=np.asarray(x, float), k=k)


def test_problem_data(prob):
    # main.m:20-56
    assert prob.nit == 150 and prob.inK == 4 and prob.Ts == 0.05
    assert np.abs(so.vandevusse_model(prob.x0, prob.u0)).max() < 1e-9          # fsolve steady state (:39)
    assert prob.r[0, 8] == prob.x0[1] and prob.r[0, 9] == 1.2 and prob.r[0, 48] == 1.2 and prob.r[0, 49] == 1.0
    assert prob.r[1, 79] == prob.x0[2] and prob.r[1, 80] == 130.0 and prob.r[1, 110] == 120.0


def test_objective_hand_evaluation(prob):
    """NMPC_Controller.m:46-141 on a case small enough to evaluate by hand: N = 2, Nu = [1 2]."""
    x = prob.x0 + np.array([0.1, -0.05, 1.5]); up = np.array([25.0, 125.0]); k = 60
    X = np.array([3.0, -2.0, 4.0])                     # input 1: one offset (held); input 2: two offsets
    Q, W = (2.0, 0.5), (1e-3, 2e-3)
    e, dU = so.objective_terms(X, _par(prob, x, up, k, 2, [1, 2], Q, W))
    x1 = rk4_sample(x, up + np.array([3.0, -2.0]), prob.Ts, prob.nsub)
    x2 = rk4_sample(x1, up + np.array([3.0, 4.0]), prob.Ts, prob.nsub)
    xk = rk4_sample(x, up, prob.Ts, prob.nsub)
    n = x[1:3] - xk[1:3]                               # :106-123 (not zero: minus the model's one-step change)
    want = np.array([prob.r[0, k] - (x1[1] + n[0]), prob.r[0, k] - (x2[1] + n[0]), prob.r[1, k] - (x1[2] + n[1]), prob.r[1, k] - (x2[2] + n[1])])
    assert np.abs(e - want).max() < 1e-14 and np.array_equal(dU, X)
    assert np.abs(n).max() > 1e-4
    # the long-double statement of the same objective (used to polish the oracle's minimiser) is the same function
    Par = _par(prob, x, up, k, 2, [1, 2], Q, W)
    f = (np.repeat(Q, 2) * e * e).sum() + (np.repeat(W, [1, 2]) * X * X).sum()
    assert abs(float(so.objective_ld(X, Par)) - f) < 1e-13 * f


CASES = [  # (state offset, uprev, k, N, Nu, Q, W)
    ((0.0, 0.0, 0.0), (20.0, 130.0), 9, 5, (2, 2), ssnmpc.BASE_Q, ssnmpc.BASE_W),          # the reference's setting at the first set-point jump
    ((0.2, -0.1, 2.0), (40.0, 120.0), 60, 8, (3, 1), (0.5, 2.0), (1e-3, 1e-2)),
    ((0.0, 0.05, -1.0), (149.5, 41.0), 120, 6, (2, 3), (3.0, 1.0), (1e-4, 1e-4)),           # both inputs close to a bound: active box
    ((-0.3, 0.0, 0.5), (5.0, 100.0), 90, 3, (1, 1), (1.0, 1.0), (1e-2, 1e-3)),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_controller_vs_scipy_minimiser(prob, case):
    dx, up, k, N, Nu, Q, W = CASES[case]
    x = prob.x0 + np.array(dx)
    Par = _par(prob, x, up, k, N, Nu, Q, W)
    Xo = so.nmpc_controller(Par)
    Xp, n_sqp, rc = nmpc_port.ssnmpc_controller(prob, x, up, prob.r[:, k], N, Nu, Q, W)
    assert rc == 0 and n_sqp < prob.max_sqp

    def f(X):
        e, d = so.objective_terms(X, Par)
        return (np.repeat(Q, N) * e * e).sum() + (np.repeat(W, Nu) * d * d).sum()
    scale = np.repeat(prob.ub - prob.lb, Nu)
    assert abs(f(Xp) - f(Xo)) <= 1e-10 * max(f(Xo), 1e-6)                     # same minimum ...
    assert (np.abs(Xp - Xo) / scale).max() < 1e-8                              # ... at the same point (flat valley: W ~ 1e-4)
    Xraw = so.nmpc_controller(Par, polish=False)                               # what a double-precision cost-based stop leaves open
    assert (np.abs(Xraw - Xo) / scale).max() < 1e-6
    lo = np.repeat(prob.lb - np.array(up), Nu); hi = np.repeat(prob.ub - np.array(up), Nu)
    assert (Xp >= lo - 1e-12).all() and (Xp <= hi + 1e-12).all()
    if case == 2:
        assert ((Xp - lo < 1e-9) | (hi - Xp < 1e-9)).any()                     # the box is active in this case


def test_port_vs_golden(prob, gold):
    """The kernel's source on the host against the committed oracle closed loops (150 samples, five settings)."""
    cost, st, y, u = nmpc_port.ssnmpc_eval_batch(prob, gold["N"], gold["Nu"], gold["Q"], gold["W"], traj=True)
    assert (st == 0).all()
    assert np.abs(y - gold["y"]).max() < 1e-5                                  # trajectories: 1e-5 (north_star tolerance)
    assert (np.abs(u - gold["u"]) / (prob.ub - prob.lb)[None, :, None]).max() < 1e-5
    assert (np.abs(cost - gold["cost"]) / gold["cost"]).max() < 1e-6           # cost: 1e-6 relative
    assert np.array_equal(y[:, :, :prob.inK - 1], np.broadcast_to(prob.x0[1:3, None], (5, 2, prob.inK - 1)))   # ClosedLoopNMPC.m:64


def test_port_noise_run_vs_golden(prob, gold):
    """ClosedLoopNMPC.m:88-90 with the noise draws as an input."""
    cost, st, y, u = nmpc_port.ssnmpc_eval_batch(prob, gold["N"][:1], gold["Nu"][:1], gold["Q"][:1], gold["W"][:1], noise=gold["noise"], traj=True)
    assert st[0] == 0
    assert np.abs(y[0] - gold["y_noise"]).max() < 1e-5
    assert (np.abs(u[0] - gold["u_noise"]) / (prob.ub - prob.lb)[:, None]).max() < 1e-5
    assert np.abs(y[0] - gold["y"][0]).max() > 1e-3                             # the noise does something


def test_invalid_horizons_and_bounds(prob):
    N = np.array([5, 0, 3, 40, 20], dtype=np.int32)
    Nu = np.array([[2, 2], [1, 1], [4, 1], [2, 2], [16, 15]], dtype=np.int32)   # Nu_j > N; N > pmax; sum Nu > 30
    cost, st, y, u = nmpc_port.ssnmpc_eval_batch(prob, N, Nu, np.ones((5, 2)), np.full((5, 2), 1e-3), traj=True)
    assert list(st) == [0, 4, 4, 4, 4] and np.isnan(cost[1:]).all() and np.isfinite(cost[0]).all()
    assert (u[0] >= prob.lb[:, None] - 1e-12).all() and (u[0] <= prob.ub[:, None] + 1e-12).all()


def test_integrator_choice(prob):
    """RK4 x nsub against a tight implicit solver (what ode23t / ode45 approximate), open loop over 20 samples after a step of
    both inputs (a harder transient than the closed loops see): the stated integrator deviation.  Measured: 4 sub-steps
    3.6e-3 (mol/l, degC / 100), 16 sub-steps 8e-6 -- fourth order (h |lambda| = 0.4 on the jacket mode at nsub = 4); a caller who
    needs the reference's integrator accuracy raises `nsub` (problem field)."""
    import dataclasses
    err = {}
    for nsub in (4, 16):
        xa = prob.x0.copy(); xb = prob.x0.copy(); u = np.array([60.0, 110.0]); e = 0.0
        for _ in range(20):
            xa = so._step(xa, u, prob.Ts, nsub, "rk4"); xb = so._step(xb, u, prob.Ts, nsub, "ivp")
            e = max(e, np.abs((xa - xb) / np.array([1.0, 1.0, 100.0])).max())
        err[nsub] = e
    assert err[4] < 5e-3 and err[16] < 2e-5 and err[16] < err[4] / 100, err


def test_steady_set_point_keeps_the_loop_at_rest(prob):
    import dataclasses
    rest = dataclasses.replace(prob, r=np.tile(prob.x0[1:3, None], (1, prob.nit)))
    cost, st, y, u = nmpc_port.ssnmpc_eval_batch(rest, [5], [[2, 2]], [ssnmpc.BASE_Q], [ssnmpc.BASE_W], traj=True)
    assert st[0] == 0 and cost.max() < 1e-12 and np.abs(u[0] - prob.u0[:, None]).max() < 1e-5


def test_abi_and_mex_fail_loudly_without_gpu(prob):
    """No CPU fallback: the single-shooting entry points refuse without a device; the gateway parses and raises like MATLAB."""
    import torch
    from test_mex import Mex
    mex = Mex()
    with pytest.raises(RuntimeError, match="ssnmpc_create: problem struct expected"):
        mex.call(1, "ssnmpc_create", 1.0)
    Ps = dict(nit=float(prob.nit), inK=float(prob.inK), Ts=float(prob.Ts), x_control=np.array([2.0, 3.0]), x0=prob.x0, u0=prob.u0,
              lb=prob.lb, ub=prob.ub, r=np.ascontiguousarray(prob.r).ravel())
    with pytest.raises(RuntimeError, match="x_control must have 2 entries"):
        mex.call(1, "ssnmpc_create", dict(Ps, x_control=np.array([2.0])))
    with pytest.raises(RuntimeError, match="bad single-shooting NMPC problem"):
        mex.call(1, "ssnmpc_create", dict(Ps, x_control=np.array([2.0, 4.0])))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="mpcgpu:create: no CUDA device"):
            mex.call(1, "ssnmpc_create", Ps)
        with pytest.raises(mpcgpu.MpcGpuError, match="no CPU fallback"):
            mpcgpu.SsnmpcEvaluator(prob)


def test_closed_loop_sensitivity_of_a_population(prob):
    """What `the same result` can mean for a closed loop: the trajectory spread of each candidate under a 1e-9 relative change
    of Q (two runs of the SAME code).  Most loops damp it; a few amplify it by 1e4 and more (long chains of controller calls
    near a bifurcation of the short-horizon loop, e.g. N = 7, Nu = [1 2], W ~ 1e-3) -- those cannot be pinned to 1e-5 by any
    two implementations, which is why the GPU parity test scales its tolerance by this spread."""
    N, Nu, Q, W = mpcgpu.synthetic_ssnmpc_population(prob, 128, seed=11)
    c0, s0, y0, u0 = nmpc_port.ssnmpc_eval_batch(prob, N, Nu, Q, W, traj=True)
    c1, s1, y1, u1 = nmpc_port.ssnmpc_eval_batch(prob, N, Nu, Q * (1 + 1e-9), W, traj=True)
    assert (s0 == 0).all() and (s1 == 0).all()
    spread = np.abs(y1 - y0).max(axis=(1, 2))
    assert np.median(spread) < 1e-8 and (spread < 1e-8).mean() > 0.6 and (spread < 1e-5).mean() > 0.95
    # the candidate that was first drawn for the golden set and replaced: its loop amplifies 1e-9 to > 1e-6
    ca, _, ya, _ = nmpc_port.ssnmpc_eval_batch(prob, [7], [[1, 2]], [[2.0, 0.5]], [[1e-3, 2e-4]], traj=True)
    cb, _, yb, _ = nmpc_port.ssnmpc_eval_batch(prob, [7], [[1, 2]], [[2.0 * (1 + 1e-9), 0.5 * (1 + 1e-9)]], [[1e-3, 2e-4]], traj=True)
    assert np.abs(ya - yb).max() > 1e-7
