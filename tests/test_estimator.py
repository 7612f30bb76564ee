"""Plant-model mismatch validation run on the CPU: the restated default estimator (mpcgpu/estimator.py) and
the oracle's estimator loop (oracle/mpc_oracle.c orc_closedloop_est).  The GPU run of the same is
tests/test_gpu_full.py::test_mismatch_validation_run."""
import numpy as np
import pytest

import mpcgpu
from mpcgpu import estimator as est
from oracle import oracle as orc


@pytest.mark.parametrize("case", ["shell3x3", "woodberry"])
def test_estimator_model_reproduces_the_channels(case):
    """(A, Bu, C) of E1 is the library's plant: its step response equals the channel recursion of mpcgpu.plant.simulate."""
    from mpcgpu.plant import simulate
    p = {"shell3x3": lambda: mpcgpu.shell3x3(2), "woodberry": mpcgpu.woodberry}[case]()
    hl = est.history_length(p)
    A, Bu, Cm, G = est.estimator_model(p, hl)
    nit = 80
    w = np.zeros((nit, p.nu + p.nd)); w[3:, 0] = 1.0; w[10:, p.nu - 1] = -0.5
    y_ref = simulate(p.plant, w)
    x = np.zeros(A.shape[0]); y = np.zeros((nit, p.ny))
    for k in range(nit):
        y[k] = Cm @ x
        x = A @ x + Bu @ w[k, :p.nu]
    np.testing.assert_allclose(y, y_ref, atol=1e-13)


def test_default_gain_is_a_stable_observer_with_integral_action():
    p = mpcgpu.shell3x3(2)
    hl = est.history_length(p)
    A, Bu, Cm, G = est.estimator_model(p, hl)
    M = est.default_estimator_gain(p, hl)
    assert M.shape == (p.ny * (p.nu + p.nd) + p.nu * hl + p.ny, p.ny)
    rho = np.abs(np.linalg.eigvals(A - A @ M @ Cm)).max()          # error dynamics of x(k+1|k)
    assert rho < 1.0 - 1e-6, rho
    # the output-disturbance states pick up a constant output offset completely (integral action -> offset-free tracking)
    x = np.zeros(A.shape[0]); off = np.array([0.3, -0.2, 0.1])
    for _ in range(3000):
        x = x + M @ (off - Cm @ x)
        x = A @ x
    np.testing.assert_allclose(Cm @ x, off, atol=1e-8)


@pytest.mark.parametrize("case", ["shell3x3", "woodberry", "shell7x5"])
def test_oracle_estimator_loop(case):
    p = {"shell3x3": lambda: mpcgpu.shell3x3(2), "woodberry": mpcgpu.woodberry, "shell7x5": mpcgpu.shell7x5}[case]()
    plant = {"shell3x3": est.shell3x3_real_plant, "woodberry": est.woodberry_real_plant, "shell7x5": est.shell7x5_real_plant}[case]()
    assert plant.a.shape == p.plant.a.shape and np.allclose(plant.a, p.plant.a)        # the reference's errors are on gains (and dead times)
    assert np.abs(plant.dcgain() / p.plant.dcgain() - 1.0).max() > 0.05
    op = orc.OracleProblem(p)
    hl = est.history_length(p, plant)
    M = est.default_estimator_gain(p, hl)
    cand = (19, 7, np.zeros(7), np.array([0.056, 0.0167, 1.61])) if case == "shell7x5" else (12, 4, np.full(p.ny, 0.5), np.full(p.nu, 0.3))
    y0, u0, _, _, rc0, _ = orc.closedloop(op, *cand, open_loop=False)
    y1, u1, rc1, _ = orc.closedloop_est(op, p.plant, M, *cand)
    assert rc0 == 0 and rc1 == 0
    assert np.abs(y1 - y0).max() < 1e-12 and np.abs(u1 - u0).max() < 1e-12           # plant == model: zero innovation
    y2, u2, rc2, _ = orc.closedloop_est(op, plant, M, *cand)
    assert rc2 == 0 and np.abs(y2 - y0).max() > 1e-3
    assert (u2 >= p.umin[:, None] - 1e-9).all() and (u2 <= p.umax[:, None] + 1e-9).all()
    if case != "shell7x5":
        assert np.abs(y2[:, -1] - p.r[-1]).max() < 5e-3                                 # offset-free despite 20-30 % gain errors
