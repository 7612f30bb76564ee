"""Known-answer tests against the only numbers the reference pins:
the scaled discrete plants, limits and scale factors stored inside MPC-Tuning/*.mat
(decoded once by tests/golden/make_fixture_kats.py into tests/golden/fixture_kats.json)."""
import numpy as np
import pytest

from mpcgpu import shell3x3, shell7x5, woodberry, c2d_fopdt, simulate, cond_min
from mpcgpu.plant import Channels


def _f(v):
    return np.array([np.inf if x == "inf" else -np.inf if x == "-inf" else x for x in v], dtype=float)


@pytest.mark.parametrize("fname,builder", [
    ("Shell3x3_Tuning_25Jul2023_12_06.mat", lambda: shell3x3(1)),
    ("Shell3x3_Tuning_Caso2.mat", lambda: shell3x3(2)),
    ("Shell7x5_Tuning_14Sep2024_14_22.mat", shell7x5),
    ("Shell7x5_Tuning_25Jul2023_12_18.mat", shell7x5),
])
def test_scaled_plant_matches_reference_object(kats, fname, builder):
    f = kats[fname]
    p = builder()
    ny, nw = f["ny"], f["nw"]
    assert p.plant.shape == (ny, nw)
    for i in range(ny):
        for j in range(nw):
            num, den = f["num"][i][j], f["den"][i][j]
            assert abs(num[0] - p.plant.b0[i, j]) < 5e-16
            assert abs(num[1] - p.plant.b1[i, j]) < 5e-16
            assert abs(-den[1] - p.plant.a[i, j]) < 5e-16 and den[0] == 1.0
            assert int(f["iodelay"][i][j]) == int(p.plant.d[i, j])
    assert f["Ts"] == p.Ts


@pytest.mark.parametrize("fname,builder", [
    ("Shell3x3_Tuning_Caso2.mat", lambda: shell3x3(2)),
    ("Shell7x5_Tuning_14Sep2024_14_22.mat", shell7x5),
])
def test_scaled_limits_and_scale_factors(kats, fname, builder):
    """MPCTuning.m:170-199."""
    f = kats[fname]
    p = builder()
    np.testing.assert_allclose(p.umin, _f(f["MV"]["Min"]), rtol=1e-14)
    np.testing.assert_allclose(p.umax, _f(f["MV"]["Max"]), rtol=1e-14)
    np.testing.assert_allclose(p.dumin, _f(f["MV"]["RateMin"]), rtol=1e-14)
    np.testing.assert_allclose(p.dumax, _f(f["MV"]["RateMax"]), rtol=1e-14)
    np.testing.assert_allclose(p.ymin, _f(f["OV"]["Min"]), rtol=1e-14)
    np.testing.assert_allclose(p.ymax, _f(f["OV"]["Max"]), rtol=1e-14)
    np.testing.assert_allclose(p.su, _f(f["MV"]["ScaleFactor"]), rtol=1e-14)
    np.testing.assert_allclose(p.sy, _f(f["OV"]["ScaleFactor"]), rtol=1e-14)
    np.testing.assert_allclose(p.ecr_min, _f(f["OV"]["MinECR"]))
    np.testing.assert_allclose(p.ecr_max, _f(f["OV"]["MaxECR"]))
    if "DV" in f:  # DV ScaleFactor = 0.5 / Rv (MPCTuning.m:195-199); v itself is mdv / Rv
        np.testing.assert_allclose(p.v[-1], _f(f["DV"]["ScaleFactor"]), rtol=1e-14)


def test_dmin_and_validity():
    p = shell3x3(2)
    assert list(p.dmin) == [6, 3, 0]          # SURVEY.md §8d (descompMPC.m:35-38 on the scaled plant)
    assert p.valid(24, 6) and p.valid(7, 2)
    assert not p.valid(6, 2)                  # N <= dmin
    assert not p.valid(8, 8) and not p.valid(8, 1)
    assert list(shell7x5().dmin) == [6, 3, 0, 0, 0, 0, 0]


def test_simulate_matches_scipy_dlsim():
    """Independent check of the channel recursion (closedloop_toolbox.m:100 `lsim`)."""
    from scipy import signal
    p = shell3x3(2)
    rng = np.random.default_rng(1)
    w = rng.standard_normal((120, 3))
    y = simulate(p.plant, w)
    yy = np.zeros_like(y)
    for i in range(3):
        for j in range(3):
            d = int(p.plant.d[i, j])
            num = np.concatenate([np.zeros(d), [p.plant.b0[i, j], p.plant.b1[i, j]]])
            den = np.concatenate([[1.0, -p.plant.a[i, j]], np.zeros(d)])
            _, yo = signal.dlsim((num, den, 1.0), w[:, j])
            yy[:, i] += yo[:, 0]
    np.testing.assert_allclose(y, yy, atol=1e-12)


def test_zoh_matches_continuous_step_response():
    """c2d(...,'zoh') must reproduce the continuous FOPDT step response at the sample instants."""
    K, tau, theta, Ts = 4.05, 50.0, 27.0, 4.0
    ch = c2d_fopdt([[K]], [[tau]], [[theta]], Ts)
    y = simulate(ch, np.ones((60, 1)))[:, 0]
    t = np.arange(60) * Ts
    yc = np.where(t >= theta, K * (1 - np.exp(-(t - theta) / tau)), 0.0)
    np.testing.assert_allclose(y, yc, atol=1e-13)


def test_yref_and_setpoint_shapes():
    p = shell3x3(2)
    assert p.r.shape == (500, 3) and p.yref.shape == (3, 500) and p.v.shape == (500, 0)
    # Xsp(1,10:80)=0.2 ... with later assignments overriding (Shell3x3.m:89-92), then * L
    np.testing.assert_allclose(p.r[8], 0.0)
    np.testing.assert_allclose(p.r[9] / p.L, [0.2, 0.2, 0.2])
    np.testing.assert_allclose(p.r[79] / p.L, [0.0, 0.4, 0.1])
    np.testing.assert_allclose(p.r[199] / p.L, [0.1, 0.3, 0.0])
    np.testing.assert_allclose(p.r[399] / p.L, [0.0, 0.0, 0.0])
    w = woodberry()
    assert w.r.shape == (400, 2) and w.v.shape == (400, 1)
    assert w.v[298, 0] == 0.0 and w.v[299, 0] == -0.25
    s = shell7x5()
    assert s.band_mask.all() and s.rho_ecr == 1e4 and not s.square


def test_cond_min_reaches_survey_condition_number():
    """cond* = 19.9495 for the Shell 3x3 gain; the minimiser is a scale family."""
    K = np.array([[4.05, 1.77, 5.88], [5.39, 5.72, 6.9], [4.38, 4.42, 7.2]])
    L, R, c = cond_min(K)
    assert abs(c - 19.9495) < 2e-3
    from mpcgpu.problems import SHELL3X3_L, SHELL3X3_R
    assert abs(np.linalg.cond(np.diag(SHELL3X3_L) @ K @ np.diag(SHELL3X3_R)) - 19.9495) < 2e-3


def test_general_order_c2d_with_fractional_delay():
    """c2d(tf(num,den,'iodelay',theta),Ts,'zoh') for any order.  First order = the closed form the kernels
    use (pinned to the reference's saved objects above); no delay = scipy's ZOH; fractional delay = a fine-grid continuous
    simulation of the delayed transfer function."""
    from scipy.signal import cont2discrete, lsim
    from mpcgpu.plant import c2d_tf, c2d_fopdt, descomp_mpc
    for K, tau, theta, Ts in ((4.05, 50.0, 27.0, 4.0), (12.8, 16.7, 1.0, 1.0), (3.8, 14.9, 8.1, 1.0), (7.2, 19.0, 0.0, 4.0)):
        bz, az, d = c2d_tf([K], [tau, 1.0], theta, Ts)
        ch = c2d_fopdt([[K]], [[tau]], [[theta]], Ts)
        assert d == ch.d[0, 0]
        np.testing.assert_allclose(bz, [ch.b0[0, 0], ch.b1[0, 0]], atol=1e-14)
        np.testing.assert_allclose(az, [1.0, -ch.a[0, 0]], atol=1e-14)
    num, den, Ts = [2.0, 1.0], [10.0, 7.0, 1.0], 0.5
    bz, az, d = c2d_tf(num, den, 0.0, Ts)
    nd, dd, _ = cont2discrete((num, den), Ts, method="zoh")
    np.testing.assert_allclose(bz, nd.ravel(), atol=1e-13); np.testing.assert_allclose(az, dd, atol=1e-13)
    theta = 1.3
    bz, az, d = c2d_tf(num, den, theta, Ts)
    assert d == 3 and len(bz) == 3
    nit = 60
    u = np.zeros(nit); u[2:] = 1.0; u[20:] = -0.5
    y = np.zeros(nit)
    for k in range(nit):
        y[k] = -sum(az[i] * y[k - i] for i in range(1, len(az)) if k - i >= 0) + sum(bz[i] * u[k - d - i] for i in range(len(bz)) if k - d - i >= 0)
    dt = 0.001
    tt = np.arange(0, nit * Ts, dt)
    uc = np.array([u[int(np.floor((x - theta) / Ts + 1e-9))] if x >= theta else 0.0 for x in tt])
    _, yc, _ = lsim((num, den), uc, tt)
    assert np.abs(y - yc[::int(round(Ts / dt))]).max() < 5e-4      # (the error is the reference simulation's grid)
    B, A, dd2 = descomp_mpc(bz, az, d)                              # descompMPC.m:35-38: leading coefficient non-zero -> d-1, zero prepended
    assert dd2 == d - 1 and B[0] == 0.0 and len(B) == len(bz) + 1
