"""Batched optimiser drivers (mpcgpu/tuner.py) against test doubles on the CPU: the search
logic is host code; the GPU run of the same drivers is tests/test_gpu_parity.py::test_batched_tuner_on_shell3x3."""
import numpy as np

from mpcgpu import tuner


def test_bit_encoding_and_legality():
    assert tuner._bits(127, 7).tolist() == [1] * 7 and tuner._value(tuner._bits(24, 7)) == 24
    assert tuner.legal(24, np.array([6, 2, 2]), [6, 3, 0]) and not tuner.legal(6, np.array([6, 2, 2]), [6, 3, 0])
    assert not tuner.legal(24, np.array([6, 1, 2]), [6, 3, 0])        # Nu > 1 (VNS2.m:135)
    assert not tuner.legal(5, np.array([2, 2, 2]), [6, 3, 0])         # N > dmin


def test_neighbourhood_is_the_hamming_sphere():
    N, Nu = 24, np.array([6, 2, 2])
    cn, cnu = tuner.neighbourhood(N, Nu, 7, 4, 1, [6, 3, 0])
    # 7 single-bit flips of N and 4 of each input's Nu, minus the illegal ones
    assert len(cn) <= 7 + 3 * 4 and len(cn) > 8
    for n2, nu2 in zip(cn, cnu):
        dist = bin(n2 ^ N).count("1") + sum(bin(int(a) ^ int(b)).count("1") for a, b in zip(nu2, Nu))
        assert dist == 1 and tuner.legal(int(n2), nu2, [6, 3, 0])
    cn3, cnu3 = tuner.neighbourhood(N, Nu, 7, 4, 3, [6, 3, 0])
    assert all(bin(a ^ N).count("1") + sum(bin(int(x) ^ int(y)).count("1") for x, y in zip(b, Nu)) == 3 for a, b in zip(cn3, cnu3))


def test_vns_search_finds_the_optimum_of_a_separable_objective():
    target_N, target_Nu = 37, 5
    calls = []

    def F(N, Nu):
        calls.append(len(N))
        return (np.asarray(N) - target_N) ** 2 + 3.0 * (np.asarray(Nu).max(axis=1) - target_Nu) ** 2 + np.asarray(N)

    N, Nu, Fv, evals = tuner.vns_search(F, 127, [2, 2, 2], 7, 4, [6, 3, 0])
    assert Nu.max() == target_Nu and abs(N - target_N) <= 1          # the +N term of VNS2.m:195 pulls N down by < 1
    assert evals == sum(calls) and max(calls) > 10                     # whole neighbourhoods went out as single batches


def test_goal_attain_reduces_the_attainment_factor():
    opt = np.array([0.05, 0.4, 0.01, 0.2])

    def g(X):
        X = np.atleast_2d(X)
        L = np.log(X / opt)
        return np.stack([L[:, 0] ** 2 + L[:, 2] ** 2, L[:, 1] ** 2 + 0.5 * L[:, 3] ** 2], axis=1) + 1e-3

    x, gam, gx, evals = tuner.goal_attain(g, np.ones(4), w=[0.5, 0.5], pop=128, iters=25)
    g0 = tuner.attainment(g(np.ones(4)), 1e-3, [0.5, 0.5])[0]
    assert gam < 1e-2 * g0 and np.abs(np.log(x / opt)).max() < 0.2 and evals > 128


def test_mpc_tfob_alternates_until_the_weights_stop_improving():
    def gam_at(N, Nu):
        return lambda X: np.stack([np.log(np.atleast_2d(X)[:, 0] / 0.1) ** 2 + 0.01 * abs(N - 30),
                                   np.log(np.atleast_2d(X)[:, 1] / 0.3) ** 2 + 1e-3], axis=1)

    def vns_at(delta, lam):
        return lambda N, Nu: (np.asarray(N) - 30.0) ** 2 + (np.asarray(Nu).max(axis=1) - 4.0) ** 2

    out = tuner.mpc_tfob(gam_at, vns_at, 1, 1, 127, [2], [1.0], [1.0], [0.5, 0.5], 7, 4, [0], pop=64, iters=15)
    assert out["N"] == 30 and out["Nu"].max() == 4 and abs(np.log(out["delta"][0] / 0.1)) < 0.3


def test_tuning_parameters_mat_round_trip(tmp_path, kats):
    """matio: the struct of MPCTuning.m:374-380 written and read back; the values are the reference's own Shell3x3
    result (tests/golden/fixture_kats.json, decoded from Shell3x3_Tuning_25Jul2023_12_06.mat)."""
    from mpcgpu import matio
    k = kats["Shell3x3_Tuning_25Jul2023_12_06.mat"]
    f = str(tmp_path / "Shell3x3_Tuning_test.mat")
    matio.save_tuning(f, k["N"], k["Nu"], k["delta"], k["lambda"], L=k["L"], R=k["R"])
    back = matio.load_tuning(f)
    assert back["N"].tolist() == [k["N"]] and back["Nu"].tolist() == k["Nu"]
    np.testing.assert_array_equal(back["delta"], k["delta"]); np.testing.assert_array_equal(back["lambda"], k["lambda"])
    np.testing.assert_array_equal(back["L"], k["L"]); np.testing.assert_array_equal(back["Ru"], k["R"][:3])


def test_load_tuning_reads_the_reference_files(kats):
    """Only where the reference tree is mounted (the build container); elsewhere the committed KAT JSON stands in."""
    import os
    import pytest
    from mpcgpu import matio
    f = "/root/reference/MPC-Tuning/Shell3x3_Tuning_Caso2.mat"
    if not os.path.exists(f):
        pytest.skip("reference tree not mounted")
    got = matio.load_tuning(f)
    k = kats["Shell3x3_Tuning_Caso2.mat"]
    assert int(got["N"].max()) == k["N"] and got["Nu"].tolist() == k["Nu"]
    np.testing.assert_allclose(got["delta"], k["delta"], rtol=0, atol=0)
    np.testing.assert_allclose(got["L"], k["L"], rtol=0, atol=0)
