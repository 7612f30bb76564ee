"""CPU-side check of the *kernel source*: csrc/mpc_core.cuh (builder) and csrc/mpc_sim.cuh (closed-loop
warp kernel) compiled for the host, the latter executed thread-per-lane through
tests/host_emulation/simt.h, against the oracle.  Debugging aid for the algorithm, not a product path;
the CUDA build of the same source is checked on the GPU box by tests/test_gpu_parity.py.
The emulation pays a barrier per warp intrinsic, so these cases are small (short nit, few candidates)."""
import copy

import numpy as np
import pytest

import emu
from mpcgpu import shell3x3, woodberry, synthetic_population
from oracle import oracle as orc
from parity_util import check_cost, oracle_sensitivity, TOL_TRAJ


def short(prob, nit):
    p = copy.copy(prob)
    p.nit = nit
    p.r = prob.r[:nit].copy(); p.v = prob.v[:nit].copy(); p.yref = prob.yref[:, :nit].copy()
    return p


@pytest.mark.parametrize("case,nit", [("shell3x3", 130), ("woodberry", 330)])
def test_gam_cost_and_trajectories(case, nit):
    p = short({"shell3x3": lambda: shell3x3(2), "woodberry": woodberry}[case](), nit)
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, 7, seed=2, wlo=1e-3, whi=3.0)
    N[0], Nu[0] = 127, 15      # P = 16, two row slots when nu = 3
    N[1], Nu[1] = 40, 3        # P = 4
    lm[0] *= 1e-2              # aggressive: long saturation, more than 16 active constraints -> global spill path
    g0, st0, stats = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    g1, st1, cnt, tr = emu.eval_batch(p, N, Nu, dl, lm, "gam", traj=True)
    assert (st0 == 0).all() and (st1 == 0).all(), (st0, st1)
    sens = oracle_sensitivity(op, N, Nu, dl, lm, "gam", g0)
    check_cost(g1, g0, sens, case, min_strict=0.5)
    assert int(cnt[0]) > 0 and int(cnt[1]) > 0   # the constrained path was exercised
    for c in np.where(sens < 1e-8)[0]:
        y, u, ys, uo, rc, _ = orc.closedloop(op, N[c], Nu[c], dl[c], lm[c])
        for a, b in zip(tr, (y, u, ys, uo)):
            assert np.abs(a[c] - b).max() < TOL_TRAJ


def test_vns_cost_parity():
    p = short(shell3x3(2), 90)
    op = orc.OracleProblem(p)
    N = np.array([24, 12, 60], dtype=np.int32); Nu = np.array([6, 4, 10], dtype=np.int32)
    dl = np.array([[0.5, 0.8, 0.3], [0.05, 0.04, 0.01], [1.0, 1.0, 1.0]])
    lm = np.array([[0.2, 0.1, 0.3], [0.065, 0.0017, 0.077], [0.1, 0.1, 0.1]])
    F0, st0, _ = orc.eval_batch(op, N, Nu, dl, lm, "vns")
    F1, st1, _, _ = emu.eval_batch(p, N, Nu, dl, lm, "vns")
    assert (st0 == 0).all() and (st1 == 0).all()
    ok = F0 < 1e9   # the Jnu division (VNS2.m:183-191) is ill-posed for some candidates, see parity_util.vns_well_posed
    assert ok.sum() >= 2
    rel = np.abs(F1 - F0)[ok] / np.abs(F0[ok])
    assert rel.max() < 1e-6, (F0, F1)


def test_invalid_horizons_are_flagged():
    p = short(shell3x3(2), 20)
    N = np.array([5, 300, 10, 20], dtype=np.int32); Nu = np.array([5, 3, 0, 4], dtype=np.int32)
    g, st, _, _ = emu.eval_batch(p, N, Nu, np.ones((4, 3)), np.ones((4, 3)), "gam")
    assert list(st) == [4, 4, 4, 0]


def test_soft_output_constraints_block_kernel():
    """csrc/mpc_soft.cuh (one CTA per run, square-root dual active set) on a shortened Shell7x5: the measured
    disturbance enters at k = 19 and pushes the band-controlled outputs into their soft limits."""
    from mpcgpu import shell7x5
    p = short(shell7x5(), 27)   # 128 host threads and a barrier per phase: keep it short
    op = orc.OracleProblem(p)
    N = np.array([12], dtype=np.int32); Nu = np.array([3], dtype=np.int32)
    dl = np.zeros((1, 7)); lm = np.array([[0.06, 0.02, 1.6]])
    g0, st0, stats = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    g1, st1, cnt, tr = emu.eval_batch(p, N, Nu, dl, lm, "gam", traj=True)
    assert (st0 == 0).all() and (st1 == 0).all(), (st0, st1)
    assert int(cnt[0]) > 0 and int(cnt[1]) > 0
    rel = np.abs(g1 - g0) / np.abs(g0)
    assert rel.max() < 1e-6, rel
    for c in range(1):
        y, u, ys, uo, rc, _ = orc.closedloop(op, N[c], Nu[c], dl[c], lm[c])
        for a, b in zip(tr, (y, u, ys, uo)):
            assert np.abs(a[c] - b).max() < TOL_TRAJ


def test_two_phase_limit_cycle_reentry():
    """A bang-bang tuning (N = 8, Nu = 6, candidate 4169 of the 32768-candidate seeded population) limit-cycles with
    period 2: the closed-loop kernel parks the factor of the other phase and re-enters it (SimWarp::exchange).
    Same optimum as the oracle, and far fewer active-set iterations than with the mode disabled (knob 64)."""
    p = short(shell3x3(2), 110)
    op = orc.OracleProblem(p)
    N = np.array([8], dtype=np.int32); Nu = np.array([6], dtype=np.int32)
    dl = np.array([[5.34255979, 0.03985835, 0.00996119]]); lm = np.array([[0.01538541, 0.00029207, 0.00017591]])
    y, u, ys, uo, rc, _ = orc.closedloop(op, 8, 6, dl[0], lm[0])
    assert rc == 0
    g1, st1, cnt, tr = emu.eval_batch(p, N, Nu, dl, lm, "gam", traj=True)
    assert st1[0] == 0
    for a, b in zip(tr[:2], (y, u)):
        assert np.abs(a[0] - b).max() < TOL_TRAJ
    emu.lib().emu_set_knob(64)
    try:
        g2, st2, cnt2, _ = emu.eval_batch(p, N, Nu, dl, lm, "gam")
    finally:
        emu.lib().emu_set_knob(0)
    assert np.allclose(g1, g2, rtol=1e-9)
    assert int(cnt2[1]) > 400 and int(cnt[1]) < int(cnt2[1]) // 3, (cnt, cnt2)


def test_validation_run_against_a_mismatched_plant():
    """The validation image of the block-per-run kernel (soft_run<NU,16,true>: real plant + state estimator, mpcgpu_set_mismatch)
    executed on the host against the oracle's estimator loop, on Wood-Berry with gain and dead-time errors (WoodBerry.m:33-47).
    With the nominal plant the run must reproduce the ordinary closed loop (zero innovation)."""
    from mpcgpu import estimator as est
    p = short(woodberry(), 120)
    op = orc.OracleProblem(p)
    plant = est.woodberry_real_plant()
    hl = est.history_length(p, plant)
    M = est.default_estimator_gain(p, hl)
    cand = (14, 3, np.array([0.5, 0.5]), np.array([0.3, 0.3]))
    y0, u0, _, _, rc0, _ = orc.closedloop(op, *cand, open_loop=False)
    ya, ua, ca, rca = emu.eval_est(p, p.plant, M, hl, *cand)
    assert rc0 == 0 and rca == 0
    assert np.abs(ya - y0).max() < 1e-9 and np.abs(ua - u0).max() < 1e-9
    y1, u1, rc1, _ = orc.closedloop_est(op, plant, M, *cand)
    yb, ub, cb, rcb = emu.eval_est(p, plant, M, hl, *cand)
    assert rc1 == 0 and rcb == 0
    assert np.abs(yb - y1).max() < TOL_TRAJ and np.abs(ub - u1).max() < TOL_TRAJ, (np.abs(yb - y1).max(), np.abs(ub - u1).max())
    np.testing.assert_allclose(cb, ((y1 - op.yref) ** 2).sum(axis=1), rtol=1e-6)
    assert np.abs(y1 - y0).max() > 1e-3


def test_nmpc_group_kernel_source_on_the_host():
    """csrc/mpc_nmpc_group.cuh (the NMPC kernel: group-wide sensitivities, packed Hessian, cooperative box-QP, model-based
    full-step test) executed by a warp of host threads (G = 32) against the CPU port of the same restated algorithm
    (oracle/nmpc_port.cpp, built from mpc_nmpc_core.h: two rollouts per SQP iteration, sequential line search) and against the
    committed output of the scipy oracle.  Same iterates, so the costs agree to rounding."""
    import os
    from mpcgpu.nmpc import vandevusse
    from oracle import nmpc_port
    p = vandevusse()
    for cand in ((6, 2, [1.0, 1.0], [0.1, 0.1]), (10, 3, [0.5, 2.0], [0.05, 0.3])):
        c, st, _, cnt = emu.nmpc_eval(p, *cand, mode="gam")
        c0, s0 = nmpc_port.eval_batch(p, [cand[0]], [cand[1]], [cand[2]], [cand[3]], "gam", 1)
        assert st == 0 and s0[0] == 0 and int(cnt[0]) == p.nit - 1 and int(cnt[1]) >= p.nit - 1
        np.testing.assert_allclose(c, c0[0], rtol=1e-9)
    f, st, _, _ = emu.nmpc_eval(p, 6, 2, [1.0, 1.0], [0.1, 0.1], mode="vns")
    f0, _ = nmpc_port.eval_batch(p, [6], [2], [[1.0, 1.0]], [[0.1, 0.1]], "vns", 1)
    np.testing.assert_allclose(f, f0, rtol=1e-6)      # (the Jnu term amplifies the last digits of the open-loop plan)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden_nmpc.npz"))
    k = int(np.argmin(gold["N"] * gold["Nu"]))
    c, st, (y, u), _ = emu.nmpc_eval(p, int(gold["N"][k]), int(gold["Nu"][k]), gold["delta"][k], gold["lam"][k], mode="gam", traj=True)
    assert st == 0
    assert (np.abs(c - gold["gam"][k]) / np.abs(gold["gam"][k])).max() < 1e-5
    assert (np.abs(y - gold["y"][k]) / p.sy[:, None]).max() < 1e-5 and (np.abs(u - gold["u"][k]) / p.su[:, None]).max() < 1e-5


@pytest.mark.parametrize("G", [8, 16])
def test_nmpc_groups_of_a_warp_may_diverge(G):
    """Several closed-loop runs share a warp in the product (G = 16: two, G = 8: four): different horizons and SQP iteration
    counts make the groups diverge, every exchange is group-wide (__shfl_sync / __syncwarp with the group's mask).  The host
    emulation gives every lane mask its own barrier, so a missing or mis-masked barrier shows as a deadlock or a wrong
    result here.  Same costs as one run per warp and as the CPU port."""
    from mpcgpu.nmpc import vandevusse
    from oracle import nmpc_port
    import os
    p = vandevusse()
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden_nmpc.npz"))
    N, Nu, dl, lm = gold["N"][:4], gold["Nu"][:4], gold["delta"][:4], gold["lam"][:4]      # horizons 3..10, Nu 2..4 in one warp
    c, st = emu.nmpc_eval_group(p, N, Nu, dl, lm, G=G)
    c1, st1 = emu.nmpc_eval_group(p, N, Nu, dl, lm, G=32)
    c0, s0 = nmpc_port.eval_batch(p, N, Nu, dl, lm, "gam", 1)
    assert (st == 0).all() and (st1 == 0).all() and (s0 == 0).all()
    assert np.array_equal(c, c1)                                            # the group width does not change the arithmetic
    assert (np.abs(c - c0) / np.abs(c0)).max() < 1e-6
    assert (np.abs(c - gold["gam"][:4]) / np.abs(gold["gam"][:4])).max() < 1e-5


@pytest.mark.parametrize("deltak,deltaL", [(0.0, 0.0), (0.1, 1.0)])
def test_dtc_kernels_on_the_host(deltak, deltaL):
    """csrc/mpc_dtc_kernel.cuh (k_dtc_filter: robustness filter per candidate; k_dtc: gain, predictor and loop, one warp per
    candidate, four candidates per CTA) executed by host threads against oracle/dtc_gpc_oracle.py, the line-by-line numpy
    restatement of DTC_GPC_WW.m -- nominal plant and the reference's mismatch knobs (unstable tunings of the sweep are
    compared relative to the size of their response)."""
    from mpcgpu.dtcgpc import woodberry_dtc, synthetic_dtc_population
    from oracle import dtc_gpc_oracle as dorc
    prob = woodberry_dtc(deltak=deltak, deltaL=deltaL)
    p, m, dl, lm, alfa, raio = synthetic_dtc_population(prob, 9, seed=3)      # 9: a partly filled CTA too
    raio[0] = 0.999      # no slow pole: Fr = 1
    raio[1] = 0.5        # every pole counts as slow
    p[2] = (3, 3); m[2] = (3, 3); dl[2] = 1.0; lm[2] = 1.0; alfa[2] = 0.7; raio[2] = 0.8     # the script's own tuning (DTC_GPC_WW.m:56-76)
    ise, st, y, u = emu.dtc_eval(prob, p, m, dl, lm, alfa, raio)
    assert (st == 0).all(), st
    for c in range(len(p)):
        fr = dorc.mimofilter_Fr(prob.pnz, alfa[c], raio[c])
        y0, u0 = dorc.dtc_gpc_closed_loop(prob, p[c], m[c], dl[c], lm[c], fr)
        sc = max(1.0, np.abs(y0).max(), np.abs(u0).max())
        assert np.abs(y[c] - y0).max() < 1e-8 * sc and np.abs(u[c] - u0).max() < 1e-8 * sc, (c, np.abs(y[c] - y0).max(), sc)
        if sc < 1e3:
            np.testing.assert_allclose(ise[c], ((y0 - prob.r) ** 2).sum(axis=1), rtol=1e-8)


@pytest.mark.parametrize("case", ["shell3x3", "shell7x5"])
def test_builder_with_real_threads(case):
    """csrc/mpc_core.cuh (k_build's body: Hessian from the prefix-Gram tables, Cholesky, [M | W] = H^-1 [-K | I]) executed by 128
    host threads with real block barriers gives, bit for bit, what the serial host build of the same source gives."""
    import mpcgpu
    p = {"shell3x3": lambda: shell3x3(2), "shell7x5": mpcgpu.shell7x5}[case]()
    for N, Nu in ((40, 6), (127, 15), (9, 2)):
        dl = np.zeros(p.ny) if case == "shell7x5" else np.linspace(0.2, 1.5, p.ny)
        lm = np.linspace(0.05, 0.7, p.nu)
        M0, W0, rc0 = emu.build_matrices(p, N, Nu, 16, dl, lm, threads=0)
        M1, W1, rc1 = emu.build_matrices(p, N, Nu, 16, dl, lm, threads=128)
        assert rc0 == 0 and rc1 == 0
        assert np.array_equal(M0, M1) and np.array_equal(W0, W1)
        assert np.abs(W0).max() > 0 and (case == "shell7x5" or np.abs(M0).max() > 0)
