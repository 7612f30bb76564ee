"""CPU-side check of the *kernel source*: csrc/mpc_core.cuh compiled lane-serialised for the host
(tests/host_emulation) against the oracle.  This is a debugging aid for the algorithm, not a product path;
the CUDA build of the same source is checked on the GPU box by tests/test_gpu_parity.py."""
import numpy as np
import pytest

import emu
from mpcgpu import shell3x3, woodberry, synthetic_population
from oracle import oracle as orc
from parity_util import check_cost, oracle_sensitivity, TOL_TRAJ, vns_well_posed


@pytest.mark.parametrize("case,n", [("shell3x3", 160), ("woodberry", 96)])
def test_gam_cost_parity(case, n):
    p = {"shell3x3": lambda: shell3x3(2), "woodberry": woodberry}[case]()
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, n, seed=2)
    g0, st0, stats = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    g1, st1, cnt, _ = emu.eval_batch(p, N, Nu, dl, lm, "gam")
    assert (st0 == 0).all() and (st1 == 0).all()
    rel, strict = check_cost(g1, g0, oracle_sensitivity(op, N, Nu, dl, lm, "gam", g0), case)
    assert np.median(rel) < 1e-10
    assert int(cnt[0]) > 0 and int(cnt[1]) <= int(stats[1])  # warm start: never more active-set iterations than the cold oracle


@pytest.mark.parametrize("case", ["shell3x3", "woodberry"])
def test_trajectory_parity(case):
    p = {"shell3x3": lambda: shell3x3(2), "woodberry": woodberry}[case]()
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, 24, seed=3, wlo=1e-3, whi=3.0)
    _, st, _, tr = emu.eval_batch(p, N, Nu, dl, lm, "raw", traj=True)
    assert (st == 0).all()
    g0, _, _ = orc.eval_batch(op, N, Nu, dl, lm, "gam")
    sens = oracle_sensitivity(op, N, Nu, dl, lm, "gam", g0)
    assert (sens < 1e-8).sum() >= 16
    for c in np.where(sens < 1e-8)[0]:
        y, u, ys, uo, rc, _ = orc.closedloop(op, N[c], Nu[c], dl[c], lm[c])
        for a, b in zip(tr, (y, u, ys, uo)):
            assert np.abs(a[c] - b).max() < TOL_TRAJ


@pytest.mark.parametrize("case", ["shell3x3", "woodberry"])
def test_vns_cost_parity(case):
    p = {"shell3x3": lambda: shell3x3(2), "woodberry": woodberry}[case]()
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, 24, seed=4, wlo=1e-3, whi=3.0)
    F0, st0, _ = orc.eval_batch(op, N, Nu, dl, lm, "vns")
    F1, st1, _, _ = emu.eval_batch(p, N, Nu, dl, lm, "vns")
    ok = vns_well_posed(p, lambda r: orc.OracleProblem(p, r=r), N, Nu, dl, lm)
    sens = oracle_sensitivity(op, N, Nu, dl, lm, "vns", F0)
    assert ok.sum() >= 12
    check_cost(F1[ok], F0[ok], sens[ok], case + " vns", min_strict=0.5)


def test_invalid_horizons_are_flagged():
    p = shell3x3(2)
    N = np.array([5, 300, 10, 20], dtype=np.int32); Nu = np.array([5, 3, 0, 4], dtype=np.int32)
    dl = np.ones((4, 3)); lm = np.ones((4, 3))
    g, st, _, _ = emu.eval_batch(p, N, Nu, dl, lm, "gam")
    assert list(st) == [4, 4, 4, 0]
