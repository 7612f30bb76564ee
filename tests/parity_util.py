"""Shared helpers for the parity tests (emulation on CPU, CUDA on the GPU box)."""
import numpy as np

from oracle import oracle as orc

EPS = 2.220446049250313e-16
TOL_COST = 1e-6      # BASELINE.json north_star: 1e-6 relative on cost
TOL_TRAJ = 1e-5      # ... 1e-5 on trajectories
COND_FACTOR = 20.0   # conditioning-limited bound: 20 * cond(H) * eps (measured worst case 1.4)


def hessian_cond(op, prob, N, Nu, delta, lam):
    """cond_2 of the candidate's QP Hessian, from the oracle's own H."""
    hl = int(prob.plant.d.max()) + 2
    nw = prob.nu + prob.nd
    out = np.zeros(len(N))
    for c in range(len(N)):
        z, H, f, G, yf, it, rc = orc.single_qp(op, int(N[c]), int(Nu[c]), delta[c], lam[c], np.zeros((prob.ny, nw)),
                                               np.zeros((nw, hl)), np.zeros(nw), np.zeros(prob.ny), np.zeros(prob.nu))
        out[c] = np.linalg.cond(H)
    return out


def cost_tolerance(cond):
    """1e-6 wherever fp64 resolves it; beyond that the QP data itself (H to 1 ulp) moves the answer by
    ~cond(H)*eps, for the oracle as much as for the GPU (tests/test_oracle.py::test_oracle_fp_noise)."""
    return np.maximum(TOL_COST, COND_FACTOR * cond * EPS)


def check_cost(cost, cost_ref, cond, what=""):
    cost = np.asarray(cost); cost_ref = np.asarray(cost_ref)
    rel = np.abs(cost - cost_ref) / np.maximum(np.abs(cost_ref), 1e-300)
    if rel.ndim == 2:
        rel = rel.max(axis=1)
    tol = cost_tolerance(cond)
    bad = np.where(rel > tol)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} candidates out of tolerance, worst rel {rel.max():.3e} (cond {cond[rel.argmax()]:.2e})"
    return rel


def vns_well_posed(prob, op_factory, N, Nu, delta, lam, thresh=1e-7):
    """The Jnu term of the VNS objective divides |uopt(0)| by |diff(uopt)| and zeroes inf/nan
    (VNS2.m:183-191): a move that is 0 in exact arithmetic but 1e-18 in floating point turns a 0 term
    into 1e+30.  Candidates whose open-loop optimum has such a near-zero (but non-zero) difference are
    ill-posed in the reference itself and are excluded from the VNS parity comparison (reported)."""
    ok = np.ones(len(N), dtype=bool)
    runs = prob.ny if prob.ny == prob.nu else 1
    for c in range(len(N)):
        for i in range(runs):
            r = prob.vns_setpoint()
            if runs > 1:
                r = r * np.eye(prob.ny)[i]
            opi = op_factory(r)
            y, u, ys, uo, rc, _ = orc.closedloop(opi, int(N[c]), int(Nu[c]), delta[c], lam[c])
            rows = [i] if runs > 1 else range(prob.nu)
            for j in rows:
                df = np.abs(np.diff(uo[j][: int(Nu[c]) + 1]))
                u0 = abs(uo[j][0])
                nz = df[df > 0]
                if len(nz) and u0 > 0 and (nz < thresh * u0).any():
                    ok[c] = False
    return ok
