"""Shared helpers for the parity tests (emulation on CPU, CUDA on the GPU box)."""
import numpy as np

from oracle import oracle as orc

EPS = 2.220446049250313e-16
TOL_COST = 1e-6      # BASELINE.json north_star: 1e-6 relative on cost
TOL_TRAJ = 1e-5      # ... 1e-5 on trajectories
SENS_FACTOR = 10.0
PERT = (3e-14, -2e-14)   # relative input perturbation (delta, lambda) used to probe well-posedness


def hessian_cond(op, prob, N, Nu, delta, lam):
    """cond_2 of the candidate's QP Hessian, from the oracle's own H (diagnostics)."""
    hl = int(prob.plant.d.max()) + 2
    nw = prob.nu + prob.nd
    out = np.zeros(len(N))
    for c in range(len(N)):
        z, H, f, G, yf, it, rc = orc.single_qp(op, int(N[c]), int(Nu[c]), delta[c], lam[c], np.zeros((prob.ny, nw)),
                                               np.zeros((nw, hl)), np.zeros(nw), np.zeros(prob.ny), np.zeros(prob.nu))
        out[c] = np.linalg.cond(H)
    return out


def oracle_sensitivity(op, N, Nu, delta, lam, mode, cost_ref):
    """How far the ORACLE's own cost moves when delta, lambda are perturbed by ~100 ulp.  Two effects make
    this large for a few candidates, for any fp64 implementation (DESIGN.md "Tolerance"):
      * cond(H) up to 4e12 in the survey's weight range: H is only known to 1 ulp, the optimum moves by cond*eps;
      * closed loops that are unstable/limit-cycling against the MV limits (e.g. N=7 is shorter than the
        7-sample dead time of output 1): differences grow exponentially over the 500 samples."""
    pert, _, _ = orc.eval_batch(op, N, Nu, np.asarray(delta) * (1 + PERT[0]), np.asarray(lam) * (1 + PERT[1]), mode)
    rel = np.abs(pert - cost_ref) / np.maximum(np.abs(cost_ref), 1e-300)
    return rel.max(axis=1) if rel.ndim == 2 else rel


def check_cost(cost, cost_ref, sens, what="", min_strict=0.75):
    """|cost - oracle| / |oracle| <= max(1e-6, 10 * oracle's own sensitivity).  At least `min_strict` of the
    population must be held to the strict 1e-6."""
    cost = np.asarray(cost); cost_ref = np.asarray(cost_ref)
    rel = np.abs(cost - cost_ref) / np.maximum(np.abs(cost_ref), 1e-300)
    if rel.ndim == 2:
        rel = rel.max(axis=1)
    tol = np.maximum(TOL_COST, SENS_FACTOR * sens)
    bad = np.where(~(rel <= tol))[0]
    assert len(bad) == 0, (f"{what}: {len(bad)} candidates out of tolerance, worst rel {rel[bad].max():.3e} "
                           f"(tol {tol[bad][rel[bad].argmax()]:.2e}, candidate {bad[rel[bad].argmax()]})")
    strict = float((tol <= TOL_COST).mean())
    assert strict >= min_strict, f"{what}: only {strict:.2%} of the candidates are well-posed enough for the 1e-6 bar"
    return rel, strict


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
