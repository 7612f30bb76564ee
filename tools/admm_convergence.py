"""ADMM variant of the box/rate-constrained MPC QP (BASELINE.json north_star item 3) on the BENCHMARK's QPs: convergence data.

For the first NCAND candidates of the seed-0 Shell3x3 population bench.py times, the QP of the first closed-loop sample after
the set-point step (state 0, u(-1) = 0, r = the scenario's set-point) is taken from the oracle (H, f, and the exact optimum of
its Goldfarb-Idnani solver with the iteration count).  Candidates whose unconstrained optimum is feasible are skipped.  The
same QP is then solved by OSQP-style ADMM (splitting on A z = [z; cumsum(z)] in [lo, hi], over-relaxation 1.6, penalty
rho = scale * sqrt(tr H / tr H^-1)) and the number of iterations is recorded at which (a) the primal/dual residuals fall
below 1e-9 and (b) the iterate is within 1e-6 (relative, inf-norm) of the exact optimum -- the cost tolerance of the
benchmark needs at least (b).  One ADMM iteration is one R x R matrix-vector product with a precomputed inverse, about 600
cycles for a warp; one active-set iteration of k_sim costs ~3700 + 140 q cycles (tools/diag_runs.py).

Writes profiles/r2_admm_convergence.csv and prints the summary quoted in DESIGN.md.  CPU only (numpy + the oracle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np
from mpcgpu.problems import shell3x3, synthetic_population
from oracle import oracle as orc

NCAND = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SCALES = (0.01, 0.03, 0.1, 0.3, 1.0, 3.0)
ITMAX = 5000
prob = shell3x3(2)
op = orc.OracleProblem(prob)
ny, nu, nd = prob.ny, prob.nu, prob.nd
nw = nu + nd
hl = int(prob.plant.d.max()) + 2
N, Nu, dl, lm = synthetic_population(prob, 4096, seed=0)
k_step = int(np.argmax(np.abs(prob.r).sum(axis=1) > 0))
rk = prob.r[k_step]


def admm(H, f, A, lo, hi, rho, z_star, alpha=1.6):
    n = len(f)
    K = np.linalg.inv(H + rho * A.T @ A)
    v = np.clip(A @ (-np.linalg.solve(H, f)), lo, hi); u = np.zeros(len(lo))
    it_res = it_acc = None
    zs = np.abs(z_star).max()
    for it in range(1, ITMAX + 1):
        x = K @ (rho * A.T @ (v - u) - f)
        Ax = A @ x
        xh = alpha * Ax + (1 - alpha) * v
        vn = np.clip(xh + u, lo, hi)
        u = u + xh - vn
        res = max(np.abs(Ax - vn).max(), rho * np.abs(A.T @ (vn - v)).max())
        v = vn
        if it_res is None and res < 1e-9: it_res = it
        if it_acc is None and np.abs(x - z_star).max() <= 1e-6 * zs: it_acc = it
        if it_res is not None and it_acc is not None: break
    return it_res or ITMAX, it_acc or ITMAX


rows = []
for c in range(NCAND):
    p, m = int(N[c]), int(Nu[c])
    z, H, f, G, yf, it_as, rc = orc.single_qp(op, p, m, dl[c], lm[c], np.zeros((ny, nw)), np.zeros((nw, hl)), np.zeros(nw), rk, np.zeros(nu))
    nz = nu * m
    H = H[:nz, :nz]; f = f[:nz]; z = z[:nz]
    # variable order of the oracle: move c of input j at index c*nu + j (checked against its constraints below)
    T = np.zeros((nz, nz))
    for cc in range(m):
        for c2 in range(cc + 1):
            for j in range(nu):
                T[cc * nu + j, c2 * nu + j] = 1.0
    A = np.vstack([np.eye(nz), T])
    lo = np.concatenate([np.tile(prob.dumin, m), np.tile(prob.umin, m)]); hi = np.concatenate([np.tile(prob.dumax, m), np.tile(prob.umax, m)])
    zu = -np.linalg.solve(H, f)
    viol = max((A @ zu - hi).max(), (lo - A @ zu).max())
    if rc != 0 or viol <= 1e-10:
        continue
    assert (A @ z <= hi + 1e-8).all() and (A @ z >= lo - 1e-8).all(), "variable order"
    ev = np.linalg.eigvalsh(H)
    r0 = np.sqrt(np.trace(H) / np.trace(np.linalg.inv(H)))
    its = [admm(H, f, A, lo, hi, r0 * s, z) for s in SCALES]
    rows.append([c, p, m, ev[-1] / ev[0], it_as] + [i for pair in its for i in pair])
    print(rows[-1], flush=True)

rows = np.array(rows, dtype=float)
hdr = "candidate,N,Nu,cond_H,active_set_iterations," + ",".join(f"admm_it_res1e-9_rho{s}xr0,admm_it_acc1e-6_rho{s}xr0" for s in SCALES)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
np.savetxt(os.path.join(ROOT, "profiles", "r2_admm_convergence.csv"), rows, delimiter=",", header=hdr, comments="", fmt="%.6g")
acc = rows[:, 6::2]; res = rows[:, 5::2]
best = acc.min(axis=1)
print("constrained QPs:", len(rows), "of", NCAND, "candidates; active-set iterations median %.0f mean %.1f max %.0f" % (np.median(rows[:, 4]), rows[:, 4].mean(), rows[:, 4].max()))
for i, s in enumerate(SCALES):
    print("rho = %5.2f r0: iterations to 1e-6 of the optimum: median %4.0f p90 %4.0f, not reached in %d: %.1f %%" % (s, np.median(acc[:, i]), np.percentile(acc[:, i], 90), ITMAX, 100 * (acc[:, i] >= ITMAX).mean()))
print("best penalty PER QP (oracle choice): median %.0f p90 %.0f not reached %.1f %%" % (np.median(best), np.percentile(best, 90), 100 * (best >= ITMAX).mean()))
print("cycles per QP at 600 / ADMM iteration vs (3700 + 140 q) / active-set iteration (q ~ iterations): median ratio %.1f" % np.median(best * 600 / (rows[:, 4] * (3700 + 140 * rows[:, 4]))))
