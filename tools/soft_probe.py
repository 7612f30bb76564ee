"""Shell7x5 (soft output bands): where does a GPU run leave the oracle's trajectory?  For every candidate whose GAM cost differs
from the oracle's by more than 1e-6: the first sample at which u differs, the size of the difference there and at the end.
usage: soft_probe.py <n> <seed> <wlo>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np, mpcgpu
from oracle import oracle as orc, parity
n = int(sys.argv[1]); seed = int(sys.argv[2]); wlo = float(sys.argv[3])
p = mpcgpu.shell7x5(); ev = mpcgpu.Evaluator(p, device=0); op = orc.OracleProblem(p)
N, Nu, dl, lm = mpcgpu.synthetic_population(p, n, seed=seed, wlo=wlo)
out = ev.eval_batch(N, Nu, dl, lm, mode="gam", traj=True)
c = ev.counters()
g0, st0, _ = orc.eval_batch(op, N, Nu, dl, lm, "gam")
sens = parity.sensitivity(op, N, Nu, dl, lm, "gam", g0)
rel = (np.abs(out["cost"] - g0) / np.maximum(np.abs(g0), 1e-300)).max(axis=1)
print(os.environ.get("MPCGPU_LIB", "default"), "kernel ms %.0f" % c["last_sim_ms"], "iterations", c["as_iterations"], "status", np.bincount(out["status"]).tolist(), np.bincount(st0).tolist())
print("  rel<=1e-6: %d of %d; out of tol (max(1e-6,10 sens)): %d; well-posed (sens<1e-7) and rel>1e-6: %d" % ((rel <= 1e-6).sum(), n, (rel > np.maximum(1e-6, 10 * sens)).sum(), ((sens < 1e-7) & (rel > 1e-6)).sum()))
bad = np.where(rel > 1e-6)[0]
bad = bad[np.argsort(-(rel[bad] / np.maximum(1e-6, 10 * sens[bad])))]   # worst offenders (relative to their tolerance) first
for cidx in bad[:int(os.environ.get('PROBE_LIST', 40))]:
    y, u, ys, uo, rc, stt = orc.closedloop(op, int(N[cidx]), int(Nu[cidx]), dl[cidx], lm[cidx])
    du = np.abs(out["u"][cidx] - u).max(axis=0); dy = np.abs(out["y"][cidx] - y).max(axis=0)
    k1 = np.argmax(du > 1e-9) if (du > 1e-9).any() else -1
    print("  cand %3d N %3d Nu %2d lam %s rel %.2e sens %.1e first u-diff>1e-9 at k=%d (du %.2e) max du %.2e max dy %.2e" % (cidx, N[cidx], Nu[cidx], np.array2string(lm[cidx], precision=4), rel[cidx], sens[cidx], k1, du[k1] if k1 >= 0 else 0.0, du.max(), dy.max()))
np.savez("gpurun_out/soft_probe_%s.npz" % os.path.basename(os.environ.get("MPCGPU_LIB", "default")), rel=rel, sens=sens, cost=out["cost"], g0=g0, u=out["u"][bad[:8]], bad=bad)
