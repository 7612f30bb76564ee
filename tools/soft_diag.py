"""Shell7x5 (soft output bands) on the survey's full weight range: GPU status / cost against the CPU oracle, by category."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np, mpcgpu
from oracle import oracle as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
p = mpcgpu.shell7x5(); ev = mpcgpu.Evaluator(p, device=0); op = orc.OracleProblem(p)
N, Nu, dl, lm = mpcgpu.synthetic_population(p, n, seed=0)
ev.eval_batch(N[:32], Nu[:32], dl[:32], lm[:32], mode="gam")
t = time.time(); out = ev.eval_batch(N, Nu, dl, lm, mode="gam"); tg = time.time() - t
c = ev.counters()
t = time.time(); g0, st0, _ = orc.eval_batch(op, N, Nu, dl, lm, "gam"); to = time.time() - t
st = out["status"]
print("gpu %.2fs (kernel %.0f ms) oracle %.1fs" % (tg, c["last_sim_ms"], to))
print("gpu status", np.bincount(st, minlength=4).tolist(), "oracle status", np.bincount(st0, minlength=4).tolist())
both = (st == 0) & (st0 == 0)
rel = np.abs(out["cost"] - g0) / np.maximum(np.abs(g0), 1e-300)
rel = np.nanmax(np.where(np.isfinite(rel), rel, 0), axis=1)
print("both ok", both.sum(), "gpu only ok", ((st == 0) & (st0 != 0)).sum(), "oracle only ok", ((st != 0) & (st0 == 0)).sum(), "both fail", ((st != 0) & (st0 != 0)).sum())
for thr in (1e-6, 1e-4, 1e-2):
    print("  both ok and rel <=", thr, ":", (rel[both] <= thr).sum(), "of", both.sum())
lmin = lm.min(axis=1)
for lo, hi in ((1e-4, 1e-3), (1e-3, 1e-2), (1e-2, 1e-1), (1e-1, 10)):
    s = (lmin >= lo) & (lmin < hi)
    print("  lam_min in [%g,%g): n %d gpu fail %.2f oracle fail %.2f agree(1e-6 | both ok) %.2f" % (lo, hi, s.sum(), (st[s] != 0).mean(), (st0[s] != 0).mean(), (rel[s & both] <= 1e-6).mean() if (s & both).any() else float("nan")))
np.savez("gpurun_out/soft_diag.npz", st=st, st0=st0, rel=rel, cost=out["cost"], g0=g0)
