import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, time
import mpcgpu
from oracle import oracle as orc
p = mpcgpu.shell7x5()
ev = mpcgpu.Evaluator(p, device=0)
op = orc.OracleProblem(p)
for wlo in (1e-4, 1e-2):
    N, Nu, dl, lm = mpcgpu.synthetic_population(p, 48, seed=5, wlo=wlo)
    t = time.time(); g0, st0, stats = orc.eval_batch(op, N, Nu, dl, lm, "gam"); t0 = time.time() - t
    t = time.time(); out = ev.eval_batch(N, Nu, dl, lm, mode="gam", traj=True); t1 = time.time() - t
    ok = (st0 == 0) & (out["status"] == 0)
    rel = np.abs(out["cost"] - g0).max(axis=1) / np.abs(g0).max(axis=1)
    print("wlo", wlo, "oracle status", np.bincount(st0, minlength=3), "gpu status", np.bincount(out["status"], minlength=3), "both ok", ok.sum(),
          "oracle s %.2f gpu s %.2f" % (t0, t1), "oracle stats", stats, "gpu counters", ev.counters()["as_iterations"], ev.counters()["last_sim_ms"])
    print(" rel err (both ok):", np.sort(rel[ok])[-8:])
    print(" status pairs:", list(zip(st0, out["status"]))[:48])
