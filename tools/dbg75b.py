import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mpcgpu
from oracle import oracle as orc
p = mpcgpu.shell7x5()
ev = mpcgpu.Evaluator(p, device=0)
op = orc.OracleProblem(p)
N, Nu, dl, lm = mpcgpu.synthetic_population(p, 48, seed=5, wlo=1e-2)
sel = [int(x) for x in sys.argv[1:]] or [42]
out = ev.eval_batch(N[sel], Nu[sel], dl[sel], lm[sel], mode="gam", traj=True)
for n, c in enumerate(sel):
    y, u, ys, uo, rc, stats = orc.closedloop(op, N[c], Nu[c], dl[c], lm[c])
    du = np.abs(out["u"][n] - u).max(axis=0)
    k = int(np.argmax(du > 1e-6)) if (du > 1e-6).any() else -1
    g0 = ((y - p.yref) ** 2).sum(axis=1)
    print("cand", c, "N", N[c], "Nu", Nu[c], "status", out["status"][n], "first |du|>1e-6 at k", k, "du around", du[max(k - 1, 0):k + 6], "max du", du.max(),
          "rel cost err per output", np.abs(out["cost"][n] - g0) / g0)
