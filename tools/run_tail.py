"""Evaluate only the heaviest candidates of the seed-0 Shell3x3 population (tools/tail_idx.npy, from
tools/diag_runs.py) -- used under ncu to profile the constrained (active-set) path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np
import mpcgpu
p = mpcgpu.shell3x3(2)
ev = mpcgpu.Evaluator(p, device=0)
N, Nu, dl, lm = mpcgpu.synthetic_population(p, 4096, seed=0)
idx = np.load(os.path.join(ROOT, "tools", "tail_idx.npy"))
for _ in range(3):
    out = ev.eval_batch(N[idx], Nu[idx], dl[idx], lm[idx], mode="gam")
c = ev.counters()
print("tail run:", len(idx), "candidates, sim ms", c["last_sim_ms"], "its", c["as_iterations"] // 3, "con", c["qp_constrained"] // 3)
