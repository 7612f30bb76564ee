"""The 64 heaviest runs of the seed-0 Shell3x3 population (tools/tail_idx.txt, from per-run cycle counters) as a population of
their own: the command profiled under ncu to attribute the ACTIVE-SET path (the whole population is dominated by the
speculative unconstrained path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np, mpcgpu
p = mpcgpu.shell3x3(2); ev = mpcgpu.Evaluator(p, device=0)
N, Nu, dl, lm = mpcgpu.synthetic_population(p, 4096, seed=0)
idx = np.array([int(a) for a in open(os.path.join(ROOT, "tools", "tail_idx.txt")).read().split()])
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    out = ev.eval_batch(N[idx], Nu[idx], dl[idx], lm[idx], mode="gam")
c = ev.counters()
print("tail population", len(idx), "sim ms", c["last_sim_ms"], "iterations", c["as_iterations"], "constrained", c["qp_constrained"])
