"""Unconstrained-only Shell3x3 population (large lambda): only the speculative loop of the closed-loop kernel runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np, mpcgpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
p = mpcgpu.shell3x3(2)
ev = mpcgpu.Evaluator(p, device=0)
N, Nu, dl, lm = mpcgpu.synthetic_population(p, n, seed=0, wlo=1e-4, whi=0.3)
lm = np.exp(np.random.default_rng(5).uniform(np.log(3.0), np.log(10.0), size=lm.shape))
for _ in range(2):
    out = ev.eval_batch(N, Nu, dl, lm, mode="gam")
c = ev.counters()
print("population", n, "sim ms", c["last_sim_ms"], "constrained", c["qp_constrained"])
