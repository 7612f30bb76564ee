"""Shell7x5 full-range population: kernel time of k_soft (A/B of builds)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np, mpcgpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
p = mpcgpu.shell7x5(); ev = mpcgpu.Evaluator(p, device=0)
P = mpcgpu.synthetic_population(p, n, seed=0)
for _ in range(2): out = ev.eval_batch(*P, mode="gam")
c = ev.counters()
print(os.environ.get("MPCGPU_LIB", "default"), "n", n, "sim ms %.1f" % c["last_sim_ms"], "failed", int((out["status"] != 0).sum()), "cost checksum %.12e" % np.nansum(out["cost"]))
