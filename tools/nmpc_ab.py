"""Van de Vusse NMPC kernel timing (device time of k_nmpc) for a few population sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np, mpcgpu
p = mpcgpu.vandevusse(); ev = mpcgpu.NmpcEvaluator(p, device=0)
for n in [int(a) for a in sys.argv[1:]] or [2048, 16384]:
    pop = mpcgpu.synthetic_nmpc_population(p, n, seed=0)
    ev.eval_batch(*[a[:64] for a in pop], mode="gam")
    out = ev.eval_batch(*pop, mode="gam")
    c = ev.counters()
    print(os.environ.get("MPCGPU_LIB", "default"), "n", n, "kernel ms", round(c["last_sim_ms"], 1), "cand/s", round(n / c["last_sim_ms"] * 1e3), "status", np.bincount(out["status"]).tolist(), c)
