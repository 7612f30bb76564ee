"""Evaluate single candidates of the seed-0 Shell3x3 population alone on the GPU (latency experiments)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import mpcgpu
npop = int(sys.argv[1]); idx = [int(a) for a in sys.argv[2:]]
p = mpcgpu.shell3x3(2)
ev = mpcgpu.Evaluator(p, device=0)
N, Nu, dl, lm = mpcgpu.synthetic_population(p, npop, seed=0)
for c in idx:
    for _ in range(2):
        out = ev.eval_batch(N[c:c + 1], Nu[c:c + 1], dl[c:c + 1], lm[c:c + 1], mode="gam")
    print("cand", c, "N", N[c], "Nu", Nu[c], "sim ms", ev.counters()["last_sim_ms"], flush=True)
