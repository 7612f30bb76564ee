"""Evaluate the seed-0 Shell3x3 population (BASELINE.json configs[1]) a few times: the command profiled under
ncu for profiles/ (run it plainly first; numbers printed under a profiler are not bench values)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import mpcgpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "gam"
p = mpcgpu.shell3x3(2)
ev = mpcgpu.Evaluator(p, device=0)
N, Nu, dl, lm = mpcgpu.synthetic_population(p, n, seed=0)
for _ in range(reps):
    out = ev.eval_batch(N, Nu, dl, lm, mode=mode)
c = ev.counters()
print("population", n, mode, "sim ms", c["last_sim_ms"], "build ms", c["last_build_ms"], "failed", int((out["status"] != 0).sum()))
