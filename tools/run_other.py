"""Evaluate one population of each of the other configurations (Shell7x5 soft constraints, DTC-GPC sweep, Van de
Vusse NMPC): the command profiled under ncu for profiles/r1g_other_kernels.json."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np
import mpcgpu
from mpcgpu.dtcgpc import woodberry_dtc, synthetic_dtc_population, DtcEvaluator
from mpcgpu.nmpc import vandevusse, synthetic_nmpc_population, NmpcEvaluator

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p7 = mpcgpu.shell7x5(); e7 = mpcgpu.Evaluator(p7, device=0)
P7 = mpcgpu.synthetic_population(p7, 2048, seed=0)
pd = woodberry_dtc(); ed = DtcEvaluator(pd, device=0)
Pd = synthetic_dtc_population(pd, 16384, seed=0)
fil = None   # the filters are designed on the device from (alfa, raio)
pn = vandevusse(); en = NmpcEvaluator(pn, device=0)
Pn = synthetic_nmpc_population(pn, 16384, seed=0)
for _ in range(reps):
    t = time.time(); o7 = e7.eval_batch(*P7, mode="gam"); t7 = time.time() - t
    t = time.time(); od = ed.eval_batch(*Pd[:4], alfa=Pd[4], raio=Pd[5]); td = time.time() - t
    t = time.time(); on = en.eval_batch(*Pn, mode="gam"); tn = time.time() - t
c7 = e7.counters()
print("Shell7x5 2048 candidates (lambda in [1e-4,10]): %.1f ms wall, sim %.1f ms -> %.0f cand/s, ok %d, as_iterations %d" % (t7 * 1e3, c7["last_sim_ms"], 2048 / t7, int((o7["status"] == 0).sum()), c7["as_iterations"] // reps))
print("DTC-GPC 16384 candidates: %.1f ms wall -> %.0f cand/s, ok %d" % (td * 1e3, 16384 / td, int((od["status"] == 0).sum())))
print("NMPC 16384 candidates: %.1f ms wall -> %.0f cand/s, ok %d" % (tn * 1e3, 16384 / tn, int((on["status"] == 0).sum())))
