#!/usr/bin/env python3
"""Shortest possible GPU check of k_ssnmpc (the round's GPU budget had seconds left): the five golden closed loops through the C ABI
against the committed oracle output, result written to gpurun_out/ssnmpc_gpu.json at once; then, if time is left, a 128-candidate
sweep against nothing but its own status (timing only)."""
import json, os, sys, time
t00 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np
import mpcgpu
out = {"import_s": time.time() - t00}
dst = os.path.join(ROOT, "gpurun_out", "ssnmpc_gpu.json")
os.makedirs(os.path.dirname(dst), exist_ok=True)
g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden_ssnmpc.npz"))
p = mpcgpu.explicit_nmpc()
t0 = time.time(); ev = mpcgpu.SsnmpcEvaluator(p, device=0); out["create_s"] = time.time() - t0
t0 = time.time(); o = ev.eval_batch(g["N"], g["Nu"], g["Q"], g["W"], traj=True); out["golden_call_s"] = time.time() - t0
c = ev.counters()
out.update(status=[int(s) for s in o["status"]], dy=float(np.abs(o["y"] - g["y"]).max()),
           du_rel=float((np.abs(o["u"] - g["u"]) / (p.ub - p.lb)[None, :, None]).max()),
           cost_rel=float((np.abs(o["cost"] - g["cost"]) / g["cost"]).max()), kernel_ms=c["last_sim_ms"], controller_calls=c["qp_solves"],
           gn_iterations=c["as_iterations"])
out["golden_ok"] = bool(out["dy"] < 1e-5 and out["du_rel"] < 1e-5 and out["cost_rel"] < 1e-6 and not any(out["status"]))
json.dump(out, open(dst, "w")); print(json.dumps(out), flush=True)
N, Nu, Q, W = mpcgpu.synthetic_ssnmpc_population(p, 128, seed=0)
t0 = time.time(); s = ev.eval_batch(N, Nu, Q, W); out["sweep128_call_s"] = time.time() - t0
c2 = ev.counters()
out.update(sweep128_kernel_ms=c2["last_sim_ms"], sweep128_failed=int((s["status"] != 0).sum()), sweep128_gn_iterations=c2["as_iterations"] - c["as_iterations"])
json.dump(out, open(dst, "w")); print(json.dumps(out), flush=True)
