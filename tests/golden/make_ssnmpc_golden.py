#!/usr/bin/env python3
"""Writes tests/golden/oracle_golden_ssnmpc.npz from the CPU oracle of the single-shooting NMPC (oracle/ssnmpc_oracle.py:
the objective function of `Explicit NMPC/NMPC_Controller.m` restated line by line, minimised by scipy trust-region least
squares).  ORACLE outputs -- the reference holds no output of this demo; candidate 0 is its own setting (main.m:59-62)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
from mpcgpu.ssnmpc import explicit_nmpc  # noqa: E402
from oracle import ssnmpc_oracle as so  # noqa: E402

prob = explicit_nmpc()
N = np.array([5, 7, 10, 3, 6], dtype=np.int32)
Nu = np.array([[2, 2], [2, 1], [4, 2], [1, 1], [3, 1]], dtype=np.int32)   # (7, [1 2]) with these weights is an ill-posed loop: see tests/test_ssnmpc_oracle.py
Q = np.array([[1.0214, 0.9999], [2.0, 0.5], [0.3, 3.0], [1.0, 1.0], [5.0, 0.2]])
W = np.array([[1e-4, 1e-4], [1e-3, 2e-4], [5e-5, 5e-3], [1e-2, 1e-3], [1e-4, 1e-2]])
rng = np.random.Generator(np.random.PCG64(7))
noise = 0.01 * rng.standard_normal((3, prob.nit))            # ClosedLoopNMPC.m:76,89: noise_magnitude * randn


def run(job):
    c, nz = job
    return so.closed_loop_nmpc(prob.x0, prob.x_control, prob.u0, prob.r, N[c], Nu[c], Q[c], W[c], prob.nit, prob.ub, prob.lb,
                               prob.inK, prob.Ts, noise=nz, nsub=prob.nsub)


if __name__ != "__main__":
    raise SystemExit
from multiprocessing import Pool  # noqa: E402
with Pool(6) as pool:                                           # the polished oracle takes ~1 s per controller call
    res = pool.map(run, [(c, None) for c in range(len(N))] + [(0, noise)])
ys = [r_[0] for r_ in res[:-1]]; us = [r_[1] for r_ in res[:-1]]
cost = [so.sweep_cost(y, prob.r, prob.inK) for y in ys]
yn, un = res[-1]
print(np.array(cost))
dst = os.path.join(ROOT, "tests", "golden", "oracle_golden_ssnmpc.npz")
np.savez_compressed(dst, N=N, Nu=Nu, Q=Q, W=W, y=np.array(ys), u=np.array(us), cost=np.array(cost), noise=noise, y_noise=yn, u_noise=un,
                    x0=prob.x0)
print("wrote", dst, os.path.getsize(dst), "bytes")
