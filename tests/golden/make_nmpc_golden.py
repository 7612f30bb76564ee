#!/usr/bin/env python3
"""Writes tests/golden/oracle_golden_nmpc.npz from the CPU oracle of the nonlinear path (oracle/nmpc_oracle.py:
scipy trust-region least squares on the restated nlmpcmove problem).  ORACLE outputs -- the reference pins only
the tuned (N, Nu, delta, lambda) of this case (candidate 0 below, VanDeVusse_NMPC_Tuning_*.mat / BASELINE.md)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
from mpcgpu.nmpc import vandevusse  # noqa: E402
from oracle import nmpc_oracle as no  # noqa: E402

prob = vandevusse()
N = np.array([3, 10, 6, 8, 5], dtype=np.int32); Nu = np.array([2, 2, 3, 4, 2], dtype=np.int32)
delta = np.array([[0.093022247804, 0.113338402058], [1, 1], [0.5, 2.0], [3.0, 0.05], [0.02, 0.7]])
lam = np.array([[0.245996189228, 0.123108010965], [0.1, 0.1], [0.01, 0.3], [1.0, 0.02], [0.005, 0.004]])
ys, us, yos, uos, gam, st = [], [], [], [], [], []
for c in range(len(N)):
    y, u, yo, uo, s = no.closedloop_toolbox_nmpc(prob, prob.r, N[c], Nu[c], delta[c], lam[c])
    ys.append(y); us.append(u); yos.append(yo); uos.append(uo); st.append(s)
    gam.append(((y - prob.yref) ** 2).sum(axis=1))
vns = [no.vns_cost(prob, N[c], Nu[c], delta[c], lam[c])[0] for c in range(2)]
dst = os.path.join(ROOT, "tests", "golden", "oracle_golden_nmpc.npz")
np.savez_compressed(dst, N=N, Nu=Nu, delta=delta, lam=lam, y=np.array(ys), u=np.array(us), yopt=np.array(yos), uopt=np.array(uos),
                    gam=np.array(gam), vns=np.array(vns), status=np.array(st), x0=prob.x0)
print("wrote", dst, os.path.getsize(dst), "bytes; gam", np.array(gam), "vns", vns, "status", st)
