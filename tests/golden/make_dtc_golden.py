#!/usr/bin/env python3
"""Writes tests/golden/oracle_golden_dtc.npz from the CPU oracle of the DTC-GPC path
(oracle/dtc_gpc_oracle.py, a restatement of /root/reference/DTC-GPC/*.m).  ORACLE outputs: the reference
only plots this path.  Candidate 0 is the reference script's own setting (DTC_GPC_WW.m:56-64,108)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
from mpcgpu.dtcgpc import woodberry_dtc, synthetic_dtc_population  # noqa: E402
from oracle import dtc_gpc_oracle as dorc  # noqa: E402

prob = woodberry_dtc()
p, m, dl, lm, alfa, raio = synthetic_dtc_population(prob, 12, seed=7)
p[0] = (3, 3); m[0] = (3, 3); dl[0] = (1, 1); lm[0] = (1, 1); alfa[0] = 0.7; raio[0] = 0.8
ys, us, ise = [], [], []
for c in range(len(p)):
    fr = dorc.mimofilter_Fr(prob.pnz, alfa[c], raio[c])
    y, u = dorc.dtc_gpc_closed_loop(prob, p[c], m[c], dl[c], lm[c], fr)
    ys.append(y[:, ::4]); us.append(u[:, ::4]); ise.append(((y - prob.r) ** 2).sum(axis=1))
dst = os.path.join(ROOT, "tests", "golden", "oracle_golden_dtc.npz")
np.savez_compressed(dst, p=p, m=m, delta=dl, lam=lm, alfa=alfa, raio=raio, ise=np.array(ise), y_sub=np.array(ys), u_sub=np.array(us))
print("wrote", dst, os.path.getsize(dst), "bytes; ise range", np.min(ise), np.max(ise))
