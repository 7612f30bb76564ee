#!/usr/bin/env python3
"""Writes tests/golden/oracle_golden_shell3x3.npz from the CPU oracle (oracle/mpc_oracle.c).
These are ORACLE outputs, not reference outputs: the reference pins no per-candidate result
(SURVEY.md §4); the file freezes the oracle so that a later change to either side is noticed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mpcgpu import shell3x3, synthetic_population  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from parity_util import vns_well_posed  # noqa: E402

p = shell3x3(2)
op = orc.OracleProblem(p)
N, Nu, dl, lm = synthetic_population(p, 14, seed=99, wlo=1e-3, whi=3.0)
# the reference's own tuned results as candidates 0 and 1
N = np.concatenate([[24, 12], N]).astype(np.int32); Nu = np.concatenate([[6, 4], Nu]).astype(np.int32)
dl = np.vstack([[0.010659948215964849, 0.004019856475662751, 0.0007926546087416782],
                [0.04984473720828972, 0.039681527811050415, 0.010544968016995362], dl])
lm = np.vstack([[9.247457388705409e-05, 0.0005523146971406108, 0.0015219790494510478],
                [0.06524463112453442, 0.0016951671507189326, 0.07656112033860867], lm])
gam, st, _ = orc.eval_batch(op, N, Nu, dl, lm, "gam")
vns, st2, _ = orc.eval_batch(op, N, Nu, dl, lm, "vns")
assert (st == 0).all() and (st2 == 0).all()
ys, us = [], []
for c in range(len(N)):
    y, u, *_ = orc.closedloop(op, N[c], Nu[c], dl[c], lm[c])
    ys.append(y[:, ::10]); us.append(u[:, ::10])
ok = vns_well_posed(p, lambda r: orc.OracleProblem(p, r=r), N, Nu, dl, lm)
dst = os.path.join(ROOT, "tests", "golden", "oracle_golden_shell3x3.npz")
np.savez_compressed(dst, N=N, Nu=Nu, delta=dl, lam=lm, gam=gam, vns=vns, vns_ok=ok, y_sub=np.array(ys), u_sub=np.array(us))
print("wrote", dst, os.path.getsize(dst), "bytes; vns well-posed:", int(ok.sum()), "of", len(ok))
