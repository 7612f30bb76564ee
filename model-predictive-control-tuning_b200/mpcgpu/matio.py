"""On-disk format either side of the path: the `Tuning_Parameters` struct the
reference writes at the end of a tuning run (MPCTuning.m:371-381: `save([callerName,'_Tuning_',date],
'Tuning_Parameters')`) and reloads with `uigetfile` + `load` when `tuning = false` (Shell3x3.m:169-185).

MAT v5 struct {mpcobj, N, Nu, delta, lambda, scale{L, R, Ru, Rv}, date}.  The `mpcobj` field is a MathWorks MCOS
object: it cannot be written without MATLAB and is skipped on reading, so a file written here carries the tuned
numbers (which is what the reference's scripts read back: Shell3x3.m:176-185) but not the controller object."""
from __future__ import annotations

import datetime

import numpy as np
import scipy.io as sio


def load_tuning(path: str) -> dict:
    """Reads N, Nu, delta, lambda and the scaling diagonals from a reference (or own) `*_Tuning_*.mat`."""
    m = sio.loadmat(path, struct_as_record=False, squeeze_me=True)
    tp = m["Tuning_Parameters"]
    out = {"N": np.atleast_1d(tp.N).astype(int), "Nu": np.atleast_1d(tp.Nu).astype(int),
           "delta": np.atleast_1d(tp.delta).astype(float), "lambda": np.atleast_1d(getattr(tp, "lambda")).astype(float)}
    sc = getattr(tp, "scale", None)
    if sc is not None:
        for k in ("L", "R", "Ru", "Rv"):
            v = getattr(sc, k, None)
            if v is not None and np.size(v):
                v = np.atleast_2d(np.asarray(v, float))
                out[k] = np.diag(v) if v.shape[0] == v.shape[1] else v.ravel()
    out["date"] = str(getattr(tp, "date", ""))
    return out


def save_tuning(path: str, N, Nu, delta, lam, L=None, R=None, nu=None) -> None:
    """Writes the struct of MPCTuning.m:374-380 (without `mpcobj`).  R = diag([Ru; Rv]) as in MPCTuning.m:156-163."""
    scale = {}
    if L is not None:
        scale["L"] = np.diag(np.asarray(L, float))
    if R is not None:
        R = np.asarray(R, float)
        scale["R"] = np.diag(R)
        nu = len(np.atleast_1d(Nu)) if nu is None else int(nu)
        scale["Ru"] = np.diag(R[:nu])
        scale["Rv"] = np.diag(R[nu:]) if len(R) > nu else np.zeros((0, 0))
    tp = {"N": np.atleast_2d(np.asarray(N, float)), "Nu": np.atleast_2d(np.asarray(Nu, float)),
          "delta": np.atleast_2d(np.asarray(delta, float)), "lambda": np.atleast_2d(np.asarray(lam, float)),
          "scale": scale, "date": datetime.datetime.now().strftime("%d-%b-%Y %H:%M:%S")}
    sio.savemat(path, {"Tuning_Parameters": tp}, format="5", oned_as="row")
