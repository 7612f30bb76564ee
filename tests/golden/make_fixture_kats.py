#!/usr/bin/env python3
"""Decode the reference's saved tuning results into a small JSON known-answer file.

Run in the BUILD container only (needs /root/reference); the output
tests/golden/fixture_kats.json is committed and is what the tests read.

Source files: MPC-Tuning/*.mat, written by MPCTuning.m:371-381
(`Tuning_Parameters` struct = {mpcobj,N,Nu,delta,lambda,scale,date}).
The `mpcobj` field is an MCOS opaque object; its property values live in the
`__function_workspace__` blob.
"""
import io, json, os, sys
import numpy as np
import scipy.io as sio
from scipy.io.matlab._mio5 import MatFile5Reader

REF = "/root/reference/MPC-Tuning"
FILES = [
    "Shell3x3_Tuning_25Jul2023_12_06.mat",
    "Shell3x3_Tuning_Caso2.mat",
    "Shell7x5_Tuning_14Sep2024_14_22.mat",
    "Shell7x5_Tuning_25Jul2023_12_18.mat",
    "VanDeVusse_NMPC_Tuning_25Jul2023_11_04.mat",
    "VanDeVusse_NMPC_Tuning_06Dec2023_09_50.mat",
]


def mcos_cells(path):
    m = sio.loadmat(path, struct_as_record=False, squeeze_me=True)
    raw = m["__function_workspace__"].tobytes()
    hdr = bytearray(128)
    hdr[:4] = b"MATL"
    hdr[124:126] = b"\x00\x01"
    hdr[126:128] = b"IM"
    bio = io.BytesIO(bytes(hdr) + raw[8:])
    rd = MatFile5Reader(bio, struct_as_record=True, squeeze_me=False)
    rd.initialize_read()
    bio.seek(128)
    h, _ = rd.read_var_header()
    arr = rd.read_var_array(h, process=False)
    cells = arr["MCOS"][0, 0]["_ObjectMetadata"][0]
    return m["Tuning_Parameters"], [c[0] if isinstance(c, np.ndarray) and c.dtype == object and c.shape == (1,) else c for c in cells]


def tolist(x):
    x = np.asarray(x, dtype=float)
    return [[(None if not np.isfinite(v) else float(v)) if not np.isinf(v) else ("inf" if v > 0 else "-inf") for v in row] for row in np.atleast_2d(x)]


def vec(x):
    return [("inf" if v == np.inf else "-inf" if v == -np.inf else float(v)) for v in np.asarray(x, dtype=float).ravel()]


def decode_linear(path):
    tp, cells = mcos_cells(path)
    out = {
        "N": int(np.max(tp.N)), "Nu": [int(v) for v in np.atleast_1d(tp.Nu)],
        "delta": vec(tp.delta), "lambda": vec(getattr(tp, "lambda")),
        "L": vec(np.diag(tp.scale.L)), "R": vec(np.diag(tp.scale.R)),
        "Ru": vec(np.diag(np.atleast_2d(tp.scale.Ru))),
        "Rv": vec(np.diag(np.atleast_2d(tp.scale.Rv))) if np.size(tp.scale.Rv) else [],
    }
    num, den, dly = cells[15], cells[16], cells[17]
    ny, nw = num.shape
    out["ny"], out["nw"] = ny, nw
    out["num"] = [[vec(num[i, j]) for j in range(nw)] for i in range(ny)]
    out["den"] = [[vec(den[i, j]) for j in range(nw)] for i in range(ny)]
    out["iodelay"] = [[float(v) for v in row] for row in dly["IO"][0, 0]]
    out["Ts"] = float(np.asarray(cells[18]).ravel()[0])
    mv, ov, dv = cells[27], cells[28], cells[29]
    def fld(s, name):
        return vec([np.asarray(s[0, k][name]).ravel()[0] for k in range(s.shape[1])])
    out["MV"] = {k: fld(mv, k) for k in ("Min", "Max", "MinECR", "MaxECR", "RateMin", "RateMax", "ScaleFactor")}
    out["OV"] = {k: fld(ov, k) for k in ("Min", "Max", "MinECR", "MaxECR", "ScaleFactor")}
    if isinstance(dv, np.ndarray) and dv.dtype.names:
        out["DV"] = {"ScaleFactor": fld(dv, "ScaleFactor")}
    w = cells[30][0, 0]
    out["Weights"] = {"MV": vec(w["ManipulatedVariables"]), "MVRate": vec(w["ManipulatedVariablesRate"]),
                      "OV": vec(w["OutputVariables"]), "ECR": float(np.asarray(w["ECR"]).ravel()[0])}
    opt = cells[31][0, 0]
    out["Optimizer"] = {"Algorithm": str(np.asarray(opt["Algorithm"]).ravel()[0]),
                        "ConstraintTolerance": float(np.asarray(opt["ActiveSetOptions"][0, 0]["ConstraintTolerance"]).ravel()[0])}
    return out


def decode_nmpc(path):
    m = sio.loadmat(path, struct_as_record=False, squeeze_me=True)
    tp = m["Tuning_Parameters"]
    return {"N": int(np.max(tp.N)), "Nu": [int(v) for v in np.atleast_1d(tp.Nu)],
            "delta": vec(tp.delta), "lambda": vec(getattr(tp, "lambda"))}


def main():
    res = {}
    for f in FILES:
        p = os.path.join(REF, f)
        res[f] = decode_nmpc(p) if f.startswith("VanDeVusse") else decode_linear(p)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixture_kats.json")
    with open(dst, "w") as fh:
        json.dump(res, fh, indent=1)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
