#!/usr/bin/env python3
"""How often are the (unenforced) OV / state bounds of the Van de Vusse NMPC active inside a controller call?  See
bound_activity.cpp.  usage: python tools/nmpc_bounds/run.py [n_candidates]"""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import mpcgpu  # noqa: E402
so = "/tmp/libbound_activity.so"
subprocess.check_call(["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "nmpc_bounds", "bound_activity.cpp")])
lib = C.CDLL(so)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
p = mpcgpu.vandevusse()
N, Nu, dl, lm = mpcgpu.synthetic_nmpc_population(p, n, seed=0)
f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
bufs = [f64(getattr(p, k)) for k in ("x0", "u0", "umin", "umax", "xmin", "xmax", "su", "sy", "r")]
P = lambda a: a.ctypes.data_as(C.c_void_p)
calls = np.zeros(n, dtype=np.int32); worst = np.zeros(n)
N = np.ascontiguousarray(N, dtype=np.int32); Nu = np.ascontiguousarray(Nu, dtype=np.int32); dl = f64(dl); lm = f64(lm)
lib.bound_activity(C.c_int(p.nit), C.c_int(p.nsub), C.c_int(p.max_sqp), C.c_double(p.Ts), *[P(b) for b in bufs], C.c_int(n), P(N), P(Nu), P(dl), P(lm),
                   C.c_double(1e-9), P(calls), P(worst))
tot = n * (p.nit - 1)
print(f"{n} candidates x {p.nit - 1} controller calls = {tot} calls")
print(f"calls whose optimal plan predicts a state outside its bounds: {int(calls.sum())} ({calls.sum() / tot:.2%}), in {int((calls > 0).sum())} candidates ({(calls > 0).mean():.1%})")
print(f"closest approach of any prediction to a bound (signed excess as a fraction of the bound's range; negative = margin left): {worst.max():.3g}")
