"""Small cases for tests/host_emulation/tsan.sh (the emulation library built with ThreadSanitizer is loaded instead of the plain one)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, emu
emu._lib = C.CDLL("/tmp/libmpcemu_tsan.so")
which = sys.argv[1]
if which == "nmpc":
    from mpcgpu.nmpc import vandevusse
    p = vandevusse()
    print('ok', emu.nmpc_eval_group(p, [3, 6], [2, 3], [[0.09, 0.11], [0.5, 2.0]], [[0.25, 0.12], [0.01, 0.3]], G=16))
elif which == "dtc":
    from mpcgpu.dtcgpc import woodberry_dtc, synthetic_dtc_population
    prob = woodberry_dtc(); P = synthetic_dtc_population(prob, 4, seed=3)
    print('ok', emu.dtc_eval(prob, *P)[0])
elif which == "sim":
    import mpcgpu, copy
    p = copy.copy(mpcgpu.shell3x3(2)); p.nit = 60; p.r = p.r[:60].copy(); p.v = p.v[:60].copy(); p.yref = p.yref[:, :60].copy()
    N, Nu, dl, lm = mpcgpu.synthetic_population(p, 2, seed=2, wlo=1e-3, whi=3.0); lm[0] *= 1e-2
    print("ok", emu.eval_batch(p, N, Nu, dl, lm, "gam")[0])
elif which == "soft":
    import mpcgpu, copy
    p = copy.copy(mpcgpu.shell7x5()); p.nit = 40; p.r = p.r[:40].copy(); p.v = p.v[:40].copy(); p.yref = p.yref[:, :40].copy()
    N, Nu, dl, lm = mpcgpu.synthetic_population(p, 1, seed=5, wlo=1e-2); N[0] = 20; Nu[0] = 4
    print("ok", emu.eval_batch(p, N, Nu, dl, lm, "gam")[0])
