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
elif which == "sim_vns":
    import mpcgpu, copy
    p = copy.copy(mpcgpu.shell3x3(2)); p.nit = 50; p.r = p.r[:50].copy(); p.v = p.v[:50].copy(); p.yref = p.yref[:, :50].copy()
    N, Nu, dl, lm = mpcgpu.synthetic_population(p, 2, seed=4, wlo=1e-3, whi=3.0); lm[1] *= 1e-2
    print("ok", emu.eval_batch(p, N, Nu, dl, lm, "vns")[0])
elif which == "sim_wb":
    import mpcgpu, copy
    p = copy.copy(mpcgpu.woodberry()); p.nit = 330; p.r = p.r[:330].copy(); p.v = p.v[:330].copy(); p.yref = p.yref[:, :330].copy()
    N, Nu, dl, lm = mpcgpu.synthetic_population(p, 1, seed=2, wlo=1e-3, whi=3.0); lm[0] *= 1e-2
    print("ok", emu.eval_batch(p, N, Nu, dl, lm, "gam", traj=True)[0])
elif which == "sim_cycle":
    import mpcgpu
    p = mpcgpu.shell3x3(2)
    Ng, Nug, dlg, lmg = mpcgpu.synthetic_population(p, 32768, seed=0)
    i = 4169      # a period-2 limit cycle: the two-phase mode (parked factors) is entered
    print("ok", emu.eval_batch(p, Ng[i:i+1], Nug[i:i+1], dlg[i:i+1], lmg[i:i+1], "gam")[0])
elif which == "est":
    import mpcgpu, copy
    from mpcgpu import estimator as est
    p = copy.copy(mpcgpu.woodberry()); p.nit = 80; p.r = p.r[:80].copy(); p.v = p.v[:80].copy(); p.yref = p.yref[:, :80].copy()
    plant = est.woodberry_real_plant(); hl = est.history_length(p, plant); M = est.default_estimator_gain(p, hl)
    print("ok", emu.eval_est(p, plant, M, hl, 14, 3, np.array([0.5, 0.5]), np.array([0.3, 0.3]))[2])
elif which == "nmpc8":
    from mpcgpu.nmpc import vandevusse
    p = vandevusse()
    print("ok", emu.nmpc_eval_group(p, [3, 6, 10, 5], [2, 3, 2, 2], [[0.09, 0.11], [0.5, 2.0], [1, 1], [0.02, 0.7]], [[0.25, 0.12], [0.01, 0.3], [0.1, 0.1], [0.005, 0.004]], G=8))
elif which == "nmpc_vns":
    from mpcgpu.nmpc import vandevusse
    p = vandevusse()
    print("ok", emu.nmpc_eval(p, 6, 2, [1.0, 1.0], [0.1, 0.1], mode="vns")[0])
elif which == "build":
    import mpcgpu
    p = mpcgpu.shell3x3(2)
    M, W, rc = emu.build_matrices(p, 60, 9, 16, np.array([0.4, 1.0, 0.2]), np.array([0.1, 0.3, 0.05]), threads=128)
    print("ok", rc, float(np.abs(W).max()))
elif which == "soft_idx":
    import mpcgpu, copy
    p = copy.copy(mpcgpu.shell7x5()); p.nit = 40; p.r = p.r[:40].copy(); p.v = p.v[:40].copy(); p.yref = p.yref[:, :40].copy()
    N, Nu, dl, lm = mpcgpu.synthetic_population(p, 1, seed=5, wlo=1e-2); N[0] = 9; Nu[0] = 7      # p < 1.5 m: horizon-order pivoting
    print("ok", emu.eval_batch(p, N, Nu, dl, lm, "gam")[0])
