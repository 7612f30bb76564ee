"""Throwaway: ADMM convergence on Shell3x3 tuning QPs (numpy), to choose the penalty rule of the kernel variant."""
import sys, numpy as np
sys.path.insert(0, "model-predictive-control-tuning_b200")
from mpcgpu.problems import shell3x3, synthetic_population
from mpcgpu.plant import simulate

prob = shell3x3()
ny, nu = prob.ny, prob.nu
ch = prob.plant
S = np.zeros((200, ny, nu))
for j in range(nu):
    w = np.zeros((200, ch.a.shape[1])); w[:, j] = 1.0
    S[:, :, j] = simulate(ch, w)

def build(p, m, delta, lam):
    nz = nu * m
    G = np.zeros((p * ny, nz))
    for t in range(1, p + 1):
        for c in range(m):
            if t - 1 - c >= 0:
                G[(t - 1) * ny:(t) * ny, c * nu:(c + 1) * nu] = S[t - 1 - c]
    wy = np.tile((delta / prob.sy) ** 2, p); wu = np.tile((lam / prob.su) ** 2, m)
    H = G.T @ (wy[:, None] * G) + np.diag(wu)
    # A = [I; T] with T = cumulative sum per input
    T = np.zeros((nz, nz))
    for c in range(m):
        for cc in range(c + 1):
            for j in range(nu):
                T[c * nu + j, cc * nu + j] = 1.0
    return H, G, T

def admm(H, T, zu, lo, hi, rho, alpha=1.6, eps=1e-9, itmax=5000):
    nz = len(zu)
    A = np.vstack([np.eye(nz), T])
    K = np.linalg.inv(H + rho * A.T @ A)
    Azu = A @ zu
    v = np.clip(Azu, lo, hi); u = np.zeros(2 * nz)
    for it in range(1, itmax + 1):
        x = zu + rho * K @ (A.T @ (v - u - Azu))
        Ax = A @ x
        xh = alpha * Ax + (1 - alpha) * v
        vn = np.clip(xh + u, lo, hi)
        u = u + xh - vn
        res = max(np.abs(Ax - vn).max(), np.abs(vn - v).max())
        v = vn
        if res < eps: break
    return x, it

EPS = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-9
ALL = []
N, Nu, dl, lm = synthetic_population(prob, 24, seed=3)
lo1 = np.concatenate; 
for c in range(24):
    p, m = int(N[c]), int(Nu[c])
    H, G, T = build(p, m, dl[c], lm[c])
    nz = nu * m
    ev = np.linalg.eigvalsh(H)
    W = np.linalg.inv(H)
    r0 = np.sqrt(np.trace(H) / np.trace(W))
    r1 = np.sqrt(ev[0] * ev[-1])
    lo = np.concatenate([np.tile(prob.dumin, m), np.tile(prob.umin, m)]); hi = np.concatenate([np.tile(prob.dumax, m), np.tile(prob.umax, m)])
    # unconstrained optimum of a set-point step: zu = W G' Wy r
    wy = np.tile((dl[c] / prob.sy) ** 2, p)
    e = np.tile(np.array([1.0, -0.5, 0.7])[:ny], p)
    zu = W @ (G.T @ (wy * e))
    A = np.vstack([np.eye(nz), T]); zu = zu * (3.0 / np.max(np.abs(A @ zu) / np.maximum(np.abs(hi), np.abs(lo))))
    out = []
    for sc in (0.001, 0.003, 0.01, 0.03, 0.1, 1.0):
        x, it = admm(H, T, zu, lo, hi, r0 * sc, eps=EPS)
        out.append(it)
    ALL.append(out)
    viol = max((A @ zu - hi).max(), (lo - A @ zu).max())
    print(f"p={p:3d} m={m:2d} cond={ev[-1]/ev[0]:.1e} r0={r0:.2e} r1={r1:.2e} viol={viol:.2e} its(r0*[.001,.003,.01,.03,.1,1])={out}")

ALL = np.array(ALL)
print('eps', EPS, 'median iterations per penalty scale', np.median(ALL, axis=0), 'share not converged in 5000', (ALL >= 5000).mean(axis=0))
