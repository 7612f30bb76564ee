"""Prototype (numpy, offline): primal-dual active-set iterations on the QPs of heavy Shell3x3 candidates,
warm-started from the previous sample's optimal set.  Exploration tool only."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-predictive-control-tuning_b200"))
import numpy as np
import mpcgpu
from oracle import oracle as orc

p = mpcgpu.shell3x3(2)
op = orc.OracleProblem(p)
N, Nu, dl, lm = mpcgpu.synthetic_population(p, 4096, seed=0)
ny, nu, nit = 3, 3, p.nit
ch = p.plant
dmax = int(ch.d.max()); hl = dmax + 2


def constraints(m, uprev):
    """rows: normals (nc x nz), bounds b: n'z >= b  (all as >=)"""
    nz = nu * m
    rows, b, ids = [], [], []
    for c in range(m):
        for j in range(nu):
            e = np.zeros(nz); e[c * nu + j] = 1.0
            s = np.zeros(nz)
            for cc in range(c + 1): s[cc * nu + j] = 1.0
            rows += [e, -e, s, -s]
            b += [p.dumin[j], -p.dumax[j], p.umin[j] - uprev[j], -(p.umax[j] - uprev[j])]
            ids += [(0, c, j), (1, c, j), (2, c, j), (3, c, j)]
    return np.array(rows), np.array(b), ids


def eqp(H, f, Nm, b, A):
    nz = len(f); q = len(A)
    if q == 0:
        return np.linalg.solve(H, -f), np.zeros(0), False
    NA = Nm[A]
    K = np.block([[H, -NA.T], [NA, np.zeros((q, q))]])
    rhs = np.concatenate([-f, b[A]])
    sing = False
    try:
        sol = np.linalg.solve(K, rhs)
        if not np.isfinite(sol).all() or np.linalg.cond(K) > 1e14: raise np.linalg.LinAlgError
    except np.linalg.LinAlgError:
        sol = np.linalg.lstsq(K, rhs, rcond=1e-12)[0]; sing = True
    return sol[:nz], sol[nz:], sing


def pdas(H, f, Nm, b, A0, itmax=40):
    A = sorted(A0); seen = []
    for it in range(1, itmax + 1):
        z, mu, sing = eqp(H, f, Nm, b, A)
        s = Nm @ z - b
        An = sorted([a for a, mval in zip(A, mu) if mval > 1e-12 * (np.abs(mu).max() + 1e-300)] +
                    [i for i in range(len(b)) if i not in A and s[i] < -1e-10])
        if An == A:
            return it, z, A, sing
        if An in seen: return -it, z, A, sing
        seen.append(A); A = An
    return -itmax, z, A, False


def analyse(c):
    m, P = int(Nu[c]), int(N[c])
    y, u, ys, uo, rc, stats = orc.closedloop(op, P, m, dl[c], lm[c], open_loop=False)
    xs = np.zeros(ny * nu); wh = np.zeros((nu, hl)); up = np.zeros(nu)
    prevA = []; res = []
    for k in range(nit):
        z, H, f, G, yf, it, rc = orc.single_qp(op, P, m, dl[c], lm[c], xs, wh, up.copy(), p.r[k], up.copy())
        Nm, b, ids = constraints(m, up)
        s = Nm @ z - b
        Aopt = [i for i in range(len(b)) if s[i] < 1e-9]
        if Aopt or prevA:
            # shifted guess
            sh = []
            for i in prevA:
                t, cc, j = ids[i]
                if cc > 0: sh.append(ids.index((t, cc - 1, j)))
            it_un, z1, A1, s1 = pdas(H, f, Nm, b, prevA)
            it_sh, z2, A2, s2 = pdas(H, f, Nm, b, sh)
            it_cold, z3, A3, s3 = pdas(H, f, Nm, b, [])
            e1 = np.abs(z1 - z).max() if it_un > 0 else np.nan
            e2 = np.abs(z2 - z).max() if it_sh > 0 else np.nan
            res.append((k, len(Aopt), it, it_un, it_sh, it_cold, e1, e2, s1 or s2))
        prevA = Aopt
        # advance plant with the oracle's move
        up = up + z[:nu]
        assert np.abs(up - u[:, k]).max() < 1e-9
        wh[:, 1:] = wh[:, :-1].copy(); wh[:, 0] = up
        for i in range(ny):
            for j in range(nu):
                dd = int(ch.d[i, j])
                w0 = wh[j, dd - 1] if dd >= 1 else 0.0
                w1 = wh[j, dd]
                xs[i * nu + j] = ch.a[i, j] * xs[i * nu + j] + ch.b0[i, j] * w0 + ch.b1[i, j] * w1
    return res


if __name__ == "__main__":
    for c in [int(a) for a in sys.argv[1:]] or [2318]:
        res = analyse(c)
        r = np.array([x[:6] for x in res])
        print("cand", c, "N", N[c], "Nu", Nu[c], "constrained samples", len(res))
        print(" GI cold iterations total", r[:, 2].sum(), " PDAS unshifted", np.abs(r[:, 3]).sum(), "fails", (r[:, 3] < 0).sum(),
              " shifted", np.abs(r[:, 4]).sum(), "fails", (r[:, 4] < 0).sum(), " cold", np.abs(r[:, 5]).sum(), "fails", (r[:, 5] < 0).sum())
        for x in res[:40]:
            print("  k %3d q* %2d GI %3d | pdas unshifted %3d shifted %3d cold %3d | err %.1e %.1e sing %s" % x)
