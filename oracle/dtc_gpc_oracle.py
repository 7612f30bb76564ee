"""CPU ORACLE for the DTC-GPC path (BASELINE.json configs[3]).  TEST INFRASTRUCTURE ONLY: may be imported
only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs; the product never calls it.

Line-by-line numpy restatement of the reference's explicit MATLAB (all under /root/reference/DTC-GPC/):
    descompMPC.m:19-43, BA_MIMO.m:16-72, diophantine.m:15-79, diophantineMIMO.m:14-22, MatG.m:38-74,
    deltaUFree.m:12-63, cell2mat2.m:25-58, filtro_siso.m:26-96, mimofilter.m:33-50 (Fr only),
    OptimalPredictor2.m:24-40, DTC_GPC_WW.m:56-164 (gain and control loop).
Restatement notes:
  * `tf`/`c2d`/`lsim`/`step`/`roots`/`poly`/`minreal` come from MathWorks toolboxes; for the
    first-order-plus-dead-time channels of the reference they have closed forms (mpcgpu/plant.py, pinned
    against the reference's saved objects).  `roots` of a product of first-order factors returns the factors'
    poles; BA_MIMO rounds them to 4 decimals before use (BA_MIMO.m:38-40), which is reproduced.
  * the reference re-simulates every model from t = 0 at every step (`lsim` over the whole history,
    DTC_GPC_WW.m:131-132, OptimalPredictor2.m:28-37); a causal LTI model gives the same samples when it
    is advanced one step at a time, which is what is done here.
PARITY: the algorithm is in the reference tree, so this file follows code, not documentation; but the
reference pins no numeric output for this path either (it only plots), so parity is pinned by the
identities in tests/test_dtc_oracle.py (Diophantine identity, G against the step response, prediction
consistency) rather than by golden vectors.
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------
# polynomial preparation
# ---------------------------------------------------------------------------------------------
def descompMPC(ch):
    """descompMPC.m:19-43 on a matrix of discrete FOPDT channels (a, b0, b1, d):
    B{i,j} = num, A{i,j} = den, d = iodelay; when the leading numerator coefficient is non-zero the delay is
    reduced by one and a zero is prepended (:35-38)."""
    ny, nu = ch.a.shape
    B = [[None] * nu for _ in range(ny)]
    A = [[None] * nu for _ in range(ny)]
    d = np.array(ch.d, dtype=int).copy()
    for i in range(ny):
        for j in range(nu):
            num = np.array([ch.b0[i, j], ch.b1[i, j]], dtype=float)
            A[i][j] = np.array([1.0, -ch.a[i, j]])
            if num[0] != 0:
                d[i, j] -= 1
                num = np.concatenate([[0.0], num])
            B[i][j] = num
    return B, A, d


def BA_MIMO(Bn, An):
    """BA_MIMO.m:16-72: common denominator per output from the poles ROUNDED to 4 decimals (:38-40),
    numerators multiplied by the missing pole factors (:45-62)."""
    p, m = len(An), len(An[0])
    Bn = [[np.array(b, float) for b in row] for row in Bn]
    for i in range(p):
        for j in range(m):
            if Bn[i][j][0] == 0:
                Bn[i][j] = Bn[i][j][1:]
    A = [None] * p
    B = [[None] * m for _ in range(p)]
    for i in range(p):
        aux = An[i][0]
        for j in range(1, m):
            aux = np.convolve(aux, An[i][j])
        A[i] = aux
        if p != 1:
            Au1 = np.round(np.real(np.roots(aux)), 4)
            Pol = np.unique(Au1)
            A[i] = np.poly(Pol)
    for i in range(p):
        for j in range(m):
            aux = Bn[i][j]
            rA = list(np.round(np.real(np.roots(A[i])), 4))
            rAn = list(np.round(np.real(np.roots(An[i][j])), 4))
            kk = 0
            while kk < len(rA):             # BA_MIMO.m:50-58 (incl. its skip-after-removal behaviour)
                for jj in range(len(rAn)):
                    if kk < len(rA) and rA[kk] == rAn[jj]:
                        val = rA[kk]
                        rA = [x for x in rA if x != val]
                kk += 1
            pA = np.poly(rA) if len(rA) else np.array([1.0])
            B[i][j] = np.convolve(aux, pA)
    na = np.array([len(A[i]) - 1 for i in range(p)])
    nb = np.array([[len(B[i][j]) - 1 for j in range(m)] for i in range(p)])
    return B, A, na, nb


def diophantine(A, N, d):
    """diophantine.m:15-79."""
    AD = np.convolve(A, [1.0, -1.0])
    nAD = len(AD)
    N1, N2 = d + 1, d + N
    f = np.zeros((N2 + 1, nAD - 1))
    f[0, 0] = 1.0
    for j in range(N2):
        for i in range(nAD - 2):
            f[j + 1, i] = f[j, i + 1] - f[j, 0] * AD[i + 1]
        f[j + 1, nAD - 2] = -f[j, 0] * AD[nAD - 1]
    F = f[N1:N2 + 1, :]
    E = np.zeros((N2, N2))
    e = np.zeros(N2)
    e[0] = 1.0
    E[0, 0] = 1.0
    for i in range(1, N2):
        e[i] = f[i, 0]
        E[i, :i + 1] = e[:i + 1]
    return E[N1 - 1:N2, :], F


def diophantineMIMO(A, N, dmin):
    """diophantineMIMO.m:14-22."""
    E, En, F = [], [], []
    for i in range(len(A)):
        En1, Fn = diophantine(A[i], int(N[i]), int(dmin[i]))
        E.append(En1[-1, :]); F.append(Fn); En.append(En1)
    return E, En, F


def step_response(ch, i, j, n):
    """step(Ps(i,j), n*Ts): samples s(0..n) of the discrete channel incl. its delay."""
    a, b0, b1, d = ch.a[i, j], ch.b0[i, j], ch.b1[i, j], int(ch.d[i, j])
    s = np.zeros(n + 1)
    for k in range(1, n + 1):
        s[k] = a * s[k - 1] + (b0 if k - d >= 0 else 0.0) + (b1 if k - d - 1 >= 0 else 0.0)
    return s


def MatG(ch, N, Nu, d):
    """MatG.m:38-74: G(r,k) = g(dmin_i + 2 + r - k) (1-based), blocks output-major / input-major."""
    s_, e_ = ch.a.shape
    dmin = d.min(axis=1) if e_ > 1 else d[:, 0]
    H = [[None] * e_ for _ in range(s_)]
    for i in range(s_):
        for j in range(e_):
            g = step_response(ch, i, j, int(N[i] + dmin[i]))
            G = np.zeros((int(N[i]), int(Nu[j])))
            for k in range(1, int(Nu[j]) + 1):
                G[k - 1:, k - 1] = g[dmin[i] + 1: dmin[i] + int(N[i]) - k + 2]
            H[i][j] = G
    return np.block(H), H


def deltaUFree(B, En, N, dp):
    """deltaUFree.m:12-63 incl. its removal of EVERY exact zero of conv(E_i, B) (:39-45)."""
    ny, nu = len(B), len(B[0])
    uG = [[None] * nu for _ in range(ny)]
    for m in range(ny):
        for n in range(nu):
            cp = int(dp[m, n]) + len(B[m][n]) - 1
            if cp < 1:
                cp = 1
            uG1 = np.zeros((int(N[m]), cp))
            for i in range(int(N[m])):
                aux = np.convolve(En[m][i, :], B[m][n])
                BE = aux[aux != 0]
                lBE = len(BE)
                if lBE < cp:
                    uG1[i, cp - lBE:] = BE
                else:
                    uG1[i, :] = BE[lBE - cp:]
            uG[m][n] = uG1
    return uG


def cell2mat2(Bc):
    """cell2mat2.m:25-58: ragged block concatenation; block (i,j) sits top-left in a cell whose height is
    the tallest block of row i and whose width is the widest block of column j (:45-55; MATLAB grows A
    past its initial n1 columns when the column maxima demand it)."""
    m, n = len(Bc), len(Bc[0])
    f1 = np.array([[Bc[i][j].shape[0] for j in range(n)] for i in range(m)])
    c1 = np.array([[Bc[i][j].shape[1] for j in range(n)] for i in range(m)])
    f1m = np.concatenate([[0], f1.max(axis=1)])
    c1m = np.concatenate([[0], c1.max(axis=0)])
    A = np.zeros((int(f1m.sum()), int(c1m.sum())))
    for i in range(m):
        for j in range(n):
            r0, c0 = int(f1m[:i + 1].sum()), int(c1m[:j + 1].sum())
            A[r0:r0 + f1[i, j], c0:c0 + c1[i, j]] = Bc[i][j]
    return A


# ---------------------------------------------------------------------------------------------
# robustness filter (input to the path: filter *design* is outside the hot path)
# ---------------------------------------------------------------------------------------------
def filtro_siso(poles, d, alfa, raio):
    """filtro_siso.m:26-96 for a delay-free model with the given poles and dead time d (samples).
    Returns (Nr, Dr) of Fr(z) = Nr(z)/Dr(z)."""
    p_ind = [pz for pz in poles if abs(pz) >= raio]
    nm = len(p_ind)
    nk = nm
    pd = 0
    if d == 0:
        pd = 2
        nk = nk + pd
    px = np.poly([1.0] + list(p_ind))
    Dr = np.array([1.0])
    for _ in range(nk):
        Dr = np.convolve(Dr, [1.0, -alfa])
    ordem = (len(Dr) - 1) + d
    lpx = len(px)
    A = np.zeros((ordem + 1, ordem + 1 + pd))
    ip = 1
    for j in range(ordem + 2 - d, ordem + 1 + pd + 1):      # 1-based columns
        pn = 1
        for i in range(ip, ordem + 2):
            if pn <= lpx:
                A[i - 1, j - 1] = px[pn - 1]
                pn += 1
        ip += 1
    j = 1
    for i in range(d + 1, ordem + 2):
        A[i - 1, j - 1] = 1.0
        j += 1
    Bv = np.zeros(ordem + 1)
    Bv[0] = 1.0
    Bv[1:len(Dr)] = Dr[1:]
    X = np.linalg.lstsq(A, Bv, rcond=None)[0] if A.shape[0] != A.shape[1] else np.linalg.solve(A, Bv)
    Nr = X[:ordem + 1 - d]
    if not p_ind:
        return np.array([1.0]), np.array([1.0])
    return Nr, Dr


def mimofilter_Fr(ch, alfa, raio):
    """mimofilter.m:33-50, diagonal Fr only: per output, the product of its channels' delay-free models
    (poles a_ij; `minreal` merges repeated factors of the *same* channel product only when they cancel,
    which first-order factors with non-zero numerators never do) with the row's minimum dead time."""
    ny, nu = ch.a.shape
    out = []
    for i in range(ny):
        poles = [ch.a[i, j] for j in range(nu) if (ch.b0[i, j] + ch.b1[i, j]) != 0]
        dmin = int(np.min(ch.d[i, :]))
        out.append(filtro_siso(poles, dmin, alfa, raio) if poles else (np.array([1.0]), np.array([1.0])))
    return out


# ---------------------------------------------------------------------------------------------
# DTC_GPC_WW.m:56-164
# ---------------------------------------------------------------------------------------------
class _Chan:
    """running simulation of a matrix of discrete FOPDT channels (lsim advanced one sample at a time)"""

    def __init__(self, ch, dshift=None):
        self.ch = ch
        self.d = np.array(ch.d, int) if dshift is None else np.array(ch.d, int) - dshift[:, None]
        self.x = np.zeros(ch.a.shape)

    def output_at(self, k, w):
        """advance to sample k (0-based) given the input history w[:, :k] ; returns y(k)"""
        ny, nu = self.x.shape
        for i in range(ny):
            for j in range(nu):
                dd = self.d[i, j]
                u0 = w[j, k - dd] if k - dd >= 0 else 0.0
                u1 = w[j, k - dd - 1] if k - dd - 1 >= 0 else 0.0
                self.x[i, j] = self.ch.a[i, j] * self.x[i, j] + self.ch.b0[i, j] * u0 + self.ch.b1[i, j] * u1
        return self.x.sum(axis=1)


def dtc_gpc_setup(pnz, p, m, delta, lam):
    """DTC_GPC_WW.m:41-105: returns dict(H, S, Hp, Km, na, duM, dmin, dnz)."""
    ny, nu = pnz.a.shape
    Bp, Ap, dp = descompMPC(pnz)
    dmin = dp.min(axis=1)
    dnz = dp - dmin[:, None]
    W = np.diag(np.concatenate([np.full(int(m[j]), lam[j]) for j in range(nu)]))      # :67-71 (weights NOT squared)
    Q = np.diag(np.concatenate([np.full(int(p[i]), delta[i]) for i in range(ny)]))    # :72-76
    B, A, na, nb = BA_MIMO(Bp, Ap)                                                      # :79
    E, En, F = diophantineMIMO(A, p, np.zeros(ny, int))                                 # :80
    S = np.zeros((int(np.sum(p)), int(np.sum(na + 1))))
    r0 = c0 = 0
    for i in range(ny):                                                                 # :83-86
        S[r0:r0 + int(p[i]), c0:c0 + na[i] + 1] = F[i][:int(p[i]), :]
        r0 += int(p[i]); c0 += na[i] + 1
    H, _ = MatG(pnz, p, m, dp)                                                          # :89
    uG = deltaUFree(B, En, p, dnz)                                                      # :92
    Hp = cell2mat2(uG)                                                                  # :93
    duM = (nb + dnz).max(axis=0)                                                        # :94
    S1 = H.T @ Q @ H + W                                                                # :98-100
    S1 = (S1 + S1.T) / 2
    K = np.linalg.solve(S1, H.T @ Q)
    rows = [0] + [int(np.sum(m[:i + 1])) for i in range(nu - 1)]                        # :102-105
    Km = K[rows, :]
    return dict(H=H, S=S, Hp=Hp, Km=Km, K=K, na=na, nb=nb, duM=duM, dmin=dmin, dnz=dnz, dp=dp, A=A, B=B, F=F, En=En)


def dtc_gpc_closed_loop(prob, p, m, delta, lam, fr, k_start=4):
    """DTC_GPC_WW.m:110-164.  prob: object with pnz (scaled discrete model), preal (unscaled discrete
    process), pq (unscaled discrete disturbance model), L, R, r (ny x nit), q (nq x nit).
    fr: list of (Nr, Dr) per output.  Returns y (ny x nit), u (nu x nit)."""
    st = dtc_gpc_setup(prob.pnz, p, m, delta, lam)
    ny, nu = prob.pnz.a.shape
    nit = prob.r.shape[1]
    na, duM, dmin = st["na"], st["duM"], st["dmin"]
    Hp, S, Km = st["Hp"], st["S"], st["Km"]
    y = np.zeros((ny, nit)); ye = np.zeros((ny, nit)); u = np.zeros((nu, nit)); ue = np.zeros((nu, nit))
    re = prob.L[:, None] * prob.r
    up = np.zeros(int(duM.sum()))
    plant, dist = _Chan(prob.preal), _Chan(prob.pq)
    model, fast = _Chan(prob.pnz), _Chan(prob.pnz, dshift=dmin)
    ypz = np.zeros((ny, nit)); ygz = np.zeros((ny, nit)); eM = np.zeros((ny, nit)); yfr = np.zeros((ny, nit))
    yp = np.zeros((ny, nit))
    for k in range(nit):                                    # 0-based sample index; MATLAB k = this + 1
        y[:, k] = plant.output_at(k, u) + dist.output_at(k, prob.q)       # :131-132
        ye[:, k] = prob.L * y[:, k]                                        # :133
        ypz[:, k] = model.output_at(k, ue)                                 # OptimalPredictor2.m:28
        ygz[:, k] = fast.output_at(k, ue)                                  # :31
        eM[:, k] = ye[:, k] - ypz[:, k]                                    # :34
        for i in range(ny):                                                # :37  lsim(Fr, eM)
            Nr, Dr = fr[i]
            acc = 0.0
            nN, nD = len(Nr), len(Dr)
            off = nD - nN                                                  # Nr(z)/Dr(z): relative degree
            for l in range(nN):
                if k - off - l >= 0:
                    acc += Nr[l] * eM[i, k - off - l]
            for l in range(1, nD):
                if k - l >= 0:
                    acc -= Dr[l] * yfr[i, k - l]
            yfr[i, k] = acc / Dr[0]
        yp[:, k] = ygz[:, k] + yfr[:, k]                                   # :40
        if k + 1 < k_start:
            continue
        Yd = np.concatenate([[yp[j, k - l] if k - l >= 0 else 0.0 for l in range(na[j] + 1)] for j in range(ny)])  # :139-142
        Ref = np.concatenate([np.full(int(p[i]), re[i, k]) for i in range(ny)])                                      # :143-145
        yf = Hp @ up + S @ Yd                                                                                         # :146
        dU = Km @ (Ref - yf)                                                                                          # :149
        off = 0
        for i in range(nu):                                                                                           # :152-155
            blk = up[off:off + int(duM[i])].copy()
            up[off:off + int(duM[i])] = np.concatenate([[dU[i]], blk[:-1]])
            off += int(duM[i])
        ue[:, k] = ue[:, k - 1] + dU if k > 0 else dU                                                                 # :158-162
        u[:, k] = prob.R * ue[:, k]                                                                                   # :163
    return y, u
