"""DTC-GPC path (BASELINE.json configs[3]) on the CPU: the oracle's restatement of the reference's MATLAB
(oracle/dtc_gpc_oracle.py) is pinned by the identities the reference's algebra must satisfy -- the
reference stores no numeric output for this path -- and the product's host-side polynomial tables
(libmpcgpu.so, mpc_dtc_tables.cpp, no CUDA call) are compared with it entry by entry."""
import ctypes as C

import numpy as np
import pytest

import mpcgpu
from mpcgpu import _capi
from mpcgpu.dtcgpc import DtcProblemStruct, woodberry_dtc, mimo_filter, robustness_filter
from oracle import dtc_gpc_oracle as dorc


@pytest.fixture(scope="module")
def prob():
    return woodberry_dtc()


def test_diophantine_identity():
    """1 = E_j(z^-1) A~(z^-1) + z^-j F_j(z^-1)  (diophantine.m header) for every row j."""
    rng = np.random.default_rng(0)
    for na in (1, 2, 3):
        A = np.poly(rng.uniform(0.3, 0.98, size=na))
        N, d = 12, 2
        E, F = dorc.diophantine(A, N, d)
        AD = np.convolve(A, [1.0, -1.0])
        for row in range(N):
            j = d + 1 + row
            lhs = np.convolve(E[row, :j], AD)
            lhs = np.concatenate([lhs, np.zeros(max(0, j + F.shape[1] - len(lhs)))])
            lhs[j:j + F.shape[1]] += F[row]
            ref = np.zeros_like(lhs); ref[0] = 1.0
            assert np.abs(lhs - ref).max() < 1e-12


def test_prediction_consistency(prob):
    """Hp*up + S*Yd must reproduce the free evolution of the polynomial model B_i/A_i (this ties BA_MIMO,
    diophantineMIMO, deltaUFree and cell2mat2 together, DTC_GPC_WW.m:79-95,146) and G*du the forced response
    of the fast model (MatG.m, DTC_GPC_WW.m:89)."""
    rng = np.random.default_rng(1)
    p = np.array([7, 5]); m = np.array([3, 2])
    st = dorc.dtc_gpc_setup(prob.pnz, p, m, np.ones(2), np.ones(2))
    ny, nu = 2, 2
    T, k = 80, 40
    A, B, dnz, duM = st["A"], st["B"], st["dnz"], st["duM"]

    def poly_model(ue):
        y = np.zeros((ny, T))
        for i in range(ny):
            for t in range(T):
                acc = 0.0
                for l in range(1, len(A[i])):
                    if t - l >= 0:
                        acc -= A[i][l] * y[i, t - l]
                for j in range(nu):
                    for l in range(len(B[i][j])):
                        tt = t - 1 - dnz[i, j] - l
                        if tt >= 0:
                            acc += B[i][j][l] * ue[j, tt]
                y[i, t] = acc
        return y

    # (a) free response: moves up to k-1, none afterwards
    du = rng.normal(size=(nu, T)) * 0.1
    du[:, k:] = 0.0
    ue = np.cumsum(du, axis=1)
    y = poly_model(ue)
    # the rounded-pole common denominator (BA_MIMO.m:38-40) keeps the polynomial model within O(1e-4) of the channels
    fast = dorc._Chan(prob.pnz, dshift=st["dmin"])
    yex = np.array([fast.output_at(t, ue) for t in range(T)]).T
    assert np.abs(yex - y).max() < 5e-3
    up = np.concatenate([[du[j, k - 1 - t] for t in range(int(duM[j]))] for j in range(nu)])
    Yd = np.concatenate([[y[i, k - l] for l in range(st["na"][i] + 1)] for i in range(ny)])
    pred = st["Hp"] @ up + st["S"] @ Yd
    truth = np.concatenate([y[i, k + 1:k + 1 + p[i]] for i in range(ny)])
    assert np.abs(pred - truth).max() < 1e-10
    # (b) forced response: moves at k .. k+m-1 only, exact fast model from rest
    du2 = np.zeros((nu, T))
    for j in range(nu):
        du2[j, k:k + m[j]] = rng.normal(size=m[j])
    ue2 = np.cumsum(du2, axis=1)
    fast2 = dorc._Chan(prob.pnz, dshift=st["dmin"])
    y2 = np.array([fast2.output_at(t, ue2) for t in range(T)]).T
    duf = np.concatenate([du2[j, k:k + m[j]] for j in range(nu)])
    truth2 = np.concatenate([y2[i, k + 1:k + 1 + p[i]] for i in range(ny)])
    assert np.abs(st["H"] @ duf - truth2).max() < 1e-12


def test_matg_is_the_step_response(prob):
    p = np.array([6, 4]); m = np.array([3, 3])
    _, _, dp = dorc.descompMPC(prob.pnz)
    H, blocks = dorc.MatG(prob.pnz, p, m, dp)
    dmin = dp.min(axis=1)
    for i in range(2):
        for j in range(2):
            s = dorc.step_response(prob.pnz, i, j, 40)
            G = blocks[i][j]
            for r in range(p[i]):
                for k in range(m[j]):
                    assert G[r, k] == (s[dmin[i] + 1 + r - k] if r >= k else 0.0)
    assert H.shape == (p.sum(), m.sum())


def test_unconstrained_gain_matches_normal_equations(prob):
    p = np.array([9, 8]); m = np.array([4, 3]); delta = np.array([0.7, 2.0]); lam = np.array([0.3, 1.5])
    st = dorc.dtc_gpc_setup(prob.pnz, p, m, delta, lam)
    H = st["H"]
    Q = np.diag(np.repeat(delta, p)); W = np.diag(np.repeat(lam, m))
    e = np.random.default_rng(3).normal(size=p.sum())
    z = st["K"] @ e
    grad = H.T @ Q @ (H @ z - e) + W @ z          # optimality of min (Hz-e)'Q(Hz-e) + z'Wz
    assert np.abs(grad).max() < 1e-10


@pytest.mark.parametrize("alfa,raio", [(0.7, 0.8), (0.9, 0.95), (0.5, 0.96), (0.6, 0.7)])
def test_filter_design(prob, alfa, raio):
    fr_o = dorc.mimofilter_Fr(prob.pnz, alfa, raio)
    fr_p = mimo_filter(prob.pnz, alfa, raio)
    for i, ((No, Do), (Np, Dp)) in enumerate(zip(fr_o, fr_p)):
        assert len(No) == len(Np) and len(Do) == len(Dp)
        assert np.allclose(No, Np, rtol=1e-9, atol=1e-12) and np.allclose(Do, Dp, rtol=0, atol=1e-15)
        assert abs(np.sum(Np) / np.sum(Dp) - 1.0) < 1e-9      # unit static gain (mimofilter.m:53-58)
        d = int(np.min(prob.pnz.d[i]))
        for pz in prob.pnz.a[i]:
            if abs(pz) >= raio:                                # slow poles are cancelled: Dr z^d - Nr = 0 there
                assert abs(np.polyval(Dp, pz) * pz ** d - np.polyval(Np, pz)) < 1e-10


def _host_tables(prob):
    lib = _capi.load_library()
    ny, nu = prob.pnz.a.shape
    nq = prob.pq.a.shape[1]
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32))
    keep = dict(ma=f64(prob.pnz.a), mb0=f64(prob.pnz.b0), mb1=f64(prob.pnz.b1), md=i32(prob.pnz.d),
                pa=f64(prob.preal.a), pb0=f64(prob.preal.b0), pb1=f64(prob.preal.b1), pd=i32(prob.preal.d),
                qa=f64(prob.pq.a), qb0=f64(prob.pq.b0), qb1=f64(prob.pq.b1), qd=i32(prob.pq.d),
                L=f64(prob.L), R=f64(prob.R), r=f64(prob.r), q=f64(prob.q))
    ps = DtcProblemStruct(ny, nu, nq, prob.nit, prob.pmax, prob.mmax, prob.k_start, 0)
    for k, a in keep.items():
        setattr(ps, k, a.ctypes.data)
    info = np.zeros(64, dtype=np.int32)
    assert lib.mpcgpu_dtc_host_tables(C.byref(ps), info.ctypes.data_as(C.c_void_p), None, None, None) == 0
    step_len, pmax, MAXNA, MAXCP = info[:4]
    step = np.zeros((ny, nu, step_len)); ftab = np.zeros((ny, pmax + 1, MAXNA)); ug = np.zeros((ny, nu, pmax + 1, MAXCP))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.mpcgpu_dtc_host_tables(C.byref(ps), ptr(info), ptr(step), ptr(ftab), ptr(ug)) == 0
    return info, step, ftab, ug


@pytest.mark.parametrize("mismatch", [False, True])
def test_product_tables_match_oracle(mismatch):
    prob = woodberry_dtc(deltak=0.1, deltaL=1.0) if mismatch else woodberry_dtc()
    info, step, ftab, ug = _host_tables(prob)
    ny, nu = 2, 2
    P = prob.pmax
    pvec = np.full(ny, P)
    st = dorc.dtc_gpc_setup(prob.pnz, pvec, np.full(nu, 2), np.ones(ny), np.ones(nu))
    assert list(info[8:8 + ny]) == list(st["na"]) and list(info[16:16 + ny]) == list(st["dmin"])
    assert list(info[48:48 + nu]) == list(st["duM"])
    for i in range(ny):
        F = st["F"][i]
        assert np.abs(ftab[i, 1:P + 1, :F.shape[1]] - F).max() < 1e-13
        for j in range(nu):
            s = dorc.step_response(prob.pnz, i, j, step.shape[2] - 1)
            assert np.abs(step[i, j] - s).max() == 0.0
    uG = dorc.deltaUFree(st["B"], st["En"], pvec, st["dnz"])
    for i in range(ny):
        for j in range(nu):
            blk = uG[i][j]
            assert info[24 + i * nu + j] == blk.shape[1]
            assert np.abs(ug[i, j, 1:P + 1, :blk.shape[1]] - blk).max() < 1e-13


def test_reference_script_run_is_sane(prob):
    """DTC_GPC_WW.m defaults (p = m = 3, delta = lambda = 1, Fr(0.7, 0.8)): nothing moves before the first
    set-point step / controller start, and -- run long enough -- the loop is offset-free on both outputs
    despite the load disturbance entering at k = 140 (integral action of the CARIMA model)."""
    import dataclasses
    nit = 900
    r = np.zeros((2, nit)); r[:, :prob.nit] = prob.r; r[:, prob.nit:] = prob.r[:, -1:]
    q = np.zeros((1, nit)); q[:, :prob.nit] = prob.q; q[:, prob.nit:] = prob.q[:, -1:]
    long = dataclasses.replace(prob, nit=nit, r=r, q=q)
    fr = dorc.mimofilter_Fr(prob.pnz, 0.7, 0.8)
    y, u = dorc.dtc_gpc_closed_loop(long, np.array([3, 3]), np.array([3, 3]), np.ones(2), np.ones(2), fr)
    assert np.abs(y[:, :10]).max() == 0.0 and np.abs(u[:, :3]).max() == 0.0
    assert np.abs(y).max() < 2.0
    assert np.abs(y[:, -1] - r[:, -1]).max() < 1e-3
    # the first 200 samples are the reference's own run
    y200, u200 = dorc.dtc_gpc_closed_loop(prob, np.array([3, 3]), np.array([3, 3]), np.ones(2), np.ones(2), fr)
    assert np.array_equal(y200, y[:, :200]) and np.array_equal(u200, u[:, :200])
