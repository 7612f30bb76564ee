"""DTC-GPC (dead-time-compensated GPC) case data and host API: BASELINE.json configs[3].

`woodberry_dtc()` restates the problem set-up of DTC-GPC/DTC_GPC_WW.m:17-64,110-124 as data.  The
robustness-filter coefficients are an INPUT to the path (filter design is one-off polynomial algebra,
SURVEY.md §2 row 6); the batched sweep itself runs in libmpcgpu.so (`DtcEvaluator`)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .plant import Channels, c2d_fopdt
from .problems import WOODBERRY_L, WOODBERRY_R


@dataclass
class DtcProblem:
    name: str
    Ts: float
    nit: int
    pnz: Channels      # conditioned discrete nominal model L*Pn*R          (DTC_GPC_WW.m:35-41)
    preal: Channels    # discrete process P (unscaled; = Pn in the nominal case, :23-25)
    pq: Channels       # discrete disturbance model Pq (unscaled, :31-32)
    L: np.ndarray
    R: np.ndarray
    r: np.ndarray      # ny x nit reference (unscaled, :117-119)
    q: np.ndarray      # nq x nit disturbance (:123-124)
    k_start: int = 4   # first sample (1-based) at which the controller acts (:128)
    pmax: int = 30
    mmax: int = 10


def woodberry_dtc(deltak: float = 0.0, deltaL: float = 0.0, L=None, R=None) -> DtcProblem:
    Ts, nit = 1.0, 200                                                                   # :19, :54
    K = np.array([[12.8, -18.9], [6.6, -19.4]])                                          # :27-28
    tau = np.array([[16.7, 21.0], [10.9, 14.4]])
    theta = np.array([[1.0, 2.0], [2.0, 1.0]])
    L = WOODBERRY_L if L is None else np.asarray(L, float)                               # :35-36 CondMin(Kn): an input
    R = (WOODBERRY_R[:2] if R is None else np.asarray(R, float))
    pn = c2d_fopdt(K, tau, theta, Ts)
    pnz = pn.scaled(L, R)                                                                # :37, :41
    preal = c2d_fopdt(K * (1 + deltak), tau, theta + deltaL, Ts)                         # :23-25
    pq = c2d_fopdt(np.array([[3.8], [4.9]]), np.array([[14.9], [13.2]]), np.array([[8.1], [3.4]]), Ts)  # :31-32
    r = np.zeros((2, nit)); r[0, 10:] = 0.8; r[1, 60:] = 0.5                             # :117-119 (1-based 11:, 61:)
    q = np.zeros((1, nit)); q[0, 140:] = -0.25                                           # :123-124
    return DtcProblem("WoodBerry-DTC-GPC", Ts, nit, pnz, preal, pq, L, R, r, q)


def synthetic_dtc_population(prob: DtcProblem, n: int, seed: int = 0):
    """p_i ~ U{1..30}, m_j ~ U{1..min(p,10)}, delta, lambda log-uniform [1e-3, 1e2],
    filter alfa ~ U[0.5, 0.95], raio ~ U[0.7, 0.95].  Returns p (n x ny), m (n x nu), delta, lam, alfa, raio."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ny, nu = prob.pnz.a.shape
    p = rng.integers(1, prob.pmax + 1, size=(n, ny)).astype(np.int32)
    m = np.stack([rng.integers(1, np.minimum(p.min(axis=1), prob.mmax) + 1) for _ in range(nu)], axis=1).astype(np.int32)
    delta = np.exp(rng.uniform(np.log(1e-3), np.log(1e2), size=(n, ny)))
    lam = np.exp(rng.uniform(np.log(1e-3), np.log(1e2), size=(n, nu)))
    alfa = rng.uniform(0.5, 0.95, size=n)
    raio = rng.uniform(0.7, 0.95, size=n)
    return p, m, delta, lam, alfa, raio


# ---------------------------------------------------------------------------------------------
# robustness filter Fr (an input of the sweep) and the evaluator on libmpcgpu.so
# ---------------------------------------------------------------------------------------------
MAXF = 8   # MPCGPU_DTC_MAXF


def robustness_filter(poles, d: int, alfa: float, raio: float):
    """Fr(z) = Nr(z)/Dr(z) of DTC-GPC/filtro_siso.m:26-96 for a delay-free model with the given poles and
    dead time d >= 1 samples.  The reference's Sylvester system states Dr(z) z^d = Nr(z) + px(z) Q(z) with
    px = (z-1) * prod(z - unwanted poles) and deg Nr < deg px, so Nr is the remainder of the polynomial
    division of Dr z^d by px -- computed here as such (unit static gain and cancellation of the slow poles
    follow from Nr = Dr z^d at the roots of px).  d == 0 makes the reference's system under-determined
    (its `A\\B` picks an arbitrary basic solution), so it is refused."""
    slow = [float(pz) for pz in poles if abs(pz) >= raio]
    if not slow:
        return np.array([1.0]), np.array([1.0])                     # filtro_siso.m:90-91
    if d < 1:
        raise ValueError("robustness_filter: dead time of at least one sample required (see docstring)")
    Dr = np.poly([alfa] * len(slow))
    px = np.poly([1.0] + slow)
    _, rem = np.polydiv(np.concatenate([Dr, np.zeros(d)]), px)
    Nr = np.zeros(len(px) - 1)
    Nr[len(Nr) - len(rem):] = rem
    return Nr, Dr


def mimo_filter(pnz: Channels, alfa: float, raio: float):
    """Diagonal Fr of DTC-GPC/mimofilter.m:33-50: output i uses the product of its non-zero channels with
    the row's minimum dead time (`Pd.iodelay`, :25-29)."""
    ny, nu = pnz.a.shape
    out = []
    for i in range(ny):
        poles = [pnz.a[i, j] for j in range(nu) if (pnz.b0[i, j] + pnz.b1[i, j]) != 0]
        out.append(robustness_filter(poles, int(np.min(pnz.d[i, :])), alfa, raio) if poles else (np.array([1.0]), np.array([1.0])))
    return out


class DtcProblemStruct(C.Structure):
    """mpcgpu_dtc_problem (include/mpcgpu.h)."""
    _fields_ = [("ny", C.c_int32), ("nu", C.c_int32), ("nq", C.c_int32), ("nit", C.c_int32),
                ("pmax", C.c_int32), ("mmax", C.c_int32), ("k_start", C.c_int32), ("reserved", C.c_int32)] + \
               [(k, C.c_void_p) for k in ("ma", "mb0", "mb1", "md", "pa", "pb0", "pb1", "pd", "qa", "qb0", "qb1", "qd",
                                          "L", "R", "r", "q")]


class DtcEvaluator:
    """One `mpcgpu_dtc_handle`: the batched form of DTC_GPC_WW.m:56-164 on one B200 (no CPU fallback)."""

    def __init__(self, prob: DtcProblem, device: int = -1):
        from . import _capi
        from .api import MpcGpuError
        self._err = MpcGpuError
        self.lib = _capi.load_library()
        self.prob = prob
        self.ny, self.nu = prob.pnz.a.shape
        self.nq = prob.pq.a.shape[1]
        self.nit = int(prob.nit)
        f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32))
        self._keep = dict(ma=f64(prob.pnz.a), mb0=f64(prob.pnz.b0), mb1=f64(prob.pnz.b1), md=i32(prob.pnz.d),
                          pa=f64(prob.preal.a), pb0=f64(prob.preal.b0), pb1=f64(prob.preal.b1), pd=i32(prob.preal.d),
                          qa=f64(prob.pq.a), qb0=f64(prob.pq.b0), qb1=f64(prob.pq.b1), qd=i32(prob.pq.d),
                          L=f64(prob.L), R=f64(prob.R), r=f64(prob.r).reshape(self.ny, self.nit),
                          q=f64(prob.q).reshape(self.nq, self.nit))
        ps = DtcProblemStruct(self.ny, self.nu, self.nq, self.nit, int(prob.pmax), int(prob.mmax), int(prob.k_start), 0)
        for k, arr in self._keep.items():
            setattr(ps, k, arr.ctypes.data if arr.size else None)
        h = C.c_void_p()
        rc = self.lib.mpcgpu_dtc_create(C.byref(ps), int(device), C.byref(h))
        if rc != 0:
            raise MpcGpuError(f"mpcgpu_dtc_create failed ({rc}): {self.lib.mpcgpu_dtc_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mpcgpu_dtc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counters(self) -> dict:
        from . import _capi
        c = _capi.Counters()
        self.lib.mpcgpu_dtc_get_counters(self.h, C.byref(c))
        return c.asdict()

    def pack_filters(self, filters):
        """filters: per candidate a list (per output) of (Nr, Dr) -> fr_num, fr_den (n x ny x MAXF), fr_len."""
        n = len(filters)
        num = np.zeros((n, self.ny, MAXF)); den = np.zeros((n, self.ny, MAXF)); ln = np.zeros((n, self.ny, 2), dtype=np.int32)
        for c, fr in enumerate(filters):
            for i, (Nr, Dr) in enumerate(fr):
                num[c, i, :len(Nr)] = Nr; den[c, i, :len(Dr)] = Dr
                ln[c, i] = (len(Nr), len(Dr))
        return num, den, ln

    def eval_batch(self, p, m, delta, lam, alfa=None, raio=None, filters=None, traj=False, design_on_device=True):
        """p: n x ny, m: n x nu, delta: n x ny, lam: n x nu; either (alfa, raio) per candidate -- the robustness filter is then
        designed per candidate ON THE DEVICE (mpcgpu_dtc_eval_batch_design: mimofilter.m / filtro_siso.m as a batched op;
        design_on_device=False: on the host with `mimo_filter`) -- or explicit `filters`.  Returns dict(ise, status[, y, u])."""
        p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.int32); n = p.shape[0]
        m = np.ascontiguousarray(np.atleast_2d(m), dtype=np.int32)
        delta = np.ascontiguousarray(delta, dtype=np.float64).reshape(n, self.ny)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(n, self.nu)
        if filters is None and design_on_device:
            al = np.ascontiguousarray(np.broadcast_to(np.asarray(alfa, float), (n,)))
            ra = np.ascontiguousarray(np.broadcast_to(np.asarray(raio, float), (n,)))
            ise = np.empty((n, self.ny)); status = np.zeros(n, dtype=np.int32)
            y = np.empty((n, self.ny, self.nit)) if traj else None
            u = np.empty((n, self.nu, self.nit)) if traj else None
            ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
            rc = self.lib.mpcgpu_dtc_eval_batch_design(self.h, n, ptr(p), ptr(m), ptr(delta), ptr(lam), ptr(al), ptr(ra),
                                                       ptr(ise), ptr(y), ptr(u), ptr(status))
            if rc != 0:
                raise self._err(f"mpcgpu_dtc_eval_batch_design failed ({rc}): {self.lib.mpcgpu_dtc_last_error(self.h).decode()}")
            out = {"ise": ise, "status": status}
            if traj:
                out.update(y=y, u=u)
            return out
        if filters is None:
            # the design depends on (alfa, raio) only through which poles count as slow: design once per distinct pair
            alfa = np.broadcast_to(np.asarray(alfa, float), (n,)); raio = np.broadcast_to(np.asarray(raio, float), (n,))
            cache = {}
            filters = []
            for a, rr in zip(alfa.tolist(), raio.tolist()):
                key = (a, rr)
                if key not in cache:
                    cache[key] = mimo_filter(self.prob.pnz, a, rr)
                filters.append(cache[key])
        num, den, ln = self.pack_filters(filters)
        ise = np.empty((n, self.ny)); status = np.zeros(n, dtype=np.int32)
        y = np.empty((n, self.ny, self.nit)) if traj else None
        u = np.empty((n, self.nu, self.nit)) if traj else None
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        rc = self.lib.mpcgpu_dtc_eval_batch(self.h, n, ptr(p), ptr(m), ptr(delta), ptr(lam), ptr(num), ptr(den), ptr(ln),
                                            ptr(ise), ptr(y), ptr(u), ptr(status))
        if rc != 0:
            raise self._err(f"mpcgpu_dtc_eval_batch failed ({rc}): {self.lib.mpcgpu_dtc_last_error(self.h).decode()}")
        out = {"ise": ise, "status": status}
        if traj:
            out.update(y=y, u=u)
        return out


def dtc_gpc_ww(prob: DtcProblem, p=(3, 3), m=(3, 3), delta=(1.0, 1.0), lam=(1.0, 1.0), alfa=0.7, raio=0.8, ev=None):
    """The reference script's run (DTC_GPC_WW.m:56-164, defaults :56-64 and :108): returns (y, u), signals x time."""
    ev = ev or DtcEvaluator(prob)
    out = ev.eval_batch([p], [m], [delta], [lam], alfa=alfa, raio=raio, traj=True)
    if out["status"][0] != 0:
        raise ev._err(f"DTC-GPC run failed with status {int(out['status'][0])}")
    return out["y"][0], out["u"][0]
