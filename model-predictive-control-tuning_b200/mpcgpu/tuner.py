"""Batched optimiser drivers on top of the evaluator.

The reference's optimisers are serial by construction: `VNS2.m:89-283` flips one bit pattern at a time and restarts
on the first improvement, `fgoalattain` (`MPC_TFob.m:66-67`) evaluates one finite-difference point at a time.  Fed
that way a GPU evaluates one candidate per call.  These drivers keep the reference's *objectives* and search
spaces and change only the evaluation order so that every step is one population:

  vns_search      the whole order-k neighbourhood (all k-bit flips of the horizon bit vectors, k = 1..3) is one
                  batch; the best improving neighbour is accepted and the search restarts at order 1
                  (best-improvement instead of the reference's first-improvement: same neighbourhoods, same
                  legality rule VNS2.m:135, same objective VNS2.m:147-195).
  goal_attain     the goal-attainment problem of MPC_TFob.m:66-67 / MPCTuning.m:88-91,
                  min gamma s.t. F_i(x) - w_i*gamma <= goal_i, x = [delta lambda] >= lb, solved by a batched
                  derivative-free contraction search in log-weights (one population per iteration) instead of
                  fgoalattain's SQP with serial finite differences.
  mpc_tfob        the outer alternation of MPC_TFob.m:56-132 (weights <-> horizons until the weights stop improving).

`evaluate_vns(N, Nu) -> F` and `evaluate_gam(X) -> g (n x ny)` are callables, so the drivers run against any
evaluator (the GPU `Evaluator`, the NMPC one, or a test double).  Host-side control logic only; all arithmetic
of the hot path stays in libmpcgpu.so.
"""
from __future__ import annotations

from itertools import combinations

import numpy as np


def _bits(value: int, nbits: int) -> np.ndarray:
    """MSB-first bit vector (MPCTuning.m:285-289: flip(de2bi(.)))."""
    return np.array([(value >> (nbits - 1 - b)) & 1 for b in range(nbits)], dtype=np.int64)


def _value(bits: np.ndarray) -> int:
    return int(sum(int(b) << (len(bits) - 1 - i) for i, b in enumerate(bits)))


def legal(N: int, Nu: np.ndarray, dmin) -> bool:
    """VNS2.m:135 with PreCon.m:23: min(N) > max(Nu), all non-zero, N above every minimum dead time, Nu > 1."""
    return N > int(np.max(Nu)) and N > int(np.max(dmin)) and int(np.min(Nu)) > 1


def neighbourhood(N: int, Nu, nbp: int, nbc: int, order: int, dmin):
    """All legal candidates at Hamming distance `order` in the N bit vector (tt = 1, VNS2.m:98-101) or in one input's
    Nu bit vector (tt = 2, :102-104).  Returns (N_list, Nu_matrix)."""
    Nu = np.asarray(Nu, dtype=np.int64)
    outN, outNu = [], []
    bn = _bits(N, nbp)
    for idx in combinations(range(nbp), order):
        b = bn.copy(); b[list(idx)] ^= 1
        n2 = _value(b)
        if legal(n2, Nu, dmin):
            outN.append(n2); outNu.append(Nu.copy())
    for h in range(len(Nu)):
        bu = _bits(int(Nu[h]), nbc)
        for idx in combinations(range(nbc), order):
            b = bu.copy(); b[list(idx)] ^= 1
            nu2 = Nu.copy(); nu2[h] = _value(b)
            if legal(N, nu2, dmin):
                outN.append(N); outNu.append(nu2)
    return np.array(outN, dtype=np.int64), np.array(outNu, dtype=np.int64).reshape(len(outN), len(Nu))


def vns_search(evaluate_vns, N0: int, Nu0, nbp: int, nbc: int, dmin, max_order: int = 3, max_rounds: int = 200, log=None):
    """Batched variable-neighbourhood search over the horizons.  evaluate_vns(N[n], Nu[n, nu]) -> F[n]
    (the evaluator collapses Nu to its max, closedloop_toolbox.m:38-40).  Returns (N, Nu, F, evaluations)."""
    N, Nu = int(N0), np.asarray(Nu0, dtype=np.int64).copy()
    F = float(np.asarray(evaluate_vns(np.array([N]), Nu[None, :]))[0])
    evals, order = 1, 1
    for _ in range(max_rounds):
        if order > max_order:
            break
        cn, cnu = neighbourhood(N, Nu, nbp, nbc, order, dmin)
        if len(cn) == 0:
            order += 1
            continue
        Fc = np.asarray(evaluate_vns(cn, cnu), dtype=float)
        evals += len(cn)
        Fc = np.where(np.isfinite(Fc), Fc, np.inf)
        best = int(np.argmin(Fc))
        if Fc[best] < F:
            N, Nu, F = int(cn[best]), cnu[best].copy(), float(Fc[best])
            if log:
                log(f"Fvns={F:.6g}; N=[{N}]; Nu=[{int(Nu.max())}]")        # VNS2.m:200
            order = 1
        else:
            order += 1
    return N, Nu, F, evals


def attainment(g: np.ndarray, goal, w) -> np.ndarray:
    """Goal-attainment factor gamma(x) = max_i (F_i - goal_i) / w_i (fgoalattain's merit)."""
    g = np.atleast_2d(g)
    return np.max((g - np.asarray(goal, float)) / np.asarray(w, float), axis=1)


def goal_attain(evaluate_gam, x0, w, goal=1e-3, lb=1e-5, ub=1e3, pop: int = 256, iters: int = 12, sigma0: float = 1.0,
                seed: int = 0, frozen=None, log=None):
    """min gamma s.t. F_i(x) - w_i gamma <= goal (MPC_TFob.m:66-67, options MPCTuning.m:88-91), x >= lb
    (MPCTuning.m:302).  One population of `pop` log-normal perturbations of the incumbent per iteration, the step
    size contracts when an iteration does not improve.  `frozen`: boolean mask of entries kept at x0 (band outputs
    keep delta = 0, GAM_fun.m:62-66).  Returns (x, gamma, g(x), evaluations)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.maximum(np.abs(np.asarray(x0, float)), 0.0)
    frozen = np.zeros(len(x), dtype=bool) if frozen is None else np.asarray(frozen, bool)
    free = ~frozen
    x[free] = np.clip(x[free], lb, ub)
    g = np.atleast_2d(evaluate_gam(x[None, :]))[0]
    gam = float(attainment(g, goal, w)[0])
    sigma, evals = float(sigma0), 1
    for it in range(iters):
        Z = rng.normal(size=(pop, len(x))) * sigma
        Z[:, frozen] = 0.0
        Z[0, :] = 0.0
        # half of the population moves one coordinate only: resolves directions the full perturbations blur
        half = pop // 2
        keep = rng.integers(0, max(int(free.sum()), 1), size=half)
        fidx = np.where(free)[0]
        mask = np.zeros((half, len(x)), dtype=bool)
        if len(fidx):
            mask[np.arange(half), fidx[keep]] = True
        Z[pop - half:] *= mask
        X = x[None, :] * np.exp(Z)
        X[:, free] = np.clip(X[:, free], lb, ub)
        G = np.atleast_2d(evaluate_gam(X))
        evals += pop
        gm = attainment(G, goal, w)
        gm = np.where(np.isfinite(gm), gm, np.inf)
        best = int(np.argmin(gm))
        if gm[best] < gam:
            x, g, gam = X[best].copy(), G[best].copy(), float(gm[best])
            if log:
                log(f"iter {it}: attainment factor {gam:.6g}, Fgam={np.sum(g):.6g}")   # MPC_TFob.m:73-77
        else:
            sigma *= 0.5
        if sigma < 1e-3:
            break
    return x, gam, g, evals


def mpc_tfob(evaluate_gam_at, evaluate_vns_at, ny: int, nu: int, N0: int, Nu0, delta0, lambda0, w, nbp: int, nbc: int, dmin,
             goal=1e-3, band_mask=None, max_outer: int = 6, log=None, **ga_kw):
    """MPC_TFob.m:56-132: alternate the weight search (GAM) and the horizon search (VNS) until the weight search
    stops improving.  evaluate_gam_at(N, Nu)(X) -> g ;  evaluate_vns_at(delta, lam)(N, Nu) -> F.
    Returns dict(N, Nu, delta, lam, Fgam, Fvns, evaluations)."""
    N, Nu = int(N0), np.asarray(Nu0, dtype=np.int64).copy()
    x = np.concatenate([np.asarray(delta0, float), np.asarray(lambda0, float)])
    frozen = np.zeros(ny + nu, dtype=bool)
    if band_mask is not None:
        frozen[:ny] = np.asarray(band_mask, bool)
    Fgam_best, Fvns, evals = np.inf, np.inf, 0
    best = None
    for outer in range(max_outer):
        x2, gam, g, e1 = goal_attain(evaluate_gam_at(N, int(Nu.max())), x, w, goal=goal, frozen=frozen, log=log, seed=outer, **ga_kw)
        evals += e1
        Fgam = round(float(np.sum(g)), 2)                                         # MPC_TFob.m:104
        improved = Fgam < Fgam_best
        if improved:
            x, Fgam_best = x2, Fgam                                               # :106-110
        N, Nu, Fvns, e2 = vns_search(evaluate_vns_at(x[:ny], x[ny:]), N, Nu, nbp, nbc, dmin, log=log)   # :118
        evals += e2
        best = dict(N=N, Nu=Nu.copy(), delta=x[:ny].copy(), lam=x[ny:].copy(), Fgam=Fgam_best, Fvns=Fvns, evaluations=evals)
        if not improved:                                                          # :128-130
            break
    return best


def tune_linear(ev, w, N0=None, Nu0=2, delta0=None, lambda0=None, log=None, **kw):
    """MPCTuning(mpcobj, Sp, true, w, nit, Yref, mdv, nbp, nbc) for a linear `Evaluator` (MPCTuning.m:283-302 start
    point: N = 2^nbp - 1, Nu = 2, delta = lambda = 1)."""
    p = ev.prob
    ny, nu = p.ny, p.nu
    N0 = 2 ** int(p.nbp) - 1 if N0 is None else int(N0)
    Nu0 = np.full(nu, int(Nu0), dtype=np.int64)
    delta0 = np.where(p.band_mask, 0.0, 1.0) if delta0 is None else np.asarray(delta0, float)
    lambda0 = np.ones(nu) if lambda0 is None else np.asarray(lambda0, float)

    def gam_at(N, Nu):
        def f(X):
            X = np.atleast_2d(X)
            n = X.shape[0]
            out = ev.eval_batch(np.full(n, N, np.int32), np.full(n, Nu, np.int32), X[:, :ny], X[:, ny:], mode="gam")
            return out["cost"]
        return f

    def vns_at(delta, lam):
        def f(Nc, Nuc):
            n = len(Nc)
            out = ev.eval_batch(np.asarray(Nc, np.int32), np.asarray(Nuc, np.int64).max(axis=1).astype(np.int32),
                                np.broadcast_to(delta, (n, ny)), np.broadcast_to(lam, (n, nu)), mode="vns")
            return out["cost"]
        return f

    return mpc_tfob(gam_at, vns_at, ny, nu, N0, Nu0, delta0, lambda0, w, int(p.nbp), int(p.nbc), p.dmin,
                    band_mask=p.band_mask, log=log, **kw)
