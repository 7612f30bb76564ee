// mpc_sim.cuh -- the closed-loop kernel body: one warp runs one closed-loop simulation
// (closedloop_toolbox.m:50-100 fused with the GAM / VNS cost sums), SIMT style.
//
// Row layout.  The decision vector is padded input-major: row r = j*P + c (input j, horizon index c),
// P = 4, 8 or 16 >= m, R = NU*P rows, rows with c >= m are inert (zero rows of M and W, infinite bounds).
// Lane l owns row l (slot 0) and row 32+l (slot 1, only when R = 48/64); everything indexed by row --
// the QP iterate z, its bounds, W n, the step direction, the active-constraint mask -- lives in that
// lane's REGISTERS.  The cumulative sums u(k+c) = u(k-1) + sum_{c'<=c} dz are segmented warp scans
// (width P) instead of loops.  Shared memory holds M (read every sample), the physical plant state with
// circular input histories, and the active-set factor:
//     Li : inverse Cholesky factor of the Schur complement S = N'WN (W = H^-1, N = active normals)
//     V  : V = W N Li'  (R x q).  With it   z = z_unc + V (Li g)   and the Goldfarb-Idnani step direction
//          W (n_p - N r) = W n_p - V (Li N'W n_p)   cost O(R q) from shared memory instead of an R^2 pass
//          over W in L2; the new column when p is added is dir/sqrt(rho), i.e. free.
// Only the first QC active constraints keep V/Li in shared memory, the rest spill to a per-run global
// scratch (rare: > 16 simultaneously active constraints).
//
// The algorithm (warm-started dual active set, tolerances, deviation coordinates) is the one validated
// against the oracle; see DESIGN.md.  This header is also compiled by tests/host_emulation (thread-per-
// lane emulation of the warp intrinsics) so that the not-gpu tests run the same source.
#pragma once
#include <math.h>

#include "mpc_layout.h"

#define SIM_VIOL_TOL 1e-10
#define SIM_DEP_TOL 1e-13
#ifndef SIM_SHIFT_DEN
#define SIM_SHIFT_DEN 2   /* guided shifted restart when more than 1/SIM_SHIFT_DEN of the carried multipliers are negative */
#endif
#define SIM_MISC_INTS 22  /* sm.misc: [2] scratch of the shifted restart, [3..20] state of the two-phase mode */
#define SIM_PP_ON 3
#define SIM_PP_QS 9       /* [9], [10]: active-set sizes of the two parked factors */
#define SIM_PP_CUR 4      /* slot that belongs to the carried factor */
#define SIM_PP_DIRTY 11   /* the carried factor differs from its slot */
#define SIM_PP_NCON 12    /* constrained-sample counter at the last bookkeeping call */
#define SIM_PP_STREAK 13  /* consecutive constrained samples with a bookkeeping call */
#define SIM_PP_FIRST 14   /* the next re-entry is the first of the mode */
#define SIM_PP_BAD 15     /* consecutive expensive re-entries */
#define SIM_PP_MATCH 16   /* the previous bookkeeping call saw the A B A pattern */
#define SIM_PP_BACK 17    /* back-off length after failed episodes */
#define SIM_PP_WAIT 18    /* constrained-sample count before which the mode is not entered again */
#define SIM_PP_GOOD 19    /* this episode had a cheap re-entry */
#define SIM_PP_EVERGOOD 20 /* some episode of this run had one */
#ifndef SIM_PP_AVG
#define SIM_PP_AVG 6      /* mean iterations per constrained sample from which a run is eligible for the mode */
#endif
#define SIM_PP_SIG1 5
#define SIM_PP_SIG2 6
#define SIM_PP_COUNT 7   /* exchanges of this run (diagnostics) */
#define SIM_PP_REENTERED 8
#ifndef SIM_PP_BADIT
#define SIM_PP_BADIT 4    /* a re-entry that needs this many iterations is no better than the carried set */
#endif
#define SIM_PP_GOODIT 2   /* a re-entry that needs at most this many paid off */
#ifndef SIM_PP_DOUBLE_MATCH
#define SIM_PP_DOUBLE_MATCH 0   /* 1: enter on A B A B instead of A B A */
#endif
#define SIM_PP_STREAK_MIN 3 /* bookkeeping calls in a row (= expensive solves at consecutive constrained samples) before the mode may be entered, counted from 0 */
#define SIM_PP_HEAVY 6    /* active-set iterations in one QP from which the warm start counts as torn down */
#define SIM_CHURN 24   /* iterations per constrained QP above which a run switches its pivot rule */
#define SIM_REFRESH 96 /* Givens removals after which the factor of the soft-constraint kernel (mpc_soft.cuh) is refreshed */
#ifndef SIM_MB
#define SIM_MB 10 /* columns of M in flight per batch when M is read from global memory (P = 16); even */
#endif
#ifndef SIM_CH
#define SIM_CH 32 /* samples of r / yref / v staged into shared memory at a time (power of two) */
#endif
#define SIM_INF (__builtin_huge_val())
#define SIM_FULL 0xffffffffu
#ifdef MPC_SIMT_EMULATION   /* experiment switches, host emulation only */
extern int g_sim_knob;
extern long long g_sim_reappends, g_sim_rotations;
#define SIM_KNOB(b) (g_sim_knob & (b))
extern int g_sim_verbose;
#include <cstdio>
#define SIM_DBG(...) do { if (g_sim_verbose && lane == 0) { printf(__VA_ARGS__); } } while (0)
#define SIM_DBGSET(tag) do { if (g_sim_verbose && lane == 0) { printf("%s q=%d:", tag, q); for (int a_ = 0; a_ < q; ++a_) printf(" %c%d.%d", "dDuU"[sm.act[a_] & 3], (sm.act[a_] >> 2) / P, (sm.act[a_] >> 2) % P); printf("\n"); } } while (0)
#else
#define SIM_KNOB(b) 0
#define SIM_DBG(...) do { } while (0)
#define SIM_DBGSET(tag) do { } while (0)
#endif

struct MpcRunOut {
    double *cost;                  // GAM: ny slots; VNS: 1 slot (partial sum of this run); RAW: nullptr
    double *y, *u, *ys, *uopt;     // trajectories of this candidate (signals x nit) or nullptr
    unsigned long long *counters;  // [0] constrained QPs, [1] active-set iterations
    unsigned long long *diag;      // optional per-run {constrained QPs, iterations, max active set, clock cycles}
    int *trace;                    // optional per-sample {active-set iterations, final active-set size} (debug)
};

static MPC_HD int sim_pad(int m) { return m <= 4 ? 4 : (m <= 8 ? 8 : 16); }
// active constraints whose factor rows live in shared memory; the rest spills to global memory
#ifndef SIM_QC_MAX
#define SIM_QC_MAX 24
#endif
static MPC_HD int sim_qc(int R) { return R <= SIM_QC_MAX ? R : SIM_QC_MAX; }
// the kernel without a spill area (mpc_sim_spec.cuh) keeps up to SIM_SPEC_QC factor rows, Li packed triangular
#ifndef SIM_SPEC_QC
#define SIM_SPEC_QC 32
#endif
static MPC_HD int sim_spec_qc(int R) { return R <= SIM_SPEC_QC ? R : SIM_SPEC_QC; }
static MPC_HD size_t sim_spec_vli_doubles(int R) { const int qc = sim_spec_qc(R); return (size_t)qc * R + (size_t)qc * (qc + 1) / 2; }
// per-run parking slots of the two-phase mode for that kernel
static MPC_HD size_t sim_spec_slot_doubles(int R) { return 2 * (sim_spec_vli_doubles(R) + 32); }
// M (nst x R) lives in shared memory for the small buckets; for P = 16 it stays in global memory (read-only,
// L1/L2 resident, loads independent of the state) so that twice as many runs fit on an SM.
#ifndef SIM_M_GLOBAL_P
#define SIM_M_GLOBAL_P 16
#endif
static MPC_HD bool sim_m_in_smem(int P) { return P < SIM_M_GLOBAL_P; }
static MPC_HD int sim_hl(const MpcLayout &L) {
    int h = 1;
    for (int j = 0; j < L.nw; ++j) h = L.hlen[j] > h ? L.hlen[j] : h;
    return h + 1;
}
// per-run global scratch (doubles) for V columns / Li rows beyond QC
static MPC_HD size_t sim_scratch_doubles(int R) {
    const int qc = sim_qc(R);
    return 2 * (size_t)(R - qc) * R;   // V columns and (full-length) Li rows beyond QC
}
// per-run parking slot of the two-phase mode (SimWarp::exchange): V columns (QC x R), Li rows (QC x QC), the
// constraint ids (QC ints), the per-lane constraint masks.  Kept in a region of its own: appending it to the spill
// area changed that area's stride to ~32 KB and cost 4 % on its own (gpurun_out/ab21.log).
static MPC_HD size_t sim_slot_doubles(int R) {
    const int qc = sim_qc(R);
    return 2 * ((size_t)qc * R + (size_t)qc * qc + 32);   // two slots, see sim_pp_exchange
}
static MPC_HD size_t sim_smem_doubles(const MpcLayout &L, int nu, int P) {
    const int R = nu * P, qc = sim_qc(R), nch = L.ny * L.nw, HL = sim_hl(L);
    size_t n = sim_m_in_smem(P) ? (size_t)L.nst * R : 0;   // M
    n += (L.nst + 1) & ~1;                    // st
    n += 2 * nch;                             // x, xol
    n += (size_t)L.nw * HL;                   // hist
    n += 4 * nch;                             // cha, chb0, chb1, chg
    n += (size_t)SIM_CH * (2 * L.ny + L.nd);  // sig
    n += (size_t)nu * P;                      // uopt
    n += 4 * nu;                              // bnd
    n += 4 * (size_t)R;                       // z, lvl, w, wsc
    n += 4 * (size_t)R;                       // g, l, rr, mu
    n += (size_t)qc * R + (size_t)qc * qc;    // V, Li (full stride)
    n += (2 * nch + L.nst + 2 * R + SIM_MISC_INTS + 1) / 2 + 1;   // ints: chd, chj, role, act, dflag, misc
    return n;
}

struct SimSm {
    double *M, *st, *x, *xol, *hist, *cha, *chb0, *chb1, *chg, *sig, *uopt, *bnd, *z, *lvl, *w, *wsc, *g, *l, *rr, *mu, *V, *Li;
    int *chd, *chj, *role, *act, *dflag, *misc;
};

// The warp reductions and the segmented scan (-DSIM_OOL_REDUCE: out of line, 1100 SASS instructions less, but slower).
#if defined(MPC_SIMT_EMULATION)
#define SIM_OOL static inline
#elif defined(SIM_OOL_REDUCE)
#define SIM_OOL static __device__ __noinline__
#else
#define SIM_OOL static __device__ __forceinline__   /* measured: out of line costs 9 % (call latency, lost interleaving of the row slots) */
#endif
SIM_OOL double sim_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SIM_FULL, v, o);
    return v;
}
SIM_OOL double sim_wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(SIM_FULL, v, o));
    return v;
}
// all lanes end with the minimum value and, among equal values, the smallest non-negative index.
// Three warp-wide integer reductions (redux.sync) on an order-preserving 64-bit key instead of a five-step shuffle
// butterfly on (double, int) pairs: this was the single hottest line of the active-set path (profiles/r1f).
struct SimArgMin { double v; int i; };
SIM_OOL SimArgMin sim_wargmin_ool(double v, int i) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long key = (b >> 63) ? ~b : (b | 0x8000000000000000ull);   // unsigned order == numeric order
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(SIM_FULL, hi);
    const unsigned mlo = __reduce_min_sync(SIM_FULL, hi == mhi ? lo : 0xffffffffu);
    const bool mine = hi == mhi && lo == mlo;
    const unsigned mi = __reduce_min_sync(SIM_FULL, (mine && i >= 0) ? (unsigned)i : 0xffffffffu);
    const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
    const unsigned long long mb = (mk >> 63) ? (mk & 0x7fffffffffffffffull) : ~mk;
    SimArgMin r;
    r.v = __longlong_as_double((long long)mb);
    r.i = mi == 0xffffffffu ? -1 : (int)mi;
    return r;
}
__device__ __forceinline__ void sim_wargmin(double &v, int &i) {
    const SimArgMin r = sim_wargmin_ool(v, i);
    v = r.v; i = r.i;
}
// inclusive scan of v over the horizon index c inside each P-wide input segment
template <int P>
SIM_OOL double sim_segscan(double v, int c) {
#pragma unroll
    for (int off = 1; off < P; off <<= 1) {
        const double t = __shfl_up_sync(SIM_FULL, v, off, P);
        if (c >= off) v += t;
    }
    return v;
}

template <int NU>
__device__ __forceinline__ double sim_pick(const double (&u)[NU], int j) {
    double r = u[0];
#pragma unroll
    for (int k = 1; k < NU; ++k) r = (j == k) ? u[k] : r;
    return r;
}

// ---- two-phase mode (period-2 limit cycles) ----------------------------------------------------------------
// A tuning that limit-cycles with period 2 (rate limits hit in alternating directions) has TWO active sets, one per
// phase; the set carried from the previous sample is then always the wrong one and is torn down and rebuilt
// constraint by constraint, 8-17 iterations per sample for 491 samples (the heaviest runs of a 32768-candidate
// population).  The factor (V, Li) depends on the set only, not on the sample's data, so the set of two samples
// ago is parked in global memory and re-entered: a warm start that is usually optimal as it stands.
// Both routines are cold and deliberately NOT inlined, with their state in shared memory (sm.misc) and the masks
// passed packed in a register: the closed-loop kernel is instruction-cache and register bound,
// anything added to its hot body costs every run.
#ifdef MPC_SIMT_EMULATION
#define SIM_COLD static
#else
#define SIM_COLD static __device__ __noinline__
#endif
// Re-enter the factor of the other phase.  Each phase has a slot in global memory: V (QC x R) followed by Li (QC x QC)
// as in shared memory (vli), then QC constraint ids and 32 packed per-lane masks.  The carried factor is written to
// its slot only if the last solve changed it (in a settled cycle both factors are read-only: loads that stay in L2,
// no stores to evict the M tables of the neighbouring runs).  Returns the new active-set size in bits [0,8) and this
// lane's constraint masks (4 bits per row slot) from bit 8.
SIM_COLD int sim_pp_exchange(double *vli, int *act, int *misc, double *slot, int lane, int q, int masks, int R, int QC, int packed) {
    // Li: QC x QC full stride, or packed lower triangular (rows [0,q) are the first q(q+1)/2 doubles)
    const size_t LS = packed ? (size_t)QC * (QC + 1) / 2 : (size_t)QC * QC;
    const size_t S = (size_t)QC * R + LS + 32;   // doubles per slot (the int part: QC + 32 ints <= 32 doubles)
    const int cur = misc[SIM_PP_CUR], dirty = misc[SIM_PP_DIRTY];
    double *sc = slot + (size_t)cur * S, *so = slot + (size_t)(1 - cur) * S;
    int *ic = (int *)(sc + (size_t)QC * R + LS), *io = (int *)(so + (size_t)QC * R + LS);
    const int first = misc[SIM_PP_FIRST];   // first sample of the mode: park the carried factor, keep it as the warm start
    const int qb = first ? 0 : misc[SIM_PP_QS + 1 - cur];
    if (dirty) {
        const int nv = q * R, n = nv + (packed ? q * (q + 1) / 2 : q * QC);
#pragma unroll 1
        for (int j = lane; j < n; j += 32) {
            const int idx = j < nv ? j : QC * R + (j - nv);
            sc[idx] = vli[idx];
        }
        if (lane < QC) ic[lane] = act[lane];
        ic[QC + lane] = masks;
    }
    __syncwarp();
    {   // rows [0, qb) of V and of Li as ONE loop over a virtual index, four independent global loads in flight per
        // pass (a dependent load per element made the re-entry cost ten active-set iterations)
        const int nv = qb * R, n = nv + (packed ? qb * (qb + 1) / 2 : qb * QC);
#pragma unroll 1
        for (int j0 = lane; j0 < n; j0 += 128) {
            double g[4];
            int idx[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + 32 * e < n ? j0 + 32 * e : j0;
                idx[e] = j < nv ? j : QC * R + (j - nv);
                g[e] = so[idx[e]];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (j0 + 32 * e < n) vli[idx[e]] = g[e];
        }
    }
    if (lane < qb) act[lane] = io[lane];
    const int theirs = qb > 0 ? io[QC + lane] : 0;   // an empty slot holds nothing yet
    if (lane == 0) {
        misc[SIM_PP_QS + cur] = q; misc[SIM_PP_CUR] = 1 - cur; misc[SIM_PP_DIRTY] = first;
        misc[SIM_PP_REENTERED] = qb > 0; misc[SIM_PP_COUNT] += qb > 0; misc[SIM_PP_FIRST] = 0;
    }
    __syncwarp();
    return first ? (q | (masks << 8)) : (qb | (theirs << 8));
}
// End of a constrained solve that took `it` iterations and ended on `q` constraints: an expensive solve that ends on
// the set of two samples ago, and not on the previous one, is a period-2 cycle -- start parking.  Leave the mode
// again when a re-entered set turns out as expensive as a torn-down one.
SIM_COLD void sim_pp_update(const int *act, int *misc, int lane, int q, int it, int QC, int ncon, int nit_run) {
    unsigned hsh = 0u;   // order-independent signature of the final active set
    for (int a = lane; a < q; a += 32) hsh ^= ((unsigned)act[a] + 1u) * 2654435761u;
    const int sig = (int)(__reduce_xor_sync(SIM_FULL, hsh) ^ ((unsigned)q << 26));
    const int s1 = misc[SIM_PP_SIG1], s2 = misc[SIM_PP_SIG2], on = misc[SIM_PP_ON], reentered = misc[SIM_PP_REENTERED];
    const int last = misc[SIM_PP_NCON], streak = misc[SIM_PP_STREAK];   // consecutive constrained samples seen here
    __syncwarp();
    if (lane == 0) {
        const int st = (ncon == last + 1) ? streak + 1 : 0;
        misc[SIM_PP_NCON] = ncon; misc[SIM_PP_STREAK] = st;
        misc[SIM_PP_SIG2] = s1; misc[SIM_PP_SIG1] = sig;
        const int match = st >= SIM_PP_STREAK_MIN && sig == s2 && sig != s1;   // A B A at the end of a row of expensive solves
        const int match_prev = misc[SIM_PP_MATCH];
        misc[SIM_PP_MATCH] = match;
        if (on) {
            if (it > 0) misc[SIM_PP_DIRTY] = 1;
            if (reentered) {   // a re-entered set that is as expensive as a torn-down one: twice in a row, or very
                const int bad = it >= SIM_PP_BADIT ? misc[SIM_PP_BAD] + 1 : 0;
                misc[SIM_PP_BAD] = bad;
                if (it <= SIM_PP_GOODIT) { misc[SIM_PP_GOOD] = 1; misc[SIM_PP_EVERGOOD] = 1; }
                if (bad >= 2 || it >= 2 * SIM_PP_HEAVY) {
                    // leave; an episode without a single cheap re-entry was a false alarm: wait 16, 32, .. 256
                    // constrained samples before the next one
                    const int back0 = misc[SIM_PP_BACK];
                    const int back = misc[SIM_PP_GOOD] ? 0 : (back0 ? (back0 < 256 ? 2 * back0 : 256) : 16);
                    misc[SIM_PP_ON] = 0; misc[SIM_PP_BACK] = back; misc[SIM_PP_WAIT] = ncon + back;
                }
            }
        } else if (match && (match_prev || !SIM_PP_DOUBLE_MATCH) && ncon >= misc[SIM_PP_WAIT] && q <= QC && !SIM_KNOB(64) &&
                   (misc[SIM_PP_EVERGOOD] || nit_run >= SIM_PP_AVG * ncon)) {
            // ... and only for a run that churns as a whole (SIM_PP_AVG iterations per constrained sample so far) or that
            // the mode has already paid off for: moderately busy runs gain nothing and a chaotic one can lose
            misc[SIM_PP_ON] = 1; misc[SIM_PP_FIRST] = 1; misc[SIM_PP_CUR] = 0; misc[SIM_PP_DIRTY] = 1; misc[SIM_PP_BAD] = 0; misc[SIM_PP_GOOD] = 0;
        }
    }
    __syncwarp();
}

// Removal of the active constraint at position a (see SimWarp::remove_at) for the kernel WITHOUT a spill area
// (every factor row in shared memory, q <= QC <= 32): one out-of-line copy shared by both call sites -- the closed-loop
// kernel is bound by instruction fetch, the routine is ~500 SASS instructions when inlined.
template <int R, int QC, int NSLOT>   // Li packed lower triangular: row i starts at i (i + 1) / 2
SIM_COLD void sim_remove_nospill(double *V, double *Li, int *act, double *mu, double *g, double *rr, int lane, int a, int q) {
    static_assert(QC <= 32, "one row slot of Li columns");
    const int nrot = q - 1 - a;
    if (nrot > 0) {
        for (int t = lane; t < nrot; t += 32) {
            double ss = 0.0;
            for (int i = a; i <= a + t; ++i) { const double v = Li[i * (i + 1) / 2 + a]; ss = fma(v, v, ss); }
            const double y = Li[(a + t + 1) * (a + t + 2) / 2 + a];
            const double inv = 1.0 / sqrt(fma(y, y, ss));
            g[t] = sqrt(ss) * inv;
            rr[t] = y * inv;
        }
        __syncwarp();
        double cv[NSLOT], cy0 = 0.0;
        const int j0 = lane;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) cv[s] = lane + 32 * s < R ? V[(size_t)a * R + lane + 32 * s] : 0.0;
        if (j0 < a) cy0 = Li[a * (a + 1) / 2 + j0];
#pragma unroll 1
        for (int t = 0; t < nrot; ++t) {
            const int k = a + t;
            const double cx = g[t], sy = rr[t];
            __syncwarp();   // row k was read by the previous step
            const double *vsrc = V + (size_t)(k + 1) * R;
            double *vdst = V + (size_t)k * R;
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                if (lane + 32 * s < R) {
                    const double o = vsrc[lane + 32 * s];
                    vdst[lane + 32 * s] = cx * o - sy * cv[s];
                    cv[s] = cx * cv[s] + sy * o;
                }
            const double *ysrc = Li + (k + 1) * (k + 2) / 2;
            double *ydst = Li + k * (k + 1) / 2;
            if (j0 != a && j0 <= k + 1) {
                const double o = ysrc[j0];
                ydst[j0 < a ? j0 : j0 - 1] = cx * o - sy * cy0;
                cy0 = cx * cy0 + sy * o;
            }
        }
        __syncwarp();
        const int pos = a + 1 + lane;
        const int ca = pos < q ? act[pos] : 0;
        const double cm = pos < q ? mu[pos] : 0.0;
        __syncwarp();
        if (pos < q) { act[pos - 1] = ca; mu[pos - 1] = cm; }
    }
    __syncwarp();
}

// Everything one warp needs to carry through the run; template so that R, P, NSLOT are constants.
// SPILL: factor rows beyond QC live in a per-run global scratch; without it a QP that needs more than QC active
// constraints returns status SIM_ST_OVERFLOW and the host re-runs the candidate on the kernel with the spill area.
#define SIM_ST_OVERFLOW 6
// MSM: M lives in shared memory (plain loads) instead of global memory (read-only path)
template <int NU, int P, bool SPILL = true, bool MSM = false>
struct SimWarp {
    static constexpr int R = NU * P;
    static constexpr int NSLOT = (R + 31) / 32;
    static constexpr int QCMAX = SPILL ? SIM_QC_MAX : SIM_SPEC_QC;
    static constexpr int QC = (R <= QCMAX) ? R : QCMAX;

    const MpcLayout &L;
    SimSm sm;
    double *gscr;   // per-run global scratch (V columns / Li rows beyond QC), may be nullptr when QC == R
    const double *W;
    const double *Mp;   // M: shared memory (P < 16) or global
    int lane, m;
    // per-lane row data
    int row[NSLOT];
    bool valid[NSLOT];
    double dlo[NSLOT], dhi[NSLOT], ulo[NSLOT], uhi[NSLOT];
    double z[NSLOT];
    int amask[NSLOT];
    double u[NU];   // uniform: MV levels u(k-1)
    int q;          // uniform: carried active-set size
    unsigned long long n_con, n_it;
    int qmax;
    int churn;      // uniform: 1 once the run has switched to most-violated-first pivoting (qp_solve)
    // period-2 limit cycles (bang-bang tunings): the factor of the OTHER phase is parked in global memory behind the
    // spill area (exchange()); the mode's state lives in shared memory (sm.misc + SIM_PP_*), not in registers
    double *slot;   // uniform: this run's parking slot (global memory), nullptr: mode off

    __device__ __forceinline__ SimWarp(const MpcLayout &L_) : L(L_) {}

    __device__ __forceinline__ double *Vcol(int a) const { return (!SPILL || a < QC) ? sm.V + (size_t)a * R : gscr + (size_t)(a - QC) * R; }
    // row a of the inverse Cholesky factor: entries [0..a]; QC doubles long in shared memory, R in the spill
    __device__ __forceinline__ double *Lirow(int a) const {
        if (!SPILL) return sm.Li + a * (a + 1) / 2;   // packed lower triangular
        return a < QC ? sm.Li + (size_t)a * QC : gscr + (size_t)(R - QC) * R + (size_t)(a - QC) * R;
    }

    // u(k+c) for this lane's rows: u_j(k-1) + inclusive scan of z over c within the input's segment
    __device__ __forceinline__ void levels(double (&lv)[NSLOT]) const {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            const double v = sim_segscan<P>(z[s], row[s] & (P - 1));
            lv[s] = sim_pick<NU>(u, (row[s] / P) < NU ? (row[s] / P) : 0) + v;
        }
    }
    // the same for any plan zz and MV levels uu (block verification of mpc_sim_spec.cuh)
    __device__ __forceinline__ void levels_of(const double (&zz)[NSLOT], const double (&uu)[NU], double (&lv)[NSLOT]) const {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            const double v = sim_segscan<P>(zz[s], row[s] & (P - 1));
            lv[s] = sim_pick<NU>(uu, (row[s] / P) < NU ? (row[s] / P) : 0) + v;
        }
    }
    // does the plan zz with levels lv violate a rate or level limit on this lane's rows?
    __device__ __forceinline__ int infeasible(const double (&zz)[NSLOT], const double (&lv)[NSLOT]) const {
        int bad = 0;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if (valid[s])
                bad |= (zz[s] - dlo[s] < -SIM_VIOL_TOL) | (dhi[s] - zz[s] < -SIM_VIOL_TOL) |
                       (lv[s] - ulo[s] < -SIM_VIOL_TOL) | (uhi[s] - lv[s] < -SIM_VIOL_TOL);
        return bad;
    }
    __device__ __forceinline__ void publish(const double (&lv)[NSLOT]) const {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if (row[s] < R) { sm.z[row[s]] = z[s]; sm.lvl[row[s]] = lv[s]; }
        __syncwarp();
    }
    // slack of constraint cid from the published z / lvl (any lane)
    __device__ __forceinline__ double slack_of(int cid) const {
        const int type = cid & 3, r = cid >> 2, j = r / P;
        const double *b = sm.bnd + 4 * j;  // dumin, dumax, umin, umax
        switch (type) {
            case 0: return sm.z[r] - b[0];
            case 1: return b[1] - sm.z[r];
            case 2: return sm.lvl[r] - b[2];
            default: return b[3] - sm.lvl[r];
        }
    }
    // n_cid' x where x and its per-input inclusive scan xs have been published (x: rate normals, xs: level)
    __device__ __forceinline__ double ndot(int cid, const double *x, const double *xs) const {
        const int type = cid & 3, r = cid >> 2;
        const double v = type < 2 ? x[r] : xs[r];
        return (type & 1) ? -v : v;
    }
    // wv = W n_cid for this lane's rows.  Wg holds 2R rows of R: rows [0,R) are W, rows [R,2R) the running
    // sums of W's rows over the horizon index inside each input block, i.e. W times a level normal: one
    // coalesced load per row whatever the constraint.  Publishes wv and its segmented scan to dst / dsts.
    __device__ __forceinline__ void w_times_normal(int cid, double (&wv)[NSLOT], double *dst, double *dsts) const {
        const int type = cid & 3, r = cid >> 2;
        const double sg = (type & 1) ? -1.0 : 1.0;
        const double *src = W + (size_t)(type < 2 ? r : R + r) * R;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) wv[s] = row[s] < R ? sg * src[row[s]] : 0.0;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            const double v = sim_segscan<P>(wv[s], row[s] & (P - 1));
            if (row[s] < R) { dst[row[s]] = wv[s]; dsts[row[s]] = v; }
        }
        __syncwarp();
    }
    // g = N' x (x, xs published), l = Li g, rr = Li' l for the first qq active constraints; returns |l|^2
    __device__ __forceinline__ double schur_vectors(int qq, const double *x, const double *xs) const {
        for (int a = lane; a < qq; a += 32) sm.g[a] = ndot(sm.act[a], x, xs);
        __syncwarp();
        const double l2 = tri_lower(qq);
        tri_upper(qq, sm.rr);
        return l2;
    }
    // l = Li g (lane a owns row a); returns |l|^2
    __device__ __forceinline__ double tri_lower(int qq) const {
        double part = 0.0;
        for (int a = lane; a < qq; a += 32) {
            const double *row_a = Lirow(a);
            double a0 = 0.0, a1 = 0.0;
            int b = 0;
#pragma unroll 1
            for (; b + 1 <= a; b += 2) { a0 = fma(row_a[b], sm.g[b], a0); a1 = fma(row_a[b + 1], sm.g[b + 1], a1); }
            if (b <= a) a0 = fma(row_a[b], sm.g[b], a0);
            const double acc = a0 + a1;
            sm.l[a] = acc;
            part += acc * acc;
        }
        const double l2 = sim_wsum(part);
        __syncwarp();
        return l2;
    }
    // out = Li' l (lane a owns column a)
    __device__ __forceinline__ void tri_upper(int qq, double *out) const {
        const int qs = qq < QC ? qq : QC;
        for (int a = lane; a < qq; a += 32) {
            double a0 = 0.0, a1 = 0.0;
            int b = a;
#pragma unroll 1
            for (; b + 1 < qs; b += 2) { a0 = fma(Lirow(b)[a], sm.l[b], a0); a1 = fma(Lirow(b + 1)[a], sm.l[b + 1], a1); }
            if (b < qs) { a0 = fma(Lirow(b)[a], sm.l[b], a0); ++b; }
            if (b < QC) b = QC;
            if (SPILL) {
#pragma unroll 1
                for (; b < qq; ++b) a0 = fma(Lirow(b)[a], sm.l[b], a0);   // spilled rows (rare)
            }
            out[a] = a0 + a1;
        }
        __syncwarp();
    }
    // l = Li g, mu_out = Li' l  (g already in sm.g)
    __device__ __forceinline__ void schur_solve(int qq, double *mu_out) const {
        tri_lower(qq);
        tri_upper(qq, mu_out);
    }
    // x_rows += sign * sum_a coef[a] * V_a   (per-lane rows)
    __device__ __forceinline__ void add_V(int qq, const double *coef, double sign, double (&x)[NSLOT]) const {
        const int qs = qq < QC ? qq : QC;
        double acc[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) acc[s] = 0.0;
#pragma unroll 4
        for (int a = 0; a < qs; ++a) {
            const double ca = coef[a];
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                if (row[s] < R) acc[s] = fma(ca, sm.V[(size_t)a * R + row[s]], acc[s]);
        }
        if (SPILL) {
#pragma unroll 1
            for (int a = QC; a < qq; ++a) {   // spilled columns (rare)
                const double ca = coef[a];
                const double *va = Vcol(a);
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (row[s] < R) acc[s] = fma(ca, va[row[s]], acc[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) x[s] += sign * acc[s];
    }
    // commit constraint cid at position a: V_a = dir / sqrt(rho), Li row a = [-rr'/sqrt(rho), 1/sqrt(rho)]
    __device__ __forceinline__ void commit(int cid, int a, double rho, double mu_new, const double (&dir)[NSLOT]) {
        const double isr = 1.0 / sqrt(rho);
        double *va = Vcol(a), *lr = Lirow(a);
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if (row[s] < R) va[row[s]] = dir[s] * isr;
        for (int b = lane; b < a; b += 32) lr[b] = -sm.rr[b] * isr;
        if (lane == 0) { lr[a] = isr; sm.act[a] = cid; sm.mu[a] = mu_new; }
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if (row[s] == (cid >> 2)) amask[s] |= (1 << (cid & 3));
        __syncwarp();
    }
    // Remove the active constraint at list position a (order of the rest kept) by Givens rotations.
    // With Y = Li (S^-1 = Y'Y, Y lower triangular) and V = W N Y': rotating rows (k, k+1) of Y for
    // k = a .. q-2 so that column a is annihilated in row k pushes that column's mass into the last row;
    // every other row then has a zero in column a, i.e. it is a valid factor row of the reduced set once
    // column a is deleted (columns right of a move one to the left, which also restores the triangle).
    // The carried entry of column a at step k is the running norm sqrt(sum_{i=a..k} Y[i][a]^2), so all
    // rotations come from one column and do not depend on each other; V's columns and Y's rows take the
    // same rotations.  O((q-a)(R+q)) work, no pass over W.
    __device__ __forceinline__ void remove_at(int a) {
        const int nrot = q - 1 - a;
        const int cid = sm.act[a];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if (row[s] == (cid >> 2)) amask[s] &= ~(1 << (cid & 3));
        if (!SPILL) {
            sim_remove_nospill<R, QC, NSLOT>(sm.V, sm.Li, sm.act, sm.mu, sm.g, sm.rr, lane, a, q);
            q -= 1;
#ifdef MPC_SIMT_EMULATION
            if (lane == 0) g_sim_rotations += nrot;
#endif
            return;
        }
        if (nrot > 0) {
            // rotation coefficients (cx, cy) = (x, y) / hypot(x, y), kept in sm.g / sm.rr
            for (int t = lane; t < nrot; t += 32) {
                double ss = 0.0;
                for (int i = a; i <= a + t; ++i) { const double v = Lirow(i)[a]; ss = fma(v, v, ss); }
                const double y = Lirow(a + t + 1)[a];
                const double inv = 1.0 / sqrt(fma(y, y, ss));
                sm.g[t] = sqrt(ss) * inv;
                sm.rr[t] = y * inv;
            }
            __syncwarp();
            // V columns (lane owns rows) and Y columns (lane owns old column j, written to j or j-1)
            double cv[NSLOT], cy0 = 0.0, cy1 = 0.0;
            const int j0 = lane, j1 = lane + 32;
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) cv[s] = row[s] < R ? Vcol(a)[row[s]] : 0.0;
            if (j0 < a) cy0 = Lirow(a)[j0];
            if (j1 < a) cy1 = Lirow(a)[j1];
            for (int t = 0; t < nrot; ++t) {
                const int k = a + t;
                const double cx = sm.g[t], sy = sm.rr[t];
                __syncwarp();   // row k was read by the previous step
                const double *vsrc = Vcol(k + 1);
                double *vdst = Vcol(k);
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (row[s] < R) {
                        const double o = vsrc[row[s]];
                        vdst[row[s]] = cx * o - sy * cv[s];
                        cv[s] = cx * cv[s] + sy * o;
                    }
                const double *ysrc = Lirow(k + 1);
                double *ydst = Lirow(k);
                if (j0 != a && j0 <= k + 1) {
                    const double o = ysrc[j0];
                    ydst[j0 < a ? j0 : j0 - 1] = cx * o - sy * cy0;
                    cy0 = cx * cy0 + sy * o;
                }
                if (j1 <= k + 1) {   // j1 >= 32 > a is not required: guard both ways
                    if (j1 != a) {
                        const double o = ysrc[j1];
                        ydst[j1 < a ? j1 : j1 - 1] = cx * o - sy * cy1;
                        cy1 = cx * cy1 + sy * o;
                    }
                }
            }
            __syncwarp();
            // compact the constraint list and its multipliers
            int ca[2]; double cm[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int pos = a + 1 + lane + 32 * s;
                ca[s] = pos < q ? sm.act[pos] : 0;
                cm[s] = pos < q ? sm.mu[pos] : 0.0;
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int pos = a + 1 + lane + 32 * s;
                if (pos < q) { sm.act[pos - 1] = ca[s]; sm.mu[pos - 1] = cm[s]; }
            }
        }
        __syncwarp();
        q -= 1;
#ifdef MPC_SIMT_EMULATION
        if (lane == 0) g_sim_rotations += nrot;
#endif
    }
    // remove every active constraint flagged in sm.dflag (highest position first, so that the positions of
    // the ones still to go do not move)
    __device__ __forceinline__ void drop_flagged() {
        for (int a = q - 1; a >= 0; --a) {
            const int fl = sm.dflag[a];
            __syncwarp();
            if (fl) remove_at(a);
        }
    }
    // Dual active-set QP, warm-started from the carried set as it is.  Shifting EVERY carried set by one sample to
    // follow the receding horizon was measured to cost more than it saves; the shifted set is tried first only when
    // most of the carried one comes back with negative multipliers (guided shifted restart below), and a run whose set
    // alternates with period 2 re-enters the parked factor of the other phase (sim_pp_exchange).
    // z (registers) in: z_unc, out: optimum.
    __device__ __forceinline__ int qp_solve() {
        int it = 0;
        const int itmax = 20 * (NU * m + 10);
        double lv[NSLOT];
        SIM_DBGSET("entry");
        // Earliest-horizon-first pivoting wins on the population as a whole, but a few very aggressive tunings churn
        // under it (60-120 add/drop iterations per QP where most-violated-first needs ~35): a run whose running
        // mean exceeds SIM_CHURN iterations per constrained QP pivots on the most violated constraint instead.
        if (!churn && n_con > 8 && n_it > (unsigned long long)SIM_CHURN * n_con) churn = 1;   // sticky for the rest of the run
        const bool most_violated = SIM_KNOB(2) || (!SIM_KNOB(16) && churn);
        if (SIM_KNOB(1)) {
            q = 0;
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) amask[s] = 0;
        }
        if (sm.misc[SIM_PP_ON]) {   // two-phase mode: re-enter the factor of two samples ago (first sample of the
            // mode: the slot is empty, the carried factor is parked and this sample starts cold -- its carried set
            // would have been torn down anyway)
            if (q > QC) { __syncwarp(); if (lane == 0) sm.misc[SIM_PP_ON] = 0; }
            if (q <= QC) {
                int masks = 0;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) masks |= amask[s] << (4 * s);
                const int r_ = sim_pp_exchange(sm.V, sm.act, sm.misc, slot, lane, q, masks, R, QC, SPILL ? 0 : 1);
                q = r_ & 255;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) amask[s] = (r_ >> (8 + 4 * s)) & 15;
            }
            __syncwarp();
        }
        // ---- warm start on the carried set: mu = S^-1 (b_A - N_A' z_unc), shed negative multipliers ----
        int npref = 0, pptr = 0;   // uniform: shifted constraints to be tried first (stored in sm.dflag)
        bool first_pass = true;
        while (q > 0) {
            levels(lv);
            publish(lv);
            for (int a = lane; a < q; a += 32) sm.g[a] = -slack_of(sm.act[a]);
            __syncwarp();
            schur_solve(q, sm.mu);
            double mumax = 0.0;
            for (int a = lane; a < q; a += 32) mumax = fmax(mumax, fabs(sm.mu[a]));
            mumax = sim_wmax(mumax);
            int nd_ = 0;
            if (SIM_KNOB(8)) {   // experiment: shed only the most negative multiplier per round
                double mv = 0.0; int mi = -1;
                for (int a = lane; a < q; a += 32) if (sm.mu[a] < mv) { mv = sm.mu[a]; mi = a; }
                sim_wargmin(mv, mi);
                for (int a = lane; a < q; a += 32) sm.dflag[a] = (a == mi) && (mv < -1e-12 * mumax);
                nd_ = mi >= 0 && (mv < -1e-12 * mumax);
            } else {
                int nneg = 0;
                for (int a0 = 0; a0 < q; a0 += 32) {
                    const int a = a0 + lane;
                    const int fl = a < q && sm.mu[a] < -1e-12 * mumax;
                    if (a < q) sm.dflag[a] = fl;
                    nneg += __popc(__ballot_sync(SIM_FULL, fl));
                }
                nd_ = nneg > 0;
                #ifndef SIM_NO_SHIFT
                if (first_pass && SIM_SHIFT_DEN * nneg > q && !SIM_KNOB(32)) {
                    // Most of the carried set has the wrong sign: the plan is one that moves along the horizon
                    // (alternating rate limits of an aggressive tuning shift by one index per sample).  Forget the
                    // set and let the iterations below try the SHIFTED set first, constraint by constraint, through
                    // the ordinary add path (any violated constraint is a valid pivot).
                    __syncwarp();
                    if (lane == 0) {
                        int np = 0;
                        for (int a = 0; a < q; ++a) {
                            const int cid = sm.act[a], r = cid >> 2;
                            if ((r & (P - 1)) > 0) sm.dflag[np++] = (cid & 3) | ((r - 1) << 2);
                        }
                        sm.misc[2] = np;
                    }
                    __syncwarp();
                    npref = sm.misc[2];
                    q = 0;
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) amask[s] = 0;
                    it += 1;
                    break;
                }
#endif
            }
            first_pass = false;
            nd_ = __any_sync(SIM_FULL, nd_);
            __syncwarp();
            if (!nd_) break;
            it += 1;
            drop_flagged();
            SIM_DBGSET(" shed-negative");
        }
        if (q > 0) {  // z = z_unc + V l  (l = Li g is still in sm.l)
            add_V(q, sm.l, 1.0, z);
            for (int a = lane; a < q; a += 32) if (sm.mu[a] < 0.0) sm.mu[a] = 0.0;
            __syncwarp();
        }
        // ---- Goldfarb-Idnani iterations ----
        for (;;) {
            levels(lv);
            double bv = -SIM_VIOL_TOL;
            int bi = -1;
            int forced = -1;
            while (pptr < npref) {   // next constraint of the shifted set that is violated and not yet active
                const int cid = sm.dflag[pptr++];
                const int r = cid >> 2, type = cid & 3;
                double mine = 0.0;
                int isact = 1;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (row[s] == r && valid[s]) {
                        mine = type == 0 ? z[s] - dlo[s] : (type == 1 ? dhi[s] - z[s] : (type == 2 ? lv[s] - ulo[s] : uhi[s] - lv[s]));
                        isact = (amask[s] >> type) & 1;
                    }
                const double v = __shfl_sync(SIM_FULL, mine, r & 31);
                const int a_ = __shfl_sync(SIM_FULL, isact, r & 31);
                if (!a_ && v < -SIM_VIOL_TOL) { forced = cid; bv = v; break; }
            }
            if (forced < 0)
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) {
                if (!valid[s]) continue;
                const double sl[4] = {z[s] - dlo[s], dhi[s] - z[s], lv[s] - ulo[s], uhi[s] - lv[s]};
#pragma unroll
                for (int type = 0; type < 4; ++type) {
                    if (amask[s] & (1 << type)) continue;
                    const int id = type | (row[s] << 2);
                    if (sl[type] < bv || (sl[type] == bv && bi >= 0 && id < bi)) { bv = sl[type]; bi = id; }
                }
            }
            if (forced >= 0) bi = forced;
            else {   // pivot rule: among the violated constraints take the EARLIEST horizon index first (then the most
                // violated).  Any violated constraint is a valid Goldfarb-Idnani pivot; walking the horizon in time
                // order follows how rate/level saturation propagates and avoids most of the add/drop churn of the
                // most-violated rule.  One integer warp reduction finds that index -- and tells when nothing is violated.
                const unsigned cmine = bi >= 0 ? (unsigned)(row[0] & (P - 1)) : 0xffffu;
                const unsigned cmin = __reduce_min_sync(SIM_FULL, cmine);
                if (cmin == 0xffffu) break;
#ifndef SIM_PIVOT_MOST_VIOLATED
                if (!most_violated && cmine != cmin) { bv = -SIM_VIOL_TOL; bi = -1; }
#endif
                sim_wargmin(bv, bi);
            }
            const int p = bi;
            SIM_DBG("  pivot %c%d.%d viol %.3e (q=%d)\n", "dDuU"[p & 3], (p >> 2) / P, (p >> 2) % P, bv, q);
            double sp = bv, mu_p = 0.0;
            double wv[NSLOT];
            w_times_normal(p, wv, sm.w, sm.wsc);
            const double gamma = ndot(p, sm.w, sm.wsc);
            for (;;) {
                if (++it > itmax) { n_it += it; return 2; }
                const double l2 = schur_vectors(q, sm.w, sm.wsc);
                const double rho = gamma - l2;
                const int dependent = !(rho > SIM_DEP_TOL * gamma);
                double t1 = SIM_INF;
                int l1 = -1;
                for (int a = lane; a < q; a += 32) {
                    const double ra = sm.rr[a];
                    if (ra > 0.0) {
                        const double t = sm.mu[a] / ra;
                        if (t < t1 || (t == t1 && l1 >= 0 && a < l1)) { t1 = t; l1 = a; }
                    }
                }
                sim_wargmin(t1, l1);
                const double t2 = dependent ? SIM_INF : -sp / rho;
                const double t = t1 < t2 ? t1 : t2;
                if (!(t < SIM_INF)) { n_it += it; return 1; }
                const int full = !(dependent || t1 < t2);
                double dir[NSLOT];
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) dir[s] = wv[s];
                if (!dependent) {  // primal step along dir = W n_p - V l
                    add_V(q, sm.l, -1.0, dir);
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) z[s] += t * dir[s];
                    sp += t * rho;
                }
                for (int a = lane; a < q; a += 32) sm.mu[a] -= t * sm.rr[a];
                mu_p += t;
                __syncwarp();
                if (!full) SIM_DBG("    partial step t=%.3e drop pos %d (%c%d.%d) dependent=%d\n", t, l1, "dDuU"[sm.act[l1 < 0 ? 0 : l1] & 3], (sm.act[l1 < 0 ? 0 : l1] >> 2) / P, (sm.act[l1 < 0 ? 0 : l1] >> 2) % P, dependent);
                if (full) {
                    if (!SPILL && q >= QC) { n_it += it; return SIM_ST_OVERFLOW; }   // re-run on the kernel with the spill area
                    commit(p, q, rho, mu_p, dir);
                    q += 1;
                    break;
                }
                remove_at(l1);   // the blocking constraint (its multiplier just reached zero)
            }
        }
        // ---- one Newton correction on the active constraints if they drifted: z += V Li (-slack_A).  Only after
        // the set changed in this solve: an unchanged set was corrected (or found clean) with these factors before ----
        if (q > 0 && it > 0) {
            publish(lv);
            double worst = 0.0;
            for (int a = lane; a < q; a += 32) {
                const double sl = slack_of(sm.act[a]);
                sm.g[a] = -sl;
                worst = fmax(worst, fabs(sl));
            }
            worst = sim_wmax(worst);
            __syncwarp();
            if (worst > 1e-13) {
                tri_lower(q);
                add_V(q, sm.l, 1.0, z);
                __syncwarp();
            }
        }
        n_it += it;
        if (q > qmax) qmax = q;
        SIM_DBGSET("final");
        // two-phase mode bookkeeping (cold, not inlined): only after an expensive solve or inside the mode -- called after
        // every constrained solve it cost all runs 4 % (gpurun_out/d28*.log)
        if (slot && (it >= SIM_PP_HEAVY || sm.misc[SIM_PP_ON])) sim_pp_update(sm.act, sm.misc, lane, q, it, QC, (int)n_con, (int)n_it);
        return 0;
    }

    // z = M st (st already in shared memory), box/rate check, QP if needed.
    __device__ __forceinline__ int controller_move() {
        const int nst = L.nst;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) z[s] = 0.0;
        if (sim_m_in_smem(P)) {
            double acc0[NSLOT], acc1[NSLOT];
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) { acc0[s] = 0.0; acc1[s] = 0.0; }
            int sg = 0;
#pragma unroll 2
            for (; sg + 1 < nst; sg += 2) {
                const double2 s01 = *reinterpret_cast<const double2 *>(sm.st + sg);   // st is 16-byte aligned
                const double s0 = s01.x, s1 = s01.y;
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (row[s] < R) {
                        acc0[s] = fma(sm.M[(size_t)sg * R + row[s]], s0, acc0[s]);
                        acc1[s] = fma(sm.M[(size_t)(sg + 1) * R + row[s]], s1, acc1[s]);
                    }
            }
            if (sg < nst) {
                const double s0 = sm.st[sg];
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    if (row[s] < R) acc0[s] = fma(sm.M[(size_t)sg * R + row[s]], s0, acc0[s]);
            }
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) z[s] = acc0[s] + acc1[s];
        } else {
            // M from global memory: SIM_MB columns (SIM_MB * NSLOT independent loads per lane) in flight at a time; the last
            // batch is predicated, not a column-at-a-time tail (each of those was a dependent L2 round trip)
            double acc0[NSLOT], acc1[NSLOT];
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) { acc0[s] = 0.0; acc1[s] = 0.0; }
            const double *mp = Mp + lane;
#pragma unroll 1
            for (int sg = 0; sg < nst; sg += SIM_MB) {
                double mv[SIM_MB][NSLOT];
#pragma unroll
                for (int e = 0; e < SIM_MB; ++e)
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s)
                        mv[e][s] = (valid[s] && sg + e < nst) ? (MSM ? mp[(size_t)(sg + e) * R + s * 32] : __ldg(mp + (size_t)(sg + e) * R + s * 32)) : 0.0;   // padded rows of M are zero: not fetched
#pragma unroll
                for (int e = 0; e < SIM_MB; e += 2) {
                    const double s0 = sg + e < nst ? sm.st[sg + e] : 0.0, s1 = sg + e + 1 < nst ? sm.st[sg + e + 1] : 0.0;
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) { acc0[s] = fma(mv[e][s], s0, acc0[s]); acc1[s] = fma(mv[e + 1][s], s1, acc1[s]); }
                }
            }
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) z[s] = acc0[s] + acc1[s];
        }
        if (q == 0) {
            double lv[NSLOT];
            levels(lv);
            int bad = 0;
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                if (valid[s])
                    bad |= (z[s] - dlo[s] < -SIM_VIOL_TOL) | (dhi[s] - z[s] < -SIM_VIOL_TOL) |
                           (lv[s] - ulo[s] < -SIM_VIOL_TOL) | (uhi[s] - lv[s] < -SIM_VIOL_TOL);
            if (!__any_sync(SIM_FULL, bad)) return 0;
        }
        n_con += 1;
        return qp_solve();
    }
};

// ------------------------------------------------------------------------------------------------
// One closed-loop run.  sel: -2 user set-point (GAM / RAW); -1 VNS unit step on every output;
// i >= 0 VNS unit step on output i only (VNS2.m:148-165).  mode: 0 RAW, 1 GAM, 2 VNS.
// Mg: nst x R (row-padded, [col][row]); Wg: 2R x R (W, then W times the level normals).
// ------------------------------------------------------------------------------------------------
// LEAN: the GAM cost-only specialisation (mode 1, no trajectories, no open-loop pass, no diagnostics) -- the image the
// tuning loop and the benchmark run; dropping the other modes' code makes it ~20 % smaller (instruction cache).
// VLEAN: the same for the VNS objective (mode 2: open-loop pass and lock-step open-loop rollout kept, no trajectories).
template <int NU, int P, bool LEAN = false, bool VLEAN = false>
__device__ __forceinline__ int sim_run(const MpcLayout &L, const MpcTables &T, int m, const double *__restrict__ Mg,
                                       const double *__restrict__ Wg, int mode_arg, int sel, double *smem, double *gscr,
                                       const MpcRunOut &out_arg, double *pslot = nullptr) {
    const int mode = LEAN ? 1 : (VLEAN ? 2 : mode_arg);
    MpcRunOut out = out_arg;
    if (LEAN || VLEAN) { out.y = nullptr; out.u = nullptr; out.ys = nullptr; out.uopt = nullptr; out.diag = nullptr; out.trace = nullptr; }
    constexpr int R = NU * P;
    constexpr int NSLOT = (R + 31) / 32;
    constexpr int QC = (R <= SIM_QC_MAX) ? R : SIM_QC_MAX;
    const int lane = threadIdx.x & 31;
    const int ny = L.ny, nd = L.nd, nw = L.nw, nch = ny * nw, nst = L.nst, nit = L.nit;
    const int HL = sim_hl(L);
    const int nsig = 2 * ny + nd;
    SimWarp<NU, P> wp(L);
    SimSm &sm = wp.sm;
    {   // carve shared memory
        double *p = smem;
        sm.M = p; if (sim_m_in_smem(P)) p += (size_t)nst * R;
        sm.st = p; p += (nst + 1) & ~1;
        sm.x = p; p += nch;
        sm.xol = p; p += nch;
        sm.hist = p; p += (size_t)nw * HL;
        sm.cha = p; p += nch; sm.chb0 = p; p += nch; sm.chb1 = p; p += nch; sm.chg = p; p += nch;
        sm.sig = p; p += (size_t)SIM_CH * nsig;
        sm.uopt = p; p += NU * P;
        sm.bnd = p; p += 4 * NU;
        sm.z = p; p += R; sm.lvl = p; p += R; sm.w = p; p += R; sm.wsc = p; p += R;
        sm.g = p; p += R; sm.l = p; p += R; sm.rr = p; p += R; sm.mu = p; p += R;
        sm.V = p; p += (size_t)QC * R;
        sm.Li = p; p += (size_t)QC * QC;
        int *ip = (int *)p;
        sm.chd = ip; ip += nch; sm.chj = ip; ip += nch; sm.role = ip; ip += nst;
        sm.act = ip; ip += R; sm.dflag = ip; ip += R; sm.misc = ip;
    }
    wp.gscr = gscr; wp.W = Wg; wp.lane = lane; wp.m = m; wp.q = 0; wp.n_con = 0; wp.n_it = 0; wp.qmax = 0; wp.churn = 0;
#ifdef SIM_NO_TWO_PHASE
    wp.slot = nullptr;
#else
    wp.slot = pslot;
#endif
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int r = s * 32 + lane;
        const int j = r / P, c = r & (P - 1);
        wp.row[s] = r;
        wp.valid[s] = (r < R) && (c < m);
        const int jj = j < NU ? j : 0;
        wp.dlo[s] = wp.valid[s] ? L.dumin[jj] : -SIM_INF;
        wp.dhi[s] = wp.valid[s] ? L.dumax[jj] : SIM_INF;
        wp.ulo[s] = wp.valid[s] ? L.umin[jj] : -SIM_INF;
        wp.uhi[s] = wp.valid[s] ? L.umax[jj] : SIM_INF;
        wp.z[s] = 0.0;
        wp.amask[s] = 0;
    }
#pragma unroll
    for (int j = 0; j < NU; ++j) wp.u[j] = 0.0;
    // ---- one-time staging ----
    if (sim_m_in_smem(P)) { for (int i = lane; i < nst * R; i += 32) sm.M[i] = Mg[i]; }
    wp.Mp = sim_m_in_smem(P) ? sm.M : Mg;
    for (int ch = lane; ch < nch; ch += 32) {
        sm.cha[ch] = L.a[ch]; sm.chb0[ch] = L.b0[ch]; sm.chb1[ch] = L.b1[ch]; sm.chg[ch] = L.gain[ch];
        sm.chd[ch] = L.d[ch]; sm.chj[ch] = ch % nw;
        sm.x[ch] = 0.0; sm.xol[ch] = 0.0;
    }
    for (int i = lane; i < nw * HL; i += 32) sm.hist[i] = 0.0;
    for (int i = lane; i < NU * P; i += 32) sm.uopt[i] = 0.0;
    for (int j = lane; j < NU; j += 32) {
        sm.bnd[4 * j + 0] = L.dumin[j]; sm.bnd[4 * j + 1] = L.dumax[j]; sm.bnd[4 * j + 2] = L.umin[j]; sm.bnd[4 * j + 3] = L.umax[j];
    }
    if (lane < SIM_MISC_INTS) sm.misc[lane] = lane == SIM_PP_SIG2 ? 1 : 0;
    // role of every deviation coordinate: kind | a << 2 | b << 12  (kind 0: x_ch, 1: hist(j=a, lag=b), 2: e_i)
    for (int col = lane; col < nst; col += 32) {
        int role;
        if (col < nch) role = 0 | (col << 2);
        else if (col >= L.stoff_e) role = 2 | ((col - L.stoff_e) << 2);
        else {
            int j = 0;
            while (j + 1 < nw && col >= L.stoff_h[j + 1]) ++j;
            role = 1 | (j << 2) | ((L.hq0[j] + (col - L.stoff_h[j])) << 12);
        }
        sm.role[col] = role;
    }
    __syncwarp();
    int head = 0;   // uniform: physical slot of lag 0 in every circular history
    int status = 0;
    const bool want_ol = LEAN ? false : (VLEAN ? true : ((mode != 1) || out.ys || out.uopt));
    double jnu = 0.0;
    double cost_acc = 0.0;  // per lane: lane i < ny accumulates output i

    // set-point of output i at sample k for this run
    auto setpoint = [&](int i, int k, double r_user) -> double {
        if (sel == -2) return r_user;
        return (sel == -1 || sel == i) ? (k >= L.inK - 1 ? 1.0 : 0.0) : 0.0;
    };
    // build st from the physical state; hv_md / r taken from `sigrow` (r at [0,ny), yref [ny,2ny), v [2ny,..))
    auto build_st = [&](const double *sigrow, int k, bool do_cost) {
        for (int col = lane; col < nst; col += 32) {
            const int role = sm.role[col];
            const int kind = role & 3, a = (role >> 2) & 1023, b = role >> 12;
            double val;
            if (kind == 0) {
                const int j = sm.chj[a];
                const double hv = j < NU ? sim_pick<NU>(wp.u, j) : sigrow[2 * ny + (j - NU)];
                val = sm.x[a] - sm.chg[a] * hv;
            } else if (kind == 1) {
                const int j = a;
                const double hv = j < NU ? sim_pick<NU>(wp.u, j) : sigrow[2 * ny + (j - NU)];
                int pos = head + b;
                if (pos >= HL) pos -= HL;
                val = sm.hist[j * HL + pos] - hv;
            } else {
                const int i = a;
                double yi = 0.0, ysi = 0.0, gsum = 0.0;
                for (int j = 0; j < nw; ++j) {
                    yi += sm.x[i * nw + j];
                    ysi += sm.xol[i * nw + j];
                    const double hv = j < NU ? sim_pick<NU>(wp.u, j) : sigrow[2 * ny + (j - NU)];
                    gsum += sm.chg[i * nw + j] * hv;
                }
                val = setpoint(i, k, sigrow[i]) - gsum;
                if (do_cost) {
                    const bool mine = (sel < 0 || sel == i);
                    if (out.y && mine) out.y[(size_t)i * nit + k] = yi;
                    if (out.ys && mine) out.ys[(size_t)i * nit + k] = ysi;
                    if (mode == 1) {
                        const double e = yi - sigrow[ny + i];
                        cost_acc += e * e;
                    } else if (mode == 2 && k >= L.inK - 1 && mine) {
                        const double e2 = yi - ysi, er = yi - sigrow[ny + i];
                        cost_acc += e2 * e2 + er * er;
                    }
                }
            }
            sm.st[col] = val;
        }
        __syncwarp();
    };
    auto stage_signals = [&](int k0) {   // lane = signal column (nsig <= 32), one row per pass: no index division, no unrolling
        const int cnt = (nit - k0) < SIM_CH ? (nit - k0) : SIM_CH;
        const int c = lane;
#pragma unroll 1
        for (int kk = 0; kk < cnt; ++kk) {
            if (c < nsig) {
                double v;
                if (c < ny) v = T.r[(size_t)(k0 + kk) * ny + c];
                else if (c < 2 * ny) v = T.yref[(size_t)(c - ny) * nit + (k0 + kk)];
                else v = T.v[(size_t)(k0 + kk) * nd + (c - 2 * ny)];
                sm.sig[kk * nsig + c] = v;
            }
        }
        __syncwarp();
    };

    // ---------------- open-loop optimum (closedloop_toolbox.m:85-98) = pass k = -1, then the closed loop (:50)
    // with the open-loop rollout (:100) in lock-step.  One loop, so that the controller (and the whole active-set
    // solver inlined into it) exists ONCE in the kernel image: code size is a first-order cost here, the
    // instruction cache does not hold the kernel. ----------------
    for (int k = want_ol ? -1 : 0; k < nit; ++k) {
        const bool ol = k < 0;
        if (ol) {
            // the fresh controller state: x = 0, histories 0, u(-1) = 0; r = last row, v = last row
            if (lane < nsig) {
                const int c = lane;
                double v;
                if (c < ny) v = T.r[(size_t)(nit - 1) * ny + c];
                else if (c < 2 * ny) v = 0.0;
                else v = T.v[(size_t)(nit - 1) * nd + (c - 2 * ny)];
                sm.sig[c] = v;
            }
            __syncwarp();
        } else if ((k & (SIM_CH - 1)) == 0) {
            stage_signals(k);
        }
        const double *sigrow = ol ? sm.sig : sm.sig + (size_t)(k & (SIM_CH - 1)) * nsig;
        build_st(sigrow, ol ? nit - 1 : k, !ol);
        const unsigned long long it_before = wp.n_it;
#ifdef MPC_SIMT_EMULATION
        if (lane == 0) g_sim_verbose = (g_sim_knob >> 16) && k >= (g_sim_knob >> 16) && k < (g_sim_knob >> 16) + 2;
        if (g_sim_verbose && lane == 0) printf("=== sample %d\n", k);
        __syncwarp();
#endif
        const int rc = wp.controller_move();
        if (rc) status = rc;
        if (ol) {
            // Uopt rows = SEQUENTIAL cumulative sum of the moves from u(-1) = 0: a move that is exactly 0 must
            // repeat the previous level bit-for-bit, because VNS2.m:183-191 divides by these differences
            // (a tree-ordered scan would turn exact zeros into 1-ulp noise and Jnu terms of 1e+30).
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
                if (wp.row[s] < R) sm.z[wp.row[s]] = wp.z[s];
            __syncwarp();
            for (int j = lane; j < NU; j += 32) {
                double lvl = 0.0;
                for (int c = 0; c < P; ++c) { lvl += sm.z[j * P + c]; sm.uopt[j * P + c] = lvl; }
            }
            __syncwarp();
            if (mode == 2) {  // Jnu (VNS2.m:183-191)
                double part = 0.0;
                for (int j = lane; j < NU; j += 32) {
                    if (sel < 0 || sel == j) {
                        const double u0 = fabs(sm.uopt[j * P]);
                        for (int c = 0; c + 1 < m && c + 1 < nit; ++c) {
                            const double df = fabs(sm.uopt[j * P + c + 1] - sm.uopt[j * P + c]);
                            const double xn = u0 / df;
                            if (fabs(xn) <= 1.7976931348623157e308) part += xn * xn;  // inf / nan -> 0 (VNS2.m:186)
                        }
                    }
                }
                jnu = sim_wsum(part);
            }
            // the closed loop starts from an empty active set
            wp.q = 0;
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) wp.amask[s] = 0;
            __syncwarp();
            continue;
        }
        if (out.trace && lane == 0) { out.trace[2 * k] = (int)(wp.n_it - it_before); out.trace[2 * k + 1] = wp.q; }
        // apply the first move of every input
#pragma unroll
        for (int j = 0; j < NU; ++j) {
            const int r0 = j * P;
            const double du = __shfl_sync(SIM_FULL, wp.z[r0 >> 5], r0 & 31);
            wp.u[j] += du;
        }
        if (lane < NU) {
            const int j = lane;
            const bool mine = (sel < 0 || sel == j);
            if (out.u && mine) out.u[(size_t)j * nit + k] = sim_pick<NU>(wp.u, j);
            if (out.uopt && mine) out.uopt[(size_t)j * nit + k] = sm.uopt[j * P + (k < m ? k : m - 1)];
        }
        // plant sample: x(k+1) = a x(k) + b0 w(k+1-d) + b1 w(k-d); lag q of the history is w(k-1-q)
        for (int ch = lane; ch < nch; ch += 32) {
            const int j = sm.chj[ch], dd = sm.chd[ch];
            const double wk = j < NU ? sim_pick<NU>(wp.u, j) : sigrow[2 * ny + (j - NU)];
            int p1 = head + (dd > 0 ? dd - 1 : 0); if (p1 >= HL) p1 -= HL;   // indices stay in range even
            int p0 = head + (dd > 1 ? dd - 2 : 0); if (p0 >= HL) p0 -= HL;   // when the value is unused
            const double w1 = dd == 0 ? wk : sm.hist[j * HL + p1];
            const double w0 = dd == 0 ? 0.0 : (dd == 1 ? wk : sm.hist[j * HL + p0]);
            sm.x[ch] = sm.cha[ch] * sm.x[ch] + sm.chb0[ch] * w0 + sm.chb1[ch] * w1;
            if (want_ol) {
                double o1, o0;
                if (j < NU) {
                    const int k1 = k - dd, k0 = k + 1 - dd;
                    o1 = k1 < 0 ? 0.0 : sm.uopt[j * P + (k1 < m ? k1 : m - 1)];
                    o0 = (dd == 0 || k0 < 0) ? 0.0 : sm.uopt[j * P + (k0 < m ? k0 : m - 1)];
                } else {  // the measured disturbance is the same signal in both simulations
                    o1 = w1; o0 = w0;
                }
                sm.xol[ch] = sm.cha[ch] * sm.xol[ch] + sm.chb0[ch] * o0 + sm.chb1[ch] * o1;
            }
        }
        // push w(k): the slot being overwritten held lag HL-1, which nobody reads
        const int nhead = head == 0 ? HL - 1 : head - 1;
        if (lane < nw) {
            const int j = lane;
            sm.hist[j * HL + nhead] = j < NU ? sim_pick<NU>(wp.u, j) : sigrow[2 * ny + (j - NU)];
        }
        head = nhead;
        __syncwarp();
    }
    // ---------------- costs ----------------
    if (out.cost) {
        if (mode == 1) {
            // lane holding column stoff_e + i accumulated output i
            for (int i = 0; i < ny; ++i) {
                const int col = L.stoff_e + i;
                const double ci = __shfl_sync(SIM_FULL, cost_acc, col & 31);
                if (lane == 0) out.cost[i] = status ? NAN : ci;
            }
        } else if (mode == 2) {
            const double tot = sim_wsum(cost_acc) + jnu;
            if (lane == 0) out.cost[0] = status ? NAN : tot;
        }
    }
    if (out.counters && lane == 0) {
        atomicAdd(out.counters + 0, wp.n_con);
        atomicAdd(out.counters + 1, wp.n_it);
    }
    if (out.diag && lane == 0) { out.diag[0] = wp.n_con; out.diag[1] = wp.n_it; out.diag[2] = (unsigned long long)wp.qmax | ((unsigned long long)sm.misc[SIM_PP_COUNT] << 32); }
    return status;
}
