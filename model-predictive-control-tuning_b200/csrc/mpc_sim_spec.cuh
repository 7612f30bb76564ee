// mpc_sim_spec.cuh -- closed-loop run with LANE-RESIDENT plant state and SPECULATIVE unconstrained stretches.
//
// Same semantics as sim_run (mpc_sim.cuh): closedloop_toolbox.m:50-100 fused with the GAM / VNS sums, one warp per
// run, the same warm-started dual active-set solver (SimWarp::controller_move) for every sample that needs a QP.
// What changes is everything around it, for plants whose deviation state fits one warp (nst <= 32):
//
// 1. Lane c owns deviation coordinate c for the whole run.  Its constants live in registers: the held-input
//    coefficients cu[j] (st_c = base_c - sum_j cu[j] u_j(k-1)), row c of the first-move rows of M, and for a channel
//    lane the pole / numerator / delay and the channel state itself.  A sample then costs one short dependent chain
//    (NU FMAs -> NU interleaved warp sums -> NU adds -> the plant FMAs) instead of a role decode through shared memory.
//
// 2. About 90 % of the samples of a tuning population need no QP: the unconstrained plan z = M st is feasible and only
//    its FIRST move per input is applied.  The feasibility test, however, needs all R rows of z, i.e. the whole
//    nst x R product with M read from L2 -- per sample, on a critical path.  While no active set is carried the loop
//    therefore SPECULATES: it advances up to SIM_SPEC_T samples applying first moves only (NU rows of M, registers),
//    logging st(k) in shared memory, and then verifies the whole block at once, Z = M [st(k) .. st(k+T-1)]: every element
//    of M is fetched once per block and feeds T independent FMAs (instruction-level parallelism T x NSLOT instead of a
//    dependent chain).  If sample k+t is the first whose plan violates a limit, the state snapshot taken at the block
//    start is restored, the t good samples are replayed (deterministic: same instructions, same values) and sample k+t
//    goes through the ordinary controller.  Results do not depend on whether a stretch was speculated: a sample
//    is only ever committed with the first move of a plan that the full product proved feasible.
//
// The speculation block and the snapshot alias the active-set factor (V, Li): they are only live while q == 0.
#pragma once
#include "mpc_sim.cuh"
#ifdef MPC_SIMT_EMULATION
extern long long g_spec_blocks, g_spec_samples, g_spec_fail, g_spec_replay;
#endif

#ifndef SIM_SPEC_T
#define SIM_SPEC_T 8      /* samples per speculative block */
#endif
#ifndef SIM_SPEC_COOL
#define SIM_SPEC_COOL 2   /* unconstrained ordinary samples before speculation resumes after a QP */
#endif
#ifndef SIM_SPEC_CB
#define SIM_SPEC_CB 6     /* columns of M in flight per batch of the block verification */
#endif

// ---- shared-memory layout of the speculative kernel: COMPILE-TIME offsets (doubles) for a given (NU, P, LEAN) ----
// Run-time strides in the hot loop cost address arithmetic on every access (and the compiler re-derives the carve offsets
// inside the loop when registers are short); with fixed offsets every access is base + immediate.  The price is a few
// size limits (sim_spec_ok): nst <= 32, ny*nw <= 32, nw*HL <= 64, 2ny+nd <= SIM_SPEC_NSIG.
#define SIM_SPEC_HMAX 64
#define SIM_SPEC_NSIG 12
#ifndef SIM_SPEC_CH
#define SIM_SPEC_CH 8     /* samples of the packed signal table staged at a time (power of two, >= SIM_SPEC_T) */
#endif
struct SimSpecLayout {
    int x, xol, st, chg, hist, bnd, uopt, z, lvl, w, wsc, g, l, rr, mu, sig, vli, ints, total;
    // speculation area, offsets relative to vli: st log [T][32], u log [T+1][4], z log [T][ZS], snapshot
    int stb, ub, zlog, snl, snh, snu, spec_end;
};
static MPC_HD constexpr int sim_spec_zs(int nu, int P) { return (nu * (P + 1)) | 1; }
static MPC_HD constexpr SimSpecLayout sim_spec_layout(int nu, int P, bool lean) {
    SimSpecLayout o{};
    const int R = nu * P, qc = R <= SIM_SPEC_QC ? R : SIM_SPEC_QC;
    int p = 0;
    o.x = p; p += 32; o.xol = p; p += 32; o.st = p; p += 32; o.chg = p; p += 32;
    o.hist = p; p += SIM_SPEC_HMAX; o.bnd = p; p += 4 * MPC_MAXU;
    o.uopt = p; p += lean ? 0 : R;
    o.z = p; p += R; o.lvl = p; p += R; o.w = p; p += R; o.wsc = p; p += R;
    o.g = p; p += qc; o.l = p; p += qc; o.rr = p; p += qc; o.mu = p; p += qc;   // indexed by active-set position (< QC)
    o.sig = p; p += SIM_SPEC_CH * SIM_SPEC_NSIG;
    p = (p + 1) & ~1;   // 16-byte alignment of the st log
    o.vli = p;
    int s = 0;
    o.stb = s; s += SIM_SPEC_T * 32;
    o.ub = s; s += (SIM_SPEC_T + 1) * MPC_MAXU;
    o.zlog = s; s += SIM_SPEC_T * sim_spec_zs(nu, P);
    o.snl = s; s += 3 * 32;
    o.snh = s; s += SIM_SPEC_HMAX;
    o.snu = s; s += MPC_MAXU;
    o.spec_end = s;
    const int vli = qc * R + qc * (qc + 1) / 2;
    p += s > vli ? s : vli;
    o.ints = p;
    p += (2 * R + SIM_MISC_INTS + 1) / 2 + 1;   // act, dflag, misc
    o.total = p;
    return o;
}
static MPC_HD bool sim_spec_ok(const MpcLayout &L) {
    return L.nst <= 32 && L.ny * L.nw <= 32 && L.nw * sim_hl(L) <= SIM_SPEC_HMAX && 2 * L.ny + L.nd <= SIM_SPEC_NSIG && SIM_SPEC_T <= 8;
}
static MPC_HD size_t sim_spec_smem_doubles(const MpcLayout &, int nu, int P, bool lean = false) { return (size_t)sim_spec_layout(nu, P, lean).total; }
// MSM image: + M itself (32 columns of R rows) behind the base layout
static MPC_HD size_t sim_spec_msm_doubles(int nu, int P) { return (size_t)32 * nu * P; }

// MSM: M is copied to shared memory at the start of the run (12 KB for Shell3x3).  For SMALL populations only, whose launch
// keeps few runs per SM anyway: 9.82 -> 9.50 ms on 4096 candidates, but 23.4 -> 26.1 ms on 16384 where residency counts.
template <int NU, int P, bool LEAN = false, bool VLEAN = false, bool MSM = false>
__device__ __forceinline__ int sim_run_spec(const MpcLayout &L, const MpcTables &T, int m, const double *__restrict__ Mg,
                                            const double *__restrict__ Wg, int mode_arg, int sel, double *smem, double *gscr,
                                            const MpcRunOut &out_arg, double *pslot = nullptr) {
    const int mode = LEAN ? 1 : (VLEAN ? 2 : mode_arg);
    MpcRunOut out = out_arg;
    if (LEAN || VLEAN) { out.y = nullptr; out.u = nullptr; out.ys = nullptr; out.uopt = nullptr; out.diag = nullptr; out.trace = nullptr; }
    constexpr int R = NU * P;
    constexpr int NSLOT = (R + 31) / 32;
    constexpr int SIM_SPEC_ZS = sim_spec_zs(NU, P);
    constexpr int NSTB = 32;
    constexpr SimSpecLayout O = sim_spec_layout(NU, P, LEAN);
    const int lane = threadIdx.x & 31;
    const int ny = L.ny, nd = L.nd, nw = L.nw, nch = ny * nw, nst = L.nst, nit = L.nit;
    const int HL = sim_hl(L);
    const int nsig = 2 * ny + nd;
    SimWarp<NU, P, false, MSM> wp(L);   // no spill area: a QP beyond QC active constraints -> SIM_ST_OVERFLOW
    SimSm &sm = wp.sm;
    double *const stb = smem + O.vli + O.stb, *const ub = smem + O.vli + O.ub, *const zlog = smem + O.vli + O.zlog;
    double *const snl = smem + O.vli + O.snl, *const snh = smem + O.vli + O.snh, *const snu = smem + O.vli + O.snu;
    {
        sm.M = nullptr; sm.st = smem + O.st; sm.x = smem + O.x; sm.xol = smem + O.xol; sm.hist = smem + O.hist;
        sm.cha = sm.chb0 = sm.chb1 = nullptr; sm.chg = smem + O.chg; sm.sig = smem + O.sig;
        sm.uopt = smem + O.uopt; sm.bnd = smem + O.bnd;
        sm.z = smem + O.z; sm.lvl = smem + O.lvl; sm.w = smem + O.w; sm.wsc = smem + O.wsc;
        sm.g = smem + O.g; sm.l = smem + O.l; sm.rr = smem + O.rr; sm.mu = smem + O.mu;
        sm.V = smem + O.vli; sm.Li = smem + O.vli + (size_t)SimWarp<NU, P, false, MSM>::QC * R;
        int *ip = (int *)(smem + O.ints);
        sm.chd = sm.chj = sm.role = nullptr;
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
    if (MSM) {
        double *msm = smem + ((O.total + 1) & ~1);
        for (int i = lane; i < nst * R; i += 32) msm[i] = Mg[i];
        wp.Mp = msm;
    } else
    wp.Mp = Mg;   // M stays in global memory (L2): read once per speculative block / per constrained sample
    sm.chg[lane] = lane < nch ? L.gain[lane] : 0.0;
    sm.x[lane] = 0.0; sm.xol[lane] = 0.0;
    for (int i = lane; i < SIM_SPEC_HMAX; i += 32) sm.hist[i] = 0.0;
    if (!LEAN) for (int i = lane; i < NU * P; i += 32) sm.uopt[i] = 0.0;
    for (int j = lane; j < NU; j += 32) {
        sm.bnd[4 * j + 0] = L.dumin[j]; sm.bnd[4 * j + 1] = L.dumax[j]; sm.bnd[4 * j + 2] = L.umin[j]; sm.bnd[4 * j + 3] = L.umax[j];
    }
    if (lane < SIM_MISC_INTS) sm.misc[lane] = lane == SIM_PP_SIG2 ? 1 : 0;
    // ---- this lane's deviation coordinate: st = base - sum_j cu[j] u_j ----
    const int col = lane;
    const int kind = col < nch ? 0 : (col < L.stoff_e ? 1 : (col < nst ? 2 : 3));   // channel state, input history, set-point error, idle
    double cu[NU], m0[NU];
#pragma unroll
    for (int j = 0; j < NU; ++j) { cu[j] = 0.0; m0[j] = col < nst ? Mg[(size_t)col * R + j * P] : 0.0; }
    double mdg = 0.0;           // coefficient of the measured disturbance this coordinate is referred to (0: none)
    int mdj = 0;                // ... and its index
    double ca = 0.0, cb0 = 0.0, cb1 = 0.0;   // channel lanes: pole and numerator
    const double *hrow = sm.hist;           // channel and history lanes: the circular history of this lane's input
    int l1 = 0, l0 = 0;                     // channel lanes: lags of w(k-d), w(k+1-d) once w(k) has been pushed
    int hlag = 0;                           // history lanes: lag of this coordinate
    int ei = 0;                             // error lanes: output index
    if (kind == 0) {
        const int cj = col % nw, cd = L.d[col];
        ca = L.a[col]; cb0 = L.b0[col]; cb1 = L.b1[col];
        hrow = sm.hist + cj * HL; l1 = cd; l0 = cd > 0 ? cd - 1 : 0;   // cd == 0: b0 == 0 (checked at create), any finite w0 does
        const double g = L.gain[col];
        if (cj < NU) {
#pragma unroll
            for (int j = 0; j < NU; ++j) cu[j] = j == cj ? g : 0.0;
        } else { mdg = g; mdj = cj - NU; }
    } else if (kind == 1) {
        int j = 0;
        while (j + 1 < nw && col >= L.stoff_h[j + 1]) ++j;
        hrow = sm.hist + j * HL; hlag = L.hq0[j] + (col - L.stoff_h[j]);
        if (j < NU) {
#pragma unroll
            for (int jj = 0; jj < NU; ++jj) cu[jj] = jj == j ? 1.0 : 0.0;
        } else { mdg = 1.0; mdj = j - NU; }
    } else if (kind == 2) {
        ei = col - L.stoff_e;
#pragma unroll
        for (int j = 0; j < NU; ++j) cu[j] = L.gain[ei * nw + j];
    }
    double xr = 0.0, xolr = 0.0;   // channel lanes: x_ch(k) of the closed loop and of the open-loop rollout
    const double *yrow = sm.x + (lane < ny ? lane : 0) * nw;   // cost lanes (lane i < ny): the channels of output i
    __syncwarp();
    int head = 0;   // uniform: physical slot of lag 0 in every circular history
    int status = 0;
    const bool want_ol = LEAN ? false : (VLEAN ? true : ((mode != 1) || out.ys || out.uopt));
    double jnu = 0.0;
    double cost_acc = 0.0;  // lane i < ny accumulates output i

    auto setpoint = [&](int i, int k, double r_user) -> double {
        if (LEAN || sel == -2) return r_user;
        return (sel == -1 || sel == i) ? (k >= L.inK - 1 ? 1.0 : 0.0) : 0.0;
    };
    auto stage_signals = [&](int k0) {   // straight copy of SIM_CH rows of the packed signal table
        const int cnt = ((nit - k0) < SIM_SPEC_CH ? (nit - k0) : SIM_SPEC_CH) * nsig;
        const double *src = T.sig + (size_t)k0 * nsig;
#pragma unroll 1
        for (int i = lane; i < cnt; i += 32) sm.sig[i] = src[i];
        __syncwarp();
    };
    // Block verification: Z = M [st(k0) .. st(k0+T-1)] from the st log, level scan and limit test per sample.
    // Returns the number of leading samples whose unconstrained plan is feasible (cnt: all of them).
    auto verify = [&](int cnt) -> int {
        double acc[SIM_SPEC_T][NSLOT];
#pragma unroll
        for (int t = 0; t < SIM_SPEC_T; ++t)
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) acc[t][s] = 0.0;
        const double *mp = wp.Mp + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < nst; c0 += SIM_SPEC_CB) {
            double mv[SIM_SPEC_CB][NSLOT];
#pragma unroll
            for (int e = 0; e < SIM_SPEC_CB; ++e)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s)
                    mv[e][s] = (wp.valid[s] && c0 + e < nst) ? (MSM ? mp[(size_t)(c0 + e) * R + s * 32] : __ldg(mp + (size_t)(c0 + e) * R + s * 32)) : 0.0;
#pragma unroll
            for (int e = 0; e < SIM_SPEC_CB; e += 2) {
#pragma unroll
                for (int t = 0; t < SIM_SPEC_T; ++t) {
                    const double2 sv = *reinterpret_cast<const double2 *>(stb + (size_t)t * NSTB + c0 + e);
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) {
                        acc[t][s] = fma(mv[e][s], sv.x, acc[t][s]);
                        acc[t][s] = fma(mv[e + 1][s], sv.y, acc[t][s]);
                    }
                }
            }
        }
        // the plans go to the z log; lane (t, j) then walks input j's moves of sample t serially and tests the rate and
        // level limits (no shuffles, one copy of the code; strides ZS and P + 1 keep the walk free of bank conflicts)
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            if (lane + 32 * s < R) {
                double *zp = zlog + ((lane + 32 * s) / P) * (P + 1) + ((lane + 32 * s) & (P - 1));
#pragma unroll
                for (int t = 0; t < SIM_SPEC_T; ++t) zp[t * SIM_SPEC_ZS] = acc[t][s];
            }
        }
        __syncwarp();
        unsigned badmask = 0u;
        {
            const int t = lane >> 2, j = lane & 3;
            if (t < cnt && j < NU) {
                double lv = ub[t * MPC_MAXU + j];
                const double dlo = sm.bnd[4 * j + 0], dhi = sm.bnd[4 * j + 1], lo = sm.bnd[4 * j + 2], hi = sm.bnd[4 * j + 3];
                const double *zr = zlog + t * SIM_SPEC_ZS + j * (P + 1);
                int bad = 0;
#pragma unroll 1
                for (int c = 0; c < m; ++c) {
                    const double zc = zr[c];
                    lv += zc;
                    bad |= (zc - dlo < -SIM_VIOL_TOL) | (dhi - zc < -SIM_VIOL_TOL) | (lv - lo < -SIM_VIOL_TOL) | (hi - lv < -SIM_VIOL_TOL);
                }
                if (bad) badmask = 1u << t;
            }
        }
        badmask = __reduce_or_sync(SIM_FULL, badmask) & ((1u << cnt) - 1u);
        return badmask ? __ffs((int)badmask) - 1 : cnt;
    };

    // ---------------- the sample loop.  k = -1: open-loop optimum (closedloop_toolbox.m:85-98), then the closed loop (:50)
    // with the open-loop rollout (:100) in lock-step.  seg_end >= 0: samples [seg_k0, seg_end) are advanced on first moves
    // only, either speculatively (verify_pending) or as the replay of the verified part of a failed block. ----------------
    int seg_k0 = 0, seg_end = -1, cool = 0, head_s = 0;
    bool verify_pending = false;
    int k = want_ol ? -1 : 0;
    for (;;) {
        if (k == seg_end) {
            if (verify_pending) {
                verify_pending = false;
                const int cnt = k - seg_k0;
                const int good = verify(cnt);
#ifdef MPC_SIMT_EMULATION
                if (lane == 0) { g_spec_blocks += 1; g_spec_samples += cnt; if (good < cnt) { g_spec_fail += 1; g_spec_replay += good; } }
#endif
                if (good < cnt) {   // restore the snapshot, replay the good samples, then the ordinary controller
                    xr = snl[lane]; xolr = snl[32 + lane]; cost_acc = snl[64 + lane];
                    if (kind == 0) { sm.x[col] = xr; sm.xol[col] = xolr; }
                    for (int i = lane; i < SIM_SPEC_HMAX; i += 32) sm.hist[i] = snh[i];
#pragma unroll
                    for (int j = 0; j < NU; ++j) wp.u[j] = snu[j];
                    head = head_s;
                    __syncwarp();
                    k = seg_k0; seg_end = seg_k0 + good; cool = SIM_SPEC_COOL;
                    continue;
                }
            }
            seg_end = -1;
        }
        if (k >= nit) break;
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
        } else if ((k & (SIM_SPEC_CH - 1)) == 0) {
            stage_signals(k);
        }
        if (seg_end < 0 && !ol && wp.q == 0 && cool == 0 && !SIM_KNOB(128)) {
            // open a speculative block: stays inside the staged signal chunk
            int cnt = SIM_SPEC_CH - (k & (SIM_SPEC_CH - 1));
            cnt = cnt < SIM_SPEC_T ? cnt : SIM_SPEC_T;
            cnt = cnt < nit - k ? cnt : nit - k;
            snl[lane] = xr; snl[32 + lane] = xolr; snl[64 + lane] = cost_acc;
            for (int i = lane; i < SIM_SPEC_HMAX; i += 32) snh[i] = sm.hist[i];
            if (lane < NU) { const double uj = sim_pick<NU>(wp.u, lane); snu[lane] = uj; ub[lane] = uj; }
            head_s = head;
            seg_k0 = k; seg_end = k + cnt; verify_pending = true;
            // (no barrier needed: the snapshot is read after the verification's reductions at the earliest)
        }
        const bool spec = seg_end >= 0;
        const double *sigrow = ol ? sm.sig : sm.sig + (k & (SIM_SPEC_CH - 1)) * nsig;
        // ---- base_c(k): the part of st_c that does not depend on the MV levels ----
        double base = xr;   // channel lanes (idle lanes: 0)
        if (kind == 1) {
            int pos = head + hlag;
            if (pos >= HL) pos -= HL;
            base = hrow[pos];
        } else if (kind == 2) base = setpoint(ei, ol ? nit - 1 : k, sigrow[ei]);
        if (nd > 0) {
            if (mdg != 0.0) base -= mdg * sigrow[2 * ny + mdj];
            if (kind == 2)
                for (int j = NU; j < nw; ++j) base -= sm.chg[ei * nw + j] * sigrow[2 * ny + (j - NU)];
        }
        // ---- outputs and cost of sample k (x(k) is in shared memory since the previous plant sample) ----
        if (!ol && lane < ny) {
            const int i = lane;
            double yi = 0.0, ysi = 0.0;
#pragma unroll 1
            for (int j = 0; j < nw; ++j) yi += yrow[j];
            if (want_ol) {
#pragma unroll 1
                for (int j = 0; j < nw; ++j) ysi += sm.xol[i * nw + j];
            }
            const bool mine = LEAN || (sel < 0 || sel == i);
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
        double st = base;
#pragma unroll
        for (int j = 0; j < NU; ++j) st = fma(-cu[j], wp.u[j], st);
        double du[NU];
        if (spec) {
            stb[(k - seg_k0) * NSTB + col] = st;   // idle lanes hold st = 0
#pragma unroll
            for (int j = 0; j < NU; ++j) du[j] = m0[j] * st;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int j = 0; j < NU; ++j) du[j] += __shfl_xor_sync(SIM_FULL, du[j], o);
        } else {
            if (col < nst) sm.st[col] = st;
            __syncwarp();
            const unsigned long long it_before = wp.n_it, con_before = wp.n_con;
#ifdef MPC_SIMT_EMULATION
            if (lane == 0) g_sim_verbose = (g_sim_knob >> 16) && k >= (g_sim_knob >> 16) && k < (g_sim_knob >> 16) + 2;
            if (g_sim_verbose && lane == 0) printf("=== sample %d\n", k);
            __syncwarp();
#endif
            const int rc = wp.controller_move();
            if (rc) status = rc;
            if (rc == SIM_ST_OVERFLOW) break;   // this run goes to the kernel with the spill area (second pass)
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
                ++k;
                continue;
            }
            if (out.trace && lane == 0) { out.trace[2 * k] = (int)(wp.n_it - it_before); out.trace[2 * k + 1] = wp.q; }
            cool = (wp.q == 0 && wp.n_con == con_before) ? (cool > 0 ? cool - 1 : 0) : SIM_SPEC_COOL;
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                const int r0 = j * P;
                du[j] = __shfl_sync(SIM_FULL, wp.z[r0 >> 5], r0 & 31);
            }
        }
        // ---- apply the first move of every input ----
#pragma unroll
        for (int j = 0; j < NU; ++j) wp.u[j] += du[j];
        if (!LEAN && !VLEAN && lane < NU) {
            const int j = lane;
            const bool mine = (sel < 0 || sel == j);
            if (out.u && mine) out.u[(size_t)j * nit + k] = sim_pick<NU>(wp.u, j);
            if (out.uopt && mine) out.uopt[(size_t)j * nit + k] = sm.uopt[j * P + (k < m ? k : m - 1)];
        }
        // ---- push w(k) (the slot being overwritten held lag HL-1, which nobody reads); MV levels also to the u log ----
        head = head == 0 ? HL - 1 : head - 1;
        if (lane == 0) {
            double *hp = sm.hist + head;
#pragma unroll
            for (int j = 0; j < NU; ++j) hp[j * HL] = wp.u[j];
            if (spec) {   // (the u log lives in the active-set factor's memory: speculative samples only)
                double *up = ub + (k - seg_k0 + 1) * MPC_MAXU;
#pragma unroll
                for (int j = 0; j < NU; ++j) up[j] = wp.u[j];
            }
        }
        if (nd > 0 && lane >= NU && lane < nw) sm.hist[lane * HL + head] = sigrow[2 * ny + (lane - NU)];
        __syncwarp();
        // ---- plant sample: x(k+1) = a x(k) + b0 w(k+1-d) + b1 w(k-d); lag q of the history is now w(k-q) ----
        if (kind == 0) {
            int p1 = head + l1; if (p1 >= HL) p1 -= HL;
            int p0 = head + l0; if (p0 >= HL) p0 -= HL;
            const double w1 = hrow[p1], w0 = hrow[p0];
            xr = ca * xr + cb0 * w0 + cb1 * w1;
            sm.x[col] = xr;
            if (want_ol) {
                const int cj = col % nw, cd = l1;
                double o1, o0;
                if (cj < NU) {
                    const int k1 = k - cd, k0 = k + 1 - cd;
                    o1 = k1 < 0 ? 0.0 : sm.uopt[cj * P + (k1 < m ? k1 : m - 1)];
                    o0 = (cd == 0 || k0 < 0) ? 0.0 : sm.uopt[cj * P + (k0 < m ? k0 : m - 1)];
                } else {  // the measured disturbance is the same signal in both simulations
                    o1 = w1; o0 = cd == 0 ? 0.0 : w0;
                }
                xolr = ca * xolr + cb0 * o0 + cb1 * o1;
                sm.xol[col] = xolr;
            }
        }
        __syncwarp();
        ++k;
    }
    // ---------------- costs ----------------
    if (out.cost) {
        if (mode == 1) {
            for (int i = 0; i < ny; ++i) {
                const double ci = __shfl_sync(SIM_FULL, cost_acc, i);
                if (lane == 0) out.cost[i] = status ? NAN : ci;
            }
        } else if (mode == 2) {
            const double tot = sim_wsum(cost_acc) + jnu;
            if (lane == 0) out.cost[0] = status ? NAN : tot;
        }
    }
    if (out.counters && lane == 0 && status != SIM_ST_OVERFLOW) {
        atomicAdd(out.counters + 0, wp.n_con);
        atomicAdd(out.counters + 1, wp.n_it);
    }
    if (out.diag && lane == 0) { out.diag[0] = wp.n_con; out.diag[1] = wp.n_it; out.diag[2] = (unsigned long long)wp.qmax | ((unsigned long long)sm.misc[SIM_PP_COUNT] << 32); }
    return status;
}
