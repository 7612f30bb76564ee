// mpc_tables.cpp -- candidate-independent part of the prediction-matrix builder (north-star item 1).
//
// The reference never materialises these: the Toolbox rebuilds its QP matrices inside every
// closedloop_toolbox call (closedloop_toolbox.m:36-43 dirty the object).  The explicit G-matrix math it
// mirrors is DTC-GPC/MatG.m:49-73 (G[r,c] = s(d+1+r-c)); here the p-dependence of G'QG and of the
// free-response map is folded into prefix sums over the horizon so that a candidate's Hessian costs
// O(nz^2 * ny) table look-ups instead of O(ny * p * nz^2) flops (SURVEY.md §7 "prefix Grams").
#include "mpc_tables.h"

#include <cmath>
#include <cstring>

static inline double chan_step(const mpcgpu_problem &pb, int nw, int i, int j, double xprev, int n) {
    const int ch = i * nw + j, dd = pb.d[ch];
    return pb.a[ch] * xprev + (n - dd >= 0 ? pb.b0[ch] : 0.0) + (n - dd - 1 >= 0 ? pb.b1[ch] : 0.0);
}

std::string mpc_set_signals(MpcHostTables &t, int nit, const double *r, const double *v, const double *yref) {
    MpcLayout &L = t.L;
    if (nit < 1) return "nit must be >= 1";
    if (!r) return "r is NULL";
    if (L.nd > 0 && !v) return "v is NULL but the plant has measured disturbances";
    L.nit = nit;
    t.r.assign(r, r + (size_t)nit * L.ny);
    if (L.nd > 0) t.v.assign(v, v + (size_t)nit * L.nd); else t.v.assign(1, 0.0);
    if (yref) t.yref.assign(yref, yref + (size_t)nit * L.ny); else t.yref.assign((size_t)nit * L.ny, 0.0);
    const int nsg = 2 * L.ny + L.nd;
    t.sig.assign((size_t)nit * nsg, 0.0);
    for (int k = 0; k < nit; ++k) {
        for (int i = 0; i < L.ny; ++i) {
            t.sig[(size_t)k * nsg + i] = t.r[(size_t)k * L.ny + i];
            t.sig[(size_t)k * nsg + L.ny + i] = t.yref[(size_t)i * nit + k];
        }
        for (int d = 0; d < L.nd; ++d) t.sig[(size_t)k * nsg + 2 * L.ny + d] = t.v[(size_t)k * L.nd + d];
    }
    return "";
}

std::string mpc_build_tables(const mpcgpu_problem &pb, MpcHostTables &out) {
    MpcLayout &L = out.L;
    std::memset(&L, 0, sizeof(L));
    if (pb.ny < 1 || pb.ny > MPC_MAXY) return "ny out of range (1..8)";
    if (pb.nu < 1 || pb.nu > MPC_MAXU) return "nu out of range (1..4)";
    if (pb.nd < 0 || pb.nu + pb.nd > MPC_MAXW) return "nu+nd out of range (<=8)";
    if (pb.pmax < 2 || pb.pmax > MPCGPU_MAX_P) return "pmax out of range (2..255)";
    if (pb.mmax < 1 || pb.mmax > MPCGPU_MAX_M) return "mmax out of range (1..15)";
    if (!pb.a || !pb.b0 || !pb.b1 || !pb.d || !pb.umin || !pb.umax || !pb.dumin || !pb.dumax || !pb.su || !pb.sy)
        return "NULL plant / limit array";
    const int ny = pb.ny, nu = pb.nu, nd = pb.nd, nw = nu + nd;
    L.ny = ny; L.nu = nu; L.nd = nd; L.nw = nw; L.pmax = pb.pmax; L.mmax = pb.mmax;
    L.inK = pb.inK > 0 ? pb.inK : 10;
    for (int c = 0; c < ny * nw; ++c) {
        if (pb.d[c] < 0 || pb.d[c] > 64) return "channel delay out of range (0..64 samples)";
        if (pb.d[c] == 0 && pb.b0[c] != 0.0) return "direct feed-through channel (d == 0 with b0 != 0) is not a valid MPC plant";
        if (!(std::fabs(pb.a[c]) < 1.0)) return "channel pole |a| >= 1: only stable first-order channels are supported";
        L.d[c] = pb.d[c]; L.a[c] = pb.a[c]; L.b0[c] = pb.b0[c]; L.b1[c] = pb.b1[c];
    }
    for (int j = 0; j < nu; ++j) {
        L.umin[j] = pb.umin[j]; L.umax[j] = pb.umax[j]; L.dumin[j] = pb.dumin[j]; L.dumax[j] = pb.dumax[j];
        L.su[j] = pb.su[j];
        if (!(pb.su[j] > 0)) return "MV ScaleFactor must be > 0";
    }
    L.has_ov_bounds = 0;
    L.rho_ecr = pb.rho_ecr;
    for (int i = 0; i < ny; ++i) {
        L.ymin[i] = pb.ymin ? pb.ymin[i] : -INFINITY;
        L.ymax[i] = pb.ymax ? pb.ymax[i] : INFINITY;
        L.emin[i] = (pb.ecr_min ? pb.ecr_min[i] : 1.0) * pb.sy[i];
        L.emax[i] = (pb.ecr_max ? pb.ecr_max[i] : 1.0) * pb.sy[i];
        L.sy[i] = pb.sy[i];
        if (!(pb.sy[i] > 0)) return "OV ScaleFactor must be > 0";
        if (pb.ymin && std::isfinite(pb.ymin[i])) L.has_ov_bounds = 1;
        if (pb.ymax && std::isfinite(pb.ymax[i])) L.has_ov_bounds = 1;
    }
    // state-vector layout
    int off = ny * nw;
    for (int j = 0; j < nw; ++j) {
        int dm = 0;
        for (int i = 0; i < ny; ++i) dm = pb.d[i * nw + j] > dm ? pb.d[i * nw + j] : dm;
        L.hlen[j] = (j < nu) ? (dm > 1 ? dm : 1) : dm;
        L.hoff[j] = off;
        off += L.hlen[j];
    }
    L.off_v = off; off += nd;
    L.nsig = off;
    L.off_r = off; off += ny;
    L.ns = off;
    // deviation coordinates
    int so = ny * nw;
    for (int j = 0; j < nw; ++j) {
        L.hq0[j] = (j < nu) ? 1 : 0;
        L.stoff_h[j] = so;
        so += L.hlen[j] - L.hq0[j];
    }
    L.stoff_e = so; so += ny;
    L.nst = so;
    for (int c = 0; c < ny * nw; ++c) L.gain[c] = (pb.b0[c] + pb.b1[c]) / (1.0 - pb.a[c]);
    out.dmin.assign(ny, 0);
    if (pb.dmin) for (int i = 0; i < ny; ++i) out.dmin[i] = pb.dmin[i];

    const int P = L.pmax, Mx = L.mmax, T = P + Mx + 1;
    // step responses of the MV channels
    out.step.assign((size_t)ny * nu * (T + 1), 0.0);
    for (int i = 0; i < ny; ++i)
        for (int j = 0; j < nu; ++j) {
            double *s = &out.step[((size_t)i * nu + j) * (T + 1)];
            s[0] = 0.0;
            for (int n = 1; n <= T; ++n) s[n] = chan_step(pb, nw, i, j, s[n - 1], n);
        }
    auto S = [&](int i, int j, int n) -> double { return out.step[((size_t)i * nu + j) * (T + 1) + n]; };
    out.pa.assign((size_t)ny * nw * (P + 1), 1.0);
    for (int c = 0; c < ny * nw; ++c)
        for (int n = 1; n <= P; ++n) out.pa[(size_t)c * (P + 1) + n] = out.pa[(size_t)c * (P + 1) + n - 1] * pb.a[c];
    // S1
    out.S1.assign((size_t)ny * nu * (P + 1), 0.0);
    for (int i = 0; i < ny; ++i)
        for (int j = 0; j < nu; ++j) {
            double acc = 0.0;
            for (int Lh = 1; Lh <= P; ++Lh) { acc += S(i, j, Lh); out.S1[mpc_s1_index(L, i, j, Lh)] = acc; }
        }
    // prefix Grams
    out.TG.assign((size_t)ny * Mx * (P + 1) * nu * nu, 0.0);
    for (int i = 0; i < ny; ++i)
        for (int D = 0; D < Mx; ++D)
            for (int j = 0; j < nu; ++j)
                for (int j2 = 0; j2 < nu; ++j2) {
                    double acc = 0.0;
                    for (int Lh = 1; Lh <= P; ++Lh) {
                        acc += S(i, j, Lh + D) * S(i, j2, Lh);
                        out.TG[mpc_tg_index(L, i, D, Lh, j, j2)] = acc;
                    }
                }
    // free-response basis phi_{i,sig}(t), t = 1..T, by simulating each unit state component
    const int nsig = L.nsig;
    std::vector<double> phi((size_t)ny * nsig * (T + 1), 0.0);
    std::vector<double> sv(L.ns);
    for (int sig = 0; sig < nsig; ++sig) {
        std::fill(sv.begin(), sv.end(), 0.0);
        sv[sig] = 1.0;
        for (int i = 0; i < ny; ++i)
            for (int j = 0; j < nw; ++j) {
                const int ch = i * nw + j, dd = pb.d[ch];
                const double hv = (j < nu) ? sv[L.hoff[j]] : sv[L.off_v + (j - nu)];
                double xf = sv[ch];
                for (int t = 1; t <= T; ++t) {
                    const int r0 = t - dd, r1 = t - dd - 1;
                    const double w0 = r0 >= 0 ? hv : sv[L.hoff[j] + (-r0 - 1)];
                    const double w1 = r1 >= 0 ? hv : sv[L.hoff[j] + (-r1 - 1)];
                    xf = pb.a[ch] * xf + pb.b0[ch] * w0 + pb.b1[ch] * w1;
                    phi[((size_t)i * nsig + sig) * (T + 1) + t] += xf;
                }
            }
    }
    out.TK.assign((size_t)ny * nu * Mx * (P + 1) * nsig, 0.0);
    for (int i = 0; i < ny; ++i)
        for (int j = 0; j < nu; ++j)
            for (int c = 0; c < Mx; ++c)
                for (int sig = 0; sig < nsig; ++sig) {
                    const double *ph = &phi[((size_t)i * nsig + sig) * (T + 1)];
                    double acc = 0.0;
                    for (int Lh = 1; Lh <= P; ++Lh) {
                        acc += S(i, j, Lh) * ph[Lh + c];
                        out.TK[mpc_tk_index(L, i, j, c, Lh) + sig] = acc;
                    }
                }
    if (pb.nit > 0 && pb.r) {
        std::string e = mpc_set_signals(out, pb.nit, pb.r, pb.v, pb.yref);
        if (!e.empty()) return e;
    } else {
        return "problem.r / nit missing";
    }
    return "";
}
