// mpc_dtc_tables.cpp -- candidate-independent polynomial preparation of the DTC-GPC sweep, written from
// the reference's MATLAB (written independently of the test oracle):
//   descompMPC.m:33-38   leading non-zero numerator coefficient => delay - 1, zero prepended
//   BA_MIMO.m:29-62      per-output common denominator from the poles ROUNDED to 4 decimals, numerators
//                        multiplied by the pole factors they lack
//   diophantine.m:35-79  f recursion on A~ = A*(1 - z^-1); F rows, e coefficients
//   deltaUFree.m:36-57   rows conv(E_i, B) with every exact zero removed, right-aligned into cp columns
//   MatG.m:51            step responses
// For the first-order channels of the reference `roots` of a product of (z - a_ij) is {a_ij}, so no
// numerical root finder is needed.
#include "mpc_dtc.h"

#include <algorithm>
#include <cmath>
#include <cstring>

static std::vector<double> conv(const std::vector<double> &a, const std::vector<double> &b) {
    std::vector<double> c(a.size() + b.size() - 1, 0.0);
    for (size_t i = 0; i < a.size(); ++i)
        for (size_t j = 0; j < b.size(); ++j) c[i + j] += a[i] * b[j];
    return c;
}
static double round4(double x) { return std::round(x * 1e4) / 1e4; }
static std::vector<double> poly_from_roots(const std::vector<double> &r) {
    std::vector<double> p{1.0};
    for (double x : r) p = conv(p, {1.0, -x});
    return p;
}

std::string dtc_build_tables(const mpcgpu_dtc_problem &pb, DtcHostTables &out) {
    DtcLayout &L = out.L;
    std::memset(&L, 0, sizeof(L));
    if (pb.ny < 1 || pb.ny > DTC_MAXY || pb.nu < 1 || pb.nu > DTC_MAXU || pb.nq < 0 || pb.nq > DTC_MAXQ)
        return "DTC-GPC: ny, nu <= 4 and nq <= 4 supported";
    if (pb.pmax < 1 || pb.pmax > 64 || pb.mmax < 1 || pb.mmax > 16) return "DTC-GPC: pmax <= 64, mmax <= 16";
    if (pb.nit < 2 || !pb.r) return "DTC-GPC: nit / r missing";
    const int ny = pb.ny, nu = pb.nu, nq = pb.nq;
    L.ny = ny; L.nu = nu; L.nq = nq; L.nit = pb.nit; L.pmax = pb.pmax; L.mmax = pb.mmax;
    L.k_start = pb.k_start > 0 ? pb.k_start : 4;
    int dmax = 0;
    for (int c = 0; c < ny * nu; ++c) {
        L.ma[c] = pb.ma[c]; L.mb0[c] = pb.mb0[c]; L.mb1[c] = pb.mb1[c]; L.md[c] = pb.md[c];
        L.pa[c] = pb.pa[c]; L.pb0[c] = pb.pb0[c]; L.pb1[c] = pb.pb1[c]; L.pd[c] = pb.pd[c];
        dmax = std::max(dmax, std::max(pb.md[c], pb.pd[c]));
        if (pb.md[c] < 0 || pb.pd[c] < 0 || pb.md[c] > 60 || pb.pd[c] > 60) return "DTC-GPC: delay out of range";
    }
    for (int c = 0; c < ny * nq; ++c) {
        L.qa[c] = pb.qa[c]; L.qb0[c] = pb.qb0[c]; L.qb1[c] = pb.qb1[c]; L.qd[c] = pb.qd[c];
        dmax = std::max(dmax, pb.qd[c]);
    }
    L.hl = dmax + 2;
    for (int i = 0; i < ny; ++i) L.L[i] = pb.L[i];
    for (int j = 0; j < nu; ++j) L.R[j] = pb.R[j];
    // descompMPC + BA_MIMO
    std::vector<std::vector<double>> Bn(ny * nu), B(ny * nu), A(ny);
    for (int i = 0; i < ny; ++i)
        for (int j = 0; j < nu; ++j) {
            const int c = i * nu + j;
            int d = pb.md[c];
            std::vector<double> num{pb.mb0[c], pb.mb1[c]};
            if (num[0] != 0.0) { d -= 1; num.insert(num.begin(), 0.0); }   // descompMPC.m:35-38
            L.dp[c] = d;
            if (num[0] == 0.0) num.erase(num.begin());                      // BA_MIMO.m:29-35
            Bn[c] = num;
        }
    for (int i = 0; i < ny; ++i) {
        int dm = 1 << 30;
        for (int j = 0; j < nu; ++j) dm = std::min(dm, L.dp[i * nu + j]);
        L.dmin[i] = dm;
        std::vector<double> poles;
        for (int j = 0; j < nu; ++j) poles.push_back(ny != 1 ? round4(pb.ma[i * nu + j]) : pb.ma[i * nu + j]);
        std::vector<double> uniq = poles;
        if (ny != 1) {                                                      // BA_MIMO.m:37-41
            std::sort(uniq.begin(), uniq.end());
            uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        }
        A[i] = poly_from_roots(uniq);
        L.na[i] = (int)A[i].size() - 1;
        if (L.na[i] + 1 > DTC_MAXNA) return "DTC-GPC: denominator order too large";
        for (int j = 0; j < nu; ++j) {                                      // BA_MIMO.m:45-62
            std::vector<double> rA;
            for (double x : uniq) rA.push_back(round4(x));
            const double rAn = round4(pb.ma[i * nu + j]);
            size_t kk = 0;
            while (kk < rA.size()) {                                        // the reference's scan-and-skip loop
                if (rA[kk] == rAn) {
                    const double val = rA[kk];
                    rA.erase(std::remove(rA.begin(), rA.end(), val), rA.end());
                }
                kk += 1;
            }
            B[i * nu + j] = conv(Bn[i * nu + j], poly_from_roots(rA));
        }
    }
    int off = 0;
    for (int i = 0; i < ny; ++i) { L.ydoff[i] = off; off += L.na[i] + 1; }
    L.ydoff[ny] = off;
    for (int j = 0; j < nu; ++j) L.duM[j] = 0;
    for (int i = 0; i < ny; ++i)
        for (int j = 0; j < nu; ++j) {
            const int c = i * nu + j;
            int cp = (L.dp[c] - L.dmin[i]) + (int)B[c].size() - 1;          // deltaUFree.m:25
            if (cp < 1) cp = 1;
            if (cp > DTC_MAXCP) return "DTC-GPC: past-control block too wide";
            L.cp[c] = cp;
            L.duM[j] = std::max(L.duM[j], (L.dp[c] - L.dmin[i]) + (int)B[c].size() - 1);   // DTC_GPC_WW.m:94
        }
    off = 0;
    for (int j = 0; j < nu; ++j) { L.duoff[j] = off; off += L.duM[j]; }
    L.duoff[nu] = off;
    const int P = L.pmax;
    // step responses of the model channels
    int dmx = 0;
    for (int i = 0; i < ny; ++i) dmx = std::max(dmx, L.dmin[i]);
    out.step_len = P + dmx + 2;
    out.step.assign((size_t)ny * nu * out.step_len, 0.0);
    for (int i = 0; i < ny; ++i)
        for (int j = 0; j < nu; ++j) {
            const int c = i * nu + j, dd = pb.md[c];
            double *s = &out.step[(size_t)c * out.step_len];
            for (int n = 1; n < out.step_len; ++n)
                s[n] = pb.ma[c] * s[n - 1] + (n - dd >= 0 ? pb.mb0[c] : 0.0) + (n - dd - 1 >= 0 ? pb.mb1[c] : 0.0);
        }
    // diophantine recursion (diophantine.m:35-79, d = 0)
    out.ftab.assign((size_t)ny * (P + 1) * DTC_MAXNA, 0.0);
    out.ug.assign((size_t)ny * nu * (P + 1) * DTC_MAXCP, 0.0);
    for (int i = 0; i < ny; ++i) {
        std::vector<double> AD = conv(A[i], {1.0, -1.0});
        const int nAD = (int)AD.size(), nf = nAD - 1;
        std::vector<std::vector<double>> f(P + 1, std::vector<double>(nf, 0.0));
        f[0][0] = 1.0;
        for (int j = 0; j < P; ++j) {
            for (int c = 0; c < nf - 1; ++c) f[j + 1][c] = f[j][c + 1] - f[j][0] * AD[c + 1];
            f[j + 1][nf - 1] = -f[j][0] * AD[nAD - 1];
        }
        for (int row = 1; row <= P; ++row)
            for (int c = 0; c < nf; ++c) out.ftab[((size_t)i * (P + 1) + row) * DTC_MAXNA + c] = f[row][c];
        // e(1) = 1, e(i) = f(i,1): row `row` of E is e(1..row)
        std::vector<double> e(P);
        e[0] = 1.0;
        for (int k = 1; k < P; ++k) e[k] = f[k][0];
        for (int j = 0; j < nu; ++j) {
            const int c = i * nu + j, cp = L.cp[c];
            for (int row = 1; row <= P; ++row) {
                std::vector<double> Erow(e.begin(), e.begin() + row);
                std::vector<double> aux = conv(Erow, B[c]);
                std::vector<double> BE;
                for (double v : aux) if (v != 0.0) BE.push_back(v);          // deltaUFree.m:39-45
                double *dst = &out.ug[(((size_t)c) * (P + 1) + row) * DTC_MAXCP];
                const int lBE = (int)BE.size();
                if (lBE < cp) for (int k = 0; k < lBE; ++k) dst[cp - lBE + k] = BE[k];
                else for (int k = 0; k < cp; ++k) dst[k] = BE[lBE - cp + k];
            }
        }
    }
    out.r.assign(pb.r, pb.r + (size_t)ny * pb.nit);
    if (nq > 0) { if (!pb.q) return "DTC-GPC: q missing"; out.q.assign(pb.q, pb.q + (size_t)nq * pb.nit); }
    else out.q.assign(1, 0.0);
    return "";
}

// Host-only view of the tables (no CUDA call): lets the CPU test-suite compare this polynomial preparation
// with the oracle's restatement of the MATLAB.  info[0..5] = {step_len, pmax, MAXNA, MAXCP, sum duM, sum(na+1)},
// info[8+i] = na_i, info[16+i] = dmin_i, info[24+i*nu+j] = cp_ij, info[48+j] = duM_j.  Any pointer may be NULL.
extern "C" int mpcgpu_dtc_host_tables(const mpcgpu_dtc_problem *problem, int32_t *info /*64*/, double *step,
                                      double *ftab, double *ug) {
    if (!problem) return MPCGPU_ERR_ARG;
    DtcHostTables t;
    if (!dtc_build_tables(*problem, t).empty()) return MPCGPU_ERR_ARG;
    const DtcLayout &L = t.L;
    if (info) {
        std::memset(info, 0, sizeof(int32_t) * 64);
        info[0] = t.step_len; info[1] = L.pmax; info[2] = DTC_MAXNA; info[3] = DTC_MAXCP;
        info[4] = L.duoff[L.nu]; info[5] = L.ydoff[L.ny];
        for (int i = 0; i < L.ny; ++i) { info[8 + i] = L.na[i]; info[16 + i] = L.dmin[i]; }
        for (int c = 0; c < L.ny * L.nu; ++c) info[24 + c] = L.cp[c];
        for (int j = 0; j < L.nu; ++j) info[48 + j] = L.duM[j];
    }
    if (step) std::memcpy(step, t.step.data(), sizeof(double) * t.step.size());
    if (ftab) std::memcpy(ftab, t.ftab.data(), sizeof(double) * t.ftab.size());
    if (ug) std::memcpy(ug, t.ug.data(), sizeof(double) * t.ug.size());
    return MPCGPU_OK;
}
