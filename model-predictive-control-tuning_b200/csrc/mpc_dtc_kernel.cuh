// mpc_dtc_kernel.cuh -- the device side of the DTC-GPC sweep (k_dtc: one warp per candidate; k_dtc_filter: robustness-filter
// design per candidate), split from mpc_dtc.cu so that tests/host_emulation can run the kernel bodies on a CPU (lanes as host
// threads).  The bodies are device functions; the __global__ entries are thin wrappers.
#pragma once
#include <math.h>

#include "mpc_dtc.h"
#define DTC_WARPS 4
#define DTC_FULL 0xffffffffu
#define DTC_MAXF MPCGPU_DTC_MAXF

struct DtcCand {
    const int *p, *m;                 // n x ny, n x nu
    const double *delta, *lambda;     // n x ny, n x nu
    const double *fr_num, *fr_den;    // n x ny x MAXF
    const int *fr_len;                // n x ny x 2
    double *ise, *y, *u;              // n x ny ; optional n x ny x nit, n x nu x nit
    int *status;
};

struct DtcSmemPlan { int SM, SP, DU, YD, HL, nch; size_t doubles; };

static __host__ __device__ inline DtcSmemPlan dtc_plan(const DtcLayout &L) {
    DtcSmemPlan P;
    P.SM = L.nu * L.mmax; P.SP = L.ny * L.pmax; P.DU = L.duoff[L.nu]; P.YD = L.ydoff[L.ny]; P.HL = L.hl;
    P.nch = 3 * L.ny * L.nu + L.ny * L.nq;
    size_t n = (size_t)P.SM * P.SM;        // S1 / Cholesky factor
    n += (size_t)L.nu * P.SM;              // X: rows of S1^-1
    n += (size_t)L.nu * P.SP;              // Km
    n += (size_t)L.nu * (P.DU + P.YD + L.ny);   // KHp, KS, Kr
    n += 2 * (size_t)L.nu * P.HL;          // u, ue histories
    n += P.DU;                             // up
    n += (size_t)L.ny * (DTC_MAXNA + 2 * DTC_MAXF);   // yp, eM, yfr histories
    n += P.nch;                            // channel states
    n += 2 * (size_t)L.ny * DTC_MAXF;      // filter coefficients
    n += L.nu;                             // dU
    P.doubles = n + 4;
    return P;
}

__device__ __forceinline__ double dtc_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DTC_FULL, v, o);
    return v;
}

__device__ __forceinline__ void dtc_run(const DtcLayout &L, const DtcTables &T, int n, const DtcCand &C, double *smem_d, int block) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = block * DTC_WARPS + warp;
    if (c >= n) return;
    const DtcSmemPlan PL = dtc_plan(L);
    const int ny = L.ny, nu = L.nu, nq = L.nq, nit = L.nit, SMx = PL.SM, SPx = PL.SP, DU = PL.DU, YD = PL.YD, HL = PL.HL;
    double *sp = smem_d + (size_t)warp * PL.doubles;
    double *S1 = sp; sp += (size_t)SMx * SMx;
    double *X = sp; sp += (size_t)nu * SMx;
    double *Km = sp; sp += (size_t)nu * SPx;
    double *KHp = sp; sp += (size_t)nu * DU;
    double *KS = sp; sp += (size_t)nu * YD;
    double *Kr = sp; sp += (size_t)nu * ny;
    double *hu = sp; sp += (size_t)nu * HL;
    double *hue = sp; sp += (size_t)nu * HL;
    double *up = sp; sp += DU;
    double *yph = sp; sp += (size_t)ny * DTC_MAXNA;
    double *emh = sp; sp += (size_t)ny * DTC_MAXF;
    double *yfh = sp; sp += (size_t)ny * DTC_MAXF;
    double *xch = sp; sp += PL.nch;
    double *fnum = sp; sp += (size_t)ny * DTC_MAXF;
    double *fden = sp; sp += (size_t)ny * DTC_MAXF;
    double *dUs = sp;

    // ---- candidate ----
    int p[DTC_MAXY], m[DTC_MAXU], poff[DTC_MAXY + 1], moff[DTC_MAXU + 1];
    double dl[DTC_MAXY], lm[DTC_MAXU];
    int bad = 0;
    poff[0] = 0; moff[0] = 0;
#pragma unroll
    for (int i = 0; i < DTC_MAXY; ++i) {
        p[i] = i < ny ? C.p[(size_t)c * ny + i] : 0;
        dl[i] = i < ny ? C.delta[(size_t)c * ny + i] : 0.0;
        if (i < ny && (p[i] < 1 || p[i] > L.pmax)) bad = 1;
        poff[i + 1] = poff[i] + p[i];
    }
#pragma unroll
    for (int j = 0; j < DTC_MAXU; ++j) {
        m[j] = j < nu ? C.m[(size_t)c * nu + j] : 0;
        lm[j] = j < nu ? C.lambda[(size_t)c * nu + j] : 0.0;
        if (j < nu && (m[j] < 1 || m[j] > L.mmax)) bad = 1;
        moff[j + 1] = moff[j] + m[j];
    }
    for (int i = 0; i < ny; ++i) {
        const int ln = C.fr_len[((size_t)c * ny + i) * 2], ld = C.fr_len[((size_t)c * ny + i) * 2 + 1];
        if (ln < 1 || ld < 1 || ln > DTC_MAXF || ld > DTC_MAXF || ln > ld) bad = 1;
    }
    if (bad) {
        if (lane == 0) {
            C.status[c] = MPCGPU_CAND_INVALID;
            for (int i = 0; i < ny; ++i) C.ise[(size_t)c * ny + i] = NAN;
        }
        return;
    }
    const int sm = moff[nu], spn = poff[ny];
    // (output, row) / (input, column) decoders
    auto out_of = [&](int row, int &i, int &r) { i = 0; while (i + 1 < ny && row >= poff[i + 1]) ++i; r = row - poff[i]; };
    auto in_of = [&](int col, int &j, int &k) { j = 0; while (j + 1 < nu && col >= moff[j + 1]) ++j; k = col - moff[j]; };
    auto step = [&](int i, int j, int nn) -> double { return T.step[(size_t)(i * nu + j) * T.step_len + nn]; };

    // ---- S1 = H'QH + W (upper triangle computed, mirrored) ----
    for (int e = lane; e < sm * sm; e += 32) {
        const int ca = e / sm, cb = e - ca * sm;
        if (ca > cb) continue;
        int ja, ka, jb, kb;
        in_of(ca, ja, ka); in_of(cb, jb, kb);
        double acc = 0.0;
        for (int i = 0; i < ny; ++i) {
            const int base = L.dmin[i] + 1;
            double a = 0.0;
            for (int r = (ka > kb ? ka : kb); r < p[i]; ++r) a = fma(step(i, ja, base + r - ka), step(i, jb, base + r - kb), a);
            acc = fma(dl[i], a, acc);
        }
        if (ca == cb) acc += lm[ja];
        S1[(size_t)ca * SMx + cb] = acc;
        S1[(size_t)cb * SMx + ca] = acc;
    }
    __syncwarp();
    // ---- Cholesky S1 = C C' (lower, in place), right-looking ----
    int notpd = 0;
    for (int k = 0; k < sm; ++k) {
        const double dkk = S1[(size_t)k * SMx + k];
        if (!(dkk > 0.0)) { notpd = 1; break; }
        const double ckk = sqrt(dkk);
        __syncwarp();
        for (int r = k + lane; r < sm; r += 32) S1[(size_t)r * SMx + k] = (r == k) ? ckk : S1[(size_t)r * SMx + k] / ckk;
        __syncwarp();
        const int nt = sm - k - 1;
        for (int e = lane; e < nt * nt; e += 32) {
            const int r = k + 1 + e / nt, cc = k + 1 + (e - (e / nt) * nt);
            if (cc <= r) S1[(size_t)r * SMx + cc] -= S1[(size_t)r * SMx + k] * S1[(size_t)cc * SMx + k];
        }
        __syncwarp();
    }
    if (notpd) {
        if (lane == 0) {
            C.status[c] = MPCGPU_CAND_NOT_PD;
            for (int i = 0; i < ny; ++i) C.ise[(size_t)c * ny + i] = NAN;
        }
        return;
    }
    // ---- X_j = S1^-1 e_{moff[j]}: lane j solves its own right-hand side (forward, then backward) ----
    if (lane < nu) {
        double *x = X + (size_t)lane * SMx;
        const int e0 = moff[lane];
        for (int r = 0; r < sm; ++r) {
            double acc = (r == e0) ? 1.0 : 0.0;
            for (int k = e0; k < r; ++k) acc -= S1[(size_t)r * SMx + k] * x[k];
            x[r] = r < e0 ? 0.0 : acc / S1[(size_t)r * SMx + r];
        }
        for (int r = sm - 1; r >= 0; --r) {
            double acc = x[r];
            for (int k = r + 1; k < sm; ++k) acc -= S1[(size_t)k * SMx + r] * x[k];
            x[r] = acc / S1[(size_t)r * SMx + r];
        }
    }
    __syncwarp();
    // ---- Km[j][(i,r)] = delta_i * sum_col X_j[col] * H[(i,r)][col] ----
    for (int e = lane; e < nu * spn; e += 32) {
        const int j = e / spn, row = e - j * spn;
        int i, r;
        out_of(row, i, r);
        const int base = L.dmin[i] + 1;
        double acc = 0.0;
        for (int jj = 0; jj < nu; ++jj) {
            const int kmax = r < m[jj] - 1 ? r : m[jj] - 1;
            for (int k = 0; k <= kmax; ++k) acc = fma(X[(size_t)j * SMx + moff[jj] + k], step(i, jj, base + r - k), acc);
        }
        Km[(size_t)j * SPx + row] = dl[i] * acc;
    }
    __syncwarp();
    // ---- Kr, KHp, KS ----
    for (int e = lane; e < nu * ny; e += 32) {
        const int j = e / ny, i = e - j * ny;
        double acc = 0.0;
        for (int r = 0; r < p[i]; ++r) acc += Km[(size_t)j * SPx + poff[i] + r];
        Kr[e] = acc;
    }
    for (int e = lane; e < nu * DU; e += 32) {
        const int j = e / DU, col = e - j * DU;
        int jj = 0;
        while (jj + 1 < nu && col >= L.duoff[jj + 1]) ++jj;
        const int t = col - L.duoff[jj];
        double acc = 0.0;
        for (int i = 0; i < ny; ++i) {
            if (t >= L.cp[i * nu + jj]) continue;                 // cell2mat2.m:45-55: block sits top-left
            const double *ug = T.ug + ((size_t)(i * nu + jj) * (L.pmax + 1)) * DTC_MAXCP + t;
            for (int r = 0; r < p[i]; ++r) acc = fma(Km[(size_t)j * SPx + poff[i] + r], ug[(size_t)(r + 1) * DTC_MAXCP], acc);
        }
        KHp[e] = acc;
    }
    for (int e = lane; e < nu * YD; e += 32) {
        const int j = e / YD, col = e - j * YD;
        int i = 0;
        while (i + 1 < ny && col >= L.ydoff[i + 1]) ++i;
        const int cc = col - L.ydoff[i];
        const double *ft = T.ftab + ((size_t)i * (L.pmax + 1)) * DTC_MAXNA + cc;
        double acc = 0.0;
        for (int r = 0; r < p[i]; ++r) acc = fma(Km[(size_t)j * SPx + poff[i] + r], ft[(size_t)(r + 1) * DTC_MAXNA], acc);
        KS[e] = acc;
    }
    // ---- loop state ----
    for (int e = lane; e < nu * HL; e += 32) { hu[e] = 0.0; hue[e] = 0.0; }
    for (int e = lane; e < DU; e += 32) up[e] = 0.0;
    for (int e = lane; e < ny * DTC_MAXNA; e += 32) yph[e] = 0.0;
    for (int e = lane; e < ny * DTC_MAXF; e += 32) {
        emh[e] = 0.0; yfh[e] = 0.0;
        fnum[e] = C.fr_num[(size_t)c * ny * DTC_MAXF + e];
        fden[e] = C.fr_den[(size_t)c * ny * DTC_MAXF + e];
    }
    for (int e = lane; e < PL.nch; e += 32) xch[e] = 0.0;
    __syncwarp();
    const int nyu = ny * nu, nyq = ny * nq;
    double ise = 0.0;
    for (int k = 0; k < nit; ++k) {
        // channels: [0,nyu) process (input u), [nyu, nyu+nyq) disturbance (input q), then model and fast model (input ue)
        for (int ch = lane; ch < PL.nch; ch += 32) {
            double a, b0, b1, w0 = 0.0, w1 = 0.0;
            int d;
            if (ch < nyu) {
                a = L.pa[ch]; b0 = L.pb0[ch]; b1 = L.pb1[ch]; d = L.pd[ch];
                const int j = ch % nu;
                if (d >= 1 && k - d >= 0) w0 = hu[j * HL + (k - d) % HL];
                if (k - d - 1 >= 0) w1 = hu[j * HL + (k - d - 1) % HL];
            } else if (ch < nyu + nyq) {
                const int cq = ch - nyu, j = cq % nq;
                a = L.qa[cq]; b0 = L.qb0[cq]; b1 = L.qb1[cq]; d = L.qd[cq];
                if (k - d >= 0) w0 = T.q[(size_t)j * nit + (k - d)];
                if (k - d - 1 >= 0) w1 = T.q[(size_t)j * nit + (k - d - 1)];
            } else {
                const int fastm = ch >= nyu + nyq + nyu;
                const int cm = ch - nyu - nyq - (fastm ? nyu : 0), j = cm % nu, i = cm / nu;
                a = L.ma[cm]; b0 = L.mb0[cm]; b1 = L.mb1[cm];
                d = L.md[cm] - (fastm ? L.dmin[i] : 0);              // Gnz: delays minus the row minimum (DTC_GPC_WW.m:50-52)
                if (d >= 1 && k - d >= 0) w0 = hue[j * HL + (k - d) % HL];
                if (k - d - 1 >= 0) w1 = hue[j * HL + (k - d - 1) % HL];   // d >= 0: iodelay >= dp >= dmin
            }
            xch[ch] = a * xch[ch] + b0 * w0 + b1 * w1;
        }
        __syncwarp();
        if (lane < ny) {
            const int i = lane;
            double yi = 0.0, ypz = 0.0, ygz = 0.0;
            for (int j = 0; j < nu; ++j) {
                yi += xch[i * nu + j];
                ypz += xch[nyu + nyq + i * nu + j];
                ygz += xch[nyu + nyq + nyu + i * nu + j];
            }
            for (int j = 0; j < nq; ++j) yi += xch[nyu + i * nq + j];
            const double ye = L.L[i] * yi;
            const double em = ye - ypz;
            // Fr = Nr(z)/Dr(z) on eM (OptimalPredictor2.m:37): histories hold lags 1..MAXF-1 at [1..]
            double *eh = emh + i * DTC_MAXF, *fh = yfh + i * DTC_MAXF;
            const int nN = C.fr_len[((size_t)c * ny + i) * 2], nD = C.fr_len[((size_t)c * ny + i) * 2 + 1];
            for (int l = DTC_MAXF - 1; l >= 1; --l) { eh[l] = eh[l - 1]; fh[l] = fh[l - 1]; }
            eh[0] = em;
            const int off = nD - nN;
            double acc = 0.0;
            for (int l = 0; l < nN; ++l) acc += fnum[i * DTC_MAXF + l] * eh[off + l];
            for (int l = 1; l < nD; ++l) acc -= fden[i * DTC_MAXF + l] * fh[l];
            const double yfr = acc / fden[i * DTC_MAXF];
            fh[0] = yfr;
            const double yp = ygz + yfr;
            double *ph = yph + i * DTC_MAXNA;
            for (int l = DTC_MAXNA - 1; l >= 1; --l) ph[l] = ph[l - 1];
            ph[0] = yp;
            const double ri = T.r[(size_t)i * nit + k];
            const double er = yi - ri;
            ise += er * er;
            if (C.y) C.y[((size_t)c * ny + i) * nit + k] = yi;
        }
        __syncwarp();
        if (lane < nu) {
            const int j = lane;
            double du = 0.0;
            if (k + 1 >= L.k_start) {
                double acc = 0.0;
                for (int i = 0; i < ny; ++i) acc = fma(Kr[j * ny + i], L.L[i] * T.r[(size_t)i * nit + k], acc);
                for (int t = 0; t < DU; ++t) acc = fma(-KHp[j * DU + t], up[t], acc);
                for (int i = 0; i < ny; ++i)
                    for (int cc = 0; cc <= L.na[i]; ++cc) acc = fma(-KS[j * YD + L.ydoff[i] + cc], yph[i * DTC_MAXNA + cc], acc);
                du = acc;
            }
            dUs[j] = du;
        }
        __syncwarp();
        if (lane < nu) {
            const int j = lane;
            const double uprev = k > 0 ? hue[j * HL + (k - 1) % HL] : 0.0;
            double uek = 0.0;
            if (k + 1 >= L.k_start) {
                for (int t = L.duM[j] - 1; t >= 1; --t) up[L.duoff[j] + t] = up[L.duoff[j] + t - 1];
                if (L.duM[j] > 0) up[L.duoff[j]] = dUs[j];
                uek = uprev + dUs[j];
            }
            hue[j * HL + k % HL] = uek;
            const double uk = L.R[j] * uek;
            hu[j * HL + k % HL] = uk;
            if (C.u) C.u[((size_t)c * nu + j) * nit + k] = uk;
        }
        __syncwarp();
    }
    if (lane < ny) C.ise[(size_t)c * ny + lane] = ise;
    if (lane == 0) C.status[c] = MPCGPU_CAND_OK;
}

// ------------------------------------------------------------------------------------------------
// Batched robustness-filter design (DTC-GPC/mimofilter.m:33-50, filtro_siso.m:26-96): one thread per (candidate, output).
// Fr = Nr/Dr with Dr = (z - alfa)^ns, ns = number of slow poles of the output's delay-free row model (|pole| >= raio), and
// Nr = the remainder of Dr z^d by px = (z - 1) prod(z - slow poles): the reference's Sylvester system Dr z^d = Nr + px Q with
// deg Nr < deg px, solved as the polynomial division it states (d = the row's minimum dead time >= 1).  No slow pole: Fr = 1.
// Writes fr_num / fr_den (descending powers, zero padded to DTC_MAXF) and fr_len, the inputs of k_dtc.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dtc_filter_item(const DtcLayout &L, int n, const double *alfa, const double *raio, double *fr_num,
                                                double *fr_den, int *fr_len, int item) {
    if (item >= n * L.ny) return;
    const int c = item / L.ny, i = item - c * L.ny;
    double *num = fr_num + ((size_t)c * L.ny + i) * DTC_MAXF, *den = fr_den + ((size_t)c * L.ny + i) * DTC_MAXF;
    for (int e = 0; e < DTC_MAXF; ++e) { num[e] = 0.0; den[e] = 0.0; }
    double slow[DTC_MAXU];
    int ns = 0, d = 1 << 30;
    for (int j = 0; j < L.nu; ++j) {
        const int ch = i * L.nu + j;
        if (L.md[ch] < d) d = L.md[ch];                                    // Pd.iodelay of the row: its minimum dead time (mimofilter.m:25-29)
        if (L.mb0[ch] + L.mb1[ch] != 0.0 && fabs(L.ma[ch]) >= raio[c]) slow[ns++] = L.ma[ch];
    }
    if (ns == 0) { num[0] = 1.0; den[0] = 1.0; fr_len[2 * item] = 1; fr_len[2 * item + 1] = 1; return; }   // filtro_siso.m:90-91
    const int lpx = ns + 2;                                                  // px = (z - 1) prod (z - slow): degree ns + 1
    if (d < 1 || lpx - 1 > DTC_MAXF || ns + 1 + d > 4 * DTC_MAXF) { fr_len[2 * item] = 0; fr_len[2 * item + 1] = 0; return; }   // k_dtc flags the candidate MPCGPU_CAND_INVALID
    double px[DTC_MAXU + 2], Dr[DTC_MAXU + 1], w[4 * DTC_MAXF];
    px[0] = 1.0; px[1] = -1.0;
    for (int k = 0; k < ns; ++k) {                                           // px *= (z - slow_k)
        px[k + 2] = 0.0;
        for (int e = k + 2; e >= 1; --e) px[e] -= slow[k] * px[e - 1];
    }
    Dr[0] = 1.0;
    for (int k = 0; k < ns; ++k) {                                           // Dr *= (z - alfa)
        Dr[k + 1] = 0.0;
        for (int e = k + 1; e >= 1; --e) Dr[e] -= alfa[c] * Dr[e - 1];
    }
    const int lw = ns + 1 + d;                                               // Dr z^d, descending powers
    for (int e = 0; e < lw; ++e) w[e] = e <= ns ? Dr[e] : 0.0;
    for (int e = 0; e + lpx <= lw; ++e) {                                    // long division by the monic px: the remainder is left in w's tail
        const double qk = w[e];
        for (int k2 = 0; k2 < lpx; ++k2) w[e + k2] -= qk * px[k2];
    }
    const int lr = lpx - 1;                                                  // deg Nr < deg px
    for (int e = 0; e < lr; ++e) num[e] = lw >= lr ? w[lw - lr + e] : 0.0;
    for (int e = 0; e <= ns; ++e) den[e] = Dr[e];
    fr_len[2 * item] = lr; fr_len[2 * item + 1] = ns + 1;
}

#ifndef MPC_SIMT_EMULATION
__global__ void __launch_bounds__(32 * DTC_WARPS) k_dtc(const DtcLayout L, const DtcTables T, int n, DtcCand C) {
    extern __shared__ double smem_d[];
    dtc_run(L, T, n, C, smem_d, (int)blockIdx.x);
}
__global__ void k_dtc_filter(const DtcLayout L, int n, const double *alfa, const double *raio, double *fr_num, double *fr_den,
                             int *fr_len, int *status) {
    (void)status;
    dtc_filter_item(L, n, alfa, raio, fr_num, fr_den, fr_len, (int)(blockIdx.x * blockDim.x + threadIdx.x));
}
#endif
