// mpc_core.cuh -- per-candidate prediction-matrix + Hessian builder (one CTA per candidate):
//   H = G'QG + R from the prefix-Gram tables, Cholesky, [M | W] = H^-1 [-K | I], written to HBM in the
//   row-padded layout the closed-loop kernel (mpc_sim.cuh) stages into shared memory.
//
// SINGLE SOURCE: compiles for sm_100a (nvcc, the product) and, with -DMPC_HOST_EMULATION, as
// thread-serialised host code used ONLY by tests/host_emulation.  The product never builds that.
#pragma once
#include <math.h>

#include "mpc_layout.h"

#ifdef MPC_HOST_EMULATION
#define MPC_FN static inline
#define TFOR(i, n) for (int i = 0; i < (n); ++i)
#define TSYNC() ((void)0)
#define IS_T0 (true)
#define MPC_LDG(p) (*(p))
#else
#define MPC_FN __device__ __forceinline__
#define TFOR(i, n) for (int i = (int)threadIdx.x; i < (n); i += (int)blockDim.x)
#define TSYNC() __syncthreads()
#define IS_T0 (threadIdx.x == 0u)
#define MPC_LDG(p) __ldg(p)
#endif

// ---------------------------------------------------------------------------------------------
// Builder: one CTA per candidate.
// Shared memory (doubles): Hs[nz*nz] | B[(nst+nz)*nz]   (B column-major: B[col*nz + row])
// Outputs (global): Mg[nst*R] = M in deviation coordinates, stored [col][row]; Wg[2*R*R] = H^-1 (padded) followed by its running row sums.
// ---------------------------------------------------------------------------------------------
static MPC_HD size_t mpc_builder_smem_doubles(int nz, int nst) { return (size_t)nz * nz + (size_t)(nst + nz) * nz; }

MPC_FN int mpc_build_candidate(const MpcLayout &L, const MpcTables &T, int p, int m, int P, const double *delta,
                               const double *lambda, double *smem, double *Mg, double *Wg, int *flag_smem) {
    const int ny = L.ny, nu = L.nu, nw = L.nw, nz = nu * m, ns = L.nst, ncol = ns + nz;
    double *Hs = smem;
    double *B = smem + (size_t)nz * nz;
    double wy2[MPC_MAXY], wu2[MPC_MAXU];
    for (int i = 0; i < ny; ++i) { double w = delta[i] / L.sy[i]; wy2[i] = w * w; }
    for (int j = 0; j < nu; ++j) { double w = lambda[j] / L.su[j]; wu2[j] = w * w; }
    if (IS_T0) *flag_smem = 0;
    // ---- H = G' Wy^2 G + Wdu^2 from the prefix-Gram table ----
    TFOR(idx, nz * nz) {
        const int e1 = idx / nz, e2 = idx - e1 * nz;
        const int c1 = e1 / nu, j1 = e1 - c1 * nu, c2 = e2 / nu, j2 = e2 - c2 * nu;
        // the move with the smaller horizon index carries the shifted step response
        const int ca = c1 <= c2 ? c1 : c2, cb = c1 <= c2 ? c2 : c1;
        const int ja = c1 <= c2 ? j1 : j2, jb = c1 <= c2 ? j2 : j1;
        const int D = cb - ca, Lh = p - cb;
        double acc = 0.0;
        if (Lh > 0)
            for (int i = 0; i < ny; ++i) acc += wy2[i] * MPC_LDG(T.TG + mpc_tg_index(L, i, D, Lh, ja, jb));
        if (e1 == e2) acc += wu2[j1];
        Hs[idx] = acc;
    }
    // ---- B = [-K | I], K columns in deviation-coordinate order (x block, history blocks, e block) ----
    TFOR(idx, nz * ny * nw) {
        const int e = idx / (ny * nw), sig = idx - e * (ny * nw);
        const int c = e / nu, j = e - c * nu, Lh = p - c;
        const int i = sig / nw;  // x_ij only reaches output i
        const double acc = wy2[i] * MPC_LDG(T.TK + mpc_tk_index(L, i, j, c, Lh) + sig);
        B[(size_t)sig * nz + e] = -acc;
    }
    for (int jw = 0; jw < nw; ++jw) {
        const int nq = L.hlen[jw] - L.hq0[jw];
        TFOR(idx, nz * nq) {
            const int e = idx / nq, qq = idx - e * nq;
            const int c = e / nu, j = e - c * nu, Lh = p - c;
            const int sig = L.hoff[jw] + L.hq0[jw] + qq;
            double acc = 0.0;
            for (int i = 0; i < ny; ++i) acc += wy2[i] * MPC_LDG(T.TK + mpc_tk_index(L, i, j, c, Lh) + sig);
            B[(size_t)(L.stoff_h[jw] + qq) * nz + e] = -acc;
        }
    }
    TFOR(idx, nz * ny) {
        const int e = idx / ny, i = idx - e * ny;
        const int c = e / nu, j = e - c * nu, Lh = p - c;
        B[(size_t)(L.stoff_e + i) * nz + e] = wy2[i] * MPC_LDG(T.S1 + mpc_s1_index(L, i, j, Lh));
    }
    TFOR(idx, nz * nz) {
        const int col = idx / nz, row = idx - col * nz;
        B[(size_t)(ns + col) * nz + row] = (row == col) ? 1.0 : 0.0;
    }
    TSYNC();
    // ---- Cholesky, right-looking, lower factor mirrored into the upper triangle ----
    for (int k = 0; k < nz; ++k) {
        if (IS_T0) {
            const double dkk = Hs[k * nz + k];
            if (!(dkk > 0.0)) *flag_smem = 1;
            Hs[k * nz + k] = sqrt(dkk > 0.0 ? dkk : 1.0);
        }
        TSYNC();
        const double inv = 1.0 / Hs[k * nz + k];
        TFOR(i2, nz - k - 1) {
            const int i = k + 1 + i2;
            const double l = Hs[i * nz + k] * inv;
            Hs[i * nz + k] = l;
            Hs[k * nz + i] = l;
        }
        TSYNC();
        const int rem = nz - k - 1;
        TFOR(idx, rem * rem) {
            const int a = idx / rem, b = idx - a * rem;
            if (b <= a) {
                const int i = k + 1 + a, j = k + 1 + b;
                Hs[i * nz + j] -= Hs[k * nz + i] * Hs[k * nz + j];
            }
        }
        TSYNC();
    }
    const int bad = *flag_smem;
    // ---- forward solve L Y = B (column sweep; all right-hand sides at once) ----
    for (int k = 0; k < nz; ++k) {
        const double inv = 1.0 / Hs[k * nz + k];
        TFOR(col, ncol) B[(size_t)col * nz + k] *= inv;
        TSYNC();
        const int rem = nz - k - 1;
        TFOR(idx, rem * ncol) {
            const int col = idx / rem, i = k + 1 + (idx - col * rem);
            B[(size_t)col * nz + i] -= Hs[k * nz + i] * B[(size_t)col * nz + k];
        }
        TSYNC();
    }
    // ---- backward solve L' X = Y ----
    for (int k = nz - 1; k >= 0; --k) {
        const double inv = 1.0 / Hs[k * nz + k];
        TFOR(col, ncol) B[(size_t)col * nz + k] *= inv;
        TSYNC();
        TFOR(idx, k * ncol) {
            const int col = idx / k, i = idx - col * k;
            B[(size_t)col * nz + i] -= Hs[k * nz + i] * B[(size_t)col * nz + k];
        }
        TSYNC();
    }
    // ---- write out, row-padded input-major: row r = j*P + c, rows with c >= m are zero ----
    {
        const int R = nu * P;   // P >= m: the padded horizon of the kernel image that will read this candidate
        TFOR(idx, ns * R) {
            const int col = idx / R, r = idx - col * R;
            const int j = r / P, c = r - j * P;
            Mg[idx] = (c < m) ? B[(size_t)col * nz + c * nu + j] : 0.0;
        }
        TFOR(idx, R * R) {
            const int r1 = idx / R, r2 = idx - r1 * R;
            const int j1 = r1 / P, c1 = r1 - j1 * P, j2 = r2 / P, c2 = r2 - j2 * P;
            double acc = 0.0, wr = 0.0;   // running sum over the horizon index of row block j1: W * (level normal)
            if (c2 < m)
                for (int cc = 0; cc <= c1 && cc < m; ++cc) {
                    const double v = B[(size_t)(ns + cc * nu + j1) * nz + (c2 * nu + j2)];
                    acc += v;
                    if (cc == c1) wr = v;
                }
            Wg[idx] = (c1 < m) ? wr : 0.0;
            Wg[(size_t)R * R + idx] = (c1 < m) ? acc : 0.0;
        }
    }
    TSYNC();
    return bad ? 3 : 0;
}

