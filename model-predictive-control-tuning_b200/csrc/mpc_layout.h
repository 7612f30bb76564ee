// mpc_layout.h -- problem layout shared by the host table builder and the device kernels.
//
// State vector s (length ns) seen by the controller at decision time k (the linear map z_unc = M*s):
//   [0, ny*nw)            X    channel states x_ij(k), index i*nw+j       (y_i(k) = sum_j x_ij(k))
//   [hoff[j], +hlen[j])   HIST input history, slot q holds w_j(k-1-q); MVs: hlen>=1 (slot 0 = u(k-1))
//   [off_v, +nd)          VCUR measured disturbance v(k)                  (held over the horizon)
//   [off_r, +ny)          R    set-point r(k)                             (held over the horizon)
// Conventions restated from closedloop_toolbox.m:50 (`sim` without look-ahead), see oracle/mpc_oracle.c T3.
//
// The controller does not multiply s directly: z_unc is linear in s and vanishes at every steady state
// (x_ij = gain_ij*hv_j, history == held value hv_j, r_i = sum_j gain_ij*hv_j), so the product is taken
// in DEVIATION coordinates st (length nst = ns - nw) whose entries are all ~0 near steady state:
//   [0, ny*nw)                xt_ij = x_ij - gain_ij*hv_j         hv_j = u_j(k-1) for MVs, v_j(k) for MDs
//   [stoff_h[j], +hlen-hq0)   ht_jq = h_jq - hv_j                 q >= hq0[j] (MV slot 0 IS hv_j)
//   [stoff_e, +ny)            et_i  = r_i - sum_j gain_ij*hv_j
// This avoids the cancellation sum_sig M[:,sig]*s[sig] -> 0 that would otherwise cost cond(H)*eps*|s|.
#pragma once
#include <stdint.h>

#define MPC_MAXY 8
#define MPC_MAXU 4
#define MPC_MAXW 8
#define MPC_MAXM 15

#ifdef __CUDACC__
#define MPC_HD __host__ __device__ __forceinline__
#else
#define MPC_HD inline
#endif

struct MpcLayout {
    int ny, nu, nd, nw, nit, pmax, mmax, inK;
    int ns, nsig, off_v, off_r;
    int has_ov_bounds, nst, stoff_e, pad0;
    int hoff[MPC_MAXW], hlen[MPC_MAXW], hq0[MPC_MAXW], stoff_h[MPC_MAXW];
    int d[MPC_MAXY * MPC_MAXW];
    double a[MPC_MAXY * MPC_MAXW], b0[MPC_MAXY * MPC_MAXW], b1[MPC_MAXY * MPC_MAXW];
    double umin[MPC_MAXU], umax[MPC_MAXU], dumin[MPC_MAXU], dumax[MPC_MAXU];
    double su[MPC_MAXU], sy[MPC_MAXY];
    double gain[MPC_MAXY * MPC_MAXW];
    // soft output constraints (T4 of the oracle header): ymin_i - eps*emin_i <= y_i <= ymax_i + eps*emax_i,
    // emin/emax = OV(i).MinECR/MaxECR * ScaleFactor; rho_ecr = Weights.ECR
    double ymin[MPC_MAXY], ymax[MPC_MAXY], emin[MPC_MAXY], emax[MPC_MAXY];
    double rho_ecr;
};

// Candidate-independent prediction tables (device or host pointers, all fp64):
//   TG[i][D][L][j][j'] = sum_{n=1..L} s_ij(n+D) * s_ij'(n)        D<mmax, L<=pmax   ("prefix Gram")
//   TK[i][j][c][L][sig] = sum_{n=1..L} s_ij(n) * phi_{i,sig}(n+c)  c<mmax, L<=pmax, sig<nsig
//   S1[i][j][L]        = sum_{n=1..L} s_ij(n)
// with s_ij the step response of MV channel (i,j) and phi_{i,sig}(t) the free response of output i at
// k+t to a unit state component sig.  With them, for a candidate (p, m, wy2, wu2):
//   H[(c,j),(c',j')] = sum_i wy2_i * TG[i][c'-c][p-c'][j][j']  (c<=c')  + [e==e'] wu2_j
//   K[(c,j),sig]     = sum_i wy2_i * TK[i][j][c][p-c][sig] ;  K[(c,j),R_i] = -wy2_i * S1[i][j][p-c]
//   z_unc = -H^-1 K s
struct MpcTables {
    const double *TG, *TK, *S1;
    const double *r;     // nit x ny
    const double *v;     // nit x nd
    const double *yref;  // ny x nit
    const double *ST;    // step responses of the MV channels s_ij(n): [(i*nu+j)*st_stride + n]
    const double *PA;    // channel pole powers a_ch^n: [ch*(pmax+1) + n]
    int st_stride;
    const double *sig;   // the three signals packed sample-major: [k*(2ny+nd) + c], c = r (ny) | yref (ny) | v (nd)
};

static MPC_HD long long mpc_tg_index(const MpcLayout &L, int i, int D, int Lh, int j, int j2) {
    return ((((long long)i * L.mmax + D) * (L.pmax + 1) + Lh) * L.nu + j) * L.nu + j2;
}
static MPC_HD long long mpc_tk_index(const MpcLayout &L, int i, int j, int c, int Lh) {
    return ((((long long)i * L.nu + j) * L.mmax + c) * (L.pmax + 1) + Lh) * (long long)L.nsig;
}
static MPC_HD long long mpc_s1_index(const MpcLayout &L, int i, int j, int Lh) {
    return ((long long)i * L.nu + j) * (L.pmax + 1) + Lh;
}
