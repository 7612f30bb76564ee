// mpc_dtc.h -- layout + host tables of the DTC-GPC sweep (include/mpcgpu.h, DTC section).
#pragma once
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"
#include "mpc_layout.h"

#define DTC_MAXY 4
#define DTC_MAXU 4
#define DTC_MAXQ 4
#define DTC_MAXNA 5   /* na_i + 1 <= nu + 1 */
#define DTC_MAXCP 40  /* past-control columns per (output,input): dnz + nb */

struct DtcLayout {
    int ny, nu, nq, nit, pmax, mmax, k_start, hl;   // hl: history length needed for delays (max d + 2)
    int na[DTC_MAXY];                 // order of A_i
    int dmin[DTC_MAXY];               // min over inputs of descompMPC delay
    int dp[DTC_MAXY * DTC_MAXU];      // descompMPC delays
    int cp[DTC_MAXY * DTC_MAXU];      // dnz + nb: width of the past-control block
    int duM[DTC_MAXU], duoff[DTC_MAXU + 1];   // per-input width / offset of the past-control vector `up`
    int ydoff[DTC_MAXY + 1];          // offsets of the (na_i + 1) blocks of Yd
    // channels: model (scaled), process, disturbance
    double ma[DTC_MAXY * DTC_MAXU], mb0[DTC_MAXY * DTC_MAXU], mb1[DTC_MAXY * DTC_MAXU];
    int md[DTC_MAXY * DTC_MAXU];
    double pa[DTC_MAXY * DTC_MAXU], pb0[DTC_MAXY * DTC_MAXU], pb1[DTC_MAXY * DTC_MAXU];
    int pd[DTC_MAXY * DTC_MAXU];
    double qa[DTC_MAXY * DTC_MAXQ], qb0[DTC_MAXY * DTC_MAXQ], qb1[DTC_MAXY * DTC_MAXQ];
    int qd[DTC_MAXY * DTC_MAXQ];
    double L[DTC_MAXY], R[DTC_MAXU];
};

// Candidate-independent tables (all fp64):
//   step[i][j][n]          n = 0..pmax+dmax+1 : step response of model channel (i,j) incl. its delay (MatG.m:51)
//   ftab[i][row][col]      row = 1..pmax, col < na_i+1 : F polynomial rows of diophantine.m:55-65 (N1 = 1)
//   ug[i][j][row][col]     row = 1..pmax, col < cp_ij  : deltaUFree.m:36-57 rows (zero-stripped, right-aligned)
struct DtcHostTables {
    DtcLayout L;
    int step_len;
    std::vector<double> step, ftab, ug, r, q;
};

std::string dtc_build_tables(const mpcgpu_dtc_problem &pb, DtcHostTables &out);

struct DtcTables {
    const double *step, *ftab, *ug, *r, *q;
    int step_len;
};
