// mpc_tables.h -- host-side problem ingest: layout + candidate-independent prediction tables.
#pragma once
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"
#include "mpc_layout.h"

struct MpcHostTables {
    MpcLayout L;
    std::vector<double> TG, TK, S1;
    std::vector<double> step;  // s_ij(n), [i][j][n], n = 0..pmax+mmax+1 (stride pmax+mmax+2)
    std::vector<double> pa;    // a_ch^n, [ch][n], n = 0..pmax
    std::vector<double> r, v, yref;
    std::vector<double> sig;   // r | yref | v packed sample-major (nit x (2ny+nd)): what the closed-loop kernel stages
    std::vector<int> dmin;
};

// Returns empty string on success, otherwise an error message.
std::string mpc_build_tables(const mpcgpu_problem &pb, MpcHostTables &out);
std::string mpc_set_signals(MpcHostTables &t, int nit, const double *r, const double *v, const double *yref);
