// emu_dtc.cpp -- TEST-ONLY host build of the DTC-GPC kernels (csrc/mpc_dtc_kernel.cuh: k_dtc, one warp per candidate, four
// warps per CTA; k_dtc_filter, one thread per (candidate, output)), lanes as host threads (simt.h), for the not-gpu test-suite
// to compare with oracle/dtc_gpc_oracle.py.  A translation unit of its own (DTC_* macros).
#include "simt.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"
#include "../../model-predictive-control-tuning_b200/csrc/mpc_dtc_kernel.cuh"

// p: n x ny, m: n x nu, delta: n x ny, lambda: n x nu, alfa / raio: n (filters designed by dtc_filter_item, like
// mpcgpu_dtc_eval_batch_design).  ise: n x ny, y: n x ny x nit, u: n x nu x nit (may be NULL), status: n.
extern "C" int emu_dtc_eval(const mpcgpu_dtc_problem *pb, int n, const int *p, const int *m, const double *delta, const double *lambda,
                            const double *alfa, const double *raio, double *ise, double *y, double *u, int *status, char *err,
                            int errlen) {
    DtcHostTables ht;
    std::string e = dtc_build_tables(*pb, ht);
    if (!e.empty()) { std::strncpy(err, e.c_str(), errlen - 1); return 1; }
    const DtcLayout &L = ht.L;
    const int ny = L.ny;
    std::vector<double> fn((size_t)n * ny * DTC_MAXF), fd((size_t)n * ny * DTC_MAXF);
    std::vector<int> fl((size_t)n * ny * 2);
    for (int item = 0; item < n * ny; ++item) dtc_filter_item(L, n, alfa, raio, fn.data(), fd.data(), fl.data(), item);
    DtcTables T{ht.step.data(), ht.ftab.data(), ht.ug.data(), ht.r.data(), ht.q.data(), ht.step_len};
    DtcCand C{p, m, delta, lambda, fn.data(), fd.data(), fl.data(), ise, y, u, status};
    const size_t per = dtc_plan(L).doubles;
    for (int block = 0; block * DTC_WARPS < n; ++block) {
        std::vector<double> smem(per * DTC_WARPS + 8, std::nan(""));     // NaN-poisoned: reads of unwritten shared memory show
        simt_run_block([&]() { dtc_run(L, T, n, C, smem.data(), block); }, 32 * DTC_WARPS);
    }
    return 0;
}
