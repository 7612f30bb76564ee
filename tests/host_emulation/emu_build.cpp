// emu_build.cpp -- TEST-ONLY: the builder (csrc/mpc_core.cuh, k_build's body) executed by 128 host threads with real block
// barriers, next to the serial host build emu.cpp uses (MPC_HOST_EMULATION).  Same source, so M and W must come out
// bit-identical; under ThreadSanitizer (tsan.sh) it checks the builder's __syncthreads placement.
#include "simt.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"
#include "../../model-predictive-control-tuning_b200/csrc/mpc_core.cuh"
#include "../../model-predictive-control-tuning_b200/csrc/mpc_tables.h"

// M: nst x R ([col][row]), W: 2R x R, for one candidate on the P-padded layout.  Returns the builder's status.
extern "C" int emu_build_threads(const mpcgpu_problem *pb, int p, int m, int P, const double *delta, const double *lambda, double *Mg,
                                 double *Wg, int nthreads, char *err, int errlen) {
    MpcHostTables ht;
    std::string e = mpc_build_tables(*pb, ht);
    if (!e.empty()) { std::strncpy(err, e.c_str(), errlen - 1); return -1; }
    const MpcLayout &L = ht.L;
    MpcTables T{ht.TG.data(), ht.TK.data(), ht.S1.data(), ht.r.data(), ht.v.data(), ht.yref.data(),
                ht.step.data(), ht.pa.data(), L.pmax + L.mmax + 2, ht.sig.data()};
    std::vector<double> smem(mpc_builder_smem_doubles(L.nu * m, L.nst) + 8, std::nan(""));
    int flag = 0;
    std::vector<int> st((size_t)nthreads, -1);
    simt_run_block([&]() { st[threadIdx.x] = mpc_build_candidate(L, T, p, m, P, delta, lambda, smem.data(), Mg, Wg, &flag); }, (unsigned)nthreads);
    for (int t = 1; t < nthreads; ++t)
        if (st[t] != st[0]) return 99;
    return st[0];
}
