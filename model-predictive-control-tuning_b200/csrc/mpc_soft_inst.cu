// mpc_soft_inst.cu -- instantiates k_soft<SIM_NU, P> (plants with soft output constraints) for P = 4, 8, 16.
// Compiled with -fmad=false: the band-constraint QPs of Shell7x5 are degenerate enough that the pivot sequence of
// the active-set method depends on the last bit of the slacks; without implicit multiply-add contraction the
// kernel rounds like plain IEEE fp64 code (explicit fma() calls in the dot products stay fused).  DESIGN.md §2.
#include "mpc_sim_kernel.cuh"

#ifndef SIM_NU
#error "compile with -DSIM_NU=1..4"
#endif
#define SIM_CAT2(a, b) a##b
#define SIM_CAT(a, b) SIM_CAT2(a, b)

sim_kernel_t SIM_CAT(soft_kernel_nu, SIM_NU)(int P) {
    switch (P) {
        case 4: return k_soft<SIM_NU, 4>;
        case 8: return k_soft<SIM_NU, 8>;
        case 16: return k_soft<SIM_NU, 16>;
    }
    return nullptr;
}
