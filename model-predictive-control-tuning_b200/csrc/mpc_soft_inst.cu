// mpc_soft_inst.cu -- instantiates k_soft<SIM_NU, P> (plants with soft output constraints) for P = 4, 8, 16, and the
// validation-run image k_soft<SIM_NU, 16, true> (mismatched plant + state estimator, any linear plant).
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

sim_kernel_t SIM_CAT(soft_est_kernel_nu, SIM_NU)() { return k_soft<SIM_NU, 16, true>; }
