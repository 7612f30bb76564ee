// mpc_sim_inst.cu -- instantiates k_sim<SIM_NU, P> for P = 4, 8, 16 (compiled once per SIM_NU = 1..4).
#include "mpc_sim_kernel.cuh"

#ifndef SIM_NU
#error "compile with -DSIM_NU=1..4"
#endif
#define SIM_CAT2(a, b) a##b
#define SIM_CAT(a, b) SIM_CAT2(a, b)

sim_kernel_t SIM_CAT(sim_kernel_nu, SIM_NU)(int P) {
    switch (P) {
        case 4: return k_sim<SIM_NU, 4>;
        case 8: return k_sim<SIM_NU, 8>;
        case 16: return k_sim<SIM_NU, 16>;
    }
    return nullptr;
}

// GAM cost-only specialisation (see sim_run)
sim_kernel_t SIM_CAT(sim_lean_nu, SIM_NU)(int P) {
    switch (P) {
        case 4: return k_sim<SIM_NU, 4, true>;
        case 8: return k_sim<SIM_NU, 8, true>;
        case 16: return k_sim<SIM_NU, 16, true>;
    }
    return nullptr;
}

// VNS cost-only specialisation (see sim_run)
sim_kernel_t SIM_CAT(sim_vlean_nu, SIM_NU)(int P) {
    switch (P) {
        case 4: return k_sim<SIM_NU, 4, false, true>;
        case 8: return k_sim<SIM_NU, 8, false, true>;
        case 16: return k_sim<SIM_NU, 16, false, true>;
    }
    return nullptr;
}

// speculative kernel (mpc_sim_spec.cuh): the single P = 16 image, full / GAM cost-only / VNS cost-only
sim_kernel_t SIM_CAT(sim_spec_nu, SIM_NU)(int variant) {
    switch (variant) {
        case 0: return k_sim<SIM_NU, 16, false, false, true>;
        case 1: return k_sim<SIM_NU, 16, true, false, true>;
        case 2: return k_sim<SIM_NU, 16, false, true, true>;
        case 3: return k_sim<SIM_NU, 16, true, false, true, true>;   // small, tail-bound populations: M in shared memory
    }
    return nullptr;
}
