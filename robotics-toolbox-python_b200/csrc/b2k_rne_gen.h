// b2k_rne_gen.h -- interface of the robot-specialised RNE code generator (b2k_rne_gen.cpp).
#pragma once

#include <string>
#include <vector>

#include "b2k_common.cuh"

enum { B2K_GEN_RNE = 0, B2K_GEN_INERTIA = 1, B2K_GEN_GRAVLOAD = 2, B2K_GEN_ITORQUE = 3, B2K_GEN_CORIOLIS = 4, B2K_GEN_ACCEL = 5 };

struct b2k_gen_opts {
    int mode = B2K_GEN_RNE;
    int grav_mask = 7;  // bit k: base acceleration component k may be non-zero
    int has_fext = 0;   // a tip wrench is given
};

struct b2k_gen_out {
    std::string source;         // one __device__ function `rne_row(C, grav, fext, st, ct, in1, in2, out)` in terms of `real`
    std::vector<double> consts; // the constant bank C the function reads
    int n_mul = 0, n_fma = 0, n_add = 0;
    std::string error;
};

// outputs of the generated function: RNE / GRAVLOAD / ITORQUE n values; INERTIA / CORIOLIS n*n; ACCEL n*n + n
// (rows of M followed by torque - rne(q, qd, 0), the wrapper solves the system)
int b2k_rne_generate(const b2k_rne_s *r, const b2k_gen_opts &o, b2k_gen_out &out);
