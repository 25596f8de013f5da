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

// A rigid-body tree in the form Robot.rne walks it (reference Robot.py:1704-1903): n joint groups in link order, group j
// hangs off group parent[j] (-1: the base) through the constant transform C[j] (3x4 row-major; the static links of the
// group and the constant part of the joint link folded) followed by ONE joint of kind axis[j] (B2K_RX..B2K_TZ, flip[j])
// reading q[jindex[j]]; I6[j] is the 6x6 spatial inertia of the group in the joint link's frame, [linear; angular] order.
#define B2K_TREE_MAX 16
struct b2k_tree_s {
    int n;
    int parent[B2K_TREE_MAX], axis[B2K_TREE_MAX], flip[B2K_TREE_MAX], jindex[B2K_TREE_MAX];
    double C[B2K_TREE_MAX][12];
    double I6[B2K_TREE_MAX][36];
    void *spec;
};
int b2k_tree_generate(const b2k_tree_s *t, const b2k_gen_opts &o, b2k_gen_out &out);

// outputs of the generated function: RNE / GRAVLOAD / ITORQUE n values; INERTIA / CORIOLIS n*n; ACCEL n*n + n
// (rows of M followed by torque - rne(q, qd, 0), the wrapper solves the system)
int b2k_rne_generate(const b2k_rne_s *r, const b2k_gen_opts &o, b2k_gen_out &out);
