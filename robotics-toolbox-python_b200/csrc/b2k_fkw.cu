// b2k_fkw.cu -- measurement variant: the literal "one warp = one joint configuration" walk.
// Filled in after the default (lane-per-configuration) path; see DESIGN.md "Variants".
#include "b2k_common.cuh"
int b2k_fkw_launch(const b2k_chain_s *, int, int, const void *, long long, long long, const double *, const double *,
                   void *, void *, cudaStream_t)
{
    b2k_set_error("warp-per-configuration variant not built into this library");
    return B2K_ERR_INVALID;
}
