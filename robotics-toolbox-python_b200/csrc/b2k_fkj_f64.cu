// fp64 instantiations of the FK / Jacobian kernels (separate TU so dtypes compile in parallel)
#include "b2k_fkj.cuh"
int b2k_fkj_launch_f64(const b2k_chain_s *c, int mode, const void *q, long long nrows, long long ldq,
                       const double *base, const double *tool, void *T, void *J, cudaStream_t st)
{
    return fkj_launch<double>(c, mode, q, nrows, ldq, base, tool, T, J, st);
}
