// fp64 instantiations of the dynamics fan-out kernels (inertia, gravload, itorque, coriolis, accel)
#include "b2k_rne.cuh"
int b2k_rne_fan_launch_f64(const b2k_rne_s *r, int mode, const void *in0, const void *in1, const void *in2, long long nrows,
                           const double *grav, void *out, cudaStream_t st)
{
    return rne_fan_launch<double>(r, mode, in0, in1, in2, nrows, grav, out, st);
}
