// b2k_selftest.cu -- test hook for the in-house sincos (see include/b2kin.h: b2k_selftest_sincos)
#include "b2k_trig.cuh"

template <typename real>
__global__ void k_selftest_sincos(const __grid_constant__ TrigC<real> tc, const real *__restrict__ x, long long n,
                                  real *__restrict__ s, real *__restrict__ c)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        b2k_sincos(x[i], tc, &s[i], &c[i]);
}

extern "C" int b2k_selftest_sincos(int dtype, const void *x, int64_t n, void *s, void *c, void *stream)
{
    if (n < 0 || (n > 0 && (!x || !s || !c))) { b2k_set_error("b2k_selftest_sincos: bad arguments"); return B2K_ERR_INVALID; }
    if (n == 0) return B2K_OK;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)((n + 255) / 256 > 65535 ? 65535 : (n + 255) / 256);
    if (dtype == B2K_F64) {
        TrigC<double> t;
        b2k_fill_trig<double>(t);
        k_selftest_sincos<double><<<grid, 256, 0, st>>>(t, (const double *)x, n, (double *)s, (double *)c);
    } else if (dtype == B2K_F32) {
        TrigC<float> t;
        b2k_fill_trig<float>(t);
        k_selftest_sincos<float><<<grid, 256, 0, st>>>(t, (const float *)x, n, (float *)s, (float *)c);
    } else {
        b2k_set_error("b2k_selftest_sincos: bad dtype");
        return B2K_ERR_INVALID;
    }
    b2k_count_launch();
    B2K_CUDA(cudaGetLastError());
    return B2K_OK;
}
