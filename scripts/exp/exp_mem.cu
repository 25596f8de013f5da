// exp_mem.cu -- measurement aid (not product code): what HBM write patterns of the shape
// "per-warp tile: read 1.8 KB q, write 4 KB T chunk + 10.75 KB J chunk" achieve on this GPU,
// against a plain linear fill.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a exp_mem.cu -o exp_mem
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_fill(double2 *p, long long n16)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        p[i] = make_double2(1.0, 2.0);
}

// mode bit0: read q; bit1: write T; bit2: write J.  rows_per_tile = 32*RPT (a warp handles RPT sub-tiles back to back)
template <int WARPS>
__global__ void k_tiles(const double *__restrict__ q, double *__restrict__ T, double *__restrict__ J, long long nrows,
                        int mode, int order)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ntiles = nrows >> 5;
    const long long stride = (long long)gridDim.x * WARPS;
    double acc = 0;
    for (long long t0 = (long long)blockIdx.x * WARPS + warp; t0 < ntiles; t0 += stride) {
        long long tile = t0;
        if (order == 1) { // blocked order: each warp owns a contiguous range of tiles
            const long long per = (ntiles + stride - 1) / stride;
            const long long w = (long long)blockIdx.x * WARPS + warp;
            tile = w * per + (t0 - w) / stride;
            if (tile >= ntiles) continue;
        }
        const long long row0 = tile << 5;
        if (mode & 1) {
            const double2 *g = reinterpret_cast<const double2 *>(q + row0 * 7);
            for (int u = lane; u < 112; u += 32) { double2 v = g[u]; acc += v.x + v.y; }
        }
        if (mode & 2) {
            double2 *g = reinterpret_cast<double2 *>(T + row0 * 16);
#pragma unroll
            for (int it = 0; it < 8; it++) g[it * 32 + lane] = make_double2(acc, 1.0);
        }
        if (mode & 4) {
            double2 *g = reinterpret_cast<double2 *>(J + row0 * 42);
#pragma unroll
            for (int it = 0; it < 21; it++) g[it * 32 + lane] = make_double2(acc, 2.0);
        }
    }
    if (acc == 12345.678) T[0] = acc;
}

int main(int argc, char **argv)
{
    const long long N = 1000000;
    double *q[4], *T, *J;
    for (int i = 0; i < 4; i++) { CK(cudaMalloc(&q[i], N * 7 * 8)); CK(cudaMemset(q[i], 0, N * 7 * 8)); }
    CK(cudaMalloc(&T, N * 16 * 8));
    CK(cudaMalloc(&J, N * 42 * 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto fn, const char *name, double bytes) {
        for (int i = 0; i < 5; i++) fn(i);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < 30; i++) fn(i);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 30;
        printf("%-44s %8.2f us  %7.1f GB/s\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    };
    timeit([&](int) { k_fill<<<148 * 8, 256>>>((double2 *)J, N * 42 / 2); }, "fill J (336 MB) linear", 336e6);
    timeit([&](int) { k_fill<<<148 * 8, 256>>>((double2 *)J, N * 42 / 2); k_fill<<<148 * 8, 256>>>((double2 *)T, N * 16 / 2); }, "fill J then T (464 MB)", 464e6);
    char name[128];
    for (int order = 0; order < 2; order++)
        for (int bps = 1; bps <= 8; bps *= 2) {
            for (int mode : {6, 7, 4, 2}) {
                double bytes = ((mode & 1) ? 56e6 : 0) + ((mode & 2) ? 128e6 : 0) + ((mode & 4) ? 336e6 : 0);
                snprintf(name, sizeof(name), "tiles w4 blocks/SM=%d mode=%d order=%d", bps, mode, order);
                timeit([&](int i) { k_tiles<4><<<148 * bps, 128>>>(q[i & 3], T, J, N, mode, order); }, name, bytes);
            }
        }
    for (int bps = 1; bps <= 4; bps *= 2) {
        snprintf(name, sizeof(name), "tiles w8 blocks/SM=%d mode=7 order=0", bps);
        timeit([&](int i) { k_tiles<8><<<148 * bps, 256>>>(q[i & 3], T, J, N, 7, 0); }, name, 520e6);
    }
    return 0;
}
