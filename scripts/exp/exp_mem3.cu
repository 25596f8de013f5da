// exp_mem3.cu -- measurement aid: persistent grid-stride vs one-shot grids for fill and for the
// per-warp tile pattern (read 1.8 KB q, write 4 KB + 10.75 KB).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_fill_persist(double2 *p, long long n16)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
        p[i] = make_double2(1.0, 2.0);
}
__global__ void k_fill_oneshot(double2 *p, long long n16)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n16) p[i] = make_double2(1.0, 2.0);
}
__global__ void k_fill_oneshot2(double2 *p, long long n16) // 32 B per thread
{
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n16) { p[i] = make_double2(1.0, 2.0); p[i + 1] = make_double2(3.0, 4.0); }
}
__global__ void k_fill_oneshot4(double2 *p, long long n16) // 4 x 16 B per thread, warp-coalesced per instruction
{
    long long base = (long long)blockIdx.x * blockDim.x * 4 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; k++) { long long i = base + (long long)k * blockDim.x; if (i < n16) p[i] = make_double2(1.0, 2.0); }
}

template <int MODE>
__global__ void k_tiles(const double *__restrict__ q, double *__restrict__ T, double *__restrict__ J, long long nrows)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ntiles = nrows >> 5;
    const long long stride = (long long)gridDim.x * 4;
    double acc = 0;
    for (long long tile = (long long)blockIdx.x * 4 + warp; tile < ntiles; tile += stride) {
        const long long row0 = tile << 5;
        if (MODE & 1) {
            const double2 *g = reinterpret_cast<const double2 *>(q + row0 * 7);
            for (int u = lane; u < 112; u += 32) { double2 v = g[u]; acc += v.x + v.y; }
        }
        double2 *gt = reinterpret_cast<double2 *>(T + row0 * 16);
#pragma unroll
        for (int it = 0; it < 8; it++) gt[it * 32 + lane] = make_double2(acc, 1.0);
        double2 *gj = reinterpret_cast<double2 *>(J + row0 * 42);
#pragma unroll
        for (int it = 0; it < 21; it++) gj[it * 32 + lane] = make_double2(acc, 2.0);
    }
    if (acc == 12345.678) T[0] = acc;
}

int main()
{
    const long long N = 1000000;
    double *q[4], *T, *J, *B;
    for (int i = 0; i < 4; i++) { cudaMalloc(&q[i], N * 7 * 8); cudaMemset(q[i], 0, N * 7 * 8); }
    cudaMalloc(&T, N * 16 * 8); cudaMalloc(&J, N * 42 * 8); cudaMalloc(&B, 520000000);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto fn, const char *name, double bytes) {
        for (int i = 0; i < 5; i++) fn(i);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < 30; i++) fn(i);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 30;
        printf("%-44s %8.2f us  %7.1f GB/s\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    };
    const long long n16 = 520000000 / 16;
    timeit([&](int) { k_fill_persist<<<148 * 8, 256>>>((double2 *)B, n16); }, "fill 520MB persistent 148x8x256", 520e6);
    timeit([&](int) { k_fill_persist<<<148 * 16, 128>>>((double2 *)B, n16); }, "fill 520MB persistent 148x16x128", 520e6);
    timeit([&](int) { k_fill_oneshot<<<(unsigned)((n16 + 255) / 256), 256>>>((double2 *)B, n16); }, "fill 520MB one-shot 16B/thr", 520e6);
    timeit([&](int) { k_fill_oneshot2<<<(unsigned)((n16 / 2 + 255) / 256), 256>>>((double2 *)B, n16); }, "fill 520MB one-shot 32B/thr", 520e6);
    timeit([&](int) { k_fill_oneshot4<<<(unsigned)((n16 + 1023) / 1024), 256>>>((double2 *)B, n16); }, "fill 520MB one-shot 4x16B/thr", 520e6);
    timeit([&](int) { cudaMemsetAsync(B, 0, 520000000); }, "cudaMemsetAsync 520MB", 520e6);
    const unsigned one = (unsigned)((N / 32 + 3) / 4);
    timeit([&](int i) { k_tiles<6><<<148 * 4, 128>>>(q[i & 3], T, J, N); }, "tiles W persistent 592 blk", 464e6);
    timeit([&](int i) { k_tiles<6><<<one, 128>>>(q[i & 3], T, J, N); }, "tiles W one-shot 7813 blk", 464e6);
    timeit([&](int i) { k_tiles<7><<<148 * 4, 128>>>(q[i & 3], T, J, N); }, "tiles R+W persistent 592 blk", 520e6);
    timeit([&](int i) { k_tiles<7><<<one, 128>>>(q[i & 3], T, J, N); }, "tiles R+W one-shot 7813 blk", 520e6);
    timeit([&](int i) { k_tiles<7><<<148 * 8, 128>>>(q[i & 3], T, J, N); }, "tiles R+W persistent 1184 blk", 520e6);
    timeit([&](int i) { k_tiles<7><<<148 * 16, 128>>>(q[i & 3], T, J, N); }, "tiles R+W persistent 2368 blk", 520e6);
    return 0;
}
