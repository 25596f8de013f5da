// exp_mem2.cu -- measurement aid: how to bring a small read stream (q, 56 MB) into a kernel whose
// warps each write 14.8 KB per tile, without the reads slowing the write stream.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_pipeline.h>

// rmode: 0 none, 1 sync LDG, 2 register prefetch one tile ahead, 3 sync LDG with evict_first,
//        4 cp.async into smem one tile ahead, 5 dedicated loader warp (warp 0 loads for all, others write)
// wmode: 0 plain st, 1 st.global.cs (streaming), 2 st L2::evict_first
template <int RMODE, int WMODE>
__global__ void __launch_bounds__(160) k(const double *__restrict__ q, double *__restrict__ T, double *__restrict__ J, long long nrows)
{
    extern __shared__ double2 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = (RMODE == 5) ? 5 : 4; // warps per block (one extra loader warp in mode 5)
    const int cw = (RMODE == 5) ? warp - 1 : warp; // compute/write warp index 0..3
    const long long ntiles = nrows >> 5;
    const long long stride = (long long)gridDim.x * 4;
    double acc = 0;
    auto st16 = [&](double2 *p, double2 v) {
        if (WMODE == 0) *p = v;
        else if (WMODE == 1) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
        else asm volatile("st.global.wt.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    };
    if (RMODE == 5 && warp == 0) { // loader warp: stream q for the 4 writer warps through smem flags-free (just touch)
        for (long long t0 = (long long)blockIdx.x * 4; t0 < ntiles; t0 += stride)
            for (int w = 0; w < 4 && t0 + w < ntiles; w++) {
                const double2 *g = reinterpret_cast<const double2 *>(q + ((t0 + w) << 5) * 7);
                for (int u = lane; u < 112; u += 32) { double2 v = g[u]; acc += v.x + v.y; }
            }
        if (acc == 1234.5) T[0] = acc;
        return;
    }
    double2 pre[4] = {};
    long long t0 = (long long)blockIdx.x * 4 + cw;
    if (RMODE == 6 || RMODE == 7) {
        for (long long t = t0; t < ntiles; t += stride) {
            const char *a = reinterpret_cast<const char *>(q + (t << 5) * 7) + lane * 128;
            if (lane < 14) {
                if (RMODE == 6) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a));
                else asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
        }
    }
    if (RMODE == 2 && t0 < ntiles) {
        const double2 *g = reinterpret_cast<const double2 *>(q + (t0 << 5) * 7);
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) if (lane + 32 * k2 < 112) pre[k2] = g[lane + 32 * k2];
    }
    if (RMODE == 4 && t0 < ntiles) {
        const double2 *g = reinterpret_cast<const double2 *>(q + (t0 << 5) * 7);
        for (int u = lane; u < 112; u += 32) __pipeline_memcpy_async(&sm[cw * 112 + u], &g[u], 16);
        __pipeline_commit();
    }
    for (; t0 < ntiles; t0 += stride) {
        const long long row0 = t0 << 5;
        if (RMODE == 1 || RMODE == 3 || RMODE == 6 || RMODE == 7) {
            const double2 *g = reinterpret_cast<const double2 *>(q + row0 * 7);
            for (int u = lane; u < 112; u += 32) {
                double2 v;
                if (RMODE != 3) v = g[u];
                else asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(g + u));
                acc += v.x + v.y;
            }
        } else if (RMODE == 2) {
#pragma unroll
            for (int k2 = 0; k2 < 4; k2++) acc += pre[k2].x + pre[k2].y;
            if (t0 + stride < ntiles) {
                const double2 *g = reinterpret_cast<const double2 *>(q + ((t0 + stride) << 5) * 7);
#pragma unroll
                for (int k2 = 0; k2 < 4; k2++) if (lane + 32 * k2 < 112) pre[k2] = g[lane + 32 * k2];
            }
        } else if (RMODE == 4) {
            __pipeline_wait_prior(0);
            __syncwarp();
            for (int u = lane; u < 112; u += 32) { double2 v = sm[cw * 112 + u]; acc += v.x + v.y; }
            __syncwarp();
            if (t0 + stride < ntiles) {
                const double2 *g = reinterpret_cast<const double2 *>(q + ((t0 + stride) << 5) * 7);
                for (int u = lane; u < 112; u += 32) __pipeline_memcpy_async(&sm[cw * 112 + u], &g[u], 16);
                __pipeline_commit();
            }
        }
        double2 *gt = reinterpret_cast<double2 *>(T + row0 * 16);
#pragma unroll
        for (int it = 0; it < 8; it++) st16(gt + it * 32 + lane, make_double2(acc, 1.0));
        double2 *gj = reinterpret_cast<double2 *>(J + row0 * 42);
#pragma unroll
        for (int it = 0; it < 21; it++) st16(gj + it * 32 + lane, make_double2(acc, 2.0));
    }
    if (acc == 12345.678) T[0] = acc;
}

int main()
{
    const long long N = 1000000;
    double *q[4], *T, *J;
    for (int i = 0; i < 4; i++) { cudaMalloc(&q[i], N * 7 * 8); cudaMemset(q[i], 0, N * 7 * 8); }
    cudaMalloc(&T, N * 16 * 8); cudaMalloc(&J, N * 42 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto fn, const char *name, double bytes) {
        for (int i = 0; i < 5; i++) fn(i);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int i = 0; i < 30; i++) fn(i);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 30;
        cudaError_t e = cudaGetLastError();
        printf("%-40s %8.2f us  %7.1f GB/s %s\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    const int G = 148 * 4;
#define RUN(R, Wm, TH) timeit([&](int i) { k<R, Wm><<<G, TH, 4 * 112 * 16>>>(q[i & 3], T, J, N); }, "rmode " #R " wmode " #Wm, (R ? 520e6 : 464e6));
    RUN(0, 0, 128) RUN(1, 0, 128) RUN(6, 0, 128) RUN(7, 0, 128) RUN(6, 1, 128) RUN(7, 1, 128) RUN(1, 0, 128) RUN(6, 0, 128) RUN(7, 0, 128)
    return 0;
}
