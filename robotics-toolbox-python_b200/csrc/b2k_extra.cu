// b2k_extra.cu -- pure functions of the Jacobian: manipulator Hessian and Yoshikawa
// manipulability.  Replaces fknm.ETS_hessian0 / ETS_hessiane (reference fknm.cpp:583-783 ->
// _ETS_hessian methods.cpp:16-32) and the yoshikawa branch of ETS.manipulability (ETS.py:1780-1787).
#include "b2k_common.cuh"

// H[a, 0:3, b] = Jw_a x Jv_b, H[a, 3:6, b] = Jw_a x Jw_b for b >= a; mirrored translational block and a zero
// rotational block for b < a (methods.cpp:18-31).  One warp per row: the 6n values of J are staged in shared
// memory, the 6 n^2 outputs of the row are produced by consecutive lanes -> fully coalesced stores.
template <typename real, int N>
__global__ void __launch_bounds__(256) k_hessian(const real *__restrict__ J, long long nrows, real *__restrict__ H)
{
    __shared__ real sJ[8][6 * N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wstride = (long long)gridDim.x * 8;
    for (long long row = (long long)blockIdx.x * 8 + warp; row < nrows; row += wstride) {
        for (int e = lane; e < 6 * N; e += 32) sJ[warp][e] = J[row * (6 * N) + e];
        __syncwarp();
        const real *j = sJ[warp];
        real *out = H + row * (6 * N * N);
        for (int e = lane; e < 6 * N * N; e += 32) {
            const int a = e / (6 * N), r = (e / N) % 6, b = e % N;
            real v = 0;
            if (b >= a || r < 3) {
                // u x w, component c: u = Jw of the lower-numbered joint, w = Jv (r < 3) or Jw (r >= 3) of the other
                const int lo = b >= a ? a : b, hi = b >= a ? b : a;
                const int c = r % 3, c1 = (c + 1) % 3, c2 = (c + 2) % 3;
                const int wrow = r < 3 ? 0 : 3;
                const real u1 = j[(3 + c1) * N + lo], u2 = j[(3 + c2) * N + lo];
                const real w1 = j[(wrow + c1) * N + hi], w2 = j[(wrow + c2) * N + hi];
                v = u1 * w2 - u2 * w1;
            }
            out[e] = v;
        }
        __syncwarp();
    }
}

// m = sqrt(|det(Ja Ja^T)|) with Ja = the selected rows of J (|det Ja| when Ja is square), ETS.py:1780-1787
template <typename real, int N>
__global__ void __launch_bounds__(128) k_yoshikawa(const real *__restrict__ J, long long nrows, unsigned axes_mask,
                                                   real *__restrict__ m)
{
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    int sel[6], na = 0;
    for (int k = 0; k < 6; k++)
        if (axes_mask & (1u << k)) sel[na++] = k;
    real A[36];
    const real *j = J + row * (6 * N);
    const bool square = (na == N);
    for (int a = 0; a < na; a++)
        for (int b = 0; b < na; b++) {
            real s = 0;
            if (square) s = j[sel[a] * N + b];
            else
                for (int k = 0; k < N; k++) s += j[sel[a] * N + k] * j[sel[b] * N + k];
            A[a * na + b] = s;
        }
    real det = 1;
    for (int c = 0; c < na; c++) { // Gaussian elimination with partial pivoting
        int p = c;
        real best = fabs(A[c * na + c]);
        for (int r = c + 1; r < na; r++)
            if (fabs(A[r * na + c]) > best) { best = fabs(A[r * na + c]); p = r; }
        if (best == 0) { det = 0; break; }
        if (p != c) {
            for (int k = 0; k < na; k++) { real t = A[c * na + k]; A[c * na + k] = A[p * na + k]; A[p * na + k] = t; }
            det = -det;
        }
        det *= A[c * na + c];
        for (int r = c + 1; r < na; r++) {
            const real f = A[r * na + c] / A[c * na + c];
            for (int k = c + 1; k < na; k++) A[r * na + k] -= f * A[c * na + k];
        }
    }
    m[row] = square ? fabs(det) : sqrt(fabs(det));
}

template <typename real>
static int extra_launch(int what, int n, const void *J, long long N, unsigned axes_mask, void *out, cudaStream_t st)
{
#define B2K_CASE(NN)                                                                                                  \
    case NN:                                                                                                          \
        if (what == 0) {                                                                                              \
            long long blocks = (N + 7) / 8;                                                                           \
            const long long cap = (long long)b2k_num_sms() * 16;                                                      \
            if (blocks > cap) blocks = cap;                                                                           \
            k_hessian<real, NN><<<(unsigned)blocks, 256, 0, st>>>((const real *)J, N, (real *)out);                   \
        } else {                                                                                                      \
            k_yoshikawa<real, NN><<<(unsigned)((N + 127) / 128), 128, 0, st>>>((const real *)J, N, axes_mask, (real *)out); \
        }                                                                                                             \
        break;
    switch (n) {
        B2K_CASE(1) B2K_CASE(2) B2K_CASE(3) B2K_CASE(4) B2K_CASE(5)
        B2K_CASE(6) B2K_CASE(7) B2K_CASE(8) B2K_CASE(9) B2K_CASE(10)
    default:
        b2k_set_error("unsupported joint count %d", n);
        return B2K_ERR_INVALID;
    }
#undef B2K_CASE
    b2k_count_launch();
    B2K_CUDA(cudaGetLastError());
    return B2K_OK;
}

extern "C" int b2k_hessian(int dtype, int n, const void *J, int64_t N, void *H, void *stream)
{
    if (N < 0 || (N > 0 && (!J || !H))) { b2k_set_error("b2k_hessian: bad arguments"); return B2K_ERR_INVALID; }
    if (dtype != B2K_F32 && dtype != B2K_F64) { b2k_set_error("b2k_hessian: bad dtype"); return B2K_ERR_INVALID; }
    if (N == 0) return B2K_OK;
    return dtype == B2K_F64 ? extra_launch<double>(0, n, J, N, 0, H, (cudaStream_t)stream)
                            : extra_launch<float>(0, n, J, N, 0, H, (cudaStream_t)stream);
}

extern "C" int b2k_manipulability(int dtype, int n, const void *J, int64_t N, uint32_t axes_mask, void *m, void *stream)
{
    if (N < 0 || (N > 0 && (!J || !m))) { b2k_set_error("b2k_manipulability: bad arguments"); return B2K_ERR_INVALID; }
    if (dtype != B2K_F32 && dtype != B2K_F64) { b2k_set_error("b2k_manipulability: bad dtype"); return B2K_ERR_INVALID; }
    if ((axes_mask & 63u) == 0) { b2k_set_error("b2k_manipulability: no Cartesian axis selected"); return B2K_ERR_INVALID; }
    if (N == 0) return B2K_OK;
    return dtype == B2K_F64 ? extra_launch<double>(1, n, J, N, axes_mask & 63u, m, (cudaStream_t)stream)
                            : extra_launch<float>(1, n, J, N, axes_mask & 63u, m, (cudaStream_t)stream);
}
