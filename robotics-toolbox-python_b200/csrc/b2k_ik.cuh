// b2k_ik.cuh -- batched Levenberg-Marquardt inverse kinematics, one fused persistent kernel (sm_100a).
//
// Replaces fknm.IK_LM_c (reference fknm.cpp:394-525 -> _IK_LM_Chan/_Wampler/_Sugihara
// ik.cpp:157-209 -> _IK_loop ik.cpp:19-75) and, with semantics = B2K_IK_SEM_PYTHON, the Python
// solver behind ETS.ikine_LM (IKSolver._solve IK.py:297-367 + IK_LM.step IK.py:994-1017).
//
// Mapping (DESIGN.md "Kernel K4"): one lane per IK problem, the whole LM loop in registers:
//   FK walk -> angle-axis error (ik.cpp:241-286) -> E = 1/2 e^T We e -> termination / wrap /
//   joint-limit test -> base-frame Jacobian -> g = J^T We e, A = J^T We J + Wn -> Cholesky
//   solve of the n x n SPD system (the reference forms A.inverse(), ik.cpp:171) -> q += dq.
// The kernel is persistent: the grid is sized to the machine and each lane pulls the next
// problem (grid-stride) as soon as its current one terminates, so lanes in a warp are always
// inside the same LM step code whatever iteration / search their problem is at.  Restarts
// draw q inside the joint limits from a counter-based generator keyed by (seed, row, search,
// joint) -- the reference uses unseeded libc rand() (ik.cpp:293) -- mirrored bit for bit by
// oracle/oracle_kin.c:orc_rand_u01 so the restart sequence is testable.
// Compute/latency-bound, not HBM-bound: per problem it reads 16 reals and writes n + 4 words.
#pragma once

#include "b2k_fkj.cuh"

template <typename real, int N>
struct IkP {
    real qlim_l[N], qlim_h[N];
    real we[6];
    real lambda, tol;
    int ilimit, slimit, method, reject_jl, semantics, rng_per_row, has_q0;
    unsigned long long seed;
};

__device__ __forceinline__ unsigned long long b2k_mix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// uniform in [0,1): 53 random bits for fp64, the top 24 of the same word for fp32
template <typename real>
__device__ __forceinline__ real b2k_rand_u01(unsigned long long seed, unsigned long long row, unsigned search,
                                             unsigned joint)
{
    unsigned long long h = b2k_mix64(seed ^ (0x5851F42D4C957F2DULL * (row + 1)));
    h = b2k_mix64(h + (((unsigned long long)search << 32) | (unsigned long long)joint));
    if (sizeof(real) == 8) return (real)((double)(h >> 11) * (1.0 / 9007199254740992.0));
    return (real)((float)(h >> 40) * (1.0f / 16777216.0f));
}

template <typename real> __device__ __forceinline__ real b2k_sqrt(real x);
template <> __device__ __forceinline__ double b2k_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float b2k_sqrt<float>(float x) { return sqrtf(x); }
template <typename real> __device__ __forceinline__ real b2k_atan2(real y, real x);
template <> __device__ __forceinline__ double b2k_atan2<double>(double y, double x) { return atan2(y, x); }
template <> __device__ __forceinline__ float b2k_atan2<float>(float y, float x) { return atan2f(y, x); }
template <typename real> __device__ __forceinline__ real b2k_fmod(real y, real x);
template <> __device__ __forceinline__ double b2k_fmod<double>(double y, double x) { return fmod(y, x); }
template <> __device__ __forceinline__ float b2k_fmod<float>(float y, float x) { return fmodf(y, x); }

// restart sample, _rand_q ik.cpp:288-299: qlim_l + (U(-1,1) + 1) * range/2
template <typename real, int N>
__device__ __forceinline__ void ik_rand_q(const IkP<real, N> &K, unsigned long long row, unsigned search, real *q)
{
#pragma unroll
    for (int i = 0; i < N; i++) {
        real r = (real)2 * b2k_rand_u01<real>(K.seed, row, search, (unsigned)i) - (real)1;
        real range2 = (K.qlim_h[i] - K.qlim_l[i]) / (real)2;
        q[i] = (r + (real)1) * range2 + K.qlim_l[i];
    }
}

// angle-axis pose error, _angle_axis ik.cpp:241-286 (thresholds kept verbatim)
template <typename real>
__device__ __forceinline__ void ik_angle_axis(const Pose<real> &Te, const real *Tp /*12: row-major 3x4*/, real *e)
{
    e[0] = Tp[3] - Te.p[0];
    e[1] = Tp[7] - Te.p[1];
    e[2] = Tp[11] - Te.p[2];
    // R = Rep * Re^T ; R(i,j) = sum_k Rep(i,k) Re(j,k) ; Re(j,k) = column k, component j
    real R[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            R[i][j] = Tp[i * 4 + 0] * Te.c0[j] + Tp[i * 4 + 1] * Te.c1[j] + Tp[i * 4 + 2] * Te.c2[j];
    real lx = R[2][1] - R[1][2], ly = R[0][2] - R[2][0], lz = R[1][0] - R[0][1];
    real ln = b2k_sqrt<real>(lx * lx + ly * ly + lz * lz);
    real tr = R[0][0] + R[1][1] + R[2][2];
    if (ln < (real)1e-6) {
        if (tr > 0) {
            e[3] = e[4] = e[5] = 0;
        } else {
            const real pi2 = (real)1.57079632679489661923132169163975144;
            e[3] = pi2 * (R[0][0] + 1);
            e[4] = pi2 * (R[1][1] + 1);
            e[5] = pi2 * (R[2][2] + 1);
        }
    } else {
        real ang = b2k_atan2<real>(ln, tr - 1);
        e[3] = ang * lx / ln;
        e[4] = ang * ly / ln;
        e[5] = ang * lz / ln;
    }
}

// In-place Cholesky solve of the packed-lower SPD system A x = b (A: N(N+1)/2 entries, row-wise
// lower triangle).  Returns false on a non-positive / non-finite pivot.
template <typename real, int N>
__device__ __forceinline__ bool ik_chol_solve(real *A, real *b)
{
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; j++) {
        real d = A[j * (j + 1) / 2 + j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= A[j * (j + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
        ok = ok && (d > 0) && isfinite(d);
        real inv = (real)1 / b2k_sqrt<real>(d);
        A[j * (j + 1) / 2 + j] = inv; // store 1/L_jj
#pragma unroll
        for (int i = j + 1; i < N; i++) {
            real s = A[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= A[i * (i + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
            A[i * (i + 1) / 2 + j] = s * inv;
        }
    }
    // forward L y = b
#pragma unroll
    for (int i = 0; i < N; i++) {
        real s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= A[i * (i + 1) / 2 + k] * b[k];
        b[i] = s * A[i * (i + 1) / 2 + i];
    }
    // backward L^T x = y
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        real s = b[i];
#pragma unroll
        for (int k = i + 1; k < N; k++) s -= A[k * (k + 1) / 2 + i] * b[k];
        b[i] = s * A[i * (i + 1) / 2 + i];
    }
    return ok;
}

template <typename real, int N, int PROF>
__global__ void __launch_bounds__(B2K_THREADS)
k_ik_lm(const __grid_constant__ ChainP<real, N> P, const __grid_constant__ IkP<real, N> K,
        const real *__restrict__ Tep, const real *__restrict__ q0, long long nprob, real *__restrict__ q_out,
        int *__restrict__ success, int *__restrict__ iterations, int *__restrict__ searches,
        real *__restrict__ residual)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const real PI = (real)3.14159265358979323846264338327950288;
    const real PI_X2 = (real)6.283185307179586;

    real q[N], Tp[12], E = 0;
    int it = 0, search = 0, iter = 0; // meaning depends on semantics, see below
    bool active = idx < nprob;

    auto begin_problem = [&]() {
        const real *t = Tep + idx * 16;
#pragma unroll
        for (int k = 0; k < 12; k++) Tp[k] = t[k];
        const unsigned long long row = K.rng_per_row ? (unsigned long long)idx : 0ULL;
        if (K.has_q0) {
#pragma unroll
            for (int i = 0; i < N; i++) q[i] = q0[idx * N + i];
        } else {
            ik_rand_q<real, N>(K, row, 0u, q);
        }
        E = 0;
        if (K.semantics == B2K_IK_SEM_CPP) { it = 0; search = 1; iter = 1; } // fknm.cpp:406, ik.cpp:39
        else { it = 0; search = 0; iter = 0; }                                  // IK.py:299-312
    };
    auto finish = [&](int ok, int its, int srch) {
#pragma unroll
        for (int i = 0; i < N; i++) q_out[idx * N + i] = q[i];
        success[idx] = ok;
        iterations[idx] = its;
        searches[idx] = srch;
        residual[idx] = E;
        idx += stride;
        active = idx < nprob;
        if (active) begin_problem();
    };
    if (active) begin_problem();

    while (active) {
        // ---- evaluate pose error at q
        Pose<real> Te;
        real zj[N][3], pj[N][3], e[6];
        // jindex is dense 0..n-1 here (checked on the host, like the reference's C++ loop assumes)
        chain_forward<real, N, true, PROF>(P, [&](int j, int) { return q[j]; }, Te, zj, pj);
        ik_angle_axis<real>(Te, Tp, e);
        real Ecur = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) Ecur += e[k] * K.we[k] * e[k];
        Ecur *= (real)0.5;
        E = Ecur;
        const unsigned long long row = K.rng_per_row ? (unsigned long long)idx : 0ULL;

        if (K.semantics == B2K_IK_SEM_CPP && Ecur < K.tol) {
            // ik.cpp:48-54: wrap with fmod (sign of the dividend), then the limit test
            bool inlim = true;
#pragma unroll
            for (int i = 0; i < N; i++) {
                q[i] = b2k_fmod<real>(q[i] + PI, PI_X2) - PI;
                inlim = inlim && !(q[i] < K.qlim_l[i] || q[i] > K.qlim_h[i]);
            }
            if (!K.reject_jl || inlim) { finish(1, it + iter, search); continue; }
            // converged outside the limits: this search is abandoned (ik.cpp:61-69)
            it += iter; iter = 0; search++;
            if (search > K.slimit) { ik_rand_q<real, N>(K, row, (unsigned)(search - 1), q); finish(0, it, search); continue; }
            ik_rand_q<real, N>(K, row, (unsigned)(search - 1), q);
            continue;
        }

        // ---- LM step: g = J^T We e ; A = J^T We J + Wn (packed lower) ; dq = A^-1 g
        real J[N][6];
#pragma unroll
        for (int j = 0; j < N; j++) {
            const bool rev = PROF == 1 ? true : (P.axis[j] < 3);
            if (rev) {
                real dx = Te.p[0] - pj[j][0], dy = Te.p[1] - pj[j][1], dz = Te.p[2] - pj[j][2];
                J[j][0] = fma(zj[j][1], dz, -(zj[j][2] * dy));
                J[j][1] = fma(zj[j][2], dx, -(zj[j][0] * dz));
                J[j][2] = fma(zj[j][0], dy, -(zj[j][1] * dx));
                J[j][3] = zj[j][0]; J[j][4] = zj[j][1]; J[j][5] = zj[j][2];
            } else {
                J[j][0] = zj[j][0]; J[j][1] = zj[j][1]; J[j][2] = zj[j][2];
                J[j][3] = 0; J[j][4] = 0; J[j][5] = 0;
            }
        }
        const real wn = (K.method == B2K_LM_CHAN) ? K.lambda * Ecur
                        : (K.method == B2K_LM_WAMPLER) ? K.lambda : (Ecur + K.lambda);
        real A[N * (N + 1) / 2], g[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            real s = 0;
#pragma unroll
            for (int k = 0; k < 6; k++) s += J[i][k] * K.we[k] * e[k];
            g[i] = s;
#pragma unroll
            for (int j = 0; j <= i; j++) {
                real a = 0;
#pragma unroll
                for (int k = 0; k < 6; k++) a += J[i][k] * K.we[k] * J[j][k];
                A[i * (i + 1) / 2 + j] = a + (i == j ? wn : (real)0);
            }
        }
        const bool ok = ik_chol_solve<real, N>(A, g);

        if (K.semantics == B2K_IK_SEM_CPP) {
            bool restart = false;
            if (ok) {
#pragma unroll
                for (int i = 0; i < N; i++) q[i] += g[i];
                iter++;
                restart = iter > K.ilimit;
            } else {
                iter++;
                restart = true; // unfactorisable normal matrix: abandon this search
            }
            if (restart) {
                it += iter; iter = 0; search++;
                ik_rand_q<real, N>(K, row, (unsigned)(search - 1), q);
                if (search > K.slimit) { finish(0, it, search); continue; }
            }
        } else {
            // Python semantics: count the step, apply it, then test the PRE-step E (IK.py:314-327)
            iter++;
            bool next_search = false;
            if (!ok) {
                next_search = true; // numpy LinAlgError -> abandon search (IK.py:321-324)
            } else {
#pragma unroll
                for (int i = 0; i < N; i++) q[i] += g[i];
                if (Ecur < K.tol) {
                    bool inlim = true;
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        real w = b2k_fmod<real>(q[i] + PI, (real)2 * PI);
                        if (w < 0) w += (real)2 * PI; // Python floor-modulo (IK.py:331)
                        q[i] = w - PI;
                        inlim = inlim && !(q[i] < K.qlim_l[i] || q[i] > K.qlim_h[i]);
                    }
                    if (inlim || !K.reject_jl) { finish(1, it + iter, search + 1); continue; }
                    next_search = true;
                } else if (iter >= K.ilimit) {
                    next_search = true;
                }
            }
            if (next_search) {
                it += iter; iter = 0; search++;
                if (search >= K.slimit) { finish(0, it, K.slimit); continue; }
                ik_rand_q<real, N>(K, row, (unsigned)search, q);
            }
        }
    }
}

template <typename real, int N>
int ik_launch_n(const b2k_chain_s *c, const real *Tep, long long nprob, const real *q0, int ilimit, int slimit,
                double tol, int reject_jl, const double *we, double lambda, int method, unsigned long long seed,
                int semantics, int rng_per_row, real *q_out, int *success, int *iterations, int *searches,
                real *residual, cudaStream_t st)
{
    ChainP<real, N> P;
    b2k_fill_chain<real, N>(c, nullptr, nullptr, true, P);
    IkP<real, N> K;
    for (int i = 0; i < N; i++) { K.qlim_l[i] = (real)c->qlim_l[i]; K.qlim_h[i] = (real)c->qlim_h[i]; }
    for (int k = 0; k < 6; k++) K.we[k] = we ? (real)we[k] : (real)1;
    K.lambda = (real)lambda; K.tol = (real)tol;
    K.ilimit = ilimit; K.slimit = slimit; K.method = method; K.reject_jl = reject_jl ? 1 : 0;
    K.semantics = semantics; K.rng_per_row = rng_per_row ? 1 : 0; K.has_q0 = q0 ? 1 : 0; K.seed = seed;
    auto launch = [&](auto kern) -> int {
        int per_sm = b2k_blocks_per_sm((const void *)kern, B2K_THREADS, 0);
        if (per_sm < 1) return per_sm < 0 ? per_sm : (b2k_set_error("ik kernel does not fit on an SM"), B2K_ERR_INVALID);
        long long grid = (long long)b2k_num_sms() * per_sm;
        long long need = (nprob + B2K_THREADS - 1) / B2K_THREADS;
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, B2K_THREADS, 0, st>>>(P, K, Tep, q0, nprob, q_out, success, iterations, searches, residual);
        b2k_count_launch();
        B2K_CUDA(cudaGetLastError());
        return B2K_OK;
    };
    if (c->dh_like) return launch(k_ik_lm<real, N, 1>);
    return launch(k_ik_lm<real, N, 0>);
}

template <typename real>
int ik_launch(const b2k_chain_s *c, const void *Tep, long long nprob, const void *q0, int ilimit, int slimit, double tol,
              int reject_jl, const double *we, double lambda, int method, unsigned long long seed, int semantics,
              int rng_per_row, void *q_out, int *success, int *iterations, int *searches, void *residual,
              cudaStream_t st)
{
#define B2K_CASE(NN)                                                                                                   \
    case NN:                                                                                                           \
        return ik_launch_n<real, NN>(c, (const real *)Tep, nprob, (const real *)q0, ilimit, slimit, tol, reject_jl, we, \
                                     lambda, method, seed, semantics, rng_per_row, (real *)q_out, success, iterations, \
                                     searches, (real *)residual, st);
    switch (c->n) {
        B2K_CASE(1) B2K_CASE(2) B2K_CASE(3) B2K_CASE(4) B2K_CASE(5)
        B2K_CASE(6) B2K_CASE(7) B2K_CASE(8) B2K_CASE(9) B2K_CASE(10)
    default:
        b2k_set_error("ik_lm: unsupported joint count %d", c->n);
        return B2K_ERR_INVALID;
    }
#undef B2K_CASE
}
