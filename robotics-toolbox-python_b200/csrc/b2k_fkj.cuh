// b2k_fkj.cuh -- forward kinematics + geometric Jacobian kernels for sm_100a.
//
// Mapping (DESIGN.md "Kernel K12"): a warp owns a TILE of 32 joint configurations.  Lane l
// walks the serial chain of row (tile*32 + l) entirely in registers -- pose as four 3-vectors
// (bottom row 0 0 0 1 implied), chain constants read straight from the constant bank (the
// chain is a __grid_constant__ kernel parameter, so every folded constant is an immediate
// c[0x0][..] operand of the FMA).  The warp as a whole moves the tile's q block and its T / J
// blocks between HBM and registers through a per-warp shared-memory stage so that every
// global access is a fully coalesced 256-byte warp transaction:
//     q tile  (32 x ldq)   : coalesced loads -> smem (odd row stride) -> one row per lane
//     T tile  (32 x 16)    : registers -> smem (odd granule stride) -> coalesced 8-byte stores
//     J tile  (32 x 6n)    : same
// No tensor cores: the products are 3x3 / 6xn (far below an MMA tile) -- this is HBM-bound
// streaming work (520 B per evaluation for Panda fp64).
//
// Arithmetic follows reference methods.cpp:318-352 (pose) and the column rules of
// methods.cpp:137-196 (Jacobian), re-associated: constants folded per joint, pose walked
// left-to-right once, base-frame Jacobian columns formed as z_j x (p_e - p_j) instead of the
// reference's end-effector-frame walk + blkdiag(R,R) rotation.  Same values to rounding.
#pragma once

#include "b2k_common.cuh"

template <typename real>
__device__ __forceinline__ void b2k_sincos(real x, real *s, real *c);
template <>
__device__ __forceinline__ void b2k_sincos<double>(double x, double *s, double *c) { sincos(x, s, c); }
template <>
__device__ __forceinline__ void b2k_sincos<float>(float x, float *s, float *c) { sincosf(x, s, c); }

// 8-byte staging granule: one double or two floats
template <typename real> struct Granule;
template <> struct Granule<double> { typedef double type; static constexpr int PER = 1; };
template <> struct Granule<float> { typedef float2 type; static constexpr int PER = 2; };

// pose T = [c0 c1 c2 p] (columns), all in registers
template <typename real>
struct Pose {
    real c0[3], c1[3], c2[3], p[3];
};

template <typename real>
__device__ __forceinline__ void pose_from_const(Pose<real> &T, const real *A)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        T.c0[i] = A[i * 4 + 0];
        T.c1[i] = A[i * 4 + 1];
        T.c2[i] = A[i * 4 + 2];
        T.p[i] = A[i * 4 + 3];
    }
}

// T <- T * A for a folded constant A of structure class `kind` (uniform across the grid)
template <typename real>
__device__ __forceinline__ void pose_mul_const_right(Pose<real> &T, const real *A, int kind)
{
    if (kind & AK_TX) {
#pragma unroll
        for (int i = 0; i < 3; i++) T.p[i] = fma(T.c0[i], A[3], T.p[i]);
    }
    if (kind & AK_TY) {
#pragma unroll
        for (int i = 0; i < 3; i++) T.p[i] = fma(T.c1[i], A[7], T.p[i]);
    }
    if (kind & AK_TZ) {
#pragma unroll
        for (int i = 0; i < 3; i++) T.p[i] = fma(T.c2[i], A[11], T.p[i]);
    }
    switch (kind & AK_ROTMASK) {
    case AK_IDENT: break;
    case AK_RX:
#pragma unroll
        for (int i = 0; i < 3; i++) {
            real a = T.c1[i], b = T.c2[i];
            T.c1[i] = fma(a, A[5], b * A[9]);
            T.c2[i] = fma(a, A[6], b * A[10]);
        }
        break;
    case AK_RY:
#pragma unroll
        for (int i = 0; i < 3; i++) {
            real a = T.c0[i], b = T.c2[i];
            T.c0[i] = fma(a, A[0], b * A[8]);
            T.c2[i] = fma(a, A[2], b * A[10]);
        }
        break;
    case AK_RZ:
#pragma unroll
        for (int i = 0; i < 3; i++) {
            real a = T.c0[i], b = T.c1[i];
            T.c0[i] = fma(a, A[0], b * A[4]);
            T.c1[i] = fma(a, A[1], b * A[5]);
        }
        break;
    default:
#pragma unroll
        for (int i = 0; i < 3; i++) {
            real a = T.c0[i], b = T.c1[i], c = T.c2[i];
            T.c0[i] = fma(a, A[0], fma(b, A[4], c * A[8]));
            T.c1[i] = fma(a, A[1], fma(b, A[5], c * A[9]));
            T.c2[i] = fma(a, A[2], fma(b, A[6], c * A[10]));
        }
        break;
    }
}

// T <- A * T (used by the backward / end-effector-frame walk and for the base)
template <typename real>
__device__ __forceinline__ void pose_mul_const_left(Pose<real> &T, const real *A, int kind)
{
    const int rk = kind & AK_ROTMASK;
    if (rk == AK_IDENT) {
    } else if (rk == AK_RX) { // rows 1,2 mix
        real *cols[4] = {T.c0, T.c1, T.c2, T.p};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            real a = cols[j][1], b = cols[j][2];
            cols[j][1] = fma(A[5], a, A[6] * b);
            cols[j][2] = fma(A[9], a, A[10] * b);
        }
    } else if (rk == AK_RY) { // rows 0,2 mix
        real *cols[4] = {T.c0, T.c1, T.c2, T.p};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            real a = cols[j][0], b = cols[j][2];
            cols[j][0] = fma(A[0], a, A[2] * b);
            cols[j][2] = fma(A[8], a, A[10] * b);
        }
    } else if (rk == AK_RZ) { // rows 0,1 mix
        real *cols[4] = {T.c0, T.c1, T.c2, T.p};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            real a = cols[j][0], b = cols[j][1];
            cols[j][0] = fma(A[0], a, A[1] * b);
            cols[j][1] = fma(A[4], a, A[5] * b);
        }
    } else {
        real *cols[4] = {T.c0, T.c1, T.c2, T.p};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            real a = cols[j][0], b = cols[j][1], c = cols[j][2];
            cols[j][0] = fma(A[0], a, fma(A[1], b, A[2] * c));
            cols[j][1] = fma(A[4], a, fma(A[5], b, A[6] * c));
            cols[j][2] = fma(A[8], a, fma(A[9], b, A[10] * c));
        }
    }
    if (kind & AK_TX) T.p[0] += A[3];
    if (kind & AK_TY) T.p[1] += A[7];
    if (kind & AK_TZ) T.p[2] += A[11];
}

// rotate the column pair (a, b) <- (c a + s b, c b - s a): the right-multiplication by an
// elementary rotation about the third axis (reference rx/ry/rz, fknm.cpp:1320-1440)
template <typename real>
__device__ __forceinline__ void rot_cols(real *a, real *b, real s, real c)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        real x = a[i], y = b[i];
        a[i] = fma(c, x, s * y);
        b[i] = fma(c, y, -(s * x));
    }
}

// T <- T * ET(eta) for a joint of the given axis (eta already sign-flipped)
template <typename real>
__device__ __forceinline__ void pose_joint_right(Pose<real> &T, int axis, real eta)
{
    if (axis < 3) {
        real s, c;
        b2k_sincos<real>(eta, &s, &c);
        if (axis == B2K_RZ) rot_cols(T.c0, T.c1, s, c);
        else if (axis == B2K_RX) rot_cols(T.c1, T.c2, s, c);
        else rot_cols(T.c2, T.c0, s, c);
    } else {
        if (axis == B2K_TX) {
#pragma unroll
            for (int i = 0; i < 3; i++) T.p[i] = fma(T.c0[i], eta, T.p[i]);
        } else if (axis == B2K_TY) {
#pragma unroll
            for (int i = 0; i < 3; i++) T.p[i] = fma(T.c1[i], eta, T.p[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++) T.p[i] = fma(T.c2[i], eta, T.p[i]);
        }
    }
}

// rows (IA, IB) of all four columns: a' = c a - s b ; b' = s a + c b  (left-multiplication by
// an elementary rotation about the remaining axis)
template <typename real, int IA, int IB>
__device__ __forceinline__ void rot_rows(Pose<real> &T, real s, real c)
{
    real *cols[4] = {T.c0, T.c1, T.c2, T.p};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        real x = cols[j][IA], y = cols[j][IB];
        cols[j][IA] = fma(c, x, -(s * y));
        cols[j][IB] = fma(s, x, c * y);
    }
}

// T <- ET(eta) * T
template <typename real>
__device__ __forceinline__ void pose_joint_left(Pose<real> &T, int axis, real eta)
{
    if (axis < 3) {
        real s, c;
        b2k_sincos<real>(eta, &s, &c);
        if (axis == B2K_RZ) rot_rows<real, 0, 1>(T, s, c);
        else if (axis == B2K_RX) rot_rows<real, 1, 2>(T, s, c);
        else rot_rows<real, 2, 0>(T, s, c);
    } else {
        if (axis == B2K_TX) T.p[0] += eta;
        else if (axis == B2K_TY) T.p[1] += eta;
        else T.p[2] += eta;
    }
}

// end-effector-frame Jacobian column read off U (reference methods.cpp:243-302):
// revolute about axis AX: lin = r_K1 * p_A - r_K2 * p_B, ang = r_AX, where r_k is row k of U's rotation
template <typename real, int AX>
__device__ __forceinline__ void je_col_rev(const Pose<real> &U, real sg, real *col)
{
    constexpr int K1 = (AX == 0) ? 2 : (AX == 1) ? 0 : 1;
    constexpr int K2 = (AX == 0) ? 1 : (AX == 1) ? 2 : 0;
    constexpr int PA = (AX == 0) ? 1 : (AX == 1) ? 2 : 0;
    constexpr int PB = (AX == 0) ? 2 : (AX == 1) ? 0 : 1;
    const real *cols[3] = {U.c0, U.c1, U.c2};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        col[i] = sg * fma(cols[i][K1], U.p[PA], -(cols[i][K2] * U.p[PB]));
        col[3 + i] = sg * cols[i][AX];
    }
}
template <typename real, int K>
__device__ __forceinline__ void je_col_pri(const Pose<real> &U, real sg, real *col)
{
    const real *cols[3] = {U.c0, U.c1, U.c2};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        col[i] = sg * cols[i][K];
        col[3 + i] = (real)0;
    }
}

// ------------------------------------------------------------------ the forward chain walk (one row, registers only)
// T = A_0 J_0(q) A_1 J_1(q) ... A_{n-1} J_{n-1}(q) A_n.  When WJ, also records for every joint the
// (signed) joint axis z_j and the joint origin p_j in the start frame, from which the base-frame
// Jacobian column is z_j x (p_e - p_j) | z_j (revolute) or z_j | 0 (prismatic).
// getq(j, col) returns this row's coordinate of joint j, stored in column col of q (shared
// memory in the FK kernels; registers, indexed by the compile-time j, in the IK kernel).
template <typename real, int N, bool WJ, bool ALLRZ, typename GetQ>
__device__ __forceinline__ void chain_forward(const ChainP<real, N> &P, GetQ getq, Pose<real> &T,
                                              real (*zj)[3], real (*pj)[3])
{
    pose_from_const(T, P.A[0]);
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (j > 0) pose_mul_const_right(T, P.A[j], P.akind[j]);
        if (ALLRZ) {
            real eta = getq(j, P.jidx[j]);
            if (WJ) {
#pragma unroll
                for (int i = 0; i < 3; i++) { zj[j][i] = T.c2[i]; pj[j][i] = T.p[i]; }
            }
            real s, c;
            b2k_sincos<real>(eta, &s, &c);
            rot_cols(T.c0, T.c1, s, c);
        } else {
            const int ax = P.axis[j];
            real eta = getq(j, P.jidx[j]);
            const real sg = P.flip[j] ? (real)-1 : (real)1;
            eta *= sg;
            if (WJ) {
                const int k = ax < 3 ? ax : ax - 3;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    real col = (k == 0) ? T.c0[i] : (k == 1) ? T.c1[i] : T.c2[i];
                    zj[j][i] = sg * col;
                    pj[j][i] = T.p[i];
                }
            }
            pose_joint_right(T, ax, eta);
        }
    }
    pose_mul_const_right(T, P.A[N], P.akind[N]);
}

// ------------------------------------------------------------------ per-warp staging helpers
// stage row stride in 8-byte granules: odd, so that one-row-per-lane accesses are conflict free
__host__ __device__ constexpr int stage_stride(int granules_per_row) { return granules_per_row | 1; }

template <typename real>
__host__ __device__ constexpr int granules(int elems) { return elems * (int)sizeof(real) / 8; }

// bytes of per-warp shared memory needed by the FK/J kernels
template <typename real>
inline size_t fkj_warp_smem(int n, int ldq, bool wt, bool wj)
{
    size_t q = (size_t)32 * (ldq | 1) * sizeof(real);
    size_t t = wt ? (size_t)32 * stage_stride(granules<real>(16)) * 8 : 0;
    size_t j = wj ? (size_t)32 * stage_stride(granules<real>(6 * n)) * 8 : 0;
    size_t m = q > t ? q : t;
    return m > j ? m : j;
}

// coalesced copy of the warp's q tile into shared memory (row stride ldq -> ldq|1)
template <typename real>
__device__ __forceinline__ void stage_q_tile(real *sq, const real *__restrict__ gq, int rows_here, int ldq,
                                             float inv_ldq, int lane)
{
    const int cnt = rows_here * ldq;
    const int ldqp = ldq | 1;
    for (int i = lane; i < cnt; i += 32) {
        int r = __float2int_rz(((float)i + 0.5f) * inv_ldq);
        int c = i - r * ldq;
        sq[r * ldqp + c] = gq[i];
    }
}

// coalesced write-out of a staged tile: ROWG granules per row, staged with stride S
template <typename real, int ROWG>
__device__ __forceinline__ void drain_tile(const typename Granule<real>::type *stage,
                                           typename Granule<real>::type *__restrict__ gout, int rows_here, int lane)
{
    constexpr int S = stage_stride(ROWG);
    const int lim = rows_here * ROWG;
#pragma unroll 4
    for (int it = 0; it < ROWG; ++it) {
        int i = it * 32 + lane;
        if (i < lim) {
            int r = i / ROWG;
            int c = i - r * ROWG;
            gout[i] = stage[r * S + c];
        }
    }
}

// ------------------------------------------------------------------ forward walk: pose and/or base-frame Jacobian
template <typename real, int N, bool WT, bool WJ, bool ALLRZ>
__global__ void __launch_bounds__(B2K_THREADS)
k_fkj_forward(const __grid_constant__ ChainP<real, N> P, const real *__restrict__ q, long long nrows, int ldq,
              float inv_ldq, real *__restrict__ Tout, real *__restrict__ Jout, int warp_smem_bytes)
{
    typedef typename Granule<real>::type gran_t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *wbase = smem_raw + (size_t)warp * warp_smem_bytes;
    real *sq = reinterpret_cast<real *>(wbase);
    real *so = reinterpret_cast<real *>(wbase);
    const int ldqp = ldq | 1;
    const long long ntiles = (nrows + 31) >> 5;

    for (long long tile = (long long)blockIdx.x * B2K_WARPS_PER_BLOCK + warp; tile < ntiles;
         tile += (long long)gridDim.x * B2K_WARPS_PER_BLOCK) {
        const long long row0 = tile << 5;
        const int rows_here = (int)((nrows - row0) < 32 ? (nrows - row0) : 32);
        stage_q_tile<real>(sq, q + row0 * ldq, rows_here, ldq, inv_ldq, lane);
        __syncwarp();
        const real *myq = sq + (lane < rows_here ? lane : 0) * ldqp;

        Pose<real> T;
        real zj[WJ ? N : 1][3], pj[WJ ? N : 1][3];
        chain_forward<real, N, WJ, ALLRZ>(P, [&](int, int col) { return myq[col]; }, T, zj, pj);
        __syncwarp(); // all lanes are done reading q: the stage may be overwritten

        if (WT) {
            Pose<real> Tb = T;
            if (P.has_base) pose_mul_const_left(Tb, P.B, AK_GEN | AK_TX | AK_TY | AK_TZ);
            constexpr int S = stage_stride(granules<real>(16)) * Granule<real>::PER; // stride in elements
            real *row = so + lane * S;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                row[i * 4 + 0] = Tb.c0[i];
                row[i * 4 + 1] = Tb.c1[i];
                row[i * 4 + 2] = Tb.c2[i];
                row[i * 4 + 3] = Tb.p[i];
            }
            row[12] = (real)0; row[13] = (real)0; row[14] = (real)0; row[15] = (real)1;
            __syncwarp();
            drain_tile<real, granules<real>(16)>(reinterpret_cast<const gran_t *>(so),
                                                 reinterpret_cast<gran_t *>(Tout + row0 * 16), rows_here, lane);
            __syncwarp();
        }
        if (WJ) {
            constexpr int S = stage_stride(granules<real>(6 * N)) * Granule<real>::PER;
            real *row = so + lane * S;
#pragma unroll
            for (int j = 0; j < N; j++) {
                const bool rev = ALLRZ ? true : (P.axis[j] < 3);
                if (rev) {
                    real dx = T.p[0] - pj[j][0], dy = T.p[1] - pj[j][1], dz = T.p[2] - pj[j][2];
                    row[0 * N + j] = fma(zj[j][1], dz, -(zj[j][2] * dy));
                    row[1 * N + j] = fma(zj[j][2], dx, -(zj[j][0] * dz));
                    row[2 * N + j] = fma(zj[j][0], dy, -(zj[j][1] * dx));
                    row[3 * N + j] = zj[j][0];
                    row[4 * N + j] = zj[j][1];
                    row[5 * N + j] = zj[j][2];
                } else {
                    row[0 * N + j] = zj[j][0];
                    row[1 * N + j] = zj[j][1];
                    row[2 * N + j] = zj[j][2];
                    row[3 * N + j] = (real)0;
                    row[4 * N + j] = (real)0;
                    row[5 * N + j] = (real)0;
                }
            }
            __syncwarp();
            drain_tile<real, granules<real>(6 * N)>(reinterpret_cast<const gran_t *>(so),
                                                    reinterpret_cast<gran_t *>(Jout + row0 * (6 * N)), rows_here, lane);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ backward walk: end-effector-frame Jacobian (+ pose)
// Reference _ETS_jacobe, methods.cpp:219-316: U starts at the tool and is left-multiplied by
// each ET walking from the tip to the base; column j is read off U before joint j is applied.
template <typename real, int N, bool WT>
__global__ void __launch_bounds__(B2K_THREADS)
k_fkj_backward(const __grid_constant__ ChainP<real, N> P, const real *__restrict__ q, long long nrows, int ldq,
               float inv_ldq, real *__restrict__ Tout, real *__restrict__ Jout, int warp_smem_bytes)
{
    typedef typename Granule<real>::type gran_t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *wbase = smem_raw + (size_t)warp * warp_smem_bytes;
    real *sq = reinterpret_cast<real *>(wbase);
    real *so = reinterpret_cast<real *>(wbase);
    const int ldqp = ldq | 1;
    const long long ntiles = (nrows + 31) >> 5;

    for (long long tile = (long long)blockIdx.x * B2K_WARPS_PER_BLOCK + warp; tile < ntiles;
         tile += (long long)gridDim.x * B2K_WARPS_PER_BLOCK) {
        const long long row0 = tile << 5;
        const int rows_here = (int)((nrows - row0) < 32 ? (nrows - row0) : 32);
        stage_q_tile<real>(sq, q + row0 * ldq, rows_here, ldq, inv_ldq, lane);
        __syncwarp();
        const real *myq = sq + (lane < rows_here ? lane : 0) * ldqp;

        Pose<real> U;
        real Je[N][6];
        pose_from_const(U, P.A[N]);
#pragma unroll
        for (int j = N - 1; j >= 0; j--) {
            const int ax = P.axis[j];
            const real sg = P.flip[j] ? (real)-1 : (real)1;
            switch (ax) {
            case B2K_RX: je_col_rev<real, 0>(U, sg, Je[j]); break;
            case B2K_RY: je_col_rev<real, 1>(U, sg, Je[j]); break;
            case B2K_RZ: je_col_rev<real, 2>(U, sg, Je[j]); break;
            case B2K_TX: je_col_pri<real, 0>(U, sg, Je[j]); break;
            case B2K_TY: je_col_pri<real, 1>(U, sg, Je[j]); break;
            default: je_col_pri<real, 2>(U, sg, Je[j]); break;
            }
            pose_joint_left(U, ax, sg * myq[P.jidx[j]]);
            pose_mul_const_left(U, P.A[j], P.akind[j]);
        }
        __syncwarp();

        if (WT) {
            Pose<real> Tb = U;
            if (P.has_base) pose_mul_const_left(Tb, P.B, AK_GEN | AK_TX | AK_TY | AK_TZ);
            constexpr int S = stage_stride(granules<real>(16)) * Granule<real>::PER;
            real *row = so + lane * S;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                row[i * 4 + 0] = Tb.c0[i];
                row[i * 4 + 1] = Tb.c1[i];
                row[i * 4 + 2] = Tb.c2[i];
                row[i * 4 + 3] = Tb.p[i];
            }
            row[12] = (real)0; row[13] = (real)0; row[14] = (real)0; row[15] = (real)1;
            __syncwarp();
            drain_tile<real, granules<real>(16)>(reinterpret_cast<const gran_t *>(so),
                                                 reinterpret_cast<gran_t *>(Tout + row0 * 16), rows_here, lane);
            __syncwarp();
        }
        {
            constexpr int S = stage_stride(granules<real>(6 * N)) * Granule<real>::PER;
            real *row = so + lane * S;
#pragma unroll
            for (int j = 0; j < N; j++)
#pragma unroll
                for (int k = 0; k < 6; k++) row[k * N + j] = Je[j][k];
            __syncwarp();
            drain_tile<real, granules<real>(6 * N)>(reinterpret_cast<const gran_t *>(so),
                                                    reinterpret_cast<gran_t *>(Jout + row0 * (6 * N)), rows_here, lane);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ launcher
enum { FKJ_T = 1, FKJ_J0 = 2, FKJ_JE = 4 };

template <typename real, int N>
int fkj_launch_n(const b2k_chain_s *c, int mode, const real *q, long long nrows, int ldq, const double *base,
                 const double *tool, real *T, real *J, cudaStream_t st)
{
    ChainP<real, N> P;
    const bool wt = mode & FKJ_T, wj0 = mode & FKJ_J0, wje = mode & FKJ_JE;
    // pose-only: base folded into the first constant; fused: base applied to the pose at the end
    b2k_fill_chain<real, N>(c, base, tool, /*base_into_chain=*/(wt && !wj0 && !wje), P);
    const size_t wsm = (fkj_warp_smem<real>(N, ldq, wt, wj0 || wje) + 15) & ~(size_t)15;
    const size_t smem = wsm * B2K_WARPS_PER_BLOCK;
    const long long ntiles = (nrows + 31) / 32;
    const long long nblk_needed = (ntiles + B2K_WARPS_PER_BLOCK - 1) / B2K_WARPS_PER_BLOCK;
    const float inv_ldq = 1.0f / (float)ldq;

    auto launch = [&](auto kern) -> int {
        int per_sm = b2k_blocks_per_sm((const void *)kern, B2K_THREADS, smem);
        if (per_sm < 1) return per_sm < 0 ? per_sm : (b2k_set_error("fkj kernel does not fit on an SM (smem %zu B)", smem), B2K_ERR_INVALID);
        long long grid = (long long)b2k_num_sms() * per_sm;
        if (grid > nblk_needed) grid = nblk_needed;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, B2K_THREADS, smem, st>>>(P, q, nrows, ldq, inv_ldq, T, J, (int)wsm);
        b2k_count_launch();
        B2K_CUDA(cudaGetLastError());
        return B2K_OK;
    };

    if (wje) {
        if (wt) return launch(k_fkj_backward<real, N, true>);
        return launch(k_fkj_backward<real, N, false>);
    }
    if (c->all_rz) {
        if (wt && wj0) return launch(k_fkj_forward<real, N, true, true, true>);
        if (wt) return launch(k_fkj_forward<real, N, true, false, true>);
        return launch(k_fkj_forward<real, N, false, true, true>);
    }
    if (wt && wj0) return launch(k_fkj_forward<real, N, true, true, false>);
    if (wt) return launch(k_fkj_forward<real, N, true, false, false>);
    return launch(k_fkj_forward<real, N, false, true, false>);
}

template <typename real>
int fkj_launch(const b2k_chain_s *c, int mode, const void *q, long long nrows, long long ldq, const double *base,
               const double *tool, void *T, void *J, cudaStream_t st)
{
#define B2K_CASE(NN) \
    case NN: return fkj_launch_n<real, NN>(c, mode, (const real *)q, nrows, (int)ldq, base, tool, (real *)T, (real *)J, st);
    switch (c->n) {
        B2K_CASE(1) B2K_CASE(2) B2K_CASE(3) B2K_CASE(4) B2K_CASE(5)
        B2K_CASE(6) B2K_CASE(7) B2K_CASE(8) B2K_CASE(9) B2K_CASE(10)
    default:
        b2k_set_error("fkj: unsupported joint count %d", c->n);
        return B2K_ERR_INVALID;
    }
#undef B2K_CASE
}
