/*
 * oracle_kin.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the reference's native hot path
 * (petercorke/robotics-toolbox-python, src/roboticstoolbox/core/), written
 * from the algorithm, one function per reference function, in the
 * reference's own operation order so it tracks the compiled reference
 * (oracle/_ref) to a few ulp.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 * The product (robotics-toolbox-python_b200/) never links or imports it.
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks every function
 * here against (a) the literal golden vectors of the reference's own tests
 * (tests/golden/reference_kats.json, transcribed with file:line) and (b)
 * fixtures produced by the compiled reference itself
 * (tests/golden/make_golden.py -> the .npz fixtures under tests/golden/).
 *
 * Conventions: all 4x4 matrices are ROW-MAJOR here (the reference keeps them
 * column-major inside Eigen; values are identical).  q batches are (N, ldq)
 * row-major, Jacobians (N, 6, n) row-major, poses (N, 4, 4) row-major - the
 * layouts the reference's batch FK returns (fknm.cpp:1005,1048-1051).
 *
 * Reference map:
 *   orc_et_T        <- _ET_T methods.cpp:354-370, rx..tz fknm.cpp:1320-1555
 *   orc_fkine       <- _ETS_fkine methods.cpp:318-352, loop fknm.cpp:1038-1052
 *   orc_jacob0      <- _ETS_jacob0 methods.cpp:112-216
 *   orc_jacobe      <- _ETS_jacobe methods.cpp:219-316
 *   orc_angle_axis  <- _angle_axis ik.cpp:241-286
 *   orc_ik_lm       <- _IK_loop ik.cpp:19-75, _IK_LM_* ik.cpp:157-209, _IK_NR/_IK_GN ik.cpp:79-155,
 *                      IK_LM_c fknm.cpp:394-525 (semantics 0);
 *                      IKSolver._solve IK.py:297-367 + IK_LM.step IK.py:994-1017
 *                      (semantics 1)
 *   orc_rne         <- rot_mat frne.c:310-351, newton_euler ne.c:62-492
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.14159265358979323846264338327950288   /* linalg.h:19 */
#define ORC_PI_2 1.57079632679489661923132169163975144 /* linalg.h:18 */
#define ORC_PI_X2 6.283185307179586                    /* linalg.h:20 */

/* axis codes as in ET.py:244-266: Rx 0, Ry 1, Rz 2, tx 3, ty 4, tz 5 */

static void mat4_identity(double *m)
{
    memset(m, 0, 16 * sizeof(double));
    m[0] = m[5] = m[10] = m[15] = 1.0;
}

/* c = a * b, full 4x4 product, sequential k (what Eigen's small fixed-size
 * product evaluates, methods.cpp:339) */
static void mat4_mul(const double *a, const double *b, double *c)
{
    double t[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = a[i * 4 + 0] * b[0 * 4 + j];
            s += a[i * 4 + 1] * b[1 * 4 + j];
            s += a[i * 4 + 2] * b[2 * 4 + j];
            s += a[i * 4 + 3] * b[3 * 4 + j];
            t[i * 4 + j] = s;
        }
    memcpy(c, t, sizeof(t));
}

/* One elementary transform evaluated at eta (methods.cpp:354-370). */
void orc_et_T(int isjoint, int axis, int flip, const double *Tconst, double eta, double *out)
{
    if (!isjoint) {
        memcpy(out, Tconst, 16 * sizeof(double));
        return;
    }
    if (flip)
        eta = -eta;
    mat4_identity(out);
    double c, s;
    switch (axis) {
    case 0: /* Rx fknm.cpp:1320-1354 */
        c = cos(eta); s = sin(eta);
        out[5] = c; out[6] = -s; out[9] = s; out[10] = c;
        break;
    case 1: /* Ry */
        c = cos(eta); s = sin(eta);
        out[0] = c; out[2] = s; out[8] = -s; out[10] = c;
        break;
    case 2: /* Rz fknm.cpp:1395-1430 */
        c = cos(eta); s = sin(eta);
        out[0] = c; out[1] = -s; out[4] = s; out[5] = c;
        break;
    case 3: out[3] = eta; break;
    case 4: out[7] = eta; break;
    case 5: out[11] = eta; break;
    default: break;
    }
}

static void fkine_one(int m, const int *isjoint, const int *axis, const int *flip,
                      const int *jindex, const double *Tc, const double *q,
                      const double *base, const double *tool, double *out)
{
    double cur[16], E[16];
    if (base)
        memcpy(cur, base, sizeof(cur));
    else
        mat4_identity(cur);
    for (int i = 0; i < m; i++) {
        /* constants carry a dummy jindex (ET.py:104-105); never dereference q for them */
        double eta = isjoint[i] ? q[jindex[i]] : 0.0;
        orc_et_T(isjoint[i], axis[i], flip[i], Tc + 16 * i, eta, E);
        mat4_mul(cur, E, cur);
    }
    if (tool)
        mat4_mul(cur, tool, out);
    else
        memcpy(out, cur, sizeof(cur));
}

void orc_fkine(int m, const int *isjoint, const int *axis, const int *flip, const int *jindex,
               const double *Tc, const double *q, long N, long ldq,
               const double *base, const double *tool, double *out)
{
#pragma omp parallel for schedule(static)
    for (long r = 0; r < N; r++)
        fkine_one(m, isjoint, axis, flip, jindex, Tc, q + r * ldq, base, tool, out + 16 * r);
}

/* Backward walk shared by jacob0/jacobe (methods.cpp:112-209 / 219-316).
 * tJ is (6, n) row-major; U (4x4) returns the full chain transform. */
static void jac_walk(int m, int n, const int *isjoint, const int *axis, const int *flip,
                     const int *jindex, const double *Tc, const double *q, const double *tool,
                     double *tJ, double *U)
{
    double E[16];
    if (tool)
        memcpy(U, tool, 16 * sizeof(double));
    else
        mat4_identity(U);
    int j = n - 1;
    for (int i = m - 1; i >= 0; i--) {
        if (isjoint[i]) {
            const double *r0 = U + 0, *r1 = U + 4, *r2 = U + 8;
            double px = U[3], py = U[7], pz = U[11];
            double col[6] = {0, 0, 0, 0, 0, 0};
            switch (axis[i]) {
            case 0:
                for (int k = 0; k < 3; k++) { col[k] = r2[k] * py - r1[k] * pz; col[3 + k] = r0[k]; }
                break;
            case 1:
                for (int k = 0; k < 3; k++) { col[k] = r0[k] * pz - r2[k] * px; col[3 + k] = r1[k]; }
                break;
            case 2:
                for (int k = 0; k < 3; k++) { col[k] = r1[k] * px - r0[k] * py; col[3 + k] = r2[k]; }
                break;
            case 3: for (int k = 0; k < 3; k++) col[k] = r0[k]; break;
            case 4: for (int k = 0; k < 3; k++) col[k] = r1[k]; break;
            case 5: for (int k = 0; k < 3; k++) col[k] = r2[k]; break;
            default: break;
            }
            if (flip[i])
                for (int k = 0; k < 6; k++) col[k] = -col[k];
            for (int k = 0; k < 6; k++) tJ[k * n + j] = col[k];
            j--;
        }
        double eta = isjoint[i] ? q[jindex[i]] : 0.0;
        orc_et_T(isjoint[i], axis[i], flip[i], Tc + 16 * i, eta, E);
        mat4_mul(E, U, U);
    }
}

void orc_jacobe(int m, int n, const int *isjoint, const int *axis, const int *flip,
                const int *jindex, const double *Tc, const double *q, long N, long ldq,
                const double *tool, double *J)
{
#pragma omp parallel for schedule(static)
    for (long r = 0; r < N; r++) {
        double U[16];
        jac_walk(m, n, isjoint, axis, flip, jindex, Tc, q + r * ldq, tool, J + r * 6 * n, U);
    }
}

static void jacob0_one(int m, int n, const int *isjoint, const int *axis, const int *flip,
                       const int *jindex, const double *Tc, const double *q, const double *tool,
                       double *J, double *tJ)
{
    double U[16];
    jac_walk(m, n, isjoint, axis, flip, jindex, Tc, q, tool, tJ, U);
    /* J0 = blkdiag(R, R) * tJ, R = U[:3,:3]  (methods.cpp:211-216) */
    for (int c = 0; c < n; c++)
        for (int half = 0; half < 2; half++)
            for (int i = 0; i < 3; i++) {
                double s = U[i * 4 + 0] * tJ[(3 * half + 0) * n + c];
                s += U[i * 4 + 1] * tJ[(3 * half + 1) * n + c];
                s += U[i * 4 + 2] * tJ[(3 * half + 2) * n + c];
                J[(3 * half + i) * n + c] = s;
            }
}

void orc_jacob0(int m, int n, const int *isjoint, const int *axis, const int *flip,
                const int *jindex, const double *Tc, const double *q, long N, long ldq,
                const double *tool, double *J)
{
#pragma omp parallel
    {
        double *tJ = (double *)malloc(sizeof(double) * 6 * (n > 0 ? n : 1));
#pragma omp for schedule(static)
        for (long r = 0; r < N; r++)
            jacob0_one(m, n, isjoint, axis, flip, jindex, Tc, q + r * ldq, tool, J + r * 6 * n, tJ);
        free(tJ);
    }
}

/* fused convenience: T (no base in J, base applied to T like Robot.fkine) */
void orc_fkine_jacob0(int m, int n, const int *isjoint, const int *axis, const int *flip,
                      const int *jindex, const double *Tc, const double *q, long N, long ldq,
                      const double *base, const double *tool, double *T, double *J)
{
    orc_fkine(m, isjoint, axis, flip, jindex, Tc, q, N, ldq, base, tool, T);
    orc_jacob0(m, n, isjoint, axis, flip, jindex, Tc, q, N, ldq, tool, J);
}

/* ik.cpp:241-286.  Te, Tep row-major 4x4. */
void orc_angle_axis(const double *Te, const double *Tep, double *e)
{
    double R[9];
    e[0] = Tep[3] - Te[3];
    e[1] = Tep[7] - Te[7];
    e[2] = Tep[11] - Te[11];
    /* R = Rep * Re^T */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = Tep[i * 4 + 0] * Te[j * 4 + 0];
            s += Tep[i * 4 + 1] * Te[j * 4 + 1];
            s += Tep[i * 4 + 2] * Te[j * 4 + 2];
            R[i * 3 + j] = s;
        }
    double li[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    double li_norm = sqrt(li[0] * li[0] + li[1] * li[1] + li[2] * li[2]);
    double tr = R[0] + R[4] + R[8];
    if (li_norm < 1e-6) {
        if (tr > 0) {
            e[3] = e[4] = e[5] = 0.0;
        } else {
            e[3] = ORC_PI_2 * (R[0] + 1);
            e[4] = ORC_PI_2 * (R[4] + 1);
            e[5] = ORC_PI_2 * (R[8] + 1);
        }
    } else {
        double ang = atan2(li_norm, tr - 1);
        e[3] = ang * li[0] / li_norm;
        e[4] = ang * li[1] / li_norm;
        e[5] = ang * li[2] / li_norm;
    }
}

/* ---- restart RNG: the counter-based generator the CUDA kernel uses (DESIGN.md
 * "IK restarts").  The reference draws from unseeded libc rand() (ik.cpp:293) /
 * numpy default_rng (IK.py:166); neither is reproducible on a device, so the
 * product defines its own stream and the oracle mirrors it. */
static uint64_t orc_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

double orc_rand_u01(uint64_t seed, uint64_t row, uint32_t search, uint32_t joint)
{
    uint64_t h = orc_mix64(seed ^ (0x5851F42D4C957F2DULL * (row + 1)));
    h = orc_mix64(h + (((uint64_t)search << 32) | (uint64_t)joint));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

/* _rand_q ik.cpp:288-299: qlim_l + (U(-1,1) + 1) * range/2 */
static void rand_q(int n, const double *qlim_l, const double *qlim_h, uint64_t seed, uint64_t row,
                   uint32_t search, double *q)
{
    for (int i = 0; i < n; i++) {
        double r = 2.0 * orc_rand_u01(seed, row, search, (uint32_t)i) - 1.0;
        double range2 = (qlim_h[i] - qlim_l[i]) / 2.0; /* fknm.cpp:1104 */
        q[i] = (r + 1.0) * range2 + qlim_l[i];
    }
}

/* Solve A x = g for the LM step.  The reference forms A.inverse()*g with
 * Eigen's dynamic-size inverse (partial-pivot LU, ik.cpp:171); here the same
 * factorisation is used as a solve.  Returns 0 on a zero/non-finite pivot. */
static int lu_solve(int n, double *A, double *b)
{
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; i++)
            if (fabs(A[i * n + k]) > best) { best = fabs(A[i * n + k]); p = i; }
        if (!(best > 0.0) || !isfinite(best))
            return 0;
        if (p != k) {
            for (int j = 0; j < n; j++) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
            double t = b[k]; b[k] = b[p]; b[p] = t;
        }
        for (int i = k + 1; i < n; i++) {
            double f = A[i * n + k] / A[k * n + k];
            for (int j = k; j < n; j++) A[i * n + j] -= f * A[k * n + j];
            b[i] -= f * b[k];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int j = i + 1; j < n; j++) s -= A[i * n + j] * b[j];
        b[i] = s / A[i * n + i];
    }
    return 1;
}

typedef struct {
    int m, n;
    const int *isjoint, *axis, *flip, *jindex;
    const double *Tc;
} orc_chain;

/* pose error and cost at q (ik.cpp:44-46 / IK.py:994-995) */
static void lm_error(const orc_chain *c, const double *Tep, const double *we, const double *q,
                     double *e, double *Eout)
{
    double Te[16];
    fkine_one(c->m, c->isjoint, c->axis, c->flip, c->jindex, c->Tc, q, NULL, NULL, Te);
    orc_angle_axis(Te, Tep, e);
    double E = 0.0;
    for (int k = 0; k < 6; k++) E += e[k] * we[k] * e[k];
    *Eout = 0.5 * E;
}

/* one LM update dq at q given (e, E) (ik.cpp:157-209 / IK.py:997-1015);
 * returns 0 when the normal matrix cannot be factorised */
static int lm_step(const orc_chain *c, const double *we, double lambda, int method,
                   const double *q, const double *e, double E, double *dq, double *work)
{
    int n = c->n;
    double *J = work;          /* 6n */
    double *tJ = work + 6 * n; /* 6n */
    double *A = work + 12 * n; /* n*n */
    jacob0_one(c->m, n, c->isjoint, c->axis, c->flip, c->jindex, c->Tc, q, NULL, J, tJ);
    double wn = (method == 0) ? lambda * E : (method == 1) ? lambda : (E + lambda);
    for (int i = 0; i < n; i++) {
        double g = 0.0;
        for (int k = 0; k < 6; k++) g += J[k * n + i] * we[k] * e[k];
        dq[i] = g;
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int k = 0; k < 6; k++) s += J[k * n + i] * we[k] * J[k * n + j];
            A[i * n + j] = s + (i == j ? wn : 0.0);
        }
    }
    return lu_solve(n, A, dq);
}

/* One-sided (Hestenes) Jacobi SVD of W (rows x cols, row-major, cols <= 6 <= rows or cols <= rows):
 * rotates column pairs until they are mutually orthogonal.  On return W = (input) * V, so the
 * columns of W are u_i * s_i and V (cols x cols) holds the right singular vectors. */
static void onesided_jacobi(int rows, int cols, double *W, double *V)
{
    for (int i = 0; i < cols; i++)
        for (int j = 0; j < cols; j++) V[i * cols + j] = (i == j);
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotated = 0;
        for (int p = 0; p < cols - 1; p++)
            for (int q = p + 1; q < cols; q++) {
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int r = 0; r < rows; r++) {
                    al += W[r * cols + p] * W[r * cols + p];
                    be += W[r * cols + q] * W[r * cols + q];
                    ga += W[r * cols + p] * W[r * cols + q];
                }
                if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
                rotated = 1;
                double zeta = (be - al) / (2.0 * ga);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int r = 0; r < rows; r++) {
                    double wp = W[r * cols + p], wq = W[r * cols + q];
                    W[r * cols + p] = c * wp - sn * wq;
                    W[r * cols + q] = sn * wp + c * wq;
                }
                for (int r = 0; r < cols; r++) {
                    double vp = V[r * cols + p], vq = V[r * cols + q];
                    V[r * cols + p] = c * vp - sn * vq;
                    V[r * cols + q] = sn * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
}

/* Newton-Raphson step dq = pinv_d(J) e (ik.cpp:121-155; _pseudo_inverse ik.cpp:211-224:
 * J = U S V^T, pinv_d = V diag(s / (s^2 + d^2)) U^T), through a one-sided Jacobi SVD (the
 * reference uses Eigen's two-sided JacobiSVD; the pseudo-inverse is unique, so any accurate SVD
 * gives the same operator).  For n >= 6 the SVD of B = J^T (n x 6): B Vb = W, w_i = u_i s_i, and
 * pinv_d(J) e = sum_i w_i (Vb[:,i] . e) / (s_i^2 + d^2).  For n < 6 the SVD of J (6 x n): J Vj = W and
 * pinv_d(J) e = sum_i Vj[:,i] (w_i . e) / (s_i^2 + d^2).
 * (NR without pinv on n == 6, J.inverse()*e, is the same vector wherever J is invertible.)
 * Gauss-Newton step (ik.cpp:79-119): the minimum-norm least-squares solution of
 * (J^T We J) dq = J^T We e (BDCSVD::solve), i.e. pinv(We^1/2 J) We^1/2 e -- same routine on the
 * weighted J and e, no damping (the reference ignores pinv_damping there), singular values below
 * eps * k * s_max treated as zero as Eigen's rank threshold does. */
static int pinv_step(const orc_chain *c, const double *ws, double damping, int truncate, const double *q,
                     const double *e, double *dq, double *work)
{
    int n = c->n;
    double *J = work;          /* 6n, (6, n) row-major */
    double *tJ = work + 6 * n;
    jacob0_one(c->m, n, c->isjoint, c->axis, c->flip, c->jindex, c->Tc, q, NULL, J, tJ);
    double ew[6], V[36], s2[6];
    for (int a = 0; a < 6; a++) {
        ew[a] = ws[a] * e[a];
        for (int j = 0; j < n; j++) J[a * n + j] *= ws[a];
    }
    int k = n >= 6 ? 6 : n;
    double *W = tJ; /* reuse: n x 6 (n >= 6) or 6 x n */
    if (n >= 6) {
        for (int j = 0; j < n; j++)
            for (int a = 0; a < 6; a++) W[j * 6 + a] = J[a * n + j];
        onesided_jacobi(n, 6, W, V);
    } else {
        memcpy(W, J, sizeof(double) * 6 * n);
        onesided_jacobi(6, n, W, V);
    }
    int rows = n >= 6 ? n : 6;
    double smax2 = 0.0;
    for (int i = 0; i < k; i++) {
        double t = 0.0;
        for (int r = 0; r < rows; r++) t += W[r * k + i] * W[r * k + i];
        s2[i] = t;
        if (t > smax2) smax2 = t;
    }
    double rel = 2.220446049250313e-16 * (n > 6 ? n : 6);
    double thr = truncate ? rel * rel * smax2 : 0.0; /* threshold on s^2 */
    for (int j = 0; j < n; j++) dq[j] = 0.0;
    for (int i = 0; i < k; i++) {
        double den = s2[i] + damping * damping;
        if (!(s2[i] > thr) || !(den > 0.0)) continue;
        if (n >= 6) {
            double t = 0.0;
            for (int a = 0; a < 6; a++) t += V[a * 6 + i] * ew[a];
            t /= den;
            for (int j = 0; j < n; j++) dq[j] += W[j * 6 + i] * t;
        } else {
            double t = 0.0;
            for (int a = 0; a < 6; a++) t += W[a * n + i] * ew[a];
            t /= den;
            for (int j = 0; j < n; j++) dq[j] += V[j * n + i] * t;
        }
    }
    for (int j = 0; j < n; j++)
        if (!isfinite(dq[j])) return 0;
    return 1;
}

/* one solver update: LM (methods 0-2), Newton-Raphson (3), Gauss-Newton (4) */
static int ik_step(const orc_chain *c, const double *we, double lambda, int method,
                   const double *q, const double *e, double E, double *dq, double *work)
{
    if (method == 3) {
        const double one[6] = {1, 1, 1, 1, 1, 1};
        return pinv_step(c, one, lambda, 0, q, e, dq, work);
    }
    if (method == 4) {
        double ws[6];
        for (int a = 0; a < 6; a++) ws[a] = sqrt(we[a]);
        return pinv_step(c, ws, 0.0, 1, q, e, dq, work);
    }
    return lm_step(c, we, lambda, method, q, e, E, dq, work);
}

static int check_lim(int n, const double *q, const double *ql, const double *qh)
{
    for (int i = 0; i < n; i++)
        if (q[i] < ql[i] || q[i] > qh[i])
            return 0;
    return 1;
}

/*
 * Batched IK.  method: 0 chan, 1 wampler, 2 sugihara (LM); 3 Newton-Raphson with the damped
 * pseudo-inverse (lambda = pinv_damping, _IK_NR ik.cpp:121-155); 4 Gauss-Newton (_IK_GN ik.cpp:79-119).
 * semantics 0 = the C++ loop (fknm.IK_LM_c): test E before stepping, wrap with fmod,
 *               counters as ik.cpp:39-69 (it, search start at 0/1; iter restarts at 0).
 * semantics 1 = the Python IK_LM solver (ikine_LM): step first, test the pre-step E,
 *               return the post-step q, wrap with Python's floor-modulo.
 * rng_per_row: 1 -> restart draws keyed by problem row; 0 -> every row shares the
 *               per-search draws (what IKSolver.solve does for a trajectory, IK.py:222-272).
 */
void orc_ik_lm(int m, int n, const int *isjoint, const int *axis, const int *flip,
               const int *jindex, const double *Tc, const double *qlim_l, const double *qlim_h,
               const double *Tep, long N, const double *q0, int ilimit, int slimit, double tol,
               int reject_jl, const double *we_in, double lambda, int method, uint64_t seed,
               int semantics, int rng_per_row, double *q_out, int *success, int *iterations,
               int *searches, double *residual)
{
    orc_chain c = {m, n, isjoint, axis, flip, jindex, Tc};
    double we[6] = {1, 1, 1, 1, 1, 1};
    if (we_in)
        memcpy(we, we_in, sizeof(we));
#pragma omp parallel
    {
        double *work = (double *)malloc(sizeof(double) * (12 * n + n * n + 2 * n + 8));
        double *dq = work + 12 * n + n * n;
        double *q = dq + n;
#pragma omp for schedule(dynamic, 64)
        for (long r = 0; r < N; r++) {
            const double *T = Tep + 16 * r;
            uint64_t row = rng_per_row ? (uint64_t)r : 0;
            double e[6], E = 0.0;
            if (semantics == 0) {
                int it = 0, search = 1, solution = 0, iter = 1;
                if (q0) memcpy(q, q0 + r * n, sizeof(double) * n);
                else rand_q(n, qlim_l, qlim_h, seed, row, 0, q);
                while (search <= slimit) {
                    while (iter <= ilimit) {
                        lm_error(&c, T, we, q, e, &E);
                        if (E < tol) {
                            for (int i = 0; i < n; i++)
                                q[i] = fmod(q[i] + ORC_PI, ORC_PI_X2) - ORC_PI;
                            solution = reject_jl ? check_lim(n, q, qlim_l, qlim_h) : 1;
                            break;
                        }
                        int ok = ik_step(&c, we, lambda, method, q, e, E, dq, work);
                        if (!ok) { iter++; break; } /* singular normal matrix: abandon this search */
                        for (int i = 0; i < n; i++) q[i] += dq[i];
                        iter++;
                    }
                    if (solution) { it += iter; break; }
                    it += iter;
                    iter = 0;
                    search++;
                    rand_q(n, qlim_l, qlim_h, seed, row, (uint32_t)(search - 1), q);
                }
                memcpy(q_out + r * n, q, sizeof(double) * n);
                success[r] = solution; iterations[r] = it; searches[r] = search; residual[r] = E;
            } else {
                int total_i = 0, done = 0;
                for (int search = 0; search < slimit && !done; search++) {
                    if (search == 0 && q0) memcpy(q, q0 + r * n, sizeof(double) * n);
                    else rand_q(n, qlim_l, qlim_h, seed, row, (uint32_t)search, q);
                    int i = 0;
                    while (i < ilimit) {
                        i++;
                        lm_error(&c, T, we, q, e, &E);
                        int ok = ik_step(&c, we, lambda, method, q, e, E, dq, work);
                        if (!ok) break; /* numpy LinAlgError: abandon search (IK.py:321-324) */
                        for (int k = 0; k < n; k++) q[k] += dq[k];
                        if (E < tol) {
                            for (int k = 0; k < n; k++) {
                                double w = fmod(q[k] + ORC_PI, 2 * ORC_PI);
                                if (w < 0) w += 2 * ORC_PI; /* Python % (IK.py:331) */
                                q[k] = w - ORC_PI;
                            }
                            int valid = check_lim(n, q, qlim_l, qlim_h);
                            if (!valid && reject_jl) break;
                            memcpy(q_out + r * n, q, sizeof(double) * n);
                            success[r] = 1; iterations[r] = total_i + i; searches[r] = search + 1;
                            residual[r] = E; done = 1;
                            break;
                        }
                    }
                    total_i += i;
                }
                if (!done) {
                    memcpy(q_out + r * n, q, sizeof(double) * n);
                    success[r] = 0; iterations[r] = total_i; searches[r] = slimit; residual[r] = E;
                }
            }
        }
        free(work);
    }
}

/* ------------------------------------------------------------------ RNE */
typedef struct { double x, y, z; } vec3;
typedef struct { vec3 n, o, a; } rot3; /* columns, as frne.h Rot */

static vec3 v_add(vec3 a, vec3 b) { vec3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static vec3 v_cross(vec3 a, vec3 b)
{
    vec3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
static vec3 v_scale(vec3 a, double s) { vec3 r = {s * a.x, s * a.y, s * a.z}; return r; }
static double v_dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static vec3 rot_mul(const rot3 *m, vec3 v) /* vmath.c rot_vect_mult */
{
    vec3 r = {m->n.x * v.x + m->o.x * v.y + m->a.x * v.z,
              m->n.y * v.x + m->o.y * v.y + m->a.y * v.z,
              m->n.z * v.x + m->o.z * v.y + m->a.z * v.z};
    return r;
}
static vec3 rot_t_mul(const rot3 *m, vec3 v) /* vmath.c rot_trans_vect_mult */
{
    vec3 r = {m->n.x * v.x + m->n.y * v.y + m->n.z * v.z,
              m->o.x * v.x + m->o.y * v.y + m->o.z * v.z,
              m->a.x * v.x + m->a.y * v.y + m->a.z * v.z};
    return r;
}
static vec3 inertia_mul(const double *I, vec3 v) /* vmath.c mat_vect_mult, column-major read */
{
    vec3 r = {I[0] * v.x + I[3] * v.y + I[6] * v.z,
              I[1] * v.x + I[4] * v.y + I[7] * v.z,
              I[2] * v.x + I[5] * v.y + I[8] * v.z};
    return r;
}

#define ORC_MAXLINKS 32

/* one row: L = 24 doubles per link as packed by DHRobot._init_rne (DHRobot.py:1340-1361);
 * grav = the vector handed to frne (i.e. -robot.gravity, DHRobot.py:1449) */
static void rne_one(int n, int mdh, const double *L, const double *grav, const double *q,
                    const double *qd, const double *qdd, const double *fext, double *tau)
{
    rot3 R[ORC_MAXLINKS];
    vec3 pstar[ORC_MAXLINKS], w[ORC_MAXLINKS], wd[ORC_MAXLINKS], acc[ORC_MAXLINKS],
        abar[ORC_MAXLINKS], f[ORC_MAXLINKS], nn[ORC_MAXLINKS];
    const vec3 z0 = {0, 0, 1}, zero = {0, 0, 0};
    vec3 gravity = {grav[0], grav[1], grav[2]};
    vec3 f_tip = zero, n_tip = zero;
    if (fext) {
        f_tip.x = fext[0]; f_tip.y = fext[1]; f_tip.z = fext[2];
        n_tip.x = fext[3]; n_tip.y = fext[4]; n_tip.z = fext[5];
    }
    /* rot_mat frne.c:310-351 */
    for (int j = 0; j < n; j++) {
        const double *l = L + 24 * j;
        double alpha = l[0], A = l[1], theta = l[2], D = l[3], offset = l[5];
        int prismatic = ((int)l[4]) != 0;
        double th = prismatic ? theta : q[j] + offset;
        double d = prismatic ? q[j] + offset : D;
        double st = sin(th), ct = cos(th), sa = sin(alpha), ca = cos(alpha);
        if (!mdh) {
            R[j].n.x = ct; R[j].o.x = -ca * st; R[j].a.x = sa * st;
            R[j].n.y = st; R[j].o.y = ca * ct;  R[j].a.y = -sa * ct;
            R[j].n.z = 0;  R[j].o.z = sa;       R[j].a.z = ca;
            pstar[j].x = A; pstar[j].y = d * sa; pstar[j].z = d * ca;
        } else {
            R[j].n.x = ct;      R[j].o.x = -st;     R[j].a.x = 0;
            R[j].n.y = st * ca; R[j].o.y = ca * ct; R[j].a.y = -sa;
            R[j].n.z = st * sa; R[j].o.z = ct * sa; R[j].a.z = ca;
            pstar[j].x = A; pstar[j].y = -d * sa; pstar[j].z = d * ca;
        }
    }
    /* forward recursion */
    for (int j = 0; j < n; j++) {
        const double *l = L + 24 * j;
        int prismatic = ((int)l[4]) != 0;
        vec3 qdv = {0, 0, qd[j]}, qddv = {0, 0, qdd[j]};
        vec3 rcog = {l[7], l[8], l[9]};
        vec3 t1, t2, t3;
        if (mdh) { /* ne.c:137-240 */
            if (!prismatic) {
                if (j == 0) {
                    w[j] = qdv;
                    wd[j] = qddv;
                    t1 = gravity;
                } else {
                    t1 = rot_t_mul(&R[j], w[j - 1]);
                    w[j] = v_add(t1, qdv);
                    t3 = rot_t_mul(&R[j], wd[j - 1]);
                    t2 = v_cross(t1, qdv);
                    t1 = v_add(t2, t3);
                    wd[j] = v_add(t1, qddv);
                    t1 = v_cross(w[j - 1], pstar[j]);
                    t2 = v_cross(w[j - 1], t1);
                    t1 = v_cross(wd[j - 1], pstar[j]);
                    t1 = v_add(t1, t2);
                    t1 = v_add(t1, acc[j - 1]);
                }
                acc[j] = rot_t_mul(&R[j], t1);
            } else {
                if (j == 0) {
                    w[j] = qdv;   /* sic: ne.c:187-188 */
                    wd[j] = qddv; /* sic: ne.c:195-196 */
                    acc[j] = gravity;
                } else {
                    w[j] = rot_t_mul(&R[j], w[j - 1]);
                    wd[j] = rot_t_mul(&R[j], wd[j - 1]);
                    t1 = v_cross(wd[j - 1], pstar[j]);
                    t3 = v_cross(w[j - 1], pstar[j]);
                    t2 = v_cross(w[j - 1], t3);
                    t1 = v_add(t1, t2);
                    t1 = v_add(t1, acc[j - 1]);
                    acc[j] = rot_t_mul(&R[j], t1);
                    t2 = rot_t_mul(&R[j], w[j - 1]);
                    t1 = v_cross(t2, qdv);
                    t1 = v_scale(t1, 2.0);
                    acc[j] = v_add(acc[j], t1);
                    acc[j] = v_add(acc[j], qddv);
                }
            }
        } else { /* ne.c:245-331 */
            if (!prismatic) {
                t1 = (j == 0) ? qdv : v_add(w[j - 1], qdv);
                w[j] = rot_t_mul(&R[j], t1);
                if (j == 0)
                    t3 = qddv;
                else {
                    t1 = v_add(wd[j - 1], qddv);
                    t2 = v_cross(w[j - 1], qdv);
                    t3 = v_add(t1, t2);
                }
                wd[j] = rot_t_mul(&R[j], t3);
                t1 = v_cross(wd[j], pstar[j]);
                t2 = v_cross(w[j], pstar[j]);
                t3 = v_cross(w[j], t2);
                acc[j] = v_add(t1, t3);
                t1 = rot_t_mul(&R[j], (j == 0) ? gravity : acc[j - 1]);
                acc[j] = v_add(acc[j], t1);
            } else {
                if (j == 0) {
                    w[j] = zero;
                    wd[j] = zero;
                    t1 = v_add(qddv, gravity);
                    acc[j] = rot_t_mul(&R[j], t1);
                } else {
                    w[j] = rot_t_mul(&R[j], w[j - 1]);
                    wd[j] = rot_t_mul(&R[j], wd[j - 1]);
                    t1 = v_add(qddv, acc[j - 1]);
                    acc[j] = rot_t_mul(&R[j], t1);
                }
                t1 = v_cross(wd[j], pstar[j]);
                acc[j] = v_add(acc[j], t1);
                t1 = rot_t_mul(&R[j], qdv);
                t2 = v_cross(w[j], t1);
                t2 = v_scale(t2, 2.0);
                acc[j] = v_add(acc[j], t2);
                t2 = v_cross(w[j], pstar[j]);
                t3 = v_cross(w[j], t2);
                acc[j] = v_add(acc[j], t3);
            }
        }
        /* abar ne.c:228-232 / 335-339 */
        t1 = v_cross(wd[j], rcog);
        t2 = v_cross(w[j], rcog);
        t3 = v_cross(w[j], t2);
        abar[j] = v_add(t1, t3);
        abar[j] = v_add(abar[j], acc[j]);
    }
    /* backward recursion */
    for (int j = n - 1; j >= 0; j--) {
        const double *l = L + 24 * j;
        double mass = l[6];
        vec3 rcog = {l[7], l[8], l[9]};
        const double *I = l + 10;
        vec3 t1, t2, t3, t4;
        if (mdh) { /* ne.c:358-403 */
            vec3 F = v_scale(abar[j], mass);
            t1 = (j == n - 1) ? f_tip : rot_mul(&R[j + 1], f[j + 1]);
            f[j] = v_add(t1, F);
            t2 = inertia_mul(I, wd[j]);
            t3 = inertia_mul(I, w[j]);
            t4 = v_cross(w[j], t3);
            vec3 Nv = v_add(t2, t4);
            if (j == n - 1)
                t1 = n_tip;
            else {
                t1 = rot_mul(&R[j + 1], nn[j + 1]);
                t4 = rot_mul(&R[j + 1], f[j + 1]);
                t3 = v_cross(pstar[j + 1], t4);
                t1 = v_add(t1, t3);
            }
            t2 = v_cross(rcog, F);
            t1 = v_add(t1, t2);
            nn[j] = v_add(t1, Nv);
        } else { /* ne.c:409-457 */
            t4 = v_scale(abar[j], mass);
            if (j != n - 1) {
                t1 = rot_mul(&R[j + 1], f[j + 1]);
                f[j] = v_add(t4, t1);
            } else
                f[j] = v_add(t4, f_tip);
            t2 = v_add(pstar[j], rcog);
            t1 = v_cross(t2, t4);
            if (j != n - 1) {
                t2 = rot_t_mul(&R[j + 1], pstar[j]);
                t3 = v_cross(t2, f[j + 1]);
                t3 = v_add(t3, nn[j + 1]);
                t2 = rot_mul(&R[j + 1], t3);
                t1 = v_add(t1, t2);
            } else {
                t2 = v_cross(pstar[j], f_tip);
                t1 = v_add(t1, t2);
                t1 = v_add(t1, n_tip);
            }
            t2 = inertia_mul(I, wd[j]);
            t3 = inertia_mul(I, w[j]);
            t4 = v_cross(w[j], t3);
            t2 = v_add(t2, t4);
            nn[j] = v_add(t1, t2);
        }
    }
    /* joint torques ne.c:464-491 */
    for (int j = 0; j < n; j++) {
        const double *l = L + 24 * j;
        int prismatic = ((int)l[4]) != 0;
        double Jm = l[19], G = l[20], B = l[21], Tcp = l[22], Tcm = l[23];
        vec3 t1 = mdh ? z0 : rot_t_mul(&R[j], z0);
        double t = prismatic ? v_dot(f[j], t1) : v_dot(nn[j], t1);
        t += G * G * Jm * qdd[j];
        t += G * G * B * qd[j];
        t += fabs(G) * ((qd[j] > 0 ? Tcp : 0.0) + (qd[j] < 0 ? Tcm : 0.0));
        tau[j] = t;
    }
}

void orc_rne(int n, int mdh, const double *L, const double *grav, const double *q,
             const double *qd, const double *qdd, long N, const double *fext, double *tau)
{
    if (n > ORC_MAXLINKS)
        return;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < N; r++)
        rne_one(n, mdh, L, grav, q + r * n, qd + r * n, qdd + r * n, fext, tau + r * n);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int t)
{
#ifdef _OPENMP
    if (t > 0)
        omp_set_num_threads(t);
#else
    (void)t;
#endif
}
