// b2k_rne.cuh -- batched recursive Newton-Euler inverse dynamics for DH / MDH arms (sm_100a).
//
// Replaces the per-row frne.frne loop of DHRobot.rne (reference DHRobot.py:1442-1451 ->
// frne.c:106-230 -> newton_euler ne.c:62-492, rot_mat frne.c:310-351).
//
// Mapping (DESIGN.md "Kernel K3"): like the FK kernel, a warp owns a tile of 32 rows.  The
// (q, qd, qdd) tiles are loaded with coalesced 8-byte-granule loads into a per-warp smem stage
// and stay there for both recursions; lane l runs the Luh-Walker-Paul recursion for row l in
// registers with all link constants (sin/cos(alpha), a, d, m, r, I, G^2 Jm, ...) coming from
// the constant bank (kernel parameter struct).  The forward recursion keeps only what the
// backward one needs per link: F_j = m a_c, N_j = I wd + w x (I w), sin/cos(theta_j).  tau is
// staged and written back coalesced.  HBM-bound: 4n reals per row (192 B for Puma560 fp64).
//
// The arithmetic mirrors ne.c operation by operation (including its quirks for a prismatic
// first joint under modified DH, ne.c:187-204) so results track the reference to rounding.
#pragma once

#include "b2k_fkj.cuh"

template <typename real, int N>
struct RneP {
    // per link
    real sa[N], ca[N]; // sin/cos(alpha) evaluated on the host with libm, as frne.c:325-326 does per call
    real A[N], D[N], st0[N], ct0[N], offset[N];
    real m[N], r[N][3], I[N][9];
    real c_jm[N];  // G*G*Jm
    real c_b[N];   // G*G*B
    real c_tcp[N]; // |G|*Tc+
    real c_tcm[N]; // |G|*Tc-
    int prismatic[N];
    real psrc[N][3]; // p* + r (standard DH backward recursion, ne.c:415-417), summed on the host
    real ps[N][3]; // p* of a revolute link: (a, d sin(alpha), d cos(alpha)) for DH, (a, -d sin(alpha), d cos(alpha)) for MDH
    real grav[3];
    real fext[6];
    TrigC<real> trig;
};

template <typename real>
struct V3 {
    real x, y, z;
};
template <typename real> __device__ __forceinline__ V3<real> vadd(V3<real> a, V3<real> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename real> __device__ __forceinline__ V3<real> vcross(V3<real> a, V3<real> b)
{
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// a x b + c with the addition folded into the products (2 FMA per component instead of MUL + FMA + ADD)
template <typename real> __device__ __forceinline__ V3<real> vcross_acc(V3<real> a, V3<real> b, V3<real> c)
{
    return {fma(a.y, b.z, fma(-a.z, b.y, c.x)), fma(a.z, b.x, fma(-a.x, b.z, c.y)), fma(a.x, b.y, fma(-a.y, b.x, c.z))};
}
template <typename real> __device__ __forceinline__ V3<real> vscale(V3<real> a, real s) { return {s * a.x, s * a.y, s * a.z}; }
template <typename real> __device__ __forceinline__ real vdot(V3<real> a, V3<real> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// link rotation R (columns n, o, a) from sin/cos(theta), sin/cos(alpha); frne.c:329-347
template <typename real, bool MDH>
struct LinkRot {
    real nx, ny, nz, ox, oy, oz, ax, ay, az;
    __device__ __forceinline__ LinkRot(real st, real ct, real sa, real ca)
    {
        if (!MDH) {
            nx = ct; ox = -ca * st; ax = sa * st;
            ny = st; oy = ca * ct;  ay = -sa * ct;
            nz = 0;  oz = sa;       az = ca;
        } else {
            nx = ct;      ox = -st;     ax = 0;
            ny = st * ca; oy = ca * ct; ay = -sa;
            nz = st * sa; oz = ct * sa; az = ca;
        }
    }
    __device__ __forceinline__ V3<real> mul(V3<real> v) const
    {
        return {nx * v.x + ox * v.y + ax * v.z, ny * v.x + oy * v.y + ay * v.z, nz * v.x + oz * v.y + az * v.z};
    }
    __device__ __forceinline__ V3<real> tmul(V3<real> v) const
    {
        return {nx * v.x + ny * v.y + nz * v.z, ox * v.x + oy * v.y + oz * v.z, ax * v.x + ay * v.y + az * v.z};
    }
};

// ------------------------------------------------------------------ all-revolute fast path
// The same recursion as ne.c with the link rotation applied in its factored form
// (R = Rz(theta) Rx(alpha) for DH, Rx(alpha) Rz(theta) for MDH: 8 flops instead of the 9-entry
// product with its structural zeros), z-only joint-rate vectors expanded by hand (the compiler
// may not drop IEEE multiplications by a literal 0), p* taken from the constant bank, and
// R_{j+1} f_{j+1} computed once per link.  Dropping exact-zero terms does not change any value.
template <typename real, bool MDH>
struct FRot { // factored link rotation
    real st, ct, sa, ca;
    __device__ __forceinline__ V3<real> tmul(V3<real> v) const // R^T v
    {
        if (!MDH) {
            const real u = fma(ct, v.x, st * v.y), w = fma(ct, v.y, -(st * v.x));
            return {u, fma(ca, w, sa * v.z), fma(ca, v.z, -(sa * w))};
        } else {
            const real wy = fma(ca, v.y, sa * v.z), wz = fma(ca, v.z, -(sa * v.y));
            return {fma(ct, v.x, st * wy), fma(ct, wy, -(st * v.x)), wz};
        }
    }
    __device__ __forceinline__ V3<real> mul(V3<real> v) const // R v
    {
        if (!MDH) {
            const real m = fma(ca, v.y, -(sa * v.z)), z = fma(sa, v.y, ca * v.z);
            return {fma(ct, v.x, -(st * m)), fma(st, v.x, ct * m), z};
        } else {
            const real u = fma(ct, v.x, -(st * v.y)), w = fma(st, v.x, ct * v.y);
            return {u, fma(ca, w, -(sa * v.z)), fma(sa, w, ca * v.z)};
        }
    }
    // R^T v + c and R v + c with c folded into the last stage of FMAs (no separate additions)
    __device__ __forceinline__ V3<real> tmul_acc(V3<real> v, V3<real> c) const
    {
        if (!MDH) {
            const real w = fma(ct, v.y, -(st * v.x));
            return {fma(ct, v.x, fma(st, v.y, c.x)), fma(ca, w, fma(sa, v.z, c.y)), fma(ca, v.z, fma(-sa, w, c.z))};
        } else {
            const real wy = fma(ca, v.y, sa * v.z);
            return {fma(ct, v.x, fma(st, wy, c.x)), fma(ct, wy, fma(-st, v.x, c.y)), fma(ca, v.z, fma(-sa, v.y, c.z))};
        }
    }
    __device__ __forceinline__ V3<real> mul_acc(V3<real> v, V3<real> c) const
    {
        if (!MDH) {
            const real m = fma(ca, v.y, -(sa * v.z));
            return {fma(ct, v.x, fma(-st, m, c.x)), fma(st, v.x, fma(ct, m, c.y)), fma(sa, v.y, fma(ca, v.z, c.z))};
        } else {
            const real w = fma(st, v.x, ct * v.y);
            return {fma(ct, v.x, fma(-st, v.y, c.x)), fma(ca, w, fma(-sa, v.z, c.y)), fma(sa, w, fma(ca, v.z, c.z))};
        }
    }
};

template <typename real>
__device__ __forceinline__ V3<real> imul(const real *I, V3<real> v) // vmath.c mat_vect_mult (column-major read)
{
    return {fma(I[0], v.x, fma(I[3], v.y, I[6] * v.z)), fma(I[1], v.x, fma(I[4], v.y, I[7] * v.z)),
            fma(I[2], v.x, fma(I[5], v.y, I[8] * v.z))};
}

// sin / cos of every link's joint angle as one interleaved batch (frne.c:193-207 rot_mat); prismatic
// links use the fixed theta of the link
template <typename real, int N, bool ALLREV>
__device__ __forceinline__ void rne_sincos(const RneP<real, N> &P, const real *mq, real *sth, real *cth)
{
    real th[N];
#pragma unroll
    for (int j = 0; j < N; j++) th[j] = mq[j] + P.offset[j];
    b2k_sincos_batch<real, N>(th, P.trig, sth, cth);
    if (!ALLREV) {
#pragma unroll
        for (int j = 0; j < N; j++)
            if (P.prismatic[j]) { sth[j] = P.st0[j]; cth[j] = P.ct0[j]; }
    }
}

template <typename real, int N, bool MDH>
__device__ __forceinline__ void rne_row_allrev(const RneP<real, N> &P, const real *sth, const real *cth, const real *mqd,
                                               const real *mqdd, V3<real> gravity, real *tq)
{
    V3<real> Fm[N], Nm[N];
    V3<real> w = {0, 0, 0}, wd = {0, 0, 0}, acc = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < N; j++) {
        const FRot<real, MDH> R = {sth[j], cth[j], P.sa[j], P.ca[j]};
        const V3<real> ps = {P.ps[j][0], P.ps[j][1], P.ps[j][2]};
        const real qd = mqd[j], qdd = mqdd[j];
        V3<real> wn, wdn, accn;
        if (MDH) { // ne.c:144-181
            if (j == 0) {
                wn = {0, 0, qd};
                wdn = {0, 0, qdd};
                accn = R.tmul(gravity);
            } else {
                const V3<real> t1 = R.tmul(w);
                wn = {t1.x, t1.y, t1.z + qd};
                const V3<real> t3 = R.tmul(wd);
                wdn = {fma(t1.y, qd, t3.x), fma(-t1.x, qd, t3.y), t3.z + qdd}; // t1 x (0,0,qd) + t3 + (0,0,qdd)
                V3<real> a = vcross(w, ps);
                a = vcross_acc(w, a, acc);
                a = vcross_acc(wd, ps, a);
                accn = R.tmul(a);
            }
        } else { // ne.c:252-288
            if (j == 0) {
                wn = R.tmul(V3<real>{0, 0, qd});
                wdn = R.tmul(V3<real>{0, 0, qdd});
            } else {
                wn = R.tmul(V3<real>{w.x, w.y, w.z + qd});
                wdn = R.tmul(V3<real>{fma(w.y, qd, wd.x), fma(-w.x, qd, wd.y), wd.z + qdd});
            }
            const V3<real> t2 = vcross(wn, ps);
            accn = R.tmul_acc(j == 0 ? gravity : acc, vcross_acc(wdn, ps, vcross(wn, t2)));
        }
        w = wn; wd = wdn; acc = accn;
        const V3<real> rc = {P.r[j][0], P.r[j][1], P.r[j][2]};
        const V3<real> abar = vcross_acc(wd, rc, vcross_acc(w, vcross(w, rc), acc));
        Fm[j] = vscale(abar, P.m[j]);
        Nm[j] = vcross_acc(w, imul(P.I[j], w), imul(P.I[j], wd));
    }
    V3<real> f = {0, 0, 0}, nn = {0, 0, 0};
    const V3<real> f_tip = {P.fext[0], P.fext[1], P.fext[2]}, n_tip = {P.fext[3], P.fext[4], P.fext[5]};
#pragma unroll
    for (int j = N - 1; j >= 0; j--) {
        const V3<real> rc = {P.r[j][0], P.r[j][1], P.r[j][2]};
        V3<real> fj, nj;
        if (MDH) { // ne.c:358-398
            if (j == N - 1) {
                fj = vadd(f_tip, Fm[j]);
                nj = vadd(n_tip, Nm[j]);
            } else {
                const FRot<real, MDH> Rn = {sth[j + 1], cth[j + 1], P.sa[j + 1], P.ca[j + 1]};
                const V3<real> psn = {P.ps[j + 1][0], P.ps[j + 1][1], P.ps[j + 1][2]};
                const V3<real> Rf = Rn.mul(f);
                fj = vadd(Rf, Fm[j]);
                nj = Rn.mul_acc(nn, vcross_acc(psn, Rf, Nm[j]));
            }
            nj = vcross_acc(rc, Fm[j], nj);
        } else { // ne.c:409-453
            const V3<real> ps = {P.ps[j][0], P.ps[j][1], P.ps[j][2]};
            const V3<real> psrc = {P.psrc[j][0], P.psrc[j][1], P.psrc[j][2]};
            V3<real> t1 = vcross_acc(psrc, Fm[j], Nm[j]);
            if (j != N - 1) {
                const FRot<real, MDH> Rn = {sth[j + 1], cth[j + 1], P.sa[j + 1], P.ca[j + 1]};
                fj = Rn.mul_acc(f, Fm[j]);
                const V3<real> t3 = vcross_acc(Rn.tmul(ps), f, nn);
                t1 = Rn.mul_acc(t3, t1);
            } else {
                fj = vadd(Fm[j], f_tip);
                t1 = vadd(vcross_acc(ps, f_tip, t1), n_tip);
            }
            nj = t1;
        }
        f = fj; nn = nj;
        real t = MDH ? nj.z : fma(nj.y, P.sa[j], nj.z * P.ca[j]); // n . (R^T z0)
        const real qdj = mqd[j];
        t = fma(P.c_jm[j], mqdd[j], t);
        t = fma(P.c_b[j], qdj, t);
        t += (qdj > 0 ? P.c_tcp[j] : (real)0) + (qdj < 0 ? P.c_tcm[j] : (real)0);
        tq[j] = t;
    }
}

// ------------------------------------------------------------------ generic row (any mix of revolute / prismatic links)
// Mirrors ne.c operation by operation, including its quirks for a prismatic first joint under
// modified DH (ne.c:187-204).
template <typename real, int N, bool MDH>
__device__ __forceinline__ void rne_row_generic(const RneP<real, N> &P, const real *mq, const real *sth, const real *cth,
                                                const real *mqd, const real *mqdd, V3<real> gravity, real *tq)
{
    // stash for the backward recursion
    V3<real> Fm[N], Nm[N];
    // (gravity: base acceleration handed in by the caller)
    V3<real> w = {0, 0, 0}, wd = {0, 0, 0}, acc = {0, 0, 0}; // of link j-1 on entry

    // ---------------- forward recursion (ne.c:137-240 MDH, 245-347 DH)
#pragma unroll
    for (int j = 0; j < N; j++) {
        const bool pris = (P.prismatic[j] != 0);
        const real st = sth[j], ct = cth[j];
        const real d = pris ? (mq[j] + P.offset[j]) : P.D[j];
        const LinkRot<real, MDH> R(st, ct, P.sa[j], P.ca[j]);
        const V3<real> pstar = MDH ? V3<real>{P.A[j], -d * P.sa[j], d * P.ca[j]} : V3<real>{P.A[j], d * P.sa[j], d * P.ca[j]};
        const V3<real> qdv = {0, 0, mqd[j]}, qddv = {0, 0, mqdd[j]};
        V3<real> wn, wdn, accn, t1, t2, t3;
        if (MDH) {
            if (!pris) {
                if (j == 0) {
                    wn = qdv; wdn = qddv; t1 = gravity;
                } else {
                    t1 = R.tmul(w);
                    wn = vadd(t1, qdv);
                    t3 = R.tmul(wd);
                    t2 = vcross(t1, qdv);
                    t1 = vadd(t2, t3);
                    wdn = vadd(t1, qddv);
                    t1 = vcross(w, pstar);
                    t2 = vcross(w, t1);
                    t1 = vcross(wd, pstar);
                    t1 = vadd(t1, t2);
                    t1 = vadd(t1, acc);
                }
                accn = R.tmul(t1);
            } else {
                if (j == 0) {
                    wn = qdv; wdn = qddv; accn = gravity; // sic, ne.c:187-204
                } else {
                    wn = R.tmul(w);
                    wdn = R.tmul(wd);
                    t1 = vcross(wd, pstar);
                    t3 = vcross(w, pstar);
                    t2 = vcross(w, t3);
                    t1 = vadd(t1, t2);
                    t1 = vadd(t1, acc);
                    accn = R.tmul(t1);
                    t2 = R.tmul(w);
                    t1 = vcross(t2, qdv);
                    t1 = vscale(t1, (real)2);
                    accn = vadd(accn, t1);
                    accn = vadd(accn, qddv);
                }
            }
        } else {
            if (!pris) {
                t1 = (j == 0) ? qdv : vadd(w, qdv);
                wn = R.tmul(t1);
                if (j == 0) t3 = qddv;
                else {
                    t1 = vadd(wd, qddv);
                    t2 = vcross(w, qdv);
                    t3 = vadd(t1, t2);
                }
                wdn = R.tmul(t3);
                t1 = vcross(wdn, pstar);
                t2 = vcross(wn, pstar);
                t3 = vcross(wn, t2);
                accn = vadd(t1, t3);
                t1 = R.tmul(j == 0 ? gravity : acc);
                accn = vadd(accn, t1);
            } else {
                if (j == 0) {
                    wn = {0, 0, 0}; wdn = {0, 0, 0};
                    t1 = vadd(qddv, gravity);
                    accn = R.tmul(t1);
                } else {
                    wn = R.tmul(w);
                    wdn = R.tmul(wd);
                    t1 = vadd(qddv, acc);
                    accn = R.tmul(t1);
                }
                t1 = vcross(wdn, pstar);
                accn = vadd(accn, t1);
                t1 = R.tmul(qdv);
                t2 = vcross(wn, t1);
                t2 = vscale(t2, (real)2);
                accn = vadd(accn, t2);
                t2 = vcross(wn, pstar);
                t3 = vcross(wn, t2);
                accn = vadd(accn, t3);
            }
        }
        w = wn; wd = wdn; acc = accn;
        // centre-of-mass acceleration, ne.c:228-232 / 335-339, then the link wrench terms
        const V3<real> rc = {P.r[j][0], P.r[j][1], P.r[j][2]};
        t1 = vcross(wd, rc);
        t2 = vcross(w, rc);
        t3 = vcross(w, t2);
        V3<real> abar = vadd(t1, t3);
        abar = vadd(abar, acc);
        Fm[j] = vscale(abar, P.m[j]);
        const real *I = P.I[j]; // read column-major like vmath.c mat_vect_mult
        t2 = {I[0] * wd.x + I[3] * wd.y + I[6] * wd.z, I[1] * wd.x + I[4] * wd.y + I[7] * wd.z,
              I[2] * wd.x + I[5] * wd.y + I[8] * wd.z};
        t3 = {I[0] * w.x + I[3] * w.y + I[6] * w.z, I[1] * w.x + I[4] * w.y + I[7] * w.z,
              I[2] * w.x + I[5] * w.y + I[8] * w.z};
        Nm[j] = vadd(t2, vcross(w, t3));
    }

    // ---------------- backward recursion (ne.c:358-403 MDH, 409-457 DH) + joint torque (ne.c:464-491)
    V3<real> f = {0, 0, 0}, nn = {0, 0, 0}; // of link j+1 on entry
    const V3<real> f_tip = {P.fext[0], P.fext[1], P.fext[2]}, n_tip = {P.fext[3], P.fext[4], P.fext[5]};
#pragma unroll
    for (int j = N - 1; j >= 0; j--) {
        const bool pris = (P.prismatic[j] != 0);
        const V3<real> rc = {P.r[j][0], P.r[j][1], P.r[j][2]};
        V3<real> fj, nj, t1, t2, t3, t4;
        if (MDH) {
            const V3<real> F = Fm[j];
            if (j == N - 1) {
                fj = vadd(f_tip, F);
                t1 = n_tip;
            } else {
                const LinkRot<real, MDH> Rn(sth[j + 1], cth[j + 1], P.sa[j + 1], P.ca[j + 1]);
                const real dn = P.prismatic[j + 1] ? (mq[j + 1] + P.offset[j + 1]) : P.D[j + 1];
                const V3<real> pstar_n = {P.A[j + 1], -dn * P.sa[j + 1], dn * P.ca[j + 1]};
                t1 = Rn.mul(f);
                fj = vadd(t1, F);
                t1 = Rn.mul(nn);
                t4 = Rn.mul(f);
                t3 = vcross(pstar_n, t4);
                t1 = vadd(t1, t3);
            }
            t2 = vcross(rc, F);
            t1 = vadd(t1, t2);
            nj = vadd(t1, Nm[j]);
        } else {
            const real dj = pris ? (mq[j] + P.offset[j]) : P.D[j];
            const V3<real> pstar = {P.A[j], dj * P.sa[j], dj * P.ca[j]};
            t4 = Fm[j];
            t2 = vadd(pstar, rc);
            t1 = vcross(t2, t4);
            if (j != N - 1) {
                const LinkRot<real, MDH> Rn(sth[j + 1], cth[j + 1], P.sa[j + 1], P.ca[j + 1]);
                fj = vadd(t4, Rn.mul(f));
                t2 = Rn.tmul(pstar);
                t3 = vcross(t2, f);
                t3 = vadd(t3, nn);
                t2 = Rn.mul(t3);
                t1 = vadd(t1, t2);
            } else {
                fj = vadd(t4, f_tip);
                t2 = vcross(pstar, f_tip);
                t1 = vadd(t1, t2);
                t1 = vadd(t1, n_tip);
            }
            nj = vadd(t1, Nm[j]);
        }
        f = fj; nn = nj;
        // torque about / force along the joint axis
        V3<real> zax;
        if (MDH) zax = {0, 0, 1};
        else zax = {0, P.sa[j], P.ca[j]}; // R_j^T z0 = (n.z, o.z, a.z)
        real t = pris ? vdot(fj, zax) : vdot(nj, zax);
        const real qdj = mqd[j];
        t += P.c_jm[j] * mqdd[j];
        t += P.c_b[j] * qdj;
        t += (qdj > 0 ? P.c_tcp[j] : (real)0) + (qdj < 0 ? P.c_tcm[j] : (real)0);
        tq[j] = t;
    }
    // ---------------- stage tau and write it back coalesced

}

// ALLREV: every joint is revolute (the usual case) -- strips the prismatic code, which the compiler
// would otherwise if-convert into always-executed select chains.
template <typename real, int N, bool MDH, bool ALLREV>
__global__ void __launch_bounds__(B2K_THREADS, (sizeof(real) == 4 || (ALLREV && N <= 6)) ? 4 : 3)
k_rne(const __grid_constant__ RneP<real, N> P, const real *__restrict__ q, const real *__restrict__ qd,
      const real *__restrict__ qdd, long long nrows, real *__restrict__ tau, int warp_smem_bytes, int in_bytes,
      int qmode)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *wbase = smem_raw + (size_t)warp * warp_smem_bytes;
    real *sq = reinterpret_cast<real *>(wbase);
    real *sqd = reinterpret_cast<real *>(wbase + in_bytes);
    real *sqdd = reinterpret_cast<real *>(wbase + 2 * in_bytes);
    unsigned char *sout = wbase + 3 * in_bytes;
    const int lds = qmode ? N : (N | 1); // smem row stride of the input tiles
    const long long ntiles = (nrows + 31) >> 5;
    const long long tstride = (long long)gridDim.x * B2K_WARPS_PER_BLOCK;
    const float inv_n = 1.0f / (float)N;

    auto load_inputs = [&](long long t) {
        const long long r0 = t << 5;
        const int rh = (int)((nrows - r0) < 32 ? (nrows - r0) : 32);
        load_q_tile<real>(sq, q + r0 * N, rh, N, inv_n, qmode, lane);
        load_q_tile<real>(sqd, qd + r0 * N, rh, N, inv_n, qmode, lane);
        load_q_tile<real>(sqdd, qdd + r0 * N, rh, N, inv_n, qmode, lane);
    };
    long long tile = (long long)blockIdx.x * B2K_WARPS_PER_BLOCK + warp;
    if (tile < ntiles) load_inputs(tile);
    for (; tile < ntiles; tile += tstride) {
        const long long row0 = tile << 5;
        const int rows_here = (int)((nrows - row0) < 32 ? (nrows - row0) : 32);
        cp_async_wait_all();
        __syncwarp();
        const int myrow = (lane < rows_here ? lane : 0) * lds;
        const real *mq = sq + myrow, *mqd = sqd + myrow, *mqdd = sqdd + myrow;

        real tq[N], sth[N], cth[N];
        rne_sincos<real, N, ALLREV>(P, mq, sth, cth);
        const V3<real> grav = {P.grav[0], P.grav[1], P.grav[2]};
        if constexpr (ALLREV) rne_row_allrev<real, N, MDH>(P, sth, cth, mqd, mqdd, grav, tq);
        else rne_row_generic<real, N, MDH>(P, mq, sth, cth, mqd, mqdd, grav, tq);
        // the inputs of this tile are dead (tq holds the results): prefetch the next tile behind the drain
        __syncwarp();
        if (tile + tstride < ntiles) load_inputs(tile + tstride);
        TileStage<real, N>::put_row(sout, lane, tq);
        __syncwarp();
        TileStage<real, N>::drain(sout, tau + row0 * N, rows_here, lane);
        __syncwarp();
    }
    cp_async_wait_all();
}

// ------------------------------------------------------------------ dynamics fan-outs of the recursion
// The reference's DynamicsMixin builds everything from repeated rne calls in Python loops
// (Dynamics.py): inertia = n calls with unit accelerations (752-758), gravload = one call with
// qd = qdd = 0 (912-915), itorque = one call without gravity (1456-1459), coriolis = n + n(n-1)/2
// calls on a friction-free copy (825-857), accel = inertia + one call + an n x n solve (490-503).
// Here one lane does all the calls of its row with the link rotations (sincos) computed once;
// the loops over unit vectors are runtime loops (code size of ONE recursion), the small n x n
// accumulators live in local memory.
enum { FAN_INERTIA = 0, FAN_GRAVLOAD = 1, FAN_ITORQUE = 2, FAN_CORIOLIS = 3, FAN_ACCEL = 4 };

template <int MODE, int N>
struct FanShape {
    static constexpr int NIN = (MODE == FAN_INERTIA || MODE == FAN_GRAVLOAD) ? 1 : (MODE == FAN_ACCEL ? 3 : 2);
    static constexpr int OUT = (MODE == FAN_INERTIA || MODE == FAN_CORIOLIS) ? N * N : N;
};

template <typename real, int N, bool MDH, bool ALLREV, int MODE>
__global__ void __launch_bounds__(B2K_THREADS, (sizeof(real) == 4) ? 3 : 2)
k_rne_fan(const __grid_constant__ RneP<real, N> P, const real *__restrict__ in0, const real *__restrict__ in1,
          const real *__restrict__ in2, long long nrows, real *__restrict__ out, int warp_smem_bytes, int in_bytes,
          int qmode)
{
    typedef FanShape<MODE, N> SH;
    typedef TileStage<real, SH::OUT> OS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *wbase = smem_raw + (size_t)warp * warp_smem_bytes;
    real *s0 = reinterpret_cast<real *>(wbase);
    real *s1 = reinterpret_cast<real *>(wbase + in_bytes);
    real *s2 = reinterpret_cast<real *>(wbase + 2 * in_bytes);
    unsigned char *sout = wbase + SH::NIN * in_bytes;
    const int lds = qmode ? N : (N | 1);
    const long long ntiles = (nrows + 31) >> 5;
    const float inv_n = 1.0f / (float)N;
    const long long tile = (long long)blockIdx.x * B2K_WARPS_PER_BLOCK + warp; // one tile per warp, one-shot grid
    if (tile >= ntiles) return;
    const long long row0 = tile << 5;
    const int rows_here = (int)((nrows - row0) < 32 ? (nrows - row0) : 32);
    load_q_tile<real>(s0, in0 + row0 * N, rows_here, N, inv_n, qmode, lane);
    if (SH::NIN >= 2) load_q_tile<real>(s1, in1 + row0 * N, rows_here, N, inv_n, qmode, lane);
    if (SH::NIN >= 3) load_q_tile<real>(s2, in2 + row0 * N, rows_here, N, inv_n, qmode, lane);
    cp_async_wait_all();
    __syncwarp();
    const int myrow = (lane < rows_here ? lane : 0) * lds;
    const real *mq = s0 + myrow;

    real sth[N], cth[N];
    rne_sincos<real, N, ALLREV>(P, mq, sth, cth);
    const V3<real> gvec = {P.grav[0], P.grav[1], P.grav[2]}, gzero = {0, 0, 0};
    auto rne_eval_g = [&](const real *qdv, const real *qddv, V3<real> g, real *tq) {
        if constexpr (ALLREV) rne_row_allrev<real, N, MDH>(P, sth, cth, qdv, qddv, g, tq);
        else rne_row_generic<real, N, MDH>(P, mq, sth, cth, qdv, qddv, g, tq);
    };
    auto rne_eval = [&](const real *qdv, const real *qddv, real *tq) { rne_eval_g(qdv, qddv, gvec, tq); };
    real res[SH::OUT];
    real zero[N];
#pragma unroll
    for (int k = 0; k < N; k++) zero[k] = 0;

    if constexpr (MODE == FAN_GRAVLOAD) {
        rne_eval(zero, zero, res);
    } else if constexpr (MODE == FAN_ITORQUE) {
        real a[N];
#pragma unroll
        for (int k = 0; k < N; k++) a[k] = s1[myrow + k];
        rne_eval(zero, a, res);
    } else if constexpr (MODE == FAN_INERTIA) {
#pragma unroll 1
        for (int i = 0; i < N; i++) {
            real a[N], tq[N];
#pragma unroll
            for (int k = 0; k < N; k++) a[k] = (k == i) ? (real)1 : (real)0;
            rne_eval(zero, a, tq);
#pragma unroll
            for (int k = 0; k < N; k++) res[i * N + k] = tq[k]; // row i = torque for a unit acceleration of joint i
        }
    } else if constexpr (MODE == FAN_CORIOLIS) {
        real Csq[N * N];
        real qd[N];
#pragma unroll
        for (int k = 0; k < N; k++) qd[k] = s1[myrow + k];
#pragma unroll
        for (int k = 0; k < N * N; k++) res[k] = 0;
#pragma unroll 1
        for (int i = 0; i < N; i++) { // centripetal terms: one joint moving at unit speed (Dynamics.py:828-833)
            real v[N], tq[N];
#pragma unroll
            for (int k = 0; k < N; k++) v[k] = (k == i) ? (real)1 : (real)0;
            rne_eval(v, zero, tq);
#pragma unroll
            for (int k = 0; k < N; k++) Csq[k * N + i] = tq[k];
        }
#pragma unroll 1
        for (int i = 0; i < N; i++) { // Coriolis terms: pairs of joints at unit speed (Dynamics.py:839-855)
#pragma unroll 1
            for (int j = i + 1; j < N; j++) {
                real v[N], tq[N];
#pragma unroll
                for (int k = 0; k < N; k++) v[k] = (k == i || k == j) ? (real)1 : (real)0;
                rne_eval(v, zero, tq);
                real qdi = 0, qdj = 0;
#pragma unroll
                for (int k = 0; k < N; k++) { qdi = (k == i) ? qd[k] : qdi; qdj = (k == j) ? qd[k] : qdj; }
#pragma unroll
                for (int k = 0; k < N; k++) {
                    const real t = tq[k] - Csq[k * N + j] - Csq[k * N + i];
                    res[k * N + j] = res[k * N + j] + t * qdi / (real)2;
                    res[k * N + i] = res[k * N + i] + t * qdj / (real)2;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < N; k++)
#pragma unroll
            for (int i = 0; i < N; i++) res[k * N + i] = res[k * N + i] + Csq[k * N + i] * qd[i]; // + Csq @ diag(qd)
    } else { // FAN_ACCEL: qdd = M^-1 (tau - rne(q, qd, 0))  (Dynamics.py:490-503, Walker & Orin method 1)
        real M[N * N], rhs[N];
        {
            real v[N], tq[N];
#pragma unroll
            for (int k = 0; k < N; k++) v[k] = s1[myrow + k];
            rne_eval(v, zero, tq); // gravity + Coriolis + friction torque at zero acceleration
#pragma unroll
            for (int k = 0; k < N; k++) rhs[k] = s2[myrow + k] - tq[k];
        }
#pragma unroll 1
        for (int i = 0; i < N; i++) { // inertia rows: unit accelerations, no velocity, no gravity (Dynamics.py:492-496)
            real a[N], tq[N];
#pragma unroll
            for (int k = 0; k < N; k++) a[k] = (k == i) ? (real)1 : (real)0;
            rne_eval_g(zero, a, gzero, tq);
#pragma unroll
            for (int k = 0; k < N; k++) M[i * N + k] = tq[k];
        }
        // Gaussian elimination with partial pivoting (numpy.linalg.solve = LAPACK gesv)
#pragma unroll 1
        for (int c = 0; c < N; c++) {
            int p = c;
            real best = fabs(M[c * N + c]);
#pragma unroll 1
            for (int r = c + 1; r < N; r++) {
                const real v = fabs(M[r * N + c]);
                if (v > best) { best = v; p = r; }
            }
            if (p != c) {
#pragma unroll 1
                for (int k = 0; k < N; k++) { const real t = M[c * N + k]; M[c * N + k] = M[p * N + k]; M[p * N + k] = t; }
                const real t = rhs[c]; rhs[c] = rhs[p]; rhs[p] = t;
            }
            const real inv = (real)1 / M[c * N + c];
#pragma unroll 1
            for (int r = c + 1; r < N; r++) {
                const real f = M[r * N + c] * inv;
#pragma unroll 1
                for (int k = c; k < N; k++) M[r * N + k] -= f * M[c * N + k];
                rhs[r] -= f * rhs[c];
            }
        }
#pragma unroll 1
        for (int r = N - 1; r >= 0; r--) {
            real acc = rhs[r];
#pragma unroll 1
            for (int k = r + 1; k < N; k++) acc -= M[r * N + k] * res[k];
            res[r] = acc / M[r * N + r];
        }
    }
    OS::put_row(sout, lane, res);
    if (!OS::drain_async(sout, out + row0 * SH::OUT, rows_here, lane)) {
        __syncwarp();
        OS::drain(sout, out + row0 * SH::OUT, rows_here, lane);
    }
    OS::wait_all();
}

template <typename real, int N>
void rne_fill_params(const b2k_rne_s *r, const double *grav, const double *fext, RneP<real, N> &P)
{
    for (int j = 0; j < N; j++) {
        const double *l = r->L[j];
        const double alpha = l[0], A = l[1], theta = l[2], D = l[3], offset = l[5];
        const double G = l[20];
        P.sa[j] = (real)sin(alpha);
        P.ca[j] = (real)cos(alpha);
        P.A[j] = (real)A;
        P.D[j] = (real)D;
        P.st0[j] = (real)sin(theta);
        P.ct0[j] = (real)cos(theta);
        P.offset[j] = (real)offset;
        P.prismatic[j] = ((int)l[4]) != 0;
        P.m[j] = (real)l[6];
        for (int k = 0; k < 3; k++) P.r[j][k] = (real)l[7 + k];
        for (int k = 0; k < 9; k++) P.I[j][k] = (real)l[10 + k];
        P.c_jm[j] = (real)(G * G * l[19]);
        P.c_b[j] = (real)(G * G * l[21]);
        P.c_tcp[j] = (real)(fabs(G) * l[22]);
        P.c_tcm[j] = (real)(fabs(G) * l[23]);
        P.ps[j][0] = (real)A;
        P.ps[j][1] = (real)(r->mdh ? -D * sin(alpha) : D * sin(alpha));
        P.ps[j][2] = (real)(D * cos(alpha));
        for (int k = 0; k < 3; k++) P.psrc[j][k] = (real)((double)P.ps[j][k] + (double)P.r[j][k]);
    }
    for (int k = 0; k < 3; k++) P.grav[k] = grav ? (real)grav[k] : (real)0;
    for (int k = 0; k < 6; k++) P.fext[k] = fext ? (real)fext[k] : (real)0;
    b2k_fill_trig<real>(P.trig);
}

template <typename real, int N>
int rne_launch_n(const b2k_rne_s *r, const real *q, const real *qd, const real *qdd, long long nrows,
                 const double *grav, const double *fext, real *tau, cudaStream_t st)
{
    RneP<real, N> P;
    rne_fill_params<real, N>(r, grav, fext, P);
    const size_t inb = fkj_q_bytes<real>(N);
    const size_t wsm = (3 * inb + (size_t)TileStage<real, N>::BYTES + 15) & ~(size_t)15;
    const size_t smem = wsm * B2K_WARPS_PER_BLOCK;
    const int qmode = (fkj_qmode<real>(q, N) && fkj_qmode<real>(qd, N) && fkj_qmode<real>(qdd, N)) ? 1 : 0;
    if (((uintptr_t)tau) % TileStage<real, N>::UB) { b2k_set_error("rne: tau must be %d-byte aligned", TileStage<real, N>::UB); return B2K_ERR_INVALID; }
    const long long ntiles = (nrows + 31) / 32;
    const long long nblk_needed = (ntiles + B2K_WARPS_PER_BLOCK - 1) / B2K_WARPS_PER_BLOCK;
    auto launch = [&](auto kern) -> int {
        int per_sm = b2k_blocks_per_sm((const void *)kern, B2K_THREADS, smem);
        if (per_sm < 1) return per_sm < 0 ? per_sm : (b2k_set_error("rne kernel does not fit on an SM"), B2K_ERR_INVALID);
        long long grid = nblk_needed; // one tile per warp, one-shot grid (see b2k_fkj.cuh launcher)
        if (b2k_get_variant() == 4) {
            grid = (long long)b2k_num_sms() * per_sm;
            if (grid > nblk_needed) grid = nblk_needed;
        }
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, B2K_THREADS, smem, st>>>(P, q, qd, qdd, nrows, tau, (int)wsm, (int)inb, qmode);
        b2k_count_launch();
        B2K_CUDA(cudaGetLastError());
        return B2K_OK;
    };
    bool allrev = true;
    for (int j = 0; j < N; j++) allrev = allrev && !P.prismatic[j];
    if (r->mdh) return allrev ? launch(k_rne<real, N, true, true>) : launch(k_rne<real, N, true, false>);
    return allrev ? launch(k_rne<real, N, false, true>) : launch(k_rne<real, N, false, false>);
}

template <typename real>
int rne_launch(const b2k_rne_s *r, const void *q, const void *qd, const void *qdd, long long nrows, const double *grav,
               const double *fext, void *tau, cudaStream_t st)
{
#define B2K_CASE(NN) \
    case NN: return rne_launch_n<real, NN>(r, (const real *)q, (const real *)qd, (const real *)qdd, nrows, grav, fext, (real *)tau, st);
    switch (r->n) {
        B2K_CASE(1) B2K_CASE(2) B2K_CASE(3) B2K_CASE(4) B2K_CASE(5)
        B2K_CASE(6) B2K_CASE(7) B2K_CASE(8) B2K_CASE(9) B2K_CASE(10)
    default:
        b2k_set_error("rne: unsupported joint count %d", r->n);
        return B2K_ERR_INVALID;
    }
#undef B2K_CASE
}

// ------------------------------------------------------------------ fan-out launcher
template <typename real, int N, int MODE>
int rne_fan_launch_mode(const b2k_rne_s *r, const real *in0, const real *in1, const real *in2, long long nrows,
                        const double *grav, real *out, cudaStream_t st)
{
    typedef FanShape<MODE, N> SH;
    RneP<real, N> P;
    rne_fill_params<real, N>(r, (MODE == FAN_GRAVLOAD || MODE == FAN_ACCEL) ? grav : nullptr, nullptr, P);
    if (MODE == FAN_CORIOLIS) { // the reference works on robot.nofriction(True, True) (Dynamics.py:818)
        for (int j = 0; j < N; j++) { P.c_b[j] = 0; P.c_tcp[j] = 0; P.c_tcm[j] = 0; }
    }
    const size_t inb = fkj_q_bytes<real>(N);
    const size_t wsm = (SH::NIN * inb + (size_t)TileStage<real, SH::OUT>::BYTES + 15) & ~(size_t)15;
    const size_t smem = wsm * B2K_WARPS_PER_BLOCK;
    int qmode = fkj_qmode<real>(in0, N);
    if (SH::NIN >= 2) qmode = qmode && fkj_qmode<real>(in1, N);
    if (SH::NIN >= 3) qmode = qmode && fkj_qmode<real>(in2, N);
    if (((uintptr_t)out) % TileStage<real, SH::OUT>::UB) { b2k_set_error("rne fan-out: output must be %d-byte aligned", TileStage<real, SH::OUT>::UB); return B2K_ERR_INVALID; }
    const long long ntiles = (nrows + 31) / 32;
    const long long grid = (ntiles + B2K_WARPS_PER_BLOCK - 1) / B2K_WARPS_PER_BLOCK;
    bool allrev = true;
    for (int j = 0; j < N; j++) allrev = allrev && !P.prismatic[j];
    auto launch = [&](auto kern) -> int {
        int per_sm = b2k_blocks_per_sm((const void *)kern, B2K_THREADS, smem);
        if (per_sm < 1) return per_sm < 0 ? per_sm : (b2k_set_error("rne fan-out kernel does not fit on an SM"), B2K_ERR_INVALID);
        kern<<<(unsigned)grid, B2K_THREADS, smem, st>>>(P, in0, in1, in2, nrows, out, (int)wsm, (int)inb, qmode);
        b2k_count_launch();
        B2K_CUDA(cudaGetLastError());
        return B2K_OK;
    };
    if (r->mdh) return allrev ? launch(k_rne_fan<real, N, true, true, MODE>) : launch(k_rne_fan<real, N, true, false, MODE>);
    return allrev ? launch(k_rne_fan<real, N, false, true, MODE>) : launch(k_rne_fan<real, N, false, false, MODE>);
}

template <typename real, int N>
int rne_fan_launch_n(const b2k_rne_s *r, int mode, const real *in0, const real *in1, const real *in2, long long nrows,
                     const double *grav, real *out, cudaStream_t st)
{
    switch (mode) {
    case FAN_INERTIA: return rne_fan_launch_mode<real, N, FAN_INERTIA>(r, in0, in1, in2, nrows, grav, out, st);
    case FAN_GRAVLOAD: return rne_fan_launch_mode<real, N, FAN_GRAVLOAD>(r, in0, in1, in2, nrows, grav, out, st);
    case FAN_ITORQUE: return rne_fan_launch_mode<real, N, FAN_ITORQUE>(r, in0, in1, in2, nrows, grav, out, st);
    case FAN_CORIOLIS: return rne_fan_launch_mode<real, N, FAN_CORIOLIS>(r, in0, in1, in2, nrows, grav, out, st);
    case FAN_ACCEL: return rne_fan_launch_mode<real, N, FAN_ACCEL>(r, in0, in1, in2, nrows, grav, out, st);
    }
    b2k_set_error("rne fan-out: bad mode %d", mode);
    return B2K_ERR_INVALID;
}

template <typename real>
int rne_fan_launch(const b2k_rne_s *r, int mode, const void *in0, const void *in1, const void *in2, long long nrows,
                   const double *grav, void *out, cudaStream_t st)
{
#define B2K_CASE(NN) \
    case NN: return rne_fan_launch_n<real, NN>(r, mode, (const real *)in0, (const real *)in1, (const real *)in2, nrows, grav, (real *)out, st);
    switch (r->n) {
        B2K_CASE(1) B2K_CASE(2) B2K_CASE(3) B2K_CASE(4) B2K_CASE(5)
        B2K_CASE(6) B2K_CASE(7) B2K_CASE(8) B2K_CASE(9) B2K_CASE(10)
    default:
        b2k_set_error("rne fan-out: unsupported joint count %d", r->n);
        return B2K_ERR_INVALID;
    }
#undef B2K_CASE
}
