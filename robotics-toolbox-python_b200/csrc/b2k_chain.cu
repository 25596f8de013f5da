// b2k_chain.cu -- host-side chain compiler and handle management.
//
// Replaces fknm.ET_init / ETS_init (reference fknm.cpp:1182-1239, 1066-1114).  The list of
// elementary transforms is compiled once into "n x (SE(3) constant, joint) + tail constant":
// runs of constant ETs are multiplied together in fp64 on the host -- the reference's own
// ETS.compile() rule (ETS.py:857-906) -- so the device walk does one structured 3x4 product
// per joint instead of one full 4x4 product per elementary transform (Panda: 22 -> 8).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2k_common.cuh"

void b2k_ident34(double *A)
{
    for (int k = 0; k < 12; k++) A[k] = 0.0;
    A[0] = A[5] = A[10] = 1.0;
}

void b2k_mat_to34(const double *T16, double *A12)
{
    for (int k = 0; k < 12; k++) A12[k] = T16[k];
}

// C = A * B for affine 3x4 matrices (bottom row 0 0 0 1 implied)
void b2k_mul34(const double *A, const double *B, double *C)
{
    double t[12];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            t[i * 4 + j] = A[i * 4 + 0] * B[0 * 4 + j] + A[i * 4 + 1] * B[1 * 4 + j] + A[i * 4 + 2] * B[2 * 4 + j];
        t[i * 4 + 3] = A[i * 4 + 0] * B[3] + A[i * 4 + 1] * B[7] + A[i * 4 + 2] * B[11] + A[i * 4 + 3];
    }
    memcpy(C, t, sizeof(t));
}

int b2k_classify34(const double *A)
{
    auto is = [&](int r, int c, double v) { return A[r * 4 + c] == v; };
    int k;
    bool row0 = is(0, 0, 1) && is(0, 1, 0) && is(0, 2, 0) && is(1, 0, 0) && is(2, 0, 0);
    bool row1 = is(1, 1, 1) && is(1, 0, 0) && is(1, 2, 0) && is(0, 1, 0) && is(2, 1, 0);
    bool row2 = is(2, 2, 1) && is(2, 0, 0) && is(2, 1, 0) && is(0, 2, 0) && is(1, 2, 0);
    if (row0 && row1 && row2) k = AK_IDENT;
    else if (row0) k = AK_RX;
    else if (row1) k = AK_RY;
    else if (row2) k = AK_RZ;
    else k = AK_GEN;
    if (A[3] != 0.0) k |= AK_TX;
    if (A[7] != 0.0) k |= AK_TY;
    if (A[11] != 0.0) k |= AK_TZ;
    return k;
}

extern "C" int b2k_chain_create(int m, const int32_t *isjoint, const int32_t *axis, const int32_t *flip,
                                const int32_t *jindex, const double *T, const double *qlim, b2k_chain_t *out)
{
    if (!out) { b2k_set_error("b2k_chain_create: out is NULL"); return B2K_ERR_INVALID; }
    *out = nullptr;
    if (m < 0 || (m > 0 && (!isjoint || !axis || !flip || !jindex || !T || !qlim))) {
        b2k_set_error("b2k_chain_create: NULL array argument");
        return B2K_ERR_INVALID;
    }
    int n = 0;
    for (int i = 0; i < m; i++) n += isjoint[i] ? 1 : 0;
    if (n < 1 || n > B2K_MAX_JOINTS) {
        b2k_set_error("b2k_chain_create: chain has %d joints; supported 1..%d", n, B2K_MAX_JOINTS);
        return B2K_ERR_INVALID;
    }
    b2k_chain_s *c = (b2k_chain_s *)calloc(1, sizeof(b2k_chain_s));
    if (!c) { b2k_set_error("b2k_chain_create: out of memory"); return B2K_ERR_ALLOC; }
    c->n = n;
    c->m = m;
    double acc[12];
    b2k_ident34(acc);
    int j = 0, qw = 0;
    c->all_rz = 1;
    c->dense_jindex = 1;
    for (int i = 0; i < m; i++) {
        if (!isjoint[i]) {
            const double *Ti = T + 16 * i;
            if (Ti[12] != 0.0 || Ti[13] != 0.0 || Ti[14] != 0.0 || Ti[15] != 1.0) {
                b2k_set_error("b2k_chain_create: ET %d constant is not affine (bottom row must be 0 0 0 1)", i);
                free(c);
                return B2K_ERR_INVALID;
            }
            double a[12];
            b2k_mat_to34(Ti, a);
            b2k_mul34(acc, a, acc);
        } else {
            if (axis[i] < 0 || axis[i] > 5) {
                b2k_set_error("b2k_chain_create: ET %d has invalid axis code %d", i, axis[i]);
                free(c);
                return B2K_ERR_INVALID;
            }
            if (jindex[i] < 0 || jindex[i] >= B2K_MAX_QWIDTH) {
                b2k_set_error("b2k_chain_create: ET %d jindex %d outside 0..%d", i, jindex[i], B2K_MAX_QWIDTH - 1);
                free(c);
                return B2K_ERR_INVALID;
            }
            memcpy(c->A[j], acc, sizeof(acc));
            b2k_ident34(acc);
            c->axis[j] = axis[i];
            c->flip[j] = flip[i] ? 1 : 0;
            c->jidx[j] = jindex[i];
            c->qlim_l[j] = qlim[2 * i];
            c->qlim_h[j] = qlim[2 * i + 1];
            if (axis[i] != B2K_RZ || flip[i]) c->all_rz = 0;
            if (jindex[i] != j) c->dense_jindex = 0;
            if (jindex[i] + 1 > qw) qw = jindex[i] + 1;
            j++;
        }
    }
    c->distinct_jindex = 1;
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
            if (c->jidx[a] == c->jidx[b]) c->distinct_jindex = 0;
    memcpy(c->A[n], acc, sizeof(acc)); // tail constant
    c->dh_like = c->all_rz;
    for (int k = 1; k < n && c->dh_like; k++) {
        const int rk = b2k_classify34(c->A[k]) & AK_ROTMASK;
        if (rk != AK_IDENT && rk != AK_RX) c->dh_like = 0;
    }
    c->q_width = qw;
    *out = c;
    return B2K_OK;
}

extern "C" int b2k_chain_destroy(b2k_chain_t chain)
{
    free(chain);
    return B2K_OK;
}

extern "C" int b2k_chain_info(b2k_chain_t c, int *n, int *m, int *q_width)
{
    if (!c) { b2k_set_error("b2k_chain_info: NULL chain"); return B2K_ERR_INVALID; }
    if (n) *n = c->n;
    if (m) *m = c->m;
    if (q_width) *q_width = c->q_width;
    return B2K_OK;
}

extern "C" int b2k_rne_create(int n, int mdh, const double *L, b2k_rne_t *out)
{
    if (!out) { b2k_set_error("b2k_rne_create: out is NULL"); return B2K_ERR_INVALID; }
    *out = nullptr;
    if (n < 1 || n > B2K_MAX_JOINTS || !L) {
        b2k_set_error("b2k_rne_create: n=%d outside 1..%d or L is NULL", n, B2K_MAX_JOINTS);
        return B2K_ERR_INVALID;
    }
    if (mdh != 0 && mdh != 1) { b2k_set_error("b2k_rne_create: mdh must be 0 or 1"); return B2K_ERR_INVALID; }
    b2k_rne_s *r = (b2k_rne_s *)calloc(1, sizeof(b2k_rne_s));
    if (!r) { b2k_set_error("b2k_rne_create: out of memory"); return B2K_ERR_ALLOC; }
    r->n = n;
    r->mdh = mdh;
    for (int j = 0; j < n; j++) {
        memcpy(r->L[j], L + 24 * j, 24 * sizeof(double));
        int sigma = (int)r->L[j][4];
        if (sigma != 0 && sigma != 1) { // frne.c:203-205 raises on anything but R / P
            b2k_set_error("b2k_rne_create: link %d has invalid joint type %d (expecting 0 = R or 1 = P)", j, sigma);
            free(r);
            return B2K_ERR_INVALID;
        }
    }
    b2k_rne_spec_attach(r);
    *out = r;
    return B2K_OK;
}

extern "C" int b2k_rne_destroy(b2k_rne_t r)
{
    if (r) b2k_rne_spec_detach(r);
    free(r);
    return B2K_OK;
}
