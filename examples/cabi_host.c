/* cabi_host.c -- the drop-in boundary used from plain C: no Python, no torch, no CUDA headers.
 *
 *   gcc -O2 -I include examples/cabi_host.c -L robotics-toolbox-python_b200/lib -lb2kin \
 *       -Wl,-rpath,$PWD/robotics-toolbox-python_b200/lib -lm -o examples/cabi_host
 *   ./examples/cabi_host [N] > out.txt
 *
 * Builds a standard-DH 3R arm as an elementary-transform chain (what DHLink._to_ets emits, reference
 * DHLink.py:204-223: Rz(q) tz(d) tx(a) Rx(alpha) per link), evaluates pose + base-frame Jacobian for N
 * configurations through the host-buffer entry point and prints every row as text:
 *   q0 q1 q2 | 16 pose values | 18 Jacobian values
 * tests/test_gpu_parity.py::test_c_program_through_the_c_abi compiles it, runs it and checks the rows
 * against the oracle.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2kin.h"

static void ident(double *T) { memset(T, 0, 16 * sizeof(double)); T[0] = T[5] = T[10] = T[15] = 1.0; }

int main(int argc, char **argv)
{
    long N = argc > 1 ? atol(argv[1]) : 8;
    /* links: (d, a, alpha) */
    const double dh[3][3] = {{0.4, 0.0, 1.5707963267948966}, {0.0, 0.35, 0.0}, {0.1, 0.25, -1.5707963267948966}};
    enum { M = 12 };
    int32_t isjoint[M], axis[M], flip[M], jindex[M];
    double T[M * 16], qlim[M * 2];
    int m = 0;
    for (int l = 0; l < 3; l++) {
        /* joint Rz(q_l) */
        isjoint[m] = 1; axis[m] = B2K_RZ; flip[m] = 0; jindex[m] = l; ident(T + 16 * m);
        qlim[2 * m] = -3.141592653589793; qlim[2 * m + 1] = 3.141592653589793; m++;
        /* constants tz(d), tx(a), Rx(alpha) */
        for (int k = 0; k < 3; k++) {
            isjoint[m] = 0; axis[m] = 0; flip[m] = 0; jindex[m] = 0; qlim[2 * m] = qlim[2 * m + 1] = 0.0;
            double *E = T + 16 * m;
            ident(E);
            if (k == 0) E[11] = dh[l][0];
            else if (k == 1) E[3] = dh[l][1];
            else { double c = cos(dh[l][2]), s = sin(dh[l][2]); E[5] = c; E[6] = -s; E[9] = s; E[10] = c; }
            m++;
        }
    }
    b2k_chain_t chain = NULL;
    if (b2k_chain_create(m, isjoint, axis, flip, jindex, T, qlim, &chain) != B2K_OK) {
        fprintf(stderr, "chain_create: %s\n", b2k_last_error());
        return 1;
    }
    int n, mm, w;
    b2k_chain_info(chain, &n, &mm, &w);
    double *q = malloc(sizeof(double) * N * 3), *Tout = malloc(sizeof(double) * N * 16), *J = malloc(sizeof(double) * N * 18);
    unsigned long long st = 88172645463325252ULL; /* xorshift64: deterministic inputs in [-pi, pi) */
    for (long i = 0; i < N * 3; i++) {
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        q[i] = ((double)(st >> 11) / 9007199254740992.0 * 2.0 - 1.0) * 3.141592653589793;
    }
    int rc = b2k_fkine_jacob0_host(chain, B2K_F64, q, N, 3, NULL, NULL, Tout, J, 0);
    if (rc != B2K_OK) {
        fprintf(stderr, "fkine_jacob0_host: %s\n", b2k_last_error());
        return 2;
    }
    fprintf(stderr, "b2kin %d: n=%d m=%d rows=%ld launches=%lld\n", b2k_version(), n, mm, N, (long long)b2k_launch_count());
    for (long r = 0; r < N; r++) {
        for (int k = 0; k < 3; k++) printf("%.17g ", q[r * 3 + k]);
        for (int k = 0; k < 16; k++) printf("%.17g ", Tout[r * 16 + k]);
        for (int k = 0; k < 18; k++) printf("%.17g%c", J[r * 18 + k], k == 17 ? '\n' : ' ');
    }
    b2k_chain_destroy(chain);
    free(q); free(Tout); free(J);
    return 0;
}
