/* abi_consumer.c -- a plain-C consumer of include/rchem_eri.h (no Python, no torch).
 *
 * Compiled by tests/test_gpu_kernels.py::test_plain_c_consumer_of_the_abi with
 *   gcc -std=c99 -Wall -Werror -I include tests/abi_consumer.c -L rchem_b200 -lrchem_b200
 * and run on the GPU box.  It walks the calls a reference-side binding makes for the hot path
 * (basis.rs: Basis::new -> JK_direct -> build_I -> JK_inmem) on water_crawford / STO-3G and
 * prints key=value checksums the test compares with the oracle.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rchem_eri.h"

#define CHECK(call)                                                            \
  do {                                                                         \
    int rc_ = (call);                                                          \
    if (rc_ < 0) {                                                             \
      fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, rchem_last_error()); \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int main(void) {
  /* water_crawford.xyz of the reference (bohr; used raw like rchem.rs:29-35) */
  const uint64_t atomnos[3] = {8, 1, 1};
  const double coords[9] = {0.000000000000,  -0.143225816552, 0.000000000000,
                            1.638036840407,  1.136548822547,  -0.000000000000,
                            -1.638036840407, 1.136548822547,  -0.000000000000};
  rchem_basis* basis = NULL;
  int n, i, j;
  size_t n2, n4, t;
  double *D, *J, *K, *I, *J2, *K2;
  double sumI = 0.0, sumJ = 0.0, sumK = 0.0, trJ = 0.0, trK = 0.0, dJ = 0.0, dK = 0.0;
  rchem_stats st;

  if (rchem_device_count() < 1) {
    fprintf(stderr, "no CUDA device: %s\n", rchem_last_error());
    return 2;
  }
  CHECK(rchem_basis_new(3, atomnos, coords, "STO-3G", &basis));
  n = rchem_basis_nbf(basis);
  n2 = (size_t)n * n;
  n4 = n2 * n2;
  D = malloc(n2 * sizeof(double));
  J = malloc(n2 * sizeof(double));
  K = malloc(n2 * sizeof(double));
  J2 = malloc(n2 * sizeof(double));
  K2 = malloc(n2 * sizeof(double));
  I = malloc(n4 * sizeof(double));
  if (!D || !J || !K || !J2 || !K2 || !I) return 3;
  for (i = 0; i < n; ++i)
    for (j = 0; j < n; ++j) {
      D[(size_t)i * n + j] = 0.1 / (1.0 + i + j);
      J[(size_t)i * n + j] = K[(size_t)i * n + j] = 1e30; /* must be overwritten */
    }

  CHECK(rchem_jk_direct(basis, D, J, K));
  CHECK(rchem_get_stats(basis, &st));
  CHECK(rchem_build_I(basis, I));
  CHECK(rchem_jk_inmem(n, I, D, J2, K2));

  for (t = 0; t < n4; ++t) sumI += I[t];
  for (t = 0; t < n2; ++t) {
    sumJ += J[t];
    sumK += K[t];
    if (fabs(J[t] - J2[t]) > dJ) dJ = fabs(J[t] - J2[t]);
    if (fabs(K[t] - K2[t]) > dK) dK = fabs(K[t] - K2[t]);
  }
  for (i = 0; i < n; ++i) {
    trJ += J[(size_t)i * n + i];
    trK += K[(size_t)i * n + i];
  }
  printf("nbf=%d\nquartets=%lld\nsumI=%.17g\nI0000=%.17g\nsumJ=%.17g\nsumK=%.17g\ntrJ=%.17g\n"
         "trK=%.17g\ninmem_dJ=%.3e\ninmem_dK=%.3e\n",
         n, (long long)st.shell_quartets, sumI, I[0], sumJ, sumK, trJ, trK, dJ, dK);
  rchem_basis_destroy(basis);
  free(D); free(J); free(K); free(J2); free(K2); free(I);
  return 0;
}
