// Host build of android_svo_b200/csrc/ldlt.cuh (the device qualifiers defined away): ldlt_factor_rcp<6> + ldlt_subst_rcp<6>, the
// pair the sparse alignment re-uses while its Hessian stays the same, against ldlt_solve_fixed<6, true> — bit for bit, on SPD and
// indefinite systems, ties in the pivot search, zero rows, the zero matrix, and a second right-hand side on the same factor.
// Driven by tests/test_abi_and_host.py::test_ldlt_split_bit_identical (g++ -ffp-contract=off, as the library's -fmad=false).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#define __device__
#define __forceinline__ inline
#include "ldlt.cuh"
static double rnd() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
int main()
{
  int bad = 0, total = 0;
  for (int trial = 0; trial < 50000; ++trial) {
    double B[36], A[36], b[6], x0[6], x1[6];
    for (int i = 0; i < 36; ++i) B[i] = rnd() * pow(10.0, (int)(rnd() * 4));
    const int kind = trial % 5;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
      double s = 0;
      if (kind <= 1) { for (int k = 0; k < 6; ++k) s += B[i * 6 + k] * B[j * 6 + k]; }      // SPD
      else s = B[i * 6 + j] + B[j * 6 + i];                                              // symmetric indefinite
      A[i * 6 + j] = s;
    }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) A[i * 6 + j] = A[j * 6 + i];
    if (kind == 3) { A[0] = A[14] = A[28]; A[7] = -A[21]; }                              // ties in |diagonal|
    if (kind == 4) { for (int j = 0; j < 6; ++j) { A[2 * 6 + j] = 0; A[j * 6 + 2] = 0; } }   // zero row/column
    if (trial == 7) memset(A, 0, sizeof A);
    for (int i = 0; i < 6; ++i) b[i] = rnd();
    ldlt_solve_fixed<6, true>(A, b, x0);
    LdltFactor<6> F;
    ldlt_factor_rcp<6>(A, &F);
    ldlt_subst_rcp<6>(&F, b, x1);
    ++total;
    if (memcmp(x0, x1, sizeof x0) != 0) {
      bool bothnan = true;
      for (int i = 0; i < 6; ++i) if (!(std::isnan(x0[i]) && std::isnan(x1[i])) && memcmp(&x0[i], &x1[i], 8) != 0) bothnan = false;
      if (!bothnan) { if (bad < 5) { printf("trial %d kind %d:", trial, kind); for (int i = 0; i < 6; ++i) printf(" %.17g/%.17g", x0[i], x1[i]); printf("\n"); } ++bad; }
    }
    // second right-hand side with the same factor
    for (int i = 0; i < 6; ++i) b[i] = rnd();
    ldlt_solve_fixed<6, true>(A, b, x0);
    ldlt_subst_rcp<6>(&F, b, x1);
    for (int i = 0; i < 6; ++i) if (memcmp(&x0[i], &x1[i], 8) != 0 && !(std::isnan(x0[i]) && std::isnan(x1[i]))) { ++bad; break; }
  }
  printf("%d systems, %d mismatches\n", total, bad);
  return bad != 0;
}
