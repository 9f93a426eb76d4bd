// ldlt.cuh — Eigen 3.4's A.ldlt().solve(b) for FIXED-size symmetric systems (N = 3, 6) on the device, bit-identical to
// the reference's x86-64 SSE2 host build (checked on the host against Eigen itself on 2,000 random systems per size,
// and through the oracle against the compiled reference: tests/test_map_oracle.py).
//
// What has to be reproduced beyond the textbook algorithm (Eigen/src/Cholesky/LDLT.h, SolveTriangular.h, Redux.h):
//   * largest-|diagonal| symmetric pivoting, unblocked in-place factorisation, sequential inner products;
//   * the triangular solves on a fixed-size right-hand side are fully unrolled and subtract ONE reduced sum per row;
//     - lower solve: rows of a column-major matrix are strided => scalar reduction by halving (redux_novec_unroller):
//       sum(0..m) = sum(0..m/2) + sum(m/2..m);
//     - upper solve: matrixU() is the adjoint view, its rows are contiguous columns => Packet2d reduction
//       (LinearVectorizedTraversal + CompleteUnrolling): the two lanes are summed by halving, then added, then the odd tail;
//   * D^-1 with the cut-off |d| > 1/highest() ~ DBL_MIN.
// The library is compiled with -fmad=false, so none of the sums below is contracted.
#pragma once

namespace ldlt_detail {
__device__ __forceinline__ void swapd(double& a, double& b) { const double t = a; a = b; b = t; }
__device__ __forceinline__ double halving_sum(const double* t, int m)
{
  switch (m) {
    case 1: return t[0];
    case 2: return t[0] + t[1];
    case 3: return t[0] + (t[1] + t[2]);
    case 4: return (t[0] + t[1]) + (t[2] + t[3]);
    case 5: return (t[0] + t[1]) + (t[2] + (t[3] + t[4]));
  }
  return 0.0;
}
__device__ __forceinline__ double packet2_sum(const double* t, int m)
{
  switch (m) {
    case 1: return t[0];
    case 2: return t[0] + t[1];
    case 3: return (t[0] + t[1]) + t[2];
    case 4: return (t[0] + t[2]) + (t[1] + t[3]);
    case 5: return ((t[0] + t[2]) + (t[1] + t[3])) + t[4];
  }
  return 0.0;
}
}  // namespace ldlt_detail

// A: row-major N x N (only its lower triangle is read after the copy), b, x: N
// SHARED_RCP = false: every quotient is an IEEE division, bit-identical to Eigen (Point::optimize, pose optimiser).
// SHARED_RCP = true : the quotients by one pivot (its column of L and its entry of D^-1) share ONE division, the
//   reciprocal of the pivot, and are corrected with two fused residual steps — q0 = a*r, q = fma(fma(-d,q0,a), r, q0) —
//   which reproduces the correctly rounded quotient except in rare last-bit cases.  An FP64 division is a ~200-cycle
//   dependent sequence with a slow-path branch, and a 6x6 solve has 21 of them on one thread: for the sparse-alignment
//   solve, whose inputs are tree-reduced sums (tolerance-matched already), this removes 15 of the 21.
template <int N, bool SHARED_RCP = false>
__device__ inline void ldlt_solve_fixed(const double* Ain, const double* b, double* x)
{
  using namespace ldlt_detail;
  double A[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) A[i][j] = Ain[i * N + j];
  int perm[N];
  double rcp[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double big = fabs(A[k][k]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) { const double v = fabs(A[i][i]); if (v > big) { big = v; piv = i; } }
    perm[k] = piv;
#pragma unroll
    for (int p = k + 1; p < N; ++p) {
      if (piv == p) {
#pragma unroll
        for (int j = 0; j < k; ++j) swapd(A[k][j], A[p][j]);
#pragma unroll
        for (int j = p + 1; j < N; ++j) swapd(A[j][k], A[j][p]);
        swapd(A[k][k], A[p][p]);
#pragma unroll
        for (int i = k + 1; i < p; ++i) swapd(A[i][k], A[p][i]);
      }
    }
    if (k > 0) {
      double temp[N];
      double sdiag = 0;
#pragma unroll
      for (int j = 0; j < k; ++j) { temp[j] = A[j][j] * A[k][j]; sdiag += A[k][j] * temp[j]; }
      A[k][k] -= sdiag;
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        double t = 0;
#pragma unroll
        for (int j = 0; j < k; ++j) t += A[i][j] * temp[j];
        A[i][k] -= t;
      }
    }
    const double d = A[k][k];
    if (fabs(d) > 2.2250738585072014e-308) {
      if (SHARED_RCP) {
        const double r = 1.0 / d;
        rcp[k] = r;
#pragma unroll
        for (int i = k + 1; i < N; ++i) { const double q0 = A[i][k] * r; A[i][k] = fma(fma(-d, q0, A[i][k]), r, q0); }
      } else {
#pragma unroll
        for (int i = k + 1; i < N; ++i) A[i][k] /= d;
      }
    } else if (SHARED_RCP) rcp[k] = 0.0;
  }
  double y[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = b[i];
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int p = k + 1; p < N; ++p) if (perm[k] == p) swapd(y[k], y[p]);
  }
#pragma unroll
  for (int i = 1; i < N; ++i) {
    double t[N];
#pragma unroll
    for (int j = 0; j < i; ++j) t[j] = A[i][j] * y[j];
    y[i] -= halving_sum(t, i);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double d = A[i][i];
    if (SHARED_RCP) { const double q0 = y[i] * rcp[i]; y[i] = (fabs(d) > 2.2250738585072014e-308) ? fma(fma(-d, q0, y[i]), rcp[i], q0) : 0.0; }
    else y[i] = (fabs(d) > 2.2250738585072014e-308) ? y[i] / d : 0.0;
  }
#pragma unroll
  for (int i = N - 2; i >= 0; --i) {
    double t[N];
#pragma unroll
    for (int j = 0; j < N - 1 - i; ++j) t[j] = A[i + 1 + j][i] * y[i + 1 + j];
    y[i] -= packet2_sum(t, N - 1 - i);
  }
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
#pragma unroll
    for (int p = k + 1; p < N; ++p) if (perm[k] == p) swapd(y[k], y[p]);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = y[i];
}

// ---------------------------------------------------------------- the same solve, split into factor and substitution
// ldlt_solve_fixed<N, true>(A, b, x) == ldlt_factor_rcp<N>(A, F); ldlt_subst_rcp<N>(F, b, x), bit for bit — for a caller
// whose matrix stays the same over several right-hand sides (the inverse-compositional Hessian of the sparse alignment).
//
// Eigen's unblocked LDLT searches its pivot among the diagonal entries k .. N-1 BEFORE it updates any of them (only entry k
// is updated, at step k, after the search): the pivot order is a selection sort of the ORIGINAL |diagonal| and does not
// depend on the factorisation.  So the transpositions are worked out first, on an index word, and the factorisation runs on
// the permuted matrix with compile-time indices — none of the predicated row / column swaps of the interleaved form.
template <int N>
struct LdltFactor {
  double L[N * (N - 1) / 2 > 0 ? N * (N - 1) / 2 : 1];   // strictly lower triangle of the permuted matrix, row i at i(i-1)/2
  double D[N];                                           // the pivots
  double rcp[N];                                         // 1 / pivot (0 below the cut-off)
  unsigned idx;                                          // nibble r = original index at permuted position r
};

// A: row-major N x N, SYMMETRIC (both triangles valid; any address space: only dynamic indexing is by the pivot order)
template <int N>
__device__ inline void ldlt_factor_rcp(const double* A, LdltFactor<N>* F)
{
  static_assert(N <= 8, "index word holds eight nibbles");
  unsigned idx = 0x76543210u;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double big = fabs(A[((idx >> (4 * k)) & 15u) * (N + 1)]);
#pragma unroll
    for (int i = k + 1; i < N; ++i) { const double v = fabs(A[((idx >> (4 * i)) & 15u) * (N + 1)]); if (v > big) { big = v; piv = i; } }
    // swap nibbles k and piv
    const unsigned a = (idx >> (4 * k)) & 15u, b = (idx >> (4 * piv)) & 15u, x = a ^ b;
    idx ^= (x << (4 * k)) | (x << (4 * piv));           // piv == k: x = 0
  }
  double M[N][N];                                        // lower triangle of P A P^T
#pragma unroll
  for (int r = 0; r < N; ++r) {
    const unsigned ir = (idx >> (4 * r)) & 15u;
#pragma unroll
    for (int c = 0; c <= r; ++c) M[r][c] = A[ir * N + ((idx >> (4 * c)) & 15u)];
  }
  double rcp[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (k > 0) {
      double temp[N];
      double sdiag = 0;
#pragma unroll
      for (int j = 0; j < k; ++j) { temp[j] = M[j][j] * M[k][j]; sdiag += M[k][j] * temp[j]; }
      M[k][k] -= sdiag;
#pragma unroll
      for (int i = k + 1; i < N; ++i) {
        double t = 0;
#pragma unroll
        for (int j = 0; j < k; ++j) t += M[i][j] * temp[j];
        M[i][k] -= t;
      }
    }
    const double d = M[k][k];
    if (fabs(d) > 2.2250738585072014e-308) {
      const double r = 1.0 / d;
      rcp[k] = r;
#pragma unroll
      for (int i = k + 1; i < N; ++i) { const double q0 = M[i][k] * r; M[i][k] = fma(fma(-d, q0, M[i][k]), r, q0); }
    } else rcp[k] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) F->L[i * (i - 1) / 2 + j] = M[i][j];
    F->D[i] = M[i][i];
    F->rcp[i] = rcp[i];
  }
  F->idx = idx;
}

template <int N>
__device__ inline void ldlt_subst_rcp(const LdltFactor<N>* F, const double* b, double* x)
{
  using namespace ldlt_detail;
  const unsigned idx = F->idx;
  double L[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) L[i][j] = F->L[i * (i - 1) / 2 + j];
  double y[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = b[(idx >> (4 * i)) & 15u];
#pragma unroll
  for (int i = 1; i < N; ++i) {
    double t[N];
#pragma unroll
    for (int j = 0; j < i; ++j) t[j] = L[i][j] * y[j];
    y[i] -= halving_sum(t, i);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double d = F->D[i], r = F->rcp[i];
    const double q0 = y[i] * r;
    y[i] = (fabs(d) > 2.2250738585072014e-308) ? fma(fma(-d, q0, y[i]), r, q0) : 0.0;
  }
#pragma unroll
  for (int i = N - 2; i >= 0; --i) {
    double t[N];
#pragma unroll
    for (int j = 0; j < N - 1 - i; ++j) t[j] = L[i + 1 + j][i] * y[i + 1 + j];
    y[i] -= packet2_sum(t, N - 1 - i);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) x[(idx >> (4 * i)) & 15u] = y[i];
}
