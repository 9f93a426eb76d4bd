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
