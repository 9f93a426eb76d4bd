/*
 * svo_oracle_map.c — CPU restatement (plain C99) of the callers and data formats either side of
 * the tracking hot path (SURVEY.md §8f "next" rows).  TEST INFRASTRUCTURE ONLY (see svo_oracle.h).
 *
 *   f1  Reprojector::reprojectMap / reprojectCell / reprojectPoint     reprojector.cpp:72-259
 *       Point::getCloseViewObs                                         point.cpp:101-125
 *   f2  pose_optimizer::optimizeGaussNewton                            pose_optimizer.cpp:31-181
 *       Point::optimize, Point::jacobian_xyz2uv                        point.cpp:130-192, point.h
 *       FrameHandlerBase::optimizeStructure (selection rule)           frame_handler_base.cpp:190-210
 *   f3  YUV_420_888 -> RGBA (ImageProcess::GetCVImage, YUV2RGB)        ../image_process.cpp:97-186
 *       cv::cvtColor(img, COLOR_RGBA2GRAY)                             ../svo_system.cpp:49-51
 *   f4  DepthFilter::initializeSeeds (grid occupancy, Seed ctor)       depth_filter.cpp:36-45, :129-151
 *       FastDetector::setExistingFeatures / setGridOccpuancy           feature_detection.cpp:40-58
 *
 * Paths are relative to /root/reference/app/src/main/cpp/svo.
 */
#include "svo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------ small helpers */
static inline double dot3(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline int in_frame(const svo_cam* cam, int x, int y, int boundary)
{ return x >= boundary && x < cam->width - boundary && y >= boundary && y < cam->height - boundary; }

/* Frame::pos() = T_f_w_.inverse().translation_vec() (frame.h:105) */
void svo_oracle_frame_pos(const double T_f_w[7], double pos[3])
{
  double inv[7];
  svo_oracle_se3_inverse(T_f_w, inv);
  pos[0] = inv[0]; pos[1] = inv[1]; pos[2] = inv[2];
}

/* SO3::getMatrix SO3.h:396-410, row-major */
static void q_matrix(const double q[4], double m[9])
{
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  m[0] = 1.0 - 2.0 * (y2 + z2); m[1] = 2.0 * (xy - wz);       m[2] = 2.0 * (xz + wy);
  m[3] = 2.0 * (xy + wz);       m[4] = 1.0 - 2.0 * (x2 + z2); m[5] = 2.0 * (yz - wx);
  m[6] = 2.0 * (xz - wy);       m[7] = 2.0 * (yz + wx);       m[8] = 1.0 - 2.0 * (x2 + y2);
}

static double halving_sum(const double* t, int m)
{
  if (m == 1) return t[0];
  const int h = m / 2;
  return halving_sum(t, h) + halving_sum(t + h, m - h);
}

static double packet2_sum(const double* t, int m)
{
  const int np = m / 2;
  if (np == 0) return halving_sum(t, m);
  double p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0};
  for (int i = 0; i < np; ++i) { p0[i] = t[2 * i]; p1[i] = t[2 * i + 1]; }
  double r = halving_sum(p0, np) + halving_sum(p1, np);
  if (m & 1) r = r + t[m - 1];
  return r;
}

/* Eigen 3.4 LDLT (pivoted, lower, unblocked: Eigen/src/Cholesky/LDLT.h) + solve for a FIXED-size system, n <= 6,
 * x86-64 SSE2 build: bit-identical to A.ldlt().solve(b) (checked on 2,000 random systems per size). */
static void ldlt_solve(int n, const double* Ain, const double* b, double* x)
{
  double A[36]; int perm[6];
  for (int i = 0; i < n * n; ++i) A[i] = Ain[i];
#define AT(i, j) A[(i) * n + (j)]
  for (int k = 0; k < n; ++k) {
    int piv = k; double big = fabs(AT(k, k));
    for (int i = k + 1; i < n; ++i) { const double v = fabs(AT(i, i)); if (v > big) { big = v; piv = i; } }
    perm[k] = piv;
    if (piv != k) {
      const int s = n - piv - 1;
      for (int j = 0; j < k; ++j) { const double t = AT(k, j); AT(k, j) = AT(piv, j); AT(piv, j) = t; }
      for (int j = 0; j < s; ++j) { const double t = AT(piv + 1 + j, k); AT(piv + 1 + j, k) = AT(piv + 1 + j, piv); AT(piv + 1 + j, piv) = t; }
      { const double t = AT(k, k); AT(k, k) = AT(piv, piv); AT(piv, piv) = t; }
      for (int i = k + 1; i < piv; ++i) { const double t = AT(i, k); AT(i, k) = AT(piv, i); AT(piv, i) = t; }
    }
    const int rs = n - k - 1;
    if (k > 0) {
      double temp[6];
      for (int j = 0; j < k; ++j) temp[j] = AT(j, j) * AT(k, j);
      double s = 0; for (int j = 0; j < k; ++j) s += AT(k, j) * temp[j];
      AT(k, k) -= s;
      for (int i = 0; i < rs; ++i) {
        double t = 0; for (int j = 0; j < k; ++j) t += AT(k + 1 + i, j) * temp[j];
        AT(k + 1 + i, k) -= t;
      }
    }
    const double d = AT(k, k);
    if (rs > 0 && fabs(d) > DBL_MIN)
      for (int i = 0; i < rs; ++i) AT(k + 1 + i, k) /= d;
  }
  double y[6];
  for (int i = 0; i < n; ++i) y[i] = b[i];
  for (int k = 0; k < n; ++k) if (perm[k] != k) { const double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
  /* fixed-size rhs: Eigen's triangular_solver_unroller subtracts ONE sum per row, and the sum of the strided
   * cwiseProduct is the unrolled halving reduction (redux_novec_unroller): sum(0..m) = sum(0..m/2) + sum(m/2..m) */
  for (int i = 1; i < n; ++i) { double t[6]; for (int j = 0; j < i; ++j) t[j] = AT(i, j) * y[j]; y[i] -= halving_sum(t, i); }
  for (int i = 0; i < n; ++i) { const double d = AT(i, i); y[i] = (fabs(d) > DBL_MIN) ? y[i] / d : 0.0; }
  /* matrixU() = adjoint view: its rows are contiguous columns of the factor, so this reduction IS vectorised
   * (Packet2d, LinearVectorizedTraversal + CompleteUnrolling): lanes summed by halving, then predux, then the odd tail */
  for (int i = n - 2; i >= 0; --i) { double t[6]; const int m = n - 1 - i; for (int j = 0; j < m; ++j) t[j] = AT(i + 1 + j, i) * y[i + 1 + j]; y[i] -= packet2_sum(t, m); }
  for (int k = n - 1; k >= 0; --k) if (perm[k] != k) { const double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
  for (int i = 0; i < n; ++i) x[i] = y[i];
#undef AT
}

static int cmp_float(const void* a, const void* b) { const float x = *(const float*)a, y = *(const float*)b; return (x > y) - (x < y); }
static int cmp_double(const void* a, const void* b) { const double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }
/* vk::getMedian math_utils.h:125-131: nth_element at floor(n/2) == element n/2 of the sorted data */
static float median_f(const float* v, int n)
{
  float* t = (float*)malloc(sizeof(float) * (size_t)n);
  memcpy(t, v, sizeof(float) * (size_t)n);
  qsort(t, (size_t)n, sizeof(float), cmp_float);
  const float m = t[n / 2];
  free(t);
  return m;
}
static double median_d(const double* v, int n)
{
  double* t = (double*)malloc(sizeof(double) * (size_t)n);
  memcpy(t, v, sizeof(double) * (size_t)n);
  qsort(t, (size_t)n, sizeof(double), cmp_double);
  const double m = t[n / 2];
  free(t);
  return m;
}

/* ------------------------------------------------------------------ f1  Reprojector */
/* reprojector.cpp:72-168 for ONE current frame.  points are given in the reference's insertion order
 * (features of the close keyframes in closeness order, then the point candidates); obs of point i are
 * obs[obs_begin[i] .. obs_end[i]) in Point::obs_ list order.  The grid, the per-cell stable sort by
 * point type (:184, :170-175), the first-success-per-cell rule (:186-240) and the maxFts break (:164)
 * are restated; the side effects on Point (n_failed_reproj_, type_ promotion, deletion) are left to the
 * caller, which reads them off the per-point status. */
int svo_oracle_reproject_map(const svo_pyr* const* ref_pyrs /* per keyframe */, const svo_pyr* cur, const svo_cam* cam,
                             const double T_cur_w[7], int n_points, const svo_map_point* points, const svo_point_obs* obs,
                             const double* T_kf_w /* 7 per keyframe */, int cell_size, int max_fts, const svo_matcher_opts* mopts,
                             svo_reproj_result* results, int* cell_winner /* n_cells, -1 = none */, int* n_matches_out, int* n_trials_out)
{
  const int cols = (int)ceil((double)cam->width / cell_size), rows = (int)ceil((double)cam->height / cell_size);
  const int n_cells = cols * rows;
  double cur_pos[3];
  svo_oracle_frame_pos(T_cur_w, cur_pos);
  /* reprojectPoint :246-259 */
  for (int i = 0; i < n_points; ++i) {
    svo_reproj_result* r = &results[i];
    memset(r, 0, sizeof(*r));
    r->status = SVO_REPROJ_NOT_IN_FRAME; r->cell = -1; r->obs = -1;
    double pf[3];
    svo_oracle_se3_transform(T_cur_w, points[i].pos, pf);
    svo_oracle_world2cam(cam, pf, r->px);
    if (in_frame(cam, (int)r->px[0], (int)r->px[1], 8)) {
      r->cell = (int)(r->px[1] / cell_size) * cols + (int)(r->px[0] / cell_size);
      r->status = SVO_REPROJ_UNTRIED;
    }
  }
  for (int c = 0; c < n_cells; ++c) cell_winner[c] = -1;
  int n_matches = 0, n_trials = 0;
  int* order = (int*)malloc(sizeof(int) * (size_t)(n_points > 0 ? n_points : 1));
  for (int c = 0; c < n_cells; ++c) {
    /* the cell's list in insertion order, then list::sort (stable) by type descending */
    int m = 0;
    for (int i = 0; i < n_points; ++i) if (results[i].cell == c) order[m++] = i;
    for (int a = 1; a < m; ++a) {
      const int v = order[a]; int b = a - 1;
      while (b >= 0 && points[order[b]].type < points[v].type) { order[b + 1] = order[b]; --b; }
      order[b + 1] = v;
    }
    int found = 0;
    for (int a = 0; a < m && !found; ++a) {
      const int i = order[a];
      svo_reproj_result* r = &results[i];
      ++n_trials;
      if (points[i].type == SVO_POINT_DELETED) { r->status = SVO_REPROJ_DELETED; continue; }
      /* Matcher::findMatchDirect (matcher.cpp:156-202) */
      int ok = 0;
      const int nobs = points[i].obs_end - points[i].obs_begin;
      if (nobs > 0) {
        double* opos = (double*)malloc(sizeof(double) * 3 * (size_t)nobs);
        for (int k = 0; k < nobs; ++k) svo_oracle_frame_pos(T_kf_w + 7 * obs[points[i].obs_begin + k].keyframe, opos + 3 * k);
        int best = 0;
        const int close = svo_oracle_close_view_obs(cur_pos, points[i].pos, nobs, opos, &best);
        const svo_point_obs* o = &obs[points[i].obs_begin + best];
        r->obs = points[i].obs_begin + best;
        if (close) {
          double inv[7], T_cur_ref[7], d[3];
          svo_oracle_se3_inverse(T_kf_w + 7 * o->keyframe, inv);
          svo_oracle_se3_mul(T_cur_w, inv, T_cur_ref);
          for (int k = 0; k < 3; ++k) d[k] = opos[3 * best + k] - points[i].pos[k];
          const double depth_ref = sqrt(dot3(d, d));
          svo_match_result mr;
          ok = svo_oracle_find_match_direct(ref_pyrs[o->keyframe], cur, cam, &o->ftr, depth_ref, T_cur_ref, mopts, r->px, &mr);
          r->search_level = mr.search_level;
          for (int k = 0; k < 4; ++k) r->A_cur_ref[k] = mr.A_cur_ref[k];
          r->px[0] = mr.px_cur[0]; r->px[1] = mr.px_cur[1];
        }
        free(opos);
      }
      if (!ok) { r->status = SVO_REPROJ_FAILED; continue; }
      r->status = SVO_REPROJ_MATCHED;
      cell_winner[c] = i;
      found = 1;
    }
    if (found) ++n_matches;
    if (n_matches > max_fts) break;
  }
  free(order);
  *n_matches_out = n_matches; *n_trials_out = n_trials;
  return n_matches;
}

/* ------------------------------------------------------------------ f2  pose optimizer */
/* vk::robust_cost::TukeyWeightFunction::value (robust_cost.cpp), b_square = DEFAULT_B^2, DEFAULT_B = 8.6851f in this fork (robust_cost.cpp:87) */
static float tukey_weight(float x, float b_square)
{
  const float x_square = x * x;
  if (x_square <= b_square) { const float tmp = 1.0f - x_square / b_square; return tmp * tmp; }
  return 0.0f;
}

/* Frame::jacobian_xyz2uv frame.h:110-132 */
static void frame_jacobian_xyz2uv(const double p[3], double J[12])
{
  const double x = p[0], y = p[1], z_inv = 1. / p[2], z_inv_2 = z_inv * z_inv;
  J[0] = -z_inv; J[1] = 0.0; J[2] = x * z_inv_2; J[3] = y * J[2]; J[4] = -(1.0 + x * J[2]); J[5] = y * z_inv;
  J[6] = 0.0; J[7] = -z_inv; J[8] = y * z_inv_2; J[9] = 1.0 + y * J[8]; J[10] = -J[3]; J[11] = -x * z_inv;
}

/* pose_optimizer.cpp:31-181.  f: 3 per feature (bearing), level: 1 per feature, pos: 3 per feature (point in
 * world); T_f_w is in/out.  outlier[i] = 1 where the reference resets ftr->point = NULL (:149-153).
 * tukey_b: TukeyWeightFunction::DEFAULT_B. */
void svo_oracle_pose_optimize(const svo_cam* cam, int n, const double* f, const int* level, const double* pos, double reproj_thresh,
                              int n_iter, double eps, float tukey_b, double T_f_w[7], svo_pose_opt_result* res, uint8_t* outlier)
{
  memset(res, 0, sizeof(*res));
  for (int i = 0; i < n; ++i) outlier[i] = 0;
  if (n <= 0) return;                                                     /* errors.empty() :56-57 */
  const double em2 = fabs(cam->fx);                                       /* errorMultiplier2 pinhole_camera.h:64-67 */
  const float b_square = tukey_b * tukey_b;
  double chi2 = 0.0, T_old[7], A[36], b[6];
  memcpy(T_old, T_f_w, sizeof(T_old));
  float* errors = (float*)malloc(sizeof(float) * (size_t)n);
  double* chi2_init = (double*)malloc(sizeof(double) * (size_t)n);
  double* chi2_final = (double*)malloc(sizeof(double) * (size_t)n);
  int n_init = 0;
  for (int i = 0; i < n; ++i) {
    double p[3];
    svo_oracle_se3_transform(T_f_w, pos + 3 * i, p);
    double e0 = f[3 * i] / f[3 * i + 2] - p[0] / p[2], e1 = f[3 * i + 1] / f[3 * i + 2] - p[1] / p[2];
    const double s = 1.0 / (1 << level[i]);
    e0 *= s; e1 *= s;
    errors[i] = (float)sqrt(e0 * e0 + e1 * e1);
  }
  const float estimated_scale_f = 1.48f * median_f(errors, n);            /* MADScaleEstimator robust_cost.cpp */
  double estimated_scale = estimated_scale_f;
  double scale = estimated_scale;
  memset(A, 0, sizeof(A));
  int iters = 0;
  for (int iter = 0; iter < n_iter; ++iter) {
    if (iter == 5) scale = 0.85 / em2;
    memset(A, 0, sizeof(A)); memset(b, 0, sizeof(b));
    double new_chi2 = 0.0;
    for (int i = 0; i < n; ++i) {
      double p[3], J[12];
      svo_oracle_se3_transform(T_f_w, pos + 3 * i, p);
      frame_jacobian_xyz2uv(p, J);
      double e0 = f[3 * i] / f[3 * i + 2] - p[0] / p[2], e1 = f[3 * i + 1] / f[3 * i + 2] - p[1] / p[2];
      const double sic = 1.0 / (1 << level[i]);
      e0 *= sic; e1 *= sic;
      if (iter == 0) chi2_init[n_init++] = e0 * e0 + e1 * e1;
      for (int k = 0; k < 12; ++k) J[k] *= sic;
      const double weight = tukey_weight((float)(sqrt(e0 * e0 + e1 * e1) / scale), b_square);
      for (int r = 0; r < 6; ++r) {
        for (int c = 0; c < 6; ++c) A[r * 6 + c] += (J[r] * J[c] + J[6 + r] * J[6 + c]) * weight;
        b[r] -= (J[r] * e0 + J[6 + r] * e1) * weight;
      }
      new_chi2 += (e0 * e0 + e1 * e1) * weight;
    }
    double dT[6];
    ldlt_solve(6, A, b, dT);
    ++iters;
    if ((iter > 0 && new_chi2 > chi2 * 1.2) || isnan(dT[0])) { memcpy(T_f_w, T_old, sizeof(T_old)); res->rolled_back = 1; break; }
    double E[7], T_new[7];
    svo_oracle_se3_exp(dT, E);
    svo_oracle_se3_mul(E, T_f_w, T_new);
    memcpy(T_old, T_f_w, sizeof(T_old));
    memcpy(T_f_w, T_new, sizeof(T_new));
    chi2 = new_chi2;
    double nm = -1;
    for (int k = 0; k < 6; ++k) { const double a = fabs(dT[k]); if (a > nm) nm = a; }
    if (nm <= eps) break;
  }
  res->iters = iters; res->chi2 = chi2;
  memcpy(res->A, A, sizeof(A));                                            /* Cov_ = (A * em2^2)^-1 :139-140 is left to the caller */
  const double thr = reproj_thresh / em2;
  int n_deleted = 0;
  for (int i = 0; i < n; ++i) {
    double p[3];
    svo_oracle_se3_transform(T_f_w, pos + 3 * i, p);
    double e0 = f[3 * i] / f[3 * i + 2] - p[0] / p[2], e1 = f[3 * i + 1] / f[3 * i + 2] - p[1] / p[2];
    const double s = 1.0 / (1 << level[i]);
    e0 *= s; e1 *= s;
    chi2_final[i] = e0 * e0 + e1 * e1;
    if (sqrt(e0 * e0 + e1 * e1) > thr) { outlier[i] = 1; ++n_deleted; }
  }
  res->error_init = n_init ? sqrt(median_d(chi2_init, n_init)) * em2 : 0.0;
  res->error_final = sqrt(median_d(chi2_final, n)) * em2;
  res->estimated_scale = estimated_scale * em2;
  res->num_obs = n - n_deleted;
  free(errors); free(chi2_init); free(chi2_final);
}

/* Point::optimize point.cpp:130-192 for one point: obs k has pose T_f_w[7k..] and bearing f[3k..] */
int svo_oracle_point_optimize(int n_obs, const double* T_f_w, const double* f, int n_iter, double eps, double pos[3])
{
  double old_point[3] = { pos[0], pos[1], pos[2] };
  double chi2 = 0.0;
  int it = 0;
  for (int i = 0; i < n_iter; ++i) {
    double A[9] = {0}, b[3] = {0}, new_chi2 = 0.0;
    for (int k = 0; k < n_obs; ++k) {
      double p[3], R[9], J0[6], J[6];
      svo_oracle_se3_transform(T_f_w + 7 * k, pos, p);
      q_matrix(T_f_w + 7 * k + 3, R);
      const double z_inv = 1.0 / p[2], z_inv_sq = z_inv * z_inv;
      J0[0] = z_inv; J0[1] = 0.0; J0[2] = -p[0] * z_inv_sq; J0[3] = 0.0; J0[4] = z_inv; J0[5] = -p[1] * z_inv_sq;
      /* point_jac = -point_jac * R_f_w (Eigen 2x3 * 3x3 lazy product: sequential sums, negated operand) */
      for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c)
        J[r * 3 + c] = ((-J0[r * 3]) * R[c] + (-J0[r * 3 + 1]) * R[3 + c]) + (-J0[r * 3 + 2]) * R[6 + c];
      const double e0 = f[3 * k] / f[3 * k + 2] - p[0] / p[2], e1 = f[3 * k + 1] / f[3 * k + 2] - p[1] / p[2];
      new_chi2 += e0 * e0 + e1 * e1;
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) A[r * 3 + c] += J[r] * J[c] + J[3 + r] * J[3 + c];
        b[r] -= J[r] * e0 + J[3 + r] * e1;
      }
    }
    double dp[3];
    ldlt_solve(3, A, b, dp);
    ++it;
    if ((i > 0 && new_chi2 > chi2) || isnan(dp[0])) { pos[0] = old_point[0]; pos[1] = old_point[1]; pos[2] = old_point[2]; break; }
    for (int k = 0; k < 3; ++k) { old_point[k] = pos[k]; pos[k] = pos[k] + dp[k]; }
    chi2 = new_chi2;
    double nm = -1;
    for (int k = 0; k < 3; ++k) { const double a = fabs(dp[k]); if (a > nm) nm = a; }
    if (nm <= eps) break;
  }
  return it;
}

/* ------------------------------------------------------------------ f3  camera input stage */
/* YUV2RGB ../image_process.cpp:97-126 -> packed 0xAARRGGBB word, stored little-endian: bytes B,G,R,A */
static inline uint32_t yuv2rgb(int nY, int nU, int nV)
{
  nY -= 16; nU -= 128; nV -= 128;
  if (nY < 0) nY = 0;
  int nR = 1192 * nY + 1634 * nV;
  int nG = 1192 * nY - 833 * nV - 400 * nU;
  int nB = 1192 * nY + 2066 * nU;
  const int kMax = 262143;
  nR = nR < 0 ? 0 : (nR > kMax ? kMax : nR);
  nG = nG < 0 ? 0 : (nG > kMax ? kMax : nG);
  nB = nB < 0 ? 0 : (nB > kMax ? kMax : nB);
  nR = (nR >> 10) & 0xff; nG = (nG >> 10) & 0xff; nB = (nB >> 10) & 0xff;
  return 0xff000000u | ((uint32_t)nR << 16) | ((uint32_t)nG << 8) | (uint32_t)nB;
}

/* ImageProcess::GetCVImage ../image_process.cpp:151-186 (crop rect = whole image): planes as AImage hands them out */
void svo_oracle_yuv420_to_rgba(const uint8_t* y, int y_stride, const uint8_t* u, const uint8_t* v, int uv_stride, int uv_pixel_stride,
                               int w, int h, uint8_t* rgba /* w*h*4 */)
{
  for (int r = 0; r < h; ++r) {
    const uint8_t* pY = y + (size_t)y_stride * r;
    const uint8_t* pU = u + (size_t)uv_stride * (r >> 1);
    const uint8_t* pV = v + (size_t)uv_stride * (r >> 1);
    for (int x = 0; x < w; ++x) {
      const int o = (x >> 1) * uv_pixel_stride;
      const uint32_t px = yuv2rgb(pY[x], pU[o], pV[o]);
      uint8_t* d = rgba + 4 * ((size_t)r * w + x);
      d[0] = (uint8_t)(px & 0xff); d[1] = (uint8_t)((px >> 8) & 0xff); d[2] = (uint8_t)((px >> 16) & 0xff); d[3] = (uint8_t)(px >> 24);
    }
  }
}

/* cv::cvtColor(COLOR_RGBA2GRAY) for 8-bit input: OpenCV 4.x RGB2Gray<uchar> fixed point,
 * gray = (c0*RY15 + c1*GY15 + c2*BY15 + 2^14) >> 15 with RY15 = 9798, GY15 = 19235, BY15 = 3735
 * (imgproc/src/color_rgb.simd.hpp, color.simd_helpers.hpp; third-party, pinned against python cv2). */
void svo_oracle_rgba_to_gray(const uint8_t* rgba, int w, int h, uint8_t* gray)
{
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    const uint8_t* p = rgba + 4 * i;
    gray[i] = (uint8_t)((p[0] * 9798 + p[1] * 19235 + p[2] * 3735 + (1 << 14)) >> 15);
  }
}

/* the app's whole input stage: YUV_420_888 -> RGBA -> gray (what FrameHandlerMono::addImage receives) */
void svo_oracle_yuv420_to_gray(const uint8_t* y, int y_stride, const uint8_t* u, const uint8_t* v, int uv_stride, int uv_pixel_stride,
                               int w, int h, uint8_t* gray)
{
  uint8_t* rgba = (uint8_t*)malloc((size_t)w * h * 4);
  svo_oracle_yuv420_to_rgba(y, y_stride, u, v, uv_stride, uv_pixel_stride, w, h, rgba);
  svo_oracle_rgba_to_gray(rgba, w, h, gray);
  free(rgba);
}

/* ------------------------------------------------------------------ f4  seed initialisation */
/* AbstractDetector::setExistingFeatures / setGridOccpuancy feature_detection.cpp:40-58: cell of a level-0 pixel */
void svo_oracle_grid_occupancy(const svo_cam* cam, int cell_size, int n, const double* px, uint8_t* occupancy)
{
  const int cols = (int)ceil((double)cam->width / cell_size);
  for (int i = 0; i < n; ++i)
    occupancy[(int)(px[2 * i + 1] / cell_size) * cols + (int)(px[2 * i] / cell_size)] = 1;
}

/* DepthFilter::initializeSeeds depth_filter.cpp:129-151: occupancy from the frame's features, detect, one Seed
 * (depth_filter.cpp:36-45) per new corner in cell order.  Returns the number of new seeds. */
int svo_oracle_initialize_seeds(const svo_pyr* pyr, const svo_cam* cam, int n_detect_levels, int cell_size, double thr,
                                int n_existing, const double* existing_px, float depth_mean, float depth_min,
                                svo_corner* corners_out, svo_seed* seeds_out)
{
  const int cols = (int)ceil((double)cam->width / cell_size), rows = (int)ceil((double)cam->height / cell_size);
  const int n_cells = cols * rows;
  uint8_t* occ = (uint8_t*)calloc((size_t)n_cells, 1);
  svo_corner* cells = (svo_corner*)malloc(sizeof(svo_corner) * (size_t)n_cells);
  svo_oracle_grid_occupancy(cam, cell_size, n_existing, existing_px, occ);
  svo_oracle_fast_detect(pyr, n_detect_levels, cell_size, thr, occ, cells);
  int m = 0;
  for (int c = 0; c < n_cells; ++c) {
    if (!((double)cells[c].score > thr)) continue;
    corners_out[m] = cells[c];
    svo_oracle_seed_init(&seeds_out[m], depth_mean, depth_min);
    ++m;
  }
  free(occ); free(cells);
  return m;
}
