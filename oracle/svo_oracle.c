/*
 * svo_oracle.c — CPU restatement of the reference's semi-direct tracking front
 * end in plain C99.  TEST INFRASTRUCTURE ONLY (see svo_oracle.h).
 *
 * Build with -O2 -ffp-contract=off: the reference's host build is baseline
 * x86-64 (no FMA), so every float/double op here must round individually.
 * Operation order and float/double promotions follow the cited reference
 * lines; Eigen's fixed-size reductions are written out in the association
 * Eigen 3.4 generates (checked bit-for-bit against oracle/_ref in tests).
 *
 * Paths are relative to /root/reference/app/src/main/cpp/svo.
 */
#include "svo_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>

/* ------------------------------------------------------------------ */
/* a1/a2  pyramid                                                     */
/* ------------------------------------------------------------------ */

/* vision.cpp:78 — SSE2 path iff 16-byte aligned and in.cols % 16 == 0 (cv::Mat
 * storage is always aligned, so only the width matters). */
int svo_oracle_half_sample_mode_x86(int in_cols) { return (in_cols % 16) == 0 ? SVO_ROUND_SSE2 : SVO_ROUND_TRUNC; }

void svo_oracle_half_sample(const uint8_t* in, int w, int h, uint8_t* out, int mode)
{
  const int ow = w / 2, oh = h / 2;
  if (mode == SVO_ROUND_SSE2) {
    /* vision.cpp:20-45: _mm_avg_epu8 of the two rows (round half up), then
     * _mm_avg_epu16 of horizontal neighbours (round half up again). */
    for (int y = 0; y < oh; ++y) {
      const uint8_t* r0 = in + (size_t)(2 * y) * w;
      const uint8_t* r1 = r0 + w;
      uint8_t* o = out + (size_t)y * ow;
      for (int x = 0; x < ow; ++x) {
        const int a = (r0[2 * x] + r1[2 * x] + 1) >> 1;
        const int b = (r0[2 * x + 1] + r1[2 * x + 1] + 1) >> 1;
        o[x] = (uint8_t)((a + b + 1) >> 1);
      }
    }
    return;
  }
  /* vision.cpp:92-109 scalar path, including its pointer walk: after a row of
   * out_width outputs `top` has advanced 2*out_width, then += stride.  For even
   * widths that is the next row pair; for odd widths it drifts by one pixel per
   * row (reproduced, not fixed). */
  const int stride = w;
  const uint8_t* top = in;
  const uint8_t* bottom = top + stride;
  const uint8_t* end = top + (size_t)stride * h;
  uint8_t* p = out;
  int rows_done = 0;
  while (bottom < end && rows_done < oh) { /* rows_done guard: never write past out (reference would) */
    for (int j = 0; j < ow; ++j) {
      *p = (uint8_t)(((uint16_t)top[0] + top[1] + bottom[0] + bottom[1]) / 4);
      p++; top += 2; bottom += 2;
    }
    top += stride; bottom += stride;
    ++rows_done;
  }
}

size_t svo_oracle_pyramid_bytes(int w, int h, int n_levels)
{
  size_t n = 0;
  for (int l = 1; l < n_levels; ++l) { w /= 2; h /= 2; n += (size_t)w * h; }
  return n;
}

void svo_oracle_build_pyramid(const uint8_t* img0, int w, int h, int n_levels, const int* modes, uint8_t* out)
{
  const uint8_t* in = img0;
  for (int l = 1; l < n_levels; ++l) {
    const int mode = modes ? modes[l - 1] : svo_oracle_half_sample_mode_x86(w);
    svo_oracle_half_sample(in, w, h, out, mode);
    in = out;
    w /= 2; h /= 2;
    out += (size_t)w * h;
  }
}

void svo_oracle_make_pyr(svo_pyr* p, const uint8_t* img0, const uint8_t* upper, int w, int h, int n_levels)
{
  memset(p, 0, sizeof(*p));
  p->n_levels = n_levels;
  p->data[0] = img0; p->w[0] = w; p->h[0] = h;
  for (int l = 1; l < n_levels; ++l) {
    w /= 2; h /= 2;
    p->data[l] = upper; p->w[l] = w; p->h[l] = h;
    upper += (size_t)w * h;
  }
}

/* ------------------------------------------------------------------ */
/* a3  cv::FAST (third party: OpenCV 4.5.4 features2d, fast.cpp +      */
/*     fast_score.cpp; not in /root/reference).  Restated from the     */
/*     published algorithm; single call site feature_detection.cpp:91. */
/* ------------------------------------------------------------------ */
static const int kRingDx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int kRingDy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

static int fast_is_corner(const int ring[25], int v, int t)
{
  int count = 0;
  const int vt = v - t;
  for (int k = 0; k < 25; ++k) {
    if (ring[k] < vt) { if (++count > 8) return 1; } else count = 0;
  }
  const int vt2 = v + t;
  count = 0;
  for (int k = 0; k < 25; ++k) {
    if (ring[k] > vt2) { if (++count > 8) return 1; } else count = 0;
  }
  return 0;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* cornerScore<16>: largest threshold for which the pixel is still a corner */
static int fast_corner_score(const int ring[25], int v, int threshold)
{
  int d[25];
  for (int k = 0; k < 25; ++k) d[k] = v - ring[k];
  int a0 = threshold;
  for (int k = 0; k < 16; k += 2) {
    int a = imin(d[k + 1], d[k + 2]);
    a = imin(a, d[k + 3]);
    if (a <= a0) continue;
    a = imin(a, d[k + 4]); a = imin(a, d[k + 5]); a = imin(a, d[k + 6]);
    a = imin(a, d[k + 7]); a = imin(a, d[k + 8]);
    a0 = imax(a0, imin(a, d[k]));
    a0 = imax(a0, imin(a, d[k + 9]));
  }
  int b0 = -a0;
  for (int k = 0; k < 16; k += 2) {
    int b = imax(d[k + 1], d[k + 2]);
    b = imax(b, d[k + 3]); b = imax(b, d[k + 4]); b = imax(b, d[k + 5]);
    if (b >= b0) continue;
    b = imax(b, d[k + 6]); b = imax(b, d[k + 7]); b = imax(b, d[k + 8]);
    b0 = imin(b0, imax(b, d[k]));
    b0 = imin(b0, imax(b, d[k + 9]));
  }
  return -b0 - 1;
}

int svo_oracle_fast(const uint8_t* img, int w, int h, int threshold, int nonmax, int cap, int* xs, int* ys, int* scores)
{
  if (w < 7 || h < 7) return 0;
  /* score map, 0 for non-corners and outside [3,rows-3)x[3,cols-3) */
  uint8_t* sc = (uint8_t*)calloc((size_t)w * h, 1);
  uint8_t* is = (uint8_t*)calloc((size_t)w * h, 1);
  if (threshold < 0) threshold = 0;
  if (threshold > 255) threshold = 255;
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      const uint8_t* p = img + (size_t)y * w + x;
      const int v = p[0];
      int ring[25];
      for (int k = 0; k < 16; ++k) ring[k] = p[kRingDy[k] * w + kRingDx[k]];
      for (int k = 16; k < 25; ++k) ring[k] = ring[k - 16];
      if (!fast_is_corner(ring, v, threshold)) continue;
      is[(size_t)y * w + x] = 1;
      sc[(size_t)y * w + x] = (uint8_t)fast_corner_score(ring, v, threshold);
    }
  int n = 0;
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      const size_t i = (size_t)y * w + x;
      if (!is[i]) continue;
      const int s = sc[i];
      if (nonmax) {
        if (!(s > sc[i - 1] && s > sc[i + 1] && s > sc[i - w - 1] && s > sc[i - w] && s > sc[i - w + 1] &&
              s > sc[i + w - 1] && s > sc[i + w] && s > sc[i + w + 1]))
          continue;
      }
      if (n < cap) { xs[n] = x; ys[n] = y; if (scores) scores[n] = s; }
      ++n;
    }
  free(sc); free(is);
  return n;
}

/* ------------------------------------------------------------------ */
/* a4  vk::shiTomasiScore (vision.cpp:113-154)                         */
/* ------------------------------------------------------------------ */
float svo_oracle_shi_tomasi(const uint8_t* img, int w, int h, int u, int v)
{
  float dXX = 0.0f, dYY = 0.0f, dXY = 0.0f;
  const int halfbox = 4, box = 8, box_area = 64;
  const int x_min = u - halfbox, x_max = u + halfbox, y_min = v - halfbox, y_max = v + halfbox;
  if (x_min < 1 || x_max >= w - 1 || y_min < 1 || y_max >= h - 1) return 0.0f;
  const int stride = w;
  for (int y = y_min; y < y_max; ++y) {
    const uint8_t* pl = img + stride * y + x_min - 1;
    const uint8_t* pr = img + stride * y + x_min + 1;
    const uint8_t* pt = img + stride * (y - 1) + x_min;
    const uint8_t* pb = img + stride * (y + 1) + x_min;
    for (int x = 0; x < box; ++x, ++pl, ++pr, ++pt, ++pb) {
      const float dx = (float)(*pr - *pl);
      const float dy = (float)(*pb - *pt);
      dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;
    }
  }
  /* `dXX / (2.0 * box_area)` is a double division stored back to float (:149-151) */
  dXX = (float)((double)dXX / (2.0 * box_area));
  dYY = (float)((double)dYY / (2.0 * box_area));
  dXY = (float)((double)dXY / (2.0 * box_area));
  /* :152 — the float sub-expressions stay float; vision.cpp sees only ::sqrt(double) (it includes
   * <cmath> via Eigen, not <math.h>), so the root and the subtraction are double. */
  const float tr = dXX + dYY;
  const float disc = tr * tr - 4 * (dXX * dYY - dXY * dXY);
  const double root = sqrt((double)disc);
  return (float)(0.5 * ((double)tr - root));
}

/* ------------------------------------------------------------------ */
/* a3  FastDetector::detect (feature_detection.cpp:77-122)             */
/* ------------------------------------------------------------------ */
int svo_oracle_fast_detect(const svo_pyr* pyr, int n_detect_levels, int cell, double thr,
                           const uint8_t* occupancy, svo_corner* cells)
{
  const int W = pyr->w[0], H = pyr->h[0];
  const int n_cols = (int)ceil((double)W / cell), n_rows = (int)ceil((double)H / cell);
  const int n_cells = n_cols * n_rows;
  for (int k = 0; k < n_cells; ++k) { cells[k].x = 0; cells[k].y = 0; cells[k].level = 0; cells[k].score = (float)thr; }
  for (int L = 0; L < n_detect_levels; ++L) {
    const int scale = 1 << L;
    const int w = pyr->w[L], h = pyr->h[L];
    const int cap = w * h;
    int* xs = (int*)malloc(sizeof(int) * (size_t)cap);
    int* ys = (int*)malloc(sizeof(int) * (size_t)cap);
    const int n = svo_oracle_fast(pyr->data[L], w, h, 10, 1, cap, xs, ys, NULL);
    for (int i = 0; i < n; ++i) {
      /* xy is a cv::Point2f: (xy.y*scale) is float*int -> float, /cell_size_ (int) -> float, then (int) */
      const float fx = (float)xs[i], fy = (float)ys[i];
      const int k = (int)((fy * (float)scale) / (float)cell) * n_cols + (int)((fx * (float)scale) / (float)cell);
      if (occupancy && occupancy[k]) continue;
      const float score = svo_oracle_shi_tomasi(pyr->data[L], w, h, xs[i], ys[i]);
      if (score > cells[k].score) {
        /* Corner(int x, int y, ...) from float xy.x*scale */
        cells[k].x = (int)(fx * (float)scale); cells[k].y = (int)(fy * (float)scale);
        cells[k].score = score; cells[k].level = L;
      }
    }
    free(xs); free(ys);
  }
  int n_feat = 0;
  for (int k = 0; k < n_cells; ++k) if ((double)cells[k].score > thr) ++n_feat;
  return n_feat;
}

/* ------------------------------------------------------------------ */
/* a18  SE3/SO3 (SE3.h, SO3.h) and pinhole camera                      */
/* ------------------------------------------------------------------ */
typedef struct { double x, y, z; } v3;
static inline v3 v3_add(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static inline v3 v3_neg(v3 a) { v3 r = { -a.x, -a.y, -a.z }; return r; }
static inline v3 v3_scale(double s, v3 a) { v3 r = { s * a.x, s * a.y, s * a.z }; return r; }
static inline v3 v3_cross(v3 a, v3 b) { v3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; return r; }

/* SO3::operator*(Point3d) SO3.h:509-520 */
static inline v3 q_rot(const double q[4], v3 p)
{
  const v3 qv = { q[0], q[1], q[2] };
  v3 uv = v3_cross(qv, p);
  uv = v3_add(uv, uv);
  return v3_add(v3_add(p, v3_scale(q[3], uv)), v3_cross(qv, uv));
}
/* SO3::operator*(SO3) SO3.h:496-503 */
static inline void q_mul(const double a[4], const double b[4], double o[4])
{
  const double x = a[0], y = a[1], z = a[2], w = a[3];
  o[0] = w * b[0] + x * b[3] + y * b[2] - z * b[1];
  o[1] = w * b[1] + y * b[3] + z * b[0] - x * b[2];
  o[2] = w * b[2] + z * b[3] + x * b[1] - y * b[0];
  o[3] = w * b[3] - x * b[0] - y * b[1] - z * b[2];
}

void svo_oracle_se3_transform(const double T[7], const double p[3], double out[3])
{
  const v3 pp = { p[0], p[1], p[2] };
  const v3 r = q_rot(T + 3, pp);
  out[0] = T[0] + r.x; out[1] = T[1] + r.y; out[2] = T[2] + r.z;   /* SE3.h:53-57 */
}
void svo_oracle_se3_mul(const double A[7], const double B[7], double out[7])
{
  /* SE3.h:45-49: SE3(rA*rB, tA + rA*tB) */
  double q[4]; q_mul(A + 3, B + 3, q);
  const v3 tb = { B[0], B[1], B[2] };
  const v3 r = q_rot(A + 3, tb);
  out[0] = A[0] + r.x; out[1] = A[1] + r.y; out[2] = A[2] + r.z;
  out[3] = q[0]; out[4] = q[1]; out[5] = q[2]; out[6] = q[3];
}
void svo_oracle_se3_inverse(const double A[7], double out[7])
{
  /* SE3.h:35-38 */
  const double qi[4] = { -A[3], -A[4], -A[5], A[6] };
  const v3 t = { A[0], A[1], A[2] };
  const v3 r = v3_neg(q_rot(qi, t));
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = qi[0]; out[4] = qi[1]; out[5] = qi[2]; out[6] = qi[3];
}
void svo_oracle_se3_exp(const double l[6], double out[7])
{
  /* SE3.h:153-182 */
  const v3 p = { l[0], l[1], l[2] };
  const v3 r = { l[3], l[4], l[5] };
  const double theta_sq = r.x * r.x + r.y * r.y + r.z * r.z;
  const double theta = sqrt(theta_sq);
  const double half_theta = 0.5 * theta;
  double imag_factor, real_factor;
  if (theta < 1e-10) {
    const double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real_factor = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    const double s = sin(half_theta);
    imag_factor = s / theta;
    real_factor = cos(half_theta);
  }
  const v3 rxp = v3_cross(r, p);
  const v3 rxrxp = v3_cross(r, rxp);
  const double c1 = (1 - cos(theta)) / theta_sq;
  const double c2 = (theta - sin(theta)) / (theta_sq * theta);
  const v3 t = v3_add(v3_add(p, v3_scale(c1, rxp)), v3_scale(c2, rxrxp));
  out[0] = t.x; out[1] = t.y; out[2] = t.z;
  out[3] = imag_factor * r.x; out[4] = imag_factor * r.y; out[5] = imag_factor * r.z; out[6] = real_factor;
}

/* SO3::getMatrix SO3.h:396-410 (row-major) */
static void q_matrix(const double q[4], double m[9])
{
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  m[0] = 1.0 - 2.0 * (y2 + z2); m[1] = 2.0 * (xy - wz);       m[2] = 2.0 * (xz + wy);
  m[3] = 2.0 * (xy + wz);       m[4] = 1.0 - 2.0 * (x2 + z2); m[5] = 2.0 * (yz - wx);
  m[6] = 2.0 * (xz - wy);       m[7] = 2.0 * (yz + wx);       m[8] = 1.0 - 2.0 * (x2 + y2);
}

/* Eigen Vector3d squaredNorm/dot: packet of (x0,x1) reduced first, then + x2 */
static inline double dot3(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline double norm3(const double a[3]) { return sqrt(dot3(a, a)); }

/* PinholeCamera::cam2world pinhole_camera.cpp:48-66 (no distortion), xyz.normalized() */
void svo_oracle_cam2world(const svo_cam* cam, double u, double v, double f[3])
{
  double xyz[3] = { (u - cam->cx) / cam->fx, (v - cam->cy) / cam->fy, 1.0 };
  const double z = dot3(xyz, xyz);
  if (z > 0.0) { const double n = sqrt(z); f[0] = xyz[0] / n; f[1] = xyz[1] / n; f[2] = xyz[2] / n; }
  else { f[0] = xyz[0]; f[1] = xyz[1]; f[2] = xyz[2]; }
}
static inline void world2cam_uv(const svo_cam* cam, const double uv[2], double px[2])
{
  px[0] = cam->fx * uv[0] + cam->cx;      /* pinhole_camera.cpp:83-87 */
  px[1] = cam->fy * uv[1] + cam->cy;
}
void svo_oracle_world2cam(const svo_cam* cam, const double xyz[3], double px[2])
{
  const double uv[2] = { xyz[0] / xyz[2], xyz[1] / xyz[2] };   /* project2d math_utils.h:104-107 */
  world2cam_uv(cam, uv, px);
}
/* AbstractCamera::isInFrame abstract_camera.h:58-72 */
static inline int in_frame(const svo_cam* cam, int x, int y, int boundary)
{ return x >= boundary && x < cam->width - boundary && y >= boundary && y < cam->height - boundary; }
static inline int in_frame_level(const svo_cam* cam, int x, int y, int boundary, int level)
{ return x >= boundary && x < cam->width / (1 << level) - boundary && y >= boundary && y < cam->height / (1 << level) - boundary; }

/* ------------------------------------------------------------------ */
/* a5-a8  sparse image alignment                                       */
/* ------------------------------------------------------------------ */

static double ldlt_halving_sum(const double* t, int m)
{
  if (m == 1) return t[0];
  const int h = m / 2;
  return ldlt_halving_sum(t, h) + ldlt_halving_sum(t + h, m - h);
}
static double ldlt_packet2_sum(const double* t, int m)
{
  const int np = m / 2;
  if (np == 0) return ldlt_halving_sum(t, m);
  double p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0};
  for (int i = 0; i < np; ++i) { p0[i] = t[2 * i]; p1[i] = t[2 * i + 1]; }
  double r = ldlt_halving_sum(p0, np) + ldlt_halving_sum(p1, np);
  if (m & 1) r = r + t[m - 1];
  return r;
}

/* Eigen LDLT (pivoted, lower) for a 6x6 SPD-ish matrix + solve; mirrors
 * Eigen/src/Cholesky/LDLT.h unblocked algorithm and the fixed-size solve's associations. */
static void ldlt6_solve(const double Hin[36], const double b[6], double x[6])
{
  double A[36]; memcpy(A, Hin, sizeof(A));
  int perm[6];
  const int n = 6;
  for (int k = 0; k < n; ++k) {
    /* pivot: largest |diag| in the remaining block */
    int piv = k; double big = fabs(A[k * 6 + k]);
    for (int i = k + 1; i < n; ++i) { const double v = fabs(A[i * 6 + i]); if (v > big) { big = v; piv = i; } }
    perm[k] = piv;
    if (piv != k) {
      /* symmetric swap of rows/cols k and piv, lower triangle only */
      const int s = n - piv - 1;
      for (int j = 0; j < k; ++j) { const double t = A[k * 6 + j]; A[k * 6 + j] = A[piv * 6 + j]; A[piv * 6 + j] = t; }
      for (int j = 0; j < s; ++j) { const double t = A[(piv + 1 + j) * 6 + k]; A[(piv + 1 + j) * 6 + k] = A[(piv + 1 + j) * 6 + piv]; A[(piv + 1 + j) * 6 + piv] = t; }
      { const double t = A[k * 6 + k]; A[k * 6 + k] = A[piv * 6 + piv]; A[piv * 6 + piv] = t; }
      for (int i = k + 1; i < piv; ++i) { const double t = A[i * 6 + k]; A[i * 6 + k] = A[piv * 6 + i]; A[piv * 6 + i] = t; }
    }
    const int rs = n - k - 1;
    if (k > 0) {
      double temp[6];
      for (int j = 0; j < k; ++j) temp[j] = A[j * 6 + j] * A[k * 6 + j];
      double s = 0; for (int j = 0; j < k; ++j) s += A[k * 6 + j] * temp[j];
      A[k * 6 + k] -= s;
      for (int i = 0; i < rs; ++i) {
        double t = 0; for (int j = 0; j < k; ++j) t += A[(k + 1 + i) * 6 + j] * temp[j];
        A[(k + 1 + i) * 6 + k] -= t;
      }
    }
    const double d = A[k * 6 + k];
    if (rs > 0 && fabs(d) > 2.2250738585072014e-308)
      for (int i = 0; i < rs; ++i) A[(k + 1 + i) * 6 + k] /= d;
  }
  /* solve: x = P^T L^-T D^-1 L^-1 P b */
  double y[6]; memcpy(y, b, sizeof(y));
  for (int k = 0; k < n; ++k) if (perm[k] != k) { const double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
  /* Eigen's unrolled triangular solves on a fixed-size rhs subtract ONE reduced sum per row: halving association for the
   * lower solve (strided rows, scalar redux), Packet2d association for the upper solve (rows of the adjoint view are
   * contiguous).  Bit-identical to A.ldlt().solve(b) of the x86-64 SSE2 build (see svo_oracle_map.c). */
  for (int i = 1; i < n; ++i) { double t[6]; for (int j = 0; j < i; ++j) t[j] = A[i * 6 + j] * y[j]; y[i] -= ldlt_halving_sum(t, i); }
  for (int i = 0; i < n; ++i) { const double d = A[i * 6 + i]; y[i] = (fabs(d) > 2.2250738585072014e-308) ? y[i] / d : 0.0; }
  for (int i = n - 2; i >= 0; --i) { double t[6]; const int m = n - 1 - i; for (int j = 0; j < m; ++j) t[j] = A[(i + 1 + j) * 6 + i] * y[i + 1 + j]; y[i] -= ldlt_packet2_sum(t, m); }
  for (int k = n - 1; k >= 0; --k) if (perm[k] != k) { const double t = y[k]; y[k] = y[perm[k]]; y[perm[k]] = t; }
  memcpy(x, y, sizeof(y));
}

typedef struct {
  const svo_pyr *ref, *cur; const svo_cam* cam; int N;
  const double *px, *xyz_ref; const uint8_t* has_point;
  float* ref_patch_cache;      /* N x 16, never cleared between levels (sparse_img_align.cpp:66) */
  double* jacobian_cache;      /* N x 16 x 6, zeroed per level (:76) */
  uint8_t* visible;            /* sticky across levels (:67) */
  int level, have_ref_patch_cache;
  double H[36], Jres[6], x[6];
  size_t n_meas;
} align_state;

/* sparse_img_align.cpp:105-178 */
static void precompute_reference_patches(align_state* s)
{
  const int border = 3;
  const uint8_t* ref_img = s->ref->data[s->level];
  const int cols = s->ref->w[s->level], rows = s->ref->h[s->level];
  const int stride = cols;
  const float scale = 1.0f / (1 << s->level);
  const double focal_length = fabs(s->cam->fx);
  for (int i = 0; i < s->N; ++i) {
    /* px[0]*scale: double * float -> double, stored to float */
    const float u_ref = (float)(s->px[2 * i] * (double)scale);
    const float v_ref = (float)(s->px[2 * i + 1] * (double)scale);
    const int u_ref_i = (int)floorf(u_ref), v_ref_i = (int)floorf(v_ref);
    if (!s->has_point[i] || u_ref_i - border < 0 || v_ref_i - border < 0 || u_ref_i + border >= cols || v_ref_i + border >= rows)
      continue;
    s->visible[i] = 1;
    /* Frame::jacobian_xyz2uv frame.h:110-132 */
    const double x = s->xyz_ref[3 * i], y = s->xyz_ref[3 * i + 1], z = s->xyz_ref[3 * i + 2];
    const double z_inv = 1. / z, z_inv_2 = z_inv * z_inv;
    double J0[6], J1[6];
    J0[0] = -z_inv; J0[1] = 0.0; J0[2] = x * z_inv_2; J0[3] = y * J0[2]; J0[4] = -(1.0 + x * J0[2]); J0[5] = y * z_inv;
    J1[0] = 0.0; J1[1] = -z_inv; J1[2] = y * z_inv_2; J1[3] = 1.0 + y * J1[2]; J1[4] = -J0[3]; J1[5] = -x * z_inv;
    const float su = u_ref - u_ref_i, sv = v_ref - v_ref_i;
    const float w_tl = (float)((1.0 - su) * (1.0 - sv));
    const float w_tr = (float)(su * (1.0 - sv));
    const float w_bl = (float)((1.0 - su) * sv);
    const float w_br = su * sv;
    float* cache = s->ref_patch_cache + 16 * i;
    double* jc = s->jacobian_cache + (size_t)96 * i;
    const double fl = focal_length / (1 << s->level);
    int pc = 0;
    for (int yy = 0; yy < 4; ++yy) {
      const uint8_t* p = ref_img + (v_ref_i + yy - 2) * stride + (u_ref_i - 2);
      for (int xx = 0; xx < 4; ++xx, ++p, ++pc) {
        cache[pc] = w_tl * p[0] + w_tr * p[1] + w_bl * p[stride] + w_br * p[stride + 1];
        const float dx = 0.5f * ((w_tl * p[1] + w_tr * p[2] + w_bl * p[stride + 1] + w_br * p[stride + 2])
                                 - (w_tl * p[-1] + w_tr * p[0] + w_bl * p[stride - 1] + w_br * p[stride]));
        const float dy = 0.5f * ((w_tl * p[stride] + w_tr * p[1 + stride] + w_bl * p[stride * 2] + w_br * p[stride * 2 + 1])
                                 - (w_tl * p[-stride] + w_tr * p[1 - stride] + w_bl * p[0] + w_br * p[1]));
        for (int k = 0; k < 6; ++k) jc[pc * 6 + k] = ((double)dx * J0[k] + (double)dy * J1[k]) * fl;
      }
    }
  }
  s->have_ref_patch_cache = 1;
}

/* sparse_img_align.cpp:184-286; returns chi2/n_meas_ */
static double compute_residuals(align_state* s, const double T[7], int linearize)
{
  const uint8_t* cur_img = s->cur->data[s->level];
  const int cols = s->cur->w[s->level], rows = s->cur->h[s->level];
  if (!s->have_ref_patch_cache) precompute_reference_patches(s);
  const int stride = cols, border = 3;
  const float scale = 1.0f / (1 << s->level);
  float chi2 = 0.0f;
  for (int i = 0; i < s->N; ++i) {
    if (!s->visible[i]) continue;
    double xyz_cur[3], pxd[2];
    svo_oracle_se3_transform(T, s->xyz_ref + 3 * i, xyz_cur);
    svo_oracle_world2cam(s->cam, xyz_cur, pxd);
    /* (Vector2d.cast<float>() * scale) */
    const float u_cur = (float)pxd[0] * scale, v_cur = (float)pxd[1] * scale;
    const int u_i = (int)floorf(u_cur), v_i = (int)floorf(v_cur);
    if (u_i < 0 || v_i < 0 || u_i - border < 0 || v_i - border < 0 || u_i + border >= cols || v_i + border >= rows) continue;
    const float su = u_cur - u_i, sv = v_cur - v_i;
    const float w_tl = (float)((1.0 - su) * (1.0 - sv));
    const float w_tr = (float)(su * (1.0 - sv));
    const float w_bl = (float)((1.0 - su) * sv);
    const float w_br = su * sv;
    const float* cache = s->ref_patch_cache + 16 * i;
    const double* jc = s->jacobian_cache + (size_t)96 * i;
    int pc = 0;
    for (int yy = 0; yy < 4; ++yy) {
      const uint8_t* p = cur_img + (v_i + yy - 2) * stride + (u_i - 2);
      for (int xx = 0; xx < 4; ++xx, ++p, ++pc) {
        const float intensity = w_tl * p[0] + w_tr * p[1] + w_bl * p[stride] + w_br * p[stride + 1];
        const float res = intensity - cache[pc];
        const float weight = 1.0f;
        chi2 += res * res * weight;
        s->n_meas++;
        if (linearize) {
          const double* J = jc + pc * 6;
          for (int a = 0; a < 6; ++a) {
            for (int b = 0; b < 6; ++b) s->H[a * 6 + b] += J[a] * J[b] * (double)weight;
            s->Jres[a] -= J[a] * (double)res * (double)weight;
          }
        }
      }
    }
  }
  /* float / size_t -> float division, returned as double */
  return (double)(chi2 / (float)s->n_meas);
}

int svo_oracle_sparse_align(const svo_pyr* ref, const svo_pyr* cur, const svo_cam* cam, int N,
                            const double* px, const double* xyz_ref, const uint8_t* has_point,
                            const double T_init[7], const svo_align_opts* opts, svo_align_result* res)
{
  memset(res, 0, sizeof(*res));
  memcpy(res->T_cur_ref, T_init, 7 * sizeof(double));
  if (N <= 0) return 0;                                           /* sparse_img_align.cpp:55-59 */
  align_state s; memset(&s, 0, sizeof(s));
  s.ref = ref; s.cur = cur; s.cam = cam; s.N = N; s.px = px; s.xyz_ref = xyz_ref; s.has_point = has_point;
  s.ref_patch_cache = (float*)calloc((size_t)N * 16, sizeof(float));
  s.jacobian_cache = (double*)calloc((size_t)N * 96, sizeof(double));
  s.visible = (uint8_t*)calloc((size_t)N, 1);
  double model[7]; memcpy(model, T_init, sizeof(model));
  /* NLLSSolver::reset nlls_solver_impl.hpp:299-309 */
  double chi2_ = 1e10; int stop_ = 0; const int n_iter = opts->n_iter;
  for (int level = opts->max_level; level >= opts->min_level; --level) {
    s.level = level;
    memset(s.jacobian_cache, 0, (size_t)N * 96 * sizeof(double));
    s.have_ref_patch_cache = 0;
    /* optimizeGaussNewton nlls_solver_impl.hpp:25-100 */
    double old_model[7]; memcpy(old_model, model, sizeof(model));
    for (int iter = 0; iter < n_iter; ++iter) {
      memset(s.H, 0, sizeof(s.H)); memset(s.Jres, 0, sizeof(s.Jres));
      s.n_meas = 0;
      const double new_chi2 = compute_residuals(&s, model, 1);
      if (level < SVO_MAX_LEVELS) res->iters[level]++;
      ldlt6_solve(s.H, s.Jres, s.x);
      if (isnan(s.x[0])) stop_ = 1;                               /* solve() sparse_img_align.cpp:291-297 */
      if ((iter > 0 && new_chi2 > chi2_) || stop_) {
        if (iter > 0 && fabs(new_chi2 - chi2_) <= 1e-4 * fabs(chi2_)) res->n_ambiguous++;
        memcpy(model, old_model, sizeof(model));                  /* rollback */
        break;
      }
      if (iter > 0 && fabs(new_chi2 - chi2_) <= 1e-4 * fabs(chi2_)) res->n_ambiguous++;
      /* update(): T_new = T_old * exp(-x) sparse_img_align.cpp:302-308 */
      double nx[6], E[7], new_model[7];
      for (int k = 0; k < 6; ++k) nx[k] = -s.x[k];
      svo_oracle_se3_exp(nx, E);
      svo_oracle_se3_mul(model, E, new_model);
      memcpy(old_model, model, sizeof(model));
      memcpy(model, new_model, sizeof(model));
      chi2_ = new_chi2;
      double nm = 0; for (int k = 0; k < 6; ++k) { const double a = fabs(s.x[k]); if (a > nm) nm = a; }
      if (nm <= opts->eps) break;
    }
  }
  memcpy(res->T_cur_ref, model, sizeof(model));
  memcpy(res->H, s.H, sizeof(s.H)); memcpy(res->Jres, s.Jres, sizeof(s.Jres)); memcpy(res->x, s.x, sizeof(s.x));
  res->chi2 = chi2_; res->n_meas = (int)s.n_meas; res->stop = stop_;
  free(s.ref_patch_cache); free(s.jacobian_cache); free(s.visible);
  return (int)(s.n_meas / 16);
}

/* ------------------------------------------------------------------ */
/* a9/a10  feature_alignment float paths                               */
/* ------------------------------------------------------------------ */

/* Eigen Matrix3f::inverse() (Eigen/src/LU/InverseImpl.h, size-3 cofactor path) */
static void inv3f(const float m[9] /* row-major */, float r[9])
{
#define M(i, j) m[(i) * 3 + (j)]
#define COF(i, j) (M(((i) + 1) % 3, ((j) + 1) % 3) * M(((i) + 2) % 3, ((j) + 2) % 3) - M(((i) + 1) % 3, ((j) + 2) % 3) * M(((i) + 2) % 3, ((j) + 1) % 3))
  const float c00 = COF(0, 0), c10 = COF(1, 0), c20 = COF(2, 0);
  /* det = (cofactors_col0 .* matrix.col(0)).sum(): x0 + (x1 + x2) */
  const float det = c00 * M(0, 0) + (c10 * M(1, 0) + c20 * M(2, 0));
  const float invdet = 1.0f / det;
  r[1 * 3 + 0] = COF(0, 1) * invdet;
  r[1 * 3 + 1] = COF(1, 1) * invdet;
  r[2 * 3 + 0] = COF(0, 2) * invdet;
  r[1 * 3 + 2] = COF(2, 1) * invdet;
  r[2 * 3 + 1] = COF(1, 2) * invdet;
  r[2 * 3 + 2] = COF(2, 2) * invdet;
  r[0] = c00 * invdet; r[1] = c10 * invdet; r[2] = c20 * invdet;
#undef COF
#undef M
}

/* feature_alignment.cpp:154-282 (float path; the one an x86 host build runs) */
int svo_oracle_align2d(const uint8_t* cur_img, int cols, int rows, const uint8_t* pwb, const uint8_t* ref_patch,
                       int n_iter, double px[2])
{
  const int halfpatch = 4, patch_size = 8;
  int converged = 0;
  float dxs[64], dys[64];
  float H[9] = { 0 };
  const int ref_step = patch_size + 2;
  int idx = 0;
  for (int y = 0; y < patch_size; ++y) {
    const uint8_t* it = pwb + (y + 1) * ref_step + 1;
    for (int x = 0; x < patch_size; ++x, ++it, ++idx) {
      /* J[0] = 0.5 * (int) : double product stored to float */
      float J[3];
      J[0] = (float)(0.5 * (it[1] - it[-1]));
      J[1] = (float)(0.5 * (it[ref_step] - it[-ref_step]));
      J[2] = 1.0f;
      dxs[idx] = J[0]; dys[idx] = J[1];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H[a * 3 + b] += J[a] * J[b];
    }
  }
  float Hinv[9]; inv3f(H, Hinv);
  float mean_diff = 0;
  float u = (float)px[0], v = (float)px[1];
  const float min_update_squared = (float)(0.5 * 0.5);
  const int cur_step = cols;
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = (int)floor((double)u), v_r = (int)floor((double)v);
    if (u_r < halfpatch || v_r < halfpatch || u_r >= cols - halfpatch || v_r >= rows - halfpatch) break;
    if (isnan(u) || isnan(v)) return 0;
    const float sx = u - u_r, sy = v - v_r;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy));
    const float wTR = (float)(sx * (1.0 - sy));
    const float wBL = (float)((1.0 - sx) * sy);
    const float wBR = sx * sy;
    float Jres[3] = { 0, 0, 0 };
    idx = 0;
    for (int y = 0; y < patch_size; ++y) {
      const uint8_t* it = cur_img + (v_r + y - halfpatch) * cur_step + u_r - halfpatch;
      for (int x = 0; x < patch_size; ++x, ++it, ++idx) {
        const float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[cur_step] + wBR * it[cur_step + 1];
        const float res = search_pixel - ref_patch[idx] + mean_diff;
        Jres[0] -= res * dxs[idx];
        Jres[1] -= res * dys[idx];
        Jres[2] -= res;
      }
    }
    /* update = Hinv * Jres : Eigen lazy 3x3*3x1 product, row sum associates x0 + (x1 + x2) */
    float upd[3];
    for (int a = 0; a < 3; ++a) upd[a] = Hinv[a * 3 + 0] * Jres[0] + (Hinv[a * 3 + 1] * Jres[1] + Hinv[a * 3 + 2] * Jres[2]);
    u += upd[0]; v += upd[1]; mean_diff += upd[2];
    if (upd[0] * upd[0] + upd[1] * upd[1] < min_update_squared) { converged = 1; break; }
  }
  px[0] = u; px[1] = v;
  return converged;
}

/* feature_alignment.cpp:35-152 */
int svo_oracle_align1d(const uint8_t* cur_img, int cols, int rows, const float dir[2], const uint8_t* pwb,
                       const uint8_t* ref_patch, int n_iter, double px[2], double* h_inv)
{
  const int halfpatch = 4, patch_size = 8;
  int converged = 0;
  float dvs[64];
  float H[4] = { 0 };
  const int ref_step = patch_size + 2;
  int idx = 0;
  for (int y = 0; y < patch_size; ++y) {
    const uint8_t* it = pwb + (y + 1) * ref_step + 1;
    for (int x = 0; x < patch_size; ++x, ++it, ++idx) {
      /* 0.5*(dir[0]*(int) + dir[1]*(int)): float sum promoted to double by 0.5, stored to float */
      float J[2];
      J[0] = (float)(0.5 * (double)(dir[0] * (float)(it[1] - it[-1]) + dir[1] * (float)(it[ref_step] - it[-ref_step])));
      J[1] = 1.0f;
      dvs[idx] = J[0];
      for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) H[a * 2 + b] += J[a] * J[b];
    }
  }
  *h_inv = 1.0 / H[0] * patch_size * patch_size;                  /* :68, double */
  /* Matrix2f::inverse(): invdet = 1/det; [d -b; -c a]*invdet */
  const float det = H[0] * H[3] - H[2] * H[1];
  const float invdet = 1.0f / det;
  const float Hinv[4] = { H[3] * invdet, -H[1] * invdet, -H[2] * invdet, H[0] * invdet };
  float mean_diff = 0;
  float u = (float)px[0], v = (float)px[1];
  const float min_update_squared = (float)(0.03 * 0.03);
  const int cur_step = cols;
  float chi2 = 0;
  float upd[2] = { 0, 0 };
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = (int)floor((double)u), v_r = (int)floor((double)v);
    if (u_r < halfpatch || v_r < halfpatch || u_r >= cols - halfpatch || v_r >= rows - halfpatch) break;
    if (isnan(u) || isnan(v)) return 0;
    const float sx = u - u_r, sy = v - v_r;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy));
    const float wTR = (float)(sx * (1.0 - sy));
    const float wBL = (float)((1.0 - sx) * sy);
    const float wBR = sx * sy;
    float new_chi2 = 0.0f;
    float Jres[2] = { 0, 0 };
    idx = 0;
    for (int y = 0; y < patch_size; ++y) {
      const uint8_t* it = cur_img + (v_r + y - halfpatch) * cur_step + u_r - halfpatch;
      for (int x = 0; x < patch_size; ++x, ++it, ++idx) {
        const float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[cur_step] + wBR * it[cur_step + 1];
        const float res = search_pixel - ref_patch[idx] + mean_diff;
        Jres[0] -= res * dvs[idx];
        Jres[1] -= res;
        new_chi2 += res * res;
      }
    }
    if (iter > 0 && new_chi2 > chi2) { u -= upd[0]; v -= upd[1]; break; }   /* sic, :122-123 */
    chi2 = new_chi2;
    upd[0] = Hinv[0] * Jres[0] + Hinv[1] * Jres[1];
    upd[1] = Hinv[2] * Jres[0] + Hinv[3] * Jres[1];
    u += upd[0] * dir[0]; v += upd[0] * dir[1]; mean_diff += upd[1];
    if (upd[0] * upd[0] + upd[1] * upd[1] < min_update_squared) { converged = 1; break; }
  }
  px[0] = u; px[1] = v;
  return converged;
}

/* ------------------------------------------------------------------ */
/* a11/a14  warp + ZMSSD                                               */
/* ------------------------------------------------------------------ */

/* matcher.cpp:36-60 */
void svo_oracle_warp_matrix_affine(const svo_cam* cam_ref, const svo_cam* cam_cur, const double px_ref[2],
                                   const double f_ref[3], double depth_ref, const double T_cur_ref[7],
                                   int level_ref, double A[4])
{
  const int halfpatch_size = 5;
  const double xyz_ref[3] = { f_ref[0] * depth_ref, f_ref[1] * depth_ref, f_ref[2] * depth_ref };
  double du[3], dv[3];
  /* px_ref + Vector2d(halfpatch_size,0)*(1<<level_ref) */
  svo_oracle_cam2world(cam_ref, px_ref[0] + (double)halfpatch_size * (1 << level_ref), px_ref[1] + 0.0 * (1 << level_ref), du);
  svo_oracle_cam2world(cam_ref, px_ref[0] + 0.0 * (1 << level_ref), px_ref[1] + (double)halfpatch_size * (1 << level_ref), dv);
  const double su = xyz_ref[2] / du[2], sv = xyz_ref[2] / dv[2];
  for (int k = 0; k < 3; ++k) { du[k] *= su; dv[k] *= sv; }
  double t[3], px_cur[2], px_du[2], px_dv[2];
  svo_oracle_se3_transform(T_cur_ref, xyz_ref, t); svo_oracle_world2cam(cam_cur, t, px_cur);
  svo_oracle_se3_transform(T_cur_ref, du, t);      svo_oracle_world2cam(cam_cur, t, px_du);
  svo_oracle_se3_transform(T_cur_ref, dv, t);      svo_oracle_world2cam(cam_cur, t, px_dv);
  /* A.col(0) = (px_du - px_cur)/halfpatch_size  (Eigen: vector / int -> elementwise double division) */
  A[0] = (px_du[0] - px_cur[0]) / halfpatch_size; A[2] = (px_du[1] - px_cur[1]) / halfpatch_size;
  A[1] = (px_dv[0] - px_cur[0]) / halfpatch_size; A[3] = (px_dv[1] - px_cur[1]) / halfpatch_size;
}

/* matcher.cpp:65-78 */
int svo_oracle_best_search_level(const double A[4], int max_level)
{
  int search_level = 0;
  double D = A[0] * A[3] - A[2] * A[1];   /* Eigen 2x2 determinant: m00*m11 - m10*m01 */
  while (D > 3.0 && search_level < max_level) { search_level += 1; D *= 0.25; }
  return search_level;
}

/* vision.h:19-36 */
static inline float interpolate_8u(const uint8_t* img, int stride, float u, float v)
{
  const int x = (int)floor((double)u), y = (int)floor((double)v);
  const float sx = u - x, sy = v - y;
  const float w00 = (1.0f - sx) * (1.0f - sy);
  const float w01 = (1.0f - sx) * sy;
  const float w10 = sx * (1.0f - sy);
  const float w11 = 1.0f - w00 - w01 - w10;
  const uint8_t* p = img + y * stride + x;
  return w00 * p[0] + w01 * p[stride] + w10 * p[1] + w11 * p[stride + 1];
}

/* matcher.cpp:83-116 */
int svo_oracle_warp_affine(const double A[4], const uint8_t* img_ref, int cols, int rows, const double px_ref[2],
                           int level_ref, int search_level, int halfpatch_size, uint8_t* patch)
{
  const int patch_size = halfpatch_size * 2;
  /* A_cur_ref.inverse().cast<float>() — Eigen 2x2 inverse */
  const double det = A[0] * A[3] - A[2] * A[1];
  const double invdet = 1.0 / det;
  const float a00 = (float)(A[3] * invdet), a01 = (float)(-A[1] * invdet);
  const float a10 = (float)(-A[2] * invdet), a11 = (float)(A[0] * invdet);
  if (isnan(a00)) return 0;
  /* px_ref.cast<float>() / (1<<level_ref) */
  const float pr0 = (float)px_ref[0] / (float)(1 << level_ref), pr1 = (float)px_ref[1] / (float)(1 << level_ref);
  uint8_t* pp = patch;
  for (int y = 0; y < patch_size; ++y)
    for (int x = 0; x < patch_size; ++x, ++pp) {
      float p0 = (float)(x - halfpatch_size), p1 = (float)(y - halfpatch_size);
      p0 *= (float)(1 << search_level); p1 *= (float)(1 << search_level);
      const float qx = (a00 * p0 + a01 * p1) + pr0;
      const float qy = (a10 * p0 + a11 * p1) + pr1;
      if (qx < 0 || qy < 0 || qx >= cols - 1 || qy >= rows - 1) *pp = 0;
      else *pp = (uint8_t)interpolate_8u(img_ref, cols, qx, qy);
    }
  return 1;
}

/* matcher.cpp:138-147 */
void svo_oracle_patch_from_border(const uint8_t* pwb, uint8_t* patch)
{
  for (int y = 1; y < 9; ++y, patch += 8) { const uint8_t* b = pwb + y * 10 + 1; for (int x = 0; x < 8; ++x) patch[x] = b[x]; }
}

/* patch_score.h:40-220 (SSE2 path == scalar path, integer) */
int svo_oracle_zmssd(const uint8_t* ref_patch, const uint8_t* cur, int stride)
{
  uint32_t sumA = 0, sumAA = 0, sumB = 0, sumBB = 0, sumAB = 0;
  for (int r = 0; r < 64; ++r) { const uint32_t n = ref_patch[r]; sumA += n; sumAA += n * n; }
  for (int y = 0, r = 0; y < 8; ++y) {
    const uint8_t* p = cur + y * stride;
    for (int x = 0; x < 8; ++x, ++r) { const uint32_t c = p[x]; sumB += c; sumBB += c * c; sumAB += c * ref_patch[r]; }
  }
  const int a = (int)sumA, aa = (int)sumAA, b = (int)sumB, bb = (int)sumBB, ab = (int)sumAB;
  return aa - 2 * ab + bb - (a * a - 2 * a * b + b * b) / 64;
}

/* matcher.cpp:123-136 */
int svo_oracle_depth_from_triangulation(const double T[7], const double f_ref[3], const double f_cur[3], double* depth)
{
  double R[9]; q_matrix(T + 3, R);
  /* A.col(0) = R*f_ref.  Eigen 3.4/SSE2 evaluates the 3x1 result as one Packet2d (rows 0,1:
   * sequential (r0*f0 + r1*f1) + r2*f2) plus one scalar coefficient (row 2: novec redux
   * r0*f0 + (r1*f1 + r2*f2)).  Verified bit-for-bit against oracle/_ref. */
  double a0[3];
  for (int i = 0; i < 2; ++i) a0[i] = (R[i * 3 + 0] * f_ref[0] + R[i * 3 + 1] * f_ref[1]) + R[i * 3 + 2] * f_ref[2];
  a0[2] = R[6] * f_ref[0] + (R[7] * f_ref[1] + R[8] * f_ref[2]);
  const double* a1 = f_cur;
  const double m00 = dot3(a0, a0), m01 = dot3(a0, a1), m11 = dot3(a1, a1);
  const double det = m00 * m11 - m01 * m01;
  if (det < 0.000001) return 0;
  const double invdet = 1.0 / det;
  const double i00 = m11 * invdet, i01 = -m01 * invdet, i11 = m00 * invdet;
  /* depth2 = -(AtA^-1 * A^T) * t : evaluate (AtA^-1 * A^T) (2x3) then times t */
  const double t[3] = { T[0], T[1], T[2] };
  double row0[3];
  for (int k = 0; k < 3; ++k) row0[k] = i00 * a0[k] + i01 * a1[k];
  const double d0 = -((row0[0] * t[0] + row0[1] * t[1]) + row0[2] * t[2]);
  (void)i11;
  *depth = fabs(d0);
  return 1;
}

void svo_oracle_matcher_opts_default(svo_matcher_opts* o, int n_pyr_levels)
{
  o->align_1d = 0; o->align_max_iter = 10; o->max_epi_search_steps = 1000; o->subpix_refinement = 1;
  o->epi_search_edgelet_filtering = 1; o->epi_search_edgelet_max_angle = 0.7; o->max_search_level = n_pyr_levels - 1;
}

/* point.cpp:101-125 */
int svo_oracle_close_view_obs(const double framepos[3], const double pos[3], int n_obs, const double* obs_pos, int* best)
{
  double od[3] = { framepos[0] - pos[0], framepos[1] - pos[1], framepos[2] - pos[2] };
  { const double z = dot3(od, od); if (z > 0) { const double n = sqrt(z); od[0] /= n; od[1] /= n; od[2] /= n; } }
  int min_i = 0; double min_cos = 0;
  for (int i = 0; i < n_obs; ++i) {
    double d[3] = { obs_pos[3 * i] - pos[0], obs_pos[3 * i + 1] - pos[1], obs_pos[3 * i + 2] - pos[2] };
    const double z = dot3(d, d); if (z > 0) { const double n = sqrt(z); d[0] /= n; d[1] /= n; d[2] /= n; }
    const double c = dot3(od, d);
    if (c > min_cos) { min_cos = c; min_i = i; }
  }
  *best = min_i;
  return min_cos < 0.5 ? 0 : 1;
}

static void normalize2f(const float in[2], float out[2])
{
  const float z = in[0] * in[0] + in[1] * in[1];
  if (z > 0.0f) { const float n = sqrtf(z); out[0] = in[0] / n; out[1] = in[1] / n; } else { out[0] = in[0]; out[1] = in[1]; }
}

/* matcher.cpp:156-202 (after getCloseViewObs picked ref_ftr_) */
int svo_oracle_find_match_direct(const svo_pyr* ref, const svo_pyr* cur, const svo_cam* cam, const svo_ref_feature* f,
                                 double depth_ref, const double T_cur_ref[7], const svo_matcher_opts* o,
                                 const double px_cur_in[2], svo_match_result* r)
{
  memset(r, 0, sizeof(*r));
  r->px_cur[0] = px_cur_in[0]; r->px_cur[1] = px_cur_in[1];
  /* ref_ftr_->px.cast<int>()/(1<<level) , boundary halfpatch+2 */
  const int pxi = (int)f->px_ref[0] / (1 << f->level_ref), pyi = (int)f->px_ref[1] / (1 << f->level_ref);
  if (!in_frame_level(cam, pxi, pyi, 4 + 2, f->level_ref)) return 0;
  svo_oracle_warp_matrix_affine(cam, cam, f->px_ref, f->f_ref, depth_ref, T_cur_ref, f->level_ref, r->A_cur_ref);
  r->search_level = svo_oracle_best_search_level(r->A_cur_ref, o->max_search_level);
  svo_oracle_warp_affine(r->A_cur_ref, ref->data[f->level_ref], ref->w[f->level_ref], ref->h[f->level_ref], f->px_ref,
                         f->level_ref, r->search_level, 5, r->patch_with_border);
  svo_oracle_patch_from_border(r->patch_with_border, r->patch);
  double px_scaled[2] = { px_cur_in[0] / (1 << r->search_level), px_cur_in[1] / (1 << r->search_level) };
  int success;
  const int L = r->search_level;
  if (f->type == 1) {
    /* dir_cur = A*grad, normalize(), cast<float> */
    double d[2] = { r->A_cur_ref[0] * f->grad[0] + r->A_cur_ref[1] * f->grad[1], r->A_cur_ref[2] * f->grad[0] + r->A_cur_ref[3] * f->grad[1] };
    const double z = d[0] * d[0] + d[1] * d[1];
    if (z > 0) { const double n = sqrt(z); d[0] /= n; d[1] /= n; }
    const float df[2] = { (float)d[0], (float)d[1] };
    success = svo_oracle_align1d(cur->data[L], cur->w[L], cur->h[L], df, r->patch_with_border, r->patch, o->align_max_iter, px_scaled, &r->h_inv);
  } else {
    success = svo_oracle_align2d(cur->data[L], cur->w[L], cur->h[L], r->patch_with_border, r->patch, o->align_max_iter, px_scaled);
  }
  r->px_cur[0] = px_scaled[0] * (1 << L); r->px_cur[1] = px_scaled[1] * (1 << L);
  r->success = success;
  return success;
}

/* matcher.cpp:207-355 */
int svo_oracle_find_epipolar_match(const svo_pyr* ref, const svo_pyr* cur, const svo_cam* cam, const svo_ref_feature* f,
                                   const double T_cur_ref[7], double d_estimate, double d_min, double d_max,
                                   const svo_matcher_opts* o, svo_epi_result* r)
{
  memset(r, 0, sizeof(*r));
  int zmssd_best = 2000 * 64;
  r->zmssd_best = zmssd_best;
  double uv_best[2] = { 0, 0 };
  double p[3], t[3];
  for (int k = 0; k < 3; ++k) p[k] = f->f_ref[k] * d_min;
  svo_oracle_se3_transform(T_cur_ref, p, t);
  const double A[2] = { t[0] / t[2], t[1] / t[2] };
  for (int k = 0; k < 3; ++k) p[k] = f->f_ref[k] * d_max;
  svo_oracle_se3_transform(T_cur_ref, p, t);
  const double B[2] = { t[0] / t[2], t[1] / t[2] };
  const double epi_dir[2] = { A[0] - B[0], A[1] - B[1] };
  svo_oracle_warp_matrix_affine(cam, cam, f->px_ref, f->f_ref, d_estimate, T_cur_ref, f->level_ref, r->A_cur_ref);
  r->reject = 0;
  if (f->type == 1 && o->epi_search_edgelet_filtering) {
    double g[2] = { r->A_cur_ref[0] * f->grad[0] + r->A_cur_ref[1] * f->grad[1], r->A_cur_ref[2] * f->grad[0] + r->A_cur_ref[3] * f->grad[1] };
    { const double z = g[0] * g[0] + g[1] * g[1]; if (z > 0) { const double n = sqrt(z); g[0] /= n; g[1] /= n; } }
    double e[2] = { epi_dir[0], epi_dir[1] };
    { const double z = e[0] * e[0] + e[1] * e[1]; if (z > 0) { const double n = sqrt(z); e[0] /= n; e[1] /= n; } }
    const double cosangle = fabs(g[0] * e[0] + g[1] * e[1]);
    if (cosangle < o->epi_search_edgelet_max_angle) { r->reject = 1; return 0; }
  }
  r->search_level = svo_oracle_best_search_level(r->A_cur_ref, o->max_search_level);
  const int L = r->search_level;
  double px_A[2], px_B[2];
  world2cam_uv(cam, A, px_A); world2cam_uv(cam, B, px_B);
  { const double dx = px_A[0] - px_B[0], dy = px_A[1] - px_B[1]; r->epi_length = sqrt(dx * dx + dy * dy) / (1 << L); }
  svo_oracle_warp_affine(r->A_cur_ref, ref->data[f->level_ref], ref->w[f->level_ref], ref->h[f->level_ref], f->px_ref,
                         f->level_ref, L, 5, r->patch_with_border);
  svo_oracle_patch_from_border(r->patch_with_border, r->patch);
  const float dirf_raw[2] = { (float)(px_A[0] - px_B[0]), (float)(px_A[1] - px_B[1]) };
  float dirf[2]; normalize2f(dirf_raw, dirf);

  if (r->epi_length < 2.0) {
    r->px_cur[0] = (px_A[0] + px_B[0]) / 2.0; r->px_cur[1] = (px_A[1] + px_B[1]) / 2.0;
    double px_scaled[2] = { r->px_cur[0] / (1 << L), r->px_cur[1] / (1 << L) };
    int res;
    if (o->align_1d) res = svo_oracle_align1d(cur->data[L], cur->w[L], cur->h[L], dirf, r->patch_with_border, r->patch, o->align_max_iter, px_scaled, &r->h_inv);
    else res = svo_oracle_align2d(cur->data[L], cur->w[L], cur->h[L], r->patch_with_border, r->patch, o->align_max_iter, px_scaled);
    if (res) {
      r->px_cur[0] = px_scaled[0] * (1 << L); r->px_cur[1] = px_scaled[1] * (1 << L);
      double fc[3]; svo_oracle_cam2world(cam, r->px_cur[0], r->px_cur[1], fc);
      if (svo_oracle_depth_from_triangulation(T_cur_ref, f->f_ref, fc, &r->depth)) { r->success = 1; return 1; }
    }
    return 0;
  }

  size_t n_steps = (size_t)(r->epi_length / 0.7);
  const double step[2] = { epi_dir[0] / (double)n_steps, epi_dir[1] / (double)n_steps };
  r->n_steps = (int)n_steps;
  if (n_steps > (size_t)o->max_epi_search_steps) return 0;

  double uv[2] = { B[0] - step[0], B[1] - step[1] };
  int last_x = 0, last_y = 0;
  ++n_steps;
  const uint8_t* cimg = cur->data[L]; const int ccols = cur->w[L];
  for (size_t i = 0; i < n_steps; ++i, uv[0] += step[0], uv[1] += step[1]) {
    double px[2]; world2cam_uv(cam, uv, px);
    const int pxi = (int)(px[0] / (1 << L) + 0.5), pyi = (int)(px[1] / (1 << L) + 0.5);
    if (pxi == last_x && pyi == last_y) continue;
    last_x = pxi; last_y = pyi;
    if (!in_frame_level(cam, pxi, pyi, 8, L)) continue;
    const int z = svo_oracle_zmssd(r->patch, cimg + (pyi - 4) * ccols + (pxi - 4), ccols);
    r->n_evals++;
    if (z < zmssd_best) { zmssd_best = z; uv_best[0] = uv[0]; uv_best[1] = uv[1]; }
  }
  r->zmssd_best = zmssd_best;
  if (zmssd_best < 2000 * 64) {
    if (o->subpix_refinement) {
      world2cam_uv(cam, uv_best, r->px_cur);
      double px_scaled[2] = { r->px_cur[0] / (1 << L), r->px_cur[1] / (1 << L) };
      int res;
      if (o->align_1d) res = svo_oracle_align1d(cur->data[L], cur->w[L], cur->h[L], dirf, r->patch_with_border, r->patch, o->align_max_iter, px_scaled, &r->h_inv);
      else res = svo_oracle_align2d(cur->data[L], cur->w[L], cur->h[L], r->patch_with_border, r->patch, o->align_max_iter, px_scaled);
      if (res) {
        r->px_cur[0] = px_scaled[0] * (1 << L); r->px_cur[1] = px_scaled[1] * (1 << L);
        double fc[3]; svo_oracle_cam2world(cam, r->px_cur[0], r->px_cur[1], fc);
        if (svo_oracle_depth_from_triangulation(T_cur_ref, f->f_ref, fc, &r->depth)) { r->success = 1; return 1; }
      }
      return 0;
    }
    world2cam_uv(cam, uv_best, r->px_cur);
    double fc[3] = { uv_best[0], uv_best[1], 1.0 };
    { const double z = dot3(fc, fc); const double n = sqrt(z); fc[0] /= n; fc[1] /= n; fc[2] /= n; }
    if (svo_oracle_depth_from_triangulation(T_cur_ref, f->f_ref, fc, &r->depth)) { r->success = 1; return 1; }
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* a15-a17  depth filter                                               */
/* ------------------------------------------------------------------ */
void svo_oracle_seed_init(svo_seed* s, float depth_mean, float depth_min)
{
  s->a = 10; s->b = 10;
  s->mu = (float)(1.0 / depth_mean);
  s->z_range = (float)(1.0 / depth_min);
  s->sigma2 = s->z_range * s->z_range / 36;
}

static double normal_pdf(double x, double mean, double std_dev)
{
  static const double SQRT_2_PI = 1.41421356237309505;       /* sic, depth_filter.cpp:360 */
  const double exponent = -0.5 * pow((x - mean) / std_dev, 2);
  return (1 / (std_dev * SQRT_2_PI)) * exp(exponent);
}

/* depth_filter.cpp:368-391 — float variables, double intermediates where a `1.` literal appears */
void svo_oracle_update_seed(float x, float tau2, svo_seed* seed)
{
  const float norm_scale = sqrtf(seed->sigma2 + tau2);
  if (isnan(norm_scale)) return;
  const float s2 = (float)(1. / (1. / seed->sigma2 + 1. / tau2));
  const float m = s2 * (seed->mu / seed->sigma2 + x / tau2);
  float C1 = (float)((double)(seed->a / (seed->a + seed->b)) * normal_pdf(x, seed->mu, norm_scale));
  float C2 = (float)((double)(seed->b / (seed->a + seed->b)) * 1. / (double)seed->z_range);
  const float normalization_constant = C1 + C2;
  C1 /= normalization_constant;
  C2 /= normalization_constant;
  const float f = (float)((double)C1 * (seed->a + 1.) / (seed->a + seed->b + 1.) + (double)(C2 * seed->a) / (seed->a + seed->b + 1.));
  const float e = (float)((double)C1 * (seed->a + 1.) * (seed->a + 2.) / ((seed->a + seed->b + 1.) * (seed->a + seed->b + 2.))
                          + (double)(C2 * seed->a * (seed->a + 1.0f) / ((seed->a + seed->b + 1.0f) * (seed->a + seed->b + 2.0f))));
  const float mu_new = C1 * m + C2 * seed->mu;
  seed->sigma2 = C1 * (s2 + m * m) + C2 * (seed->sigma2 + seed->mu * seed->mu) - mu_new * mu_new;
  seed->mu = mu_new;
  seed->a = (e - f) / (f - e / f);
  seed->b = seed->a * (1.0f - f) / f;
}

/* depth_filter.cpp:396-416 */
double svo_oracle_compute_tau(const double T_ref_cur[7], const double f[3], double z, double px_error_angle)
{
  const double PI = 3.14159265;                                   /* svo::PI global.h:92 */
  const double t[3] = { T_ref_cur[0], T_ref_cur[1], T_ref_cur[2] };
  const double a[3] = { f[0] * z - t[0], f[1] * z - t[1], f[2] * z - t[2] };
  const double t_norm = norm3(t), a_norm = norm3(a);
  const double alpha = acos(dot3(f, t) / t_norm);
  const double nt[3] = { -t[0], -t[1], -t[2] };
  const double beta = acos(dot3(a, nt) / (t_norm * a_norm));
  const double beta_plus = beta + px_error_angle;
  const double gamma_plus = PI - alpha - beta_plus;
  const double z_plus = t_norm * sin(beta_plus) / sin(gamma_plus);
  return z_plus - z;
}

/* loop body of DepthFilter::updateSeeds depth_filter.cpp:250-340 */
int svo_oracle_update_seed_with_frame(const svo_pyr* ref, const svo_pyr* cur, const svo_cam* cam, const svo_ref_feature* f,
                                      const double T_ref_w[7], const double T_cur_w[7], const svo_matcher_opts* o,
                                      double conv_thresh, svo_seed* s, svo_epi_result* epi_out)
{
  svo_epi_result local; svo_epi_result* epi = epi_out ? epi_out : &local;
  memset(epi, 0, sizeof(*epi));
  double Tcw_inv[7], T_ref_cur[7], T_cur_ref[7];
  svo_oracle_se3_inverse(T_cur_w, Tcw_inv);
  svo_oracle_se3_mul(T_ref_w, Tcw_inv, T_ref_cur);                 /* :263 */
  svo_oracle_se3_inverse(T_ref_cur, T_cur_ref);
  const double inv_mu = 1.0 / s->mu;
  const double p[3] = { inv_mu * f->f_ref[0], inv_mu * f->f_ref[1], inv_mu * f->f_ref[2] };
  double xyz_f[3]; svo_oracle_se3_transform(T_cur_ref, p, xyz_f);
  if (xyz_f[2] < 0.0) return SVO_SEED_BEHIND;
  double pxf[2]; svo_oracle_world2cam(cam, xyz_f, pxf);
  if (!in_frame(cam, (int)pxf[0], (int)pxf[1], 0)) return SVO_SEED_NOT_IN_FRAME;
  const float z_inv_min = s->mu + sqrtf(s->sigma2);
  const float z_inv_max = fmaxf(s->mu - sqrtf(s->sigma2), 0.00000001f);
  /* findEpipolarMatchDirect recomputes T_cur_ref = cur.T_f_w_ * ref.T_f_w_.inverse() (matcher.cpp:216) */
  double Trw_inv[7], T_cur_ref_m[7];
  svo_oracle_se3_inverse(T_ref_w, Trw_inv);
  svo_oracle_se3_mul(T_cur_w, Trw_inv, T_cur_ref_m);
  if (!svo_oracle_find_epipolar_match(ref, cur, cam, f, T_cur_ref_m, 1.0 / s->mu, 1.0 / z_inv_min, 1.0 / z_inv_max, o, epi)) {
    s->b++;
    return SVO_SEED_NO_MATCH;
  }
  const double z = epi->depth;
  const double focal_length = fabs(cam->fx);
  const double px_noise = 1.0;
  const double px_error_angle = atan(px_noise / (2.0 * focal_length)) * 2.0;
  const double tau = svo_oracle_compute_tau(T_ref_cur, f->f_ref, z, px_error_angle);
  const double zmt = z - tau;
  const double tau_inverse = 0.5 * (1.0 / (0.0000001 > zmt ? 0.0000001 : zmt) - 1.0 / (z + tau));
  svo_oracle_update_seed((float)(1. / z), (float)(tau_inverse * tau_inverse), s);
  if ((double)sqrtf(s->sigma2) < (double)s->z_range / conv_thresh) return SVO_SEED_CONVERGED;
  if (isnan(z_inv_min)) return SVO_SEED_NAN_ERASED;
  return SVO_SEED_UPDATED;
}
