// matcher.cu — affine patch warp, 2D/1D Lucas-Kanade refinement, epipolar ZMSSD search and the
// depth-filter seed update.
//
// reference: warp::getWarpMatrixAffine / getBestSearchLevel / warpAffine (matcher.cpp:36-116),
// vk::interpolateMat_8u (vision.h:19-36), feature_alignment::align2D / align1D float paths
// (feature_alignment.cpp:35-282), vk::patch_score::ZMSSD<4> (patch_score.h:40-220),
// Matcher::findMatchDirect / findEpipolarMatchDirect (matcher.cpp:156-355),
// depthFromTriangulation (matcher.cpp:123-136), DepthFilter::updateSeeds loop body, updateSeed,
// computeTau (depth_filter.cpp:250-416).
//
// B200 design: one thread-block per seed (epipolar search / seed update) and one warp per
// reprojection candidate (findMatchDirect).  The 10x10 affine-warped reference patch is produced
// once into shared memory (100 lanes, one bilinear tap set each); the epipolar walk is evaluated
// thread-per-candidate with the 8x8 reference patch held in 16 registers and ZMSSD computed with
// byte dot products (dp4a) on funnel-shifted aligned words; the strict-minimum "first wins" rule
// is a 64-bit min over (score << 32 | step index).
//
// Parity: the warp, ZMSSD and the LK iterations are bit-identical to the reference.  The LK sums
// (H, Jres) are float accumulations whose order matters, so they are NOT tree-reduced: lanes
// compute the 64 per-pixel terms in parallel and 3-5 lanes each replay one sequential chain (a
// 64-long dependent FADD chain is ~256 cycles — cheaper than it sounds, and exact).  The epipolar
// sample positions come from the reference's running sum `uv += step`, replayed by two threads
// (x and y are independent chains).  Only libm calls (acos/sin/atan/exp) are tolerance-matched.
#include "common.cuh"
#include "kernels.h"

namespace {

// ---------------------------------------------------------------- warp (matcher.cpp:36-116)
__device__ inline void warp_matrix_affine(const DevCam& cam, const double* px_ref, v3d f_ref, double depth_ref,
                                          const double* T_cur_ref, int level_ref, double* A /*row-major*/)
{
  const int halfpatch_size = 5;
  const v3d xyz_ref = {f_ref.x * depth_ref, f_ref.y * depth_ref, f_ref.z * depth_ref};
  v3d du = cam2world(cam, px_ref[0] + (double)halfpatch_size * (1 << level_ref), px_ref[1] + 0.0 * (1 << level_ref));
  v3d dv = cam2world(cam, px_ref[0] + 0.0 * (1 << level_ref), px_ref[1] + (double)halfpatch_size * (1 << level_ref));
  const double su = xyz_ref.z / du.z, sv = xyz_ref.z / dv.z;
  du = {du.x * su, du.y * su, du.z * su};
  dv = {dv.x * sv, dv.y * sv, dv.z * sv};
  double pcx, pcy, pux, puy, pvx, pvy;
  world2cam(cam, se3_transform(T_cur_ref, xyz_ref), pcx, pcy);
  world2cam(cam, se3_transform(T_cur_ref, du), pux, puy);
  world2cam(cam, se3_transform(T_cur_ref, dv), pvx, pvy);
  A[0] = (pux - pcx) / halfpatch_size; A[2] = (puy - pcy) / halfpatch_size;
  A[1] = (pvx - pcx) / halfpatch_size; A[3] = (pvy - pcy) / halfpatch_size;
}

__device__ __forceinline__ int best_search_level(const double* A, int max_level)
{
  int search_level = 0;
  double D = A[0] * A[3] - A[2] * A[1];
  while (D > 3.0 && search_level < max_level) { search_level += 1; D *= 0.25; }
  return search_level;
}

// vision.h:19-36
__device__ __forceinline__ float interpolate_8u(const uint8_t* img, int stride, float u, float v)
{
  const int x = (int)floorf(u), y = (int)floorf(v);     // == floor((double)u) for a float argument
  const float sx = u - x, sy = v - y;
  const float w00 = (1.0f - sx) * (1.0f - sy);
  const float w01 = (1.0f - sx) * sy;
  const float w10 = sx * (1.0f - sy);
  const float w11 = 1.0f - w00 - w01 - w10;
  const uint8_t* p = img + (size_t)y * stride + x;
  return w00 * p[0] + w01 * p[stride] + w10 * p[1] + w11 * p[stride + 1];
}

// warpAffine with halfpatch 5 -> 10x10 patch; threads [t0, t0+nthreads) of the caller cooperate.
// Returns false (patch untouched) when the warp is NaN (matcher.cpp:94-98).
__device__ inline bool warp_affine_10x10(const double* A, const uint8_t* img, int pitch, int cols, int rows,
                                         const double* px_ref, int level_ref, int search_level,
                                         uint8_t* patch, int t, int nthreads)
{
  const double det = A[0] * A[3] - A[2] * A[1];
  const double invdet = 1.0 / det;
  const float a00 = (float)(A[3] * invdet), a01 = (float)(-A[1] * invdet);
  const float a10 = (float)(-A[2] * invdet), a11 = (float)(A[0] * invdet);
  if (isnan(a00)) return false;
  const float pr0 = (float)px_ref[0] / (float)(1 << level_ref), pr1 = (float)px_ref[1] / (float)(1 << level_ref);
  for (int i = t; i < 100; i += nthreads) {
    const int y = i / 10, x = i - y * 10;
    float p0 = (float)(x - 5), p1 = (float)(y - 5);
    p0 *= (float)(1 << search_level); p1 *= (float)(1 << search_level);
    const float qx = (a00 * p0 + a01 * p1) + pr0;
    const float qy = (a10 * p0 + a11 * p1) + pr1;
    uint8_t v = 0;
    if (!(qx < 0 || qy < 0 || qx >= cols - 1 || qy >= rows - 1)) v = (uint8_t)interpolate_8u(img, pitch, qx, qy);
    patch[i] = v;
  }
  return true;
}

// ---------------------------------------------------------------- feature alignment
struct AlignSmem {
  float dx[64], dy[64];
  float term[5][64];       // per-pixel terms of the sequential float sums, one row per chain
};

// Eigen Matrix3f::inverse(), cofactor path (Eigen/src/LU/InverseImpl.h)
__device__ __forceinline__ void inv3f(const float* m, float* r)
{
#define M_(i, j) m[(i) * 3 + (j)]
#define COF_(i, j) (M_(((i) + 1) % 3, ((j) + 1) % 3) * M_(((i) + 2) % 3, ((j) + 2) % 3) - M_(((i) + 1) % 3, ((j) + 2) % 3) * M_(((i) + 2) % 3, ((j) + 1) % 3))
  const float c00 = COF_(0, 0), c10 = COF_(1, 0), c20 = COF_(2, 0);
  const float det = c00 * M_(0, 0) + (c10 * M_(1, 0) + c20 * M_(2, 0));
  const float invdet = 1.0f / det;
  r[3] = COF_(0, 1) * invdet; r[4] = COF_(1, 1) * invdet; r[6] = COF_(0, 2) * invdet;
  r[5] = COF_(2, 1) * invdet; r[7] = COF_(1, 2) * invdet; r[8] = COF_(2, 2) * invdet;
  r[0] = c00 * invdet; r[1] = c10 * invdet; r[2] = c20 * invdet;
#undef COF_
#undef M_
}

// The reference accumulates its float sums pixel by pixel (`acc += term`, `acc -= term`); float
// addition is not associative, so a tree reduction would change the result.  Lanes 0..n_chains-1
// each replay ONE chain over the 64 staged terms with identical (non-divergent) code; a 64-long
// dependent FADD chain costs ~64 x 4 cycles.  sign = +1 adds, -1 subtracts.
__device__ __forceinline__ float chain64(const AlignSmem* S, int lane, int n_chains, float sign_mask_sub)
{
  const float* v = S->term[lane < n_chains ? lane : 0];
  float a = 0.f;
  if (sign_mask_sub != 0.f) {
#pragma unroll 16
    for (int i = 0; i < 64; ++i) a -= v[i];
  } else {
#pragma unroll 16
    for (int i = 0; i < 64; ++i) a += v[i];
  }
  return a;
}

// feature_alignment::align2D float path (feature_alignment.cpp:154-282); one full warp cooperates.
// px is at the scale of `img`.  Returns converged (uniform across the warp).
__device__ bool align2d_warp(const uint8_t* img, int pitch, int cols, int rows, const uint8_t* pwb, const uint8_t* ref_patch,
                             int n_iter, double* px, AlignSmem* S, int lane)
{
  // derivative of the template: 2 pixels per lane; H += J*J^T terms (H22 = 64 exactly)
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = lane + 32 * k, y = idx >> 3, x = idx & 7;
    const uint8_t* it = pwb + (y + 1) * 10 + 1 + x;
    const float j0 = (float)(0.5 * ((int)it[1] - (int)it[-1]));
    const float j1 = (float)(0.5 * ((int)it[10] - (int)it[-10]));
    S->dx[idx] = j0; S->dy[idx] = j1;
    S->term[0][idx] = j0 * j0;      // H00
    S->term[1][idx] = j0 * j1;      // H01
    S->term[2][idx] = j0;           // H02 (J[2] = 1)
    S->term[3][idx] = j1 * j1;      // H11
    S->term[4][idx] = j1;           // H12
  }
  __syncwarp();
  const float hv = chain64(S, lane, 5, 0.f);
  const float h00 = __shfl_sync(0xffffffffu, hv, 0), h01 = __shfl_sync(0xffffffffu, hv, 1), h02 = __shfl_sync(0xffffffffu, hv, 2);
  const float h11 = __shfl_sync(0xffffffffu, hv, 3), h12 = __shfl_sync(0xffffffffu, hv, 4);
  const float H[9] = {h00, h01, h02, h01, h11, h12, h02, h12, 64.0f};
  float Hinv[9];
  inv3f(H, Hinv);
  float mean_diff = 0;
  float u = (float)px[0], v = (float)px[1];
  const float min_update_squared = (float)(0.5 * 0.5);
  bool converged = false;
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = (int)floorf(u), v_r = (int)floorf(v);          // == floor((double)u) for a float argument
    if (u_r < 4 || v_r < 4 || u_r >= cols - 4 || v_r >= rows - 4) break;
    if (isnan(u) || isnan(v)) return false;
    const float sx = u - u_r, sy = v - v_r;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy));
    const float wTR = (float)(sx * (1.0 - sy));
    const float wBL = (float)((1.0 - sx) * sy);
    const float wBR = sx * sy;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int idx = lane + 32 * k, y = idx >> 3, x = idx & 7;
      const uint8_t* it = img + (size_t)(v_r + y - 4) * pitch + (u_r - 4 + x);
      const float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[pitch] + wBR * it[pitch + 1];
      const float res = search_pixel - ref_patch[idx] + mean_diff;
      S->term[0][idx] = res * S->dx[idx];
      S->term[1][idx] = res * S->dy[idx];
      S->term[2][idx] = res;
    }
    __syncwarp();
    const float jv = chain64(S, lane, 3, 1.f);
    const float J0 = __shfl_sync(0xffffffffu, jv, 0), J1 = __shfl_sync(0xffffffffu, jv, 1), J2 = __shfl_sync(0xffffffffu, jv, 2);
    // update = Hinv * Jres (Eigen lazy product: x0 + (x1 + x2))
    const float up0 = Hinv[0] * J0 + (Hinv[1] * J1 + Hinv[2] * J2);
    const float up1 = Hinv[3] * J0 + (Hinv[4] * J1 + Hinv[5] * J2);
    const float up2 = Hinv[6] * J0 + (Hinv[7] * J1 + Hinv[8] * J2);
    u += up0; v += up1; mean_diff += up2;
    if (up0 * up0 + up1 * up1 < min_update_squared) { converged = true; break; }
  }
  px[0] = u; px[1] = v;
  return converged;
}

// feature_alignment::align1D (feature_alignment.cpp:35-152); one full warp cooperates.
__device__ bool align1d_warp(const uint8_t* img, int pitch, int cols, int rows, float dir0, float dir1, const uint8_t* pwb,
                             const uint8_t* ref_patch, int n_iter, double* px, double* h_inv, AlignSmem* S, int lane)
{
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = lane + 32 * k, y = idx >> 3, x = idx & 7;
    const uint8_t* it = pwb + (y + 1) * 10 + 1 + x;
    const float j0 = (float)(0.5 * (double)(dir0 * (float)((int)it[1] - (int)it[-1]) + dir1 * (float)((int)it[10] - (int)it[-10])));
    S->dx[idx] = j0;
    S->term[0][idx] = j0 * j0;      // H00
    S->term[1][idx] = j0;           // H01 = H10
  }
  __syncwarp();
  const float hv = chain64(S, lane, 2, 0.f);
  const float h00 = __shfl_sync(0xffffffffu, hv, 0), h01 = __shfl_sync(0xffffffffu, hv, 1), h11 = 64.0f;
  *h_inv = 1.0 / h00 * 8 * 8;
  const float det = h00 * h11 - h01 * h01;
  const float invdet = 1.0f / det;
  const float Hinv[4] = {h11 * invdet, -h01 * invdet, -h01 * invdet, h00 * invdet};
  float mean_diff = 0;
  float u = (float)px[0], v = (float)px[1];
  const float min_update_squared = (float)(0.03 * 0.03);
  float chi2 = 0;
  float up0 = 0, up1 = 0;
  bool converged = false;
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = (int)floorf(u), v_r = (int)floorf(v);
    if (u_r < 4 || v_r < 4 || u_r >= cols - 4 || v_r >= rows - 4) break;
    if (isnan(u) || isnan(v)) return false;
    const float sx = u - u_r, sy = v - v_r;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy));
    const float wTR = (float)(sx * (1.0 - sy));
    const float wBL = (float)((1.0 - sx) * sy);
    const float wBR = sx * sy;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int idx = lane + 32 * k, y = idx >> 3, x = idx & 7;
      const uint8_t* it = img + (size_t)(v_r + y - 4) * pitch + (u_r - 4 + x);
      const float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[pitch] + wBR * it[pitch + 1];
      const float res = search_pixel - ref_patch[idx] + mean_diff;
      S->term[0][idx] = res * S->dx[idx];
      S->term[1][idx] = res;
      S->term[2][idx] = -(res * res);   // chain subtracts: 0 - (-(r*r)) ... == 0 + r*r exactly
    }
    __syncwarp();
    const float jv = chain64(S, lane, 3, 1.f);
    const float J0 = __shfl_sync(0xffffffffu, jv, 0), J1 = __shfl_sync(0xffffffffu, jv, 1), new_chi2 = __shfl_sync(0xffffffffu, jv, 2);
    if (iter > 0 && new_chi2 > chi2) { u -= up0; v -= up1; break; }   // sic (:122-123)
    chi2 = new_chi2;
    up0 = Hinv[0] * J0 + Hinv[1] * J1;
    up1 = Hinv[2] * J0 + Hinv[3] * J1;
    u += up0 * dir0; v += up0 * dir1; mean_diff += up1;
    if (up0 * up0 + up1 * up1 < min_update_squared) { converged = true; break; }
  }
  px[0] = u; px[1] = v;
  return converged;
}

// ---------------------------------------------------------------- ZMSSD (patch_score.h)
struct RefPatchRegs { uint32_t w[16]; int sumA, sumAA; };

__device__ __forceinline__ void load_ref_patch(const uint8_t* patch /*64 B, 4-aligned*/, RefPatchRegs& r)
{
  const uint32_t* p = reinterpret_cast<const uint32_t*>(patch);
  uint32_t sa = 0, saa = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    r.w[i] = p[i];
    sa = __dp4a(r.w[i], 0x01010101u, sa);
    saa = __dp4a(r.w[i], r.w[i], saa);
  }
  r.sumA = (int)sa; r.sumAA = (int)saa;
}

// cur points at the top-left pixel of the 8x8 window (any alignment)
__device__ __forceinline__ int zmssd_8x8(const RefPatchRegs& r, const uint8_t* cur, int pitch)
{
  uint32_t sb = 0, sbb = 0, sab = 0;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(cur);
  const unsigned sh = (unsigned)(a0 & 3) * 8;
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>((a0 & ~(uintptr_t)3) + (size_t)y * pitch);   // pitch % 4 == 0
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    sb = __dp4a(lo, 0x01010101u, sb); sb = __dp4a(hi, 0x01010101u, sb);
    sbb = __dp4a(lo, lo, sbb); sbb = __dp4a(hi, hi, sbb);
    sab = __dp4a(lo, r.w[2 * y], sab); sab = __dp4a(hi, r.w[2 * y + 1], sab);
  }
  const int sumB = (int)sb, sumBB = (int)sbb, sumAB = (int)sab;
  return r.sumAA - 2 * sumAB + sumBB - (r.sumA * r.sumA - 2 * r.sumA * sumB + sumB * sumB) / 64;
}

// matcher.cpp:123-136 (association of the 3-term sums as generated by Eigen 3.4 / SSE2)
__device__ inline bool depth_from_triangulation(const double* T, v3d f_ref, v3d f_cur, double* depth)
{
  double R[9];
  q_matrix(T + 3, R);
  v3d a0;
  a0.x = (R[0] * f_ref.x + R[1] * f_ref.y) + R[2] * f_ref.z;
  a0.y = (R[3] * f_ref.x + R[4] * f_ref.y) + R[5] * f_ref.z;
  a0.z = R[6] * f_ref.x + (R[7] * f_ref.y + R[8] * f_ref.z);
  const v3d a1 = f_cur;
  const double m00 = dot3(a0, a0), m01 = dot3(a0, a1), m11 = dot3(a1, a1);
  const double det = m00 * m11 - m01 * m01;
  if (det < 0.000001) return false;
  const double invdet = 1.0 / det;
  const double i00 = m11 * invdet, i01 = -m01 * invdet;
  const double r0 = i00 * a0.x + i01 * a1.x, r1 = i00 * a0.y + i01 * a1.y, r2 = i00 * a0.z + i01 * a1.z;
  const double d0 = -((r0 * T[0] + r1 * T[1]) + r2 * T[2]);
  *depth = fabs(d0);
  return true;
}

// ---------------------------------------------------------------- depth filter scalars
__device__ inline double normal_pdf(double x, double mean, double std_dev)
{
  const double SQRT_2_PI = 1.41421356237309505;          // sic, depth_filter.cpp:360
  const double q = (x - mean) / std_dev;
  const double exponent = -0.5 * (q * q);                // pow(q, 2)
  return (1 / (std_dev * SQRT_2_PI)) * exp(exponent);
}

// DepthFilter::updateSeed depth_filter.cpp:368-391 (float variables, double where a `1.` literal appears)
__device__ inline void update_seed(float x, float tau2, svob200_seed* seed)
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

// DepthFilter::computeTau depth_filter.cpp:396-416
__device__ inline double compute_tau(const double* T_ref_cur, v3d f, double z, double px_error_angle)
{
  const double PI_ = 3.14159265;                         // svo::PI global.h:92
  const v3d t = {T_ref_cur[0], T_ref_cur[1], T_ref_cur[2]};
  const v3d a = {f.x * z - t.x, f.y * z - t.y, f.z * z - t.z};
  const double t_norm = norm3(t), a_norm = norm3(a);
  const double alpha = acos(dot3(f, t) / t_norm);
  const double beta = acos(dot3(a, v3_neg(t)) / (t_norm * a_norm));
  const double beta_plus = beta + px_error_angle;
  const double gamma_plus = PI_ - alpha - beta_plus;
  const double z_plus = t_norm * sin(beta_plus) / sin(gamma_plus);
  return z_plus - z;
}

// ---------------------------------------------------------------- findEpipolarMatchDirect in three phases
// Phase 1 (per-thread double geometry), phase 2 (warp-cooperative patch warp / ZMSSD walk / LK),
// phase 3 (per-thread triangulation).  The depth filter runs them as three kernels — thread per
// seed, warp per seed, thread per seed — so the FP64 pipe (64 lanes/SM) never executes the same
// geometry redundantly across a CTA; the stand-alone epipolar query kernel runs them back to back
// in one warp.
constexpr int EPI_CHUNK = 256;           // epipolar samples staged per round
constexpr int EPI_MAX_STEPS = 1023;      // max_epi_search_steps is clamped to this
enum { EPI_MODE_NONE = 0, EPI_MODE_DIRECT = 1, EPI_MODE_WALK = 2 };

struct EpiGeom {
  double T_cur_ref[7];
  double A[4];                          // A_cur_ref, row-major
  double Bx0, By0, stepx, stepy;        // epipolar sample chain: uv_0 = B - step, uv_{i+1} = uv_i + step
  double px_mid[2];                     // (px_A + px_B) / 2
  double ex, ey;                        // epi_dir_ = A - B on the unit plane (matcher.cpp:224)
  double epi_length;
  float a00, a01, a10, a11, pr0, pr1;   // A_ref_cur (float) and px_ref at the reference level
  float dirx, diry;                     // (px_A - px_B).cast<float>().normalized()
  int warp_ok, L, mode, n, n_steps_report, reject;
};

struct EpiSearch {
  int found;                            // 0 none, 1 px_cur refined by LK, 2 uv_best only (no subpixel refinement)
  int zmssd_best, n_evals;
  double px_cur[2], uv_best[2], h_inv;
};

struct EpiWarpSmem {
  __align__(16) uint8_t pwb[100];
  __align__(16) uint8_t patch[64];
  AlignSmem al;
  short2 pxi[EPI_CHUNK + 1];
};

// matcher.cpp:207-288 up to the start of the walk: pure per-thread math
__device__ inline void epi_geometry(const DevCam& cam, const svob200_feature_ref& f, const double* T_cur_ref, double d_estimate,
                                    double d_min, double d_max, const svob200_matcher_opts& o, EpiGeom* g)
{
  for (int k = 0; k < 7; ++k) g->T_cur_ref[k] = T_cur_ref[k];
  g->reject = 0; g->mode = EPI_MODE_NONE; g->n = 0; g->n_steps_report = 0; g->L = 0; g->epi_length = 0; g->warp_ok = 0;
  g->Bx0 = g->By0 = g->stepx = g->stepy = 0; g->px_mid[0] = g->px_mid[1] = 0; g->ex = g->ey = 0;
  g->a00 = g->a01 = g->a10 = g->a11 = g->pr0 = g->pr1 = g->dirx = g->diry = 0;
  const v3d f_ref = {f.f[0], f.f[1], f.f[2]};
  const v3d tA = se3_transform(T_cur_ref, {f_ref.x * d_min, f_ref.y * d_min, f_ref.z * d_min});
  const double Ax = tA.x / tA.z, Ay = tA.y / tA.z;
  const v3d tB = se3_transform(T_cur_ref, {f_ref.x * d_max, f_ref.y * d_max, f_ref.z * d_max});
  const double Bx = tB.x / tB.z, By = tB.y / tB.z;
  const double ex = Ax - Bx, ey = Ay - By;
  g->ex = ex; g->ey = ey;
  warp_matrix_affine(cam, f.px, f_ref, d_estimate, T_cur_ref, f.level, g->A);
  const double* A = g->A;
  if (f.type == 1 && o.epi_search_edgelet_filtering) {
    double gx = A[0] * f.grad[0] + A[1] * f.grad[1], gy = A[2] * f.grad[0] + A[3] * f.grad[1];
    { const double z = gx * gx + gy * gy; if (z > 0) { const double n = sqrt(z); gx /= n; gy /= n; } }
    double nx = ex, ny = ey;
    { const double z = nx * nx + ny * ny; if (z > 0) { const double n = sqrt(z); nx /= n; ny /= n; } }
    const double cosangle = fabs(gx * nx + gy * ny);
    if (cosangle < o.epi_search_edgelet_max_angle) { g->reject = 1; return; }
  }
  const int L = best_search_level(A, o.max_search_level);
  g->L = L;
  double pAx, pAy, pBx, pBy;
  world2cam_uv(cam, Ax, Ay, pAx, pAy);
  world2cam_uv(cam, Bx, By, pBx, pBy);
  { const double dx = pAx - pBx, dy = pAy - pBy; g->epi_length = sqrt(dx * dx + dy * dy) / (1 << L); }
  // warpAffine prologue (matcher.cpp:92-102)
  {
    const double det = A[0] * A[3] - A[2] * A[1];
    const double invdet = 1.0 / det;
    g->a00 = (float)(A[3] * invdet); g->a01 = (float)(-A[1] * invdet);
    g->a10 = (float)(-A[2] * invdet); g->a11 = (float)(A[0] * invdet);
    g->warp_ok = isnan(g->a00) ? 0 : 1;
    g->pr0 = (float)f.px[0] / (float)(1 << f.level); g->pr1 = (float)f.px[1] / (float)(1 << f.level);
  }
  {
    float dx = (float)(pAx - pBx), dy = (float)(pAy - pBy);
    const float z = dx * dx + dy * dy;
    if (z > 0.0f) { const float n = sqrtf(z); dx /= n; dy /= n; }
    g->dirx = dx; g->diry = dy;
  }
  g->px_mid[0] = (pAx + pBx) / 2.0; g->px_mid[1] = (pAy + pBy) / 2.0;
  if (g->epi_length < 2.0) { g->mode = EPI_MODE_DIRECT; return; }
  // x86 (size_t)(double): NaN / out of range -> 2^63 -> "skip epipolar search" (matcher.cpp:283-288)
  const double q = g->epi_length / 0.7;
  if (!(q == q) || q >= 9.2e18) { g->n_steps_report = 0x7fffffff; return; }
  const unsigned long long n_steps = (unsigned long long)q;
  g->n_steps_report = (int)(n_steps > 0x7fffffffULL ? 0x7fffffffULL : n_steps);
  g->stepx = ex / (double)n_steps; g->stepy = ey / (double)n_steps;
  int max_steps = o.max_epi_search_steps;
  if (max_steps > EPI_MAX_STEPS) max_steps = EPI_MAX_STEPS;
  if (n_steps > (unsigned long long)max_steps) return;
  g->n = (int)n_steps + 1;
  g->Bx0 = Bx - g->stepx; g->By0 = By - g->stepy;
  g->mode = EPI_MODE_WALK;
}

// matcher.cpp:251-340: warp the patch, walk the epipolar segment, refine.  One full warp cooperates.
__device__ void epi_search_warp(const DevFrame& ref, int ref_image, const DevFrame& cur, int cur_image, const DevCam& cam,
                                const svob200_feature_ref& f, const EpiGeom& g, const svob200_matcher_opts& o, EpiWarpSmem* S,
                                int lane, bool always_warp, EpiSearch* out)
{
  out->found = 0; out->zmssd_best = 2000 * 64; out->n_evals = 0; out->px_cur[0] = out->px_cur[1] = 0;
  out->uv_best[0] = out->uv_best[1] = 0; out->h_inv = 0;
  if (g.reject) return;
  if (g.mode == EPI_MODE_NONE && !always_warp) return;
  const int L = g.L;
  if (g.warp_ok) {
    const uint8_t* rimg = ref.lvl[f.level] + (size_t)ref_image * ref.img_stride[f.level];
    const int rp = ref.pitch[f.level], rc = ref.w[f.level], rr = ref.h[f.level];
    for (int i = lane; i < 100; i += 32) {
      const int y = i / 10, x = i - y * 10;
      float p0 = (float)(x - 5), p1 = (float)(y - 5);
      p0 *= (float)(1 << L); p1 *= (float)(1 << L);
      const float qx = (g.a00 * p0 + g.a01 * p1) + g.pr0;
      const float qy = (g.a10 * p0 + g.a11 * p1) + g.pr1;
      uint8_t v = 0;
      if (!(qx < 0 || qy < 0 || qx >= rc - 1 || qy >= rr - 1)) v = (uint8_t)interpolate_8u(rimg, rp, qx, qy);
      S->pwb[i] = v;
    }
  }
  __syncwarp();
  for (int k = lane; k < 64; k += 32) S->patch[k] = S->pwb[((k >> 3) + 1) * 10 + 1 + (k & 7)];
  __syncwarp();
  if (g.mode == EPI_MODE_NONE) return;
  const uint8_t* cimg = cur.lvl[L] + (size_t)cur_image * cur.img_stride[L];
  const int cpitch = cur.pitch[L], ccols = cur.w[L], crows = cur.h[L];
  double px0, px1;
  if (g.mode == EPI_MODE_DIRECT) {
    px0 = g.px_mid[0]; px1 = g.px_mid[1];
    out->px_cur[0] = px0; out->px_cur[1] = px1;
  } else {
    RefPatchRegs rpatch;
    load_ref_patch(S->patch, rpatch);
    unsigned long long best = ((unsigned long long)(2000 * 64) << 32);   // PatchScore::threshold(), strict <
    int evals = 0;
    // the reference's running sums uv += step: x chain on lane 0, y chain on lane 1
    double uv = lane == 0 ? g.Bx0 : g.By0;
    const double st = lane == 0 ? g.stepx : g.stepy, fxy = lane == 0 ? cam.fx : cam.fy, cxy = lane == 0 ? cam.cx : cam.cy;
    const double inv_scale = 1.0 / (double)(1 << L);                     // exact: dividing by 2^L == multiplying by 2^-L
    short lastv = 0;                                                     // last_checked_pxi(0,0)
    short* dst = reinterpret_cast<short*>(S->pxi) + lane;
    for (int base = 0; base < g.n; base += EPI_CHUNK) {
      const int m = min(EPI_CHUNK, g.n - base);
      if (lane < 2) {
        dst[0] = lastv;
        for (int i = 0; i < m; ++i, uv += st) {
          const double p = fxy * uv + cxy;
          const double v = p * inv_scale + 0.5;
          const int iv = (v == v) ? (v >= 32767.0 ? 32767 : (v <= -32768.0 ? -32768 : (int)v)) : -32768;
          lastv = (short)iv;
          dst[2 * (i + 1)] = lastv;
        }
      }
      __syncwarp();
      for (int i = lane; i < m; i += 32) {
        const short2 c = S->pxi[i + 1], prev = S->pxi[i];
        if (c.x == prev.x && c.y == prev.y) continue;
        if (!in_frame_level(cam, c.x, c.y, 8, L)) continue;
        const int z = zmssd_8x8(rpatch, cimg + (size_t)(c.y - 4) * cpitch + (c.x - 4), cpitch);
        ++evals;
        const unsigned long long key = ((unsigned long long)(unsigned)z << 32) | (unsigned)(base + i);
        if (key < best) best = key;
      }
      __syncwarp();
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, off);
      if (other < best) best = other;
      evals += __shfl_xor_sync(0xffffffffu, evals, off);
    }
    out->n_evals = evals;
    const int zbest = (int)(best >> 32);
    out->zmssd_best = zbest;
    if (!(zbest < 2000 * 64)) return;
    // uv_best: replay the chain up to the winning step
    const int ibest = (int)(best & 0xffffffffu);
    double ub = lane == 0 ? g.Bx0 : g.By0;
    if (lane < 2) for (int i = 0; i < ibest; ++i) ub += st;
    const double ubx = __shfl_sync(0xffffffffu, ub, 0), uby = __shfl_sync(0xffffffffu, ub, 1);
    world2cam_uv(cam, ubx, uby, px0, px1);
    out->px_cur[0] = px0; out->px_cur[1] = px1;
    out->uv_best[0] = ubx; out->uv_best[1] = uby;
    if (!o.subpix_refinement) { out->found = 2; return; }
  }
  double pxs[2] = {px0 / (1 << L), px1 / (1 << L)};
  double h_inv = 0;
  bool res;
  if (o.align_1d) res = align1d_warp(cimg, cpitch, ccols, crows, g.dirx, g.diry, S->pwb, S->patch, o.align_max_iter, pxs, &h_inv, &S->al, lane);
  else res = align2d_warp(cimg, cpitch, ccols, crows, S->pwb, S->patch, o.align_max_iter, pxs, &S->al, lane);
  out->h_inv = h_inv;
  if (res) { out->px_cur[0] = pxs[0] * (1 << L); out->px_cur[1] = pxs[1] * (1 << L); out->found = 1; }
}

// matcher.cpp:269-276 / :341-351: triangulate from the matched pixel (per thread)
__device__ inline bool epi_finish(const DevCam& cam, const svob200_feature_ref& f, const double* T_cur_ref, const EpiSearch& s, double* depth)
{
  v3d fc;
  if (s.found == 1) fc = cam2world(cam, s.px_cur[0], s.px_cur[1]);
  else if (s.found == 2) fc = normalized3({s.uv_best[0], s.uv_best[1], 1.0});
  else return false;
  return depth_from_triangulation(T_cur_ref, {f.f[0], f.f[1], f.f[2]}, fc, depth);
}

// stand-alone epipolar query: one warp per query, the three phases back to back
struct EpiQuerySmem { EpiWarpSmem w; EpiGeom g; };

__global__ void __launch_bounds__(128) epipolar_kernel(const DevFrame* frames, const int* ref_slot, int cur_slot, DevCam cam,
                                                       int n, const svob200_feature_ref* ftrs, const double* d,
                                                       svob200_matcher_opts o, svob200_epi_result* results)
{
  __shared__ EpiQuerySmem SM[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 4 + warp;
  if (i >= n) return;
  EpiQuerySmem* S = &SM[warp];
  const svob200_feature_ref f = ftrs[i];
  for (int k = lane; k < 100; k += 32) S->w.pwb[k] = 0;
  if (lane == 0) epi_geometry(cam, f, f.T_cur_ref, d[3 * i], d[3 * i + 1], d[3 * i + 2], o, &S->g);
  __syncwarp();
  const EpiGeom g = S->g;
  EpiSearch sr;
  epi_search_warp(frames[(int)f.ref_frame_id], f.ref_image, frames[cur_slot], f.cur_image, cam, f, g, o, &S->w, lane, true, &sr);
  double depth = 0;
  const bool ok = epi_finish(cam, f, g.T_cur_ref, sr, &depth);
  __syncwarp();
  svob200_epi_result* R = &results[i];
  if (lane == 0) {
    R->success = ok ? 1 : 0; R->search_level = g.L; R->reject = g.reject; R->zmssd_best = sr.zmssd_best;
    R->n_evals = sr.n_evals; R->n_steps = g.n_steps_report; R->depth = ok ? depth : 0.0;
    R->px_cur[0] = sr.px_cur[0]; R->px_cur[1] = sr.px_cur[1];
    R->epi_length = g.epi_length; R->h_inv = sr.h_inv;
    R->epi_dir[0] = g.ex; R->epi_dir[1] = g.ey;
    R->px_cur_valid = (!g.reject && (g.mode == EPI_MODE_DIRECT || (g.mode == EPI_MODE_WALK && sr.zmssd_best < 2000 * 64))) ? 1 : 0;
    for (int k = 0; k < 4; ++k) R->A_cur_ref[k] = g.A[k];
  }
  for (int k = lane; k < 100; k += 32) R->patch_with_border[k] = g.reject ? 0 : S->w.pwb[k];
  for (int k = lane; k < 64; k += 32) R->patch[k] = g.reject ? 0 : S->w.patch[k];
}

// ---------------------------------------------------------------- depth filter: DepthFilter::updateSeeds loop body
// (depth_filter.cpp:250-340) as three kernels over all seeds
struct SeedPre {
  int status;                  // 0 = run the matcher; else the final status (BEHIND / NOT_IN_FRAME)
  float z_inv_min;
  double t_ref_cur[3];         // translation of T_ref_cur (computeTau)
};

// phase 1: thread per seed — visibility, inverse-depth range, epipolar geometry
__global__ void __launch_bounds__(128) seeds_geom_kernel(DevCam cam, int n, const svob200_feature_ref* ftrs, const double* T_ref_w_all,
                                                         const double* T_cur_w_all, svob200_matcher_opts o, const svob200_seed* seeds,
                                                         SeedPre* pre, EpiGeom* geom)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const svob200_feature_ref f = ftrs[i];
  const svob200_seed s = seeds[i];
  const double* T_ref_w = T_ref_w_all + 7 * (size_t)i;
  const double* T_cur_w = T_cur_w_all + 7 * (size_t)f.cur_image;
  double Tcw_inv[7], T_ref_cur[7], T_cur_ref[7];
  se3_inverse(T_cur_w, Tcw_inv);
  se3_mul(T_ref_w, Tcw_inv, T_ref_cur);                              // depth_filter.cpp:263
  se3_inverse(T_ref_cur, T_cur_ref);
  SeedPre p;
  p.status = 0; p.z_inv_min = 0.f;
  p.t_ref_cur[0] = T_ref_cur[0]; p.t_ref_cur[1] = T_ref_cur[1]; p.t_ref_cur[2] = T_ref_cur[2];
  const double inv_mu = 1.0 / s.mu;
  const v3d xyz_f = se3_transform(T_cur_ref, {inv_mu * f.f[0], inv_mu * f.f[1], inv_mu * f.f[2]});
  if (xyz_f.z < 0.0) p.status = SVOB200_SEED_BEHIND;
  else {
    double pxf, pyf;
    world2cam(cam, xyz_f, pxf, pyf);
    if (!in_frame(cam, (int)pxf, (int)pyf, 0)) p.status = SVOB200_SEED_NOT_IN_FRAME;
  }
  if (p.status == 0) {
    const float z_inv_min = s.mu + sqrtf(s.sigma2);
    const float z_inv_max = fmaxf(s.mu - sqrtf(s.sigma2), 0.00000001f);
    p.z_inv_min = z_inv_min;
    double Trw_inv[7], T_cur_ref_m[7];
    se3_inverse(T_ref_w, Trw_inv);
    se3_mul(T_cur_w, Trw_inv, T_cur_ref_m);                          // matcher.cpp:216
    epi_geometry(cam, f, T_cur_ref_m, 1.0 / s.mu, 1.0 / z_inv_min, 1.0 / z_inv_max, o, &geom[i]);
  }
  pre[i] = p;
}

// phase 2: warp per seed — patch warp, ZMSSD walk, LK refinement
__global__ void __launch_bounds__(128) seeds_search_kernel(const DevFrame* frames, int cur_slot, DevCam cam, int n,
                                                           const svob200_feature_ref* ftrs, svob200_matcher_opts o,
                                                           const SeedPre* pre, const EpiGeom* geom, EpiSearch* search)
{
  __shared__ EpiWarpSmem SM[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 4 + warp;
  if (i >= n) return;
  if (pre[i].status != 0) return;
  const EpiGeom g = geom[i];
  if (g.reject || g.mode == EPI_MODE_NONE) {
    if (lane == 0) { EpiSearch z; z.found = 0; z.zmssd_best = 2000 * 64; z.n_evals = 0; z.px_cur[0] = z.px_cur[1] = 0; z.uv_best[0] = z.uv_best[1] = 0; z.h_inv = 0; search[i] = z; }
    return;
  }
  const svob200_feature_ref f = ftrs[i];
  EpiSearch sr;
  epi_search_warp(frames[(int)f.ref_frame_id], f.ref_image, frames[cur_slot], f.cur_image, cam, f, g, o, &SM[warp], lane, false, &sr);
  if (lane == 0) search[i] = sr;
}

// phase 3: thread per seed — triangulation, tau, Gaussian x Beta update, status
__global__ void __launch_bounds__(128) seeds_finish_kernel(DevCam cam, int n, const svob200_feature_ref* ftrs, double conv_thresh,
                                                           const SeedPre* pre, const EpiGeom* geom, const EpiSearch* search,
                                                           svob200_seed* seeds, svob200_seed_obs* obs)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SeedPre p = pre[i];
  svob200_seed_obs ob;
  ob.status = p.status; ob.search_level = 0; ob.zmssd_best = 2000 * 64; ob.n_evals = 0; ob.z = 0; ob.px_cur[0] = ob.px_cur[1] = 0; ob.epi_length = 0;
  if (p.status == 0) {
    const svob200_feature_ref f = ftrs[i];
    const EpiSearch sr = search[i];
    const EpiGeom* g = &geom[i];
    svob200_seed s = seeds[i];
    ob.search_level = g->L; ob.zmssd_best = sr.zmssd_best; ob.n_evals = sr.n_evals; ob.epi_length = g->epi_length;
    ob.px_cur[0] = sr.px_cur[0]; ob.px_cur[1] = sr.px_cur[1];
    double T_cur_ref[7];
    for (int k = 0; k < 7; ++k) T_cur_ref[k] = g->T_cur_ref[k];
    double z = 0;
    if (!epi_finish(cam, f, T_cur_ref, sr, &z)) {
      s.b++;                                                         // depth_filter.cpp:286
      ob.status = SVOB200_SEED_NO_MATCH;
    } else {
      ob.z = z;
      const double focal_length = fabs(cam.fx);
      const double px_error_angle = atan(1.0 / (2.0 * focal_length)) * 2.0;
      const double T_ref_cur[7] = {p.t_ref_cur[0], p.t_ref_cur[1], p.t_ref_cur[2], 0, 0, 0, 1};
      const double tau = compute_tau(T_ref_cur, {f.f[0], f.f[1], f.f[2]}, z, px_error_angle);
      const double zmt = z - tau;
      const double tau_inverse = 0.5 * (1.0 / (0.0000001 > zmt ? 0.0000001 : zmt) - 1.0 / (z + tau));
      update_seed((float)(1. / z), (float)(tau_inverse * tau_inverse), &s);
      if ((double)sqrtf(s.sigma2) < (double)s.z_range / conv_thresh) ob.status = SVOB200_SEED_CONVERGED;
      else if (isnan(p.z_inv_min)) ob.status = SVOB200_SEED_NAN_ERASED;
      else ob.status = SVOB200_SEED_UPDATED;
    }
    seeds[i] = s;
  }
  obs[i] = ob;
}

// ---------------------------------------------------------------- findMatchDirect: one warp per candidate
struct MatchSmem { __align__(16) uint8_t pwb[100]; __align__(16) uint8_t patch[64]; AlignSmem al; };

__global__ void __launch_bounds__(128) match_direct_kernel(const DevFrame* frames, const int* ref_slot, int cur_slot, DevCam cam, int n,
                                                           const svob200_feature_ref* ftrs, const double* depth_ref,
                                                           const double* px_in, svob200_matcher_opts o, svob200_match_result* results,
                                                           double* px_out, int* ok_out)
{
  __shared__ MatchSmem SM[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 4 + warp;
  if (i >= n) return;
  MatchSmem* S = &SM[warp];
  const svob200_feature_ref f = ftrs[i];
  svob200_match_result* R = results ? &results[i] : nullptr;
  const DevFrame& ref = frames[(int)f.ref_frame_id];
  const DevFrame& cur = frames[cur_slot];
  double px_cur[2] = {px_in[2 * i], px_in[2 * i + 1]};
  // ref_ftr_->px.cast<int>()/(1<<level), boundary halfpatch_size_+2 (matcher.cpp:165-167)
  const int pxi = (int)f.px[0] / (1 << f.level), pyi = (int)f.px[1] / (1 << f.level);
  if (!in_frame_level(cam, pxi, pyi, 6, f.level)) {
    if (lane == 0 && ok_out) { ok_out[i] = 0; px_out[2 * i] = px_cur[0]; px_out[2 * i + 1] = px_cur[1]; }
    if (!R) return;
    if (lane == 0) {
      R->success = 0; R->search_level = 0; R->px_cur[0] = px_cur[0]; R->px_cur[1] = px_cur[1]; R->h_inv = 0;
      for (int k = 0; k < 4; ++k) R->A_cur_ref[k] = 0;
    }
    for (int k = lane; k < 100; k += 32) R->patch_with_border[k] = 0;
    for (int k = lane; k < 64; k += 32) R->patch[k] = 0;
    return;
  }
  double A[4];
  warp_matrix_affine(cam, f.px, {f.f[0], f.f[1], f.f[2]}, depth_ref[i], f.T_cur_ref, f.level, A);
  const int L = best_search_level(A, o.max_search_level);
  for (int k = lane; k < 100; k += 32) S->pwb[k] = 0;
  __syncwarp();
  const uint8_t* rimg = ref.lvl[f.level] + (size_t)f.ref_image * ref.img_stride[f.level];
  warp_affine_10x10(A, rimg, ref.pitch[f.level], ref.w[f.level], ref.h[f.level], f.px, f.level, L, S->pwb, lane, 32);
  __syncwarp();
  for (int k = lane; k < 64; k += 32) S->patch[k] = S->pwb[((k >> 3) + 1) * 10 + 1 + (k & 7)];
  __syncwarp();
  double pxs[2] = {px_cur[0] / (1 << L), px_cur[1] / (1 << L)};
  const uint8_t* cimg = cur.lvl[L] + (size_t)f.cur_image * cur.img_stride[L];
  bool success;
  double h_inv = 0;
  if (f.type == 1) {
    double dx = A[0] * f.grad[0] + A[1] * f.grad[1], dy = A[2] * f.grad[0] + A[3] * f.grad[1];
    { const double z = dx * dx + dy * dy; if (z > 0) { const double nn = sqrt(z); dx /= nn; dy /= nn; } }
    success = align1d_warp(cimg, cur.pitch[L], cur.w[L], cur.h[L], (float)dx, (float)dy, S->pwb, S->patch, o.align_max_iter, pxs, &h_inv, &S->al, lane);
  } else {
    success = align2d_warp(cimg, cur.pitch[L], cur.w[L], cur.h[L], S->pwb, S->patch, o.align_max_iter, pxs, &S->al, lane);
  }
  if (lane == 0 && ok_out) { ok_out[i] = success ? 1 : 0; px_out[2 * i] = pxs[0] * (1 << L); px_out[2 * i + 1] = pxs[1] * (1 << L); }
  if (!R) return;
  if (lane == 0) {
    R->success = success ? 1 : 0; R->search_level = L; R->h_inv = h_inv;
    R->px_cur[0] = pxs[0] * (1 << L); R->px_cur[1] = pxs[1] * (1 << L);
    for (int k = 0; k < 4; ++k) R->A_cur_ref[k] = A[k];
  }
  for (int k = lane; k < 100; k += 32) R->patch_with_border[k] = S->pwb[k];
  for (int k = lane; k < 64; k += 32) R->patch[k] = S->patch[k];
}

// stand-alone align2D / align1D on caller-provided patches: one warp per problem
__global__ void __launch_bounds__(128) align_patches_kernel(DevFrame f, int level, int n, const int* image, const uint8_t* pwb,
                                                            const uint8_t* patch, const float* dir, int n_iter, double* px,
                                                            int* converged, double* h_inv)
{
  __shared__ MatchSmem SM[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 4 + warp;
  if (i >= n) return;
  MatchSmem* S = &SM[warp];
  for (int k = lane; k < 100; k += 32) S->pwb[k] = pwb[100 * (size_t)i + k];
  for (int k = lane; k < 64; k += 32) S->patch[k] = patch[64 * (size_t)i + k];
  __syncwarp();
  const uint8_t* img = f.lvl[level] + (size_t)image[i] * f.img_stride[level];
  double pxs[2] = {px[2 * i], px[2 * i + 1]};
  bool ok;
  double hi = 0;
  if (dir) ok = align1d_warp(img, f.pitch[level], f.w[level], f.h[level], dir[2 * i], dir[2 * i + 1], S->pwb, S->patch, n_iter, pxs, &hi, &S->al, lane);
  else ok = align2d_warp(img, f.pitch[level], f.w[level], f.h[level], S->pwb, S->patch, n_iter, pxs, &S->al, lane);
  if (lane == 0) { px[2 * i] = pxs[0]; px[2 * i + 1] = pxs[1]; converged[i] = ok ? 1 : 0; if (h_inv) h_inv[i] = hi; }
}

__global__ void update_seed_kernel(int n, const float* x, const float* tau2, svob200_seed* seeds)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  svob200_seed s = seeds[i];
  update_seed(x[i], tau2[i], &s);
  seeds[i] = s;
}

__global__ void compute_tau_kernel(int n, const double* T, const double* f, const double* z, double angle, double* out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = compute_tau(T + 7 * (size_t)i, {f[3 * i], f[3 * i + 1], f[3 * i + 2]}, z[i], angle);
}

// stand-alone warp::getWarpMatrixAffine (matcher.cpp:36-60): thread per item
__global__ void warp_matrix_kernel(DevCam cam, int n, const double* px_ref, const double* f_ref, const double* depth_ref,
                                   const double* T_cur_ref, const int* level_ref, double* A_out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double A[4];
  warp_matrix_affine(cam, px_ref + 2 * i, {f_ref[3 * i], f_ref[3 * i + 1], f_ref[3 * i + 2]}, depth_ref[i], T_cur_ref + 7 * (size_t)i, level_ref[i], A);
  for (int k = 0; k < 4; ++k) A_out[4 * (size_t)i + k] = A[k];
}

// stand-alone warp::warpAffine (matcher.cpp:83-116) for any halfpatch size: thread per patch pixel.
// Leaves the patch untouched when the warp is NaN, like the reference (:94-98).
__global__ void warp_affine_kernel(const uint8_t* img, int pitch, int cols, int rows, double A0, double A1, double A2, double A3,
                                   double px0, double px1, int level_ref, int search_level, int halfpatch, uint8_t* patch)
{
  const int ps = 2 * halfpatch;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ps * ps) return;
  const double det = A0 * A3 - A2 * A1;
  const double invdet = 1.0 / det;
  const float a00 = (float)(A3 * invdet), a01 = (float)(-A1 * invdet);
  const float a10 = (float)(-A2 * invdet), a11 = (float)(A0 * invdet);
  if (isnan(a00)) return;
  const float pr0 = (float)px0 / (float)(1 << level_ref), pr1 = (float)px1 / (float)(1 << level_ref);
  const int y = i / ps, x = i - y * ps;
  float p0 = (float)(x - halfpatch), p1 = (float)(y - halfpatch);
  p0 *= (float)(1 << search_level); p1 *= (float)(1 << search_level);
  const float qx = (a00 * p0 + a01 * p1) + pr0;
  const float qy = (a10 * p0 + a11 * p1) + pr1;
  uint8_t v = 0;
  if (!(qx < 0 || qy < 0 || qx >= cols - 1 || qy >= rows - 1)) v = (uint8_t)interpolate_8u(img, pitch, qx, qy);
  patch[i] = v;
}

// stand-alone depthFromTriangulation (matcher.cpp:123-136): thread per item
__global__ void triangulate_kernel(int n, const double* T, const double* f_ref, const double* f_cur, double* depth, int* ok)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = 0.0;
  const bool r = depth_from_triangulation(T + 7 * (size_t)i, {f_ref[3 * i], f_ref[3 * i + 1], f_ref[3 * i + 2]},
                                          {f_cur[3 * i], f_cur[3 * i + 1], f_cur[3 * i + 2]}, &d);
  ok[i] = r ? 1 : 0;
  if (r) depth[i] = d;
}

}  // namespace

int launch_align_patches(const DevFrame& f, int level, int n, const int* d_image, const uint8_t* d_pwb, const uint8_t* d_patch,
                         const float* d_dir, int n_iter, double* d_px, int* d_converged, double* d_h_inv,
                         cudaStream_t s, long long* launches)
{
  if (n <= 0) return 0;
  align_patches_kernel<<<(n + 3) / 4, 128, 0, s>>>(f, level, n, d_image, d_pwb, d_patch, d_dir, n_iter, d_px, d_converged, d_h_inv);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_match_direct(const DevFrame* d_frames, const int* d_ref_slot, int cur_slot, const DevCam& cam, int n,
                        const svob200_feature_ref* d_ftrs, const double* d_depth_ref, const double* d_px_in,
                        svob200_matcher_opts opts, svob200_match_result* d_results, cudaStream_t s, long long* launches)
{
  if (n <= 0) return 0;
  match_direct_kernel<<<(n + 3) / 4, 128, 0, s>>>(d_frames, d_ref_slot, cur_slot, cam, n, d_ftrs, d_depth_ref, d_px_in, opts, d_results, nullptr, nullptr);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// same kernel, compact outputs only (refined pixel + success flag): what the tracker keeps per map point
int launch_match_direct_compact(const DevFrame* d_frames, int cur_slot, const DevCam& cam, int n, const svob200_feature_ref* d_ftrs,
                                const double* d_depth_ref, const double* d_px_in, svob200_matcher_opts opts, double* d_px_out,
                                int* d_ok_out, cudaStream_t s, long long* launches)
{
  if (n <= 0) return 0;
  match_direct_kernel<<<(n + 3) / 4, 128, 0, s>>>(d_frames, nullptr, cur_slot, cam, n, d_ftrs, d_depth_ref, d_px_in, opts, nullptr, d_px_out, d_ok_out);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_epipolar(const DevFrame* d_frames, const int* d_ref_slot, int cur_slot, const DevCam& cam, int n,
                    const svob200_feature_ref* d_ftrs, const double* d_d, svob200_matcher_opts opts,
                    svob200_epi_result* d_results, cudaStream_t s, long long* launches)
{
  if (n <= 0) return 0;
  epipolar_kernel<<<(n + 3) / 4, 128, 0, s>>>(d_frames, d_ref_slot, cur_slot, cam, n, d_ftrs, d_d, opts, d_results);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

size_t seeds_scratch_bytes(int n)
{
  const size_t m = (size_t)(n > 0 ? n : 1);
  return m * (sizeof(SeedPre) + sizeof(EpiGeom) + sizeof(EpiSearch)) + 2048;
}

int launch_seeds_update(const DevFrame* d_frames, const int* d_ref_slot, int cur_slot, const DevCam& cam, int n,
                        const svob200_feature_ref* d_ftrs, const double* d_T_ref_w, const double* d_T_cur_w,
                        svob200_matcher_opts opts, double conv_thresh, svob200_seed* d_seeds, svob200_seed_obs* d_obs,
                        void* d_scratch, int scratch_total, int first, cudaStream_t s, long long* launches, cudaEvent_t* marks)
{
  // d_ftrs / d_T_ref_w / d_seeds / d_obs already point at seed `first`; the scratch was sized for
  // scratch_total seeds and is indexed by absolute seed number
  (void)d_ref_slot;
  if (n <= 0) return 0;
  const size_t m = (size_t)(scratch_total > 0 ? scratch_total : 1);
  char* p = static_cast<char*>(d_scratch);
  EpiGeom* geom = reinterpret_cast<EpiGeom*>(p) + first; p += ((m * sizeof(EpiGeom) + 255) & ~(size_t)255);
  EpiSearch* search = reinterpret_cast<EpiSearch*>(p) + first; p += ((m * sizeof(EpiSearch) + 255) & ~(size_t)255);
  SeedPre* pre = reinterpret_cast<SeedPre*>(p) + first;
  seeds_geom_kernel<<<(n + 127) / 128, 128, 0, s>>>(cam, n, d_ftrs, d_T_ref_w, d_T_cur_w, opts, d_seeds, pre, geom);
  if (marks) cudaEventRecord(marks[0], s);
  seeds_search_kernel<<<(n + 3) / 4, 128, 0, s>>>(d_frames, cur_slot, cam, n, d_ftrs, opts, pre, geom, search);
  if (marks) cudaEventRecord(marks[1], s);
  seeds_finish_kernel<<<(n + 127) / 128, 128, 0, s>>>(cam, n, d_ftrs, conv_thresh, pre, geom, search, d_seeds, d_obs);
  *launches += 3;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_update_seed(int n, const float* d_x, const float* d_tau2, svob200_seed* d_seeds, cudaStream_t s, long long* launches)
{
  if (n <= 0) return 0;
  update_seed_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, d_x, d_tau2, d_seeds);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_compute_tau(int n, const double* d_T, const double* d_f, const double* d_z, double angle, double* d_out,
                       cudaStream_t s, long long* launches)
{
  if (n <= 0) return 0;
  compute_tau_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, d_T, d_f, d_z, angle, d_out);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_warp_matrix(const DevCam& cam, int n, const double* d_px_ref, const double* d_f_ref, const double* d_depth_ref,
                       const double* d_T_cur_ref, const int* d_level_ref, double* d_A_out, cudaStream_t s, long long* launches)
{
  if (n <= 0) return 0;
  warp_matrix_kernel<<<(n + 127) / 128, 128, 0, s>>>(cam, n, d_px_ref, d_f_ref, d_depth_ref, d_T_cur_ref, d_level_ref, d_A_out);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_warp_affine(const uint8_t* d_img, int pitch, int cols, int rows, const double* A, const double* px_ref, int level_ref,
                       int search_level, int halfpatch, uint8_t* d_patch, cudaStream_t s, long long* launches)
{
  const int n = 4 * halfpatch * halfpatch;
  if (n <= 0) return 0;
  warp_affine_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_img, pitch, cols, rows, A[0], A[1], A[2], A[3], px_ref[0], px_ref[1], level_ref,
                                                     search_level, halfpatch, d_patch);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_triangulate(int n, const double* d_T, const double* d_f_ref, const double* d_f_cur, double* d_depth, int* d_ok,
                       cudaStream_t s, long long* launches)
{
  if (n <= 0) return 0;
  triangulate_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, d_T, d_f_ref, d_f_cur, d_depth, d_ok);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
