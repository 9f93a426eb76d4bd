// TEST INFRASTRUCTURE ONLY — flat C entry points over the REAL reference code.
//
// Compiled (oracle/Makefile) together with the reference's own, unmodified
// translation units from /root/reference/app/src/main/cpp/svo into
// oracle/_ref/libsvo_ref.so.  No reference source is copied into this repo:
// this file only *calls* the reference's public operator surface
// (vk::halfSample, FastDetector::detect, SparseImgAlign::run,
// feature_alignment::align2D/align1D, Matcher::find*MatchDirect,
// DepthFilter::updateSeed/computeTau/addFrame) so that tests can pin the C
// restatement (svo_oracle.c) and the CUDA path against it.
//
// Pose layout: double[7] = {tx,ty,tz,qx,qy,qz,qw}.
#include <svo/global.h>
#include <svo/config.h>
#include <svo/vision.h>
#include <svo/patch_score.h>
#include <svo/pinhole_camera.h>
#include <svo/frame.h>
#include <svo/feature.h>
#include <svo/point.h>
#include <svo/feature_detection.h>
#include <svo/feature_alignment.h>
#include <svo/sparse_img_align.h>
#include <svo/matcher.h>
#include <svo/depth_filter.h>
#include <svo/map.h>
#include <svo/reprojector.h>
#include <svo/pose_optimizer.h>
#include <map>
#include <chrono>
#include <thread>
#include <mutex>
#include <cstring>
#ifdef SVOB200_DROPIN
// Same harness, linked over android_svo_b200/host/svo_b200_dropin.cpp INSTEAD OF the reference's
// vision / feature_alignment / matcher / sparse_img_align / feature_detection TUs
// (oracle/_ref/libsvo_dropin.so): every call below then lands in CUDA through the C ABI.
#include "svo_b200_dropin.h"
typedef svo::B200DepthFilter DepthFilterT;
#else
typedef svo::DepthFilter DepthFilterT;
#endif

using namespace svo;
namespace svo { bool depthFromTriangulation(const SE3& T_search_ref, const Vector3d& f_ref, const Vector3d& f_cur, double& depth); }

namespace {

SE3 to_se3(const double* T) { return SE3(T[0], T[1], T[2], T[3], T[4], T[5], T[6]); }
void from_se3(const SE3& T, double* o)
{
  o[0] = T.get_translation().x; o[1] = T.get_translation().y; o[2] = T.get_translation().z;
  o[3] = T.get_rotation().x; o[4] = T.get_rotation().y; o[5] = T.get_rotation().z; o[6] = T.get_rotation().w;
}

cv::Mat aligned_copy(const uint8_t* img, int w, int h)
{
  cv::Mat m(h, w, CV_8UC1);
  memcpy(m.data, img, (size_t)w * h);
  return m;
}

// exposes the protected solver state of SparseImgAlign
struct AlignProbe : public SparseImgAlign {
  AlignProbe(int maxl, int minl, int n_iter) : SparseImgAlign(maxl, minl, n_iter, GaussNewton, false, false) {}
  std::vector<int> evals;
  virtual double computeResiduals(const SE3& m, bool lin, bool w = false)
  {
    if ((int)evals.size() <= level_) evals.resize(level_ + 1, 0);
    evals[level_]++;
    return SparseImgAlign::computeResiduals(m, lin, w);
  }
#ifdef SVOB200_DROPIN
  void fetch_evals() { evals.assign(svo::b200::lastAlignIterations(), svo::b200::lastAlignIterations() + 8); }
#else
  void fetch_evals() {}
#endif
  const Matrix<double, 6, 6>& H() const { return H_; }
  const Matrix<double, 6, 1>& Jres() const { return Jres_; }
  const Matrix<double, 6, 1>& x() const { return x_; }
  double chi2() const { return chi2_; }
  size_t nmeas() const { return n_meas_; }
  bool stopped() const { return stop_; }
};

vk::PinholeCamera* make_cam(const int* wh, const double* k) { return new vk::PinholeCamera(wh[0], wh[1], k[0], k[1], k[2], k[3]); }

}  // namespace

extern "C" {

void svo_ref_config(int n_pyr_levels, int klt_max_level, int klt_min_level)
{
  Config::nPyrLevels() = n_pyr_levels;
  Config::kltMaxLevel() = klt_max_level;
  Config::kltMinLevel() = klt_min_level;
}

// 1 when this library is the B200 drop-in build (the hot-path symbols resolve to the CUDA adapter)
int svo_ref_is_dropin()
{
#ifdef SVOB200_DROPIN
  return 1;
#else
  return 0;
#endif
}

long long svo_ref_dropin_launches()
{
#ifdef SVOB200_DROPIN
  return svo::b200::launchCount();
#else
  return 0;
#endif
}

void svo_ref_dropin_shutdown()
{
#ifdef SVOB200_DROPIN
  svo::b200::shutdown();
#endif
}

int svo_ref_has_sse2()
{
#ifdef __SSE2__
  return 1;
#else
  return 0;
#endif
}

// vk::halfSample on 64-byte aligned Mats (dispatch rule of the host build applies)
void svo_ref_half_sample(const uint8_t* in, int w, int h, uint8_t* out)
{
  cv::Mat mi = aligned_copy(in, w, h);
  cv::Mat mo(h / 2, w / 2, CV_8U);
  vk::halfSample(mi, mo);
  memcpy(out, mo.data, (size_t)(w / 2) * (h / 2));
}

// frame_utils::createImgPyramid; out = levels 1..n-1 concatenated
void svo_ref_pyramid(const uint8_t* img, int w, int h, int n_levels, uint8_t* out)
{
  cv::Mat m0 = aligned_copy(img, w, h);
  ImgPyr pyr;
  frame_utils::createImgPyramid(m0, n_levels, pyr);
  for (int l = 1; l < n_levels; ++l) {
    memcpy(out, pyr[l].data, (size_t)pyr[l].cols * pyr[l].rows);
    out += (size_t)pyr[l].cols * pyr[l].rows;
  }
}

float svo_ref_shi_tomasi(const uint8_t* img, int w, int h, int u, int v)
{
  cv::Mat m(h, w, CV_8UC1, (void*)img);
  return vk::shiTomasiScore(m, u, v);
}

int svo_ref_zmssd(const uint8_t* ref_patch, const uint8_t* cur, int stride)
{
  uint8_t patch[64] __attribute__((aligned(16)));
  memcpy(patch, ref_patch, 64);
  vk::patch_score::ZMSSD<4> score(patch);
  return score.computeScore((uint8_t*)cur, stride);
}

// FastDetector::detect on a Frame built from img (pyramid depth from Config).
// occ_px: n_occ pixel positions marked via setGridOccpuancy before detect.
int svo_ref_fast_detect(const uint8_t* img, const int* wh, const double* k, int n_detect_levels, int cell, double thr,
                        int n_occ, const double* occ_px, int cap, double* out_px, int* out_level)
{
  vk::PinholeCamera* cam = make_cam(wh, k);
  int n = 0;
  {
    cv::Mat m0 = aligned_copy(img, wh[0], wh[1]);
    Frame frame(cam, m0, 0.0);
    feature_detection::FastDetector det(wh[0], wh[1], cell, n_detect_levels);
    for (int i = 0; i < n_occ; ++i) det.setGridOccpuancy(Vector2d(occ_px[2 * i], occ_px[2 * i + 1]));
    Features fts;
    det.detect(&frame, frame.img_pyr_, thr, fts);
    for (auto it = fts.begin(); it != fts.end(); ++it, ++n) {
      if (n < cap) { out_px[2 * n] = (*it)->px[0]; out_px[2 * n + 1] = (*it)->px[1]; out_level[n] = (*it)->level; }
      delete *it;
    }
  }
  delete cam;
  return n;
}

int svo_ref_align2d(const uint8_t* img, int w, int h, const uint8_t* pwb, const uint8_t* patch, int n_iter, double* px)
{
  cv::Mat m(h, w, CV_8UC1, (void*)img);
  uint8_t a[100] __attribute__((aligned(16))), b[64] __attribute__((aligned(16)));
  memcpy(a, pwb, 100); memcpy(b, patch, 64);
  Vector2d p(px[0], px[1]);
  bool ok = feature_alignment::align2D(m, a, b, n_iter, p);
  px[0] = p[0]; px[1] = p[1];
  return ok ? 1 : 0;
}

int svo_ref_align1d(const uint8_t* img, int w, int h, const float* dir, const uint8_t* pwb, const uint8_t* patch,
                    int n_iter, double* px, double* h_inv)
{
  cv::Mat m(h, w, CV_8UC1, (void*)img);
  uint8_t a[100] __attribute__((aligned(16))), b[64] __attribute__((aligned(16)));
  memcpy(a, pwb, 100); memcpy(b, patch, 64);
  Vector2d p(px[0], px[1]);
  bool ok = feature_alignment::align1D(m, Vector2f(dir[0], dir[1]), a, b, n_iter, p, *h_inv);
  px[0] = p[0]; px[1] = p[1];
  return ok ? 1 : 0;
}

void svo_ref_se3_mul(const double* A, const double* B, double* o) { from_se3(to_se3(A) * to_se3(B), o); }
void svo_ref_se3_inverse(const double* A, double* o) { from_se3(to_se3(A).inverse(), o); }
void svo_ref_se3_exp(const double* x, double* o) { from_se3(SE3::exp(x), o); }
void svo_ref_se3_transform(const double* T, const double* p, double* o)
{
  Vector3d r = to_se3(T) * Vector3d(p[0], p[1], p[2]);
  o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
}
void svo_ref_cam2world(const int* wh, const double* k, double u, double v, double* f)
{
  vk::PinholeCamera* cam = make_cam(wh, k);
  Vector3d r = cam->cam2world(Vector2d(u, v));
  f[0] = r[0]; f[1] = r[1]; f[2] = r[2];
  delete cam;
}

// Builds the reference-side inputs of sparse alignment from level-0 pixels and
// world points exactly as the reference does (Feature ctor: f = cam2world(px);
// sparse_img_align.cpp:132-134: depth = |pos - ref_pos|, xyz_ref = f*depth),
// runs SparseImgAlign::run, and returns both the result and the derived
// per-feature arrays so the other implementations can be fed identical values.
// pt_world: 3N, NaN in x => feature without point.
int svo_ref_sparse_align(const uint8_t* ref_img, const uint8_t* cur_img, const int* wh, const double* k,
                         int max_level, int min_level, int n_iter,
                         const double* T_ref_w, const double* T_cur_w_init,
                         int N, const double* px, const int* level, const double* pt_world,
                         double* T_cur_w_out, double* T_cur_ref_init_out, double* T_cur_ref_out,
                         double* H_out, double* Jres_out, double* x_out, double* chi2_out, int* n_meas_out,
                         int* iters_out /*8*/, int* stop_out,
                         double* f_out /*3N*/, double* xyz_ref_out /*3N*/)
{
  vk::PinholeCamera* cam = make_cam(wh, k);
  size_t ret = 0;
  {
    FramePtr ref(new Frame(cam, aligned_copy(ref_img, wh[0], wh[1]), 0.0));
    FramePtr cur(new Frame(cam, aligned_copy(cur_img, wh[0], wh[1]), 1.0));
    ref->T_f_w_ = to_se3(T_ref_w);
    cur->T_f_w_ = to_se3(T_cur_w_init);
    std::vector<Point*> pts;
    const Vector3d ref_pos = ref->pos();
    for (int i = 0; i < N; ++i) {
      Feature* ftr = new Feature(ref.get(), Vector2d(px[2 * i], px[2 * i + 1]), level ? level[i] : 0);
      if (!std::isnan(pt_world[3 * i])) {
        Point* pt = new Point(Vector3d(pt_world[3 * i], pt_world[3 * i + 1], pt_world[3 * i + 2]));
        ftr->point = pt; pts.push_back(pt);
        const double depth((ftr->point->pos_ - ref_pos).norm());
        const Vector3d xyz_ref(ftr->f * depth);
        for (int c = 0; c < 3; ++c) xyz_ref_out[3 * i + c] = xyz_ref[c];
      } else {
        for (int c = 0; c < 3; ++c) xyz_ref_out[3 * i + c] = 0.0;
      }
      for (int c = 0; c < 3; ++c) f_out[3 * i + c] = ftr->f[c];
      ref->addFeature(ftr);
    }
    from_se3(SE3(cur->T_f_w_ * ref->T_f_w_.inverse()), T_cur_ref_init_out);
    AlignProbe al(max_level, min_level, n_iter);
    ret = al.run(ref, cur);
    al.fetch_evals();
    from_se3(cur->T_f_w_, T_cur_w_out);
    from_se3(SE3(cur->T_f_w_ * ref->T_f_w_.inverse()), T_cur_ref_out);  // informational (recomposed)
    for (int a = 0; a < 6; ++a) {
      for (int b = 0; b < 6; ++b) H_out[a * 6 + b] = al.H()(a, b);
      Jres_out[a] = al.Jres()[a]; x_out[a] = al.x()[a];
    }
    *chi2_out = al.chi2(); *n_meas_out = (int)al.nmeas(); *stop_out = al.stopped() ? 1 : 0;
    for (int l = 0; l < 8; ++l) iters_out[l] = l < (int)al.evals.size() ? al.evals[l] : 0;
    for (auto p : pts) delete p;
  }
  delete cam;
  return (int)ret;
}

// warp::getWarpMatrixAffine + getBestSearchLevel + warpAffine(halfpatch 5)
int svo_ref_warp(const uint8_t* ref_img_level, int w, int h, const int* wh, const double* k,
                 const double* px_ref, const double* f_ref, double depth_ref, const double* T_cur_ref,
                 int level_ref, int max_search_level, double* A_out /*row-major*/, int* search_level_out, uint8_t* patch100)
{
  vk::PinholeCamera* cam = make_cam(wh, k);
  Matrix2d A;
  warp::getWarpMatrixAffine(*cam, *cam, Vector2d(px_ref[0], px_ref[1]), Vector3d(f_ref[0], f_ref[1], f_ref[2]), depth_ref,
                            to_se3(T_cur_ref), level_ref, A);
  A_out[0] = A(0, 0); A_out[1] = A(0, 1); A_out[2] = A(1, 0); A_out[3] = A(1, 1);
  const int sl = warp::getBestSearchLevel(A, max_search_level);
  *search_level_out = sl;
  cv::Mat m(h, w, CV_8UC1, (void*)ref_img_level);
  warp::warpAffine(A, m, Vector2d(px_ref[0], px_ref[1]), level_ref, sl, 5, patch100);
  delete cam;
  return 1;
}

// Matcher::findMatchDirect with a Point that has exactly the given observations.
// obs k: frame image index (into imgs), pose, px, level. The reference picks the close-view obs itself.
int svo_ref_find_match_direct(int n_frames, const uint8_t* const* imgs, const double* T_f_w /*7 per frame*/,
                              const int* wh, const double* k,
                              const double* pt_world, int n_obs, const int* obs_frame, const double* obs_px, const int* obs_level,
                              const int* obs_type, const double* obs_grad,
                              int cur_frame, double* px_cur /*inout*/,
                              int* chosen_obs, int* search_level, double* A_out, double* h_inv, uint8_t* pwb_out, uint8_t* patch_out)
{
  vk::PinholeCamera* cam = make_cam(wh, k);
  bool ok = false;
  {
    std::vector<FramePtr> frames;
    for (int i = 0; i < n_frames; ++i) {
      FramePtr f(new Frame(cam, aligned_copy(imgs[i], wh[0], wh[1]), (double)i));
      f->T_f_w_ = to_se3(T_f_w + 7 * i);
      frames.push_back(f);
    }
    Point pt(Vector3d(pt_world[0], pt_world[1], pt_world[2]));
    std::vector<Feature*> fs;
    for (int i = 0; i < n_obs; ++i) {
      Feature* ftr = new Feature(frames[obs_frame[i]].get(), Vector2d(obs_px[2 * i], obs_px[2 * i + 1]), obs_level[i]);
      ftr->type = obs_type[i] ? Feature::EDGELET : Feature::CORNER;
      ftr->grad = Vector2d(obs_grad[2 * i], obs_grad[2 * i + 1]);
      ftr->point = &pt;
      pt.addFrameRef(ftr);
      fs.push_back(ftr);
    }
    Matcher matcher;
    matcher.ref_ftr_ = NULL; matcher.search_level_ = -1; matcher.h_inv_ = 0; matcher.A_cur_ref_.setZero();
    memset(matcher.patch_with_border_, 0, 100); memset(matcher.patch_, 0, 64);
    Vector2d p(px_cur[0], px_cur[1]);
    ok = matcher.findMatchDirect(pt, *frames[cur_frame], p);
    px_cur[0] = p[0]; px_cur[1] = p[1];
    *chosen_obs = -1;
    for (int i = 0; i < n_obs; ++i) if (fs[i] == matcher.ref_ftr_) *chosen_obs = i;
    *search_level = matcher.search_level_;
    A_out[0] = matcher.A_cur_ref_(0, 0); A_out[1] = matcher.A_cur_ref_(0, 1); A_out[2] = matcher.A_cur_ref_(1, 0); A_out[3] = matcher.A_cur_ref_(1, 1);
    *h_inv = matcher.h_inv_;
    memcpy(pwb_out, matcher.patch_with_border_, 100); memcpy(patch_out, matcher.patch_, 64);
    pt.obs_.clear();
    for (auto f : fs) delete f;
  }
  delete cam;
  return ok ? 1 : 0;
}

// Matcher::findEpipolarMatchDirect
int svo_ref_find_epipolar_match(const uint8_t* ref_img, const uint8_t* cur_img, const int* wh, const double* k,
                                const double* T_ref_w, const double* T_cur_w,
                                const double* px_ref, int level_ref, int type, const double* grad,
                                double d_estimate, double d_min, double d_max,
                                int align_1d, int align_max_iter, int max_epi_search_steps, int subpix_refinement,
                                double* depth, double* px_cur, double* epi_length, int* search_level, int* reject,
                                double* A_out, double* f_ref_out, uint8_t* pwb_out, uint8_t* patch_out)
{
  vk::PinholeCamera* cam = make_cam(wh, k);
  bool ok = false;
  {
    Frame ref(cam, aligned_copy(ref_img, wh[0], wh[1]), 0.0);
    Frame cur(cam, aligned_copy(cur_img, wh[0], wh[1]), 1.0);
    ref.T_f_w_ = to_se3(T_ref_w); cur.T_f_w_ = to_se3(T_cur_w);
    Feature ftr(&ref, Vector2d(px_ref[0], px_ref[1]), level_ref);
    ftr.type = type ? Feature::EDGELET : Feature::CORNER;
    ftr.grad = Vector2d(grad[0], grad[1]);
    for (int c = 0; c < 3; ++c) f_ref_out[c] = ftr.f[c];
    Matcher m;
    m.options_.align_1d = align_1d; m.options_.align_max_iter = align_max_iter;
    m.options_.max_epi_search_steps = max_epi_search_steps; m.options_.subpix_refinement = subpix_refinement;
    m.px_cur_.setZero(); m.epi_length_ = 0; m.search_level_ = -1; m.reject_ = false; m.A_cur_ref_.setZero();
    memset(m.patch_with_border_, 0, 100); memset(m.patch_, 0, 64);
    *depth = 0;
    ok = m.findEpipolarMatchDirect(ref, cur, ftr, d_estimate, d_min, d_max, *depth);
    px_cur[0] = m.px_cur_[0]; px_cur[1] = m.px_cur_[1];
    *epi_length = m.epi_length_; *search_level = m.search_level_; *reject = m.reject_ ? 1 : 0;
    A_out[0] = m.A_cur_ref_(0, 0); A_out[1] = m.A_cur_ref_(0, 1); A_out[2] = m.A_cur_ref_(1, 0); A_out[3] = m.A_cur_ref_(1, 1);
    memcpy(pwb_out, m.patch_with_border_, 100); memcpy(patch_out, m.patch_, 64);
  }
  delete cam;
  return ok ? 1 : 0;
}

int svo_ref_depth_from_triangulation(const double* T, const double* f_ref, const double* f_cur, double* depth)
{
  return svo::depthFromTriangulation(to_se3(T), Vector3d(f_ref[0], f_ref[1], f_ref[2]), Vector3d(f_cur[0], f_cur[1], f_cur[2]), *depth) ? 1 : 0;
}

void svo_ref_seed_init(float depth_mean, float depth_min, float* s /*a,b,mu,z_range,sigma2*/)
{
  Seed seed(NULL, depth_mean, depth_min);
  s[0] = seed.a; s[1] = seed.b; s[2] = seed.mu; s[3] = seed.z_range; s[4] = seed.sigma2;
}

void svo_ref_update_seed(float x, float tau2, float* s)
{
  Seed seed(NULL, 1.0f, 1.0f);
  seed.a = s[0]; seed.b = s[1]; seed.mu = s[2]; seed.z_range = s[3]; seed.sigma2 = s[4];
  DepthFilter::updateSeed(x, tau2, &seed);
  s[0] = seed.a; s[1] = seed.b; s[2] = seed.mu; s[3] = seed.z_range; s[4] = seed.sigma2;
}

double svo_ref_compute_tau(const double* T_ref_cur, const double* f, double z, double px_error_angle)
{
  return DepthFilter::computeTau(to_se3(T_ref_cur), Vector3d(f[0], f[1], f[2]), z, px_error_angle);
}

// DepthFilter::updateSeeds (through addFrame, synchronous mode) over S seeds that live on
// n_ref reference frames.  status: 0 = seed still in list, 1 = converged (callback fired), 2 = erased otherwise.
// For converged seeds the callback's sigma2 is returned in state[4] and the other fields are
// recomputed by a second run with an infinite convergence threshold.
int svo_ref_update_seeds(int n_ref, const uint8_t* const* ref_imgs, const double* T_ref_w,
                         const uint8_t* cur_img, const double* T_cur_w, const int* wh, const double* k,
                         int S, const int* seed_ref, const double* px, const int* level,
                         float* state /*5 per seed, inout*/, int* status, double conv_thresh)
{
  vk::PinholeCamera* cam = make_cam(wh, k);
  {
    std::vector<FramePtr> refs;
    for (int i = 0; i < n_ref; ++i) {
      FramePtr f(new Frame(cam, aligned_copy(ref_imgs[i], wh[0], wh[1]), (double)i));
      f->T_f_w_ = to_se3(T_ref_w + 7 * i);
      refs.push_back(f);
    }
    FramePtr cur(new Frame(cam, aligned_copy(cur_img, wh[0], wh[1]), 100.0));
    cur->T_f_w_ = to_se3(T_cur_w);
    std::vector<float> in(state, state + 5 * S), out_inf(5 * S);
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<Feature*> fs(S);
      std::map<Feature*, int> index;
      std::vector<Point*> new_points;
      std::vector<int> st(S, 2);
      std::vector<float> conv_sigma2(S, 0.f);
      DepthFilterT df(feature_detection::DetectorPtr(), [&](Point* p, double sigma2) {
        const int i = index[p->obs_.front()];
        st[i] = 1; conv_sigma2[i] = (float)sigma2; new_points.push_back(p);
      });
      df.options_.seed_convergence_sigma2_thresh = pass == 0 ? conv_thresh : 1e300;
      Seed::batch_counter = 0;
      for (int i = 0; i < S; ++i) {
        fs[i] = new Feature(refs[seed_ref[i]].get(), Vector2d(px[2 * i], px[2 * i + 1]), level[i]);
        index[fs[i]] = i;
        Seed seed(fs[i], 1.0f, 1.0f);
        seed.a = in[5 * i]; seed.b = in[5 * i + 1]; seed.mu = in[5 * i + 2]; seed.z_range = in[5 * i + 3]; seed.sigma2 = in[5 * i + 4];
        df.getSeeds().push_back(seed);
      }
      df.addFrame(cur);
      std::vector<float>& dst = out_inf;
      for (auto& s : df.getSeeds()) {
        const int i = index[s.ftr];
        if (pass == 0) st[i] = 0;
        if (pass == 1) { dst[5 * i] = s.a; dst[5 * i + 1] = s.b; dst[5 * i + 2] = s.mu; dst[5 * i + 3] = s.z_range; dst[5 * i + 4] = s.sigma2; }
      }
      if (pass == 0) for (int i = 0; i < S; ++i) status[i] = st[i];
      if (pass == 1) {
        // seeds erased in the infinite-threshold pass (NaN) keep their input state
        std::vector<char> seen(S, 0);
        for (auto& s : df.getSeeds()) seen[index[s.ftr]] = 1;
        for (int i = 0; i < S; ++i) if (!seen[i]) for (int c = 0; c < 5; ++c) dst[5 * i + c] = in[5 * i + c];
      }
      for (auto p : new_points) { p->obs_.clear(); delete p; }
      for (auto f : fs) delete f;
    }
    memcpy(state, out_inf.data(), sizeof(float) * 5 * S);
  }
  delete cam;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// One front-end step per frame with the reference's own operators, chained the way
// FrameHandlerMono::processFrame + DepthFilter do (frame_handler_mono.cpp:171-262,
// depth_filter.cpp:237-341) minus pose_optimizer / map management:
//   new Frame (pyramid) -> SparseImgAlign::run(last, cur) -> Matcher::findMatchDirect per map point
//   -> DepthFilter::addFrame(cur) (synchronous updateSeeds).
struct svo_ref_step_stats {
  double T_cur_w[7];
  double chi2;
  int n_tracked, n_matched, n_seeds_updated, n_seeds_converged, n_seeds_failed, n_seeds_skipped;
  int align_iters, n_exact_chi2;
  int n_reproj_trials, n_pose_obs;
};

// The reference's native two-thread layout: tracking in the caller's thread, DepthFilter::updateSeedsLoop in its own
// (DepthFilter::startThread, depth_filter.cpp:63-67).  pending() / wait_idle() let the harness pace the tracking thread so
// that the frame queue never drops a frame (:95-97) and drain it at the end of a timed run.
struct ThreadedDF : public DepthFilterT {
  ThreadedDF(feature_detection::DetectorPtr d, callback_t cb) : DepthFilterT(d, cb) {}
  size_t pending() { std::unique_lock<std::mutex> l(frame_queue_mut_); return frame_queue_.size(); }
  void wait_idle()
  {
    while (pending() > 0) std::this_thread::yield();
    for (int k = 0; k < 3; ++k) { { std::unique_lock<std::mutex> l(seeds_mut_); } std::this_thread::yield(); }   // the update in flight holds seeds_mut_
  }
  // DepthFilter::stopThread joins a thread that sleeps in frame_queue_cond_.wait with an empty queue and is never notified:
  // wake it with one more frame (updateSeeds returns at once: seeds_updating_halt_ is set), then join
  void shutdown(FramePtr any)
  {
    if (!thread_) return;
    seeds_updating_halt_ = true; thread_stop_ = true;
    { std::unique_lock<std::mutex> l(frame_queue_mut_); frame_queue_.push(any); }
    frame_queue_cond_.notify_one();
    if (thread_->joinable()) thread_->join();
    thread_ = NULL;
  }
};
struct RefSeq {
  vk::PinholeCamera* cam;
  FramePtr kf, last;
  std::vector<Point*> pts;
  std::vector<Feature*> seed_ftrs;
  std::map<Feature*, int> seed_index;
  DepthFilter* df;
  Matcher matcher;
  int max_level, min_level, n_iter, reseed;
  float depth_mean, depth_min;
  std::vector<Point*> conv_points;
  // chain mode: the reference's own Reprojector + pose optimiser between alignment and the depth filter
  svo::Map* map = NULL;
  Reprojector* reproj = NULL;
  int chain_pose_opt = 0;
  // std::chrono::steady_clock seconds per operator, accumulated over the steps since the last svo_ref_seq_get_timing
  // (BASELINE.md section 3): [0] Frame ctor (pyramid), [1] SparseImgAlign::run, [2] reprojection / refinement (+ pose optimiser in
  // chain mode), [3] DepthFilter::addFrame (updateSeeds); [4] = number of steps
  double timing[5] = {0, 0, 0, 0, 0};
  // keyframe insertion (svo_ref_seq_add_keyframe): the reference's own DepthFilter::addKeyframe / removeKeyframe with its FastDetector
  feature_detection::DetectorPtr det;
  std::vector<FramePtr> kfs;            // ring: index = keyframe index (0 = the keyframe of set_keyframe)
  std::vector<int> kf_batch;
  int max_kfs = 4;
  double conv_thresh = 100.0;
  std::vector<double> step_px; std::vector<int> step_ok, step_level;
  ThreadedDF* tdf = NULL;               // two-thread layout (svo_ref_seq_set_threaded)
  bool kf_dropped = false;
  long n_conv_async = 0;
};
static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void* svo_ref_seq_create(const int* wh, const double* k, int max_level, int min_level, int n_iter, double conv_thresh,
                         float depth_mean, float depth_min, int reseed)
{
  RefSeq* s = new RefSeq();
  s->cam = make_cam(wh, k);
  s->max_level = max_level; s->min_level = min_level; s->n_iter = n_iter; s->reseed = reseed;
  s->depth_mean = depth_mean; s->depth_min = depth_min;
  s->df = new DepthFilterT(feature_detection::DetectorPtr(), [s](Point* p, double) { s->conv_points.push_back(p); });
  s->df->options_.seed_convergence_sigma2_thresh = conv_thresh;
  s->conv_thresh = conv_thresh;
  return s;
}

// Before svo_ref_seq_set_keyframe: keyframe ring, DepthFilter::Options::max_n_kfs, reseed mode (3 = the reference's own list
// semantics: finished seeds simply leave the list) and the detector DepthFilter::initializeSeeds runs
// (FastDetector(width, height, Config::gridSize, Config::nPyrLevels), threshold Config::triangMinCornerScore).
void svo_ref_seq_set_pool(void* h, int max_kfs, int max_n_kfs, int reseed, int det_cell, int det_levels, double det_thr)
{
  RefSeq* s = (RefSeq*)h;
  s->max_kfs = max_kfs; s->reseed = reseed;
  delete s->df;
  s->det.reset(new feature_detection::FastDetector(s->cam->width(), s->cam->height(), det_cell, det_levels));
  s->df = new DepthFilterT(s->det, [s](Point* p, double) { s->conv_points.push_back(p); });
  s->df->options_.seed_convergence_sigma2_thresh = s->conv_thresh;
  s->df->options_.max_n_kfs = max_n_kfs;
  Config::triangMinCornerScore() = det_thr;
}

// Timing only: the depth filter runs in its own thread like in the app (call before svo_ref_seq_set_keyframe).  Finished seeds
// are re-initialised inside the convergence callback (the list is being walked under seeds_mut_ there), so the workload stays
// stationary without the tracking thread touching the list.
void svo_ref_seq_set_threaded(void* h)
{
  RefSeq* s = (RefSeq*)h;
  delete s->df;
  s->tdf = new ThreadedDF(s->det, [s](Point* p, double) {
    Feature* f = p->obs_.front(); f->point = NULL; p->obs_.clear(); delete p;
    ++s->n_conv_async;
    if (s->reseed) s->df->getSeeds().push_back(Seed(f, s->depth_mean, s->depth_min));
  });
  s->df = s->tdf;
  s->df->options_.seed_convergence_sigma2_thresh = s->conv_thresh;
  s->df->startThread();
}
// waits until the depth-filter thread has consumed every queued frame
void svo_ref_seq_drain(void* h) { RefSeq* s = (RefSeq*)h; if (s->tdf) s->tdf->wait_idle(); }

// The frame of the most recent step becomes a keyframe: FrameHandlerMono::processFrame :260-312 restricted to the depth filter's
// part — setKeyframe, DepthFilter::addKeyframe (synchronous: initializeSeeds), removeKeyframe of the oldest one when the ring is full.
int svo_ref_seq_add_keyframe(void* h, float depth_mean, float depth_min)
{
  RefSeq* s = (RefSeq*)h;
  FramePtr fr = s->last;
  // the frame's features: the map points matched in it, at their refined pixels (what Reprojector::reprojectCell adds, :219-231)
  for (auto f : fr->fts_) delete f;
  fr->fts_.clear();
  for (size_t i = 0; i < s->pts.size(); ++i) {
    if (!s->step_ok[i]) continue;
    Feature* f = new Feature(fr.get(), Vector2d(s->step_px[2 * i], s->step_px[2 * i + 1]), s->step_level[i]);
    f->point = s->pts[i];
    fr->fts_.push_back(f);
  }
  fr->setKeyframe();
  for (auto f : fr->fts_) f->point->addFrameRef(f);                  // frame_handler_mono.cpp:277-279
  if (s->kfs.empty()) { s->kfs.assign(s->max_kfs, FramePtr()); s->kf_batch.assign(s->max_kfs, 0); s->kfs[0] = s->kf; }
  int k = -1;
  for (int i = 0; i < s->max_kfs; ++i) if (!s->kfs[i]) { k = i; break; }
  if (k < 0) {
    k = 0;
    for (int i = 1; i < s->max_kfs; ++i) if (s->kf_batch[i] < s->kf_batch[k]) k = i;
    s->df->removeKeyframe(s->kfs[k]);
    for (auto p : s->pts) p->deleteFrameRef(s->kfs[k].get());           // Map::safeDeleteFrame (map.cpp): the points forget the keyframe
    if (s->kfs[k] == s->kf) s->kf_dropped = true;
  }
  s->kfs[k] = fr;
  const size_t before = s->df->getSeeds().size();
  s->df->addKeyframe(fr, depth_mean, depth_min);
  s->kf_batch[k] = Seed::batch_counter;
  return (int)(s->df->getSeeds().size() - before);
}

// the seed list as it stands: px (2), level, keyframe index, batch id and the 5 state floats per seed; returns the count
int svo_ref_seq_get_seed_list(void* h, int cap, double* px, int* level, int* kf, int* batch, float* state)
{
  RefSeq* s = (RefSeq*)h;
  int n = 0;
  for (auto& sd : s->df->getSeeds()) {
    if (n >= cap) break;
    px[2 * n] = sd.ftr->px[0]; px[2 * n + 1] = sd.ftr->px[1]; level[n] = sd.ftr->level; batch[n] = sd.batch_id;
    int ki = 0;
    for (size_t i = 0; i < s->kfs.size(); ++i) if (s->kfs[i].get() == sd.ftr->frame) ki = (int)i;
    kf[n] = ki;
    state[5 * n] = sd.a; state[5 * n + 1] = sd.b; state[5 * n + 2] = sd.mu; state[5 * n + 3] = sd.z_range; state[5 * n + 4] = sd.sigma2;
    ++n;
  }
  return n;
}

// FrameHandlerMono::processFrame's Step 2 + Step 3 (frame_handler_mono.cpp:191-222) instead of refining every map point;
// call after svo_ref_seq_set_keyframe.  Config::gridSize / maxFts are process-wide, like in the reference.
void svo_ref_seq_set_chain(void* h, int cell_size, int max_fts, int pose_opt)
{
  RefSeq* s = (RefSeq*)h;
  Config::gridSize() = cell_size;
  Config::maxFts() = max_fts;
  s->map = new svo::Map();
  s->kf->setKeyframe();
  s->map->addKeyframe(s->kf);
  s->reproj = new Reprojector(s->cam, *s->map);
  s->chain_pose_opt = pose_opt;
}

void svo_ref_seq_destroy(void* h)
{
  RefSeq* s = (RefSeq*)h;
  if (s->tdf) { s->tdf->wait_idle(); s->tdf->shutdown(s->last ? s->last : s->kf); }
  if (s->reproj) delete s->reproj;
  if (s->map) { s->map->keyframes_.clear(); delete s->map; }
  s->df->getSeeds().clear();
  delete s->df;
  for (auto f : s->seed_ftrs) delete f;
  s->kf.reset(); s->last.reset();          // ~Frame deletes its features
  for (auto p : s->pts) { p->obs_.clear(); delete p; }
  delete s->cam;
  delete s;
}

void svo_ref_seq_set_keyframe(void* h, const uint8_t* img, const double* T_kf_w, int N, const double* kf_px, const int* kf_level,
                              const double* pt_world, int S, const double* seed_px, const int* seed_level)
{
  RefSeq* s = (RefSeq*)h;
  s->kf.reset(new Frame(s->cam, aligned_copy(img, s->cam->width(), s->cam->height()), 0.0));
  s->kf->T_f_w_ = to_se3(T_kf_w);
  for (int i = 0; i < N; ++i) {
    Point* pt = new Point(Vector3d(pt_world[3 * i], pt_world[3 * i + 1], pt_world[3 * i + 2]));
    Feature* f = new Feature(s->kf.get(), Vector2d(kf_px[2 * i], kf_px[2 * i + 1]), kf_level[i]);
    f->point = pt;
    pt->addFrameRef(f);
    s->kf->addFeature(f);
    s->pts.push_back(pt);
  }
  Seed::batch_counter = 0;
  for (int i = 0; i < S; ++i) {
    Feature* f = new Feature(s->kf.get(), Vector2d(seed_px[2 * i], seed_px[2 * i + 1]), seed_level[i]);
    s->seed_ftrs.push_back(f);
    s->seed_index[f] = i;
    s->df->getSeeds().push_back(Seed(f, s->depth_mean, s->depth_min));
  }
}

void svo_ref_seq_set_last(void* h, const uint8_t* img)
{
  RefSeq* s = (RefSeq*)h;
  s->last.reset(new Frame(s->cam, aligned_copy(img, s->cam->width(), s->cam->height()), 0.0));
}

void svo_ref_seq_step(void* h, const uint8_t* cur_img, const double* T_last_w, const double* last_px,
                      svo_ref_step_stats* st, double* px_refined, int* match_ok)
{
  RefSeq* s = (RefSeq*)h;
  memset(st, 0, sizeof(*st));
  const double t0 = now_s();
  FramePtr cur(new Frame(s->cam, aligned_copy(cur_img, s->cam->width(), s->cam->height()), 1.0));
  const double t1 = now_s();
  FramePtr last = s->last;
  if (last->isKeyframe() || s->tdf) {        // (two-thread layout: the depth-filter thread may still be reading the frame)
    // a keyframe keeps the pose it was inserted with (and its features); tracking continues from a twin of the frame (same
    // image, same pyramid) that takes the caller's pose of the last frame like every other `last`
    last.reset(new Frame(s->cam, last->img_pyr_[0], 0.0));
  }
  last->T_f_w_ = to_se3(T_last_w);
  for (auto f : last->fts_) delete f;
  last->fts_.clear();
  const int N = (int)s->pts.size();
  for (int i = 0; i < N; ++i) {
    Feature* f = new Feature(last.get(), Vector2d(last_px[2 * i], last_px[2 * i + 1]), 0);
    f->point = s->pts[i];
    last->fts_.push_back(f);
  }
  cur->T_f_w_ = last->T_f_w_;                                       // frame_handler_mono.cpp:175
  const double t2 = now_s();
  AlignProbe al(s->max_level, s->min_level, s->n_iter);
  st->n_tracked = (int)al.run(last, cur);
  const double t3 = now_s();
  al.fetch_evals();
  st->chi2 = al.chi2();
  for (size_t l = 0; l < al.evals.size(); ++l) st->align_iters += al.evals[l];
  from_se3(cur->T_f_w_, st->T_cur_w);
  if (s->reproj) {
    std::vector<std::pair<FramePtr, size_t> > overlap;
    s->reproj->reprojectMap(cur, overlap);
    st->n_matched = (int)s->reproj->n_matches_; st->n_reproj_trials = (int)s->reproj->n_trials_;
    std::map<Point*, int> index;
    for (int i = 0; i < N; ++i) index[s->pts[i]] = i;
    if (match_ok) for (int i = 0; i < N; ++i) match_ok[i] = 0;
    if (px_refined) for (int i = 0; i < N; ++i) { Vector2d p(cur->w2c(s->pts[i]->pos_)); px_refined[2 * i] = p[0]; px_refined[2 * i + 1] = p[1]; }
    for (auto f : cur->fts_) {
      const int i = index[f->point];
      if (match_ok) match_ok[i] = 1;
      if (px_refined) { px_refined[2 * i] = f->px[0]; px_refined[2 * i + 1] = f->px[1]; }
    }
    st->n_pose_obs = st->n_matched;
    if (s->chain_pose_opt) {
      double scale = 0, e0 = 0, e1 = 0; size_t nobs = 0;
      pose_optimizer::optimizeGaussNewton(Config::poseOptimThresh(), Config::poseOptimNumIter(), false, cur, scale, e0, e1, nobs);
      st->n_pose_obs = (int)nobs;
      from_se3(cur->T_f_w_, st->T_cur_w);
    }
    // stationary benchmark workload: the per-point bookkeeping of the reprojector starts afresh every frame
    for (auto p : s->pts) { p->n_failed_reproj_ = 0; p->n_succeeded_reproj_ = 0; p->type_ = Point::TYPE_UNKNOWN; }
  } else {
  for (int i = 0; i < N; ++i) {
    Vector2d px(cur->w2c(s->pts[i]->pos_));                         // reprojector.cpp:131-145
    const bool ok = !s->pts[i]->obs_.empty() && s->matcher.findMatchDirect(*s->pts[i], *cur, px);   // (getCloseViewObs dereferences obs_.begin())
    st->n_matched += ok ? 1 : 0;
    if (px_refined) { px_refined[2 * i] = px[0]; px_refined[2 * i + 1] = px[1]; }
    if (match_ok) match_ok[i] = ok ? 1 : 0;
    if ((int)s->step_ok.size() != N) { s->step_ok.assign(N, 0); s->step_level.assign(N, 0); s->step_px.assign(2 * (size_t)N, 0.0); }
    s->step_px[2 * i] = px[0]; s->step_px[2 * i + 1] = px[1]; s->step_ok[i] = ok ? 1 : 0; s->step_level[i] = s->matcher.search_level_;
  }
  }
  s->conv_points.clear();
  const double t4 = now_s();
  if (s->tdf) while (s->tdf->pending() >= 2) std::this_thread::yield();   // never let the queue drop a frame (depth_filter.cpp:95-97)
  s->df->addFrame(cur);                                             // synchronous updateSeeds, or the queue of the filter thread
  const double t5 = now_s();
  s->timing[0] += t1 - t0; s->timing[1] += t3 - t2; s->timing[2] += t4 - t3; s->timing[3] += t5 - t4; s->timing[4] += 1.0;
  st->n_seeds_converged = (int)s->conv_points.size();
  st->n_seeds_updated = st->n_seeds_failed = st->n_seeds_skipped = -1;   // not observable through the reference API
  std::vector<Feature*> finished;
  for (auto p : s->conv_points) { Feature* f = p->obs_.front(); f->point = NULL; p->obs_.clear(); delete p; finished.push_back(f); }
  if (s->reseed == 3 || s->tdf) { s->last = cur; return; }         // list semantics of the reference / the filter thread owns the list
  if ((int)s->df->getSeeds().size() + (int)finished.size() != (int)s->seed_ftrs.size()) {
    std::map<Feature*, char> seen;                                   // seeds erased because z_inv_min was NaN
    for (auto& sd : s->df->getSeeds()) seen[sd.ftr] = 1;
    for (auto f : finished) seen[f] = 1;
    for (auto f : s->seed_ftrs) if (!seen.count(f)) finished.push_back(f);
  }
  if (s->reseed == 2) {
    // "young seed" regime: EVERY seed starts afresh every frame (long epipolar segments), in the original order
    s->df->getSeeds().clear();
    for (auto f : s->seed_ftrs) s->df->getSeeds().push_back(Seed(f, s->depth_mean, s->depth_min));
  } else if (s->reseed) for (auto f : finished) s->df->getSeeds().push_back(Seed(f, s->depth_mean, s->depth_min));
  s->last = cur;
}

// seconds per operator since the last call (5 doubles, see RefSeq::timing); resets the accumulators
void svo_ref_seq_get_timing(void* h, double* out)
{
  RefSeq* s = (RefSeq*)h;
  for (int k = 0; k < 5; ++k) { out[k] = s->timing[k]; s->timing[k] = 0; }
}

// state of every seed by original index; seeds no longer in the list get a = -1
void svo_ref_seq_get_seeds(void* h, float* out /*5 per seed*/)
{
  RefSeq* s = (RefSeq*)h;
  for (size_t i = 0; i < s->seed_ftrs.size(); ++i) out[5 * i] = -1.f;
  for (auto& sd : s->df->getSeeds()) {
    const int i = s->seed_index[sd.ftr];
    out[5 * i] = sd.a; out[5 * i + 1] = sd.b; out[5 * i + 2] = sd.mu; out[5 * i + 3] = sd.z_range; out[5 * i + 4] = sd.sigma2;
  }
}

}  // extern "C"
