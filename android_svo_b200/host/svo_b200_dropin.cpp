// svo_b200_dropin.cpp — the reference's hot-path C++ symbols, defined on top of libsvob200's C ABI.
//
// Compiled against the reference's UNCHANGED headers (-I /root/reference/app/src/main/cpp/svo/include
// and its vendored Eigen) and linked instead of vision.cpp / feature_alignment.cpp / matcher.cpp /
// sparse_img_align.cpp / feature_detection.cpp.  Everything arithmetic happens in CUDA kernels behind
// include/svob200.h; this file only marshals the reference's pointer-rich host model (Frame, Feature,
// Point, std::list<Seed>) into the flat arrays of the C ABI and writes the results back into the
// members the reference's callers read.  There is no CPU fallback: if no CUDA device is usable the
// first call throws std::runtime_error (the one exception type the reference itself uses, frame.cpp:55).
//
// Paths cited below are relative to /root/reference/app/src/main/cpp/svo.
#include <svo/global.h>
#include <svo/config.h>
#include <svo/vision.h>
#include <svo/aligned_mem.h>
#include <svo/abstract_camera.h>
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

#include <algorithm>
#include <cstdio>
#include <deque>
#include <list>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "svob200.h"
#include "svo_b200_dropin.h"

namespace svo {
namespace b200 {
namespace {

struct CacheEntry {
  const uint8_t* data0;      // level-0 pixel pointer of the host pyramid the mirror was made from
  int w, h, n_levels;
  unsigned long long stamp;
};

struct Runtime {
  std::recursive_mutex mu;   // one stream + one staging arena: calls from the tracking and depth-filter threads serialise here
  svob200_ctx* ctx = nullptr;
  std::unordered_map<int, CacheEntry> cache;          // Frame::id_ -> mirror
  std::unordered_map<long long, int64_t> scratch;     // (w,h,levels) -> temp frame id for bare cv::Mat arguments
  size_t capacity = 64;
  int pin_depth = 0;         // > 0 while a call is marshalling: no mirror is evicted until it is over (the cache may overshoot)
  unsigned long long clock = 0;
  int64_t next_temp_id = (int64_t)1 << 40;
};

// Two runtimes, each with its own svob200 context (= its own CUDA stream, staging arenas, scratch and frame mirrors): [0] serves
// the tracking thread's operators and, by default, the depth filter's updateSeeds too; with SVOB200_DROPIN_DF_CTX=1 updateSeeds —
// which the reference runs on a thread of its own (DepthFilter::startThread, depth_filter.cpp:63-103) — gets runtime [1], and the
// seed update of frame k and the alignment / reprojection of frame k+1 are in flight together on two streams instead of queueing
// behind one lock (a context is single-threaded, svob200.h; two contexts are independent).  Measured on one 640x480 sequence
// (bench.py --impl dropin --chain, ms per frame): two threads 0.80 with two contexts against 0.82 with one, ONE thread 0.70 against
// 0.61 — a frame both sides use is created and mirrored twice, and 0.12 ms of seed work per frame is too little to hide that.  So
// one context is the default; the second stream pays for batches, where the tracker's asynchronous depth filter provides it.
constexpr int N_RUNTIMES = 2;
thread_local int g_role = 0;

Runtime& runtime(int k)
{
  static Runtime r[N_RUNTIMES];
  return r[k];
}
Runtime& rt() { return runtime(g_role); }

int depth_filter_role()
{
  static const int v = [] { const char* e = getenv("SVOB200_DROPIN_DF_CTX"); return (e && atoi(e) == 1) ? 1 : 0; }();
  return v;
}

// everything the calling thread does through rt() until the scope ends goes to runtime `role`
struct RoleScope {
  int prev;
  explicit RoleScope(int role) : prev(g_role) { g_role = role; }
  ~RoleScope() { g_role = prev; }
};

[[noreturn]] void die(const char* what)
{
  Runtime& r = rt();
  std::string msg = std::string("svo_b200: ") + what + ": " + (r.ctx ? svob200_last_error(r.ctx) : "no usable CUDA device (there is no CPU fallback)");
  throw std::runtime_error(msg);
}

inline void check(int rc, const char* what) { if (rc != SVOB200_OK) die(what); }

svob200_ctx* ctx_locked()
{
  Runtime& r = rt();
  if (!r.ctx) {
    const char* dev = getenv("SVOB200_DEVICE");
    if (svob200_ctx_create(dev ? atoi(dev) : 0, &r.ctx) != SVOB200_OK) { r.ctx = nullptr; die("svob200_ctx_create"); }
  }
  return r.ctx;
}

inline int step_of(const cv::Mat& m) { return (int)m.step.p[0]; }

inline void pose7(const SE3& T, double* o)
{
  o[0] = T.get_translation().x; o[1] = T.get_translation().y; o[2] = T.get_translation().z;
  o[3] = T.get_rotation().x; o[4] = T.get_rotation().y; o[5] = T.get_rotation().z; o[6] = T.get_rotation().w;
}
inline SE3 se3_of(const double* T) { return SE3(T[0], T[1], T[2], T[3], T[4], T[5], T[6]); }

svob200_camera camera_of(const vk::AbstractCamera* cam)
{
  const vk::PinholeCamera* p = dynamic_cast<const vk::PinholeCamera*>(cam);
  if (!p) throw std::runtime_error("svo_b200: only vk::PinholeCamera is supported on the device path");
  if (p->d0() != 0.0 || p->d1() != 0.0 || p->d2() != 0.0 || p->d3() != 0.0 || p->d4() != 0.0)
    throw std::runtime_error("svo_b200: the device path implements the distortion-free pinhole branches only (pinhole_camera.cpp:48-53, :83-87)");
  svob200_camera c;
  c.width = p->width(); c.height = p->height(); c.fx = p->fx(); c.fy = p->fy(); c.cx = p->cx(); c.cy = p->cy();
  return c;
}

void matcher_opts_of(const Matcher::Options& o, svob200_matcher_opts* m)
{
  m->align_1d = o.align_1d ? 1 : 0;
  m->align_max_iter = o.align_max_iter;
  m->max_epi_search_steps = o.max_epi_search_steps > 0x7fffffffu ? 0x7fffffff : (int)o.max_epi_search_steps;
  m->subpix_refinement = o.subpix_refinement ? 1 : 0;
  m->epi_search_edgelet_filtering = o.epi_search_edgelet_filtering ? 1 : 0;
  m->epi_search_edgelet_max_angle = o.epi_search_edgelet_max_angle;
  m->max_search_level = (int)Config::nPyrLevels() - 1;            // matcher.cpp:174, :243
}

void evict_locked(Runtime& r)
{
  if (r.pin_depth > 0) return;
  while (r.cache.size() > r.capacity) {
    auto victim = r.cache.begin();
    for (auto it = r.cache.begin(); it != r.cache.end(); ++it) if (it->second.stamp < victim->second.stamp) victim = it;
    svob200_frame_release(r.ctx, victim->first);
    r.cache.erase(victim);
  }
}

// SVOB200_DROPIN_REBUILD_PYRAMID=0: mirror a Frame's pyramid level by level, exactly as the host holds it (five uploads, each with
// its synchronisation: ~0.17 ms per VGA frame), for callers whose Frames carry pyramids that did not come from createImgPyramid.
// Default: upload level 0 and rebuild the levels on the device (one upload + one kernel).
bool rebuild_pyramids()
{
  static const bool v = [] { const char* e = getenv("SVOB200_DROPIN_REBUILD_PYRAMID"); return !(e && atoi(e) == 0); }();
  return v;
}

// Mirror a host pyramid on the device under `id`, level by level, exactly as the host holds it.
void upload_pyramid_locked(svob200_ctx* ctx, int64_t id, const ImgPyr& pyr, int n_levels)
{
  for (int l = 0; l < n_levels; ++l)
    check(svob200_frame_upload_level(ctx, id, 0, l, pyr[l].data, step_of(pyr[l])), "svob200_frame_upload_level");
}

bool pyramid_is_halving(const ImgPyr& pyr, int n_levels)
{
  for (int l = 1; l < n_levels; ++l)
    if (pyr[l].cols != pyr[l - 1].cols / 2 || pyr[l].rows != pyr[l - 1].rows / 2) return false;
  return true;
}

// Device mirror of a Frame's pyramid, keyed by Frame::id_ (images are immutable after the ctor).
int64_t ensure_frame_locked(const Frame& f)
{
  Runtime& r = rt();
  svob200_ctx* ctx = ctx_locked();
  const int n_levels = (int)f.img_pyr_.size();
  if (n_levels < 1 || n_levels > SVOB200_MAX_LEVELS) throw std::runtime_error("svo_b200: frame pyramid depth outside [1, 8]");
  if (!pyramid_is_halving(f.img_pyr_, n_levels)) throw std::runtime_error("svo_b200: frame pyramid is not an integer-halving pyramid (frame.cpp:192)");
  const cv::Mat& l0 = f.img_pyr_[0];
  auto it = r.cache.find(f.id_);
  if (it != r.cache.end()) {
    CacheEntry& e = it->second;
    if (e.data0 == l0.data && e.w == l0.cols && e.h == l0.rows && e.n_levels == n_levels) { e.stamp = ++r.clock; return f.id_; }
    svob200_frame_release(ctx, f.id_);
    r.cache.erase(it);
  }
  check(svob200_frame_create(ctx, f.id_, 1, l0.cols, l0.rows, n_levels), "svob200_frame_create");
  if (rebuild_pyramids())
    // ONE upload (level 0) and the fused pyramid kernel: the levels the host holds came out of createImgPyramid -> vk::halfSample,
    // which in a process linked with this file is the same kernel with the same rounding rule — the mirror is bit-identical
    check(svob200_frame_upload(ctx, f.id_, l0.data, step_of(l0), nullptr, SVOB200_MEM_HOST), "svob200_frame_upload");
  else
    upload_pyramid_locked(ctx, f.id_, f.img_pyr_, n_levels);
  r.cache[f.id_] = CacheEntry{l0.data, l0.cols, l0.rows, n_levels, ++r.clock};
  evict_locked(r);
  return f.id_;
}

// Temp device frame for operators that receive bare images (align1D/2D, FastDetector::detect's img_pyr).
int64_t scratch_frame_locked(int w, int h, int n_levels)
{
  Runtime& r = rt();
  svob200_ctx* ctx = ctx_locked();
  const long long key = ((long long)w << 36) | ((long long)h << 8) | n_levels;
  auto it = r.scratch.find(key);
  if (it != r.scratch.end()) return it->second;
  if (r.scratch.size() >= 16) {                       // bounded: drop all and start over
    for (auto& kv : r.scratch) svob200_frame_release(ctx, kv.second);
    r.scratch.clear();
  }
  const int64_t id = r.next_temp_id++;
  check(svob200_frame_create(ctx, id, 1, w, h, n_levels), "svob200_frame_create");
  r.scratch[key] = id;
  return id;
}

thread_local int g_last_iters[SVOB200_MAX_LEVELS] = {0, 0, 0, 0, 0, 0, 0, 0};
thread_local int g_last_exact = 0;

typedef std::lock_guard<std::recursive_mutex> Lock;

// Every frame a call touches stays resident until the svob200_* call that uses it has returned: a call may reference more
// distinct frames than the cache holds (Reprojector::reprojectMap walks the observation lists of every candidate point, and
// the reference's Android Config keeps an unbounded map), and evicting while marshalling would release a mirror the same call
// already referenced.  Construct under the runtime lock; the eviction runs when the outermost scope ends.
struct PinFrames {
  Runtime& r;
  PinFrames() : r(rt()) { ++r.pin_depth; }
  ~PinFrames() { if (--r.pin_depth == 0 && r.ctx) evict_locked(r); }
};

}  // namespace

svob200_ctx* context() { Lock lk(rt().mu); return ctx_locked(); }

void releaseFrame(const Frame& frame)
{
  for (int k = 0; k < N_RUNTIMES; ++k) {
    Runtime& r = runtime(k);
    Lock lk(r.mu);
    auto it = r.cache.find(frame.id_);
    if (it == r.cache.end() || !r.ctx) continue;
    svob200_frame_release(r.ctx, frame.id_);
    r.cache.erase(it);
  }
}

static int g_seed_chunk = 2048;
int seedChunk() { return g_seed_chunk; }
void setSeedChunk(int seeds_per_call) { g_seed_chunk = seeds_per_call < 64 ? 64 : seeds_per_call; }

void setFrameCacheCapacity(size_t capacity)
{
  for (int k = 0; k < N_RUNTIMES; ++k) {
    Runtime& r = runtime(k);
    Lock lk(r.mu);
    r.capacity = capacity < 4 ? 4 : capacity;
    if (r.ctx) evict_locked(r);
  }
}

void shutdown()
{
  for (int k = 0; k < N_RUNTIMES; ++k) {
    Runtime& r = runtime(k);
    Lock lk(r.mu);
    if (!r.ctx) continue;
    svob200_ctx_destroy(r.ctx);     // frees every resident frame
    r.ctx = nullptr; r.cache.clear(); r.scratch.clear();
  }
}

long long launchCount()
{
  long long n = 0;
  for (int k = 0; k < N_RUNTIMES; ++k) { Runtime& r = runtime(k); Lock lk(r.mu); if (r.ctx) n += svob200_ctx_launch_count(r.ctx); }
  return n;
}
const int* lastAlignIterations() { return g_last_iters; }
int lastAlignExactChi2() { return g_last_exact; }

}  // namespace b200
}  // namespace svo

using svo::b200::rt;
using svo::b200::Lock;
using svo::b200::check;
using svo::b200::ctx_locked;
using svo::b200::step_of;

// ============================================================================ vision.h
namespace vk {

// vision.cpp:71-110.  The rounding the reference computes depends on the ISA it was compiled for:
// SSE2 builds take the double-rounding _mm_avg path iff both buffers are 16-byte aligned and
// in.cols % 16 == 0 (:78); every other case (NEON, scalar) truncates (a+b+c+d)/4.
void halfSample(const cv::Mat& in, cv::Mat& out)
{
  assert(in.rows / 2 == out.rows && in.cols / 2 == out.cols);
  assert(in.type() == CV_8U && out.type() == CV_8U);
  int mode = SVOB200_ROUND_TRUNC;
#ifdef __SSE2__
  if (aligned_mem::is_aligned16(in.data) && aligned_mem::is_aligned16(out.data) && ((in.cols % 16) == 0)) mode = SVOB200_ROUND_SSE2;
#endif
  Lock lk(rt().mu);
  check(svob200_half_sample(ctx_locked(), in.data, in.cols, in.rows, step_of(in), out.data, step_of(out), mode), "svob200_half_sample");
}

// vision.cpp:113-154
float shiTomasiScore(const cv::Mat& img, int u, int v)
{
  assert(img.type() == CV_8UC1);
  const int uv[2] = {u, v};
  float score = 0.f;
  Lock lk(rt().mu);
  check(svob200_shi_tomasi(ctx_locked(), img.data, img.cols, img.rows, step_of(img), 1, uv, &score), "svob200_shi_tomasi");
  return score;
}

}  // namespace vk

namespace svo {

// ============================================================================ feature_alignment.h
namespace feature_alignment {

namespace {
bool align_on_device(const cv::Mat& cur_img, const float* dir, uint8_t* ref_patch_with_border, uint8_t* ref_patch, int n_iter,
                     Vector2d& cur_px_estimate, double* h_inv)
{
  Lock lk(rt().mu);
  svob200_ctx* ctx = ctx_locked();
  const int64_t id = b200::scratch_frame_locked(cur_img.cols, cur_img.rows, 1);
  check(svob200_frame_upload_level(ctx, id, 0, 0, cur_img.data, step_of(cur_img)), "svob200_frame_upload_level");
  const int image = 0;
  double px[2] = {cur_px_estimate[0], cur_px_estimate[1]};
  int converged = 0;
  check(svob200_align_patches(ctx, id, 0, 1, &image, ref_patch_with_border, ref_patch, dir, n_iter, px, &converged, h_inv, SVOB200_MEM_HOST),
        "svob200_align_patches");
  cur_px_estimate << px[0], px[1];
  return converged != 0;
}
}  // namespace

// feature_alignment.cpp:35-152
bool align1D(const cv::Mat& cur_img, const Vector2f& dir, uint8_t* ref_patch_with_border, uint8_t* ref_patch, const int n_iter,
             Vector2d& cur_px_estimate, double& h_inv)
{
  const float d[2] = {dir[0], dir[1]};
  return align_on_device(cur_img, d, ref_patch_with_border, ref_patch, n_iter, cur_px_estimate, &h_inv);
}

// feature_alignment.cpp:154-282 (float path; the SSE2/NEON integer variants are compiled out of the
// reference's align2D by its `no_simd` default and are not on the path)
bool align2D(const cv::Mat& cur_img, uint8_t* ref_patch_with_border, uint8_t* ref_patch, const int n_iter, Vector2d& cur_px_estimate, bool)
{
  return align_on_device(cur_img, nullptr, ref_patch_with_border, ref_patch, n_iter, cur_px_estimate, nullptr);
}

}  // namespace feature_alignment

// ============================================================================ matcher.h: warp
namespace warp {

// matcher.cpp:36-60
void getWarpMatrixAffine(const vk::AbstractCamera& cam_ref, const vk::AbstractCamera& /*cam_cur: same camera*/, const Vector2d& px_ref,
                         const Vector3d& f_ref, const double depth_ref, const SE3& T_cur_ref, const int level_ref, Matrix2d& A_cur_ref)
{
  const svob200_camera cam = b200::camera_of(&cam_ref);
  const double px[2] = {px_ref[0], px_ref[1]}, f[3] = {f_ref[0], f_ref[1], f_ref[2]};
  double T[7], A[4];
  b200::pose7(T_cur_ref, T);
  Lock lk(rt().mu);
  check(svob200_warp_matrix_affine(ctx_locked(), &cam, 1, px, f, &depth_ref, T, &level_ref, A, SVOB200_MEM_HOST), "svob200_warp_matrix_affine");
  A_cur_ref << A[0], A[1], A[2], A[3];
}

// matcher.cpp:65-78 — three lines of integer control on a determinant, kept on the host
int getBestSearchLevel(const Matrix2d& A_cur_ref, const int max_level)
{
  int search_level = 0;
  double D = A_cur_ref(0, 0) * A_cur_ref(1, 1) - A_cur_ref(1, 0) * A_cur_ref(0, 1);
  while (D > 3.0 && search_level < max_level) { search_level += 1; D *= 0.25; }
  return search_level;
}

// matcher.cpp:83-116
void warpAffine(const Matrix2d& A_cur_ref, const cv::Mat& img_ref, const Vector2d& px_ref, const int level_ref, const int search_level,
                const int halfpatch_size, uint8_t* patch)
{
  const double A[4] = {A_cur_ref(0, 0), A_cur_ref(0, 1), A_cur_ref(1, 0), A_cur_ref(1, 1)}, px[2] = {px_ref[0], px_ref[1]};
  Lock lk(rt().mu);
  check(svob200_warp_affine(ctx_locked(), img_ref.data, img_ref.cols, img_ref.rows, step_of(img_ref), A, px, level_ref, search_level,
                            halfpatch_size, patch), "svob200_warp_affine");
}

}  // namespace warp

// matcher.cpp:123-136 (not declared in matcher.h, but an external symbol of matcher.cpp)
bool depthFromTriangulation(const SE3& T_search_ref, const Vector3d& f_ref, const Vector3d& f_cur, double& depth)
{
  double T[7];
  b200::pose7(T_search_ref, T);
  const double a[3] = {f_ref[0], f_ref[1], f_ref[2]}, b[3] = {f_cur[0], f_cur[1], f_cur[2]};
  int ok = 0;
  Lock lk(rt().mu);
  check(svob200_depth_from_triangulation(ctx_locked(), 1, T, a, b, &depth, &ok), "svob200_depth_from_triangulation");
  return ok != 0;
}

// ============================================================================ matcher.h: Matcher
namespace {

void fill_feature_ref(const Feature& ftr, int64_t ref_frame_id, const SE3& T_cur_ref, svob200_feature_ref* f)
{
  memset(f, 0, sizeof(*f));
  f->ref_frame_id = ref_frame_id; f->ref_image = 0; f->cur_image = 0;
  f->level = ftr.level; f->type = ftr.type == Feature::EDGELET ? 1 : 0;
  f->px[0] = ftr.px[0]; f->px[1] = ftr.px[1];
  f->f[0] = ftr.f[0]; f->f[1] = ftr.f[1]; f->f[2] = ftr.f[2];
  f->grad[0] = ftr.grad[0]; f->grad[1] = ftr.grad[1];
  b200::pose7(T_cur_ref, f->T_cur_ref);
}

}  // namespace

// matcher.cpp:138-147
void Matcher::createPatchFromPatchWithBorder()
{
  for (int y = 0; y < patch_size_; ++y)
    memcpy(patch_ + y * patch_size_, patch_with_border_ + (y + 1) * (patch_size_ + 2) + 1, patch_size_);
}

// matcher.cpp:156-202
bool Matcher::findMatchDirect(const Point& pt, const Frame& cur_frame, Vector2d& px_cur)
{
  if (!pt.getCloseViewObs(cur_frame.pos(), ref_ftr_)) return false;                       // host: picks the observation (point.cpp:101-125)
  const Vector2i pxi(ref_ftr_->px.cast<int>() / (1 << ref_ftr_->level));
  if (!ref_ftr_->frame->cam_->isInFrame(pxi, halfpatch_size_ + 2, ref_ftr_->level)) return false;

  const svob200_camera cam = b200::camera_of(cur_frame.cam_);
  svob200_matcher_opts mo;
  b200::matcher_opts_of(options_, &mo);
  const double depth_ref = (ref_ftr_->frame->pos() - pt.pos_).norm();
  const double px_in[2] = {px_cur[0], px_cur[1]};
  svob200_match_result res;
  {
    Lock lk(rt().mu);
    b200::PinFrames pin;
    svob200_ctx* ctx = ctx_locked();
    const int64_t ref_id = b200::ensure_frame_locked(*ref_ftr_->frame);
    const int64_t cur_id = b200::ensure_frame_locked(cur_frame);
    svob200_feature_ref f;
    fill_feature_ref(*ref_ftr_, ref_id, cur_frame.T_f_w_ * ref_ftr_->frame->T_f_w_.inverse(), &f);
    check(svob200_match_direct(ctx, cur_id, &cam, 1, &f, &depth_ref, px_in, &mo, &res, SVOB200_MEM_HOST), "svob200_match_direct");
  }
  A_cur_ref_ << res.A_cur_ref[0], res.A_cur_ref[1], res.A_cur_ref[2], res.A_cur_ref[3];
  search_level_ = res.search_level;
  memcpy(patch_with_border_, res.patch_with_border, sizeof(patch_with_border_));
  memcpy(patch_, res.patch, sizeof(patch_));
  if (ref_ftr_->type == Feature::EDGELET) h_inv_ = res.h_inv;
  px_cur << res.px_cur[0], res.px_cur[1];                                                 // written whether or not LK converged (:200)
  return res.success != 0;
}

// matcher.cpp:207-355
bool Matcher::findEpipolarMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const Feature& ref_ftr, const double d_estimate,
                                      const double d_min, const double d_max, double& depth)
{
  const svob200_camera cam = b200::camera_of(cur_frame.cam_);
  svob200_matcher_opts mo;
  b200::matcher_opts_of(options_, &mo);
  const SE3 T_cur_ref = cur_frame.T_f_w_ * ref_frame.T_f_w_.inverse();
  const double d[3] = {d_estimate, d_min, d_max};
  svob200_epi_result res;
  {
    Lock lk(rt().mu);
    b200::PinFrames pin;
    svob200_ctx* ctx = ctx_locked();
    const int64_t ref_id = b200::ensure_frame_locked(ref_frame);
    const int64_t cur_id = b200::ensure_frame_locked(cur_frame);
    svob200_feature_ref f;
    fill_feature_ref(ref_ftr, ref_id, T_cur_ref, &f);
    check(svob200_epipolar_match(ctx, cur_id, &cam, 1, &f, d, &mo, &res, SVOB200_MEM_HOST), "svob200_epipolar_match");
  }
  epi_dir_ << res.epi_dir[0], res.epi_dir[1];
  A_cur_ref_ << res.A_cur_ref[0], res.A_cur_ref[1], res.A_cur_ref[2], res.A_cur_ref[3];
  reject_ = res.reject != 0;
  if (reject_) return false;                                                              // :236-239: nothing else is touched
  search_level_ = res.search_level;
  epi_length_ = res.epi_length;
  memcpy(patch_with_border_, res.patch_with_border, sizeof(patch_with_border_));
  memcpy(patch_, res.patch, sizeof(patch_));
  if (res.px_cur_valid) px_cur_ << res.px_cur[0], res.px_cur[1];
  if (options_.align_1d && res.h_inv != 0.0) h_inv_ = res.h_inv;
  if (!res.success) return false;
  depth = res.depth;
  return true;
}

// ============================================================================ sparse_img_align.h
// sparse_img_align.cpp:29-41
SparseImgAlign::SparseImgAlign(int max_level, int min_level, int n_iter, Method method, bool display, bool verbose)
  : display_(display), max_level_(max_level), min_level_(min_level)
{
  n_iter_ = n_iter;
  n_iter_init_ = n_iter_;
  method_ = method;
  verbose_ = verbose;
  eps_ = 0.000001;
}

// sparse_img_align.cpp:51-92.  precomputeReferencePatches / computeResiduals / solve / update and the
// whole vk::NLLSSolver::optimizeGaussNewton loop (nlls_solver_impl.hpp:25-100) run inside ONE kernel launch.
size_t SparseImgAlign::run(FramePtr ref_frame, FramePtr cur_frame)
{
  reset();
  if (ref_frame->fts_.empty()) {
    fprintf(stderr, "[svo_b200] SparseImgAlign: no features to track!\n");
    return 0;
  }
  if (method_ != GaussNewton) throw std::runtime_error("svo_b200: SparseImgAlign runs Gauss-Newton only (the one method the reference uses, frame_handler_mono.cpp:186-187)");
  ref_frame_ = ref_frame;
  cur_frame_ = cur_frame;
  const int N = (int)ref_frame_->fts_.size();
  visible_fts_.resize(N, false);
  have_ref_patch_cache_ = false;

  // flatten the feature list (sparse_img_align.cpp:114-134: depth = |pos - ref_pos|, xyz_ref = f * depth)
  std::vector<double> px(2 * (size_t)N), xyz(3 * (size_t)N, 0.0);
  std::vector<uint8_t> has_point(N, 0);
  const Vector3d ref_pos = ref_frame_->pos();
  {
    int i = 0;
    for (auto it = ref_frame_->fts_.begin(); it != ref_frame_->fts_.end(); ++it, ++i) {
      const Feature* ftr = *it;
      px[2 * i] = ftr->px[0]; px[2 * i + 1] = ftr->px[1];
      if (ftr->point == NULL) continue;
      has_point[i] = 1;
      const double depth = (ftr->point->pos_ - ref_pos).norm();
      const Vector3d xyz_ref(ftr->f * depth);
      xyz[3 * i] = xyz_ref[0]; xyz[3 * i + 1] = xyz_ref[1]; xyz[3 * i + 2] = xyz_ref[2];
    }
  }
  SE3 T_cur_from_ref(cur_frame_->T_f_w_ * ref_frame_->T_f_w_.inverse());
  double T_init[7];
  b200::pose7(T_cur_from_ref, T_init);
  const svob200_camera cam = b200::camera_of(ref_frame_->cam_);
  svob200_align_opts opts;
  opts.max_level = max_level_; opts.min_level = min_level_; opts.n_iter = (int)n_iter_; opts.eps = eps_;
  const int offsets[2] = {0, N};
  svob200_align_result res;
  {
    Lock lk(rt().mu);
    b200::PinFrames pin;
    svob200_ctx* ctx = ctx_locked();
    const int64_t ref_id = b200::ensure_frame_locked(*ref_frame_);
    const int64_t cur_id = b200::ensure_frame_locked(*cur_frame_);
    check(svob200_sparse_align(ctx, ref_id, cur_id, &cam, 1, offsets, px.data(), xyz.data(), has_point.data(), T_init, &opts, &res, SVOB200_MEM_HOST),
          "svob200_sparse_align");
  }
  // solver state the reference leaves behind (read by getFisherInformation and by callers of NLLSSolver's public members)
  for (int r = 0; r < 6; ++r) {
    for (int c = 0; c < 6; ++c) H_(r, c) = res.H[r * 6 + c];
    Jres_[r] = res.Jres[r];
    x_[r] = res.x[r];
  }
  chi2_ = res.chi2;
  n_meas_ = (size_t)res.n_meas;
  stop_ = res.stop != 0;
  level_ = min_level_ - 1;                                   // value of the loop variable after run() (:65)
  iter_ = 0;
  for (int l = 0; l < SVOB200_MAX_LEVELS; ++l) { b200::g_last_iters[l] = res.iters[l]; if (l == min_level_ && res.iters[l] > 0) iter_ = (size_t)res.iters[l] - 1; }
  b200::g_last_exact = res.n_exact_chi2;

  T_cur_from_ref = b200::se3_of(res.T_cur_ref);
  cur_frame_->T_f_w_ = T_cur_from_ref * ref_frame_->T_f_w_;  // :89
  return n_meas_ / patch_area_;
}

// sparse_img_align.cpp:94-99
Matrix<double, 6, 6> SparseImgAlign::getFisherInformation()
{
  const double sigma_i_sq = 5e-4 * 255 * 255;
  Matrix<double, 6, 6> I = H_ / sigma_i_sq;
  return I;
}

// NLLSSolver's per-iteration hooks (nlls_solver.h:63-87).  The Gauss-Newton loop lives on the device, so
// run() never calls them; they exist because the class declares them (vtable) and keep the solver's contract
// for a caller that drives vk::NLLSSolver::optimize() by hand: that route is not accelerated and says so.
void SparseImgAlign::precomputeReferencePatches() {}
double SparseImgAlign::computeResiduals(const SE3&, bool, bool)
{
  throw std::runtime_error("svo_b200: SparseImgAlign::computeResiduals is fused into run() on the device; call run()");
}
int SparseImgAlign::solve()
{
  x_ = H_.ldlt().solve(Jres_);                               // sparse_img_align.cpp:291-297
  if ((bool)std::isnan((double)x_[0])) return 0;
  return 1;
}
void SparseImgAlign::update(const ModelType& T_curold_from_ref, ModelType& T_curnew_from_ref)
{
  const Matrix<double, 6, 1> neg = -x_;
  T_curnew_from_ref = T_curold_from_ref * SE3::exp(neg.data());   // sparse_img_align.cpp:302-308
}
void SparseImgAlign::startIteration() {}
void SparseImgAlign::finishIteration() {}

// ============================================================================ feature_detection.h
namespace feature_detection {

// feature_detection.cpp:24-64 — grid bookkeeping (host state of the detector object)
AbstractDetector::AbstractDetector(const int img_width, const int img_height, const int cell_size, const int n_pyr_levels)
  : cell_size_(cell_size), n_pyr_levels_(n_pyr_levels),
    grid_n_cols_(ceil(static_cast<double>(img_width) / cell_size_)),
    grid_n_rows_(ceil(static_cast<double>(img_height) / cell_size_)),
    grid_occupancy_(grid_n_cols_ * grid_n_rows_, false)
{}

void AbstractDetector::resetGrid() { std::fill(grid_occupancy_.begin(), grid_occupancy_.end(), false); }

void AbstractDetector::setExistingFeatures(const Features& fts)
{
  for (auto it = fts.begin(); it != fts.end(); ++it) setGridOccpuancy((*it)->px);
}

void AbstractDetector::setGridOccpuancy(const Vector2d& px)
{
  grid_occupancy_.at(static_cast<int>(px[1] / cell_size_) * grid_n_cols_ + static_cast<int>(px[0] / cell_size_)) = true;
}

FastDetector::FastDetector(const int img_width, const int img_height, const int cell_size, const int n_pyr_levels)
  : AbstractDetector(img_width, img_height, cell_size, n_pyr_levels)
{}

// feature_detection.cpp:77-122: cv::FAST + shiTomasiScore + per-cell strict-> argmax over all levels in one pass on the device
void FastDetector::detect(Frame* frame, const ImgPyr& img_pyr, const double detection_threshold, Features& fts)
{
  if ((int)img_pyr.size() < n_pyr_levels_) throw std::runtime_error("svo_b200: FastDetector::detect: pyramid has fewer levels than n_pyr_levels");
  if (!b200::pyramid_is_halving(img_pyr, n_pyr_levels_)) throw std::runtime_error("svo_b200: FastDetector::detect: not an integer-halving pyramid");
  const int n_cells = grid_n_cols_ * grid_n_rows_;
  const int w = img_pyr[0].cols, h = img_pyr[0].rows;
  if ((w + cell_size_ - 1) / cell_size_ != grid_n_cols_ || (h + cell_size_ - 1) / cell_size_ != grid_n_rows_)
    throw std::runtime_error("svo_b200: FastDetector::detect: image size does not match the detector grid");
  std::vector<uint8_t> occ(n_cells);
  for (int k = 0; k < n_cells; ++k) occ[k] = grid_occupancy_[k] ? 1 : 0;
  std::vector<svob200_corner> cells(n_cells);
  int n_features = 0;
  {
    Lock lk(rt().mu);
    svob200_ctx* ctx = ctx_locked();
    const int64_t id = b200::scratch_frame_locked(w, h, n_pyr_levels_);
    b200::upload_pyramid_locked(ctx, id, img_pyr, n_pyr_levels_);
    check(svob200_fast_detect(ctx, id, n_pyr_levels_, cell_size_, detection_threshold, occ.data(), cells.data(), &n_features, SVOB200_MEM_HOST),
          "svob200_fast_detect");
  }
  for (int k = 0; k < n_cells; ++k) {                          // cell-index order, like the reference's for_each (:115-119)
    const svob200_corner& c = cells[k];
    if ((double)c.score > detection_threshold) fts.push_back(new Feature(frame, Vector2d(c.x, c.y), c.level));
  }
  resetGrid();
}

}  // namespace feature_detection

// ============================================================================ depth_filter.h
// depth_filter.cpp:237-341.  Every seed's update depends only on its own state, the two frames and the
// two poses, so the sequential list walk is a batch: flatten -> three launches -> apply in list order.
void B200DepthFilter::updateSeeds(FramePtr frame)
{
  svo::b200::RoleScope role(svo::b200::depth_filter_role());          // the depth filter's own context and stream
  lock_t lock(seeds_mut_);
  last_n_seeds_ = seeds_.size(); last_n_updates_ = 0; last_n_failed_matches_ = 0;
  if (seeds_updating_halt_) return;                                   // the reference polls the flag per seed (:253-254); a batch polls it once

  // too-old seeds go first (:258-261); the reference erases them lazily while walking the list — same result
  for (auto it = seeds_.begin(); it != seeds_.end();) {
    if ((Seed::batch_counter - it->batch_id) > options_.max_n_kfs) it = seeds_.erase(it);
    else ++it;
  }
  const int n = (int)seeds_.size();
  if (n == 0) return;

  const svob200_camera cam = b200::camera_of(frame->cam_);
  svob200_matcher_opts mo;
  b200::matcher_opts_of(matcher_.options_, &mo);
  std::vector<svob200_feature_ref> ftrs(n);
  std::vector<double> T_ref_w(7 * (size_t)n);
  std::vector<svob200_seed> state(n);
  std::vector<svob200_seed_obs> obs(n);
  double T_cur_w[7];
  b200::pose7(frame->T_f_w_, T_cur_w);
  // The list goes to the device in chunks of b200::seedChunk() seeds (default 2,048).  Between two chunks seeds_updating_halt_ is
  // polled again — addKeyframe / removeKeyframe / reset set it and then wait for seeds_mut_ (depth_filter.cpp:112-117, :153-170):
  // the reference returns at the next seed (:253-254), this returns at the next chunk — and the runtime lock is released, so
  // the tracking thread's calls (SparseImgAlign::run, reprojectMap) interleave with a long update instead of queueing behind it.
  int n_done = 0;
  {
    auto it = seeds_.begin();
    const int chunk = b200::seedChunk();
    while (n_done < n) {
      if (n_done > 0 && seeds_updating_halt_) break;
      const int m = std::min(chunk, n - n_done);
      Lock lk(rt().mu);
      b200::PinFrames pin;
      svob200_ctx* ctx = ctx_locked();
      const int64_t cur_id = b200::ensure_frame_locked(*frame);
      for (int i = n_done; i < n_done + m; ++i, ++it) {
        const int64_t ref_id = b200::ensure_frame_locked(*it->ftr->frame);
        fill_feature_ref(*it->ftr, ref_id, SE3(), &ftrs[i]);              // poses travel separately (T_ref_w / T_cur_w)
        b200::pose7(it->ftr->frame->T_f_w_, &T_ref_w[7 * (size_t)i]);
        state[i].a = it->a; state[i].b = it->b; state[i].mu = it->mu; state[i].z_range = it->z_range; state[i].sigma2 = it->sigma2;
      }
      check(svob200_seeds_update(ctx, cur_id, &cam, m, ftrs.data() + n_done, T_ref_w.data() + 7 * (size_t)n_done, T_cur_w, &mo,
                                 options_.seed_convergence_sigma2_thresh, state.data() + n_done, obs.data() + n_done, SVOB200_MEM_HOST),
            "svob200_seeds_update");
      n_done += m;
    }
  }

  int i = 0;
  for (auto it = seeds_.begin(); it != seeds_.end() && i < n_done; ++i) {     // (a halted update leaves the rest of the list untouched, like :253-254)
    const svob200_seed_obs& o = obs[i];
    if (o.status == SVOB200_SEED_BEHIND || o.status == SVOB200_SEED_NOT_IN_FRAME) { ++it; continue; }     // :266-273
    it->a = state[i].a; it->b = state[i].b; it->mu = state[i].mu; it->sigma2 = state[i].sigma2;
    if (o.status == SVOB200_SEED_NO_MATCH) { ++last_n_failed_matches_; ++it; continue; }                  // b++ happened on the device (:286)
    ++last_n_updates_;
    if (frame->isKeyframe() && feature_detector_)
      feature_detector_->setGridOccpuancy(Vector2d(o.px_cur[0], o.px_cur[1]));                          // :306-310
    if (o.status == SVOB200_SEED_CONVERGED) {                                                            // :314-337
      assert(it->ftr->point == NULL);
      Vector3d xyz_world(it->ftr->frame->T_f_w_.inverse() * (it->ftr->f * (1.0 / it->mu)));
      Point* point = new Point(xyz_world, it->ftr);
      it->ftr->point = point;
      seed_converged_cb_(point, it->sigma2);
      it = seeds_.erase(it);
    } else if (o.status == SVOB200_SEED_NAN_ERASED) {                                                    // :338-342
      fprintf(stderr, "[svo_b200] z_min is NaN\n");
      it = seeds_.erase(it);
    } else {
      ++it;
    }
  }
}


// ============================================================================ reprojector.h: Reprojector
// Linked INSTEAD OF reprojector.cpp.  The walk over the map that decides WHICH points are candidates (close keyframes by
// distance, one projection per point, the candidate list under its mutex: reprojector.cpp:76-146) is host control plane
// and stays here; the geometry, the grid, every findMatchDirect and the per-cell / maxFts rules run on the device in one
// svob200_reproject_map call; the side effects on Point / Map / Frame are applied afterwards in the reference's order.
Reprojector::Reprojector(vk::AbstractCamera* cam, Map& map) : map_(map) { initializeGrid(cam); }

Reprojector::~Reprojector() { std::for_each(grid_.cells.begin(), grid_.cells.end(), [&](Cell* c) { delete c; }); }

void Reprojector::initializeGrid(vk::AbstractCamera* cam)      // reprojector.cpp:43-54 (the cell lists stay empty: the grid lives on the device)
{
  grid_.cell_size = Config::gridSize();
  grid_.grid_n_cols = ceil(static_cast<double>(cam->width()) / grid_.cell_size);
  grid_.grid_n_rows = ceil(static_cast<double>(cam->height()) / grid_.cell_size);
  grid_.cells.resize(grid_.grid_n_cols * grid_.grid_n_rows);
  std::for_each(grid_.cells.begin(), grid_.cells.end(), [&](Cell*& c) { c = new Cell; });
  grid_.cell_order.resize(grid_.cells.size());
  for (size_t i = 0; i < grid_.cells.size(); ++i) grid_.cell_order[i] = i;
}

void Reprojector::resetGrid()
{
  n_matches_ = 0;
  n_trials_ = 0;
  std::for_each(grid_.cells.begin(), grid_.cells.end(), [&](Cell* c) { c->clear(); });
}

bool Reprojector::pointQualityComparator(Candidate& lhs, Candidate& rhs) { return lhs.pt->type_ > rhs.pt->type_; }
bool Reprojector::reprojectCell(Cell&, FramePtr) { return false; }       // the cell loop runs on the device
bool Reprojector::reprojectPoint(FramePtr, Point*) { return false; }     // the projection runs on the device

void Reprojector::reprojectMap(FramePtr frame, std::vector<std::pair<FramePtr, std::size_t> >& overlap_kfs)
{
  resetGrid();
  // ---- which points (host, reprojector.cpp:78-119 and :125-131)
  // nearest keyframes first, at most options_.max_n_kfs of them; a point seen from several of them is taken once (the stamp
  // last_projected_kf_id_ carries this frame's id from the first time on)
  typedef std::pair<FramePtr, double> KfDist;
  std::list<KfDist> nearby;
  map_.getCloseKeyframes(frame, nearby);
  std::vector<KfDist> ranked(nearby.begin(), nearby.end());
  std::stable_sort(ranked.begin(), ranked.end(), [](const KfDist& a, const KfDist& b) { return a.second < b.second; });
  if (ranked.size() > options_.max_n_kfs) ranked.resize(options_.max_n_kfs);
  std::vector<Point*> pts;
  std::vector<int> owner;                       // index into overlap_kfs, -1 for point candidates
  overlap_kfs.reserve(options_.max_n_kfs);
  for (const KfDist& kd : ranked) {
    const int k = (int)overlap_kfs.size();
    overlap_kfs.push_back(std::make_pair(kd.first, (std::size_t)0));
    for (Feature* ftr : kd.first->fts_) {
      Point* pt = ftr->point;
      if (!pt || pt->last_projected_kf_id_ == frame->id_) continue;
      pt->last_projected_kf_id_ = frame->id_;
      pts.push_back(pt);
      owner.push_back(k);
    }
  }
  std::unique_lock<std::mutex> cand_lock(map_.point_candidates_.mut_);
  const size_t n_map = pts.size();
  for (auto it = map_.point_candidates_.candidates_.begin(); it != map_.point_candidates_.candidates_.end(); ++it) { pts.push_back(it->first); owner.push_back(-1); }
  const int n_points = (int)pts.size();
  const int n_cells = grid_.grid_n_cols * grid_.grid_n_rows;
  std::vector<svob200_reproj_result> res(n_points);
  std::vector<int> winner(n_cells, -1);
  svob200_reproj_stats stats = {0, 0, 0, 0};
  std::vector<Feature*> obs_ftr;
  if (n_points > 0) {
    // ---- flat records: points in insertion order, observations in Point::obs_ list order
    std::vector<svob200_map_point> mp(n_points);
    std::vector<svob200_feature_ref> obs;
    std::vector<double> T_obs;
    Lock lk(rt().mu);
    b200::PinFrames pin;
    svob200_ctx* ctx = ctx_locked();
    const int64_t cur_id = b200::ensure_frame_locked(*frame);
    const SE3 ident;
    for (int i = 0; i < n_points; ++i) {
      Point* p = pts[i];
      mp[i].pos[0] = p->pos_[0]; mp[i].pos[1] = p->pos_[1]; mp[i].pos[2] = p->pos_[2];
      mp[i].type = (int)p->type_; mp[i].obs_begin = (int)obs.size(); mp[i].reserved = 0;
      for (auto o = p->obs_.begin(); o != p->obs_.end(); ++o) {
        svob200_feature_ref f;
        fill_feature_ref(**o, b200::ensure_frame_locked(*(*o)->frame), ident, &f);
        obs.push_back(f);
        double T[7];
        b200::pose7((*o)->frame->T_f_w_, T);
        T_obs.insert(T_obs.end(), T, T + 7);
        obs_ftr.push_back(*o);
      }
      mp[i].obs_end = (int)obs.size();
    }
    const svob200_camera cam = b200::camera_of(frame->cam_);
    svob200_matcher_opts mo;
    b200::matcher_opts_of(matcher_.options_, &mo);
    double T_cur[7];
    b200::pose7(frame->T_f_w_, T_cur);
    const int off[2] = {0, n_points};
    check(svob200_reproject_map(ctx, cur_id, &cam, 1, T_cur, off, n_points, mp.data(), (int)obs.size(), obs.data(), T_obs.data(),
                                grid_.cell_size, (int)Config::maxFts(), &mo, res.data(), winner.data(), &stats, SVOB200_MEM_HOST),
          "svob200_reproject_map");
  }
  // ---- side effects, in the reference's order
  for (size_t i = 0; i < n_map; ++i)
    if (res[i].status != SVOB200_REPROJ_NOT_IN_FRAME) overlap_kfs[owner[i]].second++;                   // :116-117
  {
    size_t i = n_map;
    auto it = map_.point_candidates_.candidates_.begin();
    while (it != map_.point_candidates_.candidates_.end()) {                                            // :129-144
      const bool in_frame = res[i++].status != SVOB200_REPROJ_NOT_IN_FRAME;
      if (!in_frame) {
        it->first->n_failed_reproj_ += 3;
        if (it->first->n_failed_reproj_ > 30) {
          map_.point_candidates_.deleteCandidate(*it);
          it = map_.point_candidates_.candidates_.erase(it);
          continue;
        }
      }
      ++it;
    }
  }
  cand_lock.unlock();
  // tried candidates of every visited cell, in trial order (type descending, then insertion order: :184)
  std::vector<std::vector<int> > tried(n_cells);
  for (int i = 0; i < n_points; ++i)
    if (res[i].status == SVOB200_REPROJ_DELETED || res[i].status == SVOB200_REPROJ_FAILED || res[i].status == SVOB200_REPROJ_MATCHED)
      tried[res[i].cell].push_back(i);
  std::vector<int> type0(n_points);
  for (int i = 0; i < n_points; ++i) type0[i] = (int)pts[i]->type_;
  for (int c = 0; c < n_cells; ++c) {
    std::vector<int>& t = tried[c];
    std::stable_sort(t.begin(), t.end(), [&](int a, int b) { return type0[a] > type0[b]; });
    for (int i : t) {
      Point* pt = pts[i];
      ++n_trials_;
      if (res[i].status == SVOB200_REPROJ_DELETED) continue;                                            // :190-194
      if (res[i].status == SVOB200_REPROJ_FAILED) {                                                     // :202-211
        pt->n_failed_reproj_++;
        if (pt->type_ == Point::TYPE_UNKNOWN && pt->n_failed_reproj_ > 15) map_.safeDeletePoint(pt);
        if (pt->type_ == Point::TYPE_CANDIDATE && pt->n_failed_reproj_ > 30) map_.point_candidates_.deleteCandidatePoint(pt);
        continue;
      }
      pt->n_succeeded_reproj_++;                                                                        // :214-231
      if (pt->type_ == Point::TYPE_UNKNOWN && pt->n_succeeded_reproj_ > 10) pt->type_ = Point::TYPE_GOOD;
      Vector2d px(res[i].px[0], res[i].px[1]);
      Feature* new_feature = new Feature(frame.get(), px, res[i].search_level);
      frame->addFeature(new_feature);
      new_feature->point = pt;
      Feature* ref_ftr = obs_ftr[res[i].obs];
      if (ref_ftr->type == Feature::EDGELET) {
        Matrix2d A;
        A << res[i].A_cur_ref[0], res[i].A_cur_ref[1], res[i].A_cur_ref[2], res[i].A_cur_ref[3];
        new_feature->type = Feature::EDGELET;
        new_feature->grad = A * ref_ftr->grad;
        new_feature->grad.normalize();
      }
      ++n_matches_;
    }
  }
}

// ============================================================================ pose_optimizer.h
// Linked INSTEAD OF pose_optimizer.cpp (pose_optimizer.cpp:31-181): the whole Gauss-Newton loop, the MAD scale, the
// Tukey weights, the outlier pass and the two medians run in ONE launch (svob200_pose_optimize).
namespace pose_optimizer {

void optimizeGaussNewton(const double reproj_thresh, const size_t n_iter, const bool verbose, FramePtr& frame, double& estimated_scale,
                         double& error_init, double& error_final, size_t& num_obs)
{
  std::vector<Feature*> fs;
  for (auto it = frame->fts_.begin(); it != frame->fts_.end(); ++it) if ((*it)->point != NULL) fs.push_back(*it);
  if (fs.empty()) return;                                                                               // :56-57
  const int n = (int)fs.size();
  std::vector<double> f(3 * n), pos(3 * n);
  std::vector<int> level(n);
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) { f[3 * i + k] = fs[i]->f[k]; pos[3 * i + k] = fs[i]->point->pos_[k]; }
    level[i] = fs[i]->level;
  }
  const svob200_camera cam = b200::camera_of(frame->cam_);
  svob200_pose_opt_opts o;
  svob200_pose_opt_opts_default(&o);
  o.reproj_thresh = reproj_thresh; o.n_iter = (int)n_iter; o.eps = EPS;
  double T[7];
  b200::pose7(frame->T_f_w_, T);
  svob200_pose_opt_result r;
  std::vector<uint8_t> outlier(n);
  const int off[2] = {0, n};
  {
    Lock lk(rt().mu);
    check(svob200_pose_optimize(ctx_locked(), &cam, 1, off, f.data(), level.data(), pos.data(), &o, T, &r, outlier.data(), SVOB200_MEM_HOST),
          "svob200_pose_optimize");
  }
  frame->T_f_w_ = b200::se3_of(T);
  Matrix<double, 6, 6> A;
  for (int rr = 0; rr < 6; ++rr) for (int c = 0; c < 6; ++c) A(rr, c) = r.A[rr * 6 + c];
  const double pixel_variance = 1.0;
  frame->Cov_ = pixel_variance * (A * std::pow(frame->cam_->errorMultiplier2(), 2)).inverse();          // :139-140
  for (int i = 0; i < n; ++i) if (outlier[i]) fs[i]->point = NULL;                                      // :149-153
  estimated_scale = r.estimated_scale; error_init = r.error_init; error_final = r.error_final;
  num_obs = (size_t)r.num_obs;
  if (verbose) fprintf(stderr, "[svo_b200] pose optimizer: %d its, n obs = %d, scale = %f, error init = %f, error end = %f\n",
                       r.iters, r.num_obs, estimated_scale, error_init, error_final);
}

}  // namespace pose_optimizer

// ============================================================================ FrameHandlerBase::optimizeStructure
namespace b200 {

// frame_handler_base.cpp:190-210 with every Point::optimize (point.cpp:130-192) in one launch
void optimizeStructure(FramePtr frame, size_t max_n_pts, int max_iter)
{
  std::deque<Point*> pts;
  for (Features::iterator it = frame->fts_.begin(); it != frame->fts_.end(); ++it) if ((*it)->point != NULL) pts.push_back((*it)->point);
  max_n_pts = std::min(max_n_pts, pts.size());
  std::nth_element(pts.begin(), pts.begin() + max_n_pts, pts.end(),
                   [](Point* lhs, Point* rhs) { return lhs->last_structure_optim_ < rhs->last_structure_optim_; });
  const int n = (int)max_n_pts;
  if (n == 0) return;
  std::vector<int> off(n + 1, 0);
  std::vector<double> T, f, pos(3 * n);
  for (int i = 0; i < n; ++i) {
    for (auto o = pts[i]->obs_.begin(); o != pts[i]->obs_.end(); ++o) {
      double t[7];
      pose7((*o)->frame->T_f_w_, t);
      T.insert(T.end(), t, t + 7);
      for (int k = 0; k < 3; ++k) f.push_back((*o)->f[k]);
    }
    off[i + 1] = (int)(T.size() / 7);
    for (int k = 0; k < 3; ++k) pos[3 * i + k] = pts[i]->pos_[k];
  }
  {
    Lock lk(rt().mu);
    check(svob200_points_optimize(ctx_locked(), n, off.data(), T.data(), f.data(), max_iter, EPS, pos.data(), nullptr, SVOB200_MEM_HOST),
          "svob200_points_optimize");
  }
  for (int i = 0; i < n; ++i) {
    pts[i]->pos_ = Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    pts[i]->last_structure_optim_ = frame->id_;
  }
}

}  // namespace b200

}  // namespace svo
