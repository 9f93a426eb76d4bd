// TEST INFRASTRUCTURE ONLY — flat C entry points over the REAL reference code for the callers either
// side of the hot path: Reprojector::reprojectMap, pose_optimizer::optimizeGaussNewton,
// Point::optimize, DepthFilter::initializeSeeds.  Compiled by oracle/Makefile together with the
// reference's unmodified TUs (reprojector.cpp, map.cpp, pose_optimizer.cpp, point.cpp, ...) into
// oracle/_ref/libsvo_ref*.so; with -DSVOB200_DROPIN the same file is linked over the B200 drop-in.
// No reference source is copied: this file only builds a small Map and calls the public surface.
//
// Pose layout: double[7] = {tx,ty,tz,qx,qy,qz,qw}.
#include <svo/global.h>
#include <svo/config.h>
#include <svo/pinhole_camera.h>
#include <svo/frame.h>
#include <svo/feature.h>
#include <svo/point.h>
#include <svo/map.h>
#include <svo/reprojector.h>
#include <svo/pose_optimizer.h>
#include <svo/feature_detection.h>
#include <svo/depth_filter.h>
#include <algorithm>
#include <cstring>
#include <deque>
#include <vector>
#ifdef SVOB200_DROPIN
#include "svo_b200_dropin.h"
#endif

using namespace svo;

namespace {

SE3 to_se3(const double* T) { return SE3(T[0], T[1], T[2], T[3], T[4], T[5], T[6]); }
void from_se3(const SE3& T, double* o)
{
  o[0] = T.get_translation().x; o[1] = T.get_translation().y; o[2] = T.get_translation().z;
  o[3] = T.get_rotation().x; o[4] = T.get_rotation().y; o[5] = T.get_rotation().z; o[6] = T.get_rotation().w;
}
cv::Mat mat_copy(const uint8_t* img, int w, int h)
{
  cv::Mat m(h, w, CV_8UC1);
  memcpy(m.data, img, (size_t)w * h);
  return m;
}

// DepthFilter with its seed list readable (seeds_ is protected)
#ifdef SVOB200_DROPIN
typedef svo::B200DepthFilter DepthFilterBase;
#else
typedef svo::DepthFilter DepthFilterBase;
#endif
struct SeedProbe : public DepthFilterBase {
  SeedProbe(feature_detection::DetectorPtr d, callback_t cb) : DepthFilterBase(d, cb) {}
  std::list<Seed>& seeds() { return seeds_; }
  void init(FramePtr f, double mean, double min) { new_keyframe_mean_depth_ = mean; new_keyframe_min_depth_ = min; initializeSeeds(f); }
};

}  // namespace

extern "C" {

void svo_ref_config_map(int grid_size, int max_fts)
{
  Config::gridSize() = grid_size;
  Config::maxFts() = max_fts;
}

// Reprojector::reprojectMap on a map built from flat arrays.
//   keyframes: n_kf images + poses (ids 0..n_kf-1, all added to the map in index order)
//   points:    position, type, observations obs[obs_begin..obs_end) = (keyframe, px, level, type, grad);
//              the first n_points - n_candidates points are map points (their features sit in the keyframes'
//              fts_, appended in point-index order), the rest are MapPointCandidates (one observation each,
//              feature NOT in fts_, map.cpp:226-231)
// outputs per point: n_failed_reproj_, n_succeeded_reproj_, type_ after the call; new features of the current
// frame (fts_) in order: point index, px, level, type, grad; Reprojector::n_matches_ / n_trials_; overlap_kfs.
int svo_ref_reproject_map(int n_kf, const uint8_t* const* kf_imgs, const double* T_kf_w, const uint8_t* cur_img, const double* T_cur_w,
                          const int* wh, const double* k, int n_points, int n_candidates, const double* pos, const int* type,
                          const int* obs_begin, const int* obs_end, const int* obs_kf, const double* obs_px, const int* obs_level,
                          const int* obs_type, const double* obs_grad,
                          int* pt_failed, int* pt_succeeded, int* pt_type, int cap_new, int* new_point, double* new_px, int* new_level,
                          int* new_type, double* new_grad, int* n_matches, int* n_trials, int* overlap_kf /*n_kf*/, int* overlap_cnt /*n_kf*/)
{
  vk::PinholeCamera* cam = new vk::PinholeCamera(wh[0], wh[1], k[0], k[1], k[2], k[3]);
  int n_new = 0;
  {
    svo::Map map;
    std::vector<FramePtr> kfs;
    for (int i = 0; i < n_kf; ++i) {
      FramePtr f(new Frame(cam, mat_copy(kf_imgs[i], wh[0], wh[1]), (double)i));
      f->T_f_w_ = to_se3(T_kf_w + 7 * i);
      kfs.push_back(f);
    }
    FramePtr cur(new Frame(cam, mat_copy(cur_img, wh[0], wh[1]), (double)n_kf));
    cur->T_f_w_ = to_se3(T_cur_w);
    std::vector<Point*> pts(n_points);
    const int n_map = n_points - n_candidates;
    for (int i = 0; i < n_points; ++i) {
      Point* p = new Point(Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
      pts[i] = p;
      for (int o = obs_end[i] - 1; o >= obs_begin[i]; --o) {   // Point::addFrameRef pushes to the FRONT of obs_ (point.cpp:61-65)
        Feature* ftr = new Feature(kfs[obs_kf[o]].get(), Vector2d(obs_px[2 * o], obs_px[2 * o + 1]), obs_level[o]);
        ftr->type = obs_type[o] ? Feature::EDGELET : Feature::CORNER;
        ftr->grad = Vector2d(obs_grad[2 * o], obs_grad[2 * o + 1]);
        ftr->point = p;
        p->addFrameRef(ftr);
        if (i < n_map) kfs[obs_kf[o]]->addFeature(ftr);
      }
      if (i >= n_map) map.point_candidates_.newCandidatePoint(p, 1.0);
      p->type_ = (Point::PointType)type[i];
    }
    for (int i = 0; i < n_kf; ++i) { kfs[i]->setKeyframe(); map.addKeyframe(kfs[i]); }
    {
      Reprojector reprojector(cam, map);
      std::vector<std::pair<FramePtr, size_t> > overlap;
      reprojector.reprojectMap(cur, overlap);
      *n_matches = (int)reprojector.n_matches_; *n_trials = (int)reprojector.n_trials_;
      for (int i = 0; i < n_kf; ++i) { overlap_kf[i] = -1; overlap_cnt[i] = 0; }
      for (size_t i = 0; i < overlap.size() && (int)i < n_kf; ++i) { overlap_kf[i] = overlap[i].first->id_ % 1000000; overlap_cnt[i] = (int)overlap[i].second; }
      // Frame::id_ comes from a global counter: report the keyframe INDEX instead
      for (size_t i = 0; i < overlap.size() && (int)i < n_kf; ++i)
        for (int j = 0; j < n_kf; ++j) if (overlap[i].first == kfs[j]) overlap_kf[i] = j;
    }
    for (int i = 0; i < n_points; ++i) { pt_failed[i] = pts[i]->n_failed_reproj_; pt_succeeded[i] = pts[i]->n_succeeded_reproj_; pt_type[i] = (int)pts[i]->type_; }
    for (auto it = cur->fts_.begin(); it != cur->fts_.end(); ++it) {
      if (n_new >= cap_new) break;
      int idx = -1;
      for (int i = 0; i < n_points; ++i) if (pts[i] == (*it)->point) idx = i;
      new_point[n_new] = idx; new_px[2 * n_new] = (*it)->px[0]; new_px[2 * n_new + 1] = (*it)->px[1]; new_level[n_new] = (*it)->level;
      new_type[n_new] = (*it)->type == Feature::EDGELET ? 1 : 0; new_grad[2 * n_new] = (*it)->grad[0]; new_grad[2 * n_new + 1] = (*it)->grad[1];
      ++n_new;
    }
    // teardown: features of the current frame reference points but do not own them; the map's destructor (reset())
    // releases keyframes, candidates and trash; map points were never handed to the map, free them here
    cur.reset();
    for (int i = 0; i < n_map; ++i) {
      if (pts[i]->type_ == Point::TYPE_DELETED) continue;    // moved to the map's trash (safeDeletePoint)
      for (auto f : pts[i]->obs_) f->point = NULL;
      delete pts[i];
    }
  }
  delete cam;
  return n_new;
}

// pose_optimizer::optimizeGaussNewton on a frame holding n features (px, level) with points at pos
void svo_ref_pose_optimize(const int* wh, const double* k, const uint8_t* img, int n, const double* px, const int* level, const double* pos,
                           double reproj_thresh, int n_iter, double* T_f_w /*inout*/, double* A_out /*36: Cov^-1 / em2^2*/,
                           double* scale_init_final /*3: estimated_scale, error_init, error_final*/, int* num_obs, uint8_t* outlier)
{
  vk::PinholeCamera* cam = new vk::PinholeCamera(wh[0], wh[1], k[0], k[1], k[2], k[3]);
  {
    FramePtr frame(new Frame(cam, mat_copy(img, wh[0], wh[1]), 0.0));
    frame->T_f_w_ = to_se3(T_f_w);
    std::vector<Point*> pts(n);
    std::vector<Feature*> fs(n);
    for (int i = 0; i < n; ++i) {
      pts[i] = new Point(Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
      fs[i] = new Feature(frame.get(), Vector2d(px[2 * i], px[2 * i + 1]), level[i]);
      fs[i]->point = pts[i];
      frame->addFeature(fs[i]);
    }
    double est = 0, e0 = 0, e1 = 0; size_t nobs = 0;
    pose_optimizer::optimizeGaussNewton(reproj_thresh, (size_t)n_iter, false, frame, est, e0, e1, nobs);
    from_se3(frame->T_f_w_, T_f_w);
    // Cov_ = (A * em2^2)^-1  =>  A = Cov_^-1 / em2^2
    const double em2 = cam->errorMultiplier2();
    Matrix<double, 6, 6> A = frame->Cov_.inverse() / (em2 * em2);
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) A_out[r * 6 + c] = A(r, c);
    scale_init_final[0] = est; scale_init_final[1] = e0; scale_init_final[2] = e1;
    *num_obs = (int)nobs;
    for (int i = 0; i < n; ++i) outlier[i] = fs[i]->point == NULL ? 1 : 0;
    for (int i = 0; i < n; ++i) delete pts[i];
  }
  delete cam;
}

// Point::optimize for one point observed from n_obs frames (pose, bearing)
void svo_ref_point_optimize(const int* wh, const double* k, const uint8_t* img, int n_obs, const double* T_f_w, const double* f, int n_iter, double* pos)
{
  vk::PinholeCamera* cam = new vk::PinholeCamera(wh[0], wh[1], k[0], k[1], k[2], k[3]);
  {
    std::vector<FramePtr> frames;
    Point pt(Vector3d(pos[0], pos[1], pos[2]));
    std::vector<Feature*> fs;
    for (int i = n_obs - 1; i >= 0; --i) {      // addFrameRef pushes to the front: obs_ ends up in array order
      FramePtr fr(new Frame(cam, mat_copy(img, wh[0], wh[1]), (double)i));
      fr->T_f_w_ = to_se3(T_f_w + 7 * i);
      frames.push_back(fr);
      Feature* ftr = new Feature(fr.get(), Vector2d(0, 0), 0);
      ftr->f = Vector3d(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
      ftr->point = &pt;
      pt.addFrameRef(ftr);
      fs.push_back(ftr);
    }
    pt.optimize((size_t)n_iter);
    pos[0] = pt.pos_[0]; pos[1] = pt.pos_[1]; pos[2] = pt.pos_[2];
    pt.obs_.clear();
    for (auto x : fs) delete x;
  }
  delete cam;
}

// DepthFilter::initializeSeeds: occupancy from the frame's existing features, FAST detect, one Seed per new corner
int svo_ref_initialize_seeds(const int* wh, const double* k, const uint8_t* img, int n_detect_levels, int cell, double thr,
                             int n_existing, const double* existing_px, double depth_mean, double depth_min,
                             int cap, int* xs, int* ys, int* levels, float* seeds /*5 per seed: a,b,mu,z_range,sigma2*/)
{
  vk::PinholeCamera* cam = new vk::PinholeCamera(wh[0], wh[1], k[0], k[1], k[2], k[3]);
  int n = 0;
  {
    const double saved = Config::triangMinCornerScore();
    Config::triangMinCornerScore() = thr;
    FramePtr frame(new Frame(cam, mat_copy(img, wh[0], wh[1]), 0.0));
    for (int i = 0; i < n_existing; ++i) frame->addFeature(new Feature(frame.get(), Vector2d(existing_px[2 * i], existing_px[2 * i + 1]), 0));
    feature_detection::DetectorPtr det(new feature_detection::FastDetector(wh[0], wh[1], cell, n_detect_levels));
    SeedProbe df(det, [](Point*, double) {});
    df.init(frame, depth_mean, depth_min);
    for (auto it = df.seeds().begin(); it != df.seeds().end() && n < cap; ++it, ++n) {
      xs[n] = (int)it->ftr->px[0]; ys[n] = (int)it->ftr->px[1]; levels[n] = it->ftr->level;
      seeds[5 * n] = it->a; seeds[5 * n + 1] = it->b; seeds[5 * n + 2] = it->mu; seeds[5 * n + 3] = it->z_range; seeds[5 * n + 4] = it->sigma2;
    }
    for (auto it = df.seeds().begin(); it != df.seeds().end(); ++it) delete it->ftr;
    df.seeds().clear();
    Config::triangMinCornerScore() = saved;
  }
  delete cam;
  return n;
}


// FrameHandlerBase::optimizeStructure (frame_handler_base.cpp:190-210) on a frame whose n features reference n points;
// point i is observed from obs[obs_begin[i]..obs_end[i]) = (pose, bearing).  The reference build runs the member's body
// (the selection by last_structure_optim_ + Point::optimize; FrameHandlerBase itself is control plane and not linked);
// the drop-in build calls svo::b200::optimizeStructure.
void svo_ref_optimize_structure(const int* wh, const double* k, const uint8_t* img, int n_points, const int* obs_begin, const int* obs_end,
                                const double* T_f_w, const double* f, const int* last_optim_in, int max_n_pts, int n_iter,
                                double* pos /*inout*/, int* last_optim_out)
{
  vk::PinholeCamera* cam = new vk::PinholeCamera(wh[0], wh[1], k[0], k[1], k[2], k[3]);
  {
    const int n_obs = n_points ? obs_end[n_points - 1] : 0;
    std::vector<FramePtr> frames;
    for (int o = 0; o < n_obs; ++o) {
      FramePtr fr(new Frame(cam, mat_copy(img, wh[0], wh[1]), (double)o));
      fr->T_f_w_ = to_se3(T_f_w + 7 * o);
      frames.push_back(fr);
    }
    FramePtr frame(new Frame(cam, mat_copy(img, wh[0], wh[1]), 1e6));
    std::vector<Point*> pts(n_points);
    std::vector<Feature*> obs_ftrs;
    for (int i = 0; i < n_points; ++i) {
      pts[i] = new Point(Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
      pts[i]->last_structure_optim_ = last_optim_in[i];
      for (int o = obs_end[i] - 1; o >= obs_begin[i]; --o) {          // addFrameRef pushes to the front
        Feature* ftr = new Feature(frames[o].get(), Vector2d(0, 0), 0);
        ftr->f = Vector3d(f[3 * o], f[3 * o + 1], f[3 * o + 2]);
        ftr->point = pts[i];
        pts[i]->addFrameRef(ftr);
        obs_ftrs.push_back(ftr);
      }
      Feature* cf = new Feature(frame.get(), Vector2d(10 + i, 10), 0);
      cf->point = pts[i];
      frame->addFeature(cf);
    }
#ifdef SVOB200_DROPIN
    svo::b200::optimizeStructure(frame, (size_t)max_n_pts, n_iter);
#else
    {
      std::deque<Point*> q;
      for (Features::iterator it = frame->fts_.begin(); it != frame->fts_.end(); ++it) if ((*it)->point != NULL) q.push_back((*it)->point);
      size_t m = std::min((size_t)max_n_pts, q.size());
      std::nth_element(q.begin(), q.begin() + m, q.end(), [](Point* l, Point* r) { return l->last_structure_optim_ < r->last_structure_optim_; });
      for (std::deque<Point*>::iterator it = q.begin(); it != q.begin() + m; ++it) { (*it)->optimize(n_iter); (*it)->last_structure_optim_ = frame->id_; }
    }
#endif
    for (int i = 0; i < n_points; ++i) {
      pos[3 * i] = pts[i]->pos_[0]; pos[3 * i + 1] = pts[i]->pos_[1]; pos[3 * i + 2] = pts[i]->pos_[2];
      last_optim_out[i] = pts[i]->last_structure_optim_ == frame->id_ ? 1 : 0;
    }
    for (auto x : obs_ftrs) delete x;
    for (int i = 0; i < n_points; ++i) { pts[i]->obs_.clear(); }
    frame.reset();
    for (int i = 0; i < n_points; ++i) delete pts[i];
  }
  delete cam;
}

}  // extern "C"
