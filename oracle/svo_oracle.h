/*
 * svo_oracle.h — CPU restatement (plain C99) of the reference's semi-direct
 * tracking front end.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or call anything in oracle/.  The product path
 * (android_svo_b200/csrc, libsvob200.so) never does.
 *
 * Parity status: PINNED.  Every function here is checked
 *   (a) against the reference's own code compiled unchanged for this host
 *       (oracle/_ref/libsvo_ref.so, recipe oracle/Makefile) when /root/reference
 *       is present, and
 *   (b) against committed golden vectors generated from that build and from
 *       python cv2 4.13 (cv::FAST), tests/golden/ + tests/golden/make_golden.py.
 *
 * Paths below are relative to /root/reference/app/src/main/cpp/svo.
 * Pose layout everywhere: double[7] = {tx,ty,tz, qx,qy,qz,qw} (SE3.h:17-19).
 */
#ifndef SVO_ORACLE_H_
#define SVO_ORACLE_H_
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SVO_MAX_LEVELS 8
enum { SVO_ROUND_TRUNC = 0, SVO_ROUND_SSE2 = 1 };

typedef struct { int width, height; double fx, fy, cx, cy; } svo_cam;

/* dense image pyramid: stride == w at every level (frame.cpp:186-195) */
typedef struct { const uint8_t* data[SVO_MAX_LEVELS]; int w[SVO_MAX_LEVELS], h[SVO_MAX_LEVELS]; int n_levels; } svo_pyr;

/* ---- a1/a2: pyramid (vision.cpp:20-110, frame.cpp:186-195) ---- */
int    svo_oracle_half_sample_mode_x86(int in_cols);                 /* dispatch rule vision.cpp:78 */
void   svo_oracle_half_sample(const uint8_t* in, int w, int h, uint8_t* out, int mode);
size_t svo_oracle_pyramid_bytes(int w, int h, int n_levels);         /* bytes of levels 1..n-1 */
/* modes: n_levels-1 entries or NULL for the x86 host dispatch rule. out = levels 1..n-1 concatenated */
void   svo_oracle_build_pyramid(const uint8_t* img0, int w, int h, int n_levels, const int* modes, uint8_t* out);
void   svo_oracle_make_pyr(svo_pyr* p, const uint8_t* img0, const uint8_t* upper, int w, int h, int n_levels);

/* ---- a3/a4: FAST-9/16 (OpenCV definition), Shi-Tomasi, grid selection ---- */
int   svo_oracle_fast(const uint8_t* img, int w, int h, int threshold, int nonmax, int cap, int* x, int* y, int* score);
float svo_oracle_shi_tomasi(const uint8_t* img, int w, int h, int u, int v);        /* vision.cpp:113-154 */
typedef struct { int x, y, level; float score; } svo_corner;
/* feature_detection.cpp:77-122. cells_out has ceil(W/cell)*ceil(H/cell) entries, initialised to
 * (0,0,0,thr); returns number of cells with score > thr. occupancy may be NULL. */
int   svo_oracle_fast_detect(const svo_pyr* pyr, int n_detect_levels, int cell, double thr,
                             const uint8_t* occupancy, svo_corner* cells_out);

/* ---- a18: SE3 / camera (SE3.h, SO3.h, pinhole_camera.cpp) ---- */
void svo_oracle_se3_mul(const double A[7], const double B[7], double out[7]);
void svo_oracle_se3_inverse(const double A[7], double out[7]);
void svo_oracle_se3_exp(const double x[6], double out[7]);
void svo_oracle_se3_transform(const double T[7], const double p[3], double out[3]);
void svo_oracle_cam2world(const svo_cam* cam, double u, double v, double f[3]);
void svo_oracle_world2cam(const svo_cam* cam, const double xyz[3], double px[2]);

/* ---- a5-a8: sparse image alignment (sparse_img_align.cpp, nlls_solver_impl.hpp:25-100) ---- */
typedef struct { int max_level, min_level, n_iter; double eps; } svo_align_opts;
typedef struct {
  double T_cur_ref[7];
  double H[36];         /* row-major, last linearisation */
  double Jres[6];
  double x[6];          /* last solve */
  double chi2;          /* NLLSSolver::chi2_ at exit */
  int    n_meas;        /* n_meas_ of the last computeResiduals */
  int    iters[SVO_MAX_LEVELS]; /* residual evaluations per level */
  int    stop;          /* sticky stop_ flag */
  int    n_ambiguous;   /* rollback decisions where |new-old| <= 1e-4*old (diagnostic only) */
} svo_align_result;
/* px: 2N level-0 pixel coords; xyz_ref: 3N (= f*depth, sparse_img_align.cpp:132-134);
 * has_point: N flags (point != NULL). Returns n_meas/16 like SparseImgAlign::run. */
int svo_oracle_sparse_align(const svo_pyr* ref, const svo_pyr* cur, const svo_cam* cam, int N,
                            const double* px, const double* xyz_ref, const uint8_t* has_point,
                            const double T_cur_ref_init[7], const svo_align_opts* opts,
                            svo_align_result* res);

/* ---- a9/a10: feature alignment float paths (feature_alignment.cpp:35-282) ---- */
int svo_oracle_align2d(const uint8_t* cur_img, int w, int h, const uint8_t* patch_with_border /*10x10*/,
                       const uint8_t* patch /*8x8*/, int n_iter, double px[2]);
int svo_oracle_align1d(const uint8_t* cur_img, int w, int h, const float dir[2], const uint8_t* patch_with_border,
                       const uint8_t* patch, int n_iter, double px[2], double* h_inv);

/* ---- a11/a14: warp + ZMSSD (matcher.cpp:36-147, patch_score.h) ---- */
void svo_oracle_warp_matrix_affine(const svo_cam* cam_ref, const svo_cam* cam_cur, const double px_ref[2],
                                   const double f_ref[3], double depth_ref, const double T_cur_ref[7],
                                   int level_ref, double A_cur_ref[4] /* row-major 2x2 */);
int  svo_oracle_best_search_level(const double A_cur_ref[4], int max_level);
/* returns 0 if the warp is NaN (patch left untouched, matcher.cpp:94-98) */
int  svo_oracle_warp_affine(const double A_cur_ref[4], const uint8_t* img_ref, int w, int h, const double px_ref[2],
                            int level_ref, int search_level, int halfpatch_size, uint8_t* patch);
void svo_oracle_patch_from_border(const uint8_t* patch_with_border, uint8_t* patch);
int  svo_oracle_zmssd(const uint8_t* ref_patch /*64*/, const uint8_t* cur, int stride);
int  svo_oracle_depth_from_triangulation(const double T_search_ref[7], const double f_ref[3],
                                         const double f_cur[3], double* depth);

/* ---- a12: Matcher::findMatchDirect after getCloseViewObs (matcher.cpp:156-202) ---- */
typedef struct {
  int align_1d, align_max_iter, max_epi_search_steps, subpix_refinement, epi_search_edgelet_filtering;
  double epi_search_edgelet_max_angle;
  int max_search_level;        /* Config::nPyrLevels()-1 */
} svo_matcher_opts;
void svo_oracle_matcher_opts_default(svo_matcher_opts* o, int n_pyr_levels);
typedef struct {
  double px_ref[2]; double f_ref[3]; int level_ref; int type; /*0 corner,1 edgelet*/ double grad[2];
} svo_ref_feature;
typedef struct {
  int success; int search_level; double A_cur_ref[4]; double h_inv; double px_cur[2];
  uint8_t patch_with_border[100]; uint8_t patch[64];
} svo_match_result;
int svo_oracle_close_view_obs(const double framepos[3], const double pos[3], int n_obs,
                              const double* obs_frame_pos /*3*n*/, int* best);   /* point.cpp:101-125 */
int svo_oracle_find_match_direct(const svo_pyr* ref, const svo_pyr* cur, const svo_cam* cam,
                                 const svo_ref_feature* ftr, double depth_ref, const double T_cur_ref[7],
                                 const svo_matcher_opts* o, const double px_cur_in[2], svo_match_result* r);

/* ---- a13: Matcher::findEpipolarMatchDirect (matcher.cpp:207-355) ---- */
typedef struct {
  int success; double depth; double px_cur[2]; double epi_length; int search_level; int reject;
  int zmssd_best; int n_evals; int n_steps; double A_cur_ref[4]; double h_inv;
  uint8_t patch_with_border[100]; uint8_t patch[64];
} svo_epi_result;
int svo_oracle_find_epipolar_match(const svo_pyr* ref, const svo_pyr* cur, const svo_cam* cam,
                                   const svo_ref_feature* ftr, const double T_cur_ref[7],
                                   double d_estimate, double d_min, double d_max,
                                   const svo_matcher_opts* o, svo_epi_result* r);

/* ---- a15-a17: depth filter (depth_filter.cpp:237-416) ---- */
typedef struct { float a, b, mu, z_range, sigma2; } svo_seed;
void   svo_oracle_seed_init(svo_seed* s, float depth_mean, float depth_min);      /* depth_filter.cpp:36-45 */
void   svo_oracle_update_seed(float x, float tau2, svo_seed* s);                  /* :368-391 */
double svo_oracle_compute_tau(const double T_ref_cur[7], const double f[3], double z, double px_error_angle);
enum { SVO_SEED_BEHIND = 1, SVO_SEED_NOT_IN_FRAME = 2, SVO_SEED_NO_MATCH = 3, SVO_SEED_UPDATED = 4,
       SVO_SEED_CONVERGED = 5, SVO_SEED_NAN_ERASED = 6 };
/* One seed against one frame: the loop body of DepthFilter::updateSeeds (:250-340), without the
 * list mutation (the status tells the caller what the reference would do to the list). */
int svo_oracle_update_seed_with_frame(const svo_pyr* ref, const svo_pyr* cur, const svo_cam* cam,
                                      const svo_ref_feature* ftr, const double T_ref_w[7], const double T_cur_w[7],
                                      const svo_matcher_opts* o, double seed_convergence_sigma2_thresh,
                                      svo_seed* s, svo_epi_result* epi_out /* may be NULL */);

/* ==================================================================== SURVEY §8f "next" rows (svo_oracle_map.c) */

/* ---- f1: Reprojector::reprojectMap (reprojector.cpp:72-259) ---- */
enum { SVO_POINT_DELETED = 0, SVO_POINT_CANDIDATE = 1, SVO_POINT_UNKNOWN = 2, SVO_POINT_GOOD = 3 };   /* point.h PointType */
enum { SVO_REPROJ_NOT_IN_FRAME = 0, SVO_REPROJ_UNTRIED = 1, SVO_REPROJ_DELETED = 2, SVO_REPROJ_FAILED = 3, SVO_REPROJ_MATCHED = 4 };
typedef struct { double pos[3]; int type; int obs_begin, obs_end; } svo_map_point;
typedef struct { int keyframe; svo_ref_feature ftr; } svo_point_obs;          /* Feature of Point::obs_, in its keyframe */
typedef struct { int status, cell, obs, search_level; double px[2]; double A_cur_ref[4]; } svo_reproj_result;
void svo_oracle_frame_pos(const double T_f_w[7], double pos[3]);               /* Frame::pos() frame.h:105 */
int  svo_oracle_reproject_map(const svo_pyr* const* ref_pyrs, const svo_pyr* cur, const svo_cam* cam, const double T_cur_w[7],
                              int n_points, const svo_map_point* points, const svo_point_obs* obs, const double* T_kf_w,
                              int cell_size, int max_fts, const svo_matcher_opts* mopts, svo_reproj_result* results,
                              int* cell_winner, int* n_matches_out, int* n_trials_out);

/* ---- f2: pose_optimizer::optimizeGaussNewton (pose_optimizer.cpp:31-181), Point::optimize (point.cpp:130-192) ---- */
typedef struct { double A[36]; double chi2, estimated_scale, error_init, error_final; int iters, num_obs, rolled_back; } svo_pose_opt_result;
void svo_oracle_pose_optimize(const svo_cam* cam, int n, const double* f, const int* level, const double* pos, double reproj_thresh,
                              int n_iter, double eps, float tukey_b, double T_f_w[7], svo_pose_opt_result* res, uint8_t* outlier);
int  svo_oracle_point_optimize(int n_obs, const double* T_f_w, const double* f, int n_iter, double eps, double pos[3]);

/* ---- f3: camera input stage (../image_process.cpp:97-186, ../svo_system.cpp:49-51) ---- */
void svo_oracle_yuv420_to_rgba(const uint8_t* y, int y_stride, const uint8_t* u, const uint8_t* v, int uv_stride, int uv_pixel_stride,
                               int w, int h, uint8_t* rgba);
void svo_oracle_rgba_to_gray(const uint8_t* rgba, int w, int h, uint8_t* gray);
void svo_oracle_yuv420_to_gray(const uint8_t* y, int y_stride, const uint8_t* u, const uint8_t* v, int uv_stride, int uv_pixel_stride,
                               int w, int h, uint8_t* gray);

/* ---- f4: DepthFilter::initializeSeeds (depth_filter.cpp:129-151) ---- */
void svo_oracle_grid_occupancy(const svo_cam* cam, int cell_size, int n, const double* px, uint8_t* occupancy);
int  svo_oracle_initialize_seeds(const svo_pyr* pyr, const svo_cam* cam, int n_detect_levels, int cell_size, double thr,
                                 int n_existing, const double* existing_px, float depth_mean, float depth_min,
                                 svo_corner* corners_out, svo_seed* seeds_out);

#ifdef __cplusplus
}
#endif
#endif
