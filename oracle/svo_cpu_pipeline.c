/*
 * svo_cpu_pipeline.c — CPU restatement of one front-end step per frame, built from the oracle's
 * operator restatements (svo_oracle.c).  TEST INFRASTRUCTURE ONLY: the checker for
 * svob200_tracker_step and the "port" CPU baseline of bench.py.
 *
 * A step chains the reference's hot-path operators the way FrameHandlerMono::processFrame and
 * DepthFilter::updateSeeds do (frame_handler_mono.cpp:171-262, depth_filter.cpp:237-341), without
 * the host-only stages that are out of scope (pose_optimizer, map management):
 *   Frame ctor (pyramid) -> SparseImgAlign::run(last, cur) -> cur.T_f_w = T_cur_from_ref * last.T_f_w
 *   -> for every map point: px = cur.w2c(pos); Matcher::findMatchDirect(kf patch -> cur)
 *   -> DepthFilter::updateSeeds(cur) over the keyframe's seeds.
 * Finished seeds are re-initialised when `reseed` is set (stationary benchmark workload).
 */
#include "svo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef struct {
  double T_cur_w[7];
  double chi2;
  int n_tracked, n_matched, n_seeds_updated, n_seeds_converged, n_seeds_failed, n_seeds_skipped;
  int align_iters, n_exact_chi2;
  int n_reproj_trials, n_pose_obs;     /* chain mode: Reprojector::n_trials_, features left after the pose optimiser */
} svo_step_stats;

#define SVO_SEQ_MAX_KFS 16
#define SVO_SEED_TOO_OLD 7
/* what DepthFilter::updateSeeds saw for one seed in the last step (mirror of svob200_seed_obs) */
typedef struct svo_seq_seed_obs { int status, search_level, zmssd_best, n_evals; double z, px_cur[2], epi_length; } svo_seq_seed_obs;

typedef struct svo_seq {
  svo_cam cam;
  int n_levels, w, h;
  svo_align_opts aopts;
  svo_matcher_opts mopts;
  double conv_thresh;
  svo_seed seed_init;
  int reseed;
  /* frames: level 0 + upper levels, dense */
  uint8_t *kf0, *kfu, *last0, *lastu, *cur0, *curu;
  svo_pyr kf, last, cur;
  double T_kf_w[7];
  int N, S;                            /* S = seed SLOTS (alive or not) */
  double *kf_px, *kf_f, *pt_world; int* kf_level;
  double *seed_px, *seed_f; int* seed_level; svo_seed* seeds;
  /* keyframe insertion (svo_oracle_seq_add_keyframe): DepthFilter::addKeyframe -> initializeSeeds (depth_filter.cpp:109-151),
   * the ageing rule (:258-261), removeKeyframe (:153-170).  Keyframe 0 is the one of set_keyframe (kf above). */
  uint8_t *kfs0[SVO_SEQ_MAX_KFS], *kfsu[SVO_SEQ_MAX_KFS];
  svo_pyr kfp[SVO_SEQ_MAX_KFS];
  double T_kfs[SVO_SEQ_MAX_KFS][7];
  int kf_used[SVO_SEQ_MAX_KFS], kf_batch[SVO_SEQ_MAX_KFS];
  int max_kfs, max_n_kfs, batch_counter;
  int det_cell, det_levels; double det_thr;
  int S_cap;
  int *seed_kf, *seed_batch, *seed_alive;
  double *step_px; int* step_ok; int* step_level; double T_step[7]; int have_step;
  /* map points: one observation per keyframe the point was matched in (Point::obs_, newest first: addFrameRef pushes to the front) */
  int* pobs_n; int* pobs_kf; double* pobs_px; double* pobs_f; int* pobs_level;      /* [N][SVO_SEQ_MAX_KFS] */
  /* per-step scratch */
  double *xyz; uint8_t* has_point;
  /* chain mode (svo_oracle_seq_set_chain): Reprojector::reprojectMap + pose_optimizer::optimizeGaussNewton replace the
   * refine-every-map-point loop, as in FrameHandlerMono::processFrame (frame_handler_mono.cpp:191-222) */
  int chain_cell, chain_max_fts, chain_pose_opt;
  /* test hooks: the pose the matcher / depth-filter stages of the NEXT step use instead of the aligned one (parity tests
   * inject the device's pose to separate the tolerance-matched alignment sums from the depth filter's own arithmetic), and
   * the per-seed observation of the last step */
  double T_override[7]; int have_override;
  struct svo_seq_seed_obs* obs;
} svo_seq;

static void build_frame(svo_seq* s, const uint8_t* img, uint8_t* l0, uint8_t* up, svo_pyr* p)
{
  memcpy(l0, img, (size_t)s->w * s->h);
  svo_oracle_build_pyramid(l0, s->w, s->h, s->n_levels, NULL, up);
  svo_oracle_make_pyr(p, l0, up, s->w, s->h, s->n_levels);
}

svo_seq* svo_oracle_seq_create(const svo_cam* cam, int n_levels, int align_max_level, int align_min_level, int n_iter,
                               int n_pyr_levels_cfg, double conv_thresh, float depth_mean, float depth_min, int reseed)
{
  svo_seq* s = (svo_seq*)calloc(1, sizeof(svo_seq));
  s->cam = *cam; s->n_levels = n_levels; s->w = cam->width; s->h = cam->height;
  s->aopts.max_level = align_max_level; s->aopts.min_level = align_min_level; s->aopts.n_iter = n_iter; s->aopts.eps = 0.000001;
  svo_oracle_matcher_opts_default(&s->mopts, n_pyr_levels_cfg);
  s->conv_thresh = conv_thresh; s->reseed = reseed;
  svo_oracle_seed_init(&s->seed_init, depth_mean, depth_min);
  const size_t l0 = (size_t)s->w * s->h, up = svo_oracle_pyramid_bytes(s->w, s->h, n_levels) + 16;
  s->max_kfs = 4; s->max_n_kfs = 3; s->det_cell = 30; s->det_levels = 3; s->det_thr = 20.0;
  s->kf0 = (uint8_t*)malloc(l0); s->kfu = (uint8_t*)malloc(up);
  s->last0 = (uint8_t*)malloc(l0); s->lastu = (uint8_t*)malloc(up);
  s->cur0 = (uint8_t*)malloc(l0); s->curu = (uint8_t*)malloc(up);
  return s;
}

void svo_oracle_seq_destroy(svo_seq* s)
{
  if (!s) return;
  free(s->kf0); free(s->kfu); free(s->last0); free(s->lastu); free(s->cur0); free(s->curu);
  free(s->kf_px); free(s->kf_f); free(s->pt_world); free(s->kf_level);
  free(s->seed_px); free(s->seed_f); free(s->seed_level); free(s->seeds); free(s->xyz); free(s->has_point); free(s->obs);
  free(s->seed_kf); free(s->seed_batch); free(s->seed_alive); free(s->step_px); free(s->step_ok); free(s->step_level);
  free(s->pobs_n); free(s->pobs_kf); free(s->pobs_px); free(s->pobs_f); free(s->pobs_level);
  for (int k = 1; k < SVO_SEQ_MAX_KFS; ++k) { free(s->kfs0[k]); free(s->kfsu[k]); }
  free(s);
}

void svo_oracle_seq_set_keyframe(svo_seq* s, const uint8_t* img, const double* T_kf_w, int N, const double* kf_px,
                                 const int* kf_level, const double* pt_world, int S, const double* seed_px, const int* seed_level)
{
  build_frame(s, img, s->kf0, s->kfu, &s->kf);
  memcpy(s->T_kf_w, T_kf_w, sizeof(s->T_kf_w));
  s->N = N; s->S = S;
  s->kf_px = (double*)malloc(sizeof(double) * 2 * (N + 1)); s->kf_f = (double*)malloc(sizeof(double) * 3 * (N + 1));
  s->pt_world = (double*)malloc(sizeof(double) * 3 * (N + 1)); s->kf_level = (int*)malloc(sizeof(int) * (N + 1));
  s->S_cap = S + 1;
  s->seed_px = (double*)malloc(sizeof(double) * 2 * (S + 1)); s->seed_f = (double*)malloc(sizeof(double) * 3 * (S + 1));
  s->seed_level = (int*)malloc(sizeof(int) * (S + 1)); s->seeds = (svo_seed*)malloc(sizeof(svo_seed) * (S + 1));
  s->xyz = (double*)malloc(sizeof(double) * 3 * (N + 1)); s->has_point = (uint8_t*)malloc(N + 1);
  s->obs = (svo_seq_seed_obs*)calloc((size_t)S + 1, sizeof(svo_seq_seed_obs));
  s->seed_kf = (int*)calloc((size_t)S + 1, sizeof(int)); s->seed_batch = (int*)calloc((size_t)S + 1, sizeof(int)); s->seed_alive = (int*)malloc(sizeof(int) * (S + 1));
  for (int i = 0; i < S; ++i) s->seed_alive[i] = 1;
  s->step_px = (double*)malloc(sizeof(double) * 2 * (N + 1)); s->step_ok = (int*)calloc((size_t)N + 1, sizeof(int));
  s->step_level = (int*)calloc((size_t)N + 1, sizeof(int));
  {
    const size_t M = (size_t)(N + 1) * SVO_SEQ_MAX_KFS;
    s->pobs_n = (int*)calloc((size_t)N + 1, sizeof(int)); s->pobs_kf = (int*)calloc(M, sizeof(int)); s->pobs_level = (int*)calloc(M, sizeof(int));
    s->pobs_px = (double*)calloc(2 * M, sizeof(double)); s->pobs_f = (double*)calloc(3 * M, sizeof(double));
  }
  s->kf_used[0] = 1; s->kf_batch[0] = 0; s->kfp[0] = s->kf; memcpy(s->T_kfs[0], T_kf_w, sizeof(s->T_kfs[0]));
  memcpy(s->kf_px, kf_px, sizeof(double) * 2 * N); memcpy(s->pt_world, pt_world, sizeof(double) * 3 * N);
  memcpy(s->kf_level, kf_level, sizeof(int) * N);
  memcpy(s->seed_px, seed_px, sizeof(double) * 2 * S); memcpy(s->seed_level, seed_level, sizeof(int) * S);
  for (int i = 0; i < N; ++i) { svo_oracle_cam2world(&s->cam, kf_px[2 * i], kf_px[2 * i + 1], s->kf_f + 3 * i); s->has_point[i] = 1; }
  for (int i = 0; i < N; ++i) {                                /* every map point starts with one observation, in keyframe 0 */
    const size_t o = (size_t)i * SVO_SEQ_MAX_KFS;
    s->pobs_n[i] = 1; s->pobs_kf[o] = 0; s->pobs_level[o] = kf_level[i];
    s->pobs_px[2 * o] = kf_px[2 * i]; s->pobs_px[2 * o + 1] = kf_px[2 * i + 1];
    memcpy(s->pobs_f + 3 * o, s->kf_f + 3 * i, sizeof(double) * 3);
  }
  for (int i = 0; i < S; ++i) { svo_oracle_cam2world(&s->cam, seed_px[2 * i], seed_px[2 * i + 1], s->seed_f + 3 * i); s->seeds[i] = s->seed_init; }
}

void svo_oracle_seq_set_chain(svo_seq* s, int cell_size, int max_fts, int pose_opt)
{
  s->chain_cell = cell_size; s->chain_max_fts = max_fts; s->chain_pose_opt = pose_opt;
}

void svo_oracle_seq_set_last(svo_seq* s, const uint8_t* img) { build_frame(s, img, s->last0, s->lastu, &s->last); }

static double norm3d(double x, double y, double z) { return sqrt((x * x + y * y) + z * z); }

void svo_oracle_seq_step(svo_seq* s, const uint8_t* cur_img, const double* T_last_w, const double* last_px,
                         svo_step_stats* st, double* px_refined, int* match_ok)
{
  memset(st, 0, sizeof(*st));
  build_frame(s, cur_img, s->cur0, s->curu, &s->cur);
  /* Feature ctor + sparse_img_align.cpp:132-134 */
  double Tlw_inv[7];
  svo_oracle_se3_inverse(T_last_w, Tlw_inv);
  for (int i = 0; i < s->N; ++i) {
    double f[3];
    svo_oracle_cam2world(&s->cam, last_px[2 * i], last_px[2 * i + 1], f);
    const double depth = norm3d(s->pt_world[3 * i] - Tlw_inv[0], s->pt_world[3 * i + 1] - Tlw_inv[1], s->pt_world[3 * i + 2] - Tlw_inv[2]);
    s->xyz[3 * i] = f[0] * depth; s->xyz[3 * i + 1] = f[1] * depth; s->xyz[3 * i + 2] = f[2] * depth;
  }
  double T_init[7];
  svo_oracle_se3_mul(T_last_w, Tlw_inv, T_init);          /* cur->T_f_w_ = last->T_f_w_ */
  svo_align_result ar;
  st->n_tracked = svo_oracle_sparse_align(&s->last, &s->cur, &s->cam, s->N, last_px, s->xyz, s->has_point, T_init, &s->aopts, &ar);
  st->chi2 = ar.chi2;
  for (int l = 0; l < SVO_MAX_LEVELS; ++l) st->align_iters += ar.iters[l];
  svo_oracle_se3_mul(ar.T_cur_ref, T_last_w, st->T_cur_w);
  /* st->T_cur_w keeps this sequence's own pose; Tc is what the following stages see.  Default mode: the override replaces
   * the aligned pose for the matcher and the depth filter.  Chain mode: the reprojector and the pose optimiser run on the own
   * pose and the override (the device's pose AFTER its optimiser) replaces it for the depth filter only. */
  double T_use[7];
  const double* Tc = st->T_cur_w;
  if (s->have_override) { memcpy(T_use, s->T_override, sizeof(T_use)); if (s->chain_cell <= 0) Tc = T_use; }
  if (s->chain_cell > 0) {
    /* Reprojector::reprojectMap over the keyframe's map points (one observation each, insertion order = fts_ order), every
     * point TYPE_UNKNOWN: the benchmark resets the per-point counters every frame to keep the workload stationary */
    const int N = s->N;
    svo_map_point* pts = (svo_map_point*)malloc(sizeof(svo_map_point) * (size_t)(N + 1));
    svo_point_obs* obs = (svo_point_obs*)malloc(sizeof(svo_point_obs) * (size_t)(N + 1));
    svo_reproj_result* res = (svo_reproj_result*)malloc(sizeof(svo_reproj_result) * (size_t)(N + 1));
    const int cols = (s->cam.width + s->chain_cell - 1) / s->chain_cell, rows = (s->cam.height + s->chain_cell - 1) / s->chain_cell;
    int* winner = (int*)malloc(sizeof(int) * (size_t)cols * rows);
    for (int i = 0; i < N; ++i) {
      memcpy(pts[i].pos, s->pt_world + 3 * i, sizeof(pts[i].pos));
      pts[i].type = SVO_POINT_UNKNOWN; pts[i].obs_begin = i; pts[i].obs_end = i + 1;
      obs[i].keyframe = 0;
      obs[i].ftr.px_ref[0] = s->kf_px[2 * i]; obs[i].ftr.px_ref[1] = s->kf_px[2 * i + 1];
      memcpy(obs[i].ftr.f_ref, s->kf_f + 3 * i, sizeof(obs[i].ftr.f_ref));
      obs[i].ftr.level_ref = s->kf_level[i]; obs[i].ftr.type = 0; obs[i].ftr.grad[0] = 1.0; obs[i].ftr.grad[1] = 0.0;
    }
    const svo_pyr* kfp = &s->kf;
    int nm = 0, nt = 0;
    svo_oracle_reproject_map(&kfp, &s->cur, &s->cam, st->T_cur_w, N, pts, obs, s->T_kf_w, s->chain_cell, s->chain_max_fts, &s->mopts,
                             res, winner, &nm, &nt);
    st->n_matched = nm; st->n_reproj_trials = nt;
    for (int i = 0; i < N; ++i) {
      const int ok = res[i].status == SVO_REPROJ_MATCHED;
      if (px_refined) { px_refined[2 * i] = res[i].px[0]; px_refined[2 * i + 1] = res[i].px[1]; }
      if (match_ok) match_ok[i] = ok;
    }
    st->n_pose_obs = nm;
    if (s->chain_pose_opt && nm > 0) {
      /* the new features of the frame, in cell order: f = cam2world(px) (Feature ctor), level = search level */
      double* f = (double*)malloc(sizeof(double) * 3 * (size_t)nm); double* pos = (double*)malloc(sizeof(double) * 3 * (size_t)nm);
      int* lv = (int*)malloc(sizeof(int) * (size_t)nm); uint8_t* outl = (uint8_t*)malloc((size_t)nm);
      int m = 0;
      for (int c = 0; c < cols * rows; ++c) {
        const int i = winner[c];
        if (i < 0) continue;
        svo_oracle_cam2world(&s->cam, res[i].px[0], res[i].px[1], f + 3 * m);
        memcpy(pos + 3 * m, pts[i].pos, sizeof(double) * 3); lv[m] = res[i].search_level; ++m;
      }
      svo_pose_opt_result pr;
      svo_oracle_pose_optimize(&s->cam, m, f, lv, pos, 2.0, 10, 0.0000000001, 8.6851f, st->T_cur_w, &pr, outl);
      st->n_pose_obs = pr.num_obs;
      free(f); free(pos); free(lv); free(outl);
    }
    free(pts); free(obs); free(res); free(winner);
  } else {
  /* reprojection refinement: Matcher::findMatchDirect per map point (matcher.cpp:156-202) incl. Point::getCloseViewObs over the
   * point's observations (point.cpp:101-125) */
  double cur_pos[3];
  svo_oracle_frame_pos(Tc, cur_pos);
  for (int i = 0; i < s->N; ++i) {
    const size_t o = (size_t)i * SVO_SEQ_MAX_KFS;
    double pc[3], px_in[2];
    svo_oracle_se3_transform(Tc, s->pt_world + 3 * i, pc);
    svo_oracle_world2cam(&s->cam, pc, px_in);
    svo_match_result mr;
    memset(&mr, 0, sizeof(mr));
    mr.px_cur[0] = px_in[0]; mr.px_cur[1] = px_in[1];
    int ok = 0;
    if (s->pobs_n[i] > 0) {
      double obs_pos[3 * SVO_SEQ_MAX_KFS];
      for (int k = 0; k < s->pobs_n[i]; ++k) svo_oracle_frame_pos(s->T_kfs[s->pobs_kf[o + k]], obs_pos + 3 * k);
      int best = 0;
      if (svo_oracle_close_view_obs(cur_pos, s->pt_world + 3 * i, s->pobs_n[i], obs_pos, &best)) {
        const int kfi = s->pobs_kf[o + best];
        svo_ref_feature f;
        f.px_ref[0] = s->pobs_px[2 * (o + best)]; f.px_ref[1] = s->pobs_px[2 * (o + best) + 1];
        memcpy(f.f_ref, s->pobs_f + 3 * (o + best), sizeof(f.f_ref));
        f.level_ref = s->pobs_level[o + best]; f.type = 0; f.grad[0] = 1.0; f.grad[1] = 0.0;
        double Tkw_inv[7], T_cur_kf[7];
        svo_oracle_se3_inverse(s->T_kfs[kfi], Tkw_inv);
        svo_oracle_se3_mul(Tc, Tkw_inv, T_cur_kf);
        const double depth_ref = norm3d(obs_pos[3 * best] - s->pt_world[3 * i], obs_pos[3 * best + 1] - s->pt_world[3 * i + 1], obs_pos[3 * best + 2] - s->pt_world[3 * i + 2]);
        ok = svo_oracle_find_match_direct(&s->kfp[kfi], &s->cur, &s->cam, &f, depth_ref, T_cur_kf, &s->mopts, px_in, &mr);
      }
    }
    st->n_matched += ok;
    if (px_refined) { px_refined[2 * i] = mr.px_cur[0]; px_refined[2 * i + 1] = mr.px_cur[1]; }
    if (match_ok) match_ok[i] = ok;
    s->step_px[2 * i] = mr.px_cur[0]; s->step_px[2 * i + 1] = mr.px_cur[1]; s->step_ok[i] = ok; s->step_level[i] = mr.search_level;
  }
  }
  memcpy(s->T_step, Tc, sizeof(s->T_step)); s->have_step = 1;
  /* depth filter */
  if (s->have_override) Tc = T_use;
  s->have_override = 0;
  for (int i = 0; i < s->S; ++i) {
    if (!s->seed_alive[i]) { memset(&s->obs[i], 0, sizeof(s->obs[i])); s->obs[i].zmssd_best = 2000 * 64; continue; }   /* empty slot */
    /* "check if seed is not already too old" (depth_filter.cpp:258-261) */
    if (s->batch_counter - s->seed_batch[i] > s->max_n_kfs) {
      s->seed_alive[i] = 0;
      memset(&s->obs[i], 0, sizeof(s->obs[i])); s->obs[i].status = SVO_SEED_TOO_OLD; s->obs[i].zmssd_best = 2000 * 64;
      continue;
    }
    svo_ref_feature f;
    f.px_ref[0] = s->seed_px[2 * i]; f.px_ref[1] = s->seed_px[2 * i + 1];
    memcpy(f.f_ref, s->seed_f + 3 * i, sizeof(f.f_ref));
    f.level_ref = s->seed_level[i]; f.type = 0; f.grad[0] = 1.0; f.grad[1] = 0.0;
    svo_epi_result epi;
    const int kfi = s->seed_kf[i];
    const int status = svo_oracle_update_seed_with_frame(&s->kfp[kfi], &s->cur, &s->cam, &f, s->T_kfs[kfi], Tc, &s->mopts,
                                                         s->conv_thresh, &s->seeds[i], &epi);
    {
      svo_seq_seed_obs* ob = &s->obs[i];
      memset(ob, 0, sizeof(*ob));
      ob->status = status; ob->zmssd_best = 2000 * 64;
      if (status != SVO_SEED_BEHIND && status != SVO_SEED_NOT_IN_FRAME) {
        ob->search_level = epi.search_level; ob->zmssd_best = epi.zmssd_best; ob->n_evals = epi.n_evals; ob->epi_length = epi.epi_length;
        ob->px_cur[0] = epi.px_cur[0]; ob->px_cur[1] = epi.px_cur[1];
        if (status != SVO_SEED_NO_MATCH) ob->z = epi.depth;
      }
    }
    if (status == SVO_SEED_UPDATED) st->n_seeds_updated++;
    else if (status == SVO_SEED_CONVERGED) st->n_seeds_converged++;
    else if (status == SVO_SEED_NO_MATCH) st->n_seeds_failed++;
    else st->n_seeds_skipped++;
    /* reseed 1: finished seeds start afresh (stationary workload); 2: EVERY seed starts afresh every frame (young-seed regime) */
    /* 3: the reference's list semantics: a converged (callback + erase, :314-331) or NaN (:334-338) seed leaves the list */
    if (s->reseed == 3) { if (status == SVO_SEED_CONVERGED || status == SVO_SEED_NAN_ERASED) s->seed_alive[i] = 0; }
    else if (s->reseed == 2 || (s->reseed && (status == SVO_SEED_CONVERGED || status == SVO_SEED_NAN_ERASED))) s->seeds[i] = s->seed_init;
  }
  /* the current frame becomes the last frame */
  { uint8_t* t0 = s->last0; uint8_t* tu = s->lastu; s->last0 = s->cur0; s->lastu = s->curu; s->cur0 = t0; s->curu = tu; }
  svo_oracle_make_pyr(&s->last, s->last0, s->lastu, s->w, s->h, s->n_levels);
}

void svo_oracle_seq_get_seeds(const svo_seq* s, svo_seed* out) { memcpy(out, s->seeds, sizeof(svo_seed) * s->S); }

/* ---- keyframe insertion ---- */
void svo_oracle_seq_set_pool(svo_seq* s, int max_kfs, int max_n_kfs, int reseed)
{
  s->max_kfs = max_kfs > SVO_SEQ_MAX_KFS ? SVO_SEQ_MAX_KFS : max_kfs; s->max_n_kfs = max_n_kfs; s->reseed = reseed;
}
void svo_oracle_seq_set_detector(svo_seq* s, int cell, int levels, double thr) { s->det_cell = cell; s->det_levels = levels; s->det_thr = thr; }
int svo_oracle_seq_num_slots(const svo_seq* s) { return s->S; }
void svo_oracle_seq_get_seed_refs(const svo_seq* s, double* px, int* level, int* kf, int* batch, int* state)
{
  for (int i = 0; i < s->S; ++i) {
    px[2 * i] = s->seed_px[2 * i]; px[2 * i + 1] = s->seed_px[2 * i + 1]; level[i] = s->seed_level[i];
    kf[i] = s->seed_kf[i]; batch[i] = s->seed_batch[i]; state[i] = s->seed_alive[i] ? 0 : 1;
  }
}

static void seq_grow(svo_seq* s, int need)
{
  if (need <= s->S_cap) return;
  const int cap = need + need / 2 + 64;
  s->seed_px = (double*)realloc(s->seed_px, sizeof(double) * 2 * cap); s->seed_f = (double*)realloc(s->seed_f, sizeof(double) * 3 * cap);
  s->seed_level = (int*)realloc(s->seed_level, sizeof(int) * cap); s->seeds = (svo_seed*)realloc(s->seeds, sizeof(svo_seed) * cap);
  s->obs = (svo_seq_seed_obs*)realloc(s->obs, sizeof(svo_seq_seed_obs) * cap);
  s->seed_kf = (int*)realloc(s->seed_kf, sizeof(int) * cap); s->seed_batch = (int*)realloc(s->seed_batch, sizeof(int) * cap);
  s->seed_alive = (int*)realloc(s->seed_alive, sizeof(int) * cap);
  s->S_cap = cap;
}

/* The frame of the most recent step becomes a keyframe (FrameHandlerMono::processFrame :260-312 -> DepthFilter::addKeyframe ->
 * initializeSeeds, synchronous mode): occupancy from the frame's features (the map points matched in it), detector, one seed per new
 * corner in cell order; the oldest keyframe leaves when the ring is full (removeKeyframe: its seeds are erased).  Returns the
 * number of new seeds. */
int svo_oracle_seq_add_keyframe(svo_seq* s, float depth_mean, float depth_min)
{
  if (!s->have_step) return -1;
  int k = -1;
  for (int i = 0; i < s->max_kfs; ++i) if (!s->kf_used[i]) { k = i; break; }
  if (k < 0) {
    k = 0;
    for (int i = 1; i < s->max_kfs; ++i) if (s->kf_batch[i] < s->kf_batch[k]) k = i;
    for (int i = 0; i < s->S; ++i) if (s->seed_alive[i] && s->seed_kf[i] == k) s->seed_alive[i] = 0;   /* removeKeyframe */
    for (int i = 0; i < s->N; ++i) {                               /* Map::safeDeleteFrame: the points forget the keyframe */
      const size_t o = (size_t)i * SVO_SEQ_MAX_KFS;
      int w = 0;
      for (int j = 0; j < s->pobs_n[i]; ++j) {
        if (s->pobs_kf[o + j] == k) continue;
        if (w != j) { s->pobs_kf[o + w] = s->pobs_kf[o + j]; s->pobs_level[o + w] = s->pobs_level[o + j];
                      memcpy(s->pobs_px + 2 * (o + w), s->pobs_px + 2 * (o + j), sizeof(double) * 2); memcpy(s->pobs_f + 3 * (o + w), s->pobs_f + 3 * (o + j), sizeof(double) * 3); }
        ++w;
      }
      s->pobs_n[i] = w;
    }
  }
  const size_t l0 = (size_t)s->w * s->h, up = svo_oracle_pyramid_bytes(s->w, s->h, s->n_levels) + 16;
  if (k == 0) {                                                    /* only without map features: reuse keyframe 0's buffers */
    memcpy(s->kf0, s->last0, l0); memcpy(s->kfu, s->lastu, up - 16);
    svo_oracle_make_pyr(&s->kf, s->kf0, s->kfu, s->w, s->h, s->n_levels); s->kfp[0] = s->kf;
  } else {
    if (!s->kfs0[k]) { s->kfs0[k] = (uint8_t*)malloc(l0); s->kfsu[k] = (uint8_t*)malloc(up); }
    memcpy(s->kfs0[k], s->last0, l0); memcpy(s->kfsu[k], s->lastu, up - 16);
    svo_oracle_make_pyr(&s->kfp[k], s->kfs0[k], s->kfsu[k], s->w, s->h, s->n_levels);
  }
  memcpy(s->T_kfs[k], s->T_step, sizeof(s->T_kfs[k]));
  if (k == 0) memcpy(s->T_kf_w, s->T_step, sizeof(s->T_kf_w));
  s->kf_used[k] = 1; s->kf_batch[k] = ++s->batch_counter;          /* ++Seed::batch_counter (:139) */
  /* AbstractDetector::setExistingFeatures (feature_detection.cpp:40-58) over the frame's features */
  const int cols = (int)ceil((double)s->w / s->det_cell), rows = (int)ceil((double)s->h / s->det_cell);
  uint8_t* occ = (uint8_t*)calloc((size_t)cols * rows, 1);
  for (int i = 0; i < s->N; ++i) {
    if (!s->step_ok[i]) continue;
    const long long c = (long long)(int)(s->step_px[2 * i + 1] / s->det_cell) * cols + (int)(s->step_px[2 * i] / s->det_cell);
    if (c >= 0 && c < (long long)cols * rows) occ[c] = 1;
  }
  svo_corner* cells = (svo_corner*)malloc(sizeof(svo_corner) * (size_t)cols * rows);
  svo_oracle_fast_detect(&s->kfp[k], s->det_levels, s->det_cell, s->det_thr, occ, cells);
  int n_new = 0;
  int slot = 0;
  for (int c = 0; c < cols * rows; ++c) {
    if (!((double)cells[c].score > s->det_thr)) continue;
    while (slot < s->S && s->seed_alive[slot]) ++slot;             /* next empty slot, else a new one at the end */
    if (slot >= s->S) { seq_grow(s, s->S + 1); slot = s->S; s->S++; memset(&s->obs[slot], 0, sizeof(s->obs[slot])); s->obs[slot].zmssd_best = 2000 * 64; }
    s->seed_px[2 * slot] = (double)cells[c].x; s->seed_px[2 * slot + 1] = (double)cells[c].y; s->seed_level[slot] = cells[c].level;
    svo_oracle_cam2world(&s->cam, s->seed_px[2 * slot], s->seed_px[2 * slot + 1], s->seed_f + 3 * slot);
    svo_oracle_seed_init(&s->seeds[slot], depth_mean, depth_min);
    s->seed_kf[slot] = k; s->seed_batch[slot] = s->batch_counter; s->seed_alive[slot] = 1;     /* (obs keeps describing the last step) */
    ++n_new; ++slot;
  }
  free(occ); free(cells);
  /* frame_handler_mono.cpp:277-279: the keyframe's features with a point add a frame reference to it (push_front); the features are the
   * map points matched in the frame: Feature(frame, px_refined, search_level) (reprojector.cpp:219-231) */
  for (int i = 0; i < s->N; ++i) {
    if (!s->step_ok[i] || s->pobs_n[i] >= SVO_SEQ_MAX_KFS) continue;
    const size_t o = (size_t)i * SVO_SEQ_MAX_KFS;
    for (int j = s->pobs_n[i]; j > 0; --j) {
      s->pobs_kf[o + j] = s->pobs_kf[o + j - 1]; s->pobs_level[o + j] = s->pobs_level[o + j - 1];
      memcpy(s->pobs_px + 2 * (o + j), s->pobs_px + 2 * (o + j - 1), sizeof(double) * 2); memcpy(s->pobs_f + 3 * (o + j), s->pobs_f + 3 * (o + j - 1), sizeof(double) * 3);
    }
    s->pobs_kf[o] = k; s->pobs_level[o] = s->step_level[i];
    s->pobs_px[2 * o] = s->step_px[2 * i]; s->pobs_px[2 * o + 1] = s->step_px[2 * i + 1];
    svo_oracle_cam2world(&s->cam, s->step_px[2 * i], s->step_px[2 * i + 1], s->pobs_f + 3 * o);
    s->pobs_n[i]++;
  }
  return n_new;
}
void svo_oracle_seq_get_seed_obs(const svo_seq* s, svo_seq_seed_obs* out) { memcpy(out, s->obs, sizeof(svo_seq_seed_obs) * s->S); }
/* the matcher / depth-filter stages of the next step use T_cur_w instead of the pose sparse alignment produced */
void svo_oracle_seq_set_pose_override(svo_seq* s, const double* T_cur_w) { memcpy(s->T_override, T_cur_w, sizeof(s->T_override)); s->have_override = 1; }

/* ---- batch of independent sequences over a pthread pool (one work item = one sequence step) ---- */
typedef struct {
  svo_seq** seqs; int n; const uint8_t* const* imgs; const double* T_last_w; const double* const* last_px; svo_step_stats* stats;
  int next; pthread_mutex_t mu;
} batch_job;

static void* batch_worker(void* arg)
{
  batch_job* j = (batch_job*)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    const int i = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (i >= j->n) break;
    svo_oracle_seq_step(j->seqs[i], j->imgs[i], j->T_last_w + 7 * i, j->last_px[i], &j->stats[i], NULL, NULL);
  }
  return NULL;
}

void svo_oracle_seq_step_batch(svo_seq** seqs, int n, const uint8_t* const* imgs, const double* T_last_w,
                               const double* const* last_px, svo_step_stats* stats, int n_threads)
{
  batch_job j; j.seqs = seqs; j.n = n; j.imgs = imgs; j.T_last_w = T_last_w; j.last_px = last_px; j.stats = stats; j.next = 0;
  pthread_mutex_init(&j.mu, NULL);
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, batch_worker, &j);
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  pthread_mutex_destroy(&j.mu);
}

/* ---- synthetic plane renderer (same float64 arithmetic / operation order as
 *      android_svo_b200/synth.py:render and csrc/synth.cu) — input generation for the CPU arms ---- */
void svo_oracle_synth_render(const uint8_t* tex, int size, double ppm, double plane_z, const svo_cam* cam,
                             const double* T_f_w, uint8_t* out)
{
  double Tw[7];
  svo_oracle_se3_inverse(T_f_w, Tw);
  const double x = Tw[3], y = Tw[4], z = Tw[5], w = Tw[6];
  const double R[9] = { 1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y) };
  const double* c = Tw;
  for (int v = 0; v < cam->height; ++v)
    for (int u = 0; u < cam->width; ++u) {
      const double X = ((double)u - cam->cx) / cam->fx, Y = ((double)v - cam->cy) / cam->fy;
      const double dx = R[0] * X + R[1] * Y + R[2];
      const double dy = R[3] * X + R[4] * Y + R[5];
      const double dz = R[6] * X + R[7] * Y + R[8];
      const double s = (plane_z - c[2]) / dz;
      const double tu = (c[0] + s * dx) * ppm + (double)(size / 2);
      const double tv = (c[1] + s * dy) * ppm + (double)(size / 2);
      const double uf = floor(tu), vf = floor(tv);
      const long long ui = (long long)uf, vi = (long long)vf;
      const double fu = tu - uf, fv = tv - vf;
      long long u0 = ui % size, v0 = vi % size, u1 = (ui + 1) % size, v1 = (vi + 1) % size;
      if (u0 < 0) u0 += size; if (v0 < 0) v0 += size; if (u1 < 0) u1 += size; if (v1 < 0) v1 += size;
      const double t00 = tex[v0 * size + u0], t01 = tex[v0 * size + u1], t10 = tex[v1 * size + u0], t11 = tex[v1 * size + u1];
      const double val = (1 - fu) * (1 - fv) * t00 + fu * (1 - fv) * t01 + (1 - fu) * fv * t10 + fu * fv * t11;
      out[(size_t)v * cam->width + u] = (uint8_t)floor(val + 0.5);
    }
}
