// svo_b200_dropin.h — host-side (C++) face of the B200 front end for code written against the
// reference's own headers (/root/reference/app/src/main/cpp/svo/include/svo/*.h).
//
// svo_b200_dropin.cpp DEFINES the reference's hot-path symbols (same mangled names, same
// signatures, the reference's unchanged headers) on top of the C ABI in include/svob200.h:
//
//   vk::halfSample, vk::shiTomasiScore                               (vision.h:38-40)
//   svo::feature_alignment::align1D / align2D                        (feature_alignment.h:29-44)
//   svo::warp::getWarpMatrixAffine / getBestSearchLevel / warpAffine (matcher.h:40-61)
//   svo::Matcher::findMatchDirect / findEpipolarMatchDirect / createPatchFromPatchWithBorder (matcher.h:111-126)
//   svo::SparseImgAlign::SparseImgAlign / run / getFisherInformation (sparse_img_align.h:43-57)
//   svo::feature_detection::AbstractDetector / FastDetector          (feature_detection.h:41-101)
//
//   svo::Reprojector (ctor, dtor, reprojectMap)                     (reprojector.h:48-60)
//   svo::pose_optimizer::optimizeGaussNewton                         (pose_optimizer.h)
//
// so it is linked INSTEAD OF vision.cpp, feature_alignment.cpp, matcher.cpp, sparse_img_align.cpp,
// feature_detection.cpp, reprojector.cpp and pose_optimizer.cpp.  The depth filter keeps the reference's depth_filter.cpp (threading, queues,
// seed initialisation are control plane) and swaps only the hot loop through the protected virtual
// DepthFilter::updateSeeds (depth_filter.h:159): construct svo::B200DepthFilter where the reference
// constructs svo::DepthFilter (frame_handler_mono.cpp:43-48).  See INTEGRATION.md.
#ifndef SVO_B200_DROPIN_H_
#define SVO_B200_DROPIN_H_

#include <svo/depth_filter.h>
#include <svo/frame.h>

struct svob200_ctx;
struct svob200_align_result_tag;

namespace svo {

/// DepthFilter whose per-frame seed update runs as three CUDA launches over ALL seeds
/// (svob200_seeds_update) instead of the reference's sequential loop (depth_filter.cpp:237-341).
/// List mutations, the converged-seed callback and the detector-grid occupancy are applied on the
/// host afterwards, in list order, exactly as the reference's loop does.
class B200DepthFilter : public DepthFilter
{
public:
  B200DepthFilter(feature_detection::DetectorPtr feature_detector, callback_t seed_converged_cb)
    : DepthFilter(feature_detector, seed_converged_cb) {}
  virtual ~B200DepthFilter() {}

  /// number of seeds processed / updated / failed by the most recent updateSeeds call
  size_t last_n_seeds_ = 0, last_n_updates_ = 0, last_n_failed_matches_ = 0;

protected:
  virtual void updateSeeds(FramePtr frame);
};

namespace b200 {

/// The CUDA context behind the calling thread's role (created on first use; throws std::runtime_error when no usable CUDA
/// device exists — there is no CPU fallback).  By default ONE svob200 context serves every operator; with SVOB200_DROPIN_DF_CTX=1
/// B200DepthFilter::updateSeeds (the reference runs it on a thread of its own, depth_filter.cpp:63-103) gets a second context with
/// its own stream, staging arenas and frame mirrors, so the two threads' GPU work overlaps instead of queueing behind one lock.
/// Frame mirrors are made with one upload of level 0 and the pyramid kernel (SVOB200_DROPIN_REBUILD_PYRAMID=0: level by level).
svob200_ctx* context();

/// Device pyramids of host Frames are cached by Frame::id_ (the images are immutable after the Frame
/// constructor, frame.cpp:51-64).  A Frame that is about to be destroyed can be dropped explicitly;
/// otherwise the cache evicts least-recently-used entries beyond `capacity` (default 64).
void releaseFrame(const Frame& frame);
void setFrameCacheCapacity(size_t capacity);
/// B200DepthFilter::updateSeeds sends the seed list to the device in chunks of this many seeds (default 2,048) and polls
/// seeds_updating_halt_ between chunks (addKeyframe / removeKeyframe / reset return at the next chunk, like the reference at the next seed).
int seedChunk();
void setSeedChunk(int seeds_per_call);
void shutdown();

/// FrameHandlerBase::optimizeStructure (frame_handler_base.cpp:190-210) with every Point::optimize (point.cpp:130-192)
/// of the call in one launch.  Point::optimize itself lives in point.cpp next to code the drop-in keeps, so the member
/// symbol is not replaced: call this where the frame handler calls optimizeStructure (frame_handler_mono.cpp:233).
void optimizeStructure(FramePtr frame, size_t max_n_pts, int max_iter);

/// kernels launched so far by the drop-in's context (svob200_ctx_launch_count)
long long launchCount();

/// Gauss-Newton evaluations per pyramid level of the most recent SparseImgAlign::run on this thread
/// (the device runs the whole optimisation, so NLLSSolver's per-iteration virtual hooks are not called).
const int* lastAlignIterations();   // SVOB200_MAX_LEVELS entries
int lastAlignExactChi2();

}  // namespace b200
}  // namespace svo

#endif  // SVO_B200_DROPIN_H_
