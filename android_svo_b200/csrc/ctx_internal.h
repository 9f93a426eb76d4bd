// ctx_internal.h — context / frame-store internals shared by capi.cu and tracker.cu.
#pragma once
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <unordered_map>
#include <mutex>
#include <algorithm>
#include "common.cuh"
#include "kernels.h"

constexpr int kMaxFrames = 4096;

struct FrameRec {
  DevFrame f;
  uint8_t* base = nullptr;       // owned allocation (all levels)
  size_t bytes = 0;              // its size
  uint8_t* own_l0 = nullptr;     // owned level-0 storage (f.lvl[0] may alias caller memory after bind)
  int own_pitch0 = 0;
  int slot = -1;
};

struct GrowBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = std::max(n, (size_t)1 << 16);
    want = (want + (want >> 2) + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct svob200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  long long launches = 0;
  std::unordered_map<int64_t, FrameRec> frames;
  std::vector<int> free_slots;
  // allocations of released SMALL frames, kept for the next frame_create of the same size: a caller that mirrors one camera
  // frame per step (the C++ drop-in) would otherwise pay a cudaMalloc and a cudaFree per frame (~0.1 ms, more than the kernels)
  std::vector<std::pair<size_t, uint8_t*>> spare_frames;
  DevFrame* d_table = nullptr;
  // staging arenas (HOST mem mode)
  uint8_t* h_stage = nullptr; size_t h_cap = 0;
  GrowBuf d_stage, d_scratch, d_scratch2;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // streams on which work of this context may still be in flight after an asynchronous call returned (a tracker's
  // depth-filter stream): every synchronising entry point joins them into `stream` first (ctx_join_aux)
  std::vector<std::pair<cudaStream_t, cudaEvent_t>> aux;
  std::mutex mu;
};

inline int fail(svob200_ctx* c, int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, SVOB200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

// make ctx->stream wait for everything enqueued so far on the auxiliary streams
inline int ctx_join_aux(svob200_ctx* ctx)
{
  for (auto& a : ctx->aux) {
    CU(cudaEventRecord(a.second, a.first));
    CU(cudaStreamWaitEvent(ctx->stream, a.second, 0));
  }
  return 0;
}
inline int ctx_add_aux(svob200_ctx* ctx, cudaStream_t s)
{
  cudaEvent_t e = nullptr;
  CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  ctx->aux.push_back({s, e});
  return 0;
}
inline void ctx_remove_aux(svob200_ctx* ctx, cudaStream_t s)
{
  for (size_t i = 0; i < ctx->aux.size(); ++i)
    if (ctx->aux[i].first == s) { cudaEventDestroy(ctx->aux[i].second); ctx->aux.erase(ctx->aux.begin() + i); return; }
}

inline DevCam to_cam(const svob200_camera* c) { DevCam d; d.width = c->width; d.height = c->height; d.fx = c->fx; d.fy = c->fy; d.cx = c->cx; d.cy = c->cy; return d; }

inline FrameRec* find_frame(svob200_ctx* ctx, int64_t id)
{
  auto it = ctx->frames.find(id);
  return it == ctx->frames.end() ? nullptr : &it->second;
}

inline int ensure_host_stage(svob200_ctx* ctx, size_t n)
{
  if (n <= ctx->h_cap) return 0;
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  ctx->h_stage = nullptr; ctx->h_cap = 0;
  size_t want = std::max(n, (size_t)1 << 16);
  want = (want + (want >> 2) + 255) & ~(size_t)255;
  CU(cudaMallocHost((void**)&ctx->h_stage, want));
  ctx->h_cap = want;
  return 0;
}


extern "C" int svob200_frame_bind_only(svob200_ctx* ctx, int64_t frame_id, const uint8_t* dev_gray, int stride);
