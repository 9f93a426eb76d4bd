// fast.cu — FAST-9/16 corners (threshold 10, 3x3 non-max suppression), Shi-Tomasi
// scoring and best-corner-per-grid-cell selection.
//
// reference: FastDetector::detect feature_detection.cpp:77-122; cv::FAST(img,kps,10,true) from
// OpenCV 4.5.4 features2d (fast.cpp / fast_score.cpp — third party, restated from the published
// algorithm and pinned against python cv2 in tests/golden); vk::shiTomasiScore vision.cpp:113-154.
//
// B200 design: one pass over each detected level.  A CTA stages an 80x26 pixel tile in shared memory with aligned
// 8-byte loads, runs the 4-point quick rejection for every pixel of the tile plus a 1-px ring, compacts the survivors
// into a dense list (warp ballots), computes the 16-pixel ring test and the corner score for those only, suppresses
// non-maxima, and for each surviving keypoint evaluates the 8x8 Shi-Tomasi window straight from the staged tile.
// The reference's sequential "first strictly-greater wins" rule over (level, y, x) order becomes
// a single 64-bit atomicMax per keypoint on key = (ordered(score) << 32) | ~(level,y,x), which is
// order-independent and therefore bit-identical to the sequential loop.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int TW = 64, TH = 16, HALO = 5;
constexpr int XOFF = 8;                                  // the staged tile starts 8 columns left of the output tile: rows are 8-byte aligned
constexpr int PH = TH + 2 * HALO;                        // 26 staged rows
constexpr int PP = TW + 2 * XOFF;                        // 80 staged columns (needs TW + 2 * HALO = 74 of them)
constexpr int SW = TW + 2, SH = TH + 2;                  // score tile with 1-px ring

__device__ constexpr int c_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__device__ constexpr int c_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

__device__ __forceinline__ uint32_t ordered_bits(float f)
{
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t u)
{
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// 9 contiguous set bits in a circular 16-bit mask
__device__ __forceinline__ bool has_run9(uint32_t m16)
{
  const uint32_t m = m16 | (m16 << 16);
  uint32_t r = m & (m >> 1);
  r &= r >> 2;
  r &= r >> 4;
  r &= m >> 8;
  return (r & 0xFFFFu) != 0;
}

// FAST score of the pixel at smem position p (row stride PP); 0 if not a corner.
__device__ __forceinline__ int fast_score_at(const uint8_t* p, int threshold)
{
  const int v = p[0];
  // high-speed rejection on the 4 compass points is implied by the run test; do the full ring
  int ring[16];
  uint32_t darker = 0, brighter = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    ring[k] = p[c_ring_dy[k] * PP + c_ring_dx[k]];
    darker |= (ring[k] < v - threshold ? 1u : 0u) << k;
    brighter |= (ring[k] > v + threshold ? 1u : 0u) << k;
  }
  if (!has_run9(darker) && !has_run9(brighter)) return 0;
  // cornerScore<16> (fast_score.cpp): largest threshold for which it is still a corner
  int d[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) d[k] = v - ring[k & 15];
  int a0 = threshold;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    int a = min(d[k + 1], min(d[k + 2], d[k + 3]));
    if (a <= a0) continue;
    a = min(a, min(min(d[k + 4], d[k + 5]), min(d[k + 6], min(d[k + 7], d[k + 8]))));
    a0 = max(a0, min(a, d[k]));
    a0 = max(a0, min(a, d[k + 9]));
  }
  int b0 = -a0;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    int b = max(max(d[k + 1], d[k + 2]), max(d[k + 3], max(d[k + 4], d[k + 5])));
    if (b >= b0) continue;
    b = max(b, max(d[k + 6], max(d[k + 7], d[k + 8])));
    b0 = min(b0, max(b, d[k]));
    b0 = min(b0, max(b, d[k + 9]));
  }
  return (-b0 - 1) & 0xff;
}

// vk::shiTomasiScore vision.cpp:113-154 on the staged tile; p points at (u,v)
__device__ __forceinline__ float shi_tomasi_smem(const uint8_t* p)
{
  float dXX = 0.f, dYY = 0.f, dXY = 0.f;
  for (int y = -4; y < 4; ++y) {
    const uint8_t* r = p + y * PP;
#pragma unroll
    for (int x = -4; x < 4; ++x) {
      const float dx = (float)((int)r[x + 1] - (int)r[x - 1]);
      const float dy = (float)((int)r[x + PP] - (int)r[x - PP]);
      dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;     // exact: integers < 2^24
    }
  }
  dXX = (float)((double)dXX / 128.0);
  dYY = (float)((double)dYY / 128.0);
  dXY = (float)((double)dXY / 128.0);
  const float tr = dXX + dYY;
  const float disc = tr * tr - 4 * (dXX * dYY - dXY * dXY);
  const double root = sqrt((double)disc);                  // ::sqrt(double) in the reference TU
  return (float)(0.5 * ((double)tr - root));
}

struct FastArgs {
  DevFrame f;
  int n_levels;            // detect levels (grid.y = n_levels * batch)
  int cell, grid_cols, n_cells;
  float thr_f;
  const uint8_t* occupancy;
  unsigned long long* keys;
  // raw mode (svob200_fast_corners): single image / level, write score map instead of cells
  uint8_t* raw_scores; int raw_image, raw_level, raw_threshold, raw_nonmax;
};

// Necessary condition for a FAST-9 corner: an arc of 9 contiguous ring pixels contains two ADJACENT compass points
// (ring positions 0, 4, 8, 12), so both must be darker, or both brighter, than the centre by more than the threshold.
// Four loads and eight compares reject most pixels; only the survivors pay for the 16-pixel ring and the score.
__device__ __forceinline__ bool fast_quick_pass(const uint8_t* p, int threshold)
{
  const int v = p[0];
  const int r0 = p[3 * PP], r4 = p[3], r8 = p[-3 * PP], r12 = p[-3];
  const int lo = v - threshold, hi = v + threshold;
  const uint32_t d = (r0 < lo ? 1u : 0u) | (r4 < lo ? 2u : 0u) | (r8 < lo ? 4u : 0u) | (r12 < lo ? 8u : 0u);
  const uint32_t b = (r0 > hi ? 1u : 0u) | (r4 > hi ? 2u : 0u) | (r8 > hi ? 4u : 0u) | (r12 > hi ? 8u : 0u);
  const uint32_t dr = ((d << 1) | (d >> 3)) & 15u, br = ((b << 1) | (b >> 3)) & 15u;
  return ((d & dr) | (b & br)) != 0;
}

__global__ void __launch_bounds__(256) fast_kernel(FastArgs A)
{
  __shared__ __align__(16) uint8_t s_px[PH * PP];
  __shared__ uint8_t s_sc[SH * SW];
  __shared__ unsigned short s_list[SH * SW];
  __shared__ int s_n;
  const bool raw = A.raw_scores != nullptr;
  const int level = raw ? A.raw_level : (int)(blockIdx.y % A.n_levels);
  const int b = raw ? A.raw_image : (int)(blockIdx.y / A.n_levels);
  const int w = A.f.w[level], h = A.f.h[level], pitch = A.f.pitch[level];
  const int tiles_x = (w + TW - 1) / TW, tiles_y = (h + TH - 1) / TH;
  if ((int)blockIdx.x >= tiles_x * tiles_y) return;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int ox = tx * TW, oy = ty * TH;
  const uint8_t* img = A.f.lvl[level] + (size_t)b * A.f.img_stride[level];
  const int threshold = raw ? A.raw_threshold : 10;

  // stage the tile: 8-byte chunks (ox is a multiple of 64, the pitch a multiple of 16 => chunks are aligned); a chunk that
  // starts left of the image, or a row outside it, is zero.  Bytes at gx >= w come from the row padding / the next row:
  // no corner, score or Shi-Tomasi window that is evaluated ever reads them (gx + 5 < w for every evaluated pixel).
  if (threadIdx.x == 0) s_n = 0;
  for (int i = threadIdx.x; i < PH * (PP / 8); i += 256) {
    const int r = i / (PP / 8), c8 = i - r * (PP / 8);
    const int gx = ox - XOFF + 8 * c8, gy = oy - HALO + r;
    uint2 v = make_uint2(0u, 0u);
    if (gx >= 0 && gy >= 0 && gx < pitch && gy < h) v = __ldg(reinterpret_cast<const uint2*>(img + (size_t)gy * pitch + gx));
    *reinterpret_cast<uint2*>(&s_px[r * PP + 8 * c8]) = v;
  }
  for (int i = threadIdx.x; i < SH * SW; i += 256) s_sc[i] = 0;
  __syncthreads();

  // pass 1: the quick test for every pixel of the score tile; survivors go to a dense list
  for (int i0 = 0; i0 < SH * SW; i0 += 256) {
    const int i = i0 + threadIdx.x;
    bool pass = false;
    if (i < SH * SW) {
      const int r = i / SW, c = i - r * SW;
      const int gx = ox - 1 + c, gy = oy - 1 + r;
      if (gx >= 3 && gy >= 3 && gx < w - 3 && gy < h - 3)
        pass = fast_quick_pass(&s_px[(r + HALO - 1) * PP + (c + XOFF - 1)], threshold);
    }
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_n, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (pass) s_list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
    }
  }
  __syncthreads();
  // pass 2: full ring + cornerScore for the survivors only
  for (int j = threadIdx.x; j < s_n; j += 256) {
    const int i = s_list[j];
    const int r = i / SW, c = i - r * SW;
    s_sc[i] = (uint8_t)fast_score_at(&s_px[(r + HALO - 1) * PP + (c + XOFF - 1)], threshold);
  }
  __syncthreads();

  // pass 3: 3x3 non-max suppression; the keypoints that survive go to a dense list (the Shi-Tomasi window is ~1,000
  // instructions per keypoint: evaluated in place it would run with one or two live lanes per warp)
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int i0 = 0; i0 < TH * TW; i0 += 256) {
    const int i = i0 + threadIdx.x;
    const int r = i / TW, c = i - r * TW;
    const int gx = ox + c, gy = oy + r;
    bool keep = false;
    if (gx < w && gy < h) {
      const uint8_t* q = &s_sc[(r + 1) * SW + (c + 1)];
      const int sc = q[0];
      keep = sc > 0;
      if (keep && (!raw || A.raw_nonmax))
        keep = sc > q[-1] && sc > q[1] && sc > q[-SW - 1] && sc > q[-SW] && sc > q[-SW + 1] && sc > q[SW - 1] && sc > q[SW] && sc > q[SW + 1];
      if (raw) { A.raw_scores[(size_t)gy * w + gx] = keep ? (uint8_t)sc : 0; keep = false; }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_n, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (keep) s_list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
    }
  }
  __syncthreads();
  // pass 4: grid cell, occupancy, Shi-Tomasi score and the cell's 64-bit atomicMax, one thread per keypoint
  for (int j = threadIdx.x; j < s_n; j += 256) {
    const int i = s_list[j];
    const int r = i / TW, c = i - r * TW;
    const int gx = ox + c, gy = oy + r;
    // cell index: xy is a cv::Point2f, scale an int, cell_size_ an int (feature_detection.cpp:99-100)
    const float scale = (float)(1 << level);
    const int k = (int)(((float)gy * scale) / (float)A.cell) * A.grid_cols + (int)(((float)gx * scale) / (float)A.cell);
    if (A.occupancy && A.occupancy[(size_t)b * A.n_cells + k]) continue;
    float score = 0.0f;
    if (!(gx - 4 < 1 || gx + 4 >= w - 1 || gy - 4 < 1 || gy + 4 >= h - 1))
      score = shi_tomasi_smem(&s_px[(r + HALO) * PP + (c + XOFF)]);
    const uint32_t order = ((uint32_t)level << 28) | ((uint32_t)gy << 14) | (uint32_t)gx;
    const unsigned long long key = ((unsigned long long)ordered_bits(score) << 32) | (unsigned long long)(~order);
    atomicMax(&A.keys[(size_t)b * A.n_cells + k], key);
  }
}

__global__ void fast_init_keys_kernel(unsigned long long* keys, int n, float thr_f)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((unsigned long long)ordered_bits(thr_f) << 32) | 0xFFFFFFFFull;
}

// decode cell records; one CTA per image counts its features
__global__ void fast_finalize_kernel(const unsigned long long* keys, int n_cells, float thr_f, double thr,
                                     svob200_corner* cells, int* counts)
{
  const int b = blockIdx.x;
  int local = 0;
  for (int k = threadIdx.x; k < n_cells; k += blockDim.x) {
    const unsigned long long key = keys[(size_t)b * n_cells + k];
    svob200_corner c;
    if ((uint32_t)key == 0xFFFFFFFFu) { c.x = 0; c.y = 0; c.level = 0; c.score = thr_f; }
    else {
      const uint32_t order = ~(uint32_t)key;
      c.level = (int)(order >> 28);
      const int y = (int)((order >> 14) & 0x3FFF), x = (int)(order & 0x3FFF);
      c.x = x << c.level; c.y = y << c.level;
      c.score = from_ordered_bits((uint32_t)(key >> 32));
    }
    cells[(size_t)b * n_cells + k] = c;
    if ((double)c.score > thr) ++local;                        // feature_detection.cpp:117
  }
  local = warp_sum_i(local);
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_cnt, local);
  __syncthreads();
  if (threadIdx.x == 0 && counts) counts[b] = s_cnt;
}

// stand-alone vk::shiTomasiScore(img, u, v) (vision.cpp:113-154) for n points of one image
__global__ void shi_tomasi_points_kernel(const uint8_t* img, int pitch, int cols, int rows, int n, const int* uv, float* out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int u = uv[2 * i], v = uv[2 * i + 1];
  const int halfbox = 4;
  if (u - halfbox < 1 || u + halfbox >= cols - 1 || v - halfbox < 1 || v + halfbox >= rows - 1) { out[i] = 0.f; return; }   // :129
  float dXX = 0.f, dYY = 0.f, dXY = 0.f;
  for (int y = -4; y < 4; ++y) {
    const uint8_t* r = img + (size_t)(v + y) * pitch + u;
    for (int x = -4; x < 4; ++x) {
      const float dx = (float)((int)r[x + 1] - (int)r[x - 1]);
      const float dy = (float)((int)r[x + pitch] - (int)r[x - pitch]);
      dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;
    }
  }
  dXX = (float)((double)dXX / 128.0);
  dYY = (float)((double)dYY / 128.0);
  dXY = (float)((double)dXY / 128.0);
  const float tr = dXX + dYY;
  const float disc = tr * tr - 4 * (dXX * dYY - dXY * dXY);
  out[i] = (float)(0.5 * ((double)tr - sqrt((double)disc)));
}

}  // namespace

int launch_fast_detect(const DevFrame& f, int n_detect_levels, int cell, int grid_cols, int grid_rows, double thr,
                       const uint8_t* d_occupancy, unsigned long long* d_keys, svob200_corner* d_cells, int* d_counts,
                       cudaStream_t s, long long* launches)
{
  const int n_cells = grid_cols * grid_rows;
  const int total = n_cells * f.batch;
  const float thr_f = (float)thr;
  fast_init_keys_kernel<<<(total + 255) / 256, 256, 0, s>>>(d_keys, total, thr_f);
  FastArgs A{};
  A.f = f; A.n_levels = n_detect_levels; A.cell = cell; A.grid_cols = grid_cols; A.n_cells = n_cells; A.thr_f = thr_f;
  A.occupancy = d_occupancy; A.keys = d_keys; A.raw_scores = nullptr;
  const int tiles0 = ((f.w[0] + TW - 1) / TW) * ((f.h[0] + TH - 1) / TH);
  dim3 grid(tiles0, n_detect_levels * f.batch);
  fast_kernel<<<grid, 256, 0, s>>>(A);
  fast_finalize_kernel<<<f.batch, 256, 0, s>>>(d_keys, n_cells, thr_f, thr, d_cells, d_counts);
  *launches += 3;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_fast_raw(const DevFrame& f, int image, int level, int threshold, int nonmax, uint8_t* d_scores,
                    cudaStream_t s, long long* launches)
{
  FastArgs A{};
  A.f = f; A.n_levels = 1; A.raw_scores = d_scores; A.raw_image = image; A.raw_level = level;
  A.raw_threshold = threshold; A.raw_nonmax = nonmax;
  const int tiles = ((f.w[level] + TW - 1) / TW) * ((f.h[level] + TH - 1) / TH);
  fast_kernel<<<dim3(tiles, 1), 256, 0, s>>>(A);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_shi_tomasi_points(const uint8_t* d_img, int pitch, int cols, int rows, int n, const int* d_uv, float* d_out,
                             cudaStream_t s, long long* launches)
{
  if (n <= 0) return 0;
  shi_tomasi_points_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_img, pitch, cols, rows, n, d_uv, d_out);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
