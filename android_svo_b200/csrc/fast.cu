// fast.cu — FAST-9/16 corners (threshold 10, 3x3 non-max suppression), Shi-Tomasi
// scoring and best-corner-per-grid-cell selection.
//
// reference: FastDetector::detect feature_detection.cpp:77-122; cv::FAST(img,kps,10,true) from
// OpenCV 4.5.4 features2d (fast.cpp / fast_score.cpp — third party, restated from the published
// algorithm and pinned against python cv2 in tests/golden); vk::shiTomasiScore vision.cpp:113-154.
//
// B200 design: one pass over each detected level, a CTA per 128x32 tile staged in shared memory with 16-byte loads.
//   A. byte-SIMD candidate test, FOUR pixels per thread: the 7x12-byte neighbourhood of a 4-pixel group sits in 21 registers,
//      every ring position is one PRMT, and two pixels at a time go through VIMNMX.U16x2.  The test is OpenCV's own necessary
//      condition: an arc of 9 of 16 contains one pixel of EVERY opposite pair (k, k+8), so all 8 pairs need a darker (or all a
//      brighter) member: max over the pairs of min(pair) < v - t, or min over the pairs of max(pair) > v + t.
//      It passes 33 % of the pixels of the bench texture (24 % are corners) where two adjacent compass points passed 58 %.
//   B. thread per candidate: score = max over the 16 arcs of the minimum of |v - ring| over the arc (sliding-window minimum
//      with 3-input VIMNMX3: 40 instructions), which is both the exact FAST-9 test (score > t) and cv's cornerScore - only the
//      candidate's polarity is evaluated (a darker and a brighter arc of 9 cannot coexist on 16 pixels).
//   C. thread per corner: strict 3x3 non-max suppression on the score tile.
//   D. thread per keypoint: grid cell, occupancy, Shi-Tomasi by byte dot products (dp4a on funnel-shifted words: the
//      integer sums are exact, as the reference's float sums are), one 64-bit atomicMax per keypoint.
// The reference's sequential "first strictly-greater wins" rule over (level, y, x) order becomes
// a single 64-bit atomicMax per keypoint on key = (ordered(score) << 32) | ~(level,y,x), which is
// order-independent and therefore bit-identical to the sequential loop.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int TW = 128, TH = 32;                         // output tile
constexpr int HALO = 5;                                  // rows above / below (Shi-Tomasi reads y-5 .. y+4)
constexpr int XOFF = 16;                                 // the staged tile starts 16 columns left of the output tile (16-byte aligned chunks)
constexpr int PH = TH + 2 * HALO;                        // 42 staged rows
constexpr int PP = TW + 2 * XOFF;                        // 160 staged columns = 40 words
constexpr int PW = PP / 4;
constexpr int SG = TW / 4 + 2;                           // 34 groups of 4 pixels per score row: columns ox-4 .. ox+131
constexpr int SW = 4 * SG, SH = TH + 2;                  // score tile 136 x 34 (1-px ring for the NMS, rounded to groups)
constexpr int NT = 256;
constexpr int COL0 = XOFF - 4, ROW0 = HALO - 1;              // staged-tile position of score-tile column 0 / row 0

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

// ring position K of a 4-pixel group as two words of 16-bit lanes: E = pixels 0 and 2, O = pixels 1 and 3.
// W[r][0..2] = the aligned words left of / at / right of the group on row y - 3 + r.
template <int K>
__device__ __forceinline__ void ring_lanes(const uint32_t (&W)[7][3], uint32_t& E, uint32_t& O)
{
  constexpr int dx = c_ring_dx[K], r = 3 + c_ring_dy[K];
  constexpr int start = 4 + dx;                                   // first byte inside the 12-byte row (left word = bytes 0..3)
  uint32_t v;
  if (dx == 0) v = W[r][1];
  else if (dx > 0) v = __byte_perm(W[r][1], W[r][2], (dx) | ((dx + 1) << 4) | ((dx + 2) << 8) | ((dx + 3) << 12));
  else v = __byte_perm(W[r][0], W[r][1], (start) | ((start + 1) << 4) | ((start + 2) << 8) | ((start + 3) << 12));
  E = v & 0x00ff00ffu;
  O = __byte_perm(v, 0u, 0x4341);                                 // (v >> 8) & 0x00ff00ff
}

// one opposite pair (K, K + 8): its minimum feeds the running maximum (darker test), its maximum the running minimum (brighter
// test), two pixels per VIMNMX on 16-bit lanes.  A pixel can be a darker corner only if EVERY pair has a member < v - t, i.e.
// max over the pairs of min(pair) < v - t; and a brighter one only if min over the pairs of max(pair) > v + t.
template <int K>
__device__ __forceinline__ void pair_step(const uint32_t (&W)[7][3], uint32_t& mxE, uint32_t& mxO, uint32_t& mnE, uint32_t& mnO)
{
  uint32_t e0, o0, e1, o1;
  ring_lanes<K>(W, e0, o0);
  ring_lanes<K + 8>(W, e1, o1);
  mxE = __vmaxu2(mxE, __vminu2(e0, e1)); mxO = __vmaxu2(mxO, __vminu2(o0, o1));
  mnE = __vminu2(mnE, __vmaxu2(e0, e1)); mnO = __vminu2(mnO, __vmaxu2(o0, o1));
}

// bit 15 / 31 of the result: pixel (lane) still passes.  darker: mx + t + 1 <= v; brighter: mn >= v + t + 1 (no lane can borrow:
// the minuend carries bit 15 and the subtrahend is at most 511)
__device__ __forceinline__ uint32_t dark_flags(uint32_t mx, uint32_t v_hi, uint32_t t1) { return v_hi - (mx + t1); }
__device__ __forceinline__ uint32_t bright_flags(uint32_t mn, uint32_t v_t1) { return (mn | 0x80008000u) - v_t1; }

// FAST-9 score of polarity `bright` from the ring of the pixel at smem position p: max over the 16 arcs of 9 contiguous
// ring pixels of the arc's minimum of e_k = +-(v - ring_k).  > threshold <=> the pixel is a corner of that polarity, and then
// it is cv's cornerScore + 1 (fast_score.cpp: a0 / -b0 are exactly these max-min values, the other polarity cannot exceed t).
__device__ __forceinline__ int fast_arc_score(const uint8_t* p, bool bright)
{
  const int sgn = bright ? -1 : 1;
  const int sv = sgn * (int)p[0];
  int e[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) e[k] = sv - sgn * (int)p[c_ring_dy[k] * PP + c_ring_dx[k]];     // +-(v - ring_k): one IMAD
  int w3[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) w3[k] = __vimin3_s32(e[k], e[(k + 1) & 15], e[(k + 2) & 15]);
  int best = -256;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const int m0 = __vimin3_s32(w3[k], w3[(k + 3) & 15], w3[(k + 6) & 15]);
    const int m1 = __vimin3_s32(w3[k + 1], w3[(k + 4) & 15], w3[(k + 7) & 15]);
    best = __vimax3_s32(best, m0, m1);
  }
  return best;
}

// vk::shiTomasiScore vision.cpp:113-154 on the staged tile; p points at (u,v).  The reference's float sums of integer-valued
// terms are exact (< 2^24), so they are formed as byte dot products: with P / M / Z the 8 window columns shifted by +1 / -1 / 0,
//   dXX = sum P.P + M.M - 2 P.M     dYY = sum Zu.Zu + Zd.Zd - 2 Zu.Zd     dXY = sum (P - M).(Zu - Zd)     (u / d: row below / above)
__device__ __forceinline__ float shi_tomasi_smem(const uint8_t* p)
{
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(p - 5);
  const unsigned sh = (unsigned)(a0 & 3) * 8;
  const uint32_t* q0 = reinterpret_cast<const uint32_t*>(a0 & ~(uintptr_t)3);
  uint32_t Zl[3], Zh[3];                                   // rolling window: rows y'-1, y', y'+1
  uint32_t zz[3];
  uint32_t sxx = 0, sxy_pos = 0, sxy_neg = 0, syy = 0, pm = 0, zc = 0, Pl = 0, Ph = 0, Ml = 0, Mh = 0;
#pragma unroll
  for (int r = 0; r < 10; ++r) {                           // rows y - 5 .. y + 4
    const uint32_t* q = q0 + (r - 5) * PW;
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
    const uint32_t f0 = __funnelshift_r(w0, w1, sh), f1 = __funnelshift_r(w1, w2, sh), f2 = __funnelshift_r(w2, w3, sh);   // cols x-5 .. x+6
    const uint32_t zl = __byte_perm(f0, f1, 0x4321), zh = __byte_perm(f1, f2, 0x4321);                                    // cols x-4 .. x+3
    const int s = r % 3;
    Zl[s] = zl; Zh[s] = zh;
    zz[s] = __dp4a(zl, zl, __dp4a(zh, zh, 0u));
    if (r >= 2) {
      // centre row y' = row r - 1 (its P / M were formed in the previous round), below = row r, above = row r - 2
      const int u = s, d = (r - 2) % 3;
      sxx = __dp4a(Pl, Pl, sxx); sxx = __dp4a(Ph, Ph, sxx); sxx = __dp4a(Ml, Ml, sxx); sxx = __dp4a(Mh, Mh, sxx);
      pm = __dp4a(Pl, Ml, pm); pm = __dp4a(Ph, Mh, pm);
      syy += zz[u] + zz[d];
      zc = __dp4a(Zl[u], Zl[d], zc); zc = __dp4a(Zh[u], Zh[d], zc);
      sxy_pos = __dp4a(Pl, Zl[u], sxy_pos); sxy_pos = __dp4a(Ph, Zh[u], sxy_pos); sxy_pos = __dp4a(Ml, Zl[d], sxy_pos); sxy_pos = __dp4a(Mh, Zh[d], sxy_pos);
      sxy_neg = __dp4a(Pl, Zl[d], sxy_neg); sxy_neg = __dp4a(Ph, Zh[d], sxy_neg); sxy_neg = __dp4a(Ml, Zl[u], sxy_neg); sxy_neg = __dp4a(Mh, Zh[u], sxy_neg);
    }
    // P / M of this row, used when it is the centre row (next round): rows y - 4 .. y + 3
    Ml = f0; Mh = f1;                                                                                                      // cols x-5 .. x+2
    Pl = __byte_perm(f0, f1, 0x5432); Ph = __byte_perm(f1, f2, 0x5432);                                                    // cols x-3 .. x+4
  }
  float dXX = (float)(int)(sxx - 2u * pm);
  float dYY = (float)(int)(syy - 2u * zc);
  float dXY = (float)((int)sxy_pos - (int)sxy_neg);
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
  int n_levels;            // detect levels
  int tiles_x[SVOB_MAXL], tile_begin[SVOB_MAXL + 1];   // flat tile index -> level (grid.x = tile_begin[n_levels], grid.y = batch)
  int cell, grid_cols, n_cells;
  float thr_f;
  const uint8_t* occupancy;
  unsigned long long* keys;
  // raw mode (svob200_fast_corners): single image / level, write score map instead of cells
  uint8_t* raw_scores; int raw_image, raw_level, raw_threshold, raw_nonmax;
};

// append `mine` entries per thread to a shared list: warp prefix sum + one atomic per warp; returns this thread's first slot
__device__ __forceinline__ int list_reserve(int mine, int* counter)
{
  const int lane = threadIdx.x & 31;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 31 && total) base = atomicAdd(counter, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  return base + incl - mine;
}

__global__ void __launch_bounds__(NT, 5) fast_kernel(FastArgs A)
{
  __shared__ __align__(16) uint8_t s_px[PH * PP];
  __shared__ __align__(4) uint8_t s_sc[SH * SW];
  __shared__ unsigned short s_cand[SH * SW];              // pass A -> B: score-tile index | dark << 13 | bright << 14 ; reused C -> D
  __shared__ unsigned short s_corner[SH * SW];            // pass B -> C
  __shared__ int s_n[3];
  const bool raw = A.raw_scores != nullptr;
  int level = 0;
  if (raw) level = A.raw_level;
  else { while (level + 1 < A.n_levels && (int)blockIdx.x >= A.tile_begin[level + 1]) ++level; }
  const int b = raw ? A.raw_image : (int)blockIdx.y;
  const int w = A.f.w[level], h = A.f.h[level], pitch = A.f.pitch[level];
  const int tile = (int)blockIdx.x - (raw ? 0 : A.tile_begin[level]);
  const int tiles_x = (w + TW - 1) / TW;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int ox = tx * TW, oy = ty * TH;
  const uint8_t* img = A.f.lvl[level] + (size_t)b * A.f.img_stride[level];
  const int threshold = raw ? A.raw_threshold : 10;

  // stage the tile: 16-byte chunks (ox is a multiple of 128, the pitch a multiple of 16 => chunks are aligned); a chunk that
  // starts left of the image, or a row outside it, is zero.  Bytes at gx >= w come from the row padding / the next row:
  // no corner, score or Shi-Tomasi window that is evaluated ever reads them (gx + 5 < w for every evaluated pixel).
  if (threadIdx.x < 3) s_n[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < PH * (PP / 16); i += NT) {
    const int r = i / (PP / 16), c16 = i - r * (PP / 16);
    const int gx = ox - XOFF + 16 * c16, gy = oy - HALO + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (gx >= 0 && gy >= 0 && gx < pitch && gy < h) v = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * pitch + gx));
    *reinterpret_cast<uint4*>(&s_px[r * PP + 16 * c16]) = v;
  }
  for (int i = threadIdx.x; i < SH * SW / 4; i += NT) reinterpret_cast<uint32_t*>(s_sc)[i] = 0;
  __syncthreads();

  // pass A: candidate test, four pixels per thread.  Task i = (score row sr, group sg): pixels gx = ox - 4 + 4 sg + {0..3},
  // gy = oy - 1 + sr; in the staged tile that is word COL0 / 4 + sg of row ROW0 + sr.
  const uint32_t* s_w = reinterpret_cast<const uint32_t*>(s_px);
  const uint32_t t1 = (uint32_t)(threshold + 1) * 0x00010001u;
  for (int i0 = 0; i0 < SH * SG; i0 += NT) {
    const int i = i0 + threadIdx.x;
    // flags of the group's pixels 0..3 sit at bits 15, 14, 31, 30 (E lanes: pixels 0 and 2; O lanes, shifted right by one: 1 and 3)
    uint32_t D = 0, B = 0;
    int sr = 0, sg = 0;
    if (i < SH * SG) {
      sr = i / SG; sg = i - sr * SG;
      const int gy = oy - 1 + sr, gx0 = ox - 4 + 4 * sg;
      if (gy >= 3 && gy < h - 3 && gx0 + 3 >= 3 && gx0 < w - 3) {
        uint32_t W[7][3];
        const uint32_t* q = s_w + (sr + ROW0 - 3) * PW + (COL0 / 4 - 1) + sg;      // row gy - 3, word left of the group
#pragma unroll
        for (int r = 0; r < 7; ++r) { W[r][0] = q[r * PW]; W[r][1] = q[r * PW + 1]; W[r][2] = q[r * PW + 2]; }
        const uint32_t c = W[3][1];
        const uint32_t vE = c & 0x00ff00ffu, vO = __byte_perm(c, 0u, 0x4341);
        const uint32_t vhE = vE | 0x80008000u, vhO = vO | 0x80008000u, vtE = vE + t1, vtO = vO + t1;
        uint32_t mxE = 0u, mxO = 0u, mnE = 0x00ff00ffu, mnO = 0x00ff00ffu;
        pair_step<0>(W, mxE, mxO, mnE, mnO);
        pair_step<4>(W, mxE, mxO, mnE, mnO);
        pair_step<2>(W, mxE, mxO, mnE, mnO);
        pair_step<6>(W, mxE, mxO, mnE, mnO);
        const uint32_t alive = (dark_flags(mxE, vhE, t1) | dark_flags(mxO, vhO, t1) | bright_flags(mnE, vtE) | bright_flags(mnO, vtO)) & 0x80008000u;
        if (alive) {
          pair_step<1>(W, mxE, mxO, mnE, mnO);
          pair_step<3>(W, mxE, mxO, mnE, mnO);
          pair_step<5>(W, mxE, mxO, mnE, mnO);
          pair_step<7>(W, mxE, mxO, mnE, mnO);
          D = (dark_flags(mxE, vhE, t1) & 0x80008000u) | ((dark_flags(mxO, vhO, t1) & 0x80008000u) >> 1);
          B = (bright_flags(mnE, vtE) & 0x80008000u) | ((bright_flags(mnO, vtO) & 0x80008000u) >> 1);
          if (gx0 < 3 || gx0 + 3 >= w - 3) {               // groups that straddle the 3-pixel border
            uint32_t ok = 0;
            if (gx0 + 0 >= 3 && gx0 + 0 < w - 3) ok |= 1u << 15;
            if (gx0 + 1 >= 3 && gx0 + 1 < w - 3) ok |= 1u << 14;
            if (gx0 + 2 >= 3 && gx0 + 2 < w - 3) ok |= 1u << 31;
            if (gx0 + 3 >= 3 && gx0 + 3 < w - 3) ok |= 1u << 30;
            D &= ok; B &= ok;
          }
        }
      }
    }
    const uint32_t any = D | B;
    const int mine = __popc(any);
    if (__any_sync(0xffffffffu, mine != 0)) {
      int slot = list_reserve(mine, &s_n[0]);
      const int base = sr * SW + 4 * sg;
      if (any & (1u << 15)) s_cand[slot++] = (unsigned short)((base + 0) | (((D >> 15) & 1u) << 13) | (((B >> 15) & 1u) << 14));
      if (any & (1u << 14)) s_cand[slot++] = (unsigned short)((base + 1) | (((D >> 14) & 1u) << 13) | (((B >> 14) & 1u) << 14));
      if (any & (1u << 31)) s_cand[slot++] = (unsigned short)((base + 2) | (((D >> 31) & 1u) << 13) | (((B >> 31) & 1u) << 14));
      if (any & (1u << 30)) s_cand[slot++] = (unsigned short)((base + 3) | (((D >> 30) & 1u) << 13) | (((B >> 30) & 1u) << 14));
    }
  }
  __syncthreads();
  // pass B: exact FAST-9 test + corner score for the candidates, thread per candidate
  const int n_cand = s_n[0];
  for (int j0 = 0; j0 < n_cand; j0 += NT) {
    const int j = j0 + threadIdx.x;
    bool corner = false;
    int idx = 0;
    if (j < n_cand) {
      const int e = s_cand[j];
      idx = e & 0x1fff;
      const int sr = idx / SW, c = idx - sr * SW;
      const uint8_t* p = &s_px[(sr + ROW0) * PP + (c + COL0)];
      // one polarity per candidate (both flags together are possible in principle, a darker AND a brighter arc are not)
      int best = fast_arc_score(p, !(e & (1 << 13)));
      if ((e & (3 << 13)) == (3 << 13)) best = max(best, fast_arc_score(p, true));
      if (best > threshold) { s_sc[idx] = (uint8_t)((best - 1) & 0xff); corner = true; }
    }
    const unsigned m = __ballot_sync(0xffffffffu, corner);
    if (m) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_n[1], __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (corner) s_corner[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)idx;
    }
  }
  if (raw) {
    // score map of the tile: zero, then the corners below
    for (int i = threadIdx.x; i < TH * TW; i += NT) {
      const int r = i / TW, c = i - r * TW;
      if (ox + c < w && oy + r < h) A.raw_scores[(size_t)(oy + r) * w + ox + c] = 0;
    }
  }
  __syncthreads();
  // pass C: strict 3x3 non-max suppression, thread per corner of the output tile; keypoints -> s_cand
  const int n_corner = s_n[1];
  for (int j0 = 0; j0 < n_corner; j0 += NT) {
    const int j = j0 + threadIdx.x;
    bool keep = false;
    int idx = 0;
    if (j < n_corner) {
      idx = s_corner[j];
      const int sr = idx / SW, c = idx - sr * SW;
      const int gx = ox - 4 + c, gy = oy - 1 + sr;
      if (sr >= 1 && sr <= TH && c >= 4 && c < 4 + TW && gx < w && gy < h) {
        const uint8_t* q = &s_sc[idx];
        const int sc = q[0];
        const int m = max(__vimax3_s32(__vimax3_s32(q[-SW - 1], q[-SW], q[-SW + 1]), q[-1], q[1]), __vimax3_s32(q[SW - 1], q[SW], q[SW + 1]));
        keep = sc > m || (raw && !A.raw_nonmax);
        if (raw) { if (keep) A.raw_scores[(size_t)gy * w + gx] = (uint8_t)sc; keep = false; }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (m) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(&s_n[2], __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (keep) s_cand[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)idx;
    }
  }
  __syncthreads();
  // pass D: grid cell, occupancy, Shi-Tomasi score and the cell's 64-bit atomicMax, one thread per keypoint
  const int n_kp = s_n[2];
  for (int j = threadIdx.x; j < n_kp; j += NT) {
    const int idx = s_cand[j];
    const int sr = idx / SW, c = idx - sr * SW;
    const int gx = ox - 4 + c, gy = oy - 1 + sr;
    // cell index: xy is a cv::Point2f, scale an int, cell_size_ an int (feature_detection.cpp:99-100)
    const float scale = (float)(1 << level);
    const int k = (int)(((float)gy * scale) / (float)A.cell) * A.grid_cols + (int)(((float)gx * scale) / (float)A.cell);
    if (A.occupancy && A.occupancy[(size_t)b * A.n_cells + k]) continue;
    float score = 0.0f;
    if (!(gx - 4 < 1 || gx + 4 >= w - 1 || gy - 4 < 1 || gy + 4 >= h - 1))
      score = shi_tomasi_smem(&s_px[(sr + ROW0) * PP + (c + COL0)]);
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
  A.tile_begin[0] = 0;
  for (int l = 0; l < n_detect_levels; ++l) {
    A.tiles_x[l] = (f.w[l] + TW - 1) / TW;
    A.tile_begin[l + 1] = A.tile_begin[l] + A.tiles_x[l] * ((f.h[l] + TH - 1) / TH);
  }
  dim3 grid(A.tile_begin[n_detect_levels], f.batch);
  fast_kernel<<<grid, NT, 0, s>>>(A);
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
  fast_kernel<<<dim3(tiles, 1), NT, 0, s>>>(A);
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
