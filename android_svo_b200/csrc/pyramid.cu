// pyramid.cu — integer half-sampling image pyramid (reference: vk::halfSample
// vision.cpp:20-110, frame_utils::createImgPyramid frame.cpp:186-195).
//
// B200 design: the job is pure HBM streaming (read level 0 once, write every
// coarser level once: 408,000 B per VGA 4-level pyramid), so ONE fused kernel
// builds all levels: each CTA owns a 64x64 level-0 tile, pulls it with two
// coalesced uint8x16 loads per thread, reduces it with byte-SIMD integer ops in
// registers, and walks the remaining levels through shared memory.  No level is
// ever re-read from HBM.  Both reference roundings are implemented bit-exactly:
//   SSE2  : avg_epu8 of the row pair then avg_epu16 of neighbours (round-half-up twice)
//   TRUNC : (a+b+c+d)>>2  (scalar and NEON paths)
#include "common.cuh"
#include "kernels.h"

namespace {

// 8 level-0 bytes of the top row (t0,t1) + bottom row (b0,b1) -> 4 output bytes
__device__ __forceinline__ uint32_t half4_sse2(uint32_t t0, uint32_t t1, uint32_t b0, uint32_t b1)
{
  const uint32_t a0 = __vavgu4(t0, b0);              // (x+y+1)>>1 per byte == _mm_avg_epu8
  const uint32_t a1 = __vavgu4(t1, b1);
  const uint32_t e = __byte_perm(a0, a1, 0x6420);    // even bytes
  const uint32_t o = __byte_perm(a0, a1, 0x7531);    // odd bytes
  return __vavgu4(e, o);                             // == _mm_avg_epu16 on zero-extended lanes
}
__device__ __forceinline__ uint32_t half2_trunc(uint32_t t, uint32_t b)
{
  // 16-bit lanes: bytes (0,2) and (1,3) zero-extended, summed (max 1020), >>2
  const uint32_t s = __byte_perm(t, 0, 0x4240) + __byte_perm(t, 0, 0x4341) + __byte_perm(b, 0, 0x4240) + __byte_perm(b, 0, 0x4341);
  return (s >> 2) & 0x00FF00FFu;                     // results in bytes 0 and 2
}
__device__ __forceinline__ uint32_t half4_trunc(uint32_t t0, uint32_t t1, uint32_t b0, uint32_t b1)
{
  return __byte_perm(half2_trunc(t0, b0), half2_trunc(t1, b1), 0x6420);
}
__device__ __forceinline__ uint32_t half1(int a, int b, int c, int d, int mode)
{
  // a,b top pair; c,d bottom pair
  if (mode == SVOB200_ROUND_SSE2) return (uint32_t)(((((a + c + 1) >> 1) + ((b + d + 1) >> 1) + 1) >> 1));
  return (uint32_t)((a + b + c + d) >> 2);
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p)
{
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// One CTA = one 64x64 level-0 tile of one image; 128 threads.
// Thread t: 16-pixel segment (t&3) of row pair (t>>2).
__global__ void __launch_bounds__(128) pyramid_fused_kernel(DevFrame f, int modes_mask)
{
  __shared__ __align__(16) uint8_t s_a[32 * 32];
  __shared__ __align__(16) uint8_t s_b[16 * 16];
  const int t = threadIdx.x;
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * 64, ty0 = blockIdx.y * 64;

  // ---- level 0 -> 1 (registers)
  {
    const int seg = t & 3, rp = t >> 2;
    const int x0 = tx0 + seg * 16, y0 = ty0 + rp * 2;
    const int w1 = f.w[1], h1 = f.h[1];
    const int x1 = x0 >> 1, y1 = y0 >> 1;
    uint32_t o0 = 0, o1 = 0;
    const bool act = (y1 < h1) && (x0 < f.pitch[0]) && (x1 < w1);
    if (act) {
      const uint8_t* base = f.lvl[0] + (size_t)b * f.img_stride[0] + (size_t)y0 * f.pitch[0] + x0;
      const uint4 r0 = ldg_stream(reinterpret_cast<const uint4*>(base));
      const uint4 r1 = ldg_stream(reinterpret_cast<const uint4*>(base + f.pitch[0]));
      if (modes_mask & 1) {
        o0 = half4_sse2(r0.x, r0.y, r1.x, r1.y);
        o1 = half4_sse2(r0.z, r0.w, r1.z, r1.w);
      } else {
        o0 = half4_trunc(r0.x, r0.y, r1.x, r1.y);
        o1 = half4_trunc(r0.z, r0.w, r1.z, r1.w);
      }
      uint8_t* out = f.lvl[1] + (size_t)b * f.img_stride[1] + (size_t)y1 * f.pitch[1] + x1;
      if (x1 + 8 <= f.pitch[1]) {
        *reinterpret_cast<uint2*>(out) = make_uint2(o0, o1);     // pitch padding absorbs the tail
      } else {
        for (int k = 0; k < 8 && x1 + k < w1; ++k) out[k] = (uint8_t)(((k < 4 ? o0 : o1) >> (8 * (k & 3))) & 0xff);
      }
    }
    *reinterpret_cast<uint2*>(&s_a[rp * 32 + seg * 8]) = make_uint2(o0, o1);
  }
  if (f.n_levels <= 2) return;
  __syncthreads();

  // ---- levels 2.. (shared memory ping-pong); tile edge halves each level
  uint8_t* src = s_a;
  uint8_t* dst = s_b;
  int src_edge = 32;
  for (int l = 2; l < f.n_levels; ++l) {
    const int edge = src_edge >> 1;                  // 16, 8, 4, 2, 1
    const int mode = (modes_mask >> (l - 1)) & 1;
    const int wl = f.w[l], hl = f.h[l];
    const int ox = tx0 >> l, oy = ty0 >> l;
    uint8_t* gout = f.lvl[l] + (size_t)b * f.img_stride[l];
    for (int i = t; i < edge * edge; i += 128) {
      const int ly = i / edge, lx = i - ly * edge;
      const uint8_t* p = src + (2 * ly) * src_edge + 2 * lx;
      const uint32_t v = half1(p[0], p[1], p[src_edge], p[src_edge + 1], mode);
      dst[ly * edge + lx] = (uint8_t)v;
      const int gx = ox + lx, gy = oy + ly;
      if (gx < wl && gy < hl) gout[(size_t)gy * f.pitch[l] + gx] = (uint8_t)v;
    }
    __syncthreads();
    uint8_t* tmp = src; src = dst; dst = tmp;
    src_edge = edge;
  }
}

// Generic single-level kernel: any size, both roundings, and the reference's scalar pointer
// walk for ODD input widths (vision.cpp:92-109: `top` advances 2*out_w per row, then += stride,
// so each output row starts one pixel early — reproduced through dense flat offsets).
__global__ void half_sample_generic_kernel(const uint8_t* in, int in_pitch, unsigned long long in_img_stride, int w, int h,
                                           uint8_t* out, int out_pitch, unsigned long long out_img_stride, int mode)
{
  const int ow = w / 2, oh = h / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  const uint8_t* img = in + (size_t)blockIdx.z * in_img_stride;
  int a, b, c, d;
  if ((w & 1) == 0 || mode == SVOB200_ROUND_SSE2) {
    const uint8_t* p = img + (size_t)(2 * y) * in_pitch + 2 * x;
    a = p[0]; b = p[1]; c = p[in_pitch]; d = p[in_pitch + 1];
  } else {
    // dense offsets of the scalar walk: row y starts at y*(2*w-1)
    const long long o = (long long)y * (2 * w - 1) + 2 * x;
    const long long idx[4] = {o, o + 1, o + w, o + w + 1};
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = (int)(idx[k] / w), cc = (int)(idx[k] - (long long)r * w);
      v[k] = (r < h) ? img[(size_t)r * in_pitch + cc] : 0;
    }
    // the reference loop stops when `bottom` reaches the end of the image
    if ((long long)y * (2 * w - 1) + w >= (long long)w * h) return;
    a = v[0]; b = v[1]; c = v[2]; d = v[3];
  }
  out[(size_t)blockIdx.z * out_img_stride + (size_t)y * out_pitch + x] = (uint8_t)half1(a, b, c, d, mode);
}

}  // namespace

int launch_pyramid(const DevFrame& f, const int* modes, cudaStream_t s, long long* launches)
{
  if (f.n_levels <= 1) return 0;
  bool odd = false;
  for (int l = 0; l + 1 < f.n_levels; ++l) odd |= (f.w[l] & 1) != 0;
  if (!odd && f.n_levels <= 7) {
    int mask = 0;
    for (int l = 0; l + 1 < f.n_levels; ++l) if (modes[l] == SVOB200_ROUND_SSE2) mask |= 1 << l;
    dim3 grid((f.w[0] + 63) / 64, (f.h[0] + 63) / 64, f.batch);
    pyramid_fused_kernel<<<grid, 128, 0, s>>>(f, mask);
    ++*launches;
  } else {
    for (int l = 0; l + 1 < f.n_levels; ++l) {
      dim3 blk(32, 8), grid((f.w[l + 1] + 31) / 32, (f.h[l + 1] + 7) / 8, f.batch);
      if (f.w[l + 1] == 0 || f.h[l + 1] == 0) break;
      half_sample_generic_kernel<<<grid, blk, 0, s>>>(f.lvl[l], f.pitch[l], f.img_stride[l], f.w[l], f.h[l],
                                                      f.lvl[l + 1], f.pitch[l + 1], f.img_stride[l + 1], modes[l]);
      ++*launches;
    }
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_half_sample_single(const uint8_t* in, int in_pitch, int w, int h, uint8_t* out, int out_pitch, int mode,
                              cudaStream_t s, long long* launches)
{
  if (w / 2 == 0 || h / 2 == 0) return 0;
  dim3 blk(32, 8), grid((w / 2 + 31) / 32, (h / 2 + 7) / 8, 1);
  half_sample_generic_kernel<<<grid, blk, 0, s>>>(in, in_pitch, 0, w, h, out, out_pitch, 0, mode);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
