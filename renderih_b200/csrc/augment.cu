// Training-time image augmentation of the reference loader on the GPU (core/loader.py:122-181, utils/manoutils.py:214-261), one kernel:
//   cv2.warpAffine (INTER_LINEAR, BORDER_CONSTANT 0) -> brightness noise a*img + b, clip, uint8 -> horizontal flip
//   -> BGR2RGB, / 255, CHW, Normalize(mean, std)
// Bit-exact with OpenCV's 8-bit path: the inverse map is evaluated in double, converted to fixed point with 10 fractional bits, rounded to
// 1/32 pixel; the four taps are blended with the integer weights (32-fy)(32-fx), (32-fy)fx, fy(32-fx), fy*fx (sum 1024; OpenCV keeps them
// x32 in a 15-bit table, and (S*32 + 2^14) >> 15 == (S + 512) >> 10), taps outside the image contribute 0.
// HBM bound: reads 3 B/pixel (gathers, L2-resident 196 KB images), writes 12 B/pixel (+12 / +3 for the optional outputs).
#include "common.cuh"
using namespace rih;

__device__ __forceinline__ int fixed_coord(double mx, double my, double mc, int x, int y) {
  // saturate_cast<int>((my*y + mc) * 1024) + 16 + saturate_cast<int>(mx*x*1024), no fused multiply-add (OpenCV's C++ rounds each step)
  const int base = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(my, (double)y), mc), 1024.0)) + 16;
  const int delta = __double2int_rn(__dmul_rn(__dmul_rn(mx, (double)x), 1024.0));
  return (base + delta) >> 5;
}

__global__ void __launch_bounds__(256)
augment_u8_kernel(const unsigned char* __restrict__ src, const double* __restrict__ minv, const double* __restrict__ gain_offset,
                  const unsigned char* __restrict__ flip, float* __restrict__ dst, float* __restrict__ ori, unsigned char* __restrict__ out_u8,
                  int B, int H, int W, float m0, float m1, float m2, float s0, float s1, float s2) {
  pdl_sync();
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W); const long long t = i / W; const int y = (int)(t % H); const int b = (int)(t / H);
    const int xs = (flip && flip[b]) ? (W - 1 - x) : x;           // cv.flip(img, 1) happens after the warp / noise: final[y][x] = aug[y][W-1-x]
    const double* m = minv + (size_t)b * 6;
    const int X = fixed_coord(m[0], m[1], m[2], xs, y), Y = fixed_coord(m[3], m[4], m[5], xs, y);
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5)), fx = X & 31, fy = Y & 31;
    const int w[4] = {(32 - fy) * (32 - fx), (32 - fy) * fx, fy * (32 - fx), fy * fx};
    int acc[3] = {0, 0, 0};
    const unsigned char* img = src + (size_t)b * H * W * 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = sy + (k >> 1), xx = sx + (k & 1);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W && w[k]) {
        const unsigned char* p = img + ((size_t)yy * W + xx) * 3;
        acc[0] += w[k] * p[0]; acc[1] += w[k] * p[1]; acc[2] += w[k] * p[2];
      }
    }
    int v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[c] = (acc[c] + 512) >> 10;
      if (gain_offset) {                                          // imgUtils.add_noise (noise = 0) in float64, clip, truncate to uint8
        const double* g = gain_offset + (size_t)b * 4;
        double r = __dadd_rn(__dmul_rn(g[c], (double)v[c]), g[3]);
        r = fmin(fmax(r, 0.0), 255.0);
        v[c] = (int)r;
      }
    }
    const size_t plane = (size_t)H * W, o = (size_t)b * 3 * plane + (size_t)y * W + x;
    const float bl = (float)v[0] / 255.f, gr = (float)v[1] / 255.f, rd = (float)v[2] / 255.f;
    dst[o] = (rd - m0) / s0; dst[o + plane] = (gr - m1) / s1; dst[o + 2 * plane] = (bl - m2) / s2;
    if (ori) { ori[o] = bl; ori[o + plane] = gr; ori[o + 2 * plane] = rd; }
    if (out_u8) { unsigned char* q = out_u8 + (((size_t)b * H + y) * W + x) * 3; q[0] = (unsigned char)v[0]; q[1] = (unsigned char)v[1]; q[2] = (unsigned char)v[2]; }
  }
}

// src [B,H,W,3] uint8 BGR; minv [B,6] float64 = the INVERSE 2x3 maps (dst -> src), as cv::warpAffine computes them from the forward matrix;
// gain_offset [B,4] float64 {a_b, a_g, a_r, b} or NULL; flip [B] uint8 or NULL; dst [B,3,H,W] float32 normalised RGB;
// ori [B,3,H,W] float32 BGR / 255 or NULL (the loader's `ori_img`); out_u8 [B,H,W,3] augmented uint8 frames or NULL.
// reference: core/loader.py:122-181, utils/manoutils.py:214-261 (cv.warpAffine, imgUtils.add_noise, cv.flip, Normalize)
RIH_API int rih_augment_u8(const unsigned char* src, const double* minv, const double* gain_offset, const unsigned char* flip, float* dst, float* ori,
                           unsigned char* out_u8, int B, int H, int W, const float* mean3_host, const float* std3_host, cudaStream_t s) {
  RIH_REQUIRE(B >= 0 && H > 0 && W > 0 && H <= 32767 && W <= 32767 && src && minv && dst, "augment_u8: bad arguments");
  const long long total = (long long)B * H * W;
  if (total == 0) return 0;
  const int grid = (int)min(ew_ctas(s), (total + 255) / 256);
  launch_k(augment_u8_kernel, grid, 256, 0, s, src, minv, gain_offset, flip, dst, ori, out_u8, B, H, W, mean3_host[0], mean3_host[1], mean3_host[2],
                                         std3_host[0], std3_host[1], std3_host[2]);
  return check_launch("augment_u8");
}
