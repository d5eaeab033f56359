// Byte-image kernels next to the segmentation hot path: the
// cv2 calls the reference applies to the page before / to the label maps after each model
//   cv2.resize(INTER_NEAREST)          main.py:112-113 (get_image_and_scales :214, no-patch path :371, :378)
//   otsu_copy                          main.py:178-194 (cv2.threshold(THRESH_BINARY + THRESH_OTSU) of channel 0)
//   cv2.erode / cv2.dilate, 5x5 ones   main.py:397, 2074-2075 (iterations n == one (4n+1)^2 rectangle)
// All are HBM-bound uint8 work: coalesced row-major access, no tensor cores.  Results are
// bit-identical to OpenCV's (tests/test_prepost_*.py compare against cv2 itself).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sbb {

// dst[y][x][:] = src[ys[y]][xs[x]][:]; the index tables hold cv2's resizeNN source indices.
__global__ void resize_nearest_u8_kernel(const uint8_t* __restrict__ src, int64_t src_stride, int C,
                                         uint8_t* __restrict__ dst, int64_t dst_stride, int oh, int ow,
                                         const int32_t* __restrict__ ys, const int32_t* __restrict__ xs) {
  const int64_t total = (int64_t)oh * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / ow), x = (int)(i - (int64_t)y * ow);
    const uint8_t* s = src + (int64_t)__ldg(ys + y) * src_stride + (int64_t)__ldg(xs + x) * C;
    uint8_t* d = dst + (int64_t)y * dst_stride + (int64_t)x * C;
    if (C == 3) { d[0] = __ldg(s); d[1] = __ldg(s + 1); d[2] = __ldg(s + 2); }
    else for (int c = 0; c < C; ++c) d[c] = __ldg(s + c);
  }
}

// 256-bin histogram of channel 0 of an HWC uint8 image (per-block smem histograms, one global atomic pass).
__global__ void hist_ch0_kernel(const uint8_t* __restrict__ src, int64_t stride, int H, int W, int C,
                                unsigned int* __restrict__ hist) {
  __shared__ unsigned int h[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int64_t total = (int64_t)H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    atomicAdd(&h[__ldg(src + (int64_t)y * stride + (int64_t)x * C)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], h[i]);
}

// OpenCV's getThreshVal_Otsu_8u, statement for statement, in IEEE double without FMA contraction.
__global__ void otsu_threshold_kernel(const unsigned int* __restrict__ hist, int64_t npix, int* __restrict__ thr) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double scale = __ddiv_rn(1.0, (double)npix);
  double mu = 0.0;
  for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)hist[i]));
  mu = __dmul_rn(mu, scale);
  double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
  int max_val = 0;
  const double eps = 1.1920928955078125e-07;  // FLT_EPSILON
  for (int i = 0; i < 256; ++i) {
    const double p_i = __dmul_rn((double)hist[i], scale);
    mu1 = __dmul_rn(mu1, q1);
    q1 = __dadd_rn(q1, p_i);
    const double q2 = __dsub_rn(1.0, q1);
    if (fmin(q1, q2) < eps || fmax(q1, q2) > __dsub_rn(1.0, eps)) continue;
    mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
    const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
    const double d = __dsub_rn(mu1, mu2);
    const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
    if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
  }
  *thr = max_val;
}

// otsu_copy: dst[y][x][0..2] = src[y][x][0] > thr ? 255 : 0   (the reference writes channel 0's result to all three)
__global__ void otsu_apply_kernel(const uint8_t* __restrict__ src, int64_t src_stride, int C, uint8_t* __restrict__ dst,
                                  int64_t dst_stride, int H, int W, const int* __restrict__ thr) {
  const int t = *thr;
  const int64_t total = (int64_t)H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    const uint8_t v = __ldg(src + (int64_t)y * src_stride + (int64_t)x * C) > t ? 255 : 0;
    uint8_t* d = dst + (int64_t)y * dst_stride + (int64_t)x * 3;
    d[0] = v; d[1] = v; d[2] = v;
  }
}

// One separable pass of a rectangular min (erode) / max (dilate) of radius r over in-bounds pixels
// (cv2's default border value is +-DBL_MAX, i.e. the border never wins).  dir 0: along x, 1: along y.
template <bool DILATE>
__global__ void morph_pass_kernel(const uint8_t* __restrict__ src, int64_t src_stride, uint8_t* __restrict__ dst,
                                  int64_t dst_stride, int H, int W, int C, int r, int dir) {
  const int64_t row_elems = (int64_t)W * C;
  const int64_t total = (int64_t)H * row_elems;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / row_elems);
    const int e = (int)(i - (int64_t)y * row_elems);  // x * C + c
    int v = DILATE ? 0 : 255;
    if (dir == 0) {
      const int x = e / C, c = e - x * C;
      const int a = max(x - r, 0), b = min(x + r, W - 1);
      const uint8_t* s = src + (int64_t)y * src_stride + c;
      for (int k = a; k <= b; ++k) { const int u = __ldg(s + (int64_t)k * C); v = DILATE ? max(v, u) : min(v, u); }
    } else {
      const int a = max(y - r, 0), b = min(y + r, H - 1);
      const uint8_t* s = src + e;
      for (int k = a; k <= b; ++k) { const int u = __ldg(s + (int64_t)k * src_stride); v = DILATE ? max(v, u) : min(v, u); }
    }
    dst[(int64_t)y * dst_stride + e] = (uint8_t)v;
  }
}

// ---------------------------------------------------------------------------------------------------
// Deskew search (return_deskew_slope main.py:1601-1718): for each
// candidate angle the reference rotates the zero-padded float64 textline mask with
// cv2.warpAffine(INTER_CUBIC, BORDER_REPLICATE) (rotate_image, main.py:159-163), sets every non-zero
// output pixel to 1 and sums along x.  This kernel produces those row profiles for ALL angles in one
// launch without materialising the padded image or any rotated copy.
//
// Bit-exact restatement of OpenCV's arithmetic for this case:
//  * coordinates: fixed point, AB_BITS = 10, INTER_BITS = 5:  X = (rint((M1*y + M2)*1024) + 16 +
//    rint(M0*x*1024)) >> 5,  sx = (X >> 5) - 1,  fx = X & 31  (same for Y with M4, M5, M3)
//  * weights: float table of the A = -0.75 cubic at the 32 sub-pixel phases, 2-D weight = float product
//  * the source is a 0/1 mask, so the interpolated value is the sum of the weights over the non-zero
//    taps; every partial sum of <= 16 such floats is exact in double (magnitudes 2^-21..1, 24-bit
//    mantissas), hence (sum != 0) does not depend on the summation order.
struct CubicTab { float c[32][4]; };

__global__ void __launch_bounds__(256) rotate_rowsum_kernel(const uint8_t* __restrict__ mask, int64_t stride, int h, int w,
                                                            int S, int oy, int ox, const double* __restrict__ inv_affine,
                                                            const CubicTab tab, int32_t* __restrict__ profiles) {
  const int y = blockIdx.x, a = blockIdx.y;
  const double* M = inv_affine + 6 * a;
  const double m0 = __ldg(M), m1 = __ldg(M + 1), m2 = __ldg(M + 2), m3 = __ldg(M + 3), m4 = __ldg(M + 4), m5 = __ldg(M + 5);
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m1, (double)y), m2), 1024.0)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m4, (double)y), m5), 1024.0)) + 16;
  int count = 0;
  for (int x = threadIdx.x; x < S; x += blockDim.x) {
    const int X = (X0 + __double2int_rn(__dmul_rn(__dmul_rn(m0, (double)x), 1024.0))) >> 5;
    const int Y = (Y0 + __double2int_rn(__dmul_rn(__dmul_rn(m3, (double)x), 1024.0))) >> 5;
    const int sx = min(max(X >> 5, -32768), 32767) - 1, sy = min(max(Y >> 5, -32768), 32767) - 1;
    // taps (clamped into the S x S padded image = BORDER_REPLICATE) that can fall inside the mask rectangle?
    const int x_lo = min(max(sx, 0), S - 1), x_hi = min(max(sx + 3, 0), S - 1);
    const int y_lo = min(max(sy, 0), S - 1), y_hi = min(max(sy + 3, 0), S - 1);
    if (x_hi < ox || x_lo >= ox + w || y_hi < oy || y_lo >= oy + h) continue;
    const float* wy = tab.c[Y & 31];
    const float* wx = tab.c[X & 31];
    double sum = 0.0;
    int nz = 0;
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      const int yy = min(max(sy + k1, 0), S - 1) - oy;
      if (yy < 0 || yy >= h) continue;
      const uint8_t* row = mask + (int64_t)yy * stride;
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        const int xx = min(max(sx + k2, 0), S - 1) - ox;
        if (xx < 0 || xx >= w) continue;
        if (__ldg(row + xx)) {
          sum = __dadd_rn(sum, (double)__fmul_rn(wy[k1], wx[k2]));
          ++nz;
        }
      }
    }
    count += (nz != 0 && sum != 0.0) ? 1 : 0;
  }
  // block reduction -> one int per (angle, row)
  __shared__ int red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = count;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    profiles[(int64_t)a * S + y] = t;
  }
}

}  // namespace sbb
