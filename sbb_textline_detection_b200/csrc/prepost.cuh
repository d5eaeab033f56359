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

}  // namespace sbb
