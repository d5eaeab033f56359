// Bandwidth-side kernels around the implicit GEMMs (stem im2col, stem BN+ReLU+maxpool) and the
// SIMT cross-check convolution (SBB_BACKEND_SIMT: same plan, plain CUDA cores, no TMA/tcgen05).
#pragma once
#include "epilogue.cuh"
#include "plan.h"

namespace sbb {

struct StemParams {
  // input source (same convention as HeadParams)
  const uint8_t* page;
  int64_t page_row_stride;
  const float* tiles;
  const int32_t* tile_org;
  int32_t mode, TH, TW, H1, W1, nimg, planes;
  __half* a1;  // [nimg*H1*W1][planes*192]: k = (ky*7+kx)*3 + c for k < 147, zero for 147..191
};

// Stem im2col (K10 + ZeroPadding2D(3) + 7x7/2 patch gather): one thread per (output pixel, 8-wide k
// group); reads the uint8 page directly (tile extract and /255 fused), writes fp16 hi(/lo).
__global__ void stem_im2col_kernel(const StemParams s) {
  const int64_t total = (int64_t)s.nimg * s.H1 * s.W1 * 24;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % 24);
    const int64_t m = idx / 24;
    const int ox = (int)(m % s.W1);
    const int64_t t = m / s.W1;
    const int oy = (int)(t % s.H1);
    const int img = (int)(t / s.H1);
    int px0 = 0, py0 = 0;
    if (s.mode == 0) {
      const int4 org = __ldg(reinterpret_cast<const int4*>(s.tile_org) + img);
      px0 = org.x; py0 = org.y;
    }
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float v = 0.0f;
      if (k < 147) {
        const int tap = k / 3, c = k - 3 * tap;
        const int ky = tap / 7, kx = tap - 7 * ky;
        const int yy = 2 * oy + ky - 3, xx = 2 * ox + kx - 3;
        if (yy >= 0 && yy < s.TH && xx >= 0 && xx < s.TW) {
          if (s.mode == 0) {
            const uint8_t u = __ldg(s.page + (int64_t)(py0 + yy) * s.page_row_stride + (int64_t)(px0 + xx) * 3 + c);
            v = __fdiv_rn((float)u, 255.0f);
          } else {
            v = __ldg(s.tiles + (((int64_t)img * s.TH + yy) * s.TW + xx) * 3 + c);
          }
        }
      }
      f[j] = v;
    }
    __half* o = s.a1 + m * (int64_t)(s.planes * 192) + g * 8;
    split_store8(o, o + 192, f, s.planes == 2);
  }
}

struct PoolParams {
  const __half* in;   // f1 raw conv1 output [n][H1][W1][planes*64]
  __half* out;        // [n][H2][W2][planes*64]
  const float* scale; // bn_conv1 folded scale[64], shift[64]
  const float* shift;
  int32_t nimg, H1, W1, H2, W2, planes;
};

// bn_conv1 + ReLU + MaxPooling2D(3x3, stride 2, 'valid'); one thread per (output pixel, 8 channels).
__global__ void stem_bn_relu_maxpool_kernel(const PoolParams q) {
  const int64_t total = (int64_t)q.nimg * q.H2 * q.W2 * 8;
  const int pix = q.planes * 64;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx & 7);
    const int64_t m = idx >> 3;
    const int ox = (int)(m % q.W2);
    const int64_t t = m / q.W2;
    const int oy = (int)(t % q.H2);
    const int img = (int)(t / q.H2);
    float sc[8], sh[8], best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(q.scale + g * 8 + j);
      sh[j] = __ldg(q.shift + g * 8 + j);
      best[j] = 0.0f;  // ReLU output is >= 0 and every window is fully inside ('valid')
    }
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) {
        const __half* p = q.in + (((int64_t)img * q.H1 + (2 * oy + dy)) * q.W1 + (2 * ox + dx)) * pix + g * 8;
        float a[8], b[8];
        load8(p, a);
        if (q.planes == 2) {
          load8(p + 64, b);
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] += b[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], fmaf(a[j], sc[j], sh[j]));
      }
    __half* o = q.out + m * pix + g * 8;
    split_store8(o, o + 64, best, q.planes == 2);
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT cross-check: thread == output pixel, 32 output channels per thread (blockIdx.y picks the
// channel group).  Reads the RawViews with explicit bounds checks instead of TMA zero fill and the
// weight matrix straight from global memory.  Operands are recombined (hi+lo) in fp32.
template <bool HEAD>
__global__ void __launch_bounds__(128) conv_simt_kernel(const ConvParams p) {
  const int64_t per_img = (int64_t)p.GW * p.GH;
  const int64_t M = per_img * p.NIMG;
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int img = (int)(m / per_img);
  const int64_t r = m - img * per_img;
  const int y = (int)(r / p.GW), x = (int)(r - (int64_t)y * p.GW);
  const int n_base = blockIdx.y * 32;
  const bool split = p.planes == 2;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
  int kc = 0;
  for (int s = 0; s < p.n_segs; ++s) {
    const SegDesc sg = p.segs[s];
    const RawView v = p.views[sg.view];
    const int vx = x + sg.dx, vy = y + sg.dy;
    const bool inb = vx >= 0 && vx < v.W && vy >= 0 && vy < v.H && img < v.N;
    const __half* ap = v.base + img * v.sN + vy * v.sH + vx * v.sW + sg.c0;
    for (int c = 0; c < sg.nchunks; ++c, ++kc) {
      if (!inb) continue;
      for (int g = 0; g < 8; ++g) {
        float a[8], t[8];
        load8(ap + c * kChunk + g * 8, a);
        if (split) {
          load8(ap + v.lo_off + c * kChunk + g * 8, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] += t[j];
        }
        const int k0 = kc * kChunk + g * 8;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float w[8], wl[8];
          load8(p.wmat + (int64_t)(n_base + j) * p.Ktot + k0, w);
          if (split) {
            load8(p.wmat + (int64_t)(p.Cout + n_base + j) * p.Ktot + k0, wl);
#pragma unroll
            for (int e = 0; e < 8; ++e) w[e] += wl[e];
          }
          float sum = acc[j];
#pragma unroll
          for (int e = 0; e < 8; ++e) sum = fmaf(a[e], w[e], sum);
          acc[j] = sum;
        }
      }
    }
  }
  if (HEAD) {
    float inp[27];
    head_load_inputs(p.head, img, y, x, inp);
    head_finish(p.head, p.head.w_inp, p.head.w_cls, p.head.b_cls, p.bias, img, y, x, inp, acc);
  } else {
    epi_store32(p, img, y, x, n_base, acc);
  }
}

}  // namespace sbb
