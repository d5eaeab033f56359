// Bandwidth-side kernels around the implicit GEMMs (tile gather/pad, stem BN+ReLU+maxpool) and the
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
  int32_t mode, TH, TW, nimg;
  int32_t PH, pitch;  // padded image: PH = TH + 6 rows of `pitch` pixels (pitch >= TW + 16)
  __half* xp;         // [nimg][PH][pitch][8 halves]: {c0h c1h c2h 1 | c0l c1l c2l 0}
};

// Tile extract (K10) + /255 + ZeroPadding2D(3): gathers each tile from the uint8 page (or a float
// tile batch) into a zero-bordered packed fp16 image, 16 bytes per pixel: the three BGR samples as
// (hi, lo) fp16 pairs plus a constant-1 channel that carries conv biases through the MMA.  The 7x7/2
// stem and the 3x3 input-skip taps of the last decoder block read it through OVERLAPPING-window TMA
// views (window = 8 consecutive pixels = one 64-half K chunk), so no im2col matrix ever exists.
__global__ void stem_pad_kernel(const StemParams s) {
  const int64_t total = (int64_t)s.nimg * s.PH * s.pitch;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(idx % s.pitch);
    const int64_t t = idx / s.pitch;
    const int yy = (int)(t % s.PH);
    const int img = (int)(t / s.PH);
    const int ty = yy - 3, tx = xx - 3;
    float v[3] = {0.0f, 0.0f, 0.0f};
    if (ty >= 0 && ty < s.TH && tx >= 0 && tx < s.TW) {
      if (s.mode == 0) {
        const int4 org = __ldg(reinterpret_cast<const int4*>(s.tile_org) + img);
        const uint8_t* q = s.page + (int64_t)(org.y + ty) * s.page_row_stride + (int64_t)(org.x + tx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = norm_u8(__ldg(q + c));
      } else {
        const float* q = s.tiles + (((int64_t)img * s.TH + ty) * s.TW + tx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __ldg(q + c);
      }
    }
    __align__(16) __half o[8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = __float2half_rn(v[c]);
      o[4 + c] = __float2half_rn(v[c] - __half2float(o[c]));
    }
    o[3] = __float2half_rn(1.0f);
    o[7] = __float2half_rn(0.0f);
    *reinterpret_cast<uint4*>(s.xp + idx * 8) = *reinterpret_cast<const uint4*>(o);
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
__global__ void __launch_bounds__(128) conv_simt_kernel(const ConvParams* __restrict__ pv, const LaunchArgs a) {
  const ConvParams& p = *pv;
  const int64_t per_img = (int64_t)a.GW * a.GH;
  const int64_t M = per_img * a.NIMG;
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int img = (int)(m / per_img);
  const int64_t r = m - img * per_img;
  const int y = (int)(r / a.GW), x = (int)(r - (int64_t)y * a.GW);
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
    const int bn = p.Cout >= 128 ? 128 : p.Cout;
    const __half* ap = v.base + img * v.sN + vy * v.sH + vx * v.sW + sg.c0 + ((sg.flags & kSegNtile) ? (n_base / bn) * bn : 0);
    for (int c = 0; c < sg.nchunks; ++c, ++kc) {
      if (!inb) continue;
      for (int g = 0; g < 8; ++g) {
        float a[8], t[8];
        load8(ap + c * kChunk + g * 8, a);
        if (split && !(sg.flags & kSegPacked)) {
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
    int64_t pix;
    // merged-parity head (head_py < 0): the 32-column group is the output parity
    const int py = p.head_py < 0 ? (int)(blockIdx.y >> 1) : p.head_py, px = p.head_py < 0 ? (int)(blockIdx.y & 1) : p.head_px;
    if (head_owner(a.head, py, px, img, y, x, &pix))
      head_finish(a.head, a.head.w_cls, a.head.b_cls, pix, acc);
  } else {
    epi_store32(p, img, y, x, n_base, acc);
  }
}

}  // namespace sbb
