// Epilogue device functions shared by the tcgen05 kernel and the SIMT cross-check kernel.
// One thread owns one output pixel and 32 consecutive output channels in fp32 registers.
#pragma once
#include "plan.h"

namespace sbb {

__device__ __forceinline__ void split_store8(__half* hi_ptr, __half* lo_ptr, const float* f, bool split) {
  __align__(16) __half hi[8];
  __align__(16) __half lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    hi[j] = __float2half_rn(f[j]);
    lo[j] = __float2half_rn(f[j] - __half2float(hi[j]));
  }
  *reinterpret_cast<uint4*>(hi_ptr) = *reinterpret_cast<const uint4*>(hi);
  if (split) *reinterpret_cast<uint4*>(lo_ptr) = *reinterpret_cast<const uint4*>(lo);
}

__device__ __forceinline__ void load8(const __half* ptr, float* f) {
  uint4 raw = __ldg(reinterpret_cast<const uint4*>(ptr));
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 v = __half22float2(h2[j]);
    f[2 * j] = v.x;
    f[2 * j + 1] = v.y;
  }
}

// acc[0..32) -> + bias (+ residual) (ReLU) -> fp16 hi(/lo) store at channels [n_base, n_base+32).
__device__ __forceinline__ void epi_store32(const ConvParams& p, int img, int y, int x, int n_base, float (&f)[32]) {
  const bool split = p.planes == 2;
  const float4* b4 = reinterpret_cast<const float4*>(p.bias + n_base);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 b = __ldg(b4 + j);
    f[4 * j + 0] += b.x;
    f[4 * j + 1] += b.y;
    f[4 * j + 2] += b.z;
    f[4 * j + 3] += b.w;
  }
  if (p.res != nullptr) {
    const __half* r = p.res + img * p.rN + y * p.rH + x * p.rW + n_base;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float t[8];
      load8(r + 8 * g, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[8 * g + j] += t[j];
      if (split) {
        load8(r + p.res_lo_off + 8 * g, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[8 * g + j] += t[j];
      }
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  }
  __half* o = p.out + img * p.oN + y * p.oH + x * p.oW + n_base;
#pragma unroll
  for (int g = 0; g < 4; ++g) split_store8(o + 8 * g, o + p.out_lo_off + 8 * g, f + 8 * g, split);
}

// Normalised input sample the network sees at tile pixel (yy, xx), channel c: uint8/255 as one IEEE
// fp32 division == float32(float64(v)/255.0) for all 256 values (main.py:239 + Keras' fp32 cast).
__device__ __forceinline__ float head_input(const HeadParams& h, int img, int ox, int oy, int yy, int xx, int c) {
  if (yy < 0 || yy >= h.TH || xx < 0 || xx >= h.TW) return 0.0f;  // ZeroPadding2D at the TILE border
  if (h.mode == 0) {
    uint8_t v = __ldg(h.page + (int64_t)(oy + yy) * h.page_row_stride + (int64_t)(ox + xx) * 3 + c);
    return __fdiv_rn((float)v, 255.0f);
  }
  return __ldg(h.tiles + (((int64_t)img * h.TH + yy) * h.TW + xx) * 3 + c);
}

// The 27 input samples (3x3 taps x BGR) the last decoder block's 'inp' skip needs for tile pixel
// (2Y+py, 2X+px).  Issued BEFORE the accumulator wait so the loads overlap the MMA main loop.
__device__ __forceinline__ void head_load_inputs(const HeadParams& h, int img, int Y, int X, float (&inp)[27]) {
  const int y = 2 * Y + h.py, x = 2 * X + h.px;
  int ox = 0, oy = 0;
  if (h.mode == 0) {
    const int4 org = __ldg(reinterpret_cast<const int4*>(h.tile_org) + img);
    ox = org.x; oy = org.y;
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int ky = t / 3, kx = t - 3 * (t / 3);
#pragma unroll
    for (int c = 0; c < 3; ++c) inp[t * 3 + c] = head_input(h, img, ox, oy, y + ky - 1, x + kx - 1, c);
  }
}

// dec5 epilogue.  f = accumulator over the 64 upsampled channels for parity-grid pixel (Y, X) of
// image img.  w_inp/w_cls/b_cls/bias may point to shared or global memory.
__device__ __forceinline__ void head_finish(const HeadParams& h, const float* w_inp, const float* w_cls,
                                            const float* b_cls, const float* bias32, int img, int Y, int X,
                                            const float (&inp)[27], float (&f)[32]) {
  const int y = 2 * Y + h.py, x = 2 * X + h.px;  // tile pixel
  int ox = 0, oy = 0, ti = 0, tj = 0;
  if (h.mode == 0) {
    const int4 org = __ldg(reinterpret_cast<const int4*>(h.tile_org) + img);
    ox = org.x; oy = org.y; ti = org.z; tj = org.w;
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] += bias32[j];
  // 3x3 conv over the 3 raw input channels (the 'inp' skip of the last decoder block), fp32 FMA
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const float a = inp[k];
    const float4* w4 = reinterpret_cast<const float4*>(w_inp + k * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = w4[j];
      f[4 * j + 0] = fmaf(a, w.x, f[4 * j + 0]);
      f[4 * j + 1] = fmaf(a, w.y, f[4 * j + 1]);
      f[4 * j + 2] = fmaf(a, w.z, f[4 * j + 2]);
      f[4 * j + 3] = fmaf(a, w.w, f[4 * j + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  // 1x1 classifier (+ folded BN)
  float z[8];
  const int C = h.n_classes;
#pragma unroll
  for (int c = 0; c < 8; ++c) z[c] = (c < C) ? b_cls[c] : -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) z[c] = fmaf(f[j], w_cls[j * 8 + c], z[c]);
  }
  // np.argmax: first maximal class wins (main.py:290)
  int best = 0;
  float zb = z[0];
#pragma unroll
  for (int c = 1; c < 8; ++c)
    if (c < C && z[c] > zb) { zb = z[c]; best = c; }

  if (h.mode == 0) {
    // 9-case margin crop + last-writer-wins stitch (main.py:294-364) == separable owner test
    const int px_ = ox + x, py_ = oy + y;
    if (h.owner_x[px_] == ti && h.owner_y[py_] == tj)
      h.labels[(int64_t)py_ * h.labels_row_stride + px_] = (uint8_t)best;
  } else {
    const int64_t pix = ((int64_t)img * h.TH + y) * h.TW + x;
    if (h.labels) h.labels[pix] = (uint8_t)best;
    if (h.logits) {
      for (int c = 0; c < C; ++c) h.logits[pix * C + c] = z[c];
    }
    if (h.probs) {
      float e[8], s = 0.0f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        e[c] = (c < C) ? expf(z[c] - zb) : 0.0f;
        s += e[c];
      }
      for (int c = 0; c < C; ++c) h.probs[pix * C + c] = e[c] / s;
    }
  }
}

}  // namespace sbb
