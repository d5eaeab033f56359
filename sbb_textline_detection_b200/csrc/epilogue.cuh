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

// Normalised input sample the network sees: uint8/255 as one IEEE fp32 division ==
// float32(float64(v)/255.0) for all 256 values (main.py:239 + Keras' fp32 cast).
__device__ __forceinline__ float norm_u8(uint8_t v) { return __fdiv_rn((float)v, 255.0f); }

// Where tile pixel (2Y+py, 2X+px) of image img lands and whether this tile's write survives the
// reference's 9-case margin crop + last-writer-wins stitch (main.py:294-364 == separable owner
// test).  Called BEFORE the accumulator wait so the table loads overlap the MMA main loop.
// mode 1 (plain tile batch): every pixel is kept, index into [n][TH][TW].
__device__ __forceinline__ bool head_owner(const HeadParams& h, int py, int px, int img, int Y, int X,
                                           int64_t* index) {
  const int y = 2 * Y + py, x = 2 * X + px;
  if (h.mode == 0) {
    const int4 org = __ldg(reinterpret_cast<const int4*>(h.tile_org) + img);
    const int px_ = org.x + x, py_ = org.y + y;
    *index = (int64_t)py_ * h.labels_row_stride + px_;
    return __ldg(h.owner_x + px_) == org.z && __ldg(h.owner_y + py_) == org.w;
  }
  *index = ((int64_t)img * h.TH + y) * h.TW + x;
  return true;
}

// dec5 epilogue.  f = the complete dec5 pre-activation of one pixel (up-sampled taps, input-skip
// taps and bias all come out of the MMA: the bias rides on the constant-1 pad channel of the
// packed input image).  w_cls/b_cls may point to shared or global memory.
__device__ __forceinline__ void head_finish(const HeadParams& h, const float* w_cls, const float* b_cls,
                                            int64_t pix, float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  // 1x1 classifier (+ folded BN)
  float z[8];
  const int C = h.n_classes;
#pragma unroll
  for (int c = 0; c < 8; ++c) z[c] = (c < C) ? b_cls[c] : -INFINITY;
  if (C <= 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float2 w = *reinterpret_cast<const float2*>(w_cls + j * 8);
      z[0] = fmaf(f[j], w.x, z[0]);
      z[1] = fmaf(f[j], w.y, z[1]);
    }
    if (C < 2) z[1] = -INFINITY;
  } else if (C <= 4) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float4 w = *reinterpret_cast<const float4*>(w_cls + j * 8);
      z[0] = fmaf(f[j], w.x, z[0]);
      z[1] = fmaf(f[j], w.y, z[1]);
      z[2] = fmaf(f[j], w.z, z[2]);
      z[3] = fmaf(f[j], w.w, z[3]);
    }
#pragma unroll
    for (int c = 2; c < 4; ++c)
      if (c >= C) z[c] = -INFINITY;
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < C) z[c] = fmaf(f[j], w_cls[j * 8 + c], z[c]);
    }
  }
  // np.argmax: first maximal class wins (main.py:290)
  int best = 0;
  float zb = z[0];
#pragma unroll
  for (int c = 1; c < 8; ++c)
    if (c < C && z[c] > zb) { zb = z[c]; best = c; }

  if (h.labels) h.labels[pix] = (uint8_t)best;
  if (h.mode != 0) {
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
