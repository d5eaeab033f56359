// Launch-parameter structs shared by the tcgen05 implicit-GEMM kernel, the SIMT cross-check kernel
// and the host planner (sbb_net.cu).
//
// Every convolution of the network is expressed as ONE implicit GEMM
//     D[m, n] = sum over groups g, taps t of g, channels c:
//                   A_g[pixel(m) + (dx_t, dy_t), c] * B[n, k(g, t, c)]
// where a "view" is a 4-D strided window (C, W, H, N) onto an NHWC activation tensor (reads outside
// the window are zero -- that is the conv padding), a "tap" is one filter position and a "group" is
// a set of taps that read the SAME view over the same run of 64-channel chunks.  The A operand of a
// whole group is ONE smem tile per chunk: the output tile's pixels plus the halo the taps reach
// (rows of `P` pixels), and tap (dx, dy) is the same tile read `dy*P + dx` rows further down -- a
// 3x3 convolution fills shared memory once per chunk, not nine times.
// This one formalism covers 1x1 convs (1 group, 1 tap), strided 1x1 convs (a stride-2 view), 3x3
// 'same' convs (1 group of 9 taps), a bottleneck's expand conv K-concatenated with its projection
// shortcut or with its identity shortcut (2 groups), the 7x7/2 stem (2 groups of row taps over a
// packed image) and the decoder's upsample2x+concat+pad+conv3x3 (per output-parity class: 1 group of
// 4 merged taps on the low-res tensor + 4 groups on the parity sub-views of the skip tensor).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace sbb {

constexpr int kMaxViews = 8;
constexpr int kMaxGroups = 8;
constexpr int kMaxTaps = 16;
constexpr int kChunk = 64;  // channels per K chunk: 64 halves = one 128-byte swizzle row

// Non-flat launches tile every image into BW x BH output pixels laid out in rows of P = BW + 2 (the 2
// junk columns per row make tap shifts pure row offsets): 14 x 8 in rows of 16 (14 divides every
// feature-map width of the network, 14 * 2^k; the halo of a 3x3 group is (8+2) x 16 rows = 1.25 tiles)
// or 28 x 4 in rows of 30 where that wastes fewer MMA rows (28-pixel-wide maps).
constexpr int kHaloRows = 192;  // A stage rows: max tap shift (2*30 + 2 for a 3x3 group at P = 30) + 128

struct RawView {          // what the SIMT kernel (and the tensor-map encoder) needs to know about a view
  const __half* base;     // hi-plane channel 0 of view element (0,0,0)
  int64_t sW, sH, sN;     // strides in halves
  int32_t W, H, N;        // extents (out-of-range reads give 0)
  int32_t lo_off;         // halves from a hi channel to its lo twin (0 in single-plane mode)
};

// Group flags.  kGrpPacked: the view's 64-half chunk interleaves (hi, lo) pairs of the SAME tensor
// ([c0h c1h c2h 1 | c0l c1l c2l 0] per input pixel, see StemParams), so ONE A tile carries both
// planes: main += A*B_hi (B_hi holds w_hi at the hi AND the lo slots), cross += A*B_lo (w_lo at the
// hi slots only).  Bits 4-7: number of 16-half K steps that carry non-zero weights (0 = all 4).
constexpr int kGrpPacked = 1;
// kGrpNtile: the group's channel window follows the N tile (c0 += n-tile * BN) -- the residual of
// an identity block enters the accumulator as one more K group against an identity weight block,
// so it rides the same deep TMA pipeline as the operands instead of a latency-exposed epilogue load.
constexpr int kGrpNtile = 2;

struct TapDesc {
  int8_t dx, dy;    // tap offset in view pixels relative to the output pixel
  int16_t shift;    // rows into the group's A tile: (dy - oy) * P + (dx - ox)
  int16_t kcol;     // weight-matrix chunk column of (this tap, chunk 0); chunk c is kcol + c
};
struct GroupDesc {
  int16_t view, ox, oy;       // A box origin = tile origin + (ox, oy)  (the group's top-left tap)
  int16_t c0, nchunks, flags;
  int16_t ntaps, tap0;        // taps [tap0, tap0 + ntaps) of the variant's tap table
  int32_t a_bytes;            // bytes of ONE plane's A box
};
__host__ __device__ inline int grp_ksteps(int flags) { return (flags >> 4) & 15 ? (flags >> 4) & 15 : 4; }

struct HeadParams {       // dec5 epilogue: ReLU, 1x1 classifier (+ folded BN), (softmax), argmax,
                          // margin-crop + stitch  (main.py:287-364)
  const uint8_t* page;    // mode 0: uint8 BGR page (or tile-sized image), device pointer
  int64_t page_row_stride;
  const float* tiles;     // mode 1: float32 [n][TH][TW][3]
  const int32_t* tile_org;  // mode 0: per image {x0, y0, i, j}
  const int16_t* owner_x;   // mode 0: owner tile column per page x   (length page W)
  const int16_t* owner_y;   // mode 0: owner tile row per page y      (length page H)
  uint8_t* labels;        // mode 0: [H][W] page label map; mode 1: [n][TH][TW]
  int64_t labels_row_stride;
  float* probs;           // mode 1 optional [n][TH][TW][C]
  float* logits;          // mode 1 optional
  const float* w_cls;     // [32][8]
  const float* b_cls;     // [8]
  int32_t n_classes, TH, TW, mode;
};

// Static description of one implicit GEMM.  Lives in GLOBAL memory (the TMA descriptors are fetched
// from there); a launch may carry several "variants" that share the tile shape -- the four output
// parity classes of a decoder block run as ONE launch.
struct ConvParams {
  CUtensorMap tmapA[kMaxViews];
  CUtensorMap tmapB;
  CUtensorMap tmapOut;         // {32 ch, BW, BH, 1} boxes (64B swizzle) onto the output tensor, both planes
  RawView views[kMaxViews];
  GroupDesc groups[kMaxGroups];
  TapDesc taps[kMaxTaps];
  int32_t n_groups, n_taps, total_chunks, n_views;  // total_chunks = sum over groups of ntaps * nchunks
  int32_t wide_n;              // split mode: issue A_hi x [B_hi; B_lo] as one N = 2*BN MMA
  int32_t win_chunks;          // (tap, chunk) steps accumulated inside TMEM before a flush into fp32 registers
  int32_t BW, BH, P;           // M tile = BH rows of P pixels (BW valid) of one image; BH*P <= 128
  int32_t n_tiles_n;
  int32_t Cout, Ktot;          // Ktot = total_chunks * 64
  const __half* wmat;          // [planes*Cout][Ktot]  (rows [Cout, 2*Cout) are the lo plane)
  const float* bias;           // [Cout]
  __half* out;                 // element (img, y, x, c) at out[img*oN + y*oH + x*oW + c]
  int64_t oN, oH, oW;
  int32_t out_lo_off;
  int32_t relu;
  int32_t planes;              // 1 (fp16) or 2 (fp16x3 split)
  int32_t head_py, head_px;    // HEAD: output parity of this variant (output pixel = (2Y+py, 2X+px))
};

// Per-launch arguments (kernel parameter, by value).
struct LaunchArgs {
  const ConvParams* variants;  // device array
  int32_t n_variants;
  // explicit work list {variant | n-tile << 8, image, x0, y0}: decoder launches (4 parity variants,
  // tiles outside the region the stitch keeps are left out); nullptr = enumerate variant 0's grid
  const int4* worklist;
  int32_t total_work;
  int32_t GW, GH, NIMG;        // logical output grid of one variant
  int32_t tiles_x, tiles_y;
  int32_t BW, BH, P, n_tiles_n;  // common to all variants (copied here: no global load needed)
  int32_t debug;               // SBB_DEBUG bits (bottleneck experiments; results are WRONG when set):
                               // 1 skip the MMAs, 2 skip the A_lo loads, 4 skip the head/epilogue math,
                               // 8 skip ALL A loads (weights only)
  HeadParams head;
};

struct WorkItem {
  int32_t variant, nt, img, x0, y0;
};
__device__ __forceinline__ WorkItem get_work(const LaunchArgs& a, int w) {
  WorkItem k;
  if (a.worklist != nullptr) {
    const int4 e = __ldg(a.worklist + w);
    k.variant = e.x & 255; k.nt = e.x >> 8; k.img = e.y; k.x0 = e.z; k.y0 = e.w;
  } else {
    k.variant = 0;
    k.nt = w % a.n_tiles_n;
    const int m = w / a.n_tiles_n;
    const int tx = m % a.tiles_x;
    const int t2 = m / a.tiles_x;
    k.x0 = tx * a.BW;
    k.y0 = (t2 % a.tiles_y) * a.BH;
    k.img = t2 / a.tiles_y;
  }
  return k;
}

}  // namespace sbb
