// Launch-parameter structs shared by the tcgen05 implicit-GEMM kernel, the SIMT cross-check kernel
// and the host planner (sbb_net.cu).
//
// Every convolution of the network is expressed as ONE implicit GEMM
//     D[m, n] = sum over segments s, channel c:  A_s[pixel(m) + (dx_s, dy_s), c] * B[n, k(s, c)]
// where a "view" is a 4-D strided window (C, W, H, N) onto an NHWC activation tensor (reads outside
// the window are zero -- that is the conv padding), and a "segment" is one filter tap applied to one
// view over a run of 64-channel chunks.  This one formalism covers 1x1 convs (1 segment), strided 1x1
// convs (a stride-2 view), 3x3 'same' convs (9 segments), a bottleneck's expand conv K-concatenated
// with its projection shortcut (2 segments on 2 views) and the decoder's
// upsample2x+concat+pad+conv3x3 (18 segments per output-parity class: 9 on the low-res tensor, 9 on
// parity sub-views of the skip tensor).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stddef.h>
#include <stdint.h>

namespace sbb {

constexpr int kMaxViews = 10;
constexpr int kMaxSegs = 20;
constexpr int kChunk = 64;  // channels per K chunk: 64 halves = one 128-byte swizzle row

struct RawView {          // what the SIMT kernel (and the tensor-map encoder) needs to know about a view
  const __half* base;     // hi-plane channel 0 of view element (0,0,0)
  int64_t sW, sH, sN;     // strides in halves
  int32_t W, H, N;        // extents (out-of-range reads give 0)
  int32_t lo_off;         // halves from a hi channel to its lo twin (0 in single-plane mode)
};

// Segment flags.  kSegPacked: the view's 64-half chunk interleaves (hi, lo) pairs of the SAME tensor
// ([c0h c1h c2h 1 | c0l c1l c2l 0] per input pixel, see StemParams), so ONE A tile carries both
// planes: main += A*B_hi (B_hi holds w_hi at the hi AND the lo slots), cross += A*B_lo (w_lo at the
// hi slots only).  Bits 4-7: number of 16-half K steps that carry non-zero weights (0 = all 4).
constexpr int kSegPacked = 1;
// kSegNtile: the segment's channel window follows the N tile (c0 += n-tile * BN) -- the residual of
// an identity block enters the accumulator as one more K segment against an identity weight block,
// so it rides the same deep TMA pipeline as the operands instead of a latency-exposed epilogue load.
constexpr int kSegNtile = 2;

struct SegDesc {
  int16_t view, dx, dy, c0, nchunks, flags;
};
// conv_gemm_tc.cuh reads {nchunks, flags} of a staged segment with one aligned 32-bit shared-memory load
static_assert(sizeof(SegDesc) == 12 && offsetof(SegDesc, nchunks) == 8 && offsetof(SegDesc, flags) == 10, "SegDesc layout");
__host__ __device__ inline int seg_ksteps(int flags) { return (flags >> 4) & 15 ? (flags >> 4) & 15 : 4; }

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
  CUtensorMap tmapBh;          // same weight matrix in boxes of BN/2 rows: a CTA of a pair loads half of the N tile
  CUtensorMap tmapOut;         // {32 ch, BW, BH, 1} boxes (64B swizzle) onto the output tensor, both planes
  CUtensorMap tmapOut2;        // merged column parities (head_px == -2): where columns [Cout/2, Cout) are stored
  RawView views[kMaxViews];
  SegDesc segs[kMaxSegs];
  int32_t n_segs, total_chunks, n_views;
  int32_t wide_n;              // split mode: issue A_hi x [B_hi; B_lo] as one N = 2*BN MMA
  int32_t a_hi_only;           // precision plan: this launch reads only the hi plane of its activations (A_lo is neither
                               // loaded nor multiplied: 2 MMA units per K step instead of 3); weights keep hi + lo
  int32_t win_chunks;          // K chunks accumulated inside TMEM before a flush into fp32 registers
  int32_t BW, BH;              // M tile = BW x BH pixels of BI consecutive images (BW*BH*BI <= 128)
  int32_t BI;
  int32_t n_tiles_n;
  int32_t Cout, Ktot;          // Ktot = total_chunks * 64
  const __half* wmat;          // [planes*Cout][Ktot]  (rows [Cout, 2*Cout) are the lo plane)
  const float* bias;           // [Cout]
  __half* out;                 // element (img, y, x, c) at out[img*oN + y*oH + x*oW + c]
#ifdef SBB_X_DIRECT_STORE
  __half* out2;                // merged column parities (head_px == -2): base of the px = 1 sub-view (same strides)
#endif
  int64_t oN, oH, oW;
  int32_t out_lo_off;
  int32_t relu;
  const __half* res;           // optional residual added in the epilogue (SBB_RES_IN_MMA=0 only), same indexing with r*
  int64_t rN, rH, rW;
  int32_t res_lo_off;
  int32_t planes;              // 1 (fp16) or 2 (fp16x3 split)
  int32_t head_py, head_px;    // HEAD: output parity of this variant (output pixel = (2Y+py, 2X+px));
                               // -1: merged -- columns [32p, 32p+32) belong to parity p = 2*py + px
                               // head_px == -2 (non-head): columns [0, Cout/2) are px = 0, [Cout/2, Cout) px = 1
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
  int32_t BW, BH, n_tiles_n, has_res;  // common to all variants (copied here: no global load needed)
  int32_t img0, x_off;         // sub-batch launches: first image (non-flat grids) / first pixel (flat grids) of this launch
  int32_t BI;                  // images per M tile: small maps (28x28, 14x14) fill the 128 MMA rows with boxes
                               // that span several images, e.g. {64 ch, 4, 4, 8 images}
#ifdef SBB_X_DIRECT_STORE      // experiment build (profiles/r02ac_direct_store_abab.txt: parity green, 7 % SLOWER per page)
  int32_t direct_store;        // CTA-pair kernel: the epilogue warps write their rows with st.global instead of TMA stores
#endif
  // chained launch (variant 0 = a bottleneck's expand conv, variant 1 = the NEXT block's reduce conv, both 1x1 on
  // the same flat 128-pixel M tiles): the expand conv's store threads count finished N tiles per M tile in
  // chain_flags[m tile] (2 epilogue groups x its N tiles = chain_need); a reduce-conv item loads its A rows only
  // after its M tile is complete.  The work list keeps the consumers a few hundred items behind their producers,
  // so the 512..2048-channel tensor is read back from L2 instead of HBM and one launch disappears.
  uint32_t* chain_flags;
  int32_t chain_need;
  int32_t debug;               // SBB_DEBUG bits (bottleneck experiments; results are WRONG when set):
                               // 1 skip the MMAs, 2 skip the A_lo loads, 4 skip the head/epilogue math,
                               // 8 skip ALL A loads (weights only)
  // SBB_DEBUG bit 16: per-CTA wait-cycle counters, 16 x uint32 per CTA (results stay correct):
  // [0] producer waits for a free smem stage, [1] MMA issuer waits for operands, [2] MMA issuer waits for a
  // drained TMEM buffer, [3] epilogue (group 0) waits for a finished window, [4] unused, [5] CTA lifetime,
  // [6] work items of this CTA, [7] epilogue group 0's store thread: wait for the previous store's smem read +
  // named barriers + st.shared + proxy fence + TMA store issue,
  // [8] MMA issuer: cycles inside the tcgen05.mma / tcgen05.commit issue block of a chunk
  uint32_t* role_cycles;
  HeadParams head;
};

struct WorkItem {
  int32_t variant, nt, img, x0, y0;
};
__device__ __forceinline__ WorkItem get_work(const LaunchArgs& a, int w, int BW, int BH, int n_tiles_n) {
  WorkItem k;
  if (a.worklist != nullptr) {
    const int4 e = __ldg(a.worklist + w);
    k.variant = e.x & 255; k.nt = e.x >> 8; k.img = e.y; k.x0 = e.z; k.y0 = e.w;
  } else {
    k.variant = 0;
    k.nt = w % n_tiles_n;
    const int m = w / n_tiles_n;
    const int tx = m % a.tiles_x;
    const int t2 = m / a.tiles_x;
    k.x0 = tx * BW + a.x_off;
    k.y0 = (t2 % a.tiles_y) * BH;
    k.img = (t2 / a.tiles_y) * a.BI + a.img0;
  }
  return k;
}

}  // namespace sbb
