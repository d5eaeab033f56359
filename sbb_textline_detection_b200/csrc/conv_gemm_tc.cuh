// Persistent, warp-specialised implicit-GEMM convolution on the sm_100a tensor cores.
//
//   warp 0    : TMA producer   -- per K chunk: activation box(es) {64 ch, BW, BH, BI images} of the segment's
//                                 view shifted by the tap offset (out-of-window -> zeros = padding),
//                                 plus the matching 64-wide slab of the weight matrix
//   warp 1    : MMA issuer     -- one elected lane issues tcgen05.mma (M=128, N=BN, K=16) into a
//                                 double-buffered TMEM accumulator; tcgen05.commit frees smem stages
//   warps 2.. : epilogue       -- G groups of four warps (G = 2 for BN >= 64), each group owning BN/G accumulator
//                                 columns: tcgen05.ld (thread == output pixel) -> fp32 registers; then per
//                                 32-channel slice: + bias -> ReLU -> fp16 hi/lo split -> swizzled smem staging
//                                 -> TMA bulk-tensor STORE (the hardware clips partial tiles); or (HEAD) the
//                                 fused dec5 head: ReLU, classifier, argmax, margin-crop + stitch into the page
//                                 label map (BN = 32: one output-parity class per item; BN = 128: the four
//                                 output pixels of a low-res pixel, 32 columns each, two per group)
//
// No epilogue thread touches global memory for activations: HBM latency is carried by the TMA
// engine (stores drained asynchronously), the threads only see smem.  An identity shortcut enters the
// accumulator as one more K segment (plan.h: kSegNtile); only the SBB_RES_IN_MMA=0 experiment adds a
// residual in the epilogue, with plain global loads.
//
// SPLIT (SBB_PREC_FP16X3): every operand is an fp16 (hi, lo) pair; per K step the issuer runs
//   hi*hi + hi*lo + lo*hi into fp32 accumulators (the lo*lo term is below fp32 resolution).
//
// smem: S stages { A_hi [128 rows x 128 B] (+A_lo) | B_hi [BN rows x 128 B] (+B_lo) } written by TMA
// with the 128-byte swizzle the UMMA descriptors expect, then kNStg staging slices
// { hi [128 rows x 64 B] (+lo) } in the 64-byte swizzle of the output tensor maps.
#pragma once
#include "epilogue.cuh"
#include "plan.h"
#include "ptx.cuh"

namespace sbb {

template <int BN, bool SPLIT, bool HEAD>
struct TcCfg {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kPlanes = SPLIT ? 2 : 1;
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kSliceBytes = 128 * 64;                 // one plane of a 32-channel slice
  static constexpr int kStgBytes = kPlanes * kSliceBytes;      // one staging buffer
#ifndef SBB_NSTG_SMALL
#define SBB_NSTG_SMALL 4
#endif
  static constexpr int kNStg = HEAD ? 0 : (BN == 128 ? 2 : SBB_NSTG_SMALL); // staging ring depth
  static constexpr int kTailBytes = HEAD ? 3072 : 2048;        // barriers + tmem ptr | variant cache | head constants
  static constexpr int kAvail = 232448 - 1024 - kTailBytes - kNStg * kStgBytes;
  static constexpr int kStages = (kAvail / kStageBytes) > 6 ? 6 : (kAvail / kStageBytes);
  static_assert(kStages >= 2, "pipeline needs at least two stages");
  // An MMA into accumulator columns that the previous MMA pair is still updating cannot start for ~83
  // cycles (measured, tools/ubench/mma_rate.cu), but an N <= 64 K step only takes 48-96: narrow tiles
  // rotate their K steps over kNCH independent accumulator sets that the epilogue adds up.
  static constexpr int kNCH = BN == 128 ? 1 : (BN == 64 ? 2 : 4);
  static constexpr int kChainCols = kPlanes * BN;        // hi*hi accumulator (+ cross-term accumulator)
  static constexpr int kBufCols = kNCH * kChainCols;     // per TMEM window buffer
  static constexpr int kTmemCols = (2 * kBufCols <= 32) ? 32 : (2 * kBufCols <= 64) ? 64 : (2 * kBufCols <= 128) ? 128 : (2 * kBufCols <= 256) ? 256 : 512;
  static_assert(2 * kBufCols <= 512, "TMEM has 512 columns");
  static constexpr int kHeadFloats = 32 * 8 + 8;
  static constexpr int kSmemBytes = kStages * kStageBytes + kNStg * kStgBytes + kTailBytes + 1024;
  // epilogue: groups of four warps (one per TMEM lane quarter) that split the accumulator columns
  static constexpr int kEpiGroups = BN >= 64 ? 2 : 1;
  static constexpr int kNStgGroup = kNStg / kEpiGroups;   // staging buffers per group
  static constexpr int kThreads = 32 * (2 + 4 * kEpiGroups);
};

// Per-variant control data staged in shared memory at kernel start: the single-thread producer /
// issuer loops would otherwise chase it through global memory (L1 is carved down to almost nothing
// by the 227 KB of smem, so every such load is an L2 round trip on the critical path).
struct VarCache {
  SegDesc segs[kMaxSegs];
  int32_t lo_off[kMaxViews];
  int32_t n_segs, total_chunks, win_chunks, wide_n, Cout, relu, out_lo_off, head_py, head_px, a_hi_only;
  const float* bias;
};
static_assert(sizeof(VarCache) * 4 + 256 <= 1600, "variant cache must fit the smem tail");

// byte offset of 16-byte chunk j (0..3) of row r inside a [128 rows x 64 B] slice stored with
// CU_TENSOR_MAP_SWIZZLE_64B (address bits [4,6) ^= bits [7,9))
__device__ __forceinline__ uint32_t stg_off(int r, int j) { return r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }

template <int BN, bool SPLIT, bool HEAD>
__global__ void __launch_bounds__((TcCfg<BN, SPLIT, HEAD>::kThreads), 1) conv_gemm_tc_kernel(const __grid_constant__ LaunchArgs a) {
  using Cfg = TcCfg<BN, SPLIT, HEAD>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + S * Cfg::kStageBytes;
  uint8_t* tail = stg + Cfg::kNStg * Cfg::kStgBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tmem_full = empty_bar + S;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  VarCache* s_var = reinterpret_cast<VarCache*>(tail + 256);
  float* s_head = reinterpret_cast<float*>(tail + 1600);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool has_res = !HEAD && a.has_res;
  const int BW = a.BW, BH = a.BH, n_tiles_n = a.n_tiles_n;

  if (warp == 0 && lane == 0) {
    for (int q = 0; q < a.n_variants; ++q) {
      const ConvParams& p = a.variants[q];
      for (int v = 0; v < p.n_views; ++v) ptx::prefetch_tmap(&p.tmapA[v]);
      ptx::prefetch_tmap(&p.tmapB);
      if (!HEAD) ptx::prefetch_tmap(&p.tmapOut);
    }
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4 * Cfg::kEpiGroups);   // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < a.n_variants * 32; i += blockDim.x) {
    const int q = i >> 5, j = i & 31;
    const ConvParams& p = a.variants[q];
    VarCache& c = s_var[q];
    if (j < kMaxSegs) c.segs[j] = p.segs[j];
    else if (j < kMaxSegs + kMaxViews) c.lo_off[j - kMaxSegs] = p.views[j - kMaxSegs].lo_off;
    else if (j == 30) {
      c.n_segs = p.n_segs; c.total_chunks = p.total_chunks; c.win_chunks = p.win_chunks; c.wide_n = p.wide_n;
      c.Cout = p.Cout; c.a_hi_only = p.a_hi_only;
    } else {
      c.relu = p.relu; c.out_lo_off = p.out_lo_off; c.head_py = p.head_py; c.head_px = p.head_px; c.bias = p.bias;
    }
  }
  if (HEAD) {
    // stage the small fp32 head constants: w_cls[32*8] | b_cls[8]
    for (int i = threadIdx.x; i < Cfg::kHeadFloats; i += blockDim.x)
      s_head[i] = (i < 256) ? a.head.w_cls[i] : a.head.b_cls[i - 256];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int a_box_bytes = BW * BH * a.BI * 128;  // rows of an A box: BW x BH pixels of BI images
  uint32_t* const prof = a.role_cycles ? a.role_cycles + blockIdx.x * 16 : nullptr;
  const uint32_t t_begin = prof ? (uint32_t)clock() : 0u;
  // mbarrier wait that (when profiling) charges the cycles it blocked to *acc
  auto timed_wait = [&](uint64_t* bar, uint32_t parity, uint32_t& acc) {
    if (prof) {
      const uint32_t t0 = (uint32_t)clock();
      ptx::mbar_wait(bar, parity);
      acc += (uint32_t)clock() - t0;
    } else {
      ptx::mbar_wait(bar, parity);
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t c_empty = 0, n_items = 0;
      WorkItem nxt = get_work(a, blockIdx.x, BW, BH, n_tiles_n);
      for (int w = blockIdx.x; w < a.total_work; w += gridDim.x) {
        const WorkItem wi = nxt;
        ++n_items;
        if (w + (int)gridDim.x < a.total_work) nxt = get_work(a, w + gridDim.x, BW, BH, n_tiles_n);  // prefetch
        const ConvParams& p = a.variants[wi.variant];  // only ADDRESSES of its TMA descriptors are taken
        const VarCache& vc = s_var[wi.variant];
        const int img = wi.img, x0 = wi.x0, y0 = wi.y0, n0 = wi.nt * BN;
        const int n_segs = vc.n_segs, cout = vc.Cout;
        int kc = 0;
        for (int s = 0; s < n_segs; ++s) {
          const SegDesc sg = vc.segs[s];
          const CUtensorMap* map = &p.tmapA[sg.view];
          const int lo = vc.lo_off[sg.view];
          // packed views carry hi and lo in ONE tile; a hi-only launch (precision plan) never touches A_lo
          const bool two_a = SPLIT && !(sg.flags & kSegPacked) && !vc.a_hi_only && !(a.debug & 2);
          const bool no_a = (a.debug & 8) != 0;
          const uint32_t tx_bytes = (no_a ? 0 : (two_a ? 2 : 1) * a_box_bytes) + Cfg::kPlanes * Cfg::kBBytes;
          for (int c = 0; c < sg.nchunks; ++c, ++kc) {
            timed_wait(&empty_bar[stage], phase ^ 1, c_empty);
            ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            uint8_t* st = smem + stage * Cfg::kStageBytes;
            const int ch = sg.c0 + c * kChunk + ((sg.flags & kSegNtile) ? n0 : 0);
            if (!no_a) ptx::tma_load_4d(st, map, &full_bar[stage], ch, x0 + sg.dx, y0 + sg.dy, img);
            if (two_a && !no_a) ptx::tma_load_4d(st + Cfg::kABytes, map, &full_bar[stage], lo + ch, x0 + sg.dx, y0 + sg.dy, img);
            uint8_t* sb = st + Cfg::kPlanes * Cfg::kABytes;
            ptx::tma_load_2d(sb, &p.tmapB, &full_bar[stage], kc * kChunk, n0);
            if (SPLIT) ptx::tma_load_2d(sb + Cfg::kBBytes, &p.tmapB, &full_bar[stage], kc * kChunk, cout + n0);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (prof) { prof[0] = c_empty; prof[6] = n_items; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The tensor core adds into its fp32 accumulator with truncation, so a long K chain picks up a
    // systematic bias.  Two counter-measures keep the result fp32-grade:
    //  (1) the chain is cut into windows of win_chunks K chunks: each window accumulates in one of two
    //      TMEM buffers starting from zero and the epilogue warps fold it into round-to-nearest fp32
    //      registers while the next window is being issued;
    //  (2) the small cross terms hi*lo + lo*hi (2^-11 of the main term) get their OWN accumulator, so
    //      they are never truncated at the ulp of the large hi*hi sum and do not add truncation steps to it.
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_m128(BN);
      // SPLIT: B_hi and B_lo sit back to back in smem (2*BN rows) and the main / cross accumulators back
      // to back in TMEM, so A_hi x [B_hi; B_lo] is ONE MMA of N = 2*BN (A is read from smem once)
      constexpr uint32_t idesc_wide = ptx::make_idesc_f16_m128(2 * BN);
      // This loop is ONE thread: for the short-K launches (conv1, the stage-2 convs, dec4/dec5: 4-8 MMAs per
      // chunk) its own instruction latency, not the tensor pipe, set the pace (ncu: 600 of 1190 cycles per
      // chunk in conv1, profiles/r01i_ncu_stall_sites_conv1.txt).  So: running shared-space addresses of the
      // ring barriers and the low descriptor word of the current stage instead of per-chunk address
      // arithmetic, one 32-bit smem load per segment, and compile-time accumulator chains -- the planner
      // guarantees an even number of K steps per chunk when kNCH <= 2 (sbb_net.cu: build_conv), so step k of
      // a chunk always lands on chain k % kNCH and only the window's first chunk zero-initialises.
      constexpr uint32_t kStageStep = Cfg::kStageBytes >> 4;
      constexpr uint32_t kALo = Cfg::kABytes >> 4, kB = (Cfg::kPlanes * Cfg::kABytes) >> 4, kBLo = Cfg::kBBytes >> 4;
      constexpr bool kLean = Cfg::kNCH <= 2;
      const uint32_t desc0 = ptx::smem_desc_lo_sw128(ptx::smem_u32(smem));
      const uint32_t full0 = ptx::smem_u32(full_bar), empty0 = ptx::smem_u32(empty_bar);
      uint32_t full_a = full0, empty_a = empty0, da = desc0;
#ifdef SBB_ISSUE_PROBE
      bool next_ready = false;
#endif
      int stage = 0;
      uint32_t phase = 0;
      uint32_t wc = 0;  // running window counter -> TMEM buffer + mbarrier phase
      uint32_t c_full = 0, c_tmem = 0, c_issue = 0;
      int var_nxt = a.worklist != nullptr ? (__ldg(&a.worklist[blockIdx.x].x) & 255) : 0;
      for (int w = blockIdx.x; w < a.total_work; w += gridDim.x) {
        const VarCache& vc = s_var[var_nxt];
        if (a.worklist != nullptr && w + (int)gridDim.x < a.total_work)
          var_nxt = __ldg(&a.worklist[w + gridDim.x].x) & 255;  // prefetch
        const bool wide = SPLIT && vc.wide_n;
        const bool lean = kLean && wide && !(a.debug & 1);
        const int n_segs = vc.n_segs, win_chunks = vc.win_chunks;
        int left = vc.total_chunks;  // chunks of this work unit still to issue
        int in_win = 0;   // chunks already issued into the current window
        uint32_t ks = 0;  // generic path: K steps already issued into the current window (-> chain, zero-init)
        uint32_t d_buf = 0;
        for (int s = 0; s < n_segs; ++s) {
          // {nchunks, flags} of the segment in one load (SegDesc: int16 view, dx, dy, c0, nchunks, flags)
          const uint32_t nf = *reinterpret_cast<const uint32_t*>(&vc.segs[s].nchunks);
          const int nchunks = (int)(nf & 0xFFFFu), flags = (int)(nf >> 16);
          // `packed` below only decides whether the A_lo x B_hi product is issued: a hi-only launch skips it too
          const bool packed = (flags & kSegPacked) != 0 || vc.a_hi_only != 0;
          const int ksteps = (a.debug & 1) ? 0 : seg_ksteps(flags);
          for (int c = 0; c < nchunks; ++c) {
            const uint32_t buf = wc & 1;
            if (in_win == 0) {  // open a window: wait until the epilogue has drained this TMEM buffer
              timed_wait(&tmem_empty[buf], ((wc >> 1) & 1) ^ 1, c_tmem);
              ptx::tc_fence_after();
              d_buf = tmem_base + buf * Cfg::kBufCols;
              ks = 0;
            }
#ifdef SBB_ISSUE_PROBE
            if (!next_ready)
#endif
            {
              if (prof) {
                const uint32_t t0 = (uint32_t)clock();
                ptx::mbar_wait_addr(full_a, phase);
                c_full += (uint32_t)clock() - t0;
              } else {
                ptx::mbar_wait_addr(full_a, phase);
              }
            }
            ptx::tc_fence_after();
#ifdef SBB_ISSUE_PROBE
            // EXPERIMENT (not in the default build, untested on the GPU so far -- DESIGN.md section 7): between the
            // two halves of this chunk's MMAs, where the thread is blocked behind the UTCHMMA queue anyway, probe the
            // NEXT stage's barrier so that the wait above is skipped when its operands have already landed.
            const uint32_t full_n = (stage + 1 == S) ? full0 : full_a + 8, phase_n = (stage + 1 == S) ? (phase ^ 1) : phase;
            next_ready = false;
#define SBB_PROBE_NEXT() next_ready = ptx::mbar_test_addr(full_n, phase_n)
#else
#define SBB_PROBE_NEXT()
#endif
            const uint32_t t_is = prof ? (uint32_t)clock() : 0u;
            // descriptor low words of this stage: A_hi | A_lo | B_hi | B_lo; a K step is +32 B = +2
            const uint32_t a_hi = da, a_lo = da + kALo, b_hi = da + kB, b_lo = b_hi + kBLo;
            if (lean) {
              const uint32_t acc0 = in_win != 0 ? 1u : 0u;  // the window's first chunk zero-initialises the chains
              if (!packed && ksteps == 4) {  // the common case
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint32_t d = d_buf + (k & (Cfg::kNCH - 1)) * Cfg::kChainCols;
                  ptx::umma_f16_lo(d, a_hi + 2 * k, b_hi + 2 * k, idesc_wide, k >= Cfg::kNCH ? 1u : acc0);
                  ptx::umma_f16_lo(d + BN, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
                  if (k == 1) { SBB_PROBE_NEXT(); }
                }
              } else if (packed && ksteps == 4) {  // one A tile carries hi and lo (stem)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  ptx::umma_f16_lo(d_buf + (k & (Cfg::kNCH - 1)) * Cfg::kChainCols, a_hi + 2 * k, b_hi + 2 * k, idesc_wide,
                                   k >= Cfg::kNCH ? 1u : acc0);
                  if (k == 1) { SBB_PROBE_NEXT(); }
                }
              } else if (packed && ksteps == 2) {  // dec5's input-skip rows
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  ptx::umma_f16_lo(d_buf + (k & (Cfg::kNCH - 1)) * Cfg::kChainCols, a_hi + 2 * k, b_hi + 2 * k, idesc_wide,
                                   k >= Cfg::kNCH ? 1u : acc0);
                  if (k == 0) { SBB_PROBE_NEXT(); }
                }
              } else {  // any other even step count
#pragma unroll 1
                for (int k = 0; k < ksteps; ++k) {
                  const uint32_t d = d_buf + (k & (Cfg::kNCH - 1)) * Cfg::kChainCols;
                  ptx::umma_f16_lo(d, a_hi + 2 * k, b_hi + 2 * k, idesc_wide, k >= Cfg::kNCH ? 1u : acc0);
                  if (!packed) ptx::umma_f16_lo(d + BN, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
                }
              }
            } else {
              // K step number j of the window goes to chain j % kNCH; the first kNCH steps zero-initialise
#pragma unroll 1
              for (int k = 0; k < ksteps; ++k) {  // UMMA_K = 16 halves = 32 bytes; 4 per 64-channel chunk
                const uint32_t j = ks + k;
                const uint32_t d = d_buf + (j & (Cfg::kNCH - 1)) * Cfg::kChainCols, acc = j >= (uint32_t)Cfg::kNCH ? 1u : 0u;
                if (wide) {
                  ptx::umma_f16_lo(d, a_hi + 2 * k, b_hi + 2 * k, idesc_wide, acc);
                  if (!packed) ptx::umma_f16_lo(d + BN, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
                } else {
                  ptx::umma_f16_lo(d, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
                  if (SPLIT) {
                    ptx::umma_f16_lo(d + BN, a_hi + 2 * k, b_lo + 2 * k, idesc, acc);
                    if (!packed) ptx::umma_f16_lo(d + BN, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
                  }
                }
              }
              ks += ksteps;
            }
#undef SBB_PROBE_NEXT
            ptx::umma_commit_addr(empty_a);  // smem stage reusable once these MMAs retire
            if (prof) c_issue += (uint32_t)clock() - t_is;
            if (++stage == S) { stage = 0; phase ^= 1; full_a = full0; empty_a = empty0; da = desc0; }
            else { full_a += 8; empty_a += 8; da += kStageStep; }
            --left;
            if (++in_win == win_chunks || left == 0) {
              ptx::umma_commit(&tmem_full[buf]);  // window complete -> epilogue
              in_win = 0;
              ++wc;
            }
          }
        }
      }
      if (prof) { prof[1] = c_full; prof[2] = c_tmem; prof[8] = c_issue; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2 .. 2+4*G)
    // G groups of four warps (one warp per TMEM lane quarter) split the accumulator COLUMNS of a tile:
    // group g owns columns [g*NCOL, (g+1)*NCOL) -- its 32-channel output slices, or (HEAD) its output
    // parities -- with its own staging buffers, named barrier and bulk-store issuing thread.  The launches
    // with few K chunks per tile (1x1 expand convs, conv1) were bound by four warps doing the bias / ReLU /
    // hi-lo split / store of 128 x BN outputs while the tensor pipe waited for a drained TMEM buffer.
    constexpr int G = Cfg::kEpiGroups;
    constexpr int NSL = BN / 32 / G;  // slices (HEAD: parities) per group
    constexpr int NCOL = BN / G;
    constexpr int NSTG_G = Cfg::kNStgGroup > 0 ? Cfg::kNStgGroup : 1;
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int yl = row / BW, xl = row - yl * BW;
    const bool issuer = (threadIdx.x == 64 + 128 * g);  // the one thread that owns this group's bulk-store groups
    uint8_t* const my_stg = stg + g * (NSTG_G * Cfg::kStgBytes);
    uint32_t wc = 0;
    uint32_t si = 0;  // running slice counter of this group -> staging buffer
    uint32_t c_win = 0, c_store = 0;
    WorkItem nxt = get_work(a, blockIdx.x, BW, BH, n_tiles_n);
    for (int w = blockIdx.x; w < a.total_work; w += gridDim.x) {
      const WorkItem wi = nxt;
      if (w + (int)gridDim.x < a.total_work) nxt = get_work(a, w + gridDim.x, BW, BH, n_tiles_n);  // prefetch
      const ConvParams& p = a.variants[wi.variant];
      const VarCache& vc = s_var[wi.variant];
      const int nt = wi.nt, img = wi.img, x0 = wi.x0, y0 = wi.y0;
      const int total_chunks = vc.total_chunks, win_chunks = vc.win_chunks;
      const bool in_grid = (yl < BH) && (x0 + xl < a.GW) && (y0 + yl < a.GH);
      // HEAD: BN = 32 is one output-parity class (variant), BN = 128 all four of a low-res pixel
      // (columns [32p, 32p+32) = parity p = 2*py + px)
      int64_t head_pix[NSL];
      uint32_t head_own = 0;
      if (HEAD && in_grid) {
#pragma unroll
        for (int pp = 0; pp < NSL; ++pp) {
          const int par = g * NSL + pp;
          const int py = BN == 32 ? vc.head_py : (par >> 1), px = BN == 32 ? vc.head_px : (par & 1);
          if (head_owner(a.head, py, px, img, y0 + yl, x0 + xl, &head_pix[pp])) head_own |= 1u << pp;
        }
      }
      float acc[NCOL];
#pragma unroll
      for (int j = 0; j < NCOL; ++j) acc[j] = 0.0f;
      for (int kc0 = 0; kc0 < total_chunks; kc0 += win_chunks, ++wc) {
        const int buf = wc & 1;
        timed_wait(&tmem_full[buf], (wc >> 1) & 1, c_win);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * Cfg::kBufCols + g * NCOL;
#pragma unroll
        for (int ch = 0; ch < Cfg::kNCH; ++ch) {
#pragma unroll
          for (int sl = 0; sl < NSL; ++sl) {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(taddr + ch * Cfg::kChainCols + sl * 32, v);
            if (SPLIT) {
              uint32_t c[32];
              ptx::tmem_ld_32x32b_x32(taddr + ch * Cfg::kChainCols + BN + sl * 32, c);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[sl * 32 + j] += __uint_as_float(v[j]) + __uint_as_float(c[j]);
            } else {
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[sl * 32 + j] += __uint_as_float(v[j]);
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();   // every lane's tcgen05.ld has completed (tmem_ld_wait above): one arrive per warp
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
      }
      if (HEAD) {
        if (!(a.debug & 4)) {
#pragma unroll
          for (int pp = 0; pp < NSL; ++pp)
            if (head_own >> pp & 1)
              head_finish(a.head, s_head, s_head + 256, head_pix[pp], *reinterpret_cast<float(*)[32]>(&acc[32 * pp]));
        }
      } else {
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl, ++si) {
          uint8_t* sh = my_stg + (si % NSTG_G) * Cfg::kStgBytes;   // hi plane of the slice; lo plane follows
          float* f = &acc[sl * 32];
          const int c0 = nt * BN + (g * NSL + sl) * 32;
          const float4* b4 = reinterpret_cast<const float4*>(vc.bias + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            f[4 * j + 0] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
          }
          if (has_res && in_grid) {
            // SBB_RES_IN_MMA=0 only (by default an identity residual is one more K segment of the MMA)
            const __half* r = p.res + img * p.rN + (int64_t)(y0 + yl) * p.rH + (int64_t)(x0 + xl) * p.rW + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float t[8];
              load8(r + 8 * j, t);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[8 * j + e] += t[e];
              if (SPLIT) {
                load8(r + p.res_lo_off + 8 * j, t);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[8 * j + e] += t[e];
              }
            }
          }
          if (vc.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          uint4 oh[4], ol[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __half2* h2 = reinterpret_cast<__half2*>(&oh[j]);
            __half2* l2 = reinterpret_cast<__half2*>(&ol[j]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = f[8 * j + 2 * e], c = f[8 * j + 2 * e + 1];
              const __half2 hh = __floats2half2_rn(a, c);
              const float2 back = __half22float2(hh);
              h2[e] = hh;
              l2[e] = __floats2half2_rn(a - back.x, c - back.y);
            }
          }
          const uint32_t t_st = prof ? (uint32_t)clock() : 0u;
          // the bulk store that last used this staging buffer (NSTG_G slices ago) must have read it out; the
          // math above ran while it did
          if (issuer) ptx::tma_store_wait_read<NSTG_G - 1>();
          ptx::named_bar_sync(1 + g, 128);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<uint4*>(sh + stg_off(row, j)) = oh[j];
            if (SPLIT) *reinterpret_cast<uint4*>(sh + Cfg::kSliceBytes + stg_off(row, j)) = ol[j];
          }
          ptx::fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA engine
          ptx::named_bar_sync(1 + g, 128);
          if (issuer) {
            ptx::tma_store_4d(&p.tmapOut, sh, c0, x0, y0, img);
            if (SPLIT) ptx::tma_store_4d(&p.tmapOut, sh + Cfg::kSliceBytes, vc.out_lo_off + c0, x0, y0, img);
            ptx::tma_store_commit();
          }
          if (prof) c_store += (uint32_t)clock() - t_st;
        }
      }
    }
    if (!HEAD && issuer) ptx::tma_store_wait_all();
    if (prof && issuer && g == 0) { prof[3] = c_win; prof[7] = c_store; }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (prof && threadIdx.x == 0) prof[5] = (uint32_t)clock() - t_begin;
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace sbb
