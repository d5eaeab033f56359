// Persistent, warp-specialised implicit-GEMM convolution on the sm_100a tensor cores.
//
//   warp 0    : TMA producer   -- per K chunk: activation box(es) {64 ch, BW, BH, 1} of the segment's
//                                 view shifted by the tap offset (out-of-window -> zeros = padding),
//                                 plus the matching 64-wide slab of the weight matrix
//   warp 1    : MMA issuer     -- one elected lane issues tcgen05.mma (M=128, N=BN, K=16) into a
//                                 double-buffered TMEM accumulator; tcgen05.commit frees smem stages
//   warps 2-5 : epilogue       -- tcgen05.ld (thread == output pixel), bias/residual/ReLU, fp16 hi/lo
//                                 store; or (HEAD) the fused dec5 head: ReLU, classifier, argmax,
//                                 margin-crop + stitch into the page label map
//
// SPLIT (SBB_PREC_FP16X3): every operand is an fp16 (hi, lo) pair; per K step the issuer runs
//   hi*hi + hi*lo + lo*hi into the same fp32 accumulator (the lo*lo term is below fp32 resolution).
//
// smem per stage: A_hi [128 rows x 128 B] (+A_lo) | B_hi [BN rows x 128 B] (+B_lo), all written by TMA
// with the 128-byte swizzle the UMMA descriptors expect.
#pragma once
#include "epilogue.cuh"
#include "plan.h"
#include "ptx.cuh"

namespace sbb {

template <int BN, bool SPLIT>
struct TcCfg {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kPlanes = SPLIT ? 2 : 1;
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kBudget = 200 * 1024;
  static constexpr int kStages = (kBudget / kStageBytes) > 6 ? 6 : (kBudget / kStageBytes);
  static constexpr int kBufCols = kPlanes * BN;  // per TMEM buffer: hi*hi accumulator (+ cross-term accumulator)
  static constexpr int kTmemCols = (2 * kBufCols <= 32) ? 32 : (2 * kBufCols <= 64) ? 64 : (2 * kBufCols <= 128) ? 128 : (2 * kBufCols <= 256) ? 256 : 512;
  static_assert(2 * kBufCols <= 512, "TMEM has 512 columns");
  static constexpr int kHeadFloats = 32 * 8 + 8;
  // stages + barriers + tmem ptr + head constants + 1024 alignment slack
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + kHeadFloats * 4 + 1024;
  static constexpr int kThreads = 192;
};

template <int BN, bool SPLIT, bool HEAD>
__global__ void __launch_bounds__(192, 1) conv_gemm_tc_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = TcCfg<BN, SPLIT>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tmem_full = empty_bar + S;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_head = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int v = 0; v < p.n_views; ++v) ptx::prefetch_tmap(&p.tmapA[v]);
    ptx::prefetch_tmap(&p.tmapB);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 128);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  if (HEAD) {
    // stage the small fp32 head constants: w_cls[32*8] | b_cls[8]
    for (int i = threadIdx.x; i < Cfg::kHeadFloats; i += blockDim.x)
      s_head[i] = (i < 256) ? p.head.w_cls[i] : p.head.b_cls[i - 256];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int a_box_bytes = p.BW * p.BH * 128;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        const int nt = w % p.n_tiles_n;
        const int m = w / p.n_tiles_n;
        const int tx = m % p.tiles_x;
        const int t2 = m / p.tiles_x;
        const int ty = t2 % p.tiles_y;
        const int img = t2 / p.tiles_y;
        const int x0 = tx * p.BW, y0 = ty * p.BH, n0 = nt * BN;
        int kc = 0;
        for (int s = 0; s < p.n_segs; ++s) {
          const SegDesc sg = p.segs[s];
          const CUtensorMap* map = &p.tmapA[sg.view];
          const int lo = p.views[sg.view].lo_off;
          const bool two_a = SPLIT && !(sg.flags & kSegPacked);  // packed views carry hi and lo in ONE tile
          const uint32_t tx_bytes = (two_a ? 2 : 1) * a_box_bytes + Cfg::kPlanes * Cfg::kBBytes;
          for (int c = 0; c < sg.nchunks; ++c, ++kc) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            uint8_t* st = smem + stage * Cfg::kStageBytes;
            const int ch = sg.c0 + c * kChunk;
            ptx::tma_load_4d(st, map, &full_bar[stage], ch, x0 + sg.dx, y0 + sg.dy, img);
            if (two_a) ptx::tma_load_4d(st + Cfg::kABytes, map, &full_bar[stage], lo + ch, x0 + sg.dx, y0 + sg.dy, img);
            uint8_t* sb = st + Cfg::kPlanes * Cfg::kABytes;
            ptx::tma_load_2d(sb, &p.tmapB, &full_bar[stage], kc * kChunk, n0);
            if (SPLIT) ptx::tma_load_2d(sb + Cfg::kBBytes, &p.tmapB, &full_bar[stage], kc * kChunk, p.Cout + n0);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
      }
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
      int stage = 0;
      uint32_t phase = 0;
      uint32_t wc = 0;  // running window counter -> TMEM buffer + mbarrier phase
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        int kc = 0;       // chunk index inside this work unit
        int in_win = 0;   // chunks already issued into the current window
        uint32_t d_tmem = 0, d_cross = 0;
        for (int s = 0; s < p.n_segs; ++s) {
          const SegDesc sg = p.segs[s];
          const bool packed = (sg.flags & kSegPacked) != 0;
          const int ksteps = seg_ksteps(sg.flags);
          for (int c = 0; c < sg.nchunks; ++c, ++kc) {
            const int buf = wc & 1;
            if (in_win == 0) {  // open a window: wait until the epilogue has drained this TMEM buffer
              ptx::mbar_wait(&tmem_empty[buf], ((wc >> 1) & 1) ^ 1);
              ptx::tc_fence_after();
              d_tmem = tmem_base + buf * Cfg::kBufCols;
              d_cross = d_tmem + BN;
            }
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t a_hi = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
            const uint32_t a_lo = a_hi + Cfg::kABytes;
            const uint32_t b_hi = a_hi + Cfg::kPlanes * Cfg::kABytes;
            const uint32_t b_lo = b_hi + Cfg::kBBytes;
            for (int k = 0; k < ksteps; ++k) {  // UMMA_K = 16 halves = 32 bytes; 4 per 64-channel chunk
              const uint32_t acc = (in_win > 0 || k > 0) ? 1u : 0u;
              const uint64_t da_hi = ptx::make_smem_desc_sw128(a_hi + k * 32);
              const uint64_t db_hi = ptx::make_smem_desc_sw128(b_hi + k * 32);
              ptx::umma_f16(d_tmem, da_hi, db_hi, idesc, acc);
              if (SPLIT) {
                const uint64_t db_lo = ptx::make_smem_desc_sw128(b_lo + k * 32);
                ptx::umma_f16(d_cross, da_hi, db_lo, idesc, acc);
                if (!packed) {
                  const uint64_t da_lo = ptx::make_smem_desc_sw128(a_lo + k * 32);
                  ptx::umma_f16(d_cross, da_lo, db_hi, idesc, 1);
                }
              }
            }
            ptx::umma_commit(&empty_bar[stage]);  // smem stage reusable once these MMAs retire
            if (++stage == S) { stage = 0; phase ^= 1; }
            if (++in_win == p.win_chunks || kc + 1 == p.total_chunks) {
              ptx::umma_commit(&tmem_full[buf]);  // window complete -> epilogue
              in_win = 0;
              ++wc;
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int yl = row / p.BW, xl = row - yl * p.BW;
    uint32_t wc = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const int nt = w % p.n_tiles_n;
      const int m = w / p.n_tiles_n;
      const int tx = m % p.tiles_x;
      const int t2 = m / p.tiles_x;
      const int ty = t2 % p.tiles_y;
      const int img = t2 / p.tiles_y;
      const int x = tx * p.BW + xl, y = ty * p.BH + yl;
      const bool valid = (yl < p.BH) && (x < p.GW) && (y < p.GH);
      int64_t head_pix = 0;
      bool head_own = false;
      if (HEAD) {
        if (valid) head_own = head_owner(p.head, img, y, x, &head_pix);
      }
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.0f;
      for (int kc0 = 0; kc0 < p.total_chunks; kc0 += p.win_chunks, ++wc) {
        const int buf = wc & 1;
        ptx::mbar_wait(&tmem_full[buf], (wc >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * Cfg::kBufCols;
#pragma unroll
        for (int sl = 0; sl < BN / 32; ++sl) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(taddr + sl * 32, v);
          if (SPLIT) {
            uint32_t c[32];
            ptx::tmem_ld_32x32b_x32(taddr + BN + sl * 32, c);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[sl * 32 + j] += __uint_as_float(v[j]) + __uint_as_float(c[j]);
          } else {
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[sl * 32 + j] += __uint_as_float(v[j]);
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tmem_empty[buf]);
      }
      if (valid) {
        if (HEAD) {
          if (head_own) head_finish(p.head, s_head, s_head + 256, head_pix, *reinterpret_cast<float(*)[32]>(&acc[0]));
        } else {
#pragma unroll
          for (int sl = 0; sl < BN / 32; ++sl)
            epi_store32(p, img, y, x, nt * BN + sl * 32, *reinterpret_cast<float(*)[32]>(&acc[sl * 32]));
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace sbb
