// CTA-pair variant of the implicit-GEMM convolution (conv_gemm_tc.cuh) for the K-heavy N = 128 launches:
// two CTAs of one cluster (one TPC) run ONE tcgen05.mma.cta_group::2 stream with M = 256 -- each CTA holds its
// own 128-pixel A tile and HALF of the weight tile, the tensor cores of both SMs read both halves.
//
// Why: the single-CTA kernel streams the same [B_hi; B_lo] weight tile (256 rows of 128 B per K chunk) into
// every SM; with A_hi and A_lo that is 512 operand rows per chunk and SM, and the K-heavy launches are bound by
// that delivery (L2 -> smem), not by the tensor pipe.  In pair mode a CTA loads 128 + 128
// rows of A and only 64 + 64 rows of B per chunk: 384 rows (-25 %), a stage shrinks from 64 KB to 48 KB and the
// ring grows from 3 to 4 stages.
//
//   both CTAs, warp 0 : TMA producer   -- own A boxes (hi, lo) + own half of B_hi / B_lo, every load signals the
//                                         LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   leader,    warp 1 : MMA issuer     -- per K step three M = 256, N = 128 MMAs
//                                            main  += A_hi x B_hi      cross += A_hi x B_lo      cross += A_lo x B_hi
//                                         tcgen05.commit multicast frees the stage / publishes the window in BOTH CTAs
//   both CTAs, warps 2..9 : epilogue   -- as in the single-CTA kernel (own 128 TMEM lanes = own pixels); a drained
//                                         TMEM window is reported to the LEADER's barrier (remote arrive for the peer)
//
// fp16 (hi, lo) operand split, TMEM windows and the cross-term accumulator are those of conv_gemm_tc.cuh.
//
// Template parameters: HEAD (the fused dec5 head: packed input-skip rows, classifier/argmax/stitch epilogue), BN (N
// tile: 128, or 64 for conv1 and the stage-2 2a / 2b convs -- 32 + 32 weight rows per CTA, 40 KB stages, two
// accumulator chains; the all-packed stem issues A x [B_hi half; B_lo half] as one N = 128 MMA), RESB (see PairCfg).
// The TMEM hand-over arrive has CTA scope on purpose: a cluster-scope release costs a MEMBAR.ALL.GPU per arrive and
// 5 % of the page (mbar_arrive_cluster below).
#pragma once
#include "conv_gemm_tc.cuh"

namespace sbb {

namespace ptx {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// Cluster-wide rendezvous without a memory release: at kernel start the mbarrier inits are published by
// fence.mbarrier_init.release.cluster, at the end nothing is published at all (the barrier only keeps both CTAs
// alive) -- a releasing arrive would cost a MEMBAR.ALL.GPU each time.
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Arrive on an mbarrier of ANOTHER CTA of the cluster.  Default semantics (.release at CTA scope): what is being
// handed over is a drained TMEM window, ordered by tcgen05.fence::before/after_thread_sync on both sides -- a
// cluster-scope release would add a MEMBAR.ALL.GPU per arrive (25 % of the epilogue warps' stall samples on dec2,
// profiles/r02y_ncu_full_summary_dec2_pair.txt) for memory this hand-over does not publish.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's smem, the bytes are counted on `bar_cluster`, a
// shared::cluster address that may name the peer (leader) CTA's mbarrier.
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs, 128 rows each] * B[smem of both CTAs, N/2 rows each]^T
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint32_t desc_a_lo, uint32_t desc_b_lo, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(desc_a_lo), "r"(desc_b_lo), "r"(kSmemDescHiSw128), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in both CTAs of the pair once all prior MMAs retired
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_f16_m256(int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}
}  // namespace ptx

// RESB (experiment, SBB_PAIR_RESB=1; N = 64 launches whose whole weight matrix is at most 9 K chunks: conv1, the
// stage-2 2a / 2b convs): the CTA's half of ALL weight chunks is loaded once and stays in shared memory; the ring then
// carries activations only.  Parity green, but no gain (profiles/r02x_resident_b_abab.txt): the weight rows are a fifth
// (3x3) to a third (stem) of these launches' TMA row requests, yet what bounds the stem is the request rate of its A
// boxes (packed 8-pixel windows that start every 32 bytes straddle two lines), and the 3x3 convs lose a ring stage.
template <bool HEAD, int BN_ = 128, bool RESB_ = false>
struct PairCfg {
  static constexpr int BN = BN_;
  static constexpr bool RESB = RESB_;
  static_assert(BN == 128 || (BN == 64 && !HEAD), "N tile: 128, or 64 for the non-head launches with 64 output channels");
  static_assert(!RESB || BN == 64, "resident weights: N = 64 launches only");
  static constexpr int kABytes = 128 * 128;            // one plane of this CTA's A tile
  static constexpr int kBHalf = (BN / 2) * 128;        // this CTA's half of one plane of the weight tile
  static constexpr int kResBChunks = RESB ? 9 : 0;     // K chunks of weights kept resident
  static constexpr int kResBBytes = kResBChunks * 2 * kBHalf;
  static constexpr int kStageBytes = 2 * kABytes + (RESB ? 0 : 2 * kBHalf);   // 48 KB (N = 128), 40 KB, or 32 KB (RESB)
  static constexpr int kSliceBytes = 128 * 64;
  static constexpr int kStgBytes = 2 * kSliceBytes;
  static constexpr int kNStg = HEAD ? 0 : 2;           // the fused head stores labels from registers
  static constexpr int kTailBytes = HEAD ? 3072 : 2048;   // barriers + tmem ptr | variant cache | head constants
  static constexpr int kStages = (232448 - 1024 - kTailBytes - kNStg * kStgBytes - kResBBytes) / kStageBytes;
  static_assert(kStages == (RESB ? 3 : 4), "four stages, three next to resident weights");
  // an MMA into accumulator columns the previous MMA is still updating cannot start for ~83 cycles but an N = 64
  // step only takes 32-64: narrow tiles rotate their K steps over two accumulator sets (as the single-CTA kernel)
  static constexpr int kNCH = BN == 128 ? 1 : 2;
  static constexpr int kChainCols = 2 * BN;            // main | cross
  static constexpr int kBufCols = kNCH * kChainCols;
  static constexpr int kTmemCols = 512;
  static_assert(2 * kBufCols <= 512, "TMEM has 512 columns");
  static constexpr int kSmemBytes = kStages * kStageBytes + kResBBytes + kNStg * kStgBytes + kTailBytes + 1024;
  static constexpr int kEpiGroups = 2;
  static constexpr int kThreads = 32 * (2 + 4 * kEpiGroups);
  static constexpr int kHeadFloats = 32 * 8 + 8;
};

// Work: `total_work` single items as in the single-CTA kernel; the pair kernel takes them two at a time.
//   enumerated grid: pair q -> N tile q % n_tiles_n, M tiles 2 * (q / n_tiles_n) + rank (a missing second tile
//                    runs on out-of-range coordinates: TMA zero-fills the loads and clips the stores)
//   work list:       items 2q and 2q + 1 (the host pairs items of the same variant and N tile)
__device__ __forceinline__ WorkItem pair_work(const LaunchArgs& a, int q, int rank, int n_tiles_n) {
  if (a.worklist != nullptr) return get_work(a, 2 * q + rank, a.BW, a.BH, n_tiles_n);
  const int nt = q % n_tiles_n, m = 2 * (q / n_tiles_n) + rank;
  return get_work(a, m * n_tiles_n + nt, a.BW, a.BH, n_tiles_n);
}

template <bool HEAD, int BN_ = 128, bool RESB_ = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((PairCfg<HEAD, BN_, RESB_>::kThreads), 1)
    conv_gemm_pair_kernel(const __grid_constant__ LaunchArgs a) {
  using Cfg = PairCfg<HEAD, BN_, RESB_>;
  constexpr bool RESB = Cfg::RESB;
  constexpr int S = Cfg::kStages, BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* resb = smem + S * Cfg::kStageBytes;          // RESB: chunk kc at resb + kc * 2 * kBHalf: {B_hi half | B_lo half}
  uint8_t* stg = resb + Cfg::kResBBytes;
  uint8_t* tail = stg + Cfg::kNStg * Cfg::kStgBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tmem_full = empty_bar + S;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* resb_bar = reinterpret_cast<uint64_t*>(tail + 128);   // leader: the resident weights of BOTH CTAs have landed
  VarCache* s_var = reinterpret_cast<VarCache*>(tail + 256);
  float* s_head = reinterpret_cast<float*>(tail + 1600);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int n_tiles_n = a.n_tiles_n;
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  // pairs of work items this launch has (see pair_work)
  const int n_pairs = a.worklist != nullptr ? (a.total_work >> 1)
                                            : ((a.total_work / n_tiles_n + 1) >> 1) * n_tiles_n;

  if (warp == 0 && lane == 0) {
    for (int q = 0; q < a.n_variants; ++q) {
      const ConvParams& p = a.variants[q];
      for (int v = 0; v < p.n_views; ++v) ptx::prefetch_tmap(&p.tmapA[v]);
      ptx::prefetch_tmap(&p.tmapBh);
      if (!HEAD) ptx::prefetch_tmap(&p.tmapOut);
      if (!HEAD && p.head_px == -2) ptx::prefetch_tmap(&p.tmapOut2);
    }
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full_bar[s], 1);    // leader: its producer's arrive.expect_tx (bytes of BOTH CTAs); peer: unused
      ptx::mbar_init(&empty_bar[s], 1);   // the leader's tcgen05.commit, multicast to both CTAs
    }
    if (RESB) ptx::mbar_init(resb_bar, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full[b], 1);   // the leader's tcgen05.commit, multicast to both CTAs
      ptx::mbar_init(&tmem_empty[b], 2 * 4 * Cfg::kEpiGroups);  // leader: one arrive per epilogue WARP of both CTAs
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_ptr, Cfg::kTmemCols);
    ptx::tmem_relinquish_pair();
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
    for (int i = threadIdx.x; i < Cfg::kHeadFloats; i += blockDim.x)
      s_head[i] = (i < 256) ? a.head.w_cls[i] : a.head.b_cls[i - 256];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // both CTAs' barriers are initialised before anything is signalled across the pair
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int a_box_bytes = a.BW * a.BH * a.BI * 128;
  // experiment build -DSBB_X_ROLES + SBB_DEBUG=16: the wait-cycle counters of LaunchArgs::role_cycles for this kernel
  // ([1], [2], [8] are the leader CTA's: only it issues MMAs)
#ifdef SBB_X_ROLES
  uint32_t* const prof = a.role_cycles ? a.role_cycles + blockIdx.x * 16 : nullptr;
  const uint32_t t_begin = (uint32_t)clock();
#define SBB_ROLE_WAIT(acc, call) do { const uint32_t t0_ = (uint32_t)clock(); call; acc += (uint32_t)clock() - t0_; } while (0)
#define SBB_ROLE(stmt) stmt
#else
#define SBB_ROLE_WAIT(acc, call) call
#define SBB_ROLE(stmt)
#endif

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      SBB_ROLE(uint32_t c_empty = 0; uint32_t n_items = 0; uint32_t c_flag = 0;)
      const uint32_t full_leader0 = ptx::mapa(ptx::smem_u32(full_bar), 0);
      if (RESB) {   // one variant, one N tile (host checks): this CTA's 32 + 32 weight rows of every K chunk, once
        const ConvParams& p0 = a.variants[0];
        const int nck = s_var[0].total_chunks, cout0 = s_var[0].Cout;
        const uint32_t rbar = ptx::mapa(ptx::smem_u32(resb_bar), 0);
        if (leader) ptx::mbar_arrive_expect_tx(resb_bar, 2u * nck * 2u * Cfg::kBHalf);
        for (int kc = 0; kc < nck; ++kc) {
          const uint32_t dst = ptx::smem_u32(resb + kc * 2 * Cfg::kBHalf);
          ptx::tma_load_2d_pair(dst, &p0.tmapBh, rbar, kc * kChunk, (int)rank * (BN / 2));
          ptx::tma_load_2d_pair(dst + Cfg::kBHalf, &p0.tmapBh, rbar, kc * kChunk, cout0 + (int)rank * (BN / 2));
        }
      }
      for (int q = cluster_id; q < n_pairs; q += n_clusters) {
        const WorkItem wi = pair_work(a, q, rank, n_tiles_n);
        const ConvParams& p = a.variants[wi.variant];
        const VarCache& vc = s_var[wi.variant];
        const int img = wi.img, x0 = wi.x0, y0 = wi.y0;
        const int n_row = wi.nt * BN + (int)rank * (BN / 2);   // this CTA's 64 weight rows of the N tile
        const int n_segs = vc.n_segs, cout = vc.Cout;
        int kc = 0;
        if (!HEAD && a.chain_flags != nullptr && wi.variant != 0) {
          // chained launch: this item reads what variant 0 stored for the same M tile earlier in the work list
          const uint32_t* flag = a.chain_flags + ((x0 - a.x_off) >> 7);
          uint32_t spins = 0;
          SBB_ROLE(const uint32_t t_fl = (uint32_t)clock();)
          while (ptx::ld_acquire_gpu(flag) < static_cast<uint32_t>(a.chain_need)) {
            __nanosleep(100);
            if (++spins > (1u << 23)) __trap();   // ~1 s: a broken work-list order must not hang the GPU
          }
          SBB_ROLE(c_flag += (uint32_t)clock() - t_fl;)
          ptx::fence_proxy_async_all();   // the acquire above orders the TMA loads below
        }
        for (int s = 0; s < n_segs; ++s) {
          const SegDesc sg = vc.segs[s];
          const CUtensorMap* map = &p.tmapA[sg.view];
          const int lo = vc.lo_off[sg.view];
          // packed views carry hi and lo in ONE tile; a hi-only launch (precision plan) never touches A_lo
          const bool two_a = !(sg.flags & kSegPacked) && !vc.a_hi_only;
          // bytes of BOTH CTAs land on the leader's barrier: 2 x (A_hi (+ A_lo) + B_hi half + B_lo half)
#ifdef SBB_X_NO_B   // experiment build (WRONG results): no weight loads at all -- what the B rows cost the TMA path
          const uint32_t tx_bytes = 2u * ((two_a ? 2u : 1u) * a_box_bytes);
#else
          const uint32_t tx_bytes = 2u * ((two_a ? 2u : 1u) * a_box_bytes + (RESB ? 0u : 2u * Cfg::kBHalf));
#endif
          for (int c = 0; c < sg.nchunks; ++c, ++kc) {
            SBB_ROLE_WAIT(c_empty, ptx::mbar_wait(&empty_bar[stage], phase ^ 1));
            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            const uint32_t st = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
            const uint32_t bar = full_leader0 + 8u * stage;
            const int ch = sg.c0 + c * kChunk + ((sg.flags & kSegNtile) ? wi.nt * BN : 0);
            ptx::tma_load_4d_pair(st, map, bar, ch, x0 + sg.dx, y0 + sg.dy, img);
            if (two_a) ptx::tma_load_4d_pair(st + Cfg::kABytes, map, bar, lo + ch, x0 + sg.dx, y0 + sg.dy, img);
#ifndef SBB_X_NO_B
            if (!RESB)
#else
            if (false)
#endif
            {
              ptx::tma_load_2d_pair(st + 2 * Cfg::kABytes, &p.tmapBh, bar, kc * kChunk, n_row);
              ptx::tma_load_2d_pair(st + 2 * Cfg::kABytes + Cfg::kBHalf, &p.tmapBh, bar, kc * kChunk, cout + n_row);
            }
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
        SBB_ROLE(++n_items;)
      }
      SBB_ROLE(if (prof) { prof[0] = c_empty; prof[6] = n_items; prof[9] = c_flag; })
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_m256(BN);
      constexpr uint32_t idesc_wide = ptx::make_idesc_f16_m256(2 * BN > 256 ? 256 : 2 * BN);
      constexpr uint32_t kStageStep = Cfg::kStageBytes >> 4;
      constexpr uint32_t kALo = Cfg::kABytes >> 4, kB = (2 * Cfg::kABytes) >> 4, kBLo = Cfg::kBHalf >> 4;
      const uint32_t desc0 = ptx::smem_desc_lo_sw128(ptx::smem_u32(smem));
      const uint32_t full0 = ptx::smem_u32(full_bar), empty0 = ptx::smem_u32(empty_bar);
      uint32_t full_a = full0, empty_a = empty0, da = desc0;
      const uint32_t resb_desc0 = ptx::smem_desc_lo_sw128(ptx::smem_u32(resb));
      if (RESB) {
        ptx::mbar_wait(resb_bar, 0);
        ptx::tc_fence_after();
      }
      int stage = 0;
      uint32_t phase = 0;
      uint32_t wc = 0;  // running window counter -> TMEM buffer + mbarrier phase
      SBB_ROLE(uint32_t c_full = 0; uint32_t c_tmem = 0; uint32_t c_issue = 0;)
      for (int q = cluster_id; q < n_pairs; q += n_clusters) {
        const int variant = a.worklist != nullptr ? (__ldg(&a.worklist[2 * q].x) & 255) : 0;
        uint32_t b_res = resb_desc0;   // RESB: descriptor of the current K chunk's resident weights
        const VarCache& vc = s_var[variant];
        const int win_chunks = vc.win_chunks;
        const int n_segs = vc.n_segs;
        int left = vc.total_chunks;
        int in_win = 0;
        uint32_t d_buf = 0;
        for (int s = 0; s < n_segs; ++s) {
          // {nchunks, flags} of the segment in one load (SegDesc: int16 view, dx, dy, c0, nchunks, flags)
          const uint32_t nf = *reinterpret_cast<const uint32_t*>(&vc.segs[s].nchunks);
          const int nchunks = (int)(nf & 0xFFFFu), flags = (int)(nf >> 16);
          const bool packed = (flags & kSegPacked) != 0 || vc.a_hi_only != 0;   // no A_lo x B_hi product either way
          const int ksteps = seg_ksteps(flags);
          for (int c = 0; c < nchunks; ++c) {
            const uint32_t buf = wc & 1;
            if (in_win == 0) {  // open a window: both CTAs' epilogues have drained this TMEM buffer
              SBB_ROLE_WAIT(c_tmem, ptx::mbar_wait(&tmem_empty[buf], ((wc >> 1) & 1) ^ 1));
              ptx::tc_fence_after();
              d_buf = tmem_base + buf * Cfg::kBufCols;
            }
            SBB_ROLE_WAIT(c_full, ptx::mbar_wait_addr(full_a, phase));
            ptx::tc_fence_after();
            SBB_ROLE(const uint32_t t_is = (uint32_t)clock();)
            const uint32_t a_hi = da, a_lo = da + kALo, b_hi = RESB ? b_res : da + kB, b_lo = b_hi + kBLo;
            b_res += 2 * kBLo;
            const uint32_t acc0 = in_win != 0 ? 1u : 0u;  // the window's first chunk zero-initialises both accumulators
            // K step k of a chunk goes to accumulator chain k % kNCH; the window's first chunk zero-initialises each
            // chain with its first step (every chunk carries a multiple of kNCH steps: build_conv checks)
            if (!packed) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // UMMA_K = 16 halves = 32 bytes
                const uint32_t d = d_buf + (k & (Cfg::kNCH - 1)) * Cfg::kChainCols;
                const uint32_t acc = k >= Cfg::kNCH ? 1u : acc0;
                ptx::umma_f16_pair(d, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);        // main  += A_hi x B_hi
                ptx::umma_f16_pair(d + BN, a_hi + 2 * k, b_lo + 2 * k, idesc, acc);   // cross += A_hi x B_lo
                ptx::umma_f16_pair(d + BN, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);    // cross += A_lo x B_hi
              }
            } else if (BN == 64 && vc.wide_n == 2) {
              // every segment of the launch is packed (the stem): there is no A_lo x B_hi product that would have to
              // land in the cross columns, so A x [B_hi half; B_lo half] -- contiguous in each CTA's stage -- is ONE
              // N = 128 MMA per K step instead of two N = 64 ones (which cost the same each).  Accumulator columns
              // then read {CTA0: main 32 | cross 32, CTA1: main 32 | cross 32}; the epilogue knows (wide_n == 2).
#pragma unroll 1
              for (int k = 0; k < ksteps; ++k)
                ptx::umma_f16_pair(d_buf + (k & (Cfg::kNCH - 1)) * Cfg::kChainCols, a_hi + 2 * k, b_hi + 2 * k, idesc_wide,
                                   k >= Cfg::kNCH ? 1u : acc0);
            } else {
              // one A tile interleaves (hi, lo): B_hi holds w_hi at the hi AND lo slots, B_lo w_lo at the hi slots
#pragma unroll 1
              for (int k = 0; k < ksteps; ++k) {
                const uint32_t d = d_buf + (k & (Cfg::kNCH - 1)) * Cfg::kChainCols;
                const uint32_t acc = k >= Cfg::kNCH ? 1u : acc0;
                ptx::umma_f16_pair(d, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
                ptx::umma_f16_pair(d + BN, a_hi + 2 * k, b_lo + 2 * k, idesc, acc);
              }
            }
            ptx::umma_commit_pair(empty_a);  // the stage is reusable in BOTH CTAs once these MMAs retire
            SBB_ROLE(c_issue += (uint32_t)clock() - t_is;)
            if (++stage == S) { stage = 0; phase ^= 1; full_a = full0; empty_a = empty0; da = desc0; }
            else { full_a += 8; empty_a += 8; da += kStageStep; }
            --left;
            if (++in_win == win_chunks || left == 0) {
              ptx::umma_commit_pair(ptx::smem_u32(&tmem_full[buf]));  // window complete -> both epilogues
              in_win = 0;
              ++wc;
            }
          }
        }
      }
      SBB_ROLE(if (prof) { prof[1] = c_full; prof[2] = c_tmem; prof[8] = c_issue; })
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, warps 2..9)
    constexpr int G = Cfg::kEpiGroups;
    constexpr int NSL = BN / 32 / G;
    constexpr int NCOL = BN / G;
    const int g = (warp - 2) >> 2;
    const int q4 = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q4 * 32 + lane;
    const bool issuer = (threadIdx.x == 64 + 128 * g);
    uint8_t* const my_stg = stg + g * Cfg::kStgBytes;
    const int yl = row / a.BW, xl = row - yl * a.BW;
    const uint32_t empty_leader0 = ptx::mapa(ptx::smem_u32(tmem_empty), 0);
    uint32_t wc = 0;
    int chain_pend = -1;   // chained launch: M tile whose completion count this store thread still owes
    SBB_ROLE(uint32_t c_win = 0; uint32_t c_store = 0;)
    for (int q = cluster_id; q < n_pairs; q += n_clusters) {
      const WorkItem wi = pair_work(a, q, rank, n_tiles_n);
      const ConvParams& p = a.variants[wi.variant];
      const VarCache& vc = s_var[wi.variant];
      const int nt = wi.nt, img = wi.img, x0 = wi.x0, y0 = wi.y0;
      const int total_chunks = vc.total_chunks, win_chunks = vc.win_chunks;
      // HEAD (merged parities): columns [32p, 32p+32) = output parity p = 2*py + px of this thread's low-res pixel
      int64_t head_pix[NSL];
      uint32_t head_own = 0;
      if (HEAD && (yl < a.BH) && (x0 + xl < a.GW) && (y0 + yl < a.GH)) {
#pragma unroll
        for (int pp = 0; pp < NSL; ++pp) {
          const int par = g * NSL + pp;
          if (head_owner(a.head, par >> 1, par & 1, img, y0 + yl, x0 + xl, &head_pix[pp])) head_own |= 1u << pp;
        }
      }
      // direct stores (experiment build -DSBB_X_DIRECT_STORE, env SBB_DIRECT_STORE=1; parity green but 7 % slower per page than
      // the bulk stores, profiles/r02ac_direct_store_abab.txt): after staging, lane l writes 16-byte piece l & 3 of rows q4*32 + 8*i + (l >> 2), i = 0..3 --
      // a quarter-warp covers two whole 64-byte channel runs.  What TMA clips for the bulk stores is tested here.
#ifdef SBB_X_DIRECT_STORE
      int64_t st_off[4];
      uint32_t st_ok = 0;
      if (!HEAD && a.direct_store) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = q4 * 32 + 8 * i + (lane >> 2);
          const int ry = r / a.BW, rx = r - ry * a.BW, ri = ry / a.BH, yy = ry - ri * a.BH;
          const int X = x0 + rx, Y = y0 + yy, I = img + ri;
          if (ri < a.BI && static_cast<uint32_t>(X - a.x_off) < static_cast<uint32_t>(a.GW) &&
              static_cast<uint32_t>(Y) < static_cast<uint32_t>(a.GH) &&
              static_cast<uint32_t>(I - a.img0) < static_cast<uint32_t>(a.NIMG))
            st_ok |= 1u << i;
          st_off[i] = I * p.oN + Y * p.oH + X * p.oW;
        }
      }
#endif
      if (!HEAD && issuer && chain_pend >= 0 && wi.variant != 0) {
        // a consumer item may (transitively) wait for the count this thread still owes: post it before the item can
        // block anything -- its first window is thousands of cycles of MMAs away, the store latency hides behind them
        ptx::tma_store_wait_all();
        ptx::red_release_gpu_add(a.chain_flags + chain_pend, 1u);
        chain_pend = -1;
      }
      float acc[NCOL];
#pragma unroll
      for (int j = 0; j < NCOL; ++j) acc[j] = 0.0f;
      for (int kc0 = 0; kc0 < total_chunks; kc0 += win_chunks, ++wc) {
        const int buf = wc & 1;
        SBB_ROLE_WAIT(c_win, ptx::mbar_wait(&tmem_full[buf], (wc >> 1) & 1));
        ptx::tc_fence_after();
        // columns of this group's main / cross sums inside a chain: {main BN | cross BN}, or for the all-packed N = 64
        // launch (wide_n == 2, see the issuer) {CTA0's 32: main | cross, CTA1's 32: main | cross}
        const bool wide_packed = BN == 64 && vc.wide_n == 2;
        const uint32_t main_off = wide_packed ? g * 64 : g * NCOL, cross_off = wide_packed ? g * 64 + 32 : BN + g * NCOL;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * Cfg::kBufCols;
#pragma unroll
        for (int ch = 0; ch < Cfg::kNCH; ++ch) {
#pragma unroll
          for (int sl = 0; sl < NSL; ++sl) {
            uint32_t v[32], c[32];
            ptx::tmem_ld_32x32b_x32(taddr + ch * Cfg::kChainCols + main_off + sl * 32, v);
            ptx::tmem_ld_32x32b_x32(taddr + ch * Cfg::kChainCols + cross_off + sl * 32, c);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[sl * 32 + j] += __uint_as_float(v[j]) + __uint_as_float(c[j]);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) ptx::mbar_arrive(&tmem_empty[buf]);
          else ptx::mbar_arrive_cluster(empty_leader0 + 8u * buf);
        }
      }
      if (HEAD) {
#pragma unroll
        for (int pp = 0; pp < NSL; ++pp)
          if (head_own >> pp & 1)
            head_finish(a.head, s_head, s_head + 256, head_pix[pp], *reinterpret_cast<float(*)[32]>(&acc[32 * pp]));
      }
#pragma unroll
      for (int sl = 0; sl < (HEAD ? 0 : NSL); ++sl) {
        uint8_t* sh = my_stg;   // one staging buffer per group: hi plane of the slice, lo plane follows
        float* f = &acc[sl * 32];
        const int c0 = nt * BN + (g * NSL + sl) * 32;
        const float4* b4 = reinterpret_cast<const float4*>(vc.bias + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          f[4 * j + 0] += bb.x; f[4 * j + 1] += bb.y; f[4 * j + 2] += bb.z; f[4 * j + 3] += bb.w;
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
            const float x = f[8 * j + 2 * e], y = f[8 * j + 2 * e + 1];
            const __half2 hh = __floats2half2_rn(x, y);
            const float2 back = __half22float2(hh);
            h2[e] = hh;
            l2[e] = __floats2half2_rn(x - back.x, y - back.y);
          }
        }
#ifdef SBB_X_DIRECT_STORE
        if (a.direct_store) {
          // rows q4*32 .. q4*32+31 of the staging buffer belong to this warp alone: no CTA-level barrier, no proxy fence
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<uint4*>(sh + stg_off(row, j)) = oh[j];
            *reinterpret_cast<uint4*>(sh + Cfg::kSliceBytes + stg_off(row, j)) = ol[j];
          }
          __syncwarp();
          const bool split_out = vc.head_px == -2;
          __half* const ob = ((split_out && g == 1) ? p.out2 : p.out) + (split_out ? sl * 32 : c0) + (lane & 3) * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (st_ok >> i & 1) {
              const uint32_t so = stg_off(q4 * 32 + 8 * i + (lane >> 2), lane & 3);
              const uint4 vh = *reinterpret_cast<const uint4*>(sh + so);
              const uint4 vl = *reinterpret_cast<const uint4*>(sh + Cfg::kSliceBytes + so);
              *reinterpret_cast<uint4*>(ob + st_off[i]) = vh;
              *reinterpret_cast<uint4*>(ob + st_off[i] + vc.out_lo_off) = vl;
            }
          }
          continue;
        }
#endif
        SBB_ROLE(const uint32_t t_st = (uint32_t)clock();)
        if (issuer) ptx::tma_store_wait_read<0>();   // the previous slice's bulk store has read the buffer out
        ptx::named_bar_sync(1 + g, 128);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(sh + stg_off(row, j)) = oh[j];
          *reinterpret_cast<uint4*>(sh + Cfg::kSliceBytes + stg_off(row, j)) = ol[j];
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1 + g, 128);
        if (issuer) {
          // merged column parities (head_px == -2): group g's 64 columns are the 64 channels of parity px = g
          const bool split_out = vc.head_px == -2;
          const CUtensorMap* omap = (split_out && g == 1) ? &p.tmapOut2 : &p.tmapOut;
          const int oc = split_out ? sl * 32 : c0;
          // SBB_X_NO_STORE / SBB_X_NO_LO_STORE: experiment builds only (WRONG results): what the store path costs
#ifndef SBB_X_NO_STORE
          ptx::tma_store_4d(omap, sh, oc, x0, y0, img);
#ifndef SBB_X_NO_LO_STORE
          ptx::tma_store_4d(omap, sh + Cfg::kSliceBytes, vc.out_lo_off + oc, x0, y0, img);
#endif
#endif
          ptx::tma_store_commit();
        }
        SBB_ROLE(c_store += (uint32_t)clock() - t_st;)
      }
      if (!HEAD && a.chain_flags != nullptr && issuer) {
        // chained launch: a variant-0 item's part of the M tile counts for the tile's consumers once its bulk stores are
        // COMPLETE (not just read out of the staging buffer).  Waiting for that right away would stall this thread for
        // a store latency per item: the count is posted one item later, when only the newer item's NSL groups may
        // still be pending (the work list keeps consumers > 2 waves behind, so the delay is never waited for)
        if (chain_pend >= 0) {
          ptx::tma_store_wait_pending<NSL>();   // completion makes the bulk writes visible to this thread; the release below publishes them
          ptx::red_release_gpu_add(a.chain_flags + chain_pend, 1u);
        }
        chain_pend = wi.variant == 0 ? ((x0 - a.x_off) >> 7) : -1;
      }
    }
    if (!HEAD && issuer && chain_pend >= 0) {
      ptx::tma_store_wait_all();
      ptx::red_release_gpu_add(a.chain_flags + chain_pend, 1u);
    }
    SBB_ROLE(if (prof && issuer && g == 0) { prof[3] = c_win; prof[7] = c_store; })
    if (!HEAD && issuer) ptx::tma_store_wait_all();
  }

  SBB_ROLE(if (prof && threadIdx.x == 0) prof[5] = (uint32_t)clock() - t_begin;)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // no CTA leaves (or frees TMEM) while its partner may still signal or read it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

#undef SBB_ROLE_WAIT
#undef SBB_ROLE
}  // namespace sbb
