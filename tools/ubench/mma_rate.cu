// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (operands in shared memory, SS mode) for the
// instruction mixes the conv kernel uses.  Prints cycles per K=16 step.  Data is whatever is in smem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../sbb_textline_detection_b200/csrc/ptx.cuh"
using namespace sbb;

__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(bar)), "h"((uint16_t)1) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_m(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

// mode 0: N1 only; mode 1: alternate N1 (A0) and N2 (A1)
template <int TWO_CTA>
__global__ void __launch_bounds__(192, 1) k(int n1, int n2, int mode, int iters, int tma_noise, const __grid_constant__ CUtensorMap tm, long long* out,
                                            int commit_every, int ldtm_noise, int shift_rows, int rotate) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, nbar[4], cbar;
  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (TWO_CTA) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) { ptx::mbar_init(&cbar, 1); ptx::mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) ptx::mbar_init(&nbar[i], 1); ptx::fence_barrier_init(); }
  if (warp == 0) {
    if (TWO_CTA) { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(ptx::smem_u32(&tptr)) : "memory");
                   asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
    else { ptx::tmem_alloc(&tptr, 512); ptx::tmem_relinquish(); }
  }
  if (shift_rows < 0) {  // random fp16 operands in [-2, 2) instead of whatever smem held (zeros after reset)
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    for (int i = threadIdx.x; i < 49152 + 32768; i += blockDim.x * 1) {
      x = x * 1664525u + 1013904223u;
      if ((i & 1) == 0) reinterpret_cast<__half*>(smem)[i >> 1] = __float2half(((int)(x >> 20) - 2048) / 1024.0f);
    }
    shift_rows = 0;
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (TWO_CTA) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  ptx::tc_fence_after();
  const uint32_t tb = tptr;
  const uint32_t a0 = ptx::smem_u32(smem) + shift_rows * 128, a1 = a0 + 24576, b0 = ptx::smem_u32(smem) + 49152;  // A_hi, A_lo, B (up to 256 rows = 32 KB)
  if (warp >= 2 && ldtm_noise) {
    // epilogue-like TMEM drains running concurrently with the MMAs (4 warps, 256 columns per round)
    const uint32_t tq = tb + ((uint32_t)((warp & 3) * 32) << 16);
    float accn = 0.f;
    while (!stop_flag) {
      for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(tq + c, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) accn += __uint_as_float(v[j]);
      }
      if (ldtm_noise > 1) __nanosleep(ldtm_noise);
    }
    if (accn == 1234.5f) out[2] = 1;
  }
  long long t0 = 0, t1 = 0;
  if (warp == 1 && (threadIdx.x & 31) == 0 && tma_noise) {
    // background TMA traffic into smem beyond the operands: 32 KB per round, 4 buffers
    uint8_t* nb = smem + 196608;
    // tma_noise = number of 32 KB loads kept in flight (1..4); runs until the MMA thread is done
    int i = 0;
    for (; !stop_flag && i < 1000000; ++i) {
      const int s = i % tma_noise;
      if (i >= tma_noise) ptx::mbar_wait(&nbar[s], ((i / tma_noise) - 1) & 1);
      ptx::mbar_arrive_expect_tx(&nbar[s], 32768);
      ptx::tma_load_2d(nb, &tm, &nbar[s], 0, ((i * 148 + blockIdx.x) * 256) % 65536);
    }
    for (int j = (i > tma_noise ? i - tma_noise : 0); j < i; ++j) ptx::mbar_wait(&nbar[j % tma_noise], (j / tma_noise) & 1);
    if (blockIdx.x == 0) out[1] = i;
  }
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t i1 = idesc_m(TWO_CTA ? 256 : 128, n1), i2 = idesc_m(TWO_CTA ? 256 : 128, n2);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (commit_every && it % commit_every == commit_every - 1) ptx::umma_commit(&cbar);
      // rotate: stream through `rotate` different 80 KB operand sets... (3 x 64 KB fits): defeats any operand reuse
      const uint32_t rot = rotate ? (uint32_t)(it % rotate) * 65536u : 0u;
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t da0 = ptx::make_smem_desc_sw128(a0 + rot + kk * 32), da1 = ptx::make_smem_desc_sw128(a0 + rot + 16384 + kk * 32);
        const uint64_t db = ptx::make_smem_desc_sw128(a0 + rot + 32768 + kk * 32);
        const uint32_t d2 = mode == 2 ? tb + n1 - n2 : tb + 256;  // mode 2: narrow accumulates into the wide's upper half
        if (TWO_CTA) { umma_f16_2cta(tb, da0, db, i1, 1); if (mode) umma_f16_2cta(d2, da1, db, i2, 1); }
        else { ptx::umma_f16(tb, da0, db, i1, 1); if (mode) ptx::umma_f16(d2, da1, db, i2, 1); }
      }
    }
    if (TWO_CTA) umma_commit_2cta(&bar); else ptx::umma_commit(&bar);
    while (!ptx::mbar_try_wait(&bar, 0)) {}
    t1 = clock64();
    stop_flag = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (TWO_CTA) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (warp == 0) {
    if (TWO_CTA) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
    else ptx::tmem_dealloc(tb, 512);
  }
}

// TMEM -> register bandwidth: 4 warps each read `cols` columns of their 32 lanes, `iters` times
__global__ void __launch_bounds__(128, 1) ldtm(int cols, int iters, long long* out, float* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&tptr, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tb = tptr + ((uint32_t)(warp * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < cols; c += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tb + c, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += __uint_as_float(v[j]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tptr, 512);
}

int main() {
  {
    long long* d; cudaMalloc(&d, 8); float* sink; cudaMalloc(&sink, 4);
    for (int cols : {128, 256, 512}) {
      ldtm<<<148, 128>>>(cols, 200, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("LDTM %d cols x 128 lanes: %.0f cycles per pass -> %.1f B/clk/SM (%s)\n", cols, (double)cyc / 200, cols * 128 * 4.0 * 200 / cyc, cudaGetErrorString(e));
    }
  }
  long long* d; cudaMalloc(&d, 32);
  void* buf; cudaMalloc(&buf, 65536ull * 128); cudaMemset(buf, 0, 65536ull * 128);
  // tensor map for the noise loads: [65536 rows][64 halves], box 64 x 256 rows = 32 KB
  CUtensorMap tm;
  {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*F)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    cuuint64_t dims[2] = {64, 65536}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 256}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((F)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d\n", (int)r); return 1; }
  }
  const int smem = 226 * 1024;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  struct Cfg { const char* name; int two, n1, n2, mode, noise, commit, ldtm, shift, rotate; } cfgs[] = {
      {"BN=128 mix, operands rotate over 3 x 64 KB", 0, 256, 128, 2, 0, 0, 0, 0, 3},
      {"BN=128 mix, rotate 3 + TMA noise", 0, 256, 128, 2, 2, 0, 0, 0, 3},
      {"BN=128 mix, commit every 2 iters (= per stage of 8 MMAs)", 0, 256, 128, 2, 0, 2, 0, 0},
      {"BN=128 mix, commit every iter", 0, 256, 128, 2, 0, 1, 0, 0},
      {"BN=128 mix, RANDOM operand data", 0, 256, 128, 2, 0, 0, 0, -1},
      {"BN=128 mix, RANDOM data + TMA noise", 0, 256, 128, 2, 3, 2, 0, -1},
      {"BN=128 mix, A start shifted 17 rows", 0, 256, 128, 2, 0, 0, 0, 17},
      {"BN=128 mix, A start shifted 31 rows", 0, 256, 128, 2, 0, 0, 0, 31},
      {"BN=128 mix + continuous LDTM from 4 warps", 0, 256, 128, 2, 0, 0, 1, 0},
      {"BN=128 mix + LDTM rounds with 1 us sleeps", 0, 256, 128, 2, 0, 0, 1000, 0},
      {"BN=128 mix + commit/stage + LDTM + TMA noise + shift", 0, 256, 128, 2, 3, 2, 1000, 17},
      {"1cta N=256", 0, 256, 0, 0, 0}, {"1cta N=128", 0, 128, 0, 0, 0}, {"1cta N=64", 0, 64, 0, 0, 0}, {"1cta N=32", 0, 32, 0, 0, 0},
      {"1cta N=256+128 (BN=128 split mix)", 0, 256, 128, 1, 0}, {"1cta N=128+64 (BN=64)", 0, 128, 64, 1, 0}, {"1cta N=64+32 (BN=32)", 0, 64, 32, 1, 0},
      {"1cta N=256+128 + TMA noise depth 1", 0, 256, 128, 2, 1},
      {"1cta N=256+128 + TMA noise depth 2", 0, 256, 128, 2, 2},
      {"1cta N=256+128 + TMA noise depth 3", 0, 256, 128, 2, 3},
      {"1cta N=256+128 + TMA noise depth 4", 0, 256, 128, 2, 4},
      {"1cta N=32 (slow MMA) + TMA noise depth 4", 0, 32, 0, 0, 4},
      {"1cta N=256 then N=128 INTO its upper half (current kernel)", 0, 256, 128, 2, 0},
      {"1cta N=128 then N=64 INTO its upper half", 0, 128, 64, 2, 0},
      {"1cta N=64 then N=32 INTO its upper half", 0, 64, 32, 2, 0},
      {"2cta M=256 N=256", 1, 256, 0, 0, 0}, {"2cta M=256 N=128", 1, 128, 0, 0, 0}, {"2cta M=256 N=256+128", 1, 256, 128, 1, 0}, {"2cta M=256 N=128+64", 1, 128, 64, 1, 0},
  };
  // sustained run: much longer, random data, wall-clock timed -> does the power cap show up as cycles or as clock?
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
      const int it2 = 400000;
      cudaMemset(d, 0, 32);
      cudaEventRecord(e0);
      k<0><<<148, 192, smem>>>(256, 128, 2, it2, 0, tm, d, 2, 0, -1, 3);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      long long res[2] = {0, 0}; cudaMemcpy(res, d, 16, cudaMemcpyDeviceToHost);
      const double flops = 148.0 * it2 * 4 * 128.0 * 384 * 16 * 2;
      printf("SUSTAINED BN=128 mix, random data, %d iters: %.1f cycles per K-step, %.1f ms, %.0f TFLOP/s issued, eff clock %.0f MHz (%s)\n", it2,
             (double)res[0] / (it2 * 4.0), ms, flops / (ms * 1e-3) / 1e12, res[0] / (ms * 1e-3) / 1e6, cudaGetErrorString(e));
    }
  }
  for (auto& c : cfgs) {
    cudaMemset(d, 0, 16);
    if (c.two) {
      cudaLaunchConfig_t lc = {}; lc.gridDim = dim3(148); lc.blockDim = dim3(192); lc.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      cudaLaunchKernelEx(&lc, k<1>, c.n1, c.n2, c.mode, iters, c.noise, tm, d, c.commit, c.ldtm, c.shift, c.rotate);
    } else {
      k<0><<<148, 192, smem>>>(c.n1, c.n2, c.mode, iters, c.noise, tm, d, c.commit, c.ldtm, c.shift, c.rotate);
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long res[2] = {0, 0}; cudaMemcpy(res, d, 16, cudaMemcpyDeviceToHost);
    const long long cyc = res[0];
    if (c.noise) printf("    noise: %lld x 32 KB in %lld cycles = %.1f B/clk/SM\n", res[1], cyc, res[1] * 32768.0 / cyc);
    const double per = (double)cyc / (iters * 4);
    const double macs = (c.two ? 256.0 : 128.0) * (c.n1 + (c.mode ? c.n2 : 0)) * 16;
    printf("%-42s %8.1f cycles per K-step  -> %7.0f MAC/clk/SM   (%s)\n", c.name, per, macs / per / (c.two ? 2 : 1), cudaGetErrorString(e));
  }
  return 0;
}
