// Micro-benchmark: does sharing the weight (B) tile of a K chunk between the CTAs of a cluster through TMA
// multicast raise the operand delivery rate of the conv kernel's K loop?
//
// Every CTA models the producer of conv_gemm_tc_kernel<128, SPLIT> at half scale: per "half chunk" it needs 128 rows
// x 128 B that only it reads (the A tile of its own M tile) and 128 rows x 128 B that EVERY CTA reads at the same
// time (the B tile), into a ring of `depth` 32 KB slots (6 slots: throughput, not latency, is measured).  CL = 1:
// each CTA loads all 256 rows itself (what the kernel does today).  CL = 2 / 4: the CTAs of a cluster each load
// 128 / CL of the shared rows and
// multicast them to all CL CTAs (cp.async.bulk.tensor ... .multicast::cluster); a slot is re-filled only after
// every CTA of the cluster has released it (remote mbarrier arrive) -- the protocol the kernel would need.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_multicast tma_multicast.cu -lcuda && ./tma_multicast
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../sbb_textline_detection_b200/csrc/ptx.cuh"
using namespace sbb;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(ptx::smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

constexpr int kSlot = 32768;

// mode bit 0: load the private rows, bit 1: load the shared rows
template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(64, 1)
    k(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tmPart, int mode, int depth, int iters,
      int rows_total, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8];
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], CL); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (CL > 1) cluster_sync();  // peers' barriers are initialised before anyone multicasts into them
  if (threadIdx.x == 0) {
    const uint32_t bytes = ((mode & 1) ? 16384u : 0u) + ((mode & 2) ? 16384u : 0u);
    const int part_rows = 128 / CL;  // shared rows this CTA fetches per half chunk
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = i % depth;
      const uint32_t use = i / depth;
      if (i >= depth) ptx::mbar_wait(&empty[s], (use - 1) & 1);  // every CTA of the cluster has released the slot
      ptx::mbar_arrive_expect_tx(&full[s], bytes);
      uint8_t* slot = smem + s * kSlot;
      if (mode & 1) {  // private rows: distinct per CTA and iteration
        const int row = (int)(((long long)(i * 148 + blockIdx.x) * 128) % (rows_total / 2 - 128));
        ptx::tma_load_2d(slot, &tm128, &full[s], (i & 1) * 64, row);
      }
      if (mode & 2) {  // shared rows: the same for every CTA of the grid in iteration i
        const int row = (int)(((long long)i * 128) % 4096) + rows_total / 2;
        if (CL == 1) ptx::tma_load_2d(slot + 16384, &tm128, &full[s], 0, row);
        else
          tma_load_2d_mc(slot + 16384 + rank * part_rows * 128, &tmPart, &full[s], 0, row + rank * part_rows,
                         (uint16_t)((1u << CL) - 1));
      }
      // "consume": wait for the slot `depth - 1` chunks back, then release it in every CTA of the cluster
      if (i >= depth - 1) {
        const int j = i - (depth - 1), sj = j % depth;
        ptx::mbar_wait(&full[sj], (j / depth) & 1);
        if (CL == 1) ptx::mbar_arrive(&empty[sj]);
        else
          for (uint32_t r = 0; r < CL; ++r) mbar_arrive_remote(&empty[sj], r);
      }
    }
    for (int j = iters - (depth - 1); j < iters; ++j) {
      if (j < 0) continue;
      ptx::mbar_wait(&full[j % depth], (j / depth) & 1);
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (CL > 1) cluster_sync();  // nobody exits while a peer may still multicast into it / arrive on its barriers
}

static int encode(CUtensorMap* tm, void* d, int rows, int cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
  return r == CUDA_SUCCESS ? 0 : 1;
}

template <int CL>
static void run(const CUtensorMap& tm128, const CUtensorMap& tmPart, int rows, long long* out) {
  cudaFuncSetAttribute(k<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kSlot + 1024);
  const char* names[4] = {"", "private rows only  ", "shared rows only   ", "private + shared   "};
  for (int mode = 1; mode <= 3; ++mode) {
    const int depth = 6, iters = 6000;
    k<CL><<<148, 64, 6 * kSlot + 1024>>>(tm128, tmPart, mode, depth, iters, rows, out);  // warm-up (fills L2)
    cudaDeviceSynchronize();
    k<CL><<<148, 64, 6 * kSlot + 1024>>>(tm128, tmPart, mode, depth, iters, rows, out);
    cudaError_t e = cudaDeviceSynchronize();
    const double cyc = (double)out[0] / iters;
    const int rows_needed = ((mode & 1) ? 128 : 0) + ((mode & 2) ? 128 : 0);
    printf("cluster %d  %s: %7.1f cycles per half chunk -> %5.2f operand rows/clk/SM (%s)\n", CL, names[mode], cyc, rows_needed / cyc,
           cudaGetErrorString(e));
  }
}

int main() {
  const int rows = 65536, cols = 256;  // 32 MB fp16 matrix: L2 resident
  __half* d;
  cudaMalloc(&d, (size_t)rows * cols * 2);
  cudaMemset(d, 0, (size_t)rows * cols * 2);
  long long* out;
  cudaMallocManaged(&out, 64);
  CUtensorMap tm128, tm64part, tm32part;
  if (encode(&tm128, d, rows, cols, 128) || encode(&tm64part, d, rows, cols, 64) || encode(&tm32part, d, rows, cols, 32)) return 1;
  run<1>(tm128, tm128, rows, out);
  run<2>(tm128, tm64part, rows, out);
  run<4>(tm128, tm32part, rows, out);
  return 0;
}
