// Micro-benchmark: L2 -> shared-memory delivery rate of TMA tile loads as a function of the row width of the box
// (128 rows of 128 B with SWIZZLE_128B vs 128 rows of 64 B with SWIZZLE_64B vs 32 B), all 148 SMs pulling, `depth`
// loads in flight per SM, source = an L2-resident matrix.  Answers: would half-width K chunks (64-byte rows) keep up?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu -lcuda && ./tma_rate
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../sbb_textline_detection_b200/csrc/ptx.cuh"
using namespace sbb;

__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap tm, int box_bytes, int depth, int iters, int rows_total,
                                           long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) ptx::mbar_init(&bar[i], 1); ptx::fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = i % depth;
      if (i >= depth) ptx::mbar_wait(&bar[s], ((i / depth) - 1) & 1);
      ptx::mbar_arrive_expect_tx(&bar[s], box_bytes);
      const int row = (int)(((long long)(i * 148 + blockIdx.x) * 128) % (rows_total - 128));
      ptx::tma_load_2d(smem + s * 16384, &tm, &bar[s], (i & 1) * 64, row);
    }
    for (int j = (iters > depth ? iters - depth : 0); j < iters; ++j) ptx::mbar_wait(&bar[j % depth], (j / depth) & 1);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

int main() {
  const int rows = 65536, cols = 256;  // 32 MB fp16 matrix: L2 resident
  __half* d; cudaMalloc(&d, (size_t)rows * cols * 2); cudaMemset(d, 0, (size_t)rows * cols * 2);
  long long* out; cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { int inner; CUtensorMapSwizzle sw; const char* name; } cfgs[] = {
      {64, CU_TENSOR_MAP_SWIZZLE_128B, "128 rows x 128 B (SW128)"}, {32, CU_TENSOR_MAP_SWIZZLE_64B, "128 rows x  64 B (SW64) "},
      {16, CU_TENSOR_MAP_SWIZZLE_32B, "128 rows x  32 B (SW32) "}};
  for (auto& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)c.inner, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int box_bytes = c.inner * 2 * 128;
    for (int depth : {4, 8, 12}) {
      const int iters = 4000;
      k<<<148, 64, 200 * 1024>>>(tm, box_bytes, depth, iters, rows, out);  // warm-up (fills L2)
      cudaDeviceSynchronize();
      k<<<148, 64, 200 * 1024>>>(tm, box_bytes, depth, iters, rows, out);
      cudaError_t e = cudaDeviceSynchronize();
      const double cyc = (double)out[0] / iters;
      printf("%s depth %2d: %7.1f cycles per box -> %6.1f B/clk/SM, %5.2f rows/clk/SM (%s)\n", c.name, depth, cyc, box_bytes / cyc, 128.0 / cyc,
             cudaGetErrorString(e));
    }
  }
  return 0;
}
