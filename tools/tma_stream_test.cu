// Development micro-benchmark: how fast does ONE thread per SM stream NHWC halo tiles into shared memory with
// cp.async.bulk.tensor.4d (the loader of conv3x3_tc_kernel), nothing consuming the data?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tests/tma_stream tools/tma_stream_test.cu -lcuda
//   build/tests/tma_stream [n_images] [box_rows] [swizzle(0|3)] [stages]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { auto e_ = (x); if (e_ != 0) { printf("error %d at %s:%d\n", (int)e_, __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  uint32_t done = 0;
  while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(b), "r"(ph) : "memory");
}

__global__ void __launch_bounds__(128, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, int tiles_x, int tiles_per_img, int total, int tr, int tc, int box_rows, int stages,
              int stage_bytes, int loads_per_tile) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + stages * stage_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int my = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto issue = [&](int tl) {
    const int tile = blockIdx.x + tl * gridDim.x;
    const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
    const int y0 = (rem / tiles_x) * tr - 1, x0 = (rem % tiles_x) * tc - 1;
    const int s = tl % stages;
    mbar_expect(bars + 8 * s, (uint32_t)stage_bytes);
    const int rows_per_load = (tr + 2) / loads_per_tile;
    for (int l = 0; l < loads_per_tile; ++l)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
                   "r"(base + s * stage_bytes + l * rows_per_load * 32 * 128), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(0), "r"(x0),
                   "r"(y0 + l * rows_per_load), "r"(n), "r"(bars + 8 * s) : "memory");
  };
  for (int tl = 0; tl < my && tl < stages; ++tl) issue(tl);
  for (int tl = 0; tl < my; ++tl) {
    mbar_wait(bars + 8 * (tl % stages), (tl / stages) & 1);
    if (tl + stages < my) issue(tl + stages);
  }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 8, loads_per_tile = argc > 2 ? atoi(argv[2]) : 1;
  const int swz = argc > 3 ? atoi(argv[3]) : 3, stages = argc > 4 ? atoi(argv[4]) : 6;
  const int H = 272, W = 480, C = 64, TR = 4, TC = 30;
  void* x;
  CK(cudaMalloc(&x, (size_t)n * H * W * C * 2));
  CK(cudaMemset(x, 1, (size_t)n * H * W * C * 2));
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, 32, (cuuint32_t)((TR + 2) / loads_per_tile), 1}, es[4] = {1, 1, 1, 1};
  CK(cuInit(0));
  CK(cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  const int tiles_x = (W + TC - 1) / TC, tiles_y = (H + TR - 1) / TR, tpi = tiles_x * tiles_y, total = tpi * n;
  const int stage_bytes = (TR + 2) * 32 * 128 + 1024;     // 25 600
  const int smem = stages * stage_bytes + 64 + 1024;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 3; ++it) stream_kernel<<<148, 128, smem>>>(tm, tiles_x, tpi, total, TR, TC, 0, stages, stage_bytes - 1024, loads_per_tile);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  const int iters = 10;
  for (int it = 0; it < iters; ++it) stream_kernel<<<148, 128, smem>>>(tm, tiles_x, tpi, total, TR, TC, 0, stages, stage_bytes - 1024, loads_per_tile);
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double us = ms * 1e3 / iters, tile_bytes = (TR + 2) * 32 * 128.0;
  printf("n=%d loads/tile=%d swizzle=%d stages=%d: %.1f us per pass, %.2f us per tile and SM, %.0f GB/s of boxes (%.0f GB/s of tensor)\n", n,
         loads_per_tile, swz, stages, us, us / ((double)total / 148), total * tile_bytes / us * 1e-3, (double)n * H * W * C * 2 / us * 1e-3);
  return 0;
}
