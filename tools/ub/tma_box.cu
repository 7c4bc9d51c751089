// Micro-benchmark: HBM read rate of the access patterns of the tcgen05 kernels, nothing but the loads.
//   A  linear     : cp.async.bulk of contiguous 16 KiB pieces (what a TILED plane layout would allow)
//   B  box 128x64 : TMA 2-D tile [128 rows][64 bf16] of a row-major [N][D] bf16 plane = 128 segments of 128 bytes,
//                   32 KiB apart (phase A of sdn_umma.cu / sdn_flash.cu today)
//   C  box 64x64  : [64 rows][64 bf16], two per stage (phase B today)
// Ring of 12 x 16 KiB stages per CTA (192 KiB in flight), one CTA per SM, no compute.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box tma_box.cu -lcuda && ./tma_box
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

constexpr int kStages = 12;
constexpr uint32_t kTile = 16384;
constexpr int64_t D = 16384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mode 0: linear 16 KiB; 1: box [128][64]; 2: two boxes [64][64] (rows r and d-blocks as phase B walks them)
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap m128, const __grid_constant__ CUtensorMap m64,
                                             const uint8_t* buf, int mode, int64_t nrows, int tiles_per_cta, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int64_t row_tiles = nrows / 128, kblocks = D / 64;
  auto issue = [&](int i) {
    const int s = i % kStages;
    const uint32_t bar = smem_u32(&full[s]), dst = smem_u32(smem + (size_t)s * kTile);
    // tile index: CTA b walks tiles b, b + grid, ... in (row tile, k block) order with k fastest -- like phase A's K loop
    const int64_t t = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) % (row_tiles * kblocks);
    const int64_t rt = t / kblocks, kb = t % kblocks;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTile) : "memory");
    if (mode == 0) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(buf + (size_t)t * kTile), "r"(kTile), "r"(bar) : "memory");
    } else if (mode == 1) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(dst), "l"(&m128), "r"(bar), "r"((int)(kb * 64)), "r"((int)(rt * 128)) : "memory");
    } else {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(dst), "l"(&m64), "r"(bar), "r"((int)(kb * 64)), "r"((int)(rt * 128)) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(dst + 8192), "l"(&m64), "r"(bar), "r"((int)(kb * 64)), "r"((int)(rt * 128 + 64)) : "memory");
    }
  };
  for (int i = 0; i < kStages && i < tiles_per_cta; ++i) issue(i);
  for (int i = 0; i < tiles_per_cta; ++i) {
    const int s = i % kStages;
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&full[s])), "r"((uint32_t)((i / kStages) & 1)) : "memory");
    if (i + kStages < tiles_per_cta) issue(i + kStages);
  }
  sink[blockIdx.x] = smem[0];
}

int main() {
  const int64_t N = 32768;                                     // 1 GiB plane
  uint8_t* buf; unsigned* sink;
  cudaMalloc(&buf, (size_t)N * D * 2); cudaMemset(buf, 1, (size_t)N * D * 2);
  cudaMalloc(&sink, 4096);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  CUtensorMap m128, m64;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)N}; cuuint64_t strides[1] = {(cuuint64_t)D * 2}; cuuint32_t es[2] = {1, 1};
  cuuint32_t b128[2] = {64, 128}, b64[2] = {64, 64};
  enc(&m128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, b128, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&m64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, b64, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kTile);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"linear 16 KiB", "box [128 rows][64 bf16]", "2 x box [64 rows][64 bf16]"};
  // second part: the same patterns over a span that stays in the L2 (1024 rows = 32 MiB), few CTAs: what ONE SM's TMA
  // unit delivers per pattern when neither HBM nor the L2 is the limit
  for (int mode = 0; mode < 3; ++mode)
    for (int g : {148, 37, 8, 1}) {
      const int tiles = 4096;
      k<<<g, 128, kStages * kTile>>>(m128, m64, buf, mode, 1024, tiles, sink);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      k<<<g, 128, kStages * kTile>>>(m128, m64, buf, mode, 1024, tiles, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = (double)g * tiles * kTile;
      printf("L2-resident  %-28s %3d CTAs: %8.1f GB/s  (%5.1f GB/s per SM)\n", names[mode], g, bytes / ms / 1e6, bytes / ms / 1e6 / g);
    }
  for (int mode = 0; mode < 3; ++mode)
    for (int g : {148, 128}) {
      const int tiles = 4096;                                  // 64 MiB per CTA
      k<<<g, 128, kStages * kTile>>>(m128, m64, buf, mode, N, 256, sink);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      k<<<g, 128, kStages * kTile>>>(m128, m64, buf, mode, N, tiles, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = (double)g * tiles * kTile;
      printf("%-28s %3d CTAs: %8.1f GB/s  (%5.1f GB/s per SM)\n", names[mode], g, bytes / ms / 1e6, bytes / ms / 1e6 / g);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
