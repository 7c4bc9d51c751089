// Micro-benchmark 2: a replica of phase B's load side (sdn_umma.cu k_umma_accum at cfg3), nothing but loads, with the
// features of the real kernel added one at a time to find what costs 14 us (load-only pass over 201 MB: 35 us with a
// clean L2, tma_map.cu; the real phase B: 49.5 us).  128 CTAs, CTA = d-block, 48 stages of 64 bank rows:
//   per stage 2 boxes [64 rows][64 bf16] from the hi plane + 2 from the lo plane (plane_rows below it) = 32 KiB
//   +P       two more boxes [64 rows][64] of a small shared array (the weights: the same for every CTA)
//   +delay   the slot is re-armed `delay` ns after its data arrived (MMAs + commit + producer wake-up)
//   +spin    8 more warps wait on an mbarrier for the whole kernel (the epilogue warps)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_map2 tma_map2.cu -lcuda && ./tma_map2
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

constexpr int kStages = 4;
constexpr uint32_t kStage = 49152;       // 32 KiB bank + 16 KiB weights slot
constexpr int64_t D = 16384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void box(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(320, 1) k(const __grid_constant__ CUtensorMap mb, const __grid_constant__ CUtensorMap mp,
                                             int nblocks, int plane_rows, int with_p, int delay_ns, int spin, int nstages,
                                             unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], endbar;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&endbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp >= 2) {
    if (spin) wait(&endbar, 0);
    return;
  }
  if (threadIdx.x != 0) return;
  const int d0 = blockIdx.x * 128;
  auto issue = [&](int i) {
    const int s = i % nstages;
    const uint32_t bar = smem_u32(&full[s]), dst = smem_u32(smem + (size_t)s * kStage);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(with_p ? 49152u : 32768u) : "memory");
    const int r = i * 64;
    if (with_p) { box(dst + 32768, &mp, bar, 0, r); box(dst + 40960, &mp, bar, 64, r); }
    box(dst, &mb, bar, d0, r); box(dst + 8192, &mb, bar, d0 + 64, r);
    box(dst + 16384, &mb, bar, d0, plane_rows + r); box(dst + 24576, &mb, bar, d0 + 64, plane_rows + r);
  };
  for (int i = 0; i < nstages && i < nblocks; ++i) issue(i);
  for (int i = 0; i < nblocks; ++i) {
    wait(&full[i % nstages], (uint32_t)((i / nstages) & 1));
    if (delay_ns) {
      uint64_t t0, t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < (uint64_t)delay_ns);
    }
    if (i + nstages < nblocks) issue(i + nstages);
  }
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&endbar)) : "memory");
  sink[blockIdx.x] = smem[0];
}

__global__ void k_flush(const uint4* p, size_t n, unsigned* sink) {
  unsigned a = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a += p[i].x;
  if (a == 0x12345678u) sink[1000] = a;
}

int main() {
  const int64_t N = 8192;
  uint8_t* buf; unsigned* sink; uint8_t* flush; uint8_t* pbuf;
  cudaMalloc(&buf, (size_t)N * D * 2); cudaMemset(buf, 1, (size_t)N * D * 2);
  cudaMalloc(&pbuf, 4096 * 128 * 2); cudaMemset(pbuf, 1, 4096 * 128 * 2);
  cudaMalloc(&flush, 512u << 20); cudaMemset(flush, 3, 512u << 20);
  cudaMalloc(&sink, 8192);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  CUtensorMap mb, mp;
  cuuint32_t es[2] = {1, 1}, bx[2] = {64, 64};
  {
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)N}; cuuint64_t strides[1] = {(cuuint64_t)D * 2};
    enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t pd[2] = {128, 4096}; cuuint64_t ps[1] = {256};
    enc(&mp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pbuf, pd, ps, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kStage);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, int grid, int nblocks, int plane_rows, int with_p, int delay, int spin, int nstages) {
    float best = 1e30f, sum = 0;
    for (int rep = 0; rep < 4; ++rep) {
      k_flush<<<148 * 8, 256>>>(reinterpret_cast<const uint4*>(flush), (512u << 20) / 16, sink);
      cudaEventRecord(e0);
      k<<<grid, 320, 4 * kStage>>>(mb, mp, nblocks, plane_rows, with_p, delay, spin, nstages, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) { sum += ms; if (ms < best) best = ms; }
    }
    const double bytes = (double)grid * nblocks * 32768;
    printf("%-66s %6.0f MB: best %6.1f us, mean %6.1f us  %7.1f GB/s\n", name, bytes / 1e6, best * 1e3, sum / 3 * 1e3, bytes / best / 1e6);
  };
  run("hi + lo planes 3000 rows apart, 48 blocks", 128, 48, 3000, 0, 0, 0, 4);
  run("hi + lo planes 3072 rows apart", 128, 48, 3072, 0, 0, 0, 4);
  run("hi + lo planes 4096 rows apart", 128, 48, 4096, 0, 0, 0, 4);
  run("  3 stages", 128, 48, 3000, 0, 0, 0, 3);
  run("+ P tiles", 128, 48, 3000, 1, 0, 0, 4);
  run("+ 8 waiting warps", 128, 48, 3000, 0, 0, 1, 4);
  run("+ slot re-armed 200 ns after arrival", 128, 48, 3000, 0, 200, 0, 4);
  run("+ slot re-armed 400 ns after arrival", 128, 48, 3000, 0, 400, 0, 4);
  run("+ slot re-armed 800 ns after arrival", 128, 48, 3000, 0, 800, 0, 4);
  run("+ P + waiting warps + 400 ns", 128, 48, 3000, 1, 400, 1, 4);
  run("all 148 SMs, 41 blocks each (same bytes), plain", 148, 41, 3000, 0, 0, 0, 4);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
