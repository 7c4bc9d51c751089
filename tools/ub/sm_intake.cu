// Micro-benchmark: how many bytes per second can ONE SM take in through cp.async.bulk, from L2 and from HBM?
// Every CTA (1 per SM) streams 32 KiB chunks into a ring of 6 shared-memory stages and does nothing with them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sm_intake sm_intake.cu && ./sm_intake
// The answer decides what bounds kernels that read a bank tile AND an operand tile per stage.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kStages = 6;
constexpr uint32_t kChunk = 32768;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const uint8_t* buf, size_t span_bytes, int chunks_per_cta, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t nchunks_span = span_bytes / kChunk;
  auto issue = [&](int i) {
    const int s = i % kStages;
    // CTA b reads chunks b, b + grid, ... of the span (wrapping): every chunk of the span is read by some CTA
    const size_t c = ((size_t)blockIdx.x + (size_t)i * gridDim.x) % nchunks_span;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(kChunk) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem + (size_t)s * kChunk)), "l"(buf + c * kChunk), "r"(kChunk), "r"(smem_u32(&full[s])) : "memory");
  };
  for (int i = 0; i < kStages && i < chunks_per_cta; ++i) issue(i);
  for (int i = 0; i < chunks_per_cta; ++i) {
    const int s = i % kStages;
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&full[s])), "r"((uint32_t)((i / kStages) & 1)) : "memory");
    if (i + kStages < chunks_per_cta) issue(i + kStages);
  }
  sink[blockIdx.x] = smem[0];
}

int main() {
  const size_t big = (size_t)2 << 30;
  uint8_t* buf; unsigned* sink;
  cudaMalloc(&buf, big); cudaMemset(buf, 1, big);
  cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kChunk);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grids[] = {148, 128, 74, 37, 8, 1};
  const size_t spans[] = {(size_t)32 << 20, big};           // 32 MiB: stays in the L2 after a warm-up pass; 2 GiB: HBM
  for (size_t span : spans) {
    for (int g : grids) {
      const int chunks = 2048;                               // 64 MiB per CTA
      k<<<g, 128, kStages * kChunk>>>(buf, span, chunks, sink);   // warm-up (fills the L2 for the small span)
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      k<<<g, 128, kStages * kChunk>>>(buf, span, chunks, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = (double)g * chunks * kChunk;
      printf("span %5zu MiB  %3d CTAs (1 per SM, %d x 32 KiB in flight each): %8.1f GB/s total  %6.1f GB/s per SM\n",
             span >> 20, g, kStages, bytes / ms / 1e6, bytes / ms / 1e6 / g);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
