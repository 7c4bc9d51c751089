// Micro-benchmark: do 16 independent 16-byte loads of L2-resident lines overlap, per load flavour?  (one warp)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ld_modes ld_modes.cu && ./ld_modes
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ uint4 ld(const uint4* p) {
  uint4 v;
  if (MODE == 0) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == 1) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == 2) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == 3) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == 4) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == 5) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == 6) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

template <int MODE, int NLD>
__global__ void k(const uint4* buf, size_t stride16, long long* out, unsigned* sink, int rounds) {
  const int lane = threadIdx.x;
  unsigned acc = 0;
  long long best = 1ll << 60, total = 0;
  for (int r = 0; r < rounds; ++r) {
    const uint4* p = buf + (size_t)(r * NLD) * stride16 + lane * 2;
    __syncwarp();
    const long long c0 = clock64();
    uint4 v[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) v[i] = ld<MODE>(p + (size_t)i * stride16);
#pragma unroll
    for (int i = 0; i < NLD; ++i) acc += v[i].x + v[i].w;
    // make the timer depend on the data
    const long long c1 = clock64() + (acc == 0x12345678u ? 1 : 0);
    const long long d = c1 - c0;
    if (r > 0) { best = d < best ? d : best; total += d; }
  }
  if (lane == 0) { out[0] = best; out[1] = total / (rounds - 1); }
  sink[threadIdx.x] = acc;
}

template <int MODE, int NLD>
void run(const char* name, const uint4* buf, size_t stride16, long long* out, unsigned* sink) {
  const int rounds = 64;
  k<MODE, NLD><<<1, 32>>>(buf, stride16, out, sink, rounds);   // warm: brings the lines into L2
  cudaDeviceSynchronize();
  k<MODE, NLD><<<1, 32>>>(buf, stride16, out, sink, rounds);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s %2d loads: best %6lld clk  mean %6lld clk  (%5.0f clk per load)\n", name, NLD, h[0], h[1], (double)h[1] / NLD);
}

int main() {
  const size_t stride = 32768;                 // bytes between the lines of one round (like the cluster partials)
  const size_t bytes = stride * 16 * 64 + 4096;
  uint4* buf; long long* out; unsigned* sink;
  cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
  cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
#define RUN(M, NAME) run<M, 1>(NAME, buf, stride / 16, out, sink); run<M, 16>(NAME, buf, stride / 16, out, sink);
  RUN(0, "ld.global.cg")
  RUN(1, "ld.relaxed.gpu.global")
  RUN(2, "ld.volatile.global")
  RUN(3, "ld.global (weak, .ca)")
  RUN(4, "ld.global.L1::no_allocate")
  RUN(5, "ld.global.nc.L1::no_allocate")
  RUN(6, "ld.global.cv")
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
