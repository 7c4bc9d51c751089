// Shared helpers for the repellency-projection kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>

#include "../../include/sdn_repel.h"

namespace sdn {

extern std::atomic<uint64_t> g_launches;

#define SDN_CUDA_OK(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return static_cast<int>(_e);     \
  } while (0)

// Count the launch and surface launch-time errors as a positive cudaError_t.
#define SDN_LAUNCHED()                                      \
  do {                                                      \
    ::sdn::g_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t _e = cudaGetLastError();                    \
    if (_e != cudaSuccess) return static_cast<int>(_e);     \
  } while (0)

// Optional per-kernel CUDA-event timing of the last sdn_repel_partial call (sdn_profile_enable).
struct ProfSlot { const char* name; cudaEvent_t e0, e1; };
struct Prof {
  bool enabled = false;
  int n = 0;
  ProfSlot slot[16] = {};
  void reset() { n = 0; }
  int begin(const char* name, cudaStream_t st) {
    if (!enabled || n >= 16) return -1;
    ProfSlot& s = slot[n];
    if (!s.e0) { cudaEventCreate(&s.e0); cudaEventCreate(&s.e1); }
    s.name = name;
    record(s.e0, st);
    return n++;
  }
  void end(int id, cudaStream_t st) { if (id >= 0) record(slot[id].e1, st); }
  // inside a stream capture the record must become an EXTERNAL event node, otherwise the event cannot be
  // synchronised / timed after the graph is replayed
  static void record(cudaEvent_t e, cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
      cudaEventRecordWithFlags(e, st, cudaEventRecordExternal);
    else
      cudaEventRecord(e, st);
  }
};
extern Prof g_prof;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

// Per-device one-time setup (function attributes and occupancy answers belong to a device, and one process may
// drive several): index static state by the current device.
constexpr int kMaxDevices = 64;
inline int device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in every thread.  `red` is >= 33 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Streaming 16-byte load that does not allocate in L1 (bank rows are read once per pass).
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// distance -> kernel weight, shared by every path (fast.py:249: exp(-cdist / (2 sigma^2))).
__device__ __forceinline__ float dist_from_dot(float xsq, float nsq, float dot, float alpha, int power) {
  float d2 = fmaf(-2.f * alpha, dot, fmaf(alpha * alpha, nsq, xsq));
  d2 = fmaxf(d2, 0.f);
  return power == 1 ? sqrtf(d2) : d2;
}

// Weight of a bank row for a query: the Gaussian kernel of the projection, or -- internal mode kPowerSparse, reached
// only through sdn_sparse_* -- the SPELL force weight relu(radius / d - 1) on the un-squared distance d < radius
// (fast.py:312-326); the radius then travels in the `inv2s2` argument.
constexpr int kPowerSparse = -1;
__device__ __forceinline__ float weight_from_dot(float xsq, float nsq, float dot, float alpha, int power, float inv2s2) {
  if (power == kPowerSparse) {
    const float d = dist_from_dot(xsq, nsq, dot, alpha, 1);
    return d < inv2s2 ? fmaxf(inv2s2 / d - 1.f, 0.f) : 0.f;
  }
  return expf(-dist_from_dot(xsq, nsq, dot, alpha, power) * inv2s2);
}

}  // namespace sdn
