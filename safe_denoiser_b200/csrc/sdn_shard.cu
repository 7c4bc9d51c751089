// N-sharded banks: ONE kernel that merges the per-rank partial sums over NVLink peer memory and applies the
// correction -- reduce-scatter, epilogue and all-gather fused (replaces NCCL all-reduce + local epilogue).
//
// Every rank r holds num_r [Q,D] | z_r [Q] (its shard's un-normalised sums) in a buffer its peers have mapped.
// Rank r owns the D-slice r of the result:
//     for d in slice r:  num = sum_p num_p[q][d]  (peer loads, fixed rank order => bit-identical on every rank)
//                        x0'[q][d] = x0[q][d] - scale * num / (sum_p z_p[q] + eps)
//                        store x0'[q][d] into EVERY rank's output buffer (peer stores)
// so each element crosses NVLink once in and once out -- a two-shot all-reduce with the epilogue in the middle.
// Cross-GPU ordering uses epoch flags in peer-mapped memory (release/acquire at system scope); every wait is
// bounded by a wall-clock timeout that raises an error flag (no trap).  One kernel per rank, each on its own GPU.
#include <algorithm>
#include <cstdlib>

#include "sdn_internal.h"

namespace sdn {

constexpr int kMaxRanks = 8;

struct MergeArgs {
  const float* packed[kMaxRanks];   // rank p's [Q*D | Q] partial buffer, as mapped here
  float* out[kMaxRanks];            // rank p's [Q*D] result buffer
  uint32_t* sig[kMaxRanks];         // rank p's flag words [2 * kMaxRanks]
  int rank, world;
  uint32_t* epoch_ptr;              // local device word: epoch of THIS launch; the kernel advances it
  int64_t Q, D;
  float eps, scale, gate_thr;
  int flags;
  const float* x0;                  // local replicated query [Q,D]
  float* denom_out;                 // local [Q]
  int32_t* gate_out;                // local [Q]
  unsigned int* counter;            // local [4]: [0] block counter (zero between launches), [1] sticky error flag
  unsigned long long timeout_ns;    // how long a rank waits for a peer before it gives up (error flag, no trap)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// epochs only grow; compare with wrap-around in mind.  A peer whose host is late (image save, GC, torch.load)
// only delays this rank; after `timeout_ns` (default 60 s, SDN_SHARD_TIMEOUT_S) the rank stops waiting, raises
// the sticky error flag and lets the kernel finish -- a trap would kill the CUDA context of a healthy rank.
__device__ __forceinline__ void spin_until(const uint32_t* p, uint32_t epoch, unsigned long long timeout_ns,
                                           unsigned int* err_flag) {
  unsigned long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    if ((int32_t)(ld_acquire_sys(p) - epoch) >= 0) return;
    if (spin > 64u) __nanosleep(spin > 4096u ? 1000 : 64);
    if ((spin & 1023u) == 1023u) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) { atomicExch(err_flag, 1u); return; }
    }
  }
}

__global__ void __launch_bounds__(256) k_shard_merge_correct(const MergeArgs a) {
  __shared__ float s_denom;
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int64_t q = blockIdx.y;
  // The epoch lives in device memory so that the launch can be replayed from a CUDA graph: every block reads it
  // on entry, the last block to finish advances it (all blocks have read it by then).
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(a.epoch_ptr);

  // ---- barrier 1: every rank's partial sums are complete (they were written by earlier kernels of its stream)
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid < a.world) {
    __threadfence_system();
    st_release_sys(a.sig[tid] + a.rank, epoch);
  }
  if (tid < a.world) spin_until(a.sig[a.rank] + tid, epoch, a.timeout_ns, a.counter + 1);
  __syncthreads();

  // ---- my D-slice, in float4 units.  The peer loads of the first (usually only) iteration are issued BEFORE the
  // ---- z loads and the block barrier below, so that one NVLink round trip (~2.5 us) covers both.
  const int64_t v_all = a.D / 4;
  const int64_t v0 = v_all * a.rank / a.world, v1 = v_all * (a.rank + 1) / a.world;
  int64_t v = v0 + (int64_t)blockIdx.x * 256 + tid;
  float4 t[kMaxRanks];
  auto load_peers = [&](int64_t vv) {
    const int64_t o = q * a.D + vv * 4;
    // all peer loads in flight at once (an NVLink load is ~2 us; issued one after the other they dominated the
    // kernel), summed in rank order below
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p)
      t[p] = (p < a.world) ? __ldcv(reinterpret_cast<const float4*>(a.packed[p] + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  if (v < v1) load_peers(v);
  if (tid == 0) {
    float zp[kMaxRanks];
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p) zp[p] = (p < a.world) ? __ldcv(a.packed[p] + a.Q * a.D + q) : 0.f;
    float z = 0.f;
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p) z += zp[p];
    const float denom = z + a.eps;
    s_denom = denom;
    if (blockIdx.x == 0) {
      a.denom_out[q] = denom;
      a.gate_out[q] = (!(a.flags & SDN_EPI_GATE) || denom > a.gate_thr) ? 1 : 0;
    }
  }
  __syncthreads();
  const float denom = s_denom;
  while (v < v1) {
    const int64_t o = q * a.D + v * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p) { s.x += t[p].x; s.y += t[p].y; s.z += t[p].z; s.w += t[p].w; }
    const float4 x = *reinterpret_cast<const float4*>(a.x0 + o);
    float4 r;
    r.x = fmaf(-a.scale, s.x / denom, x.x);
    r.y = fmaf(-a.scale, s.y / denom, x.y);
    r.z = fmaf(-a.scale, s.z / denom, x.z);
    r.w = fmaf(-a.scale, s.w / denom, x.w);
#pragma unroll
    for (int p = 0; p < kMaxRanks; ++p)
      if (p < a.world) *reinterpret_cast<float4*>(a.out[p] + o) = r;
    v += (int64_t)gridDim.x * 256;
    if (v < v1) load_peers(v);
  }

  // ---- barrier 2: my stores have landed everywhere; leave only when every peer's stores have landed here
  __threadfence_system();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.counter, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (s_last) {
    if (tid < a.world) {
      __threadfence_system();
      st_release_sys(a.sig[tid] + kMaxRanks + a.rank, epoch);
      spin_until(a.sig[a.rank] + kMaxRanks + tid, epoch, a.timeout_ns, a.counter + 1);
    }
    __syncthreads();
    if (tid == 0) {
      *a.counter = 0;
      *a.epoch_ptr = epoch + 1;
    }
  }
}

}  // namespace sdn

using namespace sdn;

extern "C" int sdn_shard_merge_correct(const void* const* peer_packed, void* const* peer_out, void* const* peer_sig,
                                       int32_t rank, int32_t world, void* epoch_word, int64_t Q, int64_t D, float eps,
                                       float scale, float gate_threshold, int32_t flags, const float* x0_local,
                                       float* denom_out, int32_t* gate_out, void* counter, void* stream) {
  if (!peer_packed || !peer_out || !peer_sig || !x0_local || !denom_out || !gate_out || !counter || !epoch_word)
    return SDN_E_NULL;
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return SDN_E_PARAM;
  if (Q <= 0 || D <= 0 || Q > 65535) return SDN_E_SHAPE;
  if (D % 4 != 0 || !aligned16(x0_local)) return SDN_E_ALIGN;
  MergeArgs a{};
  for (int p = 0; p < world; ++p) {
    if (!peer_packed[p] || !peer_out[p] || !peer_sig[p]) return SDN_E_NULL;
    if (!aligned16(peer_packed[p]) || !aligned16(peer_out[p])) return SDN_E_ALIGN;
    a.packed[p] = static_cast<const float*>(peer_packed[p]);
    a.out[p] = static_cast<float*>(peer_out[p]);
    a.sig[p] = static_cast<uint32_t*>(peer_sig[p]);
  }
  a.rank = rank; a.world = world; a.epoch_ptr = static_cast<uint32_t*>(epoch_word); a.Q = Q; a.D = D; a.eps = eps; a.scale = scale;
  a.gate_thr = gate_threshold; a.flags = flags; a.x0 = x0_local; a.denom_out = denom_out; a.gate_out = gate_out;
  a.counter = static_cast<unsigned int*>(counter);
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("SDN_SHARD_TIMEOUT_S");
    const double sec = e ? atof(e) : 60.0;
    return (unsigned long long)((sec > 0.0 ? sec : 60.0) * 1e9);
  }();
  a.timeout_ns = timeout_ns;
  const int64_t slice_v = D / 4 / world + 1;
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv(slice_v, 256), 1024));
  const int pid = g_prof.begin("k_shard_merge_correct", (cudaStream_t)stream);
  k_shard_merge_correct<<<dim3(gx, (unsigned)Q), 256, 0, (cudaStream_t)stream>>>(a);
  g_prof.end(pid, (cudaStream_t)stream);
  SDN_LAUNCHED();
  return SDN_OK;
}
