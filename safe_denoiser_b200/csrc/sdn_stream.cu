// One-pass "flash-repellency" kernel for GEMV-shaped calls (Q <= 8): the bank is read from HBM ONCE.
//
// A thread-block cluster of CS CTAs owns a contiguous range of bank rows; CTA r of the cluster owns the
// D-slice [r*SLICE, (r+1)*SLICE) of every row (SLICE = D / CS floats).  Rows stream through a ring of
// shared-memory stages filled by TMA bulk copies (cp.async.bulk + mbarrier complete_tx).  Per tile of TN rows:
//
//   phase 1   partial dots x_q[slice] . n_i[slice] from smem -> warp reduce-scatter -> CTA partial
//             -> st.async into every peer CTA's receive slot through distributed shared memory; each
//                store completes bytes on the RECEIVER's mbarrier (no cluster barrier, no fence)
//   phase 2   (for the PREVIOUS tile, whose partials have all arrived)  full dot -> distance -> k = exp(.)
//             -> acc[q][slice] += k * n_i[slice]   re-reading the SAME smem tile: no second HBM pass
//
// The accumulators (Q x SLICE per CTA) and the query slice live in registers.  Each cluster writes one
// partial [Q, D] (+ Z[Q]); a small second kernel sums the per-cluster partials (deterministic).
//
// Replaces repellency_methods_fast.py:249-257 for the shapes the reference actually runs (Q = 1).
#include <algorithm>
#include <cstdlib>

#include "sdn_internal.h"

namespace sdn {

constexpr int kStCompute = 512;               // compute threads
constexpr int kStCWarps = kStCompute / 32;    // 16 compute warps
constexpr int kStThreads = kStCompute + 64;   // + TMA producer warp + weights warp
constexpr int kStMaxCluster = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a broken pipeline traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// Store one float into CTA `peer`'s copy of `local_addr` and complete 4 bytes on that CTA's copy of `local_bar`.
__device__ __forceinline__ void st_async_remote(float* local_addr, uint64_t* local_bar, uint32_t peer, float v) {
  uint32_t raddr, rbar;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_addr)), "r"(peer));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(local_bar)), "r"(peer));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(raddr), "r"(__float_as_uint(v)), "r"(rbar) : "memory");
}

__device__ __forceinline__ void st_remote(float* local_addr, uint32_t peer, float v) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_addr)), "r"(peer));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

// Sum V per-lane values across the warp with a reduce-scatter butterfly (V-1 shuffles instead of 5V).
// On return, lane l holds in v[0] the warp total of value index `warp_value_index<V>(l)`.
template <int V>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[V], int lane) {
  static_assert(V == 1 || V == 2 || V == 4 || V == 8 || V == 16 || V == 32, "V must be a power of two <= 32");
  int off = 16;
#pragma unroll
  for (int n = V; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
    const int h = n >> 1;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = upper ? v[i] : v[i + h];
      const float keep = upper ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}
// Which value index lane `lane` ends up holding: step j (offset 16 >> j) picks the upper half when the
// lane bit is set, so the index accumulates bit (log2V-1-j) from lane bit (4-j).
template <int V>
__device__ __forceinline__ int warp_value_index(int lane) {
  int idx = 0, off = 16;
#pragma unroll
  for (int n = V; n > 1; n >>= 1, off >>= 1)
    if (lane & off) idx += n >> 1;
  return idx;
}
// Lanes that hold a distinct value: those whose low (5 - log2 V) bits are zero.
template <int V>
__device__ __forceinline__ bool warp_value_owner(int lane) {
  int bits = 0;
  for (int n = V; n > 1; n >>= 1) ++bits;
  return (lane & ((1 << (5 - bits)) - 1)) == 0;
}

struct StreamArgs {
  const float* bank; const float* sqnorm; int64_t N; int64_t D;
  const float* x; const float* xsq; int q_real;   // xsq == nullptr: ||x||^2 is computed in the kernel
  float inv2s2; int power; float alpha;
  float* part_num;   // [clusters][q_real][D]
  float* part_z;     // [clusters][q_real]
  float* k_out;      // [q_real][N] or null
  float* zero_word;  // optional scalar cleared by CTA 0 (the logging mean of the fused epilogue), or null
};

template <int Q, int VPT, int TN, int kStStages, int LAG>
constexpr size_t stream_smem_bytes() {
  constexpr size_t slice = (size_t)VPT * kStCompute * 4;
  constexpr size_t V = (size_t)TN * Q;
  constexpr size_t SL = LAG + 2;             // slots of every hand-off between phase 1 and phase 2 (>= 3)
  return kStStages * TN * slice * 4          // tile ring
         + SL * kStCWarps * V * 4            // per-warp partials
         + SL * kStMaxCluster * V * 4        // receive slots
         + SL * V * 4                        // k of the tiles being accumulated
         + 64 * 4                            // z reduction scratch
         + kStMaxCluster * 8 * 4             // per-CTA ||x slice||^2 partials
         + (2 * kStStages + 3 * SL) * 8 + 64;      // mbarriers + alignment slack
}

// Warp roles: warps [0, 16) compute, warp 16 lane 0 issues the TMA bulk copies, warp 17 lanes [0, V) turn
// partial dots into weights.  All hand-offs are mbarriers; there is no __syncthreads in the loop.
// kStStages x TN rows x slice = 192 KiB of tiles in every instance.  Round 2: the DSMEM exchange is one ~1.2 us round
// trip per TILE and only one tile of phase 1 overlaps it, so tiles of one or two rows (TN = 1, 2 with six stages) made
// the kernel latency-bound -- 0.49 of the HBM roofline at Q = 1, N = 3000 (a 32 KiB tile per CTA lasts 0.7 us at full
// rate).  Three stages of TN = 4 rows x 16 KiB (VPT = 2, clusters of 4) carry 64 KiB per exchange instead.
// LAG = tiles between phase 1 and phase 2 of a tile: the exchange of tile t overlaps phase 1 of the next LAG tiles (and
// phase 2 of the previous ones); every hand-off has LAG + 2 slots (a peer or the compute warps may be that far ahead).
template <int Q, int VPT, int TN, int kStStages, int LAG>
__global__ void __launch_bounds__(kStThreads, 1) k_stream(const StreamArgs a) {
  constexpr int V = TN * Q;
  constexpr int SL = LAG + 2;
  static_assert(LAG >= 1 && LAG < kStStages, "phase 2 must free stages before the ring wraps");
  constexpr int SLICE = VPT * kStCompute * 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);                 // [stages][TN][SLICE]
  float* wp = tiles + (size_t)kStStages * TN * SLICE;                 // [SL][cwarps][V]
  float* recv = wp + SL * kStCWarps * V;                              // [SL][kStMaxCluster][V]
  float* ks = recv + SL * kStMaxCluster * V;                          // [SL][V]
  float* zred = ks + SL * V;                                          // [64]
  float* xsp = zred + 64;                                             // [kStMaxCluster][8]
  uint64_t* full = reinterpret_cast<uint64_t*>(xsp + kStMaxCluster * 8);   // [stages]  TMA landed
  uint64_t* empty = full + kStStages;                                 // [stages]  all compute warps done with the stage
  uint64_t* pr = empty + kStStages;                                   // [SL]      warp partials of a tile written
  uint64_t* kr = pr + SL;                                             // [SL]      weights of a tile written
  uint64_t* rbar = kr + SL;                                           // [SL]      peers' partials landed

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t crank = cluster_rank(), csize = cluster_size();
  const int cid = blockIdx.x / csize, ncl = gridDim.x / csize;

  // rows of this cluster: as even as possible
  const int64_t base = a.N / ncl, rem = a.N % ncl;
  const int64_t lo = cid * base + min((int64_t)cid, rem);
  const int64_t hi = lo + base + (cid < rem ? 1 : 0);
  const int ntiles = (int)((hi - lo + TN - 1) / TN);

  if (tid == 0) {
    for (int s = 0; s < kStStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kStCWarps); }
    for (int s = 0; s < SL; ++s) { mbar_init(&pr[s], kStCWarps * V); mbar_init(&kr[s], V); mbar_init(&rbar[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // The first tiles only touch this CTA's own shared memory: start them now so that their HBM latency overlaps
  // the ||x||^2 exchange and the start-up cluster barrier (the kernel is ramp-bound at N ~ 500).
  const int nprefetch = min(ntiles, kStStages);
  if (warp == kStCWarps && lane == 0) {
    const float* slice0 = a.bank + (int64_t)crank * SLICE;
    for (int t = 0; t < nprefetch; ++t) {
      const int64_t r0 = lo + (int64_t)t * TN;
      const int rows = (int)min((int64_t)TN, hi - r0);
      mbar_expect_tx(&full[t], (uint32_t)(rows * SLICE * 4));
      for (int r = 0; r < rows; ++r)
        bulk_g2s(tiles + ((size_t)t * TN + r) * SLICE, slice0 + (r0 + r) * a.D, SLICE * 4, &full[t]);
    }
  }
  if (a.zero_word && blockIdx.x == 0 && tid == 0) *a.zero_word = 0.f;
  if (a.xsq == nullptr) {
    // ||x_q||^2 in the kernel: every CTA squares its slice, the per-CTA partials cross the cluster through
    // distributed shared memory and become visible with the start-up cluster barrier below
    float loc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      loc[q] = 0.f;
      if (warp < kStCWarps) {
        const int qs = min(q, a.q_real - 1);
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          const float4 t = *reinterpret_cast<const float4*>(a.x + (int64_t)qs * a.D + (int64_t)crank * SLICE +
                                                            (v * kStCompute + tid) * 4);
          loc[q] = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(t.z, t.z, fmaf(t.w, t.w, loc[q]))));
        }
      }
      loc[q] = warp_sum(loc[q]);
      if (lane == 0 && warp < kStCWarps) wp[warp * Q + q] = loc[q];
    }
    __syncthreads();
    if (tid < Q) {
      float tot = 0.f;
      for (int w = 0; w < kStCWarps; ++w) tot += wp[w * Q + tid];
      for (uint32_t peer = 0; peer < csize; ++peer) st_remote(xsp + crank * 8 + tid, peer, tot);
    }
    __syncthreads();
  }
  // every CTA of the cluster has initialised its barriers (and delivered its ||x||^2 partial) before any
  // remote store can target them
  cluster_arrive();
  cluster_wait();

  if (warp == kStCWarps) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const float* slice0 = a.bank + (int64_t)crank * SLICE;
      for (int t = nprefetch; t < ntiles; ++t) {
        const int s = t % kStStages;
        mbar_wait(&empty[s], (uint32_t)(((t / kStStages) + 1) & 1));
        const int64_t r0 = lo + (int64_t)t * TN;
        const int rows = (int)min((int64_t)TN, hi - r0);
        mbar_expect_tx(&full[s], (uint32_t)(rows * SLICE * 4));
        for (int r = 0; r < rows; ++r)
          bulk_g2s(tiles + ((size_t)s * TN + r) * SLICE, slice0 + (r0 + r) * a.D, SLICE * 4, &full[s]);
      }
    }
  } else if (warp == kStCWarps + 1) {
    // =========================== weights ===========================
    if (lane < V) {
      const int r = lane / Q, q = lane % Q;
      float xs_q = 0.f;
      if (a.xsq) {
        xs_q = a.xsq[min(q, a.q_real - 1)];
      } else {
        for (uint32_t src = 0; src < csize; ++src) xs_q += xsp[src * 8 + q];
      }
      float zacc = 0.f;
      for (int t = 0; t < ntiles; ++t) {
        const int64_t r0 = lo + (int64_t)t * TN;
        const int rows = (int)min((int64_t)TN, hi - r0);
        const float nsq = (r < rows) ? a.sqnorm[r0 + r] : 0.f;
        const int sl = t % SL;
        const uint32_t slp = (uint32_t)((t / SL) & 1);
        mbar_wait(&pr[sl], slp);
        float sum = 0.f;
        const float* w0 = wp + (size_t)sl * kStCWarps * V + lane;
#pragma unroll
        for (int w = 0; w < kStCWarps; ++w) sum += w0[w * V];
        if (lane == 0) mbar_expect_tx(&rbar[sl], csize * V * 4);
        float* slot = recv + ((size_t)sl * kStMaxCluster + crank) * V + lane;
        for (uint32_t peer = 0; peer < csize; ++peer) st_async_remote(slot, &rbar[sl], peer, sum);
        mbar_wait(&rbar[sl], slp);     // every peer's partial of tile t has landed
        float k = 0.f;
        if (r < rows) {
          float dot = 0.f;
          const float* rs = recv + (size_t)sl * kStMaxCluster * V + lane;
          for (uint32_t src = 0; src < csize; ++src) dot += rs[src * V];
          k = weight_from_dot(xs_q, nsq, dot, a.alpha, a.power, a.inv2s2);
          if (crank == 0 && a.k_out && q < a.q_real) a.k_out[(int64_t)q * a.N + r0 + r] = k;
        }
        ks[sl * V + lane] = k;
        zacc += k;
        mbar_arrive(&kr[sl]);
      }
      if (crank == 0) zred[lane] = zacc;
    }
    __syncwarp();
    if (crank == 0 && lane < Q && lane < a.q_real) {
      float z = 0.f;
      for (int r = 0; r < TN; ++r) z += zred[r * Q + lane];
      a.part_z[(int64_t)cid * a.q_real + lane] = z;
    }
  } else {
    // =========================== compute warps ===========================
    // thread owns floats [v*kStCompute*4 + tid*4, +4) of the slice, for every query row
    float4 xr[Q][VPT], acc[Q][VPT];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int qs = min(q, a.q_real - 1);
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        xr[q][v] = *reinterpret_cast<const float4*>(a.x + (int64_t)qs * a.D + (int64_t)crank * SLICE +
                                                    (v * kStCompute + tid) * 4);
        acc[q][v] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    for (int t = 0; t < ntiles + LAG; ++t) {
      if (t < ntiles) {
        // ---------------- phase 1: partial dots of tile t ----------------
        const int s = t % kStStages;
        const int rows = (int)min((int64_t)TN, hi - (lo + (int64_t)t * TN));
        mbar_wait(&full[s], (uint32_t)((t / kStStages) & 1));
        float part[V];
#pragma unroll
        for (int i = 0; i < V; ++i) part[i] = 0.f;
        const float* tile = tiles + (size_t)s * TN * SLICE;
#pragma unroll
        for (int r = 0; r < TN; ++r) {
          if (r < rows) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
              const float4 b = *reinterpret_cast<const float4*>(tile + (size_t)r * SLICE + (v * kStCompute + tid) * 4);
#pragma unroll
              for (int q = 0; q < Q; ++q) {
                float p = part[r * Q + q];
                p = fmaf(b.x, xr[q][v].x, p);
                p = fmaf(b.y, xr[q][v].y, p);
                p = fmaf(b.z, xr[q][v].z, p);
                p = fmaf(b.w, xr[q][v].w, p);
                part[r * Q + q] = p;
              }
            }
          }
        }
        warp_reduce_scatter<V>(part, lane);
        if (warp_value_owner<V>(lane)) {
          wp[((size_t)(t % SL) * kStCWarps + warp) * V + warp_value_index<V>(lane)] = part[0];
          mbar_arrive(&pr[t % SL]);
        }
      }
      if (t >= LAG) {
        // ---------------- phase 2: accumulate tile t-LAG with its weights ----------------
        const int tp = t - LAG;
        const int s = tp % kStStages;
        const int rows = (int)min((int64_t)TN, hi - (lo + (int64_t)tp * TN));
        mbar_wait(&kr[tp % SL], (uint32_t)((tp / SL) & 1));
        const float* tile = tiles + (size_t)s * TN * SLICE;
        const float* kt = ks + (tp % SL) * V;
#pragma unroll
        for (int r = 0; r < TN; ++r) {
          if (r < rows) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
              const float4 b = *reinterpret_cast<const float4*>(tile + (size_t)r * SLICE + (v * kStCompute + tid) * 4);
#pragma unroll
              for (int q = 0; q < Q; ++q) {
                const float k = kt[r * Q + q];
                acc[q][v].x = fmaf(k, b.x, acc[q][v].x);
                acc[q][v].y = fmaf(k, b.y, acc[q][v].y);
                acc[q][v].z = fmaf(k, b.z, acc[q][v].z);
                acc[q][v].w = fmaf(k, b.w, acc[q][v].w);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
    }
    // ---------------- write this cluster's partial sums ----------------
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      if (q < a.q_real) {
#pragma unroll
        for (int v = 0; v < VPT; ++v)
          *reinterpret_cast<float4*>(a.part_num + ((int64_t)cid * a.q_real + q) * a.D + (int64_t)crank * SLICE +
                                     (v * kStCompute + tid) * 4) = acc[q][v];
      }
    }
  }
  // no CTA may exit while a peer can still write into its shared memory
  __syncwarp();
  cluster_arrive();
  cluster_wait();
}

// num[q][d] = sum_c part[c][q][d] ; z[q] = sum_c part_z[c][q].  Block = 32 float4 columns x 8 cluster
// groups: every thread sums a strided subset of the per-cluster partials (independent loads), then the
// 8 groups are combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
k_stream_reduce(const float* __restrict__ part_num, const float* __restrict__ part_z, int ncl, int64_t QD,
                int Q, float* __restrict__ num, float* __restrict__ z) {
  __shared__ float4 sh[8][32];
  const int col = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int64_t j = ((int64_t)blockIdx.x * 32 + col) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < QD) {
#pragma unroll 4
    for (int c = grp; c < ncl; c += 8) {
      const float4 p = ld_stream4(part_num + (int64_t)c * QD + j);
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
  }
  sh[grp][col] = s;
  __syncthreads();
  if (grp == 0 && j < QD) {
#pragma unroll
    for (int g = 1; g < 8; ++g) {
      const float4 p = sh[g][col];
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    *reinterpret_cast<float4*>(num + j) = s;
  }
  if (blockIdx.x == 0 && threadIdx.x < Q) {
    float t = 0.f;
    for (int c0 = 0; c0 < ncl; c0 += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = c0 + u < ncl ? __ldg(part_z + (int64_t)(c0 + u) * Q + threadIdx.x) : 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u) t += v[u];
    }
    z[threadIdx.x] = t;
  }
}

// Same reduction with the correction applied on the fly (single GPU, Q <= 8): x0 <- x0 - scale * num / (z + eps).
// Saves the num round trip and two launches; num / neg / z are still written when the caller wants them.
struct ReduceCorrectArgs {
  const float* part_num; const float* part_z; int ncl; int64_t D; int Q;
  float eps, scale, gate_thr; int flags;
  float* x0; float* num_out; float* z_out; float* neg_out; float* denom_out; int32_t* gate_out; float* mean_out;
};

__global__ void __launch_bounds__(256) k_stream_reduce_correct(const ReduceCorrectArgs a) {
  __shared__ float4 sh[8][32];
  __shared__ float red[33];
  const int col = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int64_t q = blockIdx.y;
  const int64_t QD = (int64_t)a.Q * a.D;
  const int64_t j = ((int64_t)blockIdx.x * 32 + col) * 4;
  // batches of independent loads, summed in cluster order: `z += load` in a plain loop is an L2 round trip per cluster
  float z = 0.f;
  for (int c0 = 0; c0 < a.ncl; c0 += 16) {
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = c0 + u < a.ncl ? __ldg(a.part_z + (int64_t)(c0 + u) * a.Q + q) : 0.f;
#pragma unroll
    for (int u = 0; u < 16; ++u) z += v[u];
  }
  const float denom = z + a.eps;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (a.z_out) a.z_out[q] = z;
    if (a.denom_out) a.denom_out[q] = denom;
    if (a.gate_out) a.gate_out[q] = (!(a.flags & SDN_EPI_GATE) || denom > a.gate_thr) ? 1 : 0;
  }
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < a.D) {
    for (int c0 = grp; c0 < a.ncl; c0 += 64) {
      float4 p[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        p[u] = c0 + 8 * u < a.ncl ? ld_stream4(a.part_num + (int64_t)(c0 + 8 * u) * QD + q * a.D + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += p[u].x; s.y += p[u].y; s.z += p[u].z; s.w += p[u].w; }
    }
  }
  sh[grp][col] = s;
  __syncthreads();
  float msum = 0.f;
  if (grp == 0 && j < a.D) {
#pragma unroll
    for (int g = 1; g < 8; ++g) {
      const float4 p = sh[g][col];
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    const int64_t o = q * a.D + j;
    if (a.num_out) *reinterpret_cast<float4*>(a.num_out + o) = s;
    const float4 n = make_float4(s.x / denom, s.y / denom, s.z / denom, s.w / denom);
    if (a.neg_out) *reinterpret_cast<float4*>(a.neg_out + o) = n;
    if (a.x0) {
      float4 x = *reinterpret_cast<const float4*>(a.x0 + o);
      x.x = fmaf(-a.scale, n.x, x.x);
      x.y = fmaf(-a.scale, n.y, x.y);
      x.z = fmaf(-a.scale, n.z, x.z);
      x.w = fmaf(-a.scale, n.w, x.w);
      *reinterpret_cast<float4*>(a.x0 + o) = x;
    }
    msum = fminf(fmaxf(n.x, -1e10f), 1e10f) + fminf(fmaxf(n.y, -1e10f), 1e10f) +
           fminf(fmaxf(n.z, -1e10f), 1e10f) + fminf(fmaxf(n.w, -1e10f), 1e10f);
  }
  if (a.mean_out) {
    msum = block_sum(msum, red);
    if (threadIdx.x == 0) atomicAdd(a.mean_out, msum / (float)QD);
  }
}

// ------------------------------------------------------------------------------------------ host side
struct StreamPlan {
  int qt = 0, vpt = 0, tn = 0, ns = 0, lag = 1, cs = 0;   // template instance and cluster size
  bool ok = false;
};

static StreamPlan plan_stream(int64_t Q, int64_t N, int64_t D) {
  StreamPlan p;
  if (Q < 1 || Q > 8 || N < 1) return p;
  p.qt = Q <= 1 ? 1 : (Q <= 2 ? 2 : (Q <= 4 ? 4 : 8));
  // slice = vpt * 2048 floats, cluster size = D / slice must be a power of two in [1, 8]
  static const int forced_vpt = [] { const char* e = getenv("SDN_STREAM_VPT"); return e ? atoi(e) : 0; }();
  static const int legacy = [] { const char* e = getenv("SDN_STREAM_LEGACY_TILES"); return e ? atoi(e) : 0; }();
  for (int vpt : {2, 4, 1}) {
    if (forced_vpt && vpt != forced_vpt) continue;
    if (vpt > 1 && p.qt * vpt > 8) continue;
    const int64_t slice = (int64_t)vpt * kStCompute * 4;
    if (D % slice) continue;
    const int64_t cs = D / slice;
    if (cs < 1 || cs > 8 || (cs & (cs - 1))) continue;
    p.vpt = vpt;
    p.cs = (int)cs;
    p.tn = vpt == 1 ? (p.qt == 8 ? 2 : 4) : (vpt == 2 ? 2 : 1);
    p.ns = 6;
    if (!legacy) {
      if (vpt == 2) { p.tn = 4; p.ns = 3; }                  // 3 stages x 4 rows x 16 KiB
      if (vpt == 1 && p.qt < 8) { p.tn = 8; p.ns = 3; }      // 3 x 8 x 8 KiB (Q = 8 keeps 6 x 2 x 8 KiB: 32 values per
                                                             // exchange measured slower, 151 against 119 us at N = 3000)
    }
    static const int lag = [] { const char* e = getenv("SDN_STREAM_LAG"); return e ? atoi(e) : 0; }();
    if (!legacy && lag >= 2 && lag <= 3) {
      // deeper hand-off instead of fatter tiles: six stages of 32 KiB (Q = 8: 16 or 32 KiB), phase 2 trails by `lag` tiles
      if (vpt == 2) { p.tn = 2; p.ns = 6; p.lag = lag; }
      if (vpt == 1 && p.qt == 8) { p.tn = getenv("SDN_STREAM_Q8_TN4") ? 4 : 2; p.ns = 6; p.lag = p.tn == 4 ? 2 : lag; }
    }
    p.ok = true;
    return p;
  }
  return p;
}

template <int Q, int VPT, int TN, int NS, int LAG>
static int launch_stream_t(const StreamArgs& a, int cs, int max_clusters_hint, int* ncl_out, cudaStream_t st) {
  auto kern = k_stream<Q, VPT, TN, NS, LAG>;
  constexpr size_t smem = stream_smem_bytes<Q, VPT, TN, NS, LAG>();
  static std::atomic<bool> configured[kMaxDevices];
  static std::atomic<int> max_clusters_of[kMaxDevices][kStMaxCluster + 1];
  const int dev = device_slot();
  std::atomic<int>* max_clusters_by_cs = max_clusters_of[dev];
  if (!configured[dev].load(std::memory_order_acquire)) {
    SDN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev].store(true, std::memory_order_release);
  }
  if (max_clusters_by_cs[cs] == 0) {
    cudaLaunchConfig_t probe{};
    probe.gridDim = dim3(kNumSMs / cs * cs);
    probe.blockDim = dim3(kStThreads);
    probe.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    probe.attrs = at; probe.numAttrs = 1;
    int n = 0;
    SDN_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &probe));
    max_clusters_by_cs[cs] = std::max(1, n);
  }
  const int max_clusters = max_clusters_by_cs[cs];
  int ncl = std::min(max_clusters, max_clusters_hint);
  ncl = (int)std::min<int64_t>(ncl, std::max<int64_t>(1, cdiv(a.N, TN)));
  *ncl_out = ncl;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ncl * cs);
  cfg.blockDim = dim3(kStThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  SDN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a));
  SDN_LAUNCHED();
  return SDN_OK;
}

static int max_clusters_for(int cs) { return kNumSMs / cs; }

bool stream_supported(int64_t Q, int64_t N, int64_t D) { return plan_stream(Q, N, D).ok; }

size_t stream_workspace_bytes(int64_t Q, int64_t N, int64_t D) {
  const StreamPlan p = plan_stream(Q, N, D);
  if (!p.ok) return 0;
  const size_t ncl = max_clusters_for(p.cs);
  return ncl * (size_t)Q * D * 4 + ncl * (size_t)Q * 4 + 256;
}

static int stream_launch(const float* bank, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                         const float* xsq, int64_t Q, float inv2s2, int power, float alpha, float* k_out,
                         float* zero_word, void* ws, size_t ws_bytes, cudaStream_t st, StreamArgs* a_out, int* ncl_out) {
  const StreamPlan p = plan_stream(Q, N, D);
  if (!p.ok) return SDN_E_UNSUPPORTED;
  if (!ws || ws_bytes < stream_workspace_bytes(Q, N, D)) return SDN_E_WORKSPACE;
  const int hint = max_clusters_for(p.cs);
  StreamArgs a{};
  a.bank = bank; a.sqnorm = sqnorm; a.N = N; a.D = D; a.x = xq; a.xsq = xsq; a.q_real = (int)Q;
  a.inv2s2 = inv2s2; a.power = power; a.alpha = alpha; a.k_out = k_out; a.zero_word = zero_word;
  a.part_num = static_cast<float*>(ws);
  a.part_z = a.part_num + (size_t)hint * Q * D;
  int ncl = 0, rc = SDN_E_UNSUPPORTED;
  int pid = g_prof.begin("k_stream", st);
#define SDN_ST_CASE(QQ, VV, TT, SS, LL) \
  if (p.qt == QQ && p.vpt == VV && p.tn == TT && p.ns == SS && p.lag == LL) rc = launch_stream_t<QQ, VV, TT, SS, LL>(a, p.cs, hint, &ncl, st);
  SDN_ST_CASE(1, 1, 4, 6, 1) SDN_ST_CASE(2, 1, 4, 6, 1) SDN_ST_CASE(4, 1, 4, 6, 1) SDN_ST_CASE(8, 1, 2, 6, 1)
  SDN_ST_CASE(1, 2, 2, 6, 1) SDN_ST_CASE(2, 2, 2, 6, 1) SDN_ST_CASE(4, 2, 2, 6, 1)
  SDN_ST_CASE(1, 4, 1, 6, 1) SDN_ST_CASE(2, 4, 1, 6, 1)
  SDN_ST_CASE(1, 2, 4, 3, 1) SDN_ST_CASE(2, 2, 4, 3, 1) SDN_ST_CASE(4, 2, 4, 3, 1)
  SDN_ST_CASE(1, 1, 8, 3, 1) SDN_ST_CASE(2, 1, 8, 3, 1) SDN_ST_CASE(4, 1, 8, 3, 1)
  SDN_ST_CASE(1, 2, 2, 6, 2) SDN_ST_CASE(2, 2, 2, 6, 2) SDN_ST_CASE(4, 2, 2, 6, 2)
  SDN_ST_CASE(1, 2, 2, 6, 3) SDN_ST_CASE(2, 2, 2, 6, 3) SDN_ST_CASE(4, 2, 2, 6, 3)
  SDN_ST_CASE(8, 1, 2, 6, 2) SDN_ST_CASE(8, 1, 2, 6, 3) SDN_ST_CASE(8, 1, 4, 6, 2)
#undef SDN_ST_CASE
  g_prof.end(pid, st);
  *a_out = a;
  *ncl_out = ncl;
  return rc;
}

int stream_partial(const float* bank, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                   const float* xsq, int64_t Q, float inv2s2, int power, float alpha, float* num, float* z,
                   float* k_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  StreamArgs a{};
  int ncl = 0;
  const int rc = stream_launch(bank, sqnorm, N, D, xq, xsq, Q, inv2s2, power, alpha, k_out, nullptr, ws, ws_bytes, st,
                               &a, &ncl);
  if (rc) return rc;
  const int64_t QD = Q * D;
  const int pid = g_prof.begin("k_stream_reduce", st);
  k_stream_reduce<<<(unsigned)cdiv(QD, 128), 256, 0, st>>>(a.part_num, a.part_z, ncl, QD, (int)Q, num, z);
  g_prof.end(pid, st);
  SDN_LAUNCHED();
  return SDN_OK;
}

int stream_conditioning(const float* bank, const float* sqnorm, int64_t N, int64_t D, float* x0_inout,
                        const float* xq, int64_t Q, float inv2s2, int power, float alpha, float eps, float scale,
                        float gate_thr, int flags, float* num_out, float* z_out, float* neg_out, float* denom_out,
                        int32_t* gate_out, float* mean_out, float* k_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  StreamArgs a{};
  int ncl = 0;
  const int rc = stream_launch(bank, sqnorm, N, D, xq, nullptr, Q, inv2s2, power, alpha, k_out, mean_out, ws, ws_bytes,
                               st, &a, &ncl);
  if (rc) return rc;
  ReduceCorrectArgs r{};
  r.part_num = a.part_num; r.part_z = a.part_z; r.ncl = ncl; r.D = D; r.Q = (int)Q; r.eps = eps; r.scale = scale;
  r.gate_thr = gate_thr; r.flags = flags; r.x0 = x0_inout; r.num_out = num_out; r.z_out = z_out; r.neg_out = neg_out;
  r.denom_out = denom_out; r.gate_out = gate_out; r.mean_out = mean_out;
  const int pid = g_prof.begin("k_stream_reduce_correct", st);
  k_stream_reduce_correct<<<dim3((unsigned)cdiv(D, 128), (unsigned)Q), 256, 0, st>>>(r);
  g_prof.end(pid, st);
  SDN_LAUNCHED();
  return SDN_OK;
}

}  // namespace sdn
