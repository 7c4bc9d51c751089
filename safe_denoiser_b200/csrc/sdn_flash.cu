// One-pass "flash-repellency" for batched calls: the bank is read from HBM exactly once.
//
// Replaces repellency_methods_fast.py:249-250 (cdist + the [Q,N,D+1] broadcast) for 8 < Q <= 128 query rows per pass.
// The two-phase tcgen05 path (sdn_umma.cu) streams the bank twice from HBM because the weights of a bank row need its
// dot product over ALL of D before the row can be accumulated, and its two kernels are a whole bank apart.  Here both
// contractions run in ONE persistent kernel, a few tiles apart, so that the second read of a tile hits the L2:
//
//   grid   = D / 128 CTAs (one per SM, clusters of 4); CTA j owns the d-slice [128 j, 128 j + 128) of every bank row
//            and keeps  X[:, slice]  (bf16 hi/lo, tcgen05 A operand)  and  num[:, slice]  (fp32 accumulator) in
//            TENSOR MEMORY for the whole kernel;
//   tile   = 64 bank rows x the slice, hi and lo planes (32 KiB) = one shared-memory stage filled by TMA;
//   phase A  S_j[q][i] = sum_{d in slice} X[q][d] bank[i][d]      tcgen05.mma, A = X (TMEM), B = tile (K-major);
//            the tile is loaded with an L2 evict_last policy and its stage is released at once;
//   reduce   S[q][i] = sum_j S_j[q][i] over the D/128 CTAs:
//              level 1  inside the cluster over distributed shared memory (st.async + mbarrier complete_tx):
//                       CTA c of a cluster receives and sums rows [16c, 16c+16) of the tile;
//              level 2  the cluster partials go through L2 as self-validating 16-byte lines {v, tag, v, tag}
//                       (no fence, no flag) to "jobs" of 4 rows x 64 queries dealt round-robin over ALL CTAs: the
//                       job's CTA sums the 32 cluster partials in cluster order, computes k = exp(-dist / 2 sigma^2)
//                       and publishes the weights as fp32 words whose low 4 mantissa bits carry the tile's sequence
//                       number (again no fence, no flag: every word validates itself);
//   phase B  num[q][slice] += sum_i k[q][i] bank[i][slice]        tcgen05.mma, A = weights (bf16 hi/lo written to
//            TMEM by tcgen05.st), B = the tile viewed MN-major, loaded AGAIN by TMA (evict_first: it is not needed
//            any more) when the tile's weights show up -- a hit in the 126 MB L2 as long as the exchange lags the
//            stream by less than the L2 (a resident-tile variant was measured first: with 6 stages per SM the
//            ~15 us round trip of the exchange capped it at 1.3 TB/s).
//
// Every sum has a fixed order (CTA rank, cluster index, row index): results are bit-reproducible run to run.
//
// Warp roles (16 warps, 1 CTA per SM): 0 TMA producer phase A | 1 MMA issuer (event loop over "tile ta loaded" /
// "weights and tile tb ready") | 2-3 level-1 sum | 4-7 phase-A drain: TMEM -> registers -> staging -> bulk DSMEM copies |
// 8-11 weights: global -> bf16 hi/lo -> TMEM (warp 8 also issues the second TMA read of the tile), z_q, and the final
// epilogue | 12-15 level-2 jobs: ||x||^2, sum over clusters, exp, publish.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "sdn_internal.h"
#include "sdn_ptx.cuh"

namespace sdn {

constexpr int kFR = 64;                 // bank rows per tile
constexpr int kFDS = 128;               // d per CTA
constexpr int kFCS = 4;                 // CTAs per cluster (this pool's B200s co-schedule only 15 clusters of 8 = 120 CTAs)
constexpr int kFRowsPerOwner = kFR / kFCS;   // rows of a tile that CTA c of a cluster reduces at level 1
constexpr int kFJobRows = 4;            // rows of a tile per level-2 job (one float4 per query row)
constexpr int kFJobsPerGroup = kFR / kFJobRows;   // 16 jobs per tile and group of 64 query rows
constexpr int kFRing = 32;              // slots of the global exchange rings: phase A runs at most kFWindow < kFRing tiles ahead of phase B
constexpr int kFWindow = 12;            // tiles between the two reads of the bank: 12 x 4 MiB stay in the 126 MB L2 (and < the 16 tiles between two units of a level-2 worker)
constexpr int kFThreads = 512;
constexpr int kFMaxCtas = 128;
constexpr int kFMaxClusters = kFMaxCtas / kFCS;
constexpr uint32_t kFStageBytes = 2u * kFR * kFDS * 2u;   // hi + lo tiles of [64 rows][128 d] bf16 = 32 KiB
constexpr uint32_t kFTmemCols = 512;
constexpr uint32_t kFSpinLimit = 1u << 24;

template <int G>
struct FCfg {
  static constexpr int kQ = 64 * G;                                   // query rows per pass
  static constexpr int kSA = 2;                                       // stages of phase A (HBM stream; released as soon as the MMAs have read them)
  static constexpr int kSB = 2;                                       // stages of phase B (second read, from L2)
  static constexpr int kStages = kSA + kSB;
  static constexpr int kSlots = 3;                                    // level-1 units in flight in the cluster (receive slots = staging buffers)
  static constexpr int kSBuf = G == 1 ? 4 : 2;                        // S accumulators in tensor memory
  // level-1 unit = (tile, group of 64 query rows); a slice = what one CTA of the cluster owns of a unit
  static constexpr int kSliceFloats = 64 * kFRowsPerOwner;            // [sub 4][q 64][4 rows]
  static constexpr int kSliceVec = kSliceFloats / 4 / 128;            // float4 per thread of the level-1 sum
  static constexpr uint32_t kRbufSlotBytes = (uint32_t)kFCS * kSliceFloats * 4;   // [src or owner][slice]
  static constexpr int kJobs = kFJobsPerGroup * G;                    // level-2 jobs per tile
  static constexpr int kXStages = (kQ * 132 * 4 + (int)kFStageBytes - 1) / (int)kFStageBytes;   // stages the query staging covers
  static constexpr int kXOverA = kXStages > kSB ? kXStages - kSB : 0;                           // ... of which phase-A stages
  // tensor-memory columns: X | S accumulators | weight operands | num accumulator
  static constexpr uint32_t kColX = 0;
  static constexpr uint32_t kColS = 64 * G;
  static constexpr uint32_t kColP = kColS + 64 * kSBuf;
  static constexpr uint32_t kPStride = 32 * G;                        // columns of one weight operand buffer (G = 2: hi 32 | lo 32)
  static constexpr uint32_t kColAcc = 384;
  static_assert(kColP + 2 * kPStride <= kColAcc, "tensor memory layout");
  static constexpr uint32_t kSmemBytes = kStages * kFStageBytes + 2 * kSlots * kRbufSlotBytes + 1024 /*align*/ + 1536 /*barriers, xsq*/;
  static_assert(kSmemBytes <= 232448, "more than 227 KiB of shared memory");
};

// Arena: library-owned per-device synchronisation memory of the one-pass kernel (zeroed once; tags grow
// monotonically over launches, so stale lines never match).
struct FlashArena {
  uint32_t* epoch;      // [1]  tag base of the next launch
  uint8_t* xsq_ll;      // [kFMaxCtas][128] 16-byte lines: ||x_q||^2 partial of every CTA
  uint8_t* part_ll;     // [ring][cluster 32][job 32][q 64][2 lines]   level-2 fan-in
  float* gw;            // [ring][job 32][q 64][4 rows]  published weights
  uint32_t* wflag;      // [ring][128] one release flag per (job, owner warp)
  uint32_t* diag;       // host-mapped: written before a timeout trap
  unsigned long long* trace;   // SDN_FLASH_TRACE=1: [cta][tile < 64][16 events] globaltimer ns (null otherwise)
};

struct FlashArgs {
  const float* xq;      // [Q][D] query the distances are taken on
  const float* sqnorm;  // [N]
  int Q, N, ntiles, nclusters, window;
  unsigned mma_sleep, poll_sleep;     // back-off of the polling loops (ns)
  int64_t D;
  float inv2s2, alpha; int power;
  float* num_out;       // [Q][D] or null
  float* z_out;         // [Q] or null
  float* k_out;         // [Q][N] or null
  FlashEpi epi;
  FlashArena ar;
};

__device__ __noinline__ void f_timeout(const FlashArgs& a, uint32_t code, uint32_t tile, uint32_t extra) {
  if (a.ar.diag) {
    volatile uint32_t* d = a.ar.diag;     // host-mapped: survives the trap (first reporter wins, racily)
    if (d[0] == 0u) {
      d[0] = code; d[1] = blockIdx.x; d[2] = threadIdx.x; d[3] = tile; d[4] = extra;
      __threadfence_system();
    }
  }
  __trap();
}

constexpr int kFTraceTiles = 64, kFTraceEvents = 16;
__device__ __forceinline__ void f_trace(const FlashArgs& a, int tile, int ev) {
  if (a.ar.trace && tile < kFTraceTiles) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.ar.trace[((size_t)blockIdx.x * kFTraceTiles + tile) * kFTraceEvents + ev] = t;
  }
}

__device__ __forceinline__ void f_trace_val(const FlashArgs& a, int tile, int ev, unsigned long long v) {
  if (a.ar.trace && tile < kFTraceTiles) a.ar.trace[((size_t)blockIdx.x * kFTraceTiles + tile) * kFTraceEvents + ev] = v;
}

// bounded mbarrier wait (a broken pipeline must not hang the GPU)
__device__ __forceinline__ void f_wait(const FlashArgs& a, uint64_t* bar, uint32_t parity, uint32_t code, uint32_t tile) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(u_smem(bar)), "r"(parity) : "memory");
    if (spin > kFSpinLimit) f_timeout(a, code, tile, parity);
  }
}

// same, observing arrivals made by peer CTAs of the cluster
__device__ __forceinline__ void f_wait_cluster(const FlashArgs& a, uint64_t* bar, uint32_t parity, uint32_t code, uint32_t tile) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(u_smem(bar)), "r"(parity) : "memory");
    if (spin > kFSpinLimit) f_timeout(a, code, tile, parity);
  }
}

__device__ __forceinline__ uint32_t f_pack_bf16(float lo_elem, float hi_elem) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);   // .x (low 16 bits) = first argument
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float f_bf16_hi(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// TMA tile load with an L2 eviction policy (phase A: evict_last, the tile is read again a few tiles later;
// phase B: evict_first, it is never needed again)
__device__ __forceinline__ void f_tma_2d_hint(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(u_smem(dst)), "l"(map), "r"(u_smem(bar)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}

struct FSmem {
  uint8_t* stages;      // [kSA phase-A stages | kSB phase-B stages] x 32 KiB
  float* rbuf;          // [kSlots][kFCS][slice floats] receive slots, then [kSlots][kFCS][slice floats] send staging
  uint64_t* afull; uint64_t* aempty;      // [kSA]
  uint64_t* bfull; uint64_t* bempty;      // [kSB]
  uint64_t* bgo;                          // [4] weights of tile u seen: its second load may start
  uint64_t* sfull; uint64_t* sempty;      // [kSBuf] S accumulators
  uint64_t* pfull; uint64_t* pempty;      // [2] weight operands
  uint64_t* rfull; uint64_t* rfree;       // [kSlots] level-1 receive slots
  uint64_t* xfull; uint64_t* accfull;     // [1]
  uint64_t* xload;                        // [4] query rows of one phase-A warp staged in shared memory
  uint32_t* tmem_base;
  float* xsq;           // [128]
  float* xsq_half;      // [128]  (||x||^2 halves at start, z_q at the end)
};

template <int G>
__device__ __forceinline__ FSmem f_carve(unsigned char* raw) {
  using C = FCfg<G>;
  FSmem s;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023;
  s.stages = reinterpret_cast<uint8_t*>(a);
  s.rbuf = reinterpret_cast<float*>(s.stages + (size_t)C::kStages * kFStageBytes);
  uint64_t* b = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s.rbuf) + 2 * C::kSlots * C::kRbufSlotBytes);
  s.afull = b; b += 4;
  s.aempty = b; b += 4;
  s.bfull = b; b += 4;
  s.bempty = b; b += 4;
  s.bgo = b; b += 4;
  s.sfull = b; b += 4;
  s.sempty = b; b += 4;
  s.pfull = b; b += 2;
  s.pempty = b; b += 2;
  s.rfull = b; b += 4;
  s.rfree = b; b += 4;
  s.xfull = b; b += 1;
  s.accfull = b; b += 1;
  s.xload = b; b += 4;
  s.tmem_base = reinterpret_cast<uint32_t*>(b); b += 1;
  s.xsq = reinterpret_cast<float*>(b);          // 51 x 8 = 408 bytes of barriers so far
  s.xsq_half = s.xsq + 128;
  return s;
}

// Staging of [kQ rows][128 floats] in shared memory with a row pitch of 132 floats: thread = row reads (query
// prologue) and writes (epilogue) are conflict-free per quarter warp, while the global side moves whole 512-byte rows.
constexpr int kFPitch = 132;

template <int G>
__global__ void __launch_bounds__(kFThreads, 1)
k_flash(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
        const __grid_constant__ FlashArgs a) {
  using C = FCfg<G>;
  extern __shared__ unsigned char smem_raw[];
  const FSmem sm = f_carve<G>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = u_cluster_rank();
  const int cta = blockIdx.x;
  const int cluster = cta / kFCS;
  const int d0 = cta * kFDS;
  const int ntiles = a.ntiles;
  const int nclusters = a.nclusters;
  // the query slice is staged in the LAST stages of the ring (their first TMA waits for the prologue)
  float* const xstage = reinterpret_cast<float*>(sm.stages + (size_t)(C::kStages - C::kXStages) * kFStageBytes);
  float* const estage = reinterpret_cast<float*>(sm.stages);      // epilogue staging: every stage is free by then

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kSA; ++s) { u_mbar_init(&sm.afull[s], 1); u_mbar_init(&sm.aempty[s], 1); }
    for (int s = 0; s < C::kSB; ++s) { u_mbar_init(&sm.bfull[s], 1); u_mbar_init(&sm.bempty[s], 1); }
    for (int s = 0; s < 4; ++s) u_mbar_init(&sm.bgo[s], 1);
    for (int b = 0; b < C::kSBuf; ++b) { u_mbar_init(&sm.sfull[b], 1); u_mbar_init(&sm.sempty[b], 4); }
    for (int b = 0; b < 2; ++b) { u_mbar_init(&sm.pfull[b], 4); u_mbar_init(&sm.pempty[b], 1); }
    for (int b = 0; b < C::kSlots; ++b) { u_mbar_init(&sm.rfull[b], 1); u_mbar_init(&sm.rfree[b], kFCS * 2); }
    u_mbar_init(sm.xfull, 4); u_mbar_init(sm.accfull, 1);
    for (int w = 0; w < 4; ++w) u_mbar_init(&sm.xload[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // arm the level-1 receive slots for their first tiles
    for (int b = 0; b < C::kSlots; ++b) u_mbar_expect_tx(&sm.rfull[b], C::kRbufSlotBytes);
    u_prefetch_map(&tm_hi); u_prefetch_map(&tm_lo);
    if (cta == 0 && a.epi.zero_mean && a.epi.mean_out) { *a.epi.mean_out = 0.f; __threadfence(); }
  }
  if (warp == 1) u_tmem_alloc(sm.tmem_base, kFTmemCols);
  // tag base of this launch: read by everyone before anything is exchanged (CTA 0 advances it at the very end)
  const uint32_t epoch0 = *reinterpret_cast<volatile uint32_t*>(a.ar.epoch);        // sequence number of tile 0
  const uint32_t launch0 = *reinterpret_cast<volatile uint32_t*>(a.ar.epoch + 1);    // tag of the ||x||^2 exchange
  u_fence_before();
  __syncthreads();
  u_fence_after();
  const uint32_t tmem = *sm.tmem_base;

  uint64_t pol_keep, pol_drop;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
  // tile t of this CTA's slice: hi and lo planes, two boxes of [64 rows][64 d] each
  auto load_tile = [&](uint8_t* st, uint64_t* bar, int t, uint64_t policy) {
    u_mbar_expect_tx(bar, kFStageBytes);
    const int r0 = t * kFR;
    f_tma_2d_hint(st, &tm_hi, d0, r0, bar, policy);
    f_tma_2d_hint(st + 8192, &tm_hi, d0 + 64, r0, bar, policy);
    f_tma_2d_hint(st + 16384, &tm_lo, d0, r0, bar, policy);
    f_tma_2d_hint(st + 24576, &tm_lo, d0 + 64, r0, bar, policy);
  };
  auto load_a = [&](int t) {
    f_trace(a, t, 0);
    load_tile(sm.stages + (size_t)(t % C::kSA) * kFStageBytes, &sm.afull[t % C::kSA], t, pol_keep);
  };
  // the first stages depend on nothing but this CTA's own barriers: start the stream before the cluster barrier
  const int npre = min(ntiles, C::kSA - C::kXOverA);
  if (threadIdx.x == 0)
    for (int t = 0; t < npre; ++t) load_a(t);

  // every CTA of the cluster has initialised its barriers before any peer stores into its shared memory
  u_cluster_arrive();
  u_cluster_wait();

  if (warp == 0) {
    // ============================================================ TMA producer, phase A (HBM stream)
    if (lane == 0) {
      if (C::kXOverA > 0 && ntiles > npre) f_wait(a, sm.xfull, 0, 0x110, 0);      // the query staging area becomes a stage
      for (int t = npre; t < ntiles; ++t) {
        if (t >= C::kSA) f_wait(a, &sm.aempty[t % C::kSA], (uint32_t)(((t / C::kSA) + 1) & 1), 0x100, t);
        load_a(t);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idA = u_idesc(128, kFR, 0, 0);      // S[128][64 rows]  = X (TMEM) x tile (K-major)
      constexpr uint32_t idB = u_idesc(128, kFDS, 0, 1);     // num[128][128 d] += P (TMEM) x tile (MN-major)
      f_wait(a, sm.xfull, 0, 0x200, 0);
      u_fence_after();
      int ta = 0, tb = 0;
      uint32_t idle = 0;
      while (tb < ntiles) {
        bool did = false;
        if (tb < ta && u_mbar_test(&sm.pfull[tb & 1], (uint32_t)((tb >> 1) & 1)) &&
            u_mbar_test(&sm.bfull[tb % C::kSB], (uint32_t)((tb / C::kSB) & 1))) {
          // ---- phase B of tile tb: 4 K steps of 16 bank rows
          u_fence_after();
          f_trace(a, tb, 8);
          const uint32_t base = u_smem(sm.stages + (size_t)(C::kSA + tb % C::kSB) * kFStageBytes);
          const uint32_t acc = tmem + C::kColAcc;
          const uint32_t pb = tmem + C::kColP + (uint32_t)(tb & 1) * C::kPStride;
#pragma unroll
          for (int kk = 0; kk < kFR / 16; ++kk) {
            const uint64_t bh = u_desc(base + kk * 2048, 8192, 1024);
            const uint64_t bl = u_desc(base + 16384 + kk * 2048, 8192, 1024);
            const uint32_t first = (tb > 0 || kk > 0) ? 1u : 0u;
            u_mma_ts(acc, pb + kk * 8, bh, idB, first);
            u_mma_ts(acc, pb + kk * 8, bl, idB, 1u);
            if constexpr (G == 2) u_mma_ts(acc, pb + 32 + kk * 8, bh, idB, 1u);
          }
          u_commit(&sm.bempty[tb % C::kSB]);
          u_commit(&sm.pempty[tb & 1]);
          if (tb == ntiles - 1) u_commit(sm.accfull);
          ++tb;
          did = true;
        }
        // phase A stays within `window` tiles of phase B: the exchange rings (kFRing slots) are reused safely and the
        // tiles waiting for their second read fit in the L2
        if (ta < ntiles && ta - tb < a.window && u_mbar_test(&sm.afull[ta % C::kSA], (uint32_t)((ta / C::kSA) & 1)) &&
            (ta < C::kSBuf || u_mbar_test(&sm.sempty[ta % C::kSBuf], (uint32_t)(((ta / C::kSBuf) + 1) & 1)))) {
          // ---- phase A of tile ta: 8 K steps of 16 d; its stage is free again as soon as these MMAs have read it
          u_fence_after();
          f_trace(a, ta, 1);
          const uint32_t base = u_smem(sm.stages + (size_t)(ta % C::kSA) * kFStageBytes);
          const uint32_t acc = tmem + C::kColS + (uint32_t)(ta % C::kSBuf) * 64;
#pragma unroll
          for (int kk = 0; kk < kFDS / 16; ++kk) {
            const uint32_t off = (uint32_t)(kk >> 2) * 8192 + (uint32_t)(kk & 3) * 32;
            const uint64_t bh = u_desc(base + off, 16, 1024);
            const uint64_t bl = u_desc(base + 16384 + off, 16, 1024);
            u_mma_ts(acc, tmem + C::kColX + kk * 8, bh, idA, kk > 0 ? 1u : 0u);
            u_mma_ts(acc, tmem + C::kColX + kk * 8, bl, idA, 1u);
            if constexpr (G == 2) u_mma_ts(acc, tmem + C::kColX + 64 + kk * 8, bh, idA, 1u);
          }
          u_commit(&sm.sfull[ta % C::kSBuf]);
          u_commit(&sm.aempty[ta % C::kSA]);
          ++ta;
          did = true;
        }
        if (did) idle = 0;
        else {
          // this thread shares its scheduler with a drain, a weights and a level-2 warp: do not spin at full rate
          if (a.mma_sleep) __nanosleep(a.mma_sleep);
          if (++idle > kFSpinLimit) f_timeout(a, 0x210, (uint32_t)ta, (uint32_t)tb);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================================================ phase-A drain + level-1 reduction
    const int lq = warp & 3;
    const int L = lq * 32 + lane;                       // TMEM lane of this thread
    const int part = G == 1 ? (lane >> 4) : 0;          // G = 1: lanes 0-15 hold hi parts, 16-31 lo parts of the same q
    const int q = G == 1 ? (lq * 16 + (lane & 15)) : L;
    const uint32_t tlane = tmem + ((uint32_t)(lq * 32) << 16);
    // ---- query slice -> bf16 hi / lo -> tensor memory; ||x_q||^2 partial of this slice
    {
      // whole 512-byte rows by bulk copies (one per lane), then thread = row out of shared memory: a warp-wide load of
      // 32 different rows costs 32 L1 wavefronts per instruction (measured: 10 us of prologue)
      constexpr int kRowsPerWarp = 16 * G;
      const int row0 = lq * kRowsPerWarp;
      const int nvalid = max(0, min(kRowsPerWarp, a.Q - row0));
      if (lane == 0) u_mbar_expect_tx(&sm.xload[lq], (uint32_t)nvalid * (kFDS * 4));
      __syncwarp();
      if (lane < nvalid) {
        const int r = row0 + lane;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(u_smem(xstage + (size_t)r * kFPitch)), "l"(a.xq + (int64_t)r * a.D + d0), "r"(kFDS * 4),
                       "r"(u_smem(&sm.xload[lq])) : "memory");
      }
      f_wait(a, &sm.xload[lq], 0, 0x340, 0);
      const bool valid = q < a.Q;
      const float* xr = xstage + (size_t)q * kFPitch;
      float ss = 0.f;
#pragma unroll 1
      for (int c4 = 0; c4 < kFDS / 32; ++c4) {           // 32 d = 16 columns per step
        float f[32];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const float4 t4 = valid ? *reinterpret_cast<const float4*>(xr + c4 * 32 + v * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          f[v * 4 + 0] = t4.x; f[v * 4 + 1] = t4.y; f[v * 4 + 2] = t4.z; f[v * 4 + 3] = t4.w;
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float h0 = f_bf16_hi(f[2 * j]), h1 = f_bf16_hi(f[2 * j + 1]);
          hi[j] = f_pack_bf16(h0, h1);
          lo[j] = f_pack_bf16(f[2 * j] - h0, f[2 * j + 1] - h1);
          ss = fmaf(f[2 * j], f[2 * j], ss);
          ss = fmaf(f[2 * j + 1], f[2 * j + 1], ss);
        }
        if constexpr (G == 1) {
          uint32_t sel[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) sel[j] = part ? lo[j] : hi[j];
          u_tmem_st16(tlane + C::kColX + c4 * 16, sel);
        } else {
          u_tmem_st16(tlane + C::kColX + c4 * 16, hi);
          u_tmem_st16(tlane + C::kColX + 64 + c4 * 16, lo);
        }
      }
      u_tmem_st_wait();
      if (part == 0)
        u_ll_store(a.ar.xsq_ll + ((size_t)cta * 128 + q) * 16, ss, 0.f, launch0);
      u_fence_before();
      __syncwarp();
      if (lane == 0) u_mbar_arrive(sm.xfull);
    }

    // ---- level 1, per unit u = (tile t, group g of 64 query rows):
    //   drain   S(t) rows of the group: TMEM -> registers -> send staging in shared memory, laid out per owner CTA
    //           [owner 4][sub 4][q 64][4 rows] (the owner of rows [16c, 16c+16) of the tile is CTA c of the cluster);
    //   send    one 4 KiB bulk copy (cp.async.bulk shared::cta -> shared::cluster) per owner: ONE transaction on the
    //           owner's mbarrier instead of 256 (16-byte st.async from registers kept that barrier busy for 2.6 us a tile);
    //   sum     the previous unit's slices of this CTA over the 4 senders, published as LL lines for the level-2 jobs.
    // Phase A runs several tiles ahead of the exchange, so this loop never waits for HBM.
    float* const stg = sm.rbuf + (size_t)C::kSlots * (C::kRbufSlotBytes / 4);
    const int myg = G == 1 ? 0 : (L >> 6);
    const int q64 = G == 1 ? q : (L & 63);
    const int nunits = ntiles * G;
#pragma unroll 1
    for (int u = 0; u < nunits; ++u) {
      const int t = u / G, g = u - t * G;
      const int sb = u % C::kSlots;
      if (warp == 4 && lane == 0) f_trace(a, t, 12);
      // the staging buffer and the owners' receive slot are free once every owner has summed unit u - kSlots
      if (u >= C::kSlots) f_wait_cluster(a, &sm.rfree[sb], (uint32_t)(((u / C::kSlots) + 1) & 1), 0x320, u);
      if (warp == 4 && lane == 0) f_trace(a, t, 13);
      float* sbuf = stg + (size_t)sb * (C::kRbufSlotBytes / 4);
      if (G == 1 || g == myg) {
        const int b = t % C::kSBuf;
        f_wait(a, &sm.sfull[b], (uint32_t)((t / C::kSBuf) & 1), 0x310, t);
        u_fence_after();
        if (warp == 4 && lane == 0) f_trace(a, t, 2);
        uint32_t r0[32], r1[32];
        u_tmem_ld32_nowait(tlane + C::kColS + (uint32_t)b * 64, r0);
        u_tmem_ld32_nowait(tlane + C::kColS + (uint32_t)b * 64 + 32, r1);
        u_tmem_ld_wait();
        u_fence_before();
        __syncwarp();
        if (lane == 0) u_mbar_arrive(&sm.sempty[b]);
        constexpr int kSubs = kFRowsPerOwner / 4;
        if constexpr (G == 1) {
          // lanes l and l+16 hold the hi-part and lo-part rows of the same query: add them, then each stages one half
          constexpr int kOwnersPerHalf = 32 / kFRowsPerOwner;
#pragma unroll
          for (int oo = 0; oo < kOwnersPerHalf; ++oo) {
#pragma unroll
            for (int v = 0; v < kSubs; ++v) {
              float4 o4;
              float* o = &o4.x;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = oo * kFRowsPerOwner + 4 * v + e;
                const float v0 = __uint_as_float(r0[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r0[j]), 16);
                const float v1 = __uint_as_float(r1[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r1[j]), 16);
                o[e] = part ? v1 : v0;
              }
              const int owner = part * kOwnersPerHalf + oo;
              *reinterpret_cast<float4*>(sbuf + ((size_t)(owner * kSubs + v) * 64 + q64) * 4) = o4;
            }
          }
        } else {
#pragma unroll
          for (int oo = 0; oo < kFCS; ++oo) {
#pragma unroll
            for (int v = 0; v < kSubs; ++v) {
              const int e = oo * kFRowsPerOwner + 4 * v;       // row of the tile, compile-time
              const uint32_t* r = e < 32 ? r0 : r1;
              const int o = e & 31;
              *reinterpret_cast<float4*>(sbuf + ((size_t)(oo * kSubs + v) * 64 + q64) * 4) =
                  make_float4(__uint_as_float(r[o]), __uint_as_float(r[o + 1]), __uint_as_float(r[o + 2]), __uint_as_float(r[o + 3]));
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the bulk copies read what the threads wrote
      asm volatile("bar.sync 3, 128;" ::: "memory");
      if (warp == 4 && lane == 0) f_trace(a, t, 14);
      if (warp == 4 && lane < kFCS) {
        const uint32_t src = u_smem(sbuf + (size_t)lane * C::kSliceFloats);
        const uint32_t dst = u_mapa(u_smem(sm.rbuf + (size_t)sb * (C::kRbufSlotBytes / 4) + (size_t)crank * C::kSliceFloats), (uint32_t)lane);
        const uint32_t rbar = u_mapa(u_smem(&sm.rfull[sb]), (uint32_t)lane);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "r"(src), "r"(C::kSliceFloats * 4), "r"(rbar) : "memory");
      }
      if (warp == 4 && lane == 0) f_trace(a, t, 3);
    }
  } else if (warp == 2 || warp == 3) {
    // ============================================================ level-1 sum: the rows this CTA owns, over the cluster
    const int lt = (warp - 2) * 32 + lane;            // 0..63
    const int nunits = ntiles * G;
#pragma unroll 1
    for (int v = 0; v < nunits; ++v) {
      const int t = v / G, g = v - t * G;
      const int slot = v % C::kSlots;
      f_wait_cluster(a, &sm.rfull[slot], (uint32_t)((v / C::kSlots) & 1), 0x300, v);
      if (warp == 2 && lane == 0 && v + C::kSlots < nunits) u_mbar_expect_tx(&sm.rfull[slot], C::kRbufSlotBytes);   // next use
      if (warp == 2 && lane == 0) f_trace(a, t, 4);
      const uint32_t tag = epoch0 + (uint32_t)t;            // the tile's sequence number (gap-free over launches)
      const float* rb = sm.rbuf + (size_t)slot * (C::kRbufSlotBytes / 4);
      // cluster partial as LL lines, laid out per level-2 job: [job 32][q 64][2 lines of 2 rows]
      uint8_t* dst = a.ar.part_ll + ((size_t)(tag % kFRing) * kFMaxClusters + cluster) * (size_t)(32 * 2048);
#pragma unroll
      for (int f = 0; f < C::kSliceFloats / 4 / 64; ++f) {
        const int idx4 = lt + 64 * f;
        const int sub = idx4 >> 6, qq = idx4 & 63;
        float4 s4 = *reinterpret_cast<const float4*>(rb + (size_t)idx4 * 4);
#pragma unroll
        for (int src = 1; src < kFCS; ++src) {
          const float4 p = *reinterpret_cast<const float4*>(rb + (size_t)src * C::kSliceFloats + (size_t)idx4 * 4);
          s4.x += p.x; s4.y += p.y; s4.z += p.z; s4.w += p.w;
        }
        const int jj = g * kFJobsPerGroup + (int)crank * (kFRowsPerOwner / kFJobRows) + sub;
        uint8_t* o = dst + (size_t)jj * 2048 + (size_t)qq * 32;
        u_ll_store(o, s4.x, s4.y, tag);
        u_ll_store(o + 16, s4.z, s4.w, tag);
      }
      if (warp == 2 && lane == 0) f_trace(a, t, 5);
      // the slot may be refilled (and the senders' staging rewritten) once both warps of every owner have read it
      __syncwarp();
      if (lane < kFCS) u_mbar_arrive_remote(u_mapa(u_smem(&sm.rfree[slot]), (uint32_t)lane));
    }
  } else if (warp >= 8 && warp < 12) {
    // ============================================================ weights -> tensor memory, z, final epilogue
    const int lq = warp & 3;
    const int L = lq * 32 + lane;
    const int part = G == 1 ? (lane >> 4) : 0;
    const int q = G == 1 ? (lq * 16 + (lane & 15)) : L;
    const uint32_t tlane = tmem + ((uint32_t)(lq * 32) << 16);
    float z = 0.f;
#pragma unroll 1
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t seq = epoch0 + (uint32_t)t;
      const int ring = (int)(seq % kFRing);
      const uint32_t tag4 = (seq / kFRing) & 15u;    // the slot's previous occupant carries tag4 - 1
      // published weights: [ring][job 32][q 64][4 rows] fp32 whose low 4 mantissa bits carry the tile's round (seq / ring):
      // every word validates itself, so the jobs need no fence and no flag and the consumers no acquire -- the slot's
      // previous occupant is exactly kFRing tiles older (ring slot = seq mod kFRing, gap-free over launches).
      const uint32_t* wr = reinterpret_cast<const uint32_t*>(a.ar.gw) + (size_t)ring * (32 * 64 * 4);
      const int b = t & 1;
      if constexpr (G == 1) {
        const uint32_t* wq = wr + ((size_t)(part * 8) * 64 + q) * 4;
        {   // one lane probes one word until the tile shows up; the full-width loads below then (nearly) always pass
          uint32_t spin = 0;
          int seen = 0;
          do {
            if (lane == 0) {
              uint32_t w0;
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w0) : "l"(wq) : "memory");
              seen = (w0 & 15u) == tag4;
            }
            seen = __shfl_sync(0xffffffffu, seen, 0);
            if (!seen && ++spin > kFSpinLimit) f_timeout(a, 0x400, (uint32_t)t, tag4);
          } while (!seen);
          if (warp == 8 && lane == 0) {     // the weights are on their way: start the second read of the tile (L2)
            if (t >= C::kSB) f_wait(a, &sm.bempty[t % C::kSB], (uint32_t)(((t / C::kSB) + 1) & 1), 0x121, t);
            load_tile(sm.stages + (size_t)(C::kSA + t % C::kSB) * kFStageBytes, &sm.bfull[t % C::kSB], t, pol_drop);
          }
        }
        uint4 wv[8];
        {
          uint32_t spin = 0;
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              wv[v] = u_ll_load(wq + (size_t)v * 64 * 4);
              ok = ok && ((wv[v].x & 15u) == tag4) && ((wv[v].y & 15u) == tag4) && ((wv[v].z & 15u) == tag4) && ((wv[v].w & 15u) == tag4);
            }
            ok = __all_sync(0xffffffffu, ok);
            if (!ok && ++spin > kFSpinLimit) f_timeout(a, 0x402, (uint32_t)t, tag4);
          } while (!ok);
        }
        if (warp == 8 && lane == 0) f_trace(a, t, 6);
        float own[32], oth[32];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          own[v * 4 + 0] = __uint_as_float(wv[v].x & ~15u); own[v * 4 + 1] = __uint_as_float(wv[v].y & ~15u);
          own[v * 4 + 2] = __uint_as_float(wv[v].z & ~15u); own[v * 4 + 3] = __uint_as_float(wv[v].w & ~15u);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) oth[j] = __shfl_xor_sync(0xffffffffu, own[j], 16);
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {            // rows 0..31 of the tile
          const float k0 = part ? oth[2 * j] : own[2 * j], k1 = part ? oth[2 * j + 1] : own[2 * j + 1];
          z += k0; z += k1;
          const float h0 = f_bf16_hi(k0), h1 = f_bf16_hi(k1);
          pk[j] = part ? f_pack_bf16(k0 - h0, k1 - h1) : f_pack_bf16(h0, h1);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {            // rows 32..63
          const float k0 = part ? own[2 * j] : oth[2 * j], k1 = part ? own[2 * j + 1] : oth[2 * j + 1];
          z += k0; z += k1;
          const float h0 = f_bf16_hi(k0), h1 = f_bf16_hi(k1);
          pk[16 + j] = part ? f_pack_bf16(k0 - h0, k1 - h1) : f_pack_bf16(h0, h1);
        }
        if (t >= 2) { f_wait(a, &sm.pempty[b], (uint32_t)(((t >> 1) + 1) & 1), 0x410, t); u_fence_after(); }
        u_tmem_st32(tlane + C::kColP + (uint32_t)b * C::kPStride, pk);
      } else {
        uint32_t ph[32], pl[32];
        const uint32_t* wq = wr + ((size_t)(q >> 6) * kFJobsPerGroup * 64 + (q & 63)) * 4;
        {
          uint32_t spin = 0;
          int seen = 0;
          do {
            if (lane == 0) {
              uint32_t w0;
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w0) : "l"(wq) : "memory");
              seen = (w0 & 15u) == tag4;
            }
            seen = __shfl_sync(0xffffffffu, seen, 0);
            if (!seen && ++spin > kFSpinLimit) f_timeout(a, 0x400, (uint32_t)t, tag4);
          } while (!seen);
          if (warp == 8 && lane == 0) {     // the weights are on their way: start the second read of the tile (L2)
            if (t >= C::kSB) f_wait(a, &sm.bempty[t % C::kSB], (uint32_t)(((t / C::kSB) + 1) & 1), 0x121, t);
            load_tile(sm.stages + (size_t)(C::kSA + t % C::kSB) * kFStageBytes, &sm.bfull[t % C::kSB], t, pol_drop);
          }
        }
#pragma unroll
        for (int hv = 0; hv < 2; ++hv) {          // rows [32 hv, 32 hv + 32): 8 lines of 4 rows
          uint4 wv[8];
          uint32_t spin = 0;
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              wv[v] = u_ll_load(wq + (size_t)(hv * 8 + v) * 64 * 4);
              ok = ok && ((wv[v].x & 15u) == tag4) && ((wv[v].y & 15u) == tag4) && ((wv[v].z & 15u) == tag4) && ((wv[v].w & 15u) == tag4);
            }
            ok = __all_sync(0xffffffffu, ok);
            if (!ok && ++spin > kFSpinLimit) f_timeout(a, 0x402, (uint32_t)t, tag4);
          } while (!ok);
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float k0 = __uint_as_float(wv[v].x & ~15u), k1 = __uint_as_float(wv[v].y & ~15u);
            const float k2 = __uint_as_float(wv[v].z & ~15u), k3 = __uint_as_float(wv[v].w & ~15u);
            z += k0; z += k1; z += k2; z += k3;
            const float h0 = f_bf16_hi(k0), h1 = f_bf16_hi(k1), h2 = f_bf16_hi(k2), h3 = f_bf16_hi(k3);
            ph[hv * 16 + 2 * v] = f_pack_bf16(h0, h1); ph[hv * 16 + 2 * v + 1] = f_pack_bf16(h2, h3);
            pl[hv * 16 + 2 * v] = f_pack_bf16(k0 - h0, k1 - h1); pl[hv * 16 + 2 * v + 1] = f_pack_bf16(k2 - h2, k3 - h3);
          }
        }
        if (warp == 8 && lane == 0) f_trace(a, t, 6);
        if (t >= 2) { f_wait(a, &sm.pempty[b], (uint32_t)(((t >> 1) + 1) & 1), 0x410, t); u_fence_after(); }
        u_tmem_st32(tlane + C::kColP + (uint32_t)b * C::kPStride, ph);
        u_tmem_st32(tlane + C::kColP + (uint32_t)b * C::kPStride + 32, pl);
      }
      u_tmem_st_wait();
      u_fence_before();
      __syncwarp();
      if (lane == 0) u_mbar_arrive(&sm.pfull[b]);
      if (warp == 8 && lane == 0) f_trace(a, t, 7);
    }

    // ---- final epilogue: num[q][slice] from tensor memory -> shared memory (thread = row) -> whole 512-byte rows
    f_wait(a, sm.accfull, 0, 0x420, ntiles);
    u_fence_after();
    {
      float* er = estage + (size_t)q * kFPitch + part * 16;
#pragma unroll 1
      for (int c = 0; c < kFDS / 32; ++c) {
        uint32_t r[32];
        u_tmem_ld32_nowait(tlane + C::kColAcc + c * 32, r);
        u_tmem_ld_wait();
        if constexpr (G == 1) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float lo16 = __uint_as_float(r[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r[j]), 16);
            const float hi16 = __uint_as_float(r[16 + j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r[16 + j]), 16);
            v[j] = part ? hi16 : lo16;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(er + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(er + c * 32 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                     __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
      if (part == 0) sm.xsq_half[q] = z;
    }
    u_fence_before();
    asm volatile("bar.sync 2, 128;" ::: "memory");
    const int nrows = min(a.Q, C::kQ);
    float msum = 0.f;
#pragma unroll 1
    for (int r = warp - 8; r < nrows; r += 4) {
      const float4 n4 = *reinterpret_cast<const float4*>(estage + (size_t)r * kFPitch + lane * 4);
      const int64_t o = (int64_t)r * a.D + d0 + lane * 4;
      if (a.num_out) *reinterpret_cast<float4*>(a.num_out + o) = n4;
      if (a.epi.fused) {
        const float denom = sm.xsq_half[r] + a.epi.eps;
        const float4 g4 = make_float4(n4.x / denom, n4.y / denom, n4.z / denom, n4.w / denom);
        msum += fminf(fmaxf(g4.x, -1e10f), 1e10f) + fminf(fmaxf(g4.y, -1e10f), 1e10f) +
                fminf(fmaxf(g4.z, -1e10f), 1e10f) + fminf(fmaxf(g4.w, -1e10f), 1e10f);
        if (a.epi.neg_out) *reinterpret_cast<float4*>(a.epi.neg_out + o) = g4;
        if (a.epi.x0) {
          float4 x4 = __ldcg(reinterpret_cast<const float4*>(a.epi.x0 + o));
          x4.x = fmaf(-a.epi.scale, g4.x, x4.x); x4.y = fmaf(-a.epi.scale, g4.y, x4.y);
          x4.z = fmaf(-a.epi.scale, g4.z, x4.z); x4.w = fmaf(-a.epi.scale, g4.w, x4.w);
          *reinterpret_cast<float4*>(a.epi.x0 + o) = x4;
        }
      }
    }
    if (cta == 0) {
      for (int r = (warp - 8) * 32 + lane; r < nrows; r += 128) {
        const float zr = sm.xsq_half[r], denom = zr + a.epi.eps;
        if (a.z_out) a.z_out[r] = zr;
        if (a.epi.fused) {
          if (a.epi.denom_out) a.epi.denom_out[r] = denom;
          if (a.epi.gate_out) a.epi.gate_out[r] = (!(a.epi.flags & SDN_EPI_GATE) || denom > a.epi.gate_thr) ? 1 : 0;
        }
      }
    }
    if (a.epi.fused && a.epi.mean_out) {
      msum = warp_sum(msum);
      if (lane == 0) atomicAdd(a.epi.mean_out, msum * a.epi.inv_qd);
    }
  } else if (warp >= 12) {
    // ============================================================ level-2 jobs: ||x||^2, sum over clusters, exp, publish
    const int ow = warp - 12;
    const int tid = ow * 32 + lane;                    // 0..127
    const int nctas = gridDim.x;
    // ---- ||x_q||^2 = sum over the CTAs' slices, fixed order
    {
      const int qx = G == 1 ? (tid & 63) : tid;
      const int half = G == 1 ? (tid >> 6) : 0;
      const int per = G == 1 ? nctas / 2 : nctas;
      const int j0 = half * per;
      float acc = 0.f;
      for (int jb = 0; jb < per; jb += 8) {
        uint4 ln[8];
        uint32_t spin = 0;
        bool ok;
        do {
          ok = true;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = min(j0 + jb + u, nctas - 1);
            ln[u] = u_ll_load(a.ar.xsq_ll + ((size_t)j * 128 + qx) * 16);
            ok = ok && ln[u].y == launch0 && ln[u].w == launch0;
          }
          if (!ok) {
            __nanosleep(128);
            if (++spin > kFSpinLimit) f_timeout(a, 0x500, (uint32_t)jb, launch0);
          }
        } while (!ok);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (jb + u < per) acc += __uint_as_float(ln[u].x);
      }
      if constexpr (G == 1) {
        sm.xsq_half[tid] = acc;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid < 64) sm.xsq[tid] = sm.xsq_half[tid] + sm.xsq_half[64 + tid];
      } else {
        sm.xsq[tid] = acc;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // ---- level-2 work units (tile t, query group g, rows [4j, 4j+4), half h of the group's 64 query rows) dealt
    //      round-robin over all 4 x #CTAs warps of this role: every warp is an independent worker (a worker blocks on
    //      its unit until the slowest cluster has delivered, so the tiles between two units of one worker bound the
    //      rate: 16 G tiles here).  Lane l reads lines l and 32 + l of the unit's 1 KiB block of every cluster (512
    //      contiguous bytes per warp instruction) = (query row, row pair) twice, and sums the clusters in order.
    {
      const int lh = lane & 1;
      const int nunits = ntiles * C::kJobs * 2;
      const int nworkers = nctas * 4;
#pragma unroll 1
      for (int U = cta * 4 + ow; U < nunits; U += nworkers) {
        const int t = U / (C::kJobs * 2), rem = U - t * (C::kJobs * 2);
        const int jj = rem >> 1, qh = rem & 1;
        const int j = jj & (kFJobsPerGroup - 1), g = jj / kFJobsPerGroup;
        const uint32_t tag = epoch0 + (uint32_t)t;
        const int ring = (int)(tag % kFRing);
        const int row = t * kFR + j * kFJobRows + lh * 2;
        // ||n||^2 of the two rows: issued before the wait, the line is cold
        const float sq0 = row < a.N ? __ldg(a.sqnorm + row) : 0.f;
        const float sq1 = row + 1 < a.N ? __ldg(a.sqnorm + row + 1) : 0.f;
        const size_t cl_stride = (size_t)32 * 2048;
        const uint8_t* src0 = a.ar.part_ll + (size_t)ring * kFMaxClusters * cl_stride + (size_t)jj * 2048 +
                              (size_t)(qh * 64 + lane) * 16;
        // A unit usually arrives long before its tile: ONE lane probes one line, with back-off, until the tile shows up
        // (every worker spinning on full-width loads saturates the L1s and L2)
        {
          uint32_t spin = 0;
          int seen = 0;
          do {
            if (lane == 0) {
              const uint4 l0 = u_ll_load(src0);
              seen = (l0.y == tag && l0.w == tag) ? 1 : 0;
              if (!seen) __nanosleep(a.poll_sleep);
            }
            seen = __shfl_sync(0xffffffffu, seen, 0);
            if (!seen && ++spin > kFSpinLimit) f_timeout(a, 0x508, (uint32_t)t, (uint32_t)rem);
          } while (!seen);
          if (ow == 0 && lane == 0) f_trace(a, t, 9);
        }
        float da0 = 0.f, da1 = 0.f, db0 = 0.f, db1 = 0.f;      // lines l (a) and 32 + l (b), rows lh*2 and lh*2+1
        for (int k0 = 0; k0 < nclusters; k0 += 8) {
          uint4 la[8], lb[8];
          uint32_t spin = 0;
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int k = min(k0 + u, nclusters - 1);
              la[u] = u_ll_load(src0 + (size_t)k * cl_stride);
              lb[u] = u_ll_load(src0 + (size_t)k * cl_stride + 512);
              ok = ok && la[u].y == tag && la[u].w == tag && lb[u].y == tag && lb[u].w == tag;
            }
            if (!ok && ++spin > kFSpinLimit) f_timeout(a, 0x510, (uint32_t)t, (uint32_t)(rem * 256 + k0));
          } while (!ok);
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (k0 + u < nclusters) {
              da0 += __uint_as_float(la[u].x); da1 += __uint_as_float(la[u].z);
              db0 += __uint_as_float(lb[u].x); db1 += __uint_as_float(lb[u].z);
            }
        }
        if (ow == 0 && lane == 0) f_trace(a, t, 10);
        const uint32_t tagbits = (tag / kFRing) & 15u;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int q64 = qh * 32 + i * 16 + (lane >> 1);
          const int q = g * 64 + q64;
          const float d0v = i ? db0 : da0, d1v = i ? db1 : da1;
          const float xs = sm.xsq[q];
          float k0v = 0.f, k1v = 0.f;
          if (q < a.Q) {
            if (row < a.N) k0v = expf(-dist_from_dot(xs, sq0, d0v, a.alpha, a.power) * a.inv2s2);
            if (row + 1 < a.N) k1v = expf(-dist_from_dot(xs, sq1, d1v, a.alpha, a.power) * a.inv2s2);
          }
          // the weights carry the tile's round (seq / ring) in their low 4 mantissa bits (2^-19 relative): no fence, no
          // flag -- each 32-bit word is valid on its own
          const uint32_t b0 = (__float_as_uint(k0v) & ~15u) | tagbits, b1 = (__float_as_uint(k1v) & ~15u) | tagbits;
          asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};"
                       ::"l"(a.ar.gw + (((size_t)ring * 32 + jj) * 64 + q64) * 4 + lh * 2), "r"(b0), "r"(b1) : "memory");
          if (a.k_out && q < a.Q) {
            if (row < a.N) a.k_out[(int64_t)q * a.N + row] = k0v;
            if (row + 1 < a.N) a.k_out[(int64_t)q * a.N + row + 1] = k1v;
          }
        }
        if (ow == 0 && lane == 0) f_trace(a, t, 11);
      }
    }
  }

  // no CTA of the cluster leaves while a peer may still store into its shared memory or arrive on its barriers
  u_fence_before();
  __syncthreads();
  u_cluster_arrive();
  u_cluster_wait();
  if (warp == 1) {
    u_fence_after();
    u_tmem_dealloc(tmem, kFTmemCols);
  }
  // every CTA has read the tag base long before any CTA can get here (each contributed to the last tile)
  if (cta == 0 && threadIdx.x == 0) {
    *reinterpret_cast<volatile uint32_t*>(a.ar.epoch) = epoch0 + (uint32_t)ntiles;      // gap-free: ring slot = seq mod kFRing
    *reinterpret_cast<volatile uint32_t*>(a.ar.epoch + 1) = launch0 + 1u;
  }
}

// ------------------------------------------------------------------------------------------ host side
namespace {

struct ArenaHost {
  std::mutex mu;
  bool ready = false;
  void* dev = nullptr;
  uint32_t* diag_host = nullptr;
  FlashArena ar{};
  int max_clusters[3] = {0, 0, 0};     // co-resident clusters of k_flash<G>, indexed by G
  bool coop_ok = true;
};
ArenaHost g_arena[kMaxDevices];

constexpr size_t kXsqBytes = (size_t)kFMaxCtas * 128 * 16;
constexpr size_t kPartBytes = (size_t)kFRing * kFMaxClusters * 32 * 2048;
constexpr size_t kGwBytes = (size_t)kFRing * 32 * 64 * 4 * 4;
constexpr size_t kFlagBytes = (size_t)kFRing * 128 * 4;

int arena_get(int dev, ArenaHost** out) {
  ArenaHost& h = g_arena[dev];
  std::lock_guard<std::mutex> lk(h.mu);
  if (!h.ready) {
    const size_t total = 256 + kXsqBytes + kPartBytes + kGwBytes + 256 + kFlagBytes;
    SDN_CUDA_OK(cudaMalloc(&h.dev, total));
    SDN_CUDA_OK(cudaMemset(h.dev, 0, total));
    uint8_t* p = static_cast<uint8_t*>(h.dev);
    h.ar.epoch = reinterpret_cast<uint32_t*>(p); p += 256;
    h.ar.xsq_ll = p; p += kXsqBytes;
    h.ar.part_ll = p; p += kPartBytes;
    h.ar.gw = reinterpret_cast<float*>(p); p += kGwBytes;
    p += 256;
    h.ar.wflag = reinterpret_cast<uint32_t*>(p);
    const uint32_t one[2] = {kFRing, 1};     // sequence numbers start in round 1: the zeroed rings (round-tag 0, LL tag 0) never match
    SDN_CUDA_OK(cudaMemcpy(h.ar.epoch, one, sizeof(one), cudaMemcpyHostToDevice));
    void* dh = nullptr;
    if (cudaHostAlloc(&dh, 64, cudaHostAllocMapped) == cudaSuccess) {
      memset(dh, 0, 64);
      void* dd = nullptr;
      if (cudaHostGetDevicePointer(&dd, dh, 0) == cudaSuccess) {
        h.diag_host = static_cast<uint32_t*>(dh);
        h.ar.diag = static_cast<uint32_t*>(dd);
      }
    }
    if (getenv("SDN_FLASH_TRACE")) {
      void* tr = nullptr;
      const size_t tb = (size_t)kFMaxCtas * kFTraceTiles * kFTraceEvents * 8;
      if (cudaMalloc(&tr, tb) == cudaSuccess) { cudaMemset(tr, 0, tb); h.ar.trace = static_cast<unsigned long long*>(tr); }
    }
    cudaGetLastError();
    SDN_CUDA_OK(cudaFuncSetAttribute(k_flash<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCfg<1>::kSmemBytes));
    SDN_CUDA_OK(cudaFuncSetAttribute(k_flash<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FCfg<2>::kSmemBytes));
    for (int G = 1; G <= 2; ++G) {
      cudaLaunchConfig_t probe{};
      probe.gridDim = dim3(kFMaxCtas);
      probe.blockDim = dim3(kFThreads);
      probe.dynamicSmemBytes = G == 1 ? FCfg<1>::kSmemBytes : FCfg<2>::kSmemBytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = kFCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      probe.attrs = at; probe.numAttrs = 1;
      int n = 0;
      const cudaError_t e = G == 1 ? cudaOccupancyMaxActiveClusters(&n, k_flash<1>, &probe)
                                   : cudaOccupancyMaxActiveClusters(&n, k_flash<2>, &probe);
      if (getenv("SDN_FLASH_DEBUG")) fprintf(stderr, "[sdn_flash] G=%d occupancy query: %s, max active clusters of %d = %d\n", G, cudaGetErrorString(e), kFCS, n);
      if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
      h.max_clusters[G] = n;
    }
    h.ready = true;
  }
  *out = &h;
  return SDN_OK;
}

// tensor maps of the bank planes: [N][D] bf16, box [64 rows][64 d]; a few entries per process, keyed on the bank
struct FlashMaps { int dev; const void* planes; int64_t N, D; CUtensorMap hi, lo; uint64_t stamp; };
std::mutex g_maps_mu;
FlashMaps g_maps[8];
int g_maps_n = 0;
uint64_t g_maps_clock = 0;

int maps_get(int dev, const void* planes, int64_t N, int64_t D, CUtensorMap* hi, CUtensorMap* lo) {
  std::lock_guard<std::mutex> lk(g_maps_mu);
  for (int i = 0; i < g_maps_n; ++i) {
    FlashMaps& m = g_maps[i];
    if (m.dev == dev && m.planes == planes && m.N == N && m.D == D) {
      m.stamp = ++g_maps_clock; *hi = m.hi; *lo = m.lo;
      return SDN_OK;
    }
  }
  int slot = g_maps_n < 8 ? g_maps_n++ : 0;
  if (slot == 0 && g_maps_n == 8)
    for (int i = 1; i < 8; ++i) if (g_maps[i].stamp < g_maps[slot].stamp) slot = i;
  FlashMaps& m = g_maps[slot];
  const __nv_bfloat16* h = static_cast<const __nv_bfloat16*>(planes);
  int rc;
  if ((rc = tmap_bf16_2d(&m.hi, h, (uint64_t)N, (uint64_t)D, kFR, 64))) { g_maps_n = std::min(g_maps_n, slot); return rc; }
  if ((rc = tmap_bf16_2d(&m.lo, h + N * D, (uint64_t)N, (uint64_t)D, kFR, 64))) { g_maps_n = std::min(g_maps_n, slot); return rc; }
  m.dev = dev; m.planes = planes; m.N = N; m.D = D; m.stamp = ++g_maps_clock;
  *hi = m.hi; *lo = m.lo;
  return SDN_OK;
}

template <int G>
int launch_flash(const CUtensorMap& hi, const CUtensorMap& lo, const FlashArgs& a, int nctas, bool coop, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nctas);
  cfg.blockDim = dim3(kFThreads);
  cfg.dynamicSmemBytes = FCfg<G>::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kFCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = coop ? 2 : 1;
  return (int)cudaLaunchKernelEx(&cfg, k_flash<G>, hi, lo, a);
}

}  // namespace

bool flash_shape_ok(int64_t Q, int64_t N, int64_t D) {
  if (Q < 1 || N < 1 || N >= (1ll << 30) || D % 1024) return false;   // an even number of clusters
  const int64_t nctas = D / kFDS;
  return nctas >= 64 && nctas <= kFMaxCtas;      // fewer CTAs cannot keep enough bytes in flight (two-phase path wins)
}

bool flash_supported(int64_t Q, int64_t N, int64_t D, const void* planes) {
  if (!planes || !flash_shape_ok(Q, N, D)) return false;
  static const bool off = [] { const char* e = getenv("SDN_FLASH_OFF"); return e && atoi(e) != 0; }();
  return !off;
}

int flash_diag_read(uint32_t* out, int n) {
  ArenaHost& h = g_arena[device_slot()];
  if (!h.ready || !h.diag_host) return 0;
  for (int i = 0; i < n && i < 16; ++i) out[i] = h.diag_host[i];
  return 1;
}

size_t flash_trace_read(void* host_out, size_t bytes) {
  ArenaHost& h = g_arena[device_slot()];
  const size_t tb = (size_t)kFMaxCtas * kFTraceTiles * kFTraceEvents * 8;
  if (!h.ready || !h.ar.trace || bytes < tb) return 0;
  if (cudaMemcpy(host_out, h.ar.trace, tb, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return tb;
}

// One pass over the bank per <= 128 query rows.  epi == nullptr: partial sums (num_out, z_out).
int flash_run(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq, int64_t Q,
              float inv2s2, int power, float alpha, float* num_out, float* z_out, float* k_out,
              const FlashEpi* epi, cudaStream_t st) {
  if (!flash_supported(Q, N, D, planes)) return SDN_E_UNSUPPORTED;
  const int dev = device_slot();
  ArenaHost* h = nullptr;
  int rc = arena_get(dev, &h);
  if (rc) return rc;
  const int nctas = (int)(D / kFDS);
  const int nclusters = nctas / kFCS;
  CUtensorMap hi, lo;
  if ((rc = maps_get(dev, planes, N, D, &hi, &lo))) return rc;
  for (int64_t q0 = 0; q0 < Q; q0 += 128) {
    const int qn = (int)std::min<int64_t>(128, Q - q0);
    const int G = qn > 64 ? 2 : 1;
    if (h->max_clusters[G] < nclusters) return SDN_E_UNSUPPORTED;    // the grid must be co-resident
    FlashArgs a{};
    a.xq = xq + q0 * D; a.sqnorm = sqnorm; a.Q = qn; a.N = (int)N; a.ntiles = (int)cdiv(N, kFR); a.nclusters = nclusters;
    static const int window = [] { const char* e = getenv("SDN_FLASH_WINDOW"); const int v = e ? atoi(e) : 0; return v >= 1 && v < kFRing ? v : kFWindow; }();
    static const int mma_sleep = [] { const char* e = getenv("SDN_FLASH_MMA_SLEEP"); return e ? atoi(e) : 40; }();
    static const int poll_sleep = [] { const char* e = getenv("SDN_FLASH_POLL_SLEEP"); return e ? atoi(e) : 100; }();
    a.mma_sleep = (unsigned)mma_sleep; a.poll_sleep = (unsigned)poll_sleep;
    a.window = window; a.D = D; a.inv2s2 = inv2s2; a.alpha = alpha; a.power = power;
    a.num_out = num_out ? num_out + q0 * D : nullptr;
    a.z_out = z_out ? z_out + q0 : nullptr;
    a.k_out = k_out ? k_out + q0 * N : nullptr;
    if (epi) {
      a.epi = *epi;
      a.epi.fused = 1;
      a.epi.x0 = epi->x0 ? epi->x0 + q0 * D : nullptr;
      a.epi.neg_out = epi->neg_out ? epi->neg_out + q0 * D : nullptr;
      a.epi.denom_out = epi->denom_out ? epi->denom_out + q0 : nullptr;
      a.epi.gate_out = epi->gate_out ? epi->gate_out + q0 : nullptr;
      a.epi.zero_mean = q0 == 0 ? 1 : 0;
    }
    a.ar = h->ar;
    const int pid = g_prof.begin("k_flash", st);
    int e = G == 1 ? launch_flash<1>(hi, lo, a, nctas, h->coop_ok, st) : launch_flash<2>(hi, lo, a, nctas, h->coop_ok, st);
    if (e != 0 && h->coop_ok) {
      // cooperative + cluster launch refused: co-residency is still guaranteed by the occupancy check above as long
      // as nothing else runs on the device
      cudaGetLastError();
      h->coop_ok = false;
      e = G == 1 ? launch_flash<1>(hi, lo, a, nctas, false, st) : launch_flash<2>(hi, lo, a, nctas, false, st);
    }
    g_prof.end(pid, st);
    if (e != 0) return e;
    SDN_LAUNCHED();
  }
  return SDN_OK;
}

}  // namespace sdn
