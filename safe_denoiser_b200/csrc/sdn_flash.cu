// One-pass "flash-repellency" for batched calls: the bank is read from HBM exactly once.
//
// Replaces repellency_methods_fast.py:249-250 (cdist + the [Q,N,D+1] broadcast) for 8 < Q <= 128 query rows per pass.
// The two-phase tcgen05 path (sdn_umma.cu) streams the bank twice because the weights of a bank row need its dot
// product over ALL of D before the row can be accumulated.  Here the row tile stays in shared memory between the
// two contractions instead:
//
//   grid   = D / 128 CTAs (one per SM, clusters of 4); CTA j owns the d-slice [128 j, 128 j + 128) of every bank row
//            and keeps  X[:, slice]  (bf16 hi/lo, tcgen05 A operand)  and  num[:, slice]  (fp32 accumulator) in
//            TENSOR MEMORY for the whole kernel;
//   tile   = 64 bank rows x the slice, hi and lo planes (32 KiB) -> one shared-memory stage filled by TMA;
//   phase A  S_j[q][i] = sum_{d in slice} X[q][d] bank[i][d]      tcgen05.mma, A = X (TMEM), B = tile (K-major)
//   reduce   S[q][i] = sum_j S_j[q][i] over the D/128 CTAs:
//              level 1  inside the cluster over distributed shared memory (st.async + mbarrier complete_tx):
//                       CTA c of a cluster receives and sums rows [16c, 16c+16) of the tile;
//              level 2  the cluster partials go through L2 as self-validating 16-byte lines {v, tag, v, tag}
//                       (no fence, no flag) to "jobs" of 4 rows x 64 queries dealt round-robin over ALL CTAs: the
//                       job's CTA sums the 32 cluster partials in cluster order, computes k = exp(-dist / 2 sigma^2)
//                       and publishes the weights into [Q][64] (+ one release flag per warp);
//   phase B  num[q][slice] += sum_i k[q][i] bank[i][slice]        tcgen05.mma, A = weights (bf16 hi/lo written to
//            TMEM by tcgen05.st), B = THE SAME shared-memory tile viewed MN-major; its commit frees the stage.
//
// Every sum has a fixed order (CTA rank, cluster index, row index): results are bit-reproducible run to run.
// A stage lives from its TMA issue to the commit of phase B, i.e. through the cross-CTA reduction (a few us), so
// the achievable HBM rate is (stages x 32 KiB x CTAs) / that latency; everything in the exchange is built to keep
// it short (no fences on the fan-in, DSMEM for the first 8:1).
//
// Warp roles (16 warps, 1 CTA per SM): 0 TMA producer | 1 MMA issuer (event loop over "phase A of tile ta ready" /
// "weights of tile tb ready") | 4-7 phase-A drain: TMEM -> registers -> DSMEM scatter, then the level-1 sum |
// 8-11 weights: global -> bf16 hi/lo -> TMEM, z_q, and the final epilogue | 12-15 tile owner: ||x||^2, level-2 sum,
// exp, publish.
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
constexpr int kFRing = 8;               // tiles in flight in the global exchange rings (> stages)
constexpr int kFThreads = 512;
constexpr int kFMaxCtas = 128;
constexpr int kFMaxClusters = kFMaxCtas / kFCS;
constexpr uint32_t kFStageBytes = 2u * kFR * kFDS * 2u;   // hi + lo tiles of [64 rows][128 d] bf16 = 32 KiB
constexpr uint32_t kFTmemCols = 512;
constexpr uint32_t kFColX = 0;          // X operand: G = 1: 64 columns (stacked hi|lo rows); G = 2: hi 64 | lo 64
constexpr uint32_t kFColS = 128;        // S accumulators, 2 x 64 columns
constexpr uint32_t kFColP = 256;        // weight operands, 2 x 64 columns (G = 2: hi 32 | lo 32)
constexpr uint32_t kFColAcc = 384;      // num accumulator, 128 columns
constexpr uint32_t kFSpinLimit = 1u << 24;

template <int G>
struct FCfg {
  static constexpr int kQ = 64 * G;                                   // query rows per pass
  static constexpr int kStages = G == 1 ? 6 : 5;
  static constexpr int kSliceFloats = kQ * kFRowsPerOwner;            // one (tile, level-1 owner) slice: [kQ][rows per owner]
  static constexpr int kSliceVec = kSliceFloats / 4 / 128;            // float4 per thread of the level-1 sum
  static constexpr uint32_t kRbufSlotBytes = (uint32_t)kFCS * kSliceFloats * 4;   // [src][kQ][rows per owner]
  static constexpr int kJobs = kFJobsPerGroup * G;                    // level-2 jobs per tile
  static constexpr int kXStages = (kQ * 132 * 4 + (int)kFStageBytes - 1) / (int)kFStageBytes;   // stages the query staging covers
  static constexpr uint32_t kSmemBytes = kStages * kFStageBytes + 2 * kRbufSlotBytes + 1024 /*align*/ + 1536 /*barriers, xsq*/;
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
  int Q, N, ntiles, nclusters;
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

struct FSmem {
  uint8_t* stages;
  float* rbuf;          // [2][kFCS][slice floats]
  uint64_t* full; uint64_t* empty;        // [stages]
  uint64_t* sfull; uint64_t* sempty;      // [2] S accumulators
  uint64_t* pfull; uint64_t* pempty;      // [2] weight operands
  uint64_t* rfull; uint64_t* rfree;       // [2] level-1 receive slots
  uint64_t* xfull; uint64_t* accfull;     // [1]
  uint64_t* wready;                       // [kFRing] weights of tile t published (signalled by the cluster's rank-0 CTA)
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
  uint64_t* b = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s.rbuf) + 2 * C::kRbufSlotBytes);
  s.full = b; b += 8;
  s.empty = b; b += 8;
  s.sfull = b; b += 2;
  s.sempty = b; b += 2;
  s.pfull = b; b += 2;
  s.pempty = b; b += 2;
  s.rfull = b; b += 2;
  s.rfree = b; b += 2;
  s.xfull = b; b += 1;
  s.accfull = b; b += 1;
  s.wready = b; b += kFRing;
  s.xload = b; b += 4;
  s.tmem_base = reinterpret_cast<uint32_t*>(b); b += 1;
  s.xsq = reinterpret_cast<float*>(b);          // 43 x 8 = 344 bytes of barriers so far
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
    for (int s = 0; s < C::kStages; ++s) { u_mbar_init(&sm.full[s], 1); u_mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      u_mbar_init(&sm.sfull[b], 1); u_mbar_init(&sm.sempty[b], 4);
      u_mbar_init(&sm.pfull[b], 4); u_mbar_init(&sm.pempty[b], 1);
      u_mbar_init(&sm.rfull[b], 1); u_mbar_init(&sm.rfree[b], kFCS * 4);
    }
    u_mbar_init(sm.xfull, 4); u_mbar_init(sm.accfull, 1);
    for (int w = 0; w < 4; ++w) u_mbar_init(&sm.xload[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // arm the two level-1 receive slots for their first tiles
    u_mbar_expect_tx(&sm.rfull[0], C::kRbufSlotBytes);
    u_mbar_expect_tx(&sm.rfull[1], C::kRbufSlotBytes);
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

  auto load_stage = [&](int t) {
    const int s = t % C::kStages;
    uint8_t* st = sm.stages + (size_t)s * kFStageBytes;
    u_mbar_expect_tx(&sm.full[s], kFStageBytes);
    const int r0 = t * kFR;
    f_trace(a, t, 0);
    u_tma_2d(st, &tm_hi, d0, r0, &sm.full[s]);
    u_tma_2d(st + 8192, &tm_hi, d0 + 64, r0, &sm.full[s]);
    u_tma_2d(st + 16384, &tm_lo, d0, r0, &sm.full[s]);
    u_tma_2d(st + 24576, &tm_lo, d0 + 64, r0, &sm.full[s]);
  };
  // the first stages depend on nothing but this CTA's own barriers: start the stream before the cluster barrier
  const int npre = min(ntiles, C::kStages - C::kXStages);
  if (threadIdx.x == 0)
    for (int t = 0; t < npre; ++t) load_stage(t);

  // every CTA of the cluster has initialised its barriers before any peer stores into its shared memory
  u_cluster_arrive();
  u_cluster_wait();

  if (warp == 0) {
    // ============================================================ TMA producer
    if (lane == 0) {
      if (ntiles > npre) f_wait(a, sm.xfull, 0, 0x110, 0);      // the query staging area becomes pipeline stages
      for (int t = npre; t < ntiles; ++t) {
        if (t >= C::kStages) f_wait(a, &sm.empty[t % C::kStages], (uint32_t)(((t / C::kStages) + 1) & 1), 0x100, t);
        load_stage(t);
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
        if (tb < ta && u_mbar_test(&sm.pfull[tb & 1], (uint32_t)((tb >> 1) & 1))) {
          // ---- phase B of tile tb: 4 K steps of 16 bank rows
          u_fence_after();
          f_trace(a, tb, 8);
          const uint32_t base = u_smem(sm.stages + (size_t)(tb % C::kStages) * kFStageBytes);
          const uint32_t acc = tmem + kFColAcc;
          const uint32_t pb = tmem + kFColP + (uint32_t)(tb & 1) * 64;
#pragma unroll
          for (int kk = 0; kk < kFR / 16; ++kk) {
            const uint64_t bh = u_desc(base + kk * 2048, 8192, 1024);
            const uint64_t bl = u_desc(base + 16384 + kk * 2048, 8192, 1024);
            const uint32_t first = (tb > 0 || kk > 0) ? 1u : 0u;
            u_mma_ts(acc, pb + kk * 8, bh, idB, first);
            u_mma_ts(acc, pb + kk * 8, bl, idB, 1u);
            if constexpr (G == 2) u_mma_ts(acc, pb + 32 + kk * 8, bh, idB, 1u);
          }
          u_commit(&sm.empty[tb % C::kStages]);
          u_commit(&sm.pempty[tb & 1]);
          if (tb == ntiles - 1) u_commit(sm.accfull);
          ++tb;
          did = true;
        }
        if (ta < ntiles && u_mbar_test(&sm.full[ta % C::kStages], (uint32_t)((ta / C::kStages) & 1)) &&
            (ta < 2 || u_mbar_test(&sm.sempty[ta & 1], (uint32_t)(((ta >> 1) + 1) & 1)))) {
          // ---- phase A of tile ta: 8 K steps of 16 d
          u_fence_after();
          f_trace(a, ta, 1);
          const uint32_t base = u_smem(sm.stages + (size_t)(ta % C::kStages) * kFStageBytes);
          const uint32_t acc = tmem + kFColS + (uint32_t)(ta & 1) * 64;
#pragma unroll
          for (int kk = 0; kk < kFDS / 16; ++kk) {
            const uint32_t off = (uint32_t)(kk >> 2) * 8192 + (uint32_t)(kk & 3) * 32;
            const uint64_t bh = u_desc(base + off, 16, 1024);
            const uint64_t bl = u_desc(base + 16384 + off, 16, 1024);
            u_mma_ts(acc, tmem + kFColX + kk * 8, bh, idA, kk > 0 ? 1u : 0u);
            u_mma_ts(acc, tmem + kFColX + kk * 8, bl, idA, 1u);
            if constexpr (G == 2) u_mma_ts(acc, tmem + kFColX + 64 + kk * 8, bh, idA, 1u);
          }
          u_commit(&sm.sfull[ta & 1]);
          ++ta;
          did = true;
        }
        if (did) idle = 0;
        else {
          // nothing ready: sleep in hardware on the barrier most likely to fire next instead of spinning on probes
          if (tb < ta) u_mbar_try(&sm.pfull[tb & 1], (uint32_t)((tb >> 1) & 1), 200);
          else if (ta < ntiles) u_mbar_try(&sm.full[ta % C::kStages], (uint32_t)((ta / C::kStages) & 1), 200);
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
          u_tmem_st16(tlane + kFColX + c4 * 16, sel);
        } else {
          u_tmem_st16(tlane + kFColX + c4 * 16, hi);
          u_tmem_st16(tlane + kFColX + 64 + c4 * 16, lo);
        }
      }
      u_tmem_st_wait();
      if (part == 0)
        u_ll_store(a.ar.xsq_ll + ((size_t)cta * 128 + q) * 16, ss, 0.f, launch0);
      u_fence_before();
      __syncwarp();
      if (lane == 0) u_mbar_arrive(sm.xfull);
    }

    // Level-1 slice of CTA c of the cluster: rows [16c, 16c+16) of the tile for all query rows, stored as
    // [sub 4][q kQ][4 rows]; float4 index = sub * kQ + q.
    const uint32_t rbuf_s = u_smem(sm.rbuf);
    auto l1sum = [&](int u) {
      const int slot = u & 1;
      f_wait(a, &sm.rfull[slot], (uint32_t)((u >> 1) & 1), 0x300, u);
      if (warp == 4 && lane == 0 && u + 2 < ntiles) u_mbar_expect_tx(&sm.rfull[slot], C::kRbufSlotBytes);   // next use
      if (warp == 4 && lane == 0) f_trace(a, u, 4);
      const uint32_t tag = epoch0 + (uint32_t)u;            // the tile's sequence number (gap-free over launches)
      const float* rb = sm.rbuf + (size_t)slot * (C::kRbufSlotBytes / 4);
      // cluster partial as LL lines, laid out per level-2 job: [job 32][q 64][2 lines of 2 rows]
      uint8_t* dst = a.ar.part_ll + ((size_t)(tag % kFRing) * kFMaxClusters + cluster) * (size_t)(32 * 2048);
#pragma unroll
      for (int f = 0; f < C::kSliceVec; ++f) {
        const int idx4 = L + 128 * f;
        const int sub = idx4 / C::kQ, qq = idx4 - sub * C::kQ;
        float4 s4 = *reinterpret_cast<const float4*>(rb + (size_t)idx4 * 4);
#pragma unroll
        for (int src = 1; src < kFCS; ++src) {
          const float4 p = *reinterpret_cast<const float4*>(rb + (size_t)src * C::kSliceFloats + (size_t)idx4 * 4);
          s4.x += p.x; s4.y += p.y; s4.z += p.z; s4.w += p.w;
        }
        const int jj = (qq >> 6) * kFJobsPerGroup + (int)crank * (kFRowsPerOwner / kFJobRows) + sub;
        uint8_t* o = dst + (size_t)jj * 2048 + (size_t)(qq & 63) * 32;
        u_ll_store(o, s4.x, s4.y, tag);
        u_ll_store(o + 16, s4.z, s4.w, tag);
      }
      if (warp == 4 && lane == 0) f_trace(a, u, 5);
      // the slot may be refilled once every warp of every owner has read it
      __syncwarp();
      if (lane < kFCS) u_mbar_arrive_remote(u_mapa(u_smem(&sm.rfree[slot]), (uint32_t)lane));
    };
    // drain S(td) -> DSMEM scatter, and the level-1 sum of tile ts as soon as its slot is complete: an event loop per
    // warp (lane 0 probes, the warp follows), older tile first -- a sum must not wait behind the next tile's HBM data
    auto drain = [&](int t) {
      const int b = t & 1;
      f_wait(a, &sm.sfull[b], (uint32_t)((t >> 1) & 1), 0x310, t);
      u_fence_after();
      if (warp == 4 && lane == 0) f_trace(a, t, 2);
      uint32_t r0[32], r1[32];
      u_tmem_ld32_nowait(tlane + kFColS + (uint32_t)b * 64, r0);
      u_tmem_ld32_nowait(tlane + kFColS + (uint32_t)b * 64 + 32, r1);
      u_tmem_ld_wait();
      u_fence_before();
      __syncwarp();
      if (lane == 0) u_mbar_arrive(&sm.sempty[b]);
      const int slot = t & 1;
      if (t >= 2) f_wait(a, &sm.rfree[slot], (uint32_t)(((t >> 1) + 1) & 1), 0x320, t);   // (probed complete by lane 0)
      // address of (sub 0, q) of this sender's slice in the owner's slot; sub s is kQ * 16 bytes further
      const uint32_t my_off = ((uint32_t)slot * (C::kRbufSlotBytes / 4) + crank * (uint32_t)C::kSliceFloats + (uint32_t)q * 4u) * 4u;
      const uint32_t bar_local = u_smem(&sm.rfull[slot]);
      constexpr int kSubs = kFRowsPerOwner / 4;
      if constexpr (G == 1) {
        // lanes l and l+16 hold the hi-part and lo-part rows of the same query: add them, then each sends one half
        float s[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v0 = __uint_as_float(r0[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r0[j]), 16);
          const float v1 = __uint_as_float(r1[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r1[j]), 16);
          s[j] = part ? v1 : v0;
        }
        constexpr int kOwnersPerHalf = 32 / kFRowsPerOwner;
#pragma unroll
        for (int oo = 0; oo < kOwnersPerHalf; ++oo) {
          const uint32_t owner = (uint32_t)(part * kOwnersPerHalf + oo);
          const uint32_t ra = u_mapa(rbuf_s + my_off, owner), rbar = u_mapa(bar_local, owner);
#pragma unroll
          for (int v = 0; v < kSubs; ++v) {
            const int e = oo * kFRowsPerOwner + 4 * v;
            u_st_async_v4(ra + (uint32_t)v * (C::kQ * 16), rbar, s[e], s[e + 1], s[e + 2], s[e + 3]);
          }
        }
      } else {
#pragma unroll
        for (int oo = 0; oo < kFCS; ++oo) {
          const uint32_t ra = u_mapa(rbuf_s + my_off, (uint32_t)oo), rbar = u_mapa(bar_local, (uint32_t)oo);
#pragma unroll
          for (int v = 0; v < kSubs; ++v) {
            const int e = oo * kFRowsPerOwner + 4 * v;       // row of the tile, compile-time
            const uint32_t* r = e < 32 ? r0 : r1;
            const int o = e & 31;
            u_st_async_v4(ra + (uint32_t)v * (C::kQ * 16), rbar, __uint_as_float(r[o]), __uint_as_float(r[o + 1]),
                          __uint_as_float(r[o + 2]), __uint_as_float(r[o + 3]));
          }
        }
      }
      if (warp == 4 && lane == 0) f_trace(a, t, 3);
    };
    int td = 0, ts = 0;
    uint32_t idle = 0;
#pragma unroll 1
    while (ts < ntiles) {
      int act = 0;
      if (lane == 0) {
        if (ts < td && u_mbar_test(&sm.rfull[ts & 1], (uint32_t)((ts >> 1) & 1))) act = 1;
        // a slot is refilled only after this warp's own sum of its previous tile (td - ts < 2) and after every owner
        // has released it (rfree); both are probed here so that a pending sum is never stuck behind a blocked drain
        else if (td < ntiles && td - ts < 2 && u_mbar_test(&sm.sfull[td & 1], (uint32_t)((td >> 1) & 1)) &&
                 (td < 2 || u_mbar_test(&sm.rfree[td & 1], (uint32_t)(((td >> 1) + 1) & 1)))) act = 2;
        if (!act) {   // sleep in hardware on the more likely event
          if (ts < td) u_mbar_try(&sm.rfull[ts & 1], (uint32_t)((ts >> 1) & 1), 150);
          else if (td < ntiles) u_mbar_try(&sm.sfull[td & 1], (uint32_t)((td >> 1) & 1), 150);
        }
      }
      act = __shfl_sync(0xffffffffu, act, 0);
      if (act == 1) { l1sum(ts); ++ts; idle = 0; }
      else if (act == 2) { drain(td); ++td; idle = 0; }
      else if (++idle > kFSpinLimit) f_timeout(a, 0x330, (uint32_t)td, (uint32_t)ts);
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
      const uint32_t tag4 = seq & 15u;
      // published weights: [ring][job 32][q 64][4 rows] fp32 whose low 4 mantissa bits carry the tile's sequence number:
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
        u_tmem_st32(tlane + kFColP + (uint32_t)b * 64, pk);
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
        u_tmem_st32(tlane + kFColP + (uint32_t)b * 64, ph);
        u_tmem_st32(tlane + kFColP + (uint32_t)b * 64 + 32, pl);
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
        u_tmem_ld32_nowait(tlane + kFColAcc + c * 32, r);
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
    // ---- jobs (tile t, query group g, rows [4j, 4j+4)) dealt round-robin over all CTAs.  Thread = (query row,
    //      row pair): lane l of warp w reads line 32 w + l of the job's 2 KiB block of every cluster (512 contiguous
    //      bytes per warp instruction) and sums the clusters in order.
    {
      const int q64 = ow * 16 + (lane >> 1);
      const int lh = lane & 1;
      const int njobs = ntiles * C::kJobs;
#pragma unroll 1
      for (int J = cta; J < njobs; J += nctas) {
        const int t = J / C::kJobs, jj = J - t * C::kJobs;
        const int j = jj & (kFJobsPerGroup - 1), g = jj / kFJobsPerGroup;
        const int q = g * 64 + q64;
        const uint32_t tag = epoch0 + (uint32_t)t;
        const int ring = (int)(tag % kFRing);
        const int row = t * kFR + j * kFJobRows + lh * 2;
        // ||n||^2 of the two rows: issued before the wait, the line is cold
        const float sq0 = row < a.N ? __ldg(a.sqnorm + row) : 0.f;
        const float sq1 = row + 1 < a.N ? __ldg(a.sqnorm + row + 1) : 0.f;
        const size_t cl_stride = (size_t)32 * 2048;
        const uint8_t* src0 = a.ar.part_ll + (size_t)ring * kFMaxClusters * cl_stride + (size_t)jj * 2048 + (size_t)tid * 16;
        // A job usually arrives long before its tile: ONE lane per warp probes one line, with back-off, until the tile
        // shows up (every owner warp of every CTA spinning on full-width loads saturates the L1s and L2)
        {
          uint32_t spin = 0;
          int seen = 0;
          do {
            if (lane == 0) {
              const uint4 l0 = u_ll_load(src0);
              seen = (l0.y == tag && l0.w == tag) ? 1 : 0;
              if (!seen) __nanosleep(20);
            }
            seen = __shfl_sync(0xffffffffu, seen, 0);
            if (!seen && ++spin > kFSpinLimit) f_timeout(a, 0x508, (uint32_t)t, (uint32_t)jj);
          } while (!seen);
          if (ow == 0 && lane == 0) f_trace(a, t, 9);
        }
        float dv0 = 0.f, dv1 = 0.f;
        for (int k0 = 0; k0 < nclusters; k0 += 16) {
          uint4 ln[16];
          uint32_t spin = 0;
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int k = min(k0 + u, nclusters - 1);
              ln[u] = u_ll_load(src0 + (size_t)k * cl_stride);
              ok = ok && ln[u].y == tag && ln[u].w == tag;
            }
            if (!ok && ++spin > kFSpinLimit) f_timeout(a, 0x510, (uint32_t)t, (uint32_t)(jj * 256 + k0));
          } while (!ok);
#pragma unroll
          for (int u = 0; u < 16; ++u)
            if (k0 + u < nclusters) { dv0 += __uint_as_float(ln[u].x); dv1 += __uint_as_float(ln[u].z); }
        }
        if (ow == 0 && lane == 0) f_trace(a, t, 10);
        const float xs = sm.xsq[q];
        float k0v = 0.f, k1v = 0.f;
        if (q < a.Q) {
          if (row < a.N) k0v = expf(-dist_from_dot(xs, sq0, dv0, a.alpha, a.power) * a.inv2s2);
          if (row + 1 < a.N) k1v = expf(-dist_from_dot(xs, sq1, dv1, a.alpha, a.power) * a.inv2s2);
        }
        // the weights carry the low 4 bits of the sequence number in their low mantissa bits (2^-19 relative): no fence,
        // no flag -- each 32-bit word is valid on its own
        {
          const uint32_t b0 = (__float_as_uint(k0v) & ~15u) | (tag & 15u), b1 = (__float_as_uint(k1v) & ~15u) | (tag & 15u);
          asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};"
                       ::"l"(a.ar.gw + (((size_t)ring * 32 + jj) * 64 + q64) * 4 + lh * 2), "r"(b0), "r"(b1) : "memory");
        }
        if (a.k_out && q < a.Q) {
          if (row < a.N) a.k_out[(int64_t)q * a.N + row] = k0v;
          if (row + 1 < a.N) a.k_out[(int64_t)q * a.N + row + 1] = k1v;
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
    const uint32_t one[2] = {1, 1};
    SDN_CUDA_OK(cudaMemcpy(h.ar.epoch, one, sizeof(one), cudaMemcpyHostToDevice));   // tags start at 1: zeroed lines never match
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
    a.D = D; a.inv2s2 = inv2s2; a.alpha = alpha; a.power = power;
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
