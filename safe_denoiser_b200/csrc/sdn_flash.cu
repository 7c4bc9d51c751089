// One-pass "flash-repellency" for batched calls: the bank is read from HBM exactly once.
//
// Replaces repellency_methods_fast.py:249-250 (cdist + the [Q,N,D+1] broadcast) for 8 < Q <= 64 query rows per pass.
// The two-phase tcgen05 path (sdn_umma.cu) streams the bank twice from HBM because the weights of a bank row need its
// dot product over ALL of D before the row can be accumulated, and its two kernels are a whole bank apart.  Here both
// contractions run in ONE persistent kernel, a few tiles apart, so that the second read of a tile hits the L2:
//
//   grid     D / 128 CTAs, one per SM (co-resident: they exchange data while running).  Four consecutive CTAs form a
//            group that covers a 512-wide "super-slice" of D; the query rows of a pass are stacked as 64 hi parts +
//            64 lo parts = the 128 lanes of tensor memory.
//   phase A  (distances) CTA r of a group takes the tiles t = r (mod 4) over the WHOLE super-slice, in four chunks of
//            128 d (one 32 KiB stage each: hi and lo planes of [64 rows][128 d], TMA, L2 evict_last):
//            S[q][i] = sum_{d in super-slice} X[q][d] bank[i][d], tcgen05.mma with A = X[:, super-slice] (bf16 hi/lo,
//            resident in TENSOR MEMORY for the whole kernel), B = the chunk (K-major).  The group's 4:1 reduction over
//            d thus happens inside the tensor core.  (Built and measured before: a DSMEM exchange between the four
//            CTAs, 1.9 us per tile in mbarrier round trips; and 16 rows x 512 d per CTA and tile, 3.2 us per tile
//            because a tcgen05.mma with N = 16 costs as much as one with N = 64 -- the A operand bounds it.)
//            The accumulator is drained every two chunks (32 MMAs): the tensor core adds with truncation.
//   exchange S[q][i] = sum over the D/512 groups.  The group partials go through L2 as self-validating 16-byte lines
//            {v, tag, v, tag} (no fence, no flag) to work units of 4 rows x 32 queries dealt round-robin over ALL
//            4 x #CTAs warps of the level-2 role: a unit sums the groups in order, computes k = exp(-dist / 2 sigma^2)
//            and publishes the weights as fp32 words whose low 4 mantissa bits carry the tile's round number (again no
//            fence, no flag: every word validates itself).
//   phase B  (accumulate) CTA j owns the d-slice [128 j, 128 j + 128): num[q][slice] += sum_i k[q][i] bank[i][slice],
//            tcgen05.mma with A = the weights (bf16 hi/lo written to TMEM by tcgen05.st), B = the tile [64 rows][128 d]
//            viewed MN-major, loaded AGAIN by TMA (evict_first) when the tile's weights show up: a hit in the 126 MB L2
//            because phase A never runs more than `window` tiles ahead.  The accumulator stays in tensor memory for
//            the whole kernel; the epilogue applies the correction of conditioning() when fused.
//
// Every sum has a fixed order (group index, row index): results are bit-reproducible run to run.
//
// Warp roles (16 warps): 0 TMA producer phase A | 1 MMA issuer (event loop over "tile ta loaded" / "weights and tile
// tb ready") | 4-7 query prologue, then phase-A drain: TMEM -> registers -> LL lines | 8-11 weights: global -> bf16
// hi/lo -> TMEM (warp 8 also issues the second TMA read of the tile), z_q, and the final epilogue | 12-15 level-2 work
// units: ||x||^2, sum over groups, exp, publish.
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
constexpr int kFDS = 128;               // d per CTA in phase B
constexpr int kFGS = 4;                 // CTAs per group
constexpr int kFSuper = kFDS * kFGS;    // d per group in phase A (512)
constexpr int kFQ = 64;                 // query rows per pass
constexpr int kFJobRows = 4;            // rows of a tile per level-2 job
constexpr int kFJobs = kFR / kFJobRows; // 16 jobs per tile, two work units (halves of the query rows) each
constexpr int kFRing = 32;              // slots of the global exchange rings: phase A runs at most kFWindow < kFRing tiles ahead of phase B
constexpr int kFWindow = 12;            // tiles between the two reads of the bank: 12 x 4 MiB stay in the 126 MB L2 (and < the 16 tiles between two units of a level-2 worker)
constexpr int kFThreads = 512;
constexpr int kFMaxCtas = 128;
constexpr int kFMaxGroups = kFMaxCtas / kFGS;
constexpr uint32_t kFStageBytes = 32768;   // hi + lo planes of [64 rows][128 d] bf16 (phase A: one chunk of the super-slice; phase B: the CTA's slice)
constexpr int kFSA = 3;                 // stages of phase A (HBM stream)
constexpr int kFSB = 3;                 // stages of phase B (second read: L2)
constexpr uint32_t kFTmemCols = 512;
constexpr uint32_t kFColX = 0;          // X operand [128 stacked rows][512 d] bf16: 256 columns
constexpr uint32_t kFColS = 256;        // S accumulator [128][64 bank rows]
constexpr uint32_t kFColP = 320;        // weight operands: 2 x 32 columns
constexpr uint32_t kFColAcc = 384;      // num accumulator [128][128 d]
constexpr uint32_t kFSpinLimit = 1u << 24;
constexpr int kFPitch = 132;            // floats per staged row (thread = row accesses are conflict-free per quarter warp)
constexpr uint32_t kFSmemBytes = (kFSA + kFSB) * kFStageBytes + 1024 /*align*/ + 1536 /*barriers, xsq*/;

// Arena: library-owned per-device synchronisation memory of the one-pass kernel (zeroed once; tags grow
// monotonically over launches, so stale lines never match).
struct FlashArena {
  uint32_t* epoch;      // [0] sequence number of the next launch's tile 0, [1] launch counter
  uint8_t* xsq_ll;      // [kFMaxCtas][128] 16-byte lines: ||x_q||^2 partial of every CTA
  uint8_t* part_ll;     // [ring][group 32][job 16][q 64][2 lines]   level-2 fan-in
  float* gw;            // [ring][job 16][q 64][4 rows]  published weights
  uint32_t* diag;       // host-mapped: written before a timeout trap
  unsigned long long* trace;   // SDN_FLASH_TRACE=1: [cta][tile < 64][16 events] globaltimer ns (null otherwise)
};

struct FlashArgs {
  const float* xq;      // [Q][D] query the distances are taken on
  const float* sqnorm;  // [N]
  int Q, N, ntiles, ngroups, window;
  unsigned mma_sleep, poll_sleep;     // back-off of the polling loops (ns)
  unsigned dbg;                        // SDN_FLASH_DBG experiments: 1 = one phase-A MMA per chunk, 2 = one phase-B MMA per tile (wrong results)
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
  uint8_t* stages;      // [kFSA phase-A stages | kFSB phase-B stages] x 32 KiB
  uint64_t* afull; uint64_t* aempty;      // [kFSA]
  uint64_t* bfull; uint64_t* bempty;      // [kFSB]
  uint64_t* sfull; uint64_t* sempty;      // [1] S accumulator (one drain per two chunks)
  uint64_t* pfull; uint64_t* pempty;      // [2] weight operands
  uint64_t* xfull; uint64_t* accfull;     // [1]
  uint64_t* xload;                        // [4 warps][2] query rows of one prologue warp staged in shared memory
  uint32_t* tmem_base;
  float* xsq;           // [64]
  float* xsq_half;      // [128]  (||x||^2 halves at start, z_q at the end)
};

__device__ __forceinline__ FSmem f_carve(unsigned char* raw) {
  FSmem s;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023;
  s.stages = reinterpret_cast<uint8_t*>(a);
  uint64_t* b = reinterpret_cast<uint64_t*>(s.stages + (size_t)(kFSA + kFSB) * kFStageBytes);
  s.afull = b; b += 4;
  s.aempty = b; b += 4;
  s.bfull = b; b += 4;
  s.bempty = b; b += 4;
  s.sfull = b; b += 4;
  s.sempty = b; b += 4;
  s.pfull = b; b += 2;
  s.pempty = b; b += 2;
  s.xfull = b; b += 1;
  s.accfull = b; b += 1;
  s.xload = b; b += 8;
  s.tmem_base = reinterpret_cast<uint32_t*>(b); b += 1;
  s.xsq = reinterpret_cast<float*>(b);          // 39 x 8 = 312 bytes of barriers so far
  s.xsq_half = s.xsq + 128;
  return s;
}

__global__ void __launch_bounds__(kFThreads, 1)
k_flash(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
        const __grid_constant__ FlashArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const FSmem sm = f_carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int grp = cta / kFGS, rk = cta % kFGS;
  const int d0 = cta * kFDS;            // phase-B slice
  const int ds0 = grp * kFSuper;        // phase-A super-slice
  const int ntiles = a.ntiles;
  const int ngroups = a.ngroups;
  // query staging (two buffers of [64 rows][132 floats]) lives in the phase-B stages, which are idle until the first
  // weights arrive; the epilogue staging in the phase-A stages, which are idle by then
  float* const xstage = reinterpret_cast<float*>(sm.stages + (size_t)kFSA * kFStageBytes);
  float* const estage = reinterpret_cast<float*>(sm.stages);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFSA; ++s) { u_mbar_init(&sm.afull[s], 1); u_mbar_init(&sm.aempty[s], 1); }
    for (int s = 0; s < kFSB; ++s) { u_mbar_init(&sm.bfull[s], 1); u_mbar_init(&sm.bempty[s], 1); }
    u_mbar_init(sm.sfull, 1); u_mbar_init(sm.sempty, 4);
    for (int b = 0; b < 2; ++b) { u_mbar_init(&sm.pfull[b], 4); u_mbar_init(&sm.pempty[b], 1); }
    u_mbar_init(sm.xfull, 4); u_mbar_init(sm.accfull, 1);
    for (int w = 0; w < 8; ++w) u_mbar_init(&sm.xload[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    u_prefetch_map(&tm_hi); u_prefetch_map(&tm_lo);
    if (cta == 0 && a.epi.zero_mean && a.epi.mean_out) { *a.epi.mean_out = 0.f; __threadfence(); }
  }
  if (warp == 1) u_tmem_alloc(sm.tmem_base, kFTmemCols);
  // tag bases of this launch: read by everyone before anything is exchanged (CTA 0 advances them at the very end)
  const uint32_t epoch0 = *reinterpret_cast<volatile uint32_t*>(a.ar.epoch);        // sequence number of tile 0
  const uint32_t launch0 = *reinterpret_cast<volatile uint32_t*>(a.ar.epoch + 1);    // tag of the ||x||^2 exchange
  u_fence_before();
  __syncthreads();
  u_fence_after();
  const uint32_t tmem = *sm.tmem_base;

  uint64_t pol_keep, pol_drop;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
  // a stage: hi and lo planes of [64 rows][128 d], two boxes of [64 rows][64 d] per plane (8 KiB each)
  auto load_tile = [&](uint8_t* st, uint64_t* bar, int t, int dd, uint64_t policy) {
    u_mbar_expect_tx(bar, kFStageBytes);
    const int r0 = t * kFR;
    f_tma_2d_hint(st, &tm_hi, dd, r0, bar, policy);
    f_tma_2d_hint(st + 8192, &tm_hi, dd + 64, r0, bar, policy);
    f_tma_2d_hint(st + 16384, &tm_lo, dd, r0, bar, policy);
    f_tma_2d_hint(st + 24576, &tm_lo, dd + 64, r0, bar, policy);
  };
  // phase A unit ua = (own tile i, chunk c): tile t = rk + 4 i, d in [ds0 + 128 c, +128)
  const int nown = (ntiles - rk + kFGS - 1) / kFGS;
  const int nua = nown * kFGS;
  auto load_a = [&](int ua) {
    const int t = rk + kFGS * (ua >> 2), c = ua & 3;
    if (c == 0) f_trace(a, t, 0);
    load_tile(sm.stages + (size_t)(ua % kFSA) * kFStageBytes, &sm.afull[ua % kFSA], t, ds0 + c * kFDS, pol_keep);
  };
  auto load_b = [&](int t) {
    load_tile(sm.stages + (size_t)(kFSA + t % kFSB) * kFStageBytes, &sm.bfull[t % kFSB], t, d0, pol_drop);
  };

  if (warp == 0) {
    // ============================================================ TMA producer, phase A (HBM stream)
    if (lane == 0) {
      for (int ua = 0; ua < nua; ++ua) {
        if (ua >= kFSA) f_wait(a, &sm.aempty[ua % kFSA], (uint32_t)(((ua / kFSA) + 1) & 1), 0x100, ua);
        load_a(ua);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idA = u_idesc(128, kFR, 0, 0);      // S[128][64 rows]   = X (TMEM) x chunk (K-major)
      constexpr uint32_t idB = u_idesc(128, kFDS, 0, 1);     // num[128][128 d] += P (TMEM) x tile (MN-major)
      f_wait(a, sm.xfull, 0, 0x200, 0);
      u_fence_after();
      int ua = 0, tb = 0;                 // next phase-A unit (own tile, chunk) / next phase-B tile
      uint32_t idle = 0;
      while (tb < ntiles) {
        bool did = false;
        if (u_mbar_test(&sm.pfull[tb & 1], (uint32_t)((tb >> 1) & 1)) &&
            u_mbar_test(&sm.bfull[tb % kFSB], (uint32_t)((tb / kFSB) & 1))) {
          // ---- phase B of tile tb: 4 K steps of 16 bank rows
          u_fence_after();
          f_trace(a, tb, 8);
          const uint32_t base = u_smem(sm.stages + (size_t)(kFSA + tb % kFSB) * kFStageBytes);
          const uint32_t acc = tmem + kFColAcc;
          const uint32_t pb = tmem + kFColP + (uint32_t)(tb & 1) * 32;
#pragma unroll
          for (int kk = 0; kk < kFR / 16; ++kk) {
            if ((a.dbg & 2) && kk > 0) break;
            const uint64_t bh = u_desc(base + kk * 2048, 8192, 1024);
            const uint64_t bl = u_desc(base + 16384 + kk * 2048, 8192, 1024);
            u_mma_ts(acc, pb + kk * 8, bh, idB, (tb > 0 || kk > 0) ? 1u : 0u);
            u_mma_ts(acc, pb + kk * 8, bl, idB, 1u);
          }
          u_commit(&sm.bempty[tb % kFSB]);
          u_commit(&sm.pempty[tb & 1]);
          if (tb == ntiles - 1) u_commit(sm.accfull);
          ++tb;
          did = true;
        }
        // phase A stays within `window` tiles of phase B: the exchange rings (kFRing slots) are reused safely and the
        // tiles waiting for their second read fit in the L2.  The S accumulator is drained every two chunks (half h).
        if (ua < nua) {
          const int ta = rk + kFGS * (ua >> 2), c = ua & 3;
          const int sh = ua >> 1;           // drains so far = halves started
          if (ta - tb < a.window && u_mbar_test(&sm.afull[ua % kFSA], (uint32_t)((ua / kFSA) & 1)) &&
              ((c & 1) || sh == 0 || u_mbar_test(sm.sempty, (uint32_t)((sh + 1) & 1)))) {
            u_fence_after();
            if (c == 0) f_trace(a, ta, 1);
            const uint32_t base = u_smem(sm.stages + (size_t)(ua % kFSA) * kFStageBytes);
            const uint32_t acc = tmem + kFColS;
#pragma unroll
            for (int kk = 0; kk < kFDS / 16; ++kk) {
              if ((a.dbg & 1) && kk > 0) break;
              const uint32_t off = (uint32_t)(kk >> 2) * 8192 + (uint32_t)(kk & 3) * 32;
              const uint64_t bh = u_desc(base + off, 16, 1024);
              const uint64_t bl = u_desc(base + 16384 + off, 16, 1024);
              const uint32_t xa = tmem + kFColX + (uint32_t)(c * 8 + kk) * 8;
              u_mma_ts(acc, xa, bh, idA, ((c & 1) || kk > 0) ? 1u : 0u);
              u_mma_ts(acc, xa, bl, idA, 1u);
            }
            if (c & 1) u_commit(sm.sfull);
            u_commit(&sm.aempty[ua % kFSA]);
            ++ua;
            did = true;
          }
        }
        if (did) idle = 0;
        else {
          if (a.mma_sleep) __nanosleep(a.mma_sleep);
          if (++idle > kFSpinLimit) f_timeout(a, 0x210, (uint32_t)ua, (uint32_t)tb);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================================================ query prologue, then phase-A drain
    const int lq = warp & 3;
    const int part = lane >> 4;                         // lanes 0-15 hold hi parts, 16-31 lo parts of the same q
    const int q = lq * 16 + (lane & 15);
    const uint32_t tlane = tmem + ((uint32_t)(lq * 32) << 16);
    // ---- X[:, super-slice] -> bf16 hi / lo -> tensor memory, in four chunks of 128 d staged through shared memory
    //      (whole 512-byte rows by bulk copies, one per lane, then thread = row: a warp-wide global load of 32
    //      different rows costs 32 L1 wavefronts per instruction); ||x_q||^2 partial of this CTA's own slice
    {
      const int row0 = lq * 16;
      const int nvalid = max(0, min(16, a.Q - row0));
      const bool valid = q < a.Q;
      auto fetch = [&](int c) {
        uint64_t* bar = &sm.xload[lq * 2 + (c & 1)];
        if (lane == 0) u_mbar_expect_tx(bar, (uint32_t)nvalid * (kFDS * 4));
        __syncwarp();
        if (lane < nvalid) {
          const int r = row0 + lane;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(u_smem(xstage + ((size_t)(c & 1) * kFQ + r) * kFPitch)),
                         "l"(a.xq + (int64_t)r * a.D + ds0 + c * kFDS), "r"(kFDS * 4), "r"(u_smem(bar)) : "memory");
        }
      };
      fetch(0);
      float ss = 0.f;
#pragma unroll 1
      for (int c = 0; c < kFGS; ++c) {
        if (c + 1 < kFGS) fetch(c + 1);                 // the other buffer: its previous chunk was consumed by this warp
        f_wait(a, &sm.xload[lq * 2 + (c & 1)], (uint32_t)((c >> 1) & 1), 0x340, c);
        const float* xr = xstage + ((size_t)(c & 1) * kFQ + q) * kFPitch;
#pragma unroll 1
        for (int c4 = 0; c4 < kFDS / 32; ++c4) {           // 32 d = 16 columns per step
          float f[32];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 t4 = valid ? *reinterpret_cast<const float4*>(xr + c4 * 32 + v * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            f[v * 4 + 0] = t4.x; f[v * 4 + 1] = t4.y; f[v * 4 + 2] = t4.z; f[v * 4 + 3] = t4.w;
          }
          uint32_t sel[16];
          float s2 = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float h0 = f_bf16_hi(f[2 * j]), h1 = f_bf16_hi(f[2 * j + 1]);
            sel[j] = part ? f_pack_bf16(f[2 * j] - h0, f[2 * j + 1] - h1) : f_pack_bf16(h0, h1);
            s2 = fmaf(f[2 * j], f[2 * j], s2);
            s2 = fmaf(f[2 * j + 1], f[2 * j + 1], s2);
          }
          if (c == rk) ss += s2;
          u_tmem_st16(tlane + kFColX + c * 64 + c4 * 16, sel);
        }
        __syncwarp();                                   // every lane has read buffer c & 1 before it is refilled
      }
      u_tmem_st_wait();
      if (part == 0)
        u_ll_store(a.ar.xsq_ll + ((size_t)cta * 128 + q) * 16, ss, 0.f, launch0);
      u_fence_before();
      __syncwarp();
      if (lane == 0) u_mbar_arrive(sm.xfull);
    }
    // ---- drain: S [128 stacked rows][64 bank rows] of an own tile, twice (after chunks 1 and 3) -> registers ->
    //      LL lines for the level-2 work units.  Layout per (ring slot, group): [job 16][q 64][2 lines of 2 rows].
#pragma unroll 1
    for (int i = 0; i < nown; ++i) {
      const int t = rk + kFGS * i;
      float acc32[32];
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int sh = i * 2 + h;
        f_wait(a, sm.sfull, (uint32_t)(sh & 1), 0x310, t);
        u_fence_after();
        if (warp == 4 && lane == 0 && h == 1) f_trace(a, t, 2);
        uint32_t r0[32], r1[32];
        u_tmem_ld32_nowait(tlane + kFColS, r0);
        u_tmem_ld32_nowait(tlane + kFColS + 32, r1);
        u_tmem_ld_wait();
        u_fence_before();
        __syncwarp();
        if (lane == 0) u_mbar_arrive(sm.sempty);
        // hi-part row + lo-part row of the same query; lanes l / l+16 keep bank rows [0, 32) / [32, 64) of the tile
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v0 = __uint_as_float(r0[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r0[j]), 16);
          const float v1 = __uint_as_float(r1[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r1[j]), 16);
          const float v = part ? v1 : v0;
          acc32[j] = h ? acc32[j] + v : v;
        }
      }
      const uint32_t tag = epoch0 + (uint32_t)t;            // the tile's sequence number (gap-free over launches)
      uint8_t* dst = a.ar.part_ll + ((size_t)(tag % kFRing) * kFMaxGroups + grp) * (size_t)(kFJobs * 2048) +
                     (size_t)(part * 8) * 2048 + (size_t)q * 32;
#pragma unroll
      for (int sub = 0; sub < 8; ++sub) {
        uint8_t* o = dst + (size_t)sub * 2048;
        u_ll_store(o, acc32[4 * sub], acc32[4 * sub + 1], tag);
        u_ll_store(o + 16, acc32[4 * sub + 2], acc32[4 * sub + 3], tag);
      }
      if (warp == 4 && lane == 0) f_trace(a, t, 5);
    }
  } else if (warp >= 8 && warp < 12) {
    // ============================================================ weights -> tensor memory, z, final epilogue
    const int lq = warp & 3;
    const int part = lane >> 4;
    const int q = lq * 16 + (lane & 15);
    const uint32_t tlane = tmem + ((uint32_t)(lq * 32) << 16);
    float z = 0.f;
    int nb = 0;                           // next tile whose second read warp 8 has to issue
    // the phase-B stages double as the staging area of the query prologue: no TMA into them before it is done
    if (warp == 8 && lane == 0) f_wait(a, sm.xfull, 0, 0x122, 0);
#pragma unroll 1
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t seq = epoch0 + (uint32_t)t;
      const int ring = (int)(seq % kFRing);
      const uint32_t tag4 = (seq / kFRing) & 15u;    // the slot's previous occupant carries tag4 - 1
      // published weights: [ring][job 16][q 64][4 rows] fp32 whose low 4 mantissa bits carry the tile's round (seq / ring):
      // every word validates itself, so the units need no fence and no flag and the consumers no acquire -- the slot's
      // previous occupant is exactly kFRing tiles older (ring slot = seq mod kFRing, gap-free over launches).
      const uint32_t* wq = reinterpret_cast<const uint32_t*>(a.ar.gw) + (size_t)ring * (kFJobs * 64 * 4) + ((size_t)(part * 8) * 64 + q) * 4;
      const int b = t & 1;
      if (warp == 8 && lane == 0) f_trace(a, t, 12);
      // warp 8 keeps the second reads of the next tiles in flight (L2 -> phase-B stages)
      if (warp == 8 && lane == 0) {
        while (nb < ntiles && nb < t + kFSB &&
               (nb < kFSB || u_mbar_test(&sm.bempty[nb % kFSB], (uint32_t)(((nb / kFSB) + 1) & 1)))) load_b(nb++);
      }
      // try the full-width loads first (in steady state the weights are already there: one round trip, not two);
      // while the tile is missing, ONE lane probes one word -- every warp spinning on 4 KiB loads would flood the L2
      uint4 wv[8];
      {
        uint32_t spin = 0;
        bool ok;
        for (;;) {
          ok = true;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            wv[v] = u_ll_load(wq + (size_t)v * 64 * 4);
            ok = ok && ((wv[v].x & 15u) == tag4) && ((wv[v].y & 15u) == tag4) && ((wv[v].z & 15u) == tag4) && ((wv[v].w & 15u) == tag4);
          }
          ok = __all_sync(0xffffffffu, ok);
          if (ok) break;
          int seen = 0;
          do {
            if (lane == 0) {
              uint32_t w0;
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w0) : "l"(wq + (size_t)7 * 64 * 4) : "memory");
              seen = (w0 & 15u) == tag4;
            }
            seen = __shfl_sync(0xffffffffu, seen, 0);
            if (++spin > kFSpinLimit) f_timeout(a, 0x400, (uint32_t)t, tag4);
          } while (!seen);
        }
      }
      if (warp == 8 && lane == 0) {     // the tile's own second read must be on its way by now
        while (nb <= t) {
          if (nb >= kFSB) f_wait(a, &sm.bempty[nb % kFSB], (uint32_t)(((nb / kFSB) + 1) & 1), 0x121, nb);
          load_b(nb++);
        }
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
      if (warp == 8 && lane == 0) f_trace(a, t, 13);
      if (t >= 2) { f_wait(a, &sm.pempty[b], (uint32_t)(((t >> 1) + 1) & 1), 0x410, t); u_fence_after(); }
      if (warp == 8 && lane == 0) f_trace(a, t, 14);
      u_tmem_st32(tlane + kFColP + (uint32_t)b * 32, pk);
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
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float lo16 = __uint_as_float(r[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r[j]), 16);
          const float hi16 = __uint_as_float(r[16 + j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r[16 + j]), 16);
          v[j] = part ? hi16 : lo16;
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(er + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      if (part == 0) sm.xsq_half[q] = z;
    }
    u_fence_before();
    asm volatile("bar.sync 2, 128;" ::: "memory");
    const int nrows = min(a.Q, kFQ);
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
    // ============================================================ level-2 work units: ||x||^2, sum over groups, exp, publish
    const int ow = warp - 12;
    const int tid = ow * 32 + lane;                    // 0..127
    const int nctas = gridDim.x;
    // ---- ||x_q||^2 = sum over the CTAs' slices, fixed order
    {
      const int qx = tid & 63, half = tid >> 6;
      const int per = nctas / 2;
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
      sm.xsq_half[tid] = acc;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid < 64) sm.xsq[tid] = sm.xsq_half[tid] + sm.xsq_half[64 + tid];
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    // ---- work units (tile t, rows [4j, 4j+4), half h of the 64 query rows) dealt round-robin over all 4 x #CTAs
    //      warps of this role: every warp is an independent worker (a worker blocks on its unit until the slowest
    //      group has delivered, so the tiles between two units of one worker bound the rate: 16 here).  Lane l reads
    //      lines l and 32 + l of the unit's 1 KiB block of every group (512 contiguous bytes per warp instruction)
    //      = (query row, row pair) twice, and sums the groups in order.
    {
      const int lh = lane & 1;
      const int nunits = ntiles * kFJobs * 2;
      const int nworkers = nctas * 4;
#pragma unroll 1
      for (int U = cta * 4 + ow; U < nunits; U += nworkers) {
        const int t = U / (kFJobs * 2), rem = U - t * (kFJobs * 2);
        const int j = rem >> 1, qh = rem & 1;
        const uint32_t tag = epoch0 + (uint32_t)t;
        const int ring = (int)(tag % kFRing);
        const int row = t * kFR + j * kFJobRows + lh * 2;
        // ||n||^2 of the two rows: issued before the wait, the line is cold
        const float sq0 = row < a.N ? __ldg(a.sqnorm + row) : 0.f;
        const float sq1 = row + 1 < a.N ? __ldg(a.sqnorm + row + 1) : 0.f;
        const size_t g_stride = (size_t)kFJobs * 2048;
        const uint8_t* src0 = a.ar.part_ll + (size_t)ring * kFMaxGroups * g_stride + (size_t)j * 2048 +
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
        for (int k0 = 0; k0 < ngroups; k0 += 8) {
          uint4 la[8], lb[8];
          uint32_t spin = 0;
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int k = min(k0 + u, ngroups - 1);
              la[u] = u_ll_load(src0 + (size_t)k * g_stride);
              lb[u] = u_ll_load(src0 + (size_t)k * g_stride + 512);
              ok = ok && la[u].y == tag && la[u].w == tag && lb[u].y == tag && lb[u].w == tag;
            }
            if (!ok && ++spin > kFSpinLimit) f_timeout(a, 0x510, (uint32_t)t, (uint32_t)(rem * 256 + k0));
          } while (!ok);
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (k0 + u < ngroups) {
              da0 += __uint_as_float(la[u].x); da1 += __uint_as_float(la[u].z);
              db0 += __uint_as_float(lb[u].x); db1 += __uint_as_float(lb[u].z);
            }
        }
        if (ow == 0 && lane == 0) f_trace(a, t, 10);
        const uint32_t tagbits = (tag / kFRing) & 15u;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int q = qh * 32 + i * 16 + (lane >> 1);
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
                       ::"l"(a.ar.gw + (((size_t)ring * kFJobs + j) * 64 + q) * 4 + lh * 2), "r"(b0), "r"(b1) : "memory");
          if (a.k_out && q < a.Q) {
            if (row < a.N) a.k_out[(int64_t)q * a.N + row] = k0v;
            if (row + 1 < a.N) a.k_out[(int64_t)q * a.N + row + 1] = k1v;
          }
        }
        if (ow == 0 && lane == 0) f_trace(a, t, 11);
      }
    }
  }

  u_fence_before();
  __syncthreads();
  if (warp == 1) {
    u_fence_after();
    u_tmem_dealloc(tmem, kFTmemCols);
  }
  // every CTA has read the tag bases long before any CTA can get here (each contributed to the last tile)
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
  int max_ctas = 0;       // co-resident CTAs of k_flash on this device
  bool coop_ok = true;
};
ArenaHost g_arena[kMaxDevices];

constexpr size_t kXsqBytes = (size_t)kFMaxCtas * 128 * 16;
constexpr size_t kPartBytes = (size_t)kFRing * kFMaxGroups * kFJobs * 2048;
constexpr size_t kGwBytes = (size_t)kFRing * kFJobs * 64 * 4 * 4;

int arena_get(int dev, ArenaHost** out) {
  ArenaHost& h = g_arena[dev];
  std::lock_guard<std::mutex> lk(h.mu);
  if (!h.ready) {
    const size_t total = 256 + kXsqBytes + kPartBytes + kGwBytes;
    SDN_CUDA_OK(cudaMalloc(&h.dev, total));
    SDN_CUDA_OK(cudaMemset(h.dev, 0, total));
    uint8_t* p = static_cast<uint8_t*>(h.dev);
    h.ar.epoch = reinterpret_cast<uint32_t*>(p); p += 256;
    h.ar.xsq_ll = p; p += kXsqBytes;
    h.ar.part_ll = p; p += kPartBytes;
    h.ar.gw = reinterpret_cast<float*>(p); p += kGwBytes;
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
    SDN_CUDA_OK(cudaFuncSetAttribute(k_flash, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFSmemBytes));
    int per_sm = 0, sms = 0;
    SDN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_flash, kFThreads, kFSmemBytes));
    SDN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    h.max_ctas = per_sm * sms;
    if (getenv("SDN_FLASH_DEBUG")) fprintf(stderr, "[sdn_flash] co-resident CTAs: %d x %d SMs\n", per_sm, sms);
    if (getenv("SDN_FLASH_NOCOOP")) h.coop_ok = false;     // profilers may refuse cooperative launches
    h.ready = true;
  }
  *out = &h;
  return SDN_OK;
}

// tensor maps of the bank planes ([N][D] bf16), boxes [64 rows][64 d];
// a few entries per process, keyed on the bank
struct FlashMaps { int dev; const void* planes; int64_t N, D; CUtensorMap m[2]; uint64_t stamp; };
std::mutex g_maps_mu;
FlashMaps g_maps[8];
int g_maps_n = 0;
uint64_t g_maps_clock = 0;

int maps_get(int dev, const void* planes, int64_t N, int64_t D, CUtensorMap* out2) {
  std::lock_guard<std::mutex> lk(g_maps_mu);
  for (int i = 0; i < g_maps_n; ++i) {
    FlashMaps& m = g_maps[i];
    if (m.dev == dev && m.planes == planes && m.N == N && m.D == D) {
      m.stamp = ++g_maps_clock;
      memcpy(out2, m.m, sizeof(m.m));
      return SDN_OK;
    }
  }
  int slot = g_maps_n < 8 ? g_maps_n : 0;
  if (g_maps_n == 8)
    for (int i = 1; i < 8; ++i) if (g_maps[i].stamp < g_maps[slot].stamp) slot = i;
  FlashMaps m{};
  const __nv_bfloat16* h = static_cast<const __nv_bfloat16*>(planes);
  int rc;
  if ((rc = tmap_bf16_2d(&m.m[0], h, (uint64_t)N, (uint64_t)D, kFR, 64))) return rc;
  if ((rc = tmap_bf16_2d(&m.m[1], h + N * D, (uint64_t)N, (uint64_t)D, kFR, 64))) return rc;
  m.dev = dev; m.planes = planes; m.N = N; m.D = D; m.stamp = ++g_maps_clock;
  g_maps[slot] = m;
  if (slot == g_maps_n) ++g_maps_n;
  memcpy(out2, m.m, sizeof(m.m));
  return SDN_OK;
}

int launch_flash(const CUtensorMap* m, const FlashArgs& a, int nctas, bool coop, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nctas);
  cfg.blockDim = dim3(kFThreads);
  cfg.dynamicSmemBytes = kFSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;     // every CTA must be resident: they wait on one another
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = coop ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, k_flash, m[0], m[1], a);
}

}  // namespace

bool flash_shape_ok(int64_t Q, int64_t N, int64_t D) {
  if (Q < 1 || N < 1 || N >= (1ll << 30) || D % 1024) return false;   // whole groups of 4 CTAs, an even number of them
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

// One pass over the bank per <= 64 query rows.  epi == nullptr: partial sums (num_out, z_out).
int flash_run(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq, int64_t Q,
              float inv2s2, int power, float alpha, float* num_out, float* z_out, float* k_out,
              const FlashEpi* epi, cudaStream_t st) {
  if (!flash_supported(Q, N, D, planes)) return SDN_E_UNSUPPORTED;
  const int dev = device_slot();
  ArenaHost* h = nullptr;
  int rc = arena_get(dev, &h);
  if (rc) return rc;
  const int nctas = (int)(D / kFDS);
  if (h->max_ctas < nctas) return SDN_E_UNSUPPORTED;    // the grid must be co-resident
  CUtensorMap maps[2];
  if ((rc = maps_get(dev, planes, N, D, maps))) return rc;
  static const int window = [] { const char* e = getenv("SDN_FLASH_WINDOW"); const int v = e ? atoi(e) : 0; return v >= 1 && v < kFRing ? v : kFWindow; }();
  static const int mma_sleep = [] { const char* e = getenv("SDN_FLASH_MMA_SLEEP"); return e ? atoi(e) : 0; }();
  static const int poll_sleep = [] { const char* e = getenv("SDN_FLASH_POLL_SLEEP"); return e ? atoi(e) : 100; }();
  for (int64_t q0 = 0; q0 < Q; q0 += kFQ) {
    const int qn = (int)std::min<int64_t>(kFQ, Q - q0);
    FlashArgs a{};
    a.xq = xq + q0 * D; a.sqnorm = sqnorm; a.Q = qn; a.N = (int)N; a.ntiles = (int)cdiv(N, kFR); a.ngroups = nctas / kFGS;
    static const int dbg = [] { const char* e = getenv("SDN_FLASH_DBG"); return e ? atoi(e) : 0; }();
    a.dbg = (unsigned)dbg;
    a.window = window; a.mma_sleep = (unsigned)mma_sleep; a.poll_sleep = (unsigned)poll_sleep;
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
    int e = launch_flash(maps, a, nctas, h->coop_ok, st);
    if (e != 0 && h->coop_ok) {
      // cooperative launch refused: co-residency still holds by the occupancy check above as long as nothing else
      // runs on the device
      cudaGetLastError();
      h->coop_ok = false;
      e = launch_flash(maps, a, nctas, false, st);
    }
    g_prof.end(pid, st);
    if (e != 0) return e;
    SDN_LAUNCHED();
  }
  return SDN_OK;
}

}  // namespace sdn
