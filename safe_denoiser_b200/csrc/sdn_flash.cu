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
//   phase A  (distances) CTA r of a group takes the 128-row tiles t = r (mod 4) over the WHOLE super-slice, in eight
//            chunks of 64 d (one 32 KiB stage each: hi and lo planes of [128 rows][64 d], TMA, L2 evict_last):
//            S[q][i] = sum_{d in super-slice} X[q][d] bank[i][d], tcgen05.mma M128 x N128 x K16 with A = X[:, super-
//            slice] (bf16 hi/lo, resident in TENSOR MEMORY for the whole kernel), B = the chunk (K-major).  The
//            group's 4:1 reduction over d thus happens inside the tensor core.  Round-2 measurements that shaped this:
//            an MMA whose A operand comes from tensor memory costs >= 64 clk whatever N is (A-read bound), so 64-row
//            tiles (N = 64) made phase A alone tensor-bound at 80 us for cfg3 (56 us with 1/16 of the MMAs);
//            N = 128 halves the MMA count per byte.  The accumulator is drained every four chunks (32 MMAs): the
//            tensor core adds with truncation.
//   exchange S[q][i] = sum over the D/512 groups.  The group partials go through L2 as self-validating 16-byte lines
//            {v, tag, v, tag} (no fence, no flag) to work units of 4 rows x 32 queries dealt round-robin over ALL
//            4 x #CTAs warps of the level-2 role: a unit sums the groups in order, computes k = exp(-dist / 2 sigma^2),
//            writes the weights as bf16 hi | lo into the tile's slot of a global ring in the layout phase B wants as
//            an MMA operand, writes its z partial, and then releases the slot: fence + one add on the tile's counter.
//   phase B  (accumulate) CTA j owns the d-slice [128 j, 128 j + 128): num[slice][q] += sum_i bank[i][slice] k[q][i],
//            tcgen05.mma with A = the tile [64 rows][128 d] viewed MN-major, loaded AGAIN by TMA (evict_first) -- a
//            hit in the 126 MB L2 because phase A never runs more than `window` tiles ahead -- and B = the weights
//            [64 rows][hi 64 | lo 64], loaded by TMA from the ring once the tile's counter says all 64 units are
//            done.  Nothing of phase B passes through registers (the first version converted the weights in every CTA
//            and stored them to tensor memory: a serial 1.9 us per 64 rows and CTA, the limiter of that version).
//            The accumulator stays in tensor memory for the whole kernel; the epilogue applies the correction of
//            conditioning() when fused.
//
// Every sum has a fixed order (group index, row index, unit index): results are bit-reproducible run to run.
//
// Warp roles (16 warps): 0 TMA producer phase A | 1 MMA issuer (event loop over "chunk loaded" / "weights and half
// tile loaded") | 2 phase-B producer: polls the tile counters, issues the TMA loads of weights + second read |
// 4-7 query prologue (chunks 0-1), then phase-A drain: TMEM -> registers -> LL lines | 8-11 query prologue (chunks
// 2-3), then z_q from the units' partials, then the final epilogue | 12-15 level-2 work units.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "sdn_internal.h"
#include "sdn_ptx.cuh"

namespace sdn {

constexpr int kFR = 128;                // bank rows per tile (N of the phase-A MMA)
constexpr int kFHB = 64;                // bank rows per phase-B stage (K of its MMAs: 4 steps of 16)
constexpr int kFDS = 128;               // d per CTA in phase B
constexpr int kFGS = 4;                 // CTAs per group
constexpr int kFSuper = kFDS * kFGS;    // d per group in phase A (512)
constexpr int kFChunkD = 64;            // d per phase-A stage
constexpr int kFChunks = kFSuper / kFChunkD;   // 8 chunks per own tile
constexpr int kFQ = 64;                 // query rows per pass
constexpr int kFJobRows = 4;            // rows of a tile per level-2 job
constexpr int kFJobs = kFR / kFJobRows; // 32 jobs per tile, two work units (halves of the query rows) each
constexpr int kFUnits = kFJobs * 2;     // counter increments per tile
constexpr int kFRingL = 16;             // slots of the level-1 partial ring (2 MiB each)
constexpr int kFRingP = 32;             // slots of the weight / z-partial / counter rings
constexpr int kFWindow = 6;             // tiles between the two reads of the bank: 6 x 8 MiB stay in the 126 MB L2.  Measured at
                                        // cfg3 (ncu dram__bytes_read, profiles/r02_flash2_cfg3_window_dram.csv): window 6 -> 234.9 MB
                                        // (1.15 x the algorithmic 205 MB), 7 -> 255.7 MB, 8 -> 281.1 MB (the second read starts to miss)
                                        // while the step goes 132 -> 116 us (SDN_FLASH_WINDOW, <= 13).  Also what makes ring reuse
                                        // safe (2 window + 6 <= kFRingP, window <= kFRingL)
constexpr int kFThreads = 512;
constexpr int kFMaxCtas = 128;
constexpr int kFMaxGroups = kFMaxCtas / kFGS;
constexpr uint32_t kFStageA = 32768;    // hi + lo planes of [128 rows][64 d] bf16
constexpr uint32_t kFStageB = 49152;    // weights [64 rows][128] bf16 (16 KiB) + hi + lo planes of [64 rows][128 d]
constexpr uint32_t kFStageRegion = 224 * 1024;
constexpr int kFMaxStages = 4;
constexpr uint32_t kFTmemCols = 512;
constexpr uint32_t kFColX = 0;          // X operand [128 stacked rows][512 d] bf16: 256 columns
constexpr uint32_t kFColS = 256;        // S accumulator [128][128 bank rows]
constexpr uint32_t kFColAcc = 384;      // num accumulator [128 d][64 hi-part columns | 64 lo-part columns]
constexpr uint32_t kFSpinLimit = 1u << 24;
constexpr int kFPitch = 132;            // floats per staged query row (thread = row accesses are conflict-free per quarter warp)
constexpr uint32_t kFXStageBytes = 4u * kFQ * kFPitch * 4u;   // four chunks of [64 rows][128 d] staged at start-up
constexpr uint32_t kFXStageOff = kFStageRegion - kFXStageBytes;
constexpr uint32_t kFSmemBytes = kFStageRegion + 1024 /*align*/ + 2048 /*barriers, xsq, z*/;

// Arena: library-owned per-device synchronisation memory of the one-pass kernel (zeroed once; tags and counters grow
// monotonically over launches, so stale data never matches).
struct FlashArena {
  uint32_t* epoch;      // [0] sequence number of the next launch's tile 0, [1] launch counter
  uint8_t* xsq_ll;      // [kFMaxCtas][128] 16-byte lines: ||x_q||^2 partial of every CTA
  uint8_t* part_ll;     // [kFRingL][group 32][job 32][q 64][2 lines]   level-2 fan-in
  __nv_bfloat16* gp;    // [kFRingP][128 rows][hi 64 | lo 64]   published weights (phase-B MMA operand)
  float* gz;            // [kFRingP][job 32][q 64]   z partials of the units
  uint32_t* counter;    // [kFRingP]   units done, 64 per tile and round
  uint32_t* diag;       // host-mapped: written before a timeout trap
  unsigned long long* trace;   // SDN_FLASH_TRACE=1: [cta][tile < 64][16 events] globaltimer ns (null otherwise)
};

struct FlashArgs {
  const float* xq;      // [Q][D] query the distances are taken on
  const float* sqnorm;  // [N]
  int Q, N, ntiles, nhb, ngroups, window, nsa, nsb;
  unsigned poll_sleep;                // back-off of the polling loops (ns)
  unsigned dbg;                       // SDN_FLASH_DBG experiments (wrong results): 1 = one phase-A MMA per chunk, 2 = one phase-B MMA per stage
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

__device__ __forceinline__ void f_st_release_cta(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(u_smem(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int f_ld_acquire_cta(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(u_smem(p)) : "memory");
  return v;
}

struct FSmem {
  uint8_t* stages;      // [nsa phase-A stages x 32 KiB | nsb phase-B stages x 48 KiB]
  uint64_t* afull; uint64_t* aempty;      // [kFMaxStages]
  uint64_t* bfull; uint64_t* bempty;      // [kFMaxStages]
  uint64_t* sfull; uint64_t* sempty;      // [1] S accumulator (one drain per four chunks)
  uint64_t* xfull; uint64_t* accfull;     // [1]
  uint64_t* xload;                        // [8 warps][2 chunks] query rows of one prologue warp staged in shared memory
  uint32_t* tmem_base;
  int* pub_upto;        // tiles < this have all their weights published (phase-B producer -> z warps)
  int* zdone;           // tiles < this have their z partials summed (z warps -> MMA issuer: ring reuse)
  float* xsq;           // [64]
  float* xsq_half;      // [128]  (||x||^2 halves at start, z_q halves at the end)
  float* zq;            // [64]
};

__device__ __forceinline__ FSmem f_carve(unsigned char* raw) {
  FSmem s;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023;
  s.stages = reinterpret_cast<uint8_t*>(a);
  uint64_t* b = reinterpret_cast<uint64_t*>(s.stages + kFStageRegion);
  s.afull = b; b += kFMaxStages;
  s.aempty = b; b += kFMaxStages;
  s.bfull = b; b += kFMaxStages;
  s.bempty = b; b += kFMaxStages;
  s.sfull = b; b += 1;
  s.sempty = b; b += 1;
  s.xfull = b; b += 1;
  s.accfull = b; b += 1;
  s.xload = b; b += 16;
  s.tmem_base = reinterpret_cast<uint32_t*>(b); b += 1;
  s.pub_upto = reinterpret_cast<int*>(b); b += 1;
  s.zdone = reinterpret_cast<int*>(b); b += 1;     // 39 x 8 = 312 bytes so far
  s.xsq = reinterpret_cast<float*>(b);
  s.xsq_half = s.xsq + 64;
  s.zq = s.xsq_half + 128;                        // + 256 floats = 1336 bytes
  return s;
}

__global__ void __launch_bounds__(kFThreads, 1)
k_flash(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
        const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ FlashArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const FSmem sm = f_carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const int grp = cta / kFGS, rk = cta % kFGS;
  const int d0 = cta * kFDS;            // phase-B slice
  const int ds0 = grp * kFSuper;        // phase-A super-slice
  const int ntiles = a.ntiles, nhb = a.nhb;
  const int nsa = a.nsa, nsb = a.nsb;
  const int ngroups = a.ngroups;
  uint8_t* const bstages = sm.stages + (size_t)nsa * kFStageA;
  // query staging (four buffers of [64 rows][132 floats]) lives at the end of the stage region until the prologue is
  // done: the stages that overlap it take their first load after `xfull`
  float* const xstage = reinterpret_cast<float*>(sm.stages + kFXStageOff);
  const int first_gated_a = (int)(kFXStageOff / kFStageA);   // phase-A stages >= this overlap the staging area

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFMaxStages; ++s) {
      u_mbar_init(&sm.afull[s], 1); u_mbar_init(&sm.aempty[s], 1);
      u_mbar_init(&sm.bfull[s], 1); u_mbar_init(&sm.bempty[s], 1);
    }
    u_mbar_init(sm.sfull, 1); u_mbar_init(sm.sempty, 4);
    u_mbar_init(sm.xfull, 8); u_mbar_init(sm.accfull, 1);
    for (int w = 0; w < 16; ++w) u_mbar_init(&sm.xload[w], 1);
    *sm.pub_upto = 0; *sm.zdone = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    u_prefetch_map(&tm_hi); u_prefetch_map(&tm_lo); u_prefetch_map(&tm_p);
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
  // phase A unit ua = (own tile i, chunk c): tile t = rk + 4 i, d in [ds0 + 64 c, +64)
  const int nown = (ntiles - rk + kFGS - 1) / kFGS;
  const int nua = nown * kFChunks;

  if (warp == 0) {
    // ============================================================ TMA producer, phase A (HBM stream)
    if (lane == 0) {
      for (int ua = 0; ua < nua; ++ua) {
        const int s = ua % nsa;
        if (ua >= nsa) f_wait(a, &sm.aempty[s], (uint32_t)(((ua / nsa) + 1) & 1), 0x100, ua);
        else if (s >= first_gated_a) f_wait(a, sm.xfull, 0, 0x101, ua);     // this stage doubles as query staging
        const int t = rk + kFGS * (ua / kFChunks), c = ua % kFChunks;
        if (c == 0) f_trace(a, t, 0);
        if (c == 4) f_trace(a, t, 14);
        if (c == 7) f_trace(a, t, 15);
        uint8_t* st = sm.stages + (size_t)s * kFStageA;
        const int r0 = t * kFR, dd = ds0 + c * kFChunkD;
        // a stage: hi and lo planes of [128 rows][64 d], two boxes of [64 rows][64 d] per plane (8 KiB each); a box
        // entirely past the last bank row is skipped (its S columns are never used)
        const bool two = r0 + 64 < a.N;
        u_mbar_expect_tx(&sm.afull[s], two ? kFStageA : kFStageA / 2);
        f_tma_2d_hint(st, &tm_hi, dd, r0, &sm.afull[s], pol_keep);
        f_tma_2d_hint(st + 16384, &tm_lo, dd, r0, &sm.afull[s], pol_keep);
        if (two) {
          f_tma_2d_hint(st + 8192, &tm_hi, dd, r0 + 64, &sm.afull[s], pol_keep);
          f_tma_2d_hint(st + 24576, &tm_lo, dd, r0 + 64, &sm.afull[s], pol_keep);
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idA = u_idesc(128, kFR, 0, 0);       // S[128 stacked q][128 rows] = X (TMEM) x chunk (K-major)
      constexpr uint32_t idBf = u_idesc(kFDS, 128, 1, 1);     // num[128 d][hi | lo] += bank_hi^T x [P_hi | P_lo]
      constexpr uint32_t idBh = u_idesc(kFDS, 64, 1, 1);      // num[128 d][hi]      += bank_lo^T x P_hi
      f_wait(a, sm.xfull, 0, 0x200, 0);
      u_fence_after();
      int ua = 0, hb = 0;                 // next phase-A unit (own tile, chunk) / next phase-B half tile
      uint32_t idle = 0;
      while (hb < nhb || ua < nua) {
        bool did = false;
        if (hb < nhb && u_mbar_test(&sm.bfull[hb % nsb], (uint32_t)((hb / nsb) & 1))) {
          // ---- phase B of half tile hb: 4 K steps of 16 bank rows
          u_fence_after();
          f_trace(a, hb >> 1, (hb & 1) ? 13 : 8);
          const uint32_t base = u_smem(bstages + (size_t)(hb % nsb) * kFStageB);
          const uint32_t acc = tmem + kFColAcc;
#pragma unroll
          for (int kk = 0; kk < kFHB / 16; ++kk) {
            if ((a.dbg & 2) && kk > 0) break;
            const uint64_t pw = u_desc(base + kk * 2048, 8192, 1024);               // P^T, MN-major: hi box | lo box
            const uint64_t ah = u_desc(base + 16384 + kk * 2048, 8192, 1024);       // bank^T, MN-major
            const uint64_t al = u_desc(base + 32768 + kk * 2048, 8192, 1024);
            u_mma(acc, ah, pw, idBf, (hb > 0 || kk > 0) ? 1u : 0u);
            u_mma(acc, al, pw, idBh, 1u);
          }
          u_commit(&sm.bempty[hb % nsb]);
          ++hb;
          if (hb == nhb) u_commit(sm.accfull);
          did = true;
        }
        // phase A stays within `window` tiles of phase B (and of the z warps): the exchange rings are reused safely and
        // the tiles waiting for their second read fit in the L2.  The S accumulator is drained every four chunks.
        if (ua < nua) {
          const int ta = rk + kFGS * (ua / kFChunks), c = ua % kFChunks;
          const int sh = ua >> 2;           // drains so far = halves started
          const int tdone = min(hb >> 1, *reinterpret_cast<volatile int*>(sm.zdone));
          if (ta - tdone < a.window && u_mbar_test(&sm.afull[ua % nsa], (uint32_t)((ua / nsa) & 1)) &&
              ((c & 3) || sh == 0 || u_mbar_test(sm.sempty, (uint32_t)((sh + 1) & 1)))) {
            u_fence_after();
            if (c == 0) f_trace(a, ta, 1);
            if (c == 4) f_trace(a, ta, 7);
            if (c == 7) f_trace(a, ta, 12);
            const uint32_t base = u_smem(sm.stages + (size_t)(ua % nsa) * kFStageA);
            const uint32_t acc = tmem + kFColS;
#pragma unroll
            for (int kk = 0; kk < kFChunkD / 16; ++kk) {
              if ((a.dbg & 1) && kk > 0) break;
              const uint64_t bh = u_desc(base + kk * 32, 16, 1024);
              const uint64_t bl = u_desc(base + 16384 + kk * 32, 16, 1024);
              const uint32_t xa = tmem + kFColX + (uint32_t)(c * 4 + kk) * 8;
              u_mma_ts(acc, xa, bh, idA, ((c & 3) || kk > 0) ? 1u : 0u);
              u_mma_ts(acc, xa, bl, idA, 1u);
            }
            if ((c & 3) == 3) u_commit(sm.sfull);
            u_commit(&sm.aempty[ua % nsa]);
            ++ua;
            did = true;
          }
        }
        if (did) idle = 0;
        else if (++idle > kFSpinLimit) f_timeout(a, 0x210, (uint32_t)ua, (uint32_t)hb);
      }
    }
  } else if (warp == 2) {
    // ============================================================ phase-B producer: weights + second read of the tile
    // The phase-B stages double as the staging area of the query prologue: no TMA into them before it is done.
    if (lane == 0) f_wait(a, sm.xfull, 0, 0x122, 0);
    __syncwarp();
    int hb = 0;
    uint32_t spin = 0;
    while (hb < nhb) {
      // lanes 0..7 look at the counters of the next tiles: one round trip tells how far the weights are published
      const int t0 = hb >> 1;
      int ok = 0;
      if (lane < 8 && t0 + lane < ntiles) {
        const uint32_t seq = epoch0 + (uint32_t)(t0 + lane);
        const uint32_t c = u_ld_acquire(a.ar.counter + (seq % kFRingP));
        ok = (int32_t)(c - (seq / kFRingP) * (uint32_t)kFUnits) >= 0;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      __syncwarp();                               // lane 0 issues the loads on behalf of the lanes that acquired
      const int nready = __ffs(~m) - 1;           // consecutive published tiles from t0 on
      if (nready == 0) {
        if (a.poll_sleep) __nanosleep(a.poll_sleep);
        if (++spin > kFSpinLimit) f_timeout(a, 0x400, (uint32_t)t0, 0);
        continue;
      }
      spin = 0;
      if (lane == 0) {
        f_st_release_cta(sm.pub_upto, t0 + nready);
        asm volatile("fence.proxy.async;" ::: "memory");      // the weights were written by generic stores of other SMs
        const int hb_end = min(nhb, (t0 + nready) * 2);
        for (; hb < hb_end; ++hb) {
          const int s = hb % nsb;
          if (hb >= nsb) f_wait(a, &sm.bempty[s], (uint32_t)(((hb / nsb) + 1) & 1), 0x121, hb);
          if (!(hb & 1)) f_trace(a, hb >> 1, 6);
          uint8_t* st = bstages + (size_t)s * kFStageB;
          const uint32_t seq = epoch0 + (uint32_t)(hb >> 1);
          const int prow = (int)(seq % kFRingP) * kFR + (hb & 1) * kFHB;
          const int r0 = hb * kFHB;
          u_mbar_expect_tx(&sm.bfull[s], kFStageB);
          u_tma_2d(st, &tm_p, 0, prow, &sm.bfull[s]);
          u_tma_2d(st + 8192, &tm_p, 64, prow, &sm.bfull[s]);
          f_tma_2d_hint(st + 16384, &tm_hi, d0, r0, &sm.bfull[s], pol_drop);
          f_tma_2d_hint(st + 24576, &tm_hi, d0 + 64, r0, &sm.bfull[s], pol_drop);
          f_tma_2d_hint(st + 32768, &tm_lo, d0, r0, &sm.bfull[s], pol_drop);
          f_tma_2d_hint(st + 40960, &tm_lo, d0 + 64, r0, &sm.bfull[s], pol_drop);
        }
      }
      hb = __shfl_sync(0xffffffffu, hb, 0);
    }
  } else if (warp >= 4 && warp < 12) {
    const int lq = warp & 3;
    const int set = (warp - 4) >> 2;                    // 0: warps 4-7, 1: warps 8-11
    const int part = lane >> 4;                         // lanes 0-15 hold hi parts, 16-31 lo parts of the same q
    const int q = lq * 16 + (lane & 15);
    const uint32_t tlane = tmem + ((uint32_t)(lq * 32) << 16);
    // ============================================================ query prologue (8 warps)
    // X[:, super-slice] -> bf16 hi / lo -> tensor memory: warp set s converts the 128-wide chunks 2s and 2s + 1 of its
    // 16 query rows, staged through shared memory (whole 512-byte rows by bulk copies, one per lane, then thread =
    // row: a warp-wide global load of 32 different rows costs 32 L1 wavefronts per instruction); ||x_q||^2 partial of
    // this CTA's own slice
    {
      const int row0 = lq * 16;
      const int nvalid = max(0, min(16, a.Q - row0));
      const bool valid = q < a.Q;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = set * 2 + cc;
        uint64_t* bar = &sm.xload[(warp - 4) * 2 + cc];
        if (lane == 0) u_mbar_expect_tx(bar, (uint32_t)nvalid * (kFDS * 4));
        __syncwarp();
        if (lane < nvalid) {
          const int r = row0 + lane;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(u_smem(xstage + ((size_t)c * kFQ + r) * kFPitch)),
                         "l"(a.xq + (int64_t)r * a.D + ds0 + c * kFDS), "r"(kFDS * 4), "r"(u_smem(bar)) : "memory");
        }
      }
      float ss = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = set * 2 + cc;
        if (nvalid > 0) f_wait(a, &sm.xload[(warp - 4) * 2 + cc], 0, 0x340, c);
        const float* xr = xstage + ((size_t)c * kFQ + q) * kFPitch;
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
      }
      u_tmem_st_wait();
      if (part == 0 && set == (rk >> 1))
        u_ll_store(a.ar.xsq_ll + ((size_t)cta * 128 + q) * 16, ss, 0.f, launch0);
      u_fence_before();
      __syncwarp();
      if (lane == 0) u_mbar_arrive(sm.xfull);
    }
    if (set == 0) {
      // ========================================================== phase-A drain
      // S [128 stacked rows][128 bank rows] of an own tile, twice (after chunks 3 and 7) -> registers -> LL lines for
      // the level-2 work units.  Layout per (ring slot, group): [job 32][q 64][2 lines of 2 rows].
#pragma unroll 1
      for (int i = 0; i < nown; ++i) {
        const int t = rk + kFGS * i;
        float acc[64];
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int sh = i * 2 + h;
          f_wait(a, sm.sfull, (uint32_t)(sh & 1), 0x310, t);
          u_fence_after();
          if (warp == 4 && lane == 0) f_trace(a, t, h ? 2 : 3);
          // hi-part row + lo-part row of the same query; lanes l / l+16 keep bank rows [64 ch, +32) / [64 ch + 32, +32)
#pragma unroll
          for (int cq = 0; cq < 4; ++cq) {                 // 32 columns at a time: ch = cq >> 1, kept by part = cq & 1
            uint32_t r[32];
            u_tmem_ld32_nowait(tlane + kFColS + cq * 32, r);
            u_tmem_ld_wait();
            if (cq == 3) {
              u_fence_before();
              __syncwarp();
              if (lane == 0) u_mbar_arrive(sm.sempty);
              if (warp == 4 && lane == 0 && h == 0) f_trace(a, t, 4);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = __uint_as_float(r[j]) + __shfl_xor_sync(0xffffffffu, __uint_as_float(r[j]), 16);
              if (part == (cq & 1)) acc[(cq >> 1) * 32 + j] = h ? acc[(cq >> 1) * 32 + j] + v : v;
            }
          }
        }
        const uint32_t tag = epoch0 + (uint32_t)t;            // the tile's sequence number (gap-free over launches)
        uint8_t* dst = a.ar.part_ll + ((size_t)(tag % kFRingL) * kFMaxGroups + grp) * (size_t)(kFJobs * 2048) + (size_t)q * 32;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
          for (int sub = 0; sub < 8; ++sub) {
            uint8_t* o = dst + (size_t)(ch * 16 + part * 8 + sub) * 2048;
            u_ll_store(o, acc[ch * 32 + 4 * sub], acc[ch * 32 + 4 * sub + 1], tag);
            u_ll_store(o + 16, acc[ch * 32 + 4 * sub + 2], acc[ch * 32 + 4 * sub + 3], tag);
          }
        }
        if (warp == 4 && lane == 0) f_trace(a, t, 5);
      }
    } else {
      // ========================================================== z_q from the units' partials, then the final epilogue
      const int tid = (warp - 8) * 32 + lane;             // 0..127
      const int zq_q = tid & 63, zhalf = tid >> 6;        // thread = (query row, half of the jobs)
      float z = 0.f;
#pragma unroll 1
      for (int t = 0; t < ntiles;) {
        int upto;
        uint32_t spin = 0;
        while ((upto = f_ld_acquire_cta(sm.pub_upto)) <= t) {
          __nanosleep(200);
          if (++spin > kFSpinLimit) f_timeout(a, 0x430, (uint32_t)t, 0);
        }
        for (; t < upto; ++t) {
          const uint32_t seq = epoch0 + (uint32_t)t;
          const float* zp = a.ar.gz + ((size_t)(seq % kFRingP) * kFJobs + zhalf * (kFJobs / 2)) * 64 + zq_q;
          float v[kFJobs / 2];
#pragma unroll
          for (int j = 0; j < kFJobs / 2; ++j) {
            uint32_t w;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w) : "l"(zp + (size_t)j * 64) : "memory");
            v[j] = __uint_as_float(w);
          }
#pragma unroll
          for (int j = 0; j < kFJobs / 2; ++j) z += v[j];
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");      // every z thread has read the slots of the tiles < t
        if (tid == 0) f_st_release_cta(sm.zdone, t);
      }
      sm.xsq_half[tid] = z;
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tid < 64) sm.zq[tid] = sm.xsq_half[tid] + sm.xsq_half[64 + tid];
      asm volatile("bar.sync 2, 128;" ::: "memory");

      // ---- final epilogue: TMEM lane = d within the slice, column = query row (hi-part sums | lo-part sums)
      f_wait(a, sm.accfull, 0, 0x420, ntiles);
      u_fence_after();
      const int nrows = min(a.Q, kFQ);
      const int64_t d = (int64_t)d0 + lq * 32 + lane;
      float msum = 0.f;
#pragma unroll 1
      for (int c = 0; c < kFQ / 32; ++c) {
        float va[32];
        {
          float vb[32];
          u_tmem_ld32(tlane + kFColAcc + (uint32_t)(c * 32), va);            // hi*P_hi + lo*P_hi, queries [32c, 32c+32)
          u_tmem_ld32(tlane + kFColAcc + (uint32_t)(kFQ + c * 32), vb);      // hi*P_lo
#pragma unroll
          for (int j = 0; j < 32; ++j) va[j] += vb[j];
        }
        if (a.epi.fused) {
          // all loads of the chunk first: the stores below may alias them as far as the compiler knows
          float xv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int qq = min(c * 32 + j, nrows - 1);
            xv[j] = a.epi.x0 ? __ldcg(a.epi.x0 + (int64_t)qq * a.D + d) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int qq = c * 32 + j;
            if (qq < nrows) {
              const int64_t o = (int64_t)qq * a.D + d;
              const float n = va[j] / (sm.zq[qq] + a.epi.eps);
              if (a.num_out) a.num_out[o] = va[j];
              if (a.epi.neg_out) a.epi.neg_out[o] = n;
              if (a.epi.x0) a.epi.x0[o] = fmaf(-a.epi.scale, n, xv[j]);
              msum += fminf(fmaxf(n, -1e10f), 1e10f);
            }
          }
        } else if (a.num_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int qq = c * 32 + j;
            if (qq < nrows) a.num_out[(int64_t)qq * a.D + d] = va[j];
          }
        }
      }
      if (cta == 0) {
        for (int r = tid; r < nrows; r += 128) {
          const float zr = sm.zq[r], denom = zr + a.epi.eps;
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
    //      group has delivered, so the tiles between two units of one worker bound the rate: 8 here).  Lane l reads
    //      lines l and 32 + l of the unit's 1 KiB block of every group (512 contiguous bytes per warp instruction)
    //      = (query row, row pair) twice, and sums the groups in order.
    {
      const int lh = lane & 1;
      const int nunits = ntiles * kFUnits;
      const int nworkers = nctas * 4;
#pragma unroll 1
      for (int U = cta * 4 + ow; U < nunits; U += nworkers) {
        const int t = U / kFUnits, rem = U - t * kFUnits;
        const int j = rem >> 1, qh = rem & 1;
        const uint32_t tag = epoch0 + (uint32_t)t;
        const int ringl = (int)(tag % kFRingL), ringp = (int)(tag % kFRingP);
        const int trow = j * kFJobRows + lh * 2;          // row within the tile
        const int row = t * kFR + trow;
        // ||n||^2 of the two rows: issued before the wait, the line is cold
        const float sq0 = row < a.N ? __ldg(a.sqnorm + row) : 0.f;
        const float sq1 = row + 1 < a.N ? __ldg(a.sqnorm + row + 1) : 0.f;
        const size_t g_stride = (size_t)kFJobs * 2048;
        const uint8_t* src0 = a.ar.part_ll + (size_t)ringl * kFMaxGroups * g_stride + (size_t)j * 2048 +
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
        __nv_bfloat16* prow = a.ar.gp + ((size_t)ringp * kFR + trow) * 128;
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
          // bf16 hi | lo, [row][hi parts of the 64 queries | lo parts]: the B operand of phase B as TMA loads it
          const __nv_bfloat16 h0 = __float2bfloat16_rn(k0v), h1 = __float2bfloat16_rn(k1v);
          prow[q] = h0;
          prow[64 + q] = __float2bfloat16_rn(k0v - __bfloat162float(h0));
          prow[128 + q] = h1;
          prow[192 + q] = __float2bfloat16_rn(k1v - __bfloat162float(h1));
          // z partial of the unit: its four rows in order (two here, two in the neighbour lane)
          float zu = k0v + k1v;
          const float zo = __shfl_xor_sync(0xffffffffu, zu, 1);
          if (lh == 0) a.ar.gz[((size_t)ringp * kFJobs + j) * 64 + q] = zu + zo;      // (rows 0,1) + (rows 2,3)
          if (a.k_out && q < a.Q) {
            if (row < a.N) a.k_out[(int64_t)q * a.N + row] = k0v;
            if (row + 1 < a.N) a.k_out[(int64_t)q * a.N + row + 1] = k1v;
          }
        }
        // release the unit: every lane's stores happen before lane 0's gpu-scope fence, which happens before the add
        __syncwarp();
        if (lane == 0) {
          u_fence_gpu();
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(a.ar.counter + ringp) : "memory");
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
    *reinterpret_cast<volatile uint32_t*>(a.ar.epoch) = epoch0 + (uint32_t)ntiles;      // gap-free: ring slot = seq mod ring
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
  CUtensorMap tm_p{};     // weight ring [kFRingP x 128 rows][128] bf16, boxes [64 rows][64]
  int max_ctas = 0;       // co-resident CTAs of k_flash on this device
  bool coop_ok = true;
};
ArenaHost g_arena[kMaxDevices];

constexpr size_t kXsqBytes = (size_t)kFMaxCtas * 128 * 16;
constexpr size_t kPartBytes = (size_t)kFRingL * kFMaxGroups * kFJobs * 2048;
constexpr size_t kGpBytes = (size_t)kFRingP * kFR * 128 * 2;
constexpr size_t kGzBytes = (size_t)kFRingP * kFJobs * 64 * 4;

int arena_get(int dev, ArenaHost** out) {
  ArenaHost& h = g_arena[dev];
  std::lock_guard<std::mutex> lk(h.mu);
  if (!h.ready) {
    const size_t total = 256 + 256 + kXsqBytes + kPartBytes + kGpBytes + kGzBytes;
    SDN_CUDA_OK(cudaMalloc(&h.dev, total));
    SDN_CUDA_OK(cudaMemset(h.dev, 0, total));
    uint8_t* p = static_cast<uint8_t*>(h.dev);
    h.ar.epoch = reinterpret_cast<uint32_t*>(p); p += 256;
    h.ar.counter = reinterpret_cast<uint32_t*>(p); p += 256;
    h.ar.xsq_ll = p; p += kXsqBytes;
    h.ar.part_ll = p; p += kPartBytes;
    h.ar.gp = reinterpret_cast<__nv_bfloat16*>(p); p += kGpBytes;
    h.ar.gz = reinterpret_cast<float*>(p); p += kGzBytes;
    {
      const int rc = tmap_bf16_2d(&h.tm_p, h.ar.gp, (uint64_t)kFRingP * kFR, 128, kFHB, 64);
      if (rc) return rc;
    }
    const uint32_t one[2] = {kFRingP, 1};    // sequence numbers start in round 1: zeroed LL tags never match, counters reach 64 x round
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
  if ((rc = tmap_bf16_2d(&m.m[0], h, (uint64_t)N, (uint64_t)D, kFHB, 64))) return rc;
  if ((rc = tmap_bf16_2d(&m.m[1], h + N * D, (uint64_t)N, (uint64_t)D, kFHB, 64))) return rc;
  m.dev = dev; m.planes = planes; m.N = N; m.D = D; m.stamp = ++g_maps_clock;
  g_maps[slot] = m;
  if (slot == g_maps_n) ++g_maps_n;
  memcpy(out2, m.m, sizeof(m.m));
  return SDN_OK;
}

int launch_flash(const CUtensorMap* m, const CUtensorMap& mp, const FlashArgs& a, int nctas, bool coop, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nctas);
  cfg.blockDim = dim3(kFThreads);
  cfg.dynamicSmemBytes = kFSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;     // every CTA must be resident: they wait on one another
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = coop ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, k_flash, m[0], m[1], mp, a);
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
  // SDN_FLASH_WINDOW may only shrink the window: ring reuse is safe for window <= kFWindow
  static const int window = [] { const char* e = getenv("SDN_FLASH_WINDOW"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 13 ? v : kFWindow; }();
  // split of the 224 KiB stage region: nsa phase-A stages of 32 KiB, the rest phase-B stages of 48 KiB
  static const int nsa = [] { const char* e = getenv("SDN_FLASH_NSA"); const int v = e ? atoi(e) : 0; return v >= 2 && v <= 4 ? v : 4; }();
  static const int poll_sleep = [] { const char* e = getenv("SDN_FLASH_POLL_SLEEP"); return e ? atoi(e) : 100; }();
  for (int64_t q0 = 0; q0 < Q; q0 += kFQ) {
    const int qn = (int)std::min<int64_t>(kFQ, Q - q0);
    FlashArgs a{};
    a.xq = xq + q0 * D; a.sqnorm = sqnorm; a.Q = qn; a.N = (int)N; a.ntiles = (int)cdiv(N, kFR); a.nhb = (int)cdiv(N, kFHB);
    a.ngroups = nctas / kFGS;
    a.nsa = nsa; a.nsb = std::min<int>(kFMaxStages, (int)((kFStageRegion - (uint32_t)nsa * kFStageA) / kFStageB));
    static const int dbg = [] { const char* e = getenv("SDN_FLASH_DBG"); return e ? atoi(e) : 0; }();
    a.dbg = (unsigned)dbg;
    a.window = window; a.poll_sleep = (unsigned)poll_sleep;
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
    int e = launch_flash(maps, h->tm_p, a, nctas, h->coop_ok, st);
    if (e != 0 && h->coop_ok) {
      // cooperative launch refused: co-residency still holds by the occupancy check above as long as nothing else
      // runs on the device
      cudaGetLastError();
      h->coop_ok = false;
      e = launch_flash(maps, h->tm_p, a, nctas, false, st);
    }
    g_prof.end(pid, st);
    if (e != 0) return e;
    SDN_LAUNCHED();
  }
  return SDN_OK;
}

}  // namespace sdn
