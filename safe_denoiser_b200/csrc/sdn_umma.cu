// tcgen05 / TMEM / TMA path for batched calls (8 < Q <= 64): the two contractions of the projection run on
// the 5th-generation tensor cores.
//
// Operands are bf16 hi/lo pairs (v = hi + lo to ~2^-17): three bf16 MMAs per contraction give fp32-class
// dot products, which the un-squared distance needs ("exactness vs tensor cores").  The bank planes
// are made once by sdn_bank_prepare; the query planes and the weight planes are made per call.
//
//   k_umma_xprep    xq [Q,D] fp32 -> X planes [128][D] bf16 (rows 0..63 hi, 64..127 lo, zero padded)
//   k_umma_dots     phase A   S^T[i][r] += sum_d X[r][d] * (hi+lo)[i][d]      A = X (K-major), B = bank rows
//                   (K-major), accumulator [128 stacked query rows x 128 bank rows] in TMEM, split-K over D
//   k_umma_weights  k_qi = exp(-dist/2sigma^2) from S^T, z_q, and the weight planes P [Npad][128] bf16
//   k_umma_accum    phase B   num[q][d] = sum_i P[q][i] * (hi+lo)[i][d]      A = bank^T (MN-major, straight
//                   from the row-major planes), B = P (MN-major), accumulator [128 d x 128 stacked q] in TMEM
//
// Both GEMM kernels: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (tcgen05.ld -> global).  All smem tiles are 128B-swizzled as written by TMA.
//
// Replaces repellency_methods_fast.py:249-250 (cdist + the [Q,N,D+1] broadcast) for batched queries.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "sdn_internal.h"
#include "sdn_ptx.cuh"

namespace sdn {

constexpr int kUK = 64;          // bf16 elements per K block = one 128-byte swizzle row
constexpr int kUStack = 128;     // stacked query rows: [0,64) hi parts, [64,128) lo parts
constexpr int kUQ = 64;          // max query rows per launch
constexpr int kUBankTile = 128;  // bank rows per phase-A tile (MMA N)
constexpr int kUDBlock = 128;    // d per phase-B CTA (MMA M)
constexpr int kUMaxStages = 4;
constexpr uint32_t kTileBytes = 128 * 128;   // a [128 rows][64 bf16] tile = 16 KiB
constexpr uint32_t kPipeBytes = 12 * kTileBytes;   // 192 KiB of pipeline stages, however they are cut
constexpr size_t kUSmemBytes = kPipeBytes + 1024 /*align*/ + 256 /*barriers*/;

// One pass serves G groups of 64 query rows (G = 2 when a call brings more than 64): the bank tile of a stage is
// shared by the groups, so 65..128 query rows cost one read of the bank per phase instead of two.  A stage holds
// G query-side tiles (X planes in phase A, weight planes in phase B) + the hi and lo bank tiles; every group has
// its own TMEM accumulator pair and its own four epilogue warps.
template <int G>
struct UCfg {
  static constexpr int kStages = G == 1 ? 4 : 3;                       // 4 x 48 KiB or 3 x 64 KiB
  static constexpr uint32_t kStageBytes = (uint32_t)(G + 2) * kTileBytes;
  static constexpr uint32_t kHiOff = (uint32_t)G * kTileBytes;         // bank hi tile within a stage
  static constexpr uint32_t kLoOff = (uint32_t)(G + 1) * kTileBytes;   // bank lo tile
  static constexpr int kThreads = 64 + 128 * G;                        // TMA warp, MMA warp, 4 G epilogue warps
  static constexpr uint32_t kAccCols = (uint32_t)G * 128;              // TMEM columns of one accumulator buffer
  static_assert(kStages * kStageBytes == kPipeBytes, "stages must fill the pipeline area");
};

// Programmatic dependent launch (round 2): the kernels of a pass are launched with the stream-serialisation attribute,
// every one lets its successor start at once (`launch_dependents` on entry) and waits for its predecessor's memory
// (`wait`) only where it first needs it -- so the successor's block scheduling, barrier/TMEM set-up and the first
// TMA loads of BANK tiles (which no kernel of the pass writes) overlap the predecessor's tail.  Both instructions are
// no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct USmem {
  uint8_t* tiles;       // [stages][G + 2][16 KiB], 1024-byte aligned
  uint64_t* full;       // [stages]
  uint64_t* empty;      // [stages]
  uint64_t* acc_full;   // [2]
  uint64_t* acc_empty;  // [2]
  uint32_t* tmem_base;  // [1]
};
__device__ __forceinline__ USmem u_carve(unsigned char* raw) {
  USmem s;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023;
  s.tiles = reinterpret_cast<uint8_t*>(a);
  s.full = reinterpret_cast<uint64_t*>(s.tiles + kPipeBytes);
  s.empty = s.full + kUMaxStages;
  s.acc_full = s.empty + kUMaxStages;
  s.acc_empty = s.acc_full + 2;
  s.tmem_base = reinterpret_cast<uint32_t*>(s.acc_empty + 2);
  return s;
}

constexpr int kListCap = 32;     // significant rows per query the listed accumulate handles

struct SigLists {
  int* flags;        // [Npad/64]      row block holds a significant weight
  int* count;        // [64]           significant rows per query (may exceed kListCap)
  int ncount;        //                entries of `count` the start-of-pass clear covers (64 G)
  int* dense;        // [1]            set by k_umma_zreduce when some query's weights are flat (z / kmax > cap):
                     //                its list must overflow, so nobody builds lists or flags and phase B is dense
  int* rows;         // [64][kListCap] their bank row indices (unsorted, first kListCap arrivals)
  float* ks;         // [64][kListCap] their weights
};

// zeroes the lists (launched as part of k_umma_weights' grid: block 0 does it before anyone appends)
__device__ __forceinline__ void siglist_clear(const SigLists& L, int nflags) {
  for (int i = threadIdx.x; i < nflags; i += blockDim.x) L.flags[i] = 0;
  if (threadIdx.x < L.ncount) L.count[threadIdx.x] = 0;
  if (threadIdx.x == 0) *L.dense = 0;
}

// What the first query-prepare block of a pass resets before anything else of the pass runs: the significant-row
// lists and the arrival counters of phase A's fused weights step.
struct StartClear {
  SigLists lists; int nflags;      // lists.flags == nullptr: no lists in this pass
  int* counters; int ncounters;    // [row tiles + 1]
};
__device__ __forceinline__ void start_clear(const StartClear& c) {
  if (c.lists.flags) siglist_clear(c.lists, c.nflags);
  for (int i = threadIdx.x; i < c.ncounters; i += blockDim.x) c.counters[i] = 0;
}

// ------------------------------------------------------------------------------------------ query planes
// bf16 hi/lo planes of 64 query rows: hi part of row q at plane row q, lo part at plane row lo_rows + q.
//   one group per pass : [64 hi | 64 lo] rows = ONE stacked 128-row MMA operand           (lo_rows = 64)
//   two groups per pass: [128 hi] [128 lo]   = two 128-row operands, group g at rows 64 g (lo_rows = 128)
__global__ void __launch_bounds__(256)
k_umma_xprep(const float* __restrict__ xq, int Q, int64_t D, __nv_bfloat16* __restrict__ planes, int lo_rows,
             const StartClear clr) {
  pdl_launch_dependents();
  if (blockIdx.x == 0 && blockIdx.y == 0) start_clear(clr);
  const int q = blockIdx.y;   // 0..63
  const int64_t j = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (j >= D) return;
  __nv_bfloat16 h[4], l[4];
  if (q < Q) {
    const float4 v = *reinterpret_cast<const float4*>(xq + (int64_t)q * D + j);
    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      h[u] = __float2bfloat16_rn(f[u]);
      l[u] = __float2bfloat16_rn(f[u] - __bfloat162float(h[u]));
    }
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) h[u] = l[u] = __float2bfloat16_rn(0.f);
  }
  *reinterpret_cast<uint2*>(planes + (int64_t)q * D + j) = *reinterpret_cast<const uint2*>(h);
  *reinterpret_cast<uint2*>(planes + (int64_t)(lo_rows + q) * D + j) = *reinterpret_cast<const uint2*>(l);
}

// Same, plus ||x_q||^2 partials (one per 1024-element chunk, summed by k_umma_weights) so that the batched
// conditioning call needs no separate query-prepare launch.  grid (D/1024, 64).
__global__ void __launch_bounds__(256)
k_umma_qprep(const float* __restrict__ xq, int Q, int64_t D, __nv_bfloat16* __restrict__ planes, int lo_rows,
             float* __restrict__ xsq_part, float* __restrict__ zero_word, const StartClear clr) {
  __shared__ float red[33];
  pdl_launch_dependents();
  if (blockIdx.x == 0 && blockIdx.y == 0) start_clear(clr);
  if (zero_word && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *zero_word = 0.f;
  const int q = blockIdx.y;
  const int64_t j = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  float ss = 0.f;
  if (j < D) {
    __nv_bfloat16 h[4], l[4];
    if (q < Q) {
      const float4 v = *reinterpret_cast<const float4*>(xq + (int64_t)q * D + j);
      const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        h[u] = __float2bfloat16_rn(f[u]);
        l[u] = __float2bfloat16_rn(f[u] - __bfloat162float(h[u]));
        ss = fmaf(f[u], f[u], ss);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) h[u] = l[u] = __float2bfloat16_rn(0.f);
    }
    *reinterpret_cast<uint2*>(planes + (int64_t)q * D + j) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(planes + (int64_t)(lo_rows + q) * D + j) = *reinterpret_cast<const uint2*>(l);
  }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) xsq_part[(int64_t)blockIdx.x * kUQ + q] = ss;
}

// ------------------------------------------------------------------------------------------ phase A
// grid (row tiles, k splits).  S_T [ksplit][Npad][128] fp32: S_T[s][i][r] = sum over K split s of X[r][d] * bank[i][d].
//   G = 1: the 128 operand rows are 64 hi parts stacked on 64 lo parts; B = hi tile, then lo tile (2 MMAs per K step,
//          the lo x lo product comes for free) and the dot of query q is S_T[..][q] + S_T[..][64 + q];
//   G = 2: two 128-row operands, the hi parts and the lo parts of 128 queries; hi x hi, hi x lo and lo x hi go into
//          ONE accumulator (3 MMAs per K step for twice the queries -- the tensor pipe, not HBM, would bound four)
//          and S_T[..][r] is the complete dot of query r.
//
// The tensor core adds into its fp32 accumulator with truncation, and near a negative the distance is the
// small difference of large dot products, so long accumulation chains cost accuracy (measured: 43 K-blocks in
// one chain -> 3.6e-4 on the weights).  The MMA warp therefore alternates between two TMEM accumulators every
// kUChunk K-blocks and the epilogue warps drain the finished one into fp32 registers (round-to-nearest adds).
constexpr int kUChunk = 4;

constexpr int kUThreadsA = 192;   // TMA warp, MMA warp, 4 epilogue warps
constexpr uint32_t kAccColsA = 128;

// Weights step fused into phase A (round 2): the CTA that delivers the LAST K-split partial of a row tile turns the
// tile's dots into weights -- sums the K-split partials in split order, k = exp(-dist / 2 sigma^2), writes the weight
// planes P [Npad][128] bf16 (hi parts in columns [0,64), lo parts in [64,128): phase B's MN-major B operand) and the
// tile's z sums / maxima; the CTA that finishes the last tile sums those over the tiles in tile order (z, kmax, the
// flat-regime mark).  Replaces the k_umma_weights + k_umma_zreduce launches (11.6 + 4 us at cfg3, 13.7 us for a
// 375-row shard) by a ~3 us tail on 24 CTAs in parallel.  Arrival counters are zeroed by the query-prepare kernel.
struct DotsTail {
  int enabled;
  int* tile_count;          // [row tiles] K splits that have delivered, then [1] tiles whose weights are done
  const float* sqnorm; const float* xsq; const float* xsq_part; int xsq_nparts;
  int N, Q, row_tiles;
  float inv2s2, alpha; int power;
  __nv_bfloat16* P; int64_t p_group_stride;      // [G][Npad][128]
  int interleave_k;         // K splits take every ksplit-th K block instead of a contiguous range
  int merged;               // tm_bank is the 3-D map [D][N][2 planes]: hi + lo tile in one TMA request
  int dbg_noshared;         // experiment (WRONG results): the X tiles are loaded for the first stages only
  int dbg_nomma;            // experiment (WRONG results): stages are released without any MMA
  int dbg_nostore;          // experiment (WRONG results): the partial dots are not stored
  int keep_from_row;        // bank rows >= this are loaded with L2 evict_last, the others evict_first (-1: no hints):
                            // phase B starts with the rows phase A read last and finds them in the L2
  float* zpart; int64_t zpart_stride;            // [G][row tiles][sums 64 | maxima 64]
  float* z; float* kmax; int* dense_flag;        // [64 G], [64 G], [1] or null
  float* k_out;                                  // [Q][N] or null
};

template <int G>
__device__ __forceinline__ void dots_tail(const DotsTail& T, const float* __restrict__ S_T, int64_t split_stride, int ksplit,
                                          int row0, float (*zs)[kUQ], float (*zm)[kUQ], int* s_flag, uint8_t* stage,
                                          uint64_t* bar, float* sq_s) {
  const int t = threadIdx.x, q = t & (kUQ - 1), rsub = t >> 6;       // 192 threads: 3 row subsets x 64 query rows
  constexpr int kSub = kUThreadsA / kUQ;
  // The K-split partials of the tile come into the (now idle) pipeline stages by bulk copies -- a round of R rows is
  // ksplit contiguous pieces of R x 512 bytes: register loads from L2 were latency-bound (54 us for this step at cfg3).
  const int R = min(kUBankTile, (int)(kPipeBytes / ((uint32_t)ksplit * kUStack * 4)));
  const float* sst = reinterpret_cast<const float*>(stage);          // [ksplit][R][128]
  float xs[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int Qg = min(kUQ, T.Q - g * kUQ);
    xs[g] = 0.f;
    if (q < Qg) {
      if (T.xsq) {
        xs[g] = T.xsq[g * kUQ + q];
      } else {
        // batches of independent loads: a loop of dependent L2 loads costs a round trip per iteration
        const float* xp = T.xsq_part + (int64_t)g * T.xsq_nparts * kUQ;
        for (int c0 = 0; c0 < T.xsq_nparts; c0 += 16) {
          float v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = c0 + u < T.xsq_nparts ? __ldcg(xp + (int64_t)(c0 + u) * kUQ + q) : 0.f;
#pragma unroll
          for (int u = 0; u < 16; ++u) xs[g] += v[u];
        }
      }
    }
  }
  // ||n||^2 of the tile's rows: one parallel load (a load per row inside the loop serialises on L2 latency)
  if (t < kUBankTile) sq_s[t] = row0 + t < T.N ? __ldg(T.sqnorm + row0 + t) : 0.f;
  __syncthreads();
  float zsum[G], zmax[G];
#pragma unroll
  for (int g = 0; g < G; ++g) zsum[g] = zmax[g] = 0.f;
  int round = 0;
#pragma unroll 1
  for (int r0 = 0; r0 < kUBankTile; r0 += R, ++round) {
    const int nr = min(R, kUBankTile - r0);
    if (t == 0) {
      const uint32_t bytes = (uint32_t)nr * kUStack * 4;
      u_mbar_expect_tx(bar, bytes * (uint32_t)ksplit);
      for (int sp = 0; sp < ksplit; ++sp)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(u_smem(stage + (size_t)sp * R * kUStack * 4)),
                       "l"(S_T + (int64_t)sp * split_stride + (int64_t)(row0 + r0) * kUStack), "r"(bytes), "r"(u_smem(bar))
                     : "memory");
    }
    u_mbar_wait(bar, (uint32_t)(round & 1));
#pragma unroll 1
    for (int i = rsub; i < nr; i += kSub) {
      const int row = row0 + r0 + i;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int Qg = min(kUQ, T.Q - g * kUQ);
        const float* sp0 = sst + (size_t)i * kUStack + (G == 1 ? q : g * kUQ + q);
        float dot = 0.f;
#pragma unroll 4
        for (int sp = 0; sp < ksplit; ++sp) {
          const float* p = sp0 + (size_t)sp * R * kUStack;
          dot += G == 1 ? p[0] + p[kUQ] : p[0];
        }
        float k = 0.f;
        if (row < T.N && q < Qg) k = weight_from_dot(xs[g], sq_s[r0 + i], dot, T.alpha, T.power, T.inv2s2);
        __nv_bfloat16* Pg = T.P + (int64_t)g * T.p_group_stride;
        const __nv_bfloat16 h = __float2bfloat16_rn(k);
        Pg[(int64_t)row * kUStack + q] = h;
        Pg[(int64_t)row * kUStack + kUQ + q] = __float2bfloat16_rn(k - __bfloat162float(h));
        if (T.k_out && row < T.N && q < Qg) T.k_out[(int64_t)(g * kUQ + q) * T.N + row] = k;
        zsum[g] += k;
        zmax[g] = fmaxf(zmax[g], k);
      }
    }
    __syncthreads();                     // the stage is read out before the next round's copies land
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    zs[rsub][q] = zsum[g]; zm[rsub][q] = zmax[g];
    __syncthreads();
    if (t < kUQ) {
      float a = 0.f, m = 0.f;
#pragma unroll
      for (int r = 0; r < kSub; ++r) { a += zs[r][t]; m = fmaxf(m, zm[r][t]); }
      float* zp = T.zpart + (int64_t)g * T.zpart_stride + (int64_t)blockIdx.x * kUStack;
      zp[t] = a; zp[kUQ + t] = m;
    }
    __syncthreads();
  }
  // ---- the CTA that completes the last row tile sums the tiles' partials in tile order
  __threadfence();
  __syncthreads();
  if (t == 0) *s_flag = atomicAdd(T.tile_count + T.row_tiles, 1) == T.row_tiles - 1 ? 1 : 0;
  __syncthreads();
  if (!*s_flag) return;
  __threadfence();
#pragma unroll 1
  for (int g = 0; g < G; ++g) {
    const float* zp = T.zpart + (int64_t)g * T.zpart_stride;
    float a = 0.f, m = 0.f;
    for (int b0 = rsub; b0 < T.row_tiles; b0 += 8 * kSub) {
      float va[8], vm[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = b0 + u * kSub;
        va[u] = b < T.row_tiles ? __ldcg(zp + (int64_t)b * kUStack + q) : 0.f;
        vm[u] = b < T.row_tiles ? __ldcg(zp + (int64_t)b * kUStack + kUQ + q) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { a += va[u]; m = fmaxf(m, vm[u]); }
    }
    zs[rsub][q] = a; zm[rsub][q] = m;
    __syncthreads();
    if (t < kUQ && t < T.Q - g * kUQ) {
      float zz = 0.f, mm = 0.f;
#pragma unroll
      for (int r = 0; r < kSub; ++r) { zz += zs[r][t]; mm = fmaxf(mm, zm[r][t]); }
      T.z[g * kUQ + t] = zz;
      T.kmax[g * kUQ + t] = mm;
      // sum_i k_i / kmax <= (#rows with k_i >= tau kmax) + N tau: more than kListCap "effective rows" means the
      // query's significant-row list must overflow
      if (T.dense_flag && zz > (float)kListCap * mm) atomicOr(T.dense_flag, 1);
    }
    __syncthreads();
  }
}

template <int G>
__global__ void __launch_bounds__(kUThreadsA, 1)
k_umma_dots(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hi,
            const __grid_constant__ CUtensorMap tm_lo, const __grid_constant__ CUtensorMap tm_bank,
            float* __restrict__ S_T, int64_t split_stride,
            int kblocks_total, int ksplit, int use_lo, const DotsTail tail) {
  using C = UCfg<G>;
  extern __shared__ unsigned char smem_raw[];
  __shared__ float zs[kUThreadsA / kUQ][kUQ], zm[kUThreadsA / kUQ][kUQ];
  __shared__ int s_flag;
  __shared__ uint64_t tail_bar;
  __shared__ float sq_s[kUBankTile];
  const USmem sm = u_carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kUBankTile;
  // K blocks of this split: a contiguous range, or (tail.interleave_k) every ksplit-th block -- then the CTAs of a row
  // tile read ADJACENT 128-byte pieces of the same bank rows at the same time (DRAM page locality)
  const int kstride = tail.interleave_k ? ksplit : 1;
  const int kb0 = tail.interleave_k ? (int)blockIdx.y : (int)((int64_t)blockIdx.y * kblocks_total / ksplit);
  const int kb1 = (int)((int64_t)(blockIdx.y + 1) * kblocks_total / ksplit);
  const int nkb = tail.interleave_k ? (kblocks_total - (int)blockIdx.y + ksplit - 1) / ksplit
                                    : kb1 - (int)((int64_t)blockIdx.y * kblocks_total / ksplit);
  const int nchunks = (nkb + kUChunk - 1) / kUChunk;

  uint64_t bank_policy = 0;
  if (tail.keep_from_row >= 0) {
    if (row0 >= tail.keep_from_row) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(bank_policy));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(bank_policy));
  }
  // what: 1 = arm the barrier + the bank tiles, 2 = the X tiles, 3 = both
  auto load_stage = [&](int i, int what = 3) {
    const int s = i % C::kStages;
    uint8_t* st = sm.tiles + (size_t)s * C::kStageBytes;
    const int kc = (kb0 + i * kstride) * kUK;
    const bool skip_x = tail.dbg_noshared && i >= C::kStages;
    if ((what & 2) && !skip_x) {
#pragma unroll
      for (int g = 0; g < G; ++g) u_tma_2d(st + (size_t)g * kTileBytes, &tm_x, kc, g * kUStack, &sm.full[s]);
    }
    if (!(what & 1)) return;
    u_mbar_expect_tx(&sm.full[s], (uint32_t)((skip_x ? 0 : G) + 1 + (use_lo ? 1 : 0)) * kTileBytes);
    if (tail.merged && use_lo && tail.keep_from_row < 0) {
      u_tma_3d(st + C::kHiOff, &tm_bank, kc, row0, 0, &sm.full[s]);      // hi tile | lo tile: one 32 KiB request
    } else if (tail.keep_from_row >= 0) {
      u_tma_2d_hint(st + C::kHiOff, &tm_hi, kc, row0, &sm.full[s], bank_policy);
      if (use_lo) u_tma_2d_hint(st + C::kLoOff, &tm_lo, kc, row0, &sm.full[s], bank_policy);
    } else {
      u_tma_2d(st + C::kHiOff, &tm_hi, kc, row0, &sm.full[s]);
      if (use_lo) u_tma_2d(st + C::kLoOff, &tm_lo, kc, row0, &sm.full[s]);
    }
  };
  const int npre = min(nkb, C::kStages);
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) { u_mbar_init(&sm.full[s], 1); u_mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { u_mbar_init(&sm.acc_full[b], 1); u_mbar_init(&sm.acc_empty[b], 4); }
    u_mbar_init(&tail_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the first stages need nothing but this thread's own barriers: their HBM latency overlaps the TMEM
    // allocation and the start-up barrier (the kernel is fill/drain bound at N ~ 3000); the bank tiles do not even
    // need the query-prepare kernel to have finished, only the X tiles do
    u_prefetch_map(&tm_x); u_prefetch_map(&tm_hi); u_prefetch_map(&tm_lo); u_prefetch_map(&tm_bank);
    for (int i = 0; i < npre; ++i) load_stage(i, 1);
    pdl_wait();
    for (int i = 0; i < npre; ++i) load_stage(i, 2);
  }
  pdl_wait();
  if (warp == 1) u_tmem_alloc(sm.tmem_base, 2 * kAccColsA);
  u_fence_before();
  __syncthreads();
  u_fence_after();
  const uint32_t tmem = *sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = npre; i < nkb; ++i) {
        u_mbar_wait(&sm.empty[i % C::kStages], (uint32_t)(((i / C::kStages) + 1) & 1));
        load_stage(i);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = u_idesc(kUStack, kUBankTile, 0, 0);
      for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c >= 2) {
          u_mbar_wait(&sm.acc_empty[buf], (uint32_t)(((c >> 1) + 1) & 1));
          u_fence_after();
        }
        const uint32_t acc = tmem + (uint32_t)buf * kAccColsA;
        const int i1 = min(nkb, (c + 1) * kUChunk);
        for (int i = c * kUChunk; i < i1; ++i) {
          const int s = i % C::kStages;
          u_mbar_wait(&sm.full[s], (uint32_t)((i / C::kStages) & 1));
          u_fence_after();
          const uint32_t base = u_smem(sm.tiles + (size_t)s * C::kStageBytes);
#pragma unroll
          for (int kk = 0; kk < (tail.dbg_nomma ? (i == c * kUChunk ? 1 : 0) : kUK / 16); ++kk) {
            const uint64_t bh = u_desc(base + C::kHiOff + kk * 32, 16, 1024);
            const uint64_t bl = u_desc(base + C::kLoOff + kk * 32, 16, 1024);
            const uint64_t a0 = u_desc(base + kk * 32, 16, 1024);                 // G = 1: stacked hi|lo; G = 2: hi parts
            u_mma(acc, a0, bh, idesc, (i > c * kUChunk || kk > 0) ? 1u : 0u);
            if (use_lo) u_mma(acc, a0, bl, idesc, 1u);
            if constexpr (G == 2) {
              const uint64_t a1 = u_desc(base + kTileBytes + kk * 32, 16, 1024);  // lo parts of the 128 queries
              u_mma(acc, a1, bh, idesc, 1u);
            }
          }
          u_commit(&sm.empty[s]);
        }
        u_commit(&sm.acc_full[buf]);
      }
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32); lane index = operand row (stacked / plain query row)
    const int lq = warp & 3;
    float sum[kUBankTile];
#pragma unroll
    for (int j = 0; j < kUBankTile; ++j) sum[j] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      u_mbar_wait(&sm.acc_full[buf], (uint32_t)((c >> 1) & 1));
      u_fence_after();
      const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)buf * kAccColsA;
#pragma unroll
      for (int cc = 0; cc < kUBankTile / 32; ++cc) {
        float v[32];
        u_tmem_ld32(tl + (uint32_t)(cc * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[cc * 32 + j] += v[j];
      }
      u_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(u_smem(&sm.acc_empty[buf])) : "memory");
    }
    // every K split owns its own partial buffer: plain coalesced stores, summed by k_umma_weights
    const int r = lq * 32 + lane;
    float* dst = S_T + (int64_t)blockIdx.y * split_stride + (int64_t)row0 * kUStack + r;
    if (tail.dbg_nostore) {                  // experiment (WRONG results): one store instead of 128
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < kUBankTile; ++j) v += sum[j];
      dst[0] = v;
    } else {
#pragma unroll
      for (int j = 0; j < kUBankTile; ++j) dst[(int64_t)j * kUStack] = sum[j];
    }
  }
  u_fence_before();
  __syncthreads();
  if (warp == 1) {
    u_fence_after();
    u_tmem_dealloc(tmem, 2 * kAccColsA);
  }
  if (tail.enabled) {
    // the partial of this K split is in global memory (epilogue stores above): is this CTA the tile's last arrival?
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_flag = atomicAdd(tail.tile_count + blockIdx.x, 1) == ksplit - 1 ? 1 : 0;
    __syncthreads();
    if (s_flag) {
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");    // the partials were written by generic stores of other SMs
      dots_tail<G>(tail, S_T, split_stride, ksplit, row0, zs, zm, &s_flag, sm.tiles, &tail_bar, sq_s);
    }
  }
}

// ------------------------------------------------------------------------------------------ weights
// Thread = (bank row, query row): sum the K-split partials in split order, k = exp(-dist / 2 sigma^2) (or the SPELL
// weight, weight_from_dot), write the weight planes P [Npad][128] bf16 (stacked query index contiguous: hi parts in
// columns [0,64), lo parts in [64,128)), which phase B reads as an MN-major B operand.  Everything is coalesced and
// every load of a row is issued before the first use.  A block has 4 or 16 row lanes (256 / 1024 threads) and walks
// `rows_per_block` bank rows; per block one z partial (sums | maxima), summed in block order by
//   - phase B's idle epilogue warps (dense accumulate, AccumEpi::zpart: 1024-thread blocks of 16-32 rows, few partials),
//   - k_umma_zreduce (block-sparse mode, which needs kmax before phase B; z-only passes of empirical_beta), or
//   - the last block of this kernel (`arrivals` != null; measured slower with few fat blocks, not used).
constexpr int kWRows = 4;          // bank rows a block handles per step (and per k_umma_siglist block)

// The dot of query row q is S_T[..][col0 + q] (+ S_T[..][col0 + lo_off + q] when lo_off > 0: stacked operand).
constexpr int kWMaxLanes = 16;     // row lanes of the fat-block variant (1024 threads)
__global__ void __launch_bounds__(kWMaxLanes * kUQ)
k_umma_weights(const float* __restrict__ S_T, int64_t split_stride, int ksplit, int col0, int lo_off,
               const float* __restrict__ sqnorm,
               const float* __restrict__ xsq, const float* __restrict__ xsq_part, int xsq_nparts, int N, int Q,
               float inv2s2, int power, float alpha, __nv_bfloat16* __restrict__ P, float* __restrict__ zpart,
               float* __restrict__ k_out, int rows_per_block, int npad, int* __restrict__ arrivals, float* __restrict__ z,
               float* __restrict__ kmax, int* __restrict__ dense_flag) {
  __shared__ float zs[kWMaxLanes][kUQ], zm[kWMaxLanes][kUQ];
  __shared__ int s_last;
  pdl_launch_dependents();
  pdl_wait();
  const int q = threadIdx.x & (kUQ - 1);
  const int rsub = threadIdx.x >> 6;
  const int lanes = (int)blockDim.x >> 6;           // bank rows the block handles per step
  float zsum = 0.f, zmax = 0.f;
  const int i0 = blockIdx.x * rows_per_block;
  float xs = 0.f;
  bool have_xs = false;
#pragma unroll 1
  for (int r = rsub; r < rows_per_block && i0 + r < npad; r += lanes) {
    const int i = i0 + r;
    // every load of the row is issued before the first use: the kernel is one L2 round trip deep, not three
    const float* s = S_T + (int64_t)i * kUStack + col0 + q;
    float a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float* p = s + (int64_t)u * split_stride;
      a[u] = u < ksplit ? __ldg(p) : 0.f;
      b[u] = (u < ksplit && lo_off > 0) ? __ldg(p + lo_off) : 0.f;
    }
    const float sq = i < N ? __ldg(sqnorm + i) : 0.f;
    if (!have_xs) {
      have_xs = true;
      if (q < Q) {
        if (xsq) {
          xs = xsq[q];
        } else {
          // batches of independent loads: `xs += load` in a plain loop is one L2 round trip per part (8 us of this
          // kernel at D = 16384 in round 1)
          for (int c0 = 0; c0 < xsq_nparts; c0 += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = c0 + u < xsq_nparts ? __ldg(xsq_part + (int64_t)(c0 + u) * kUQ + q) : 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u) xs += v[u];
          }
        }
      }
    }
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (u < ksplit) dot += lo_off > 0 ? a[u] + b[u] : a[u];
#pragma unroll 1
    for (int sp0 = 8; sp0 < ksplit; sp0 += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float* p = s + (int64_t)(sp0 + u) * split_stride;
        a[u] = sp0 + u < ksplit ? __ldg(p) : 0.f;
        b[u] = (sp0 + u < ksplit && lo_off > 0) ? __ldg(p + lo_off) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (sp0 + u < ksplit) dot += lo_off > 0 ? a[u] + b[u] : a[u];
    }
    float k = 0.f;
    if (i < N && q < Q) k = weight_from_dot(xs, sq, dot, alpha, power, inv2s2);
    const __nv_bfloat16 h = __float2bfloat16_rn(k);
    P[(int64_t)i * kUStack + q] = h;
    P[(int64_t)i * kUStack + kUQ + q] = __float2bfloat16_rn(k - __bfloat162float(h));
    if (k_out && i < N && q < Q) k_out[(int64_t)q * N + i] = k;
    zsum += k;
    zmax = fmaxf(zmax, k);
  }
  zs[rsub][q] = zsum; zm[rsub][q] = zmax;
  __syncthreads();
  if (threadIdx.x < kUQ) {
    float t = 0.f, m = 0.f;
    for (int r = 0; r < lanes; ++r) {
      t += zs[r][threadIdx.x];
      m = fmaxf(m, zm[r][threadIdx.x]);
    }
    zpart[(int64_t)blockIdx.x * kUStack + threadIdx.x] = t;          // [block][0..63]  sums
    zpart[(int64_t)blockIdx.x * kUStack + kUQ + threadIdx.x] = m;    // [block][64..127] maxima
  }
  if (!arrivals) return;               // k_umma_zreduce (or phase B's epilogue warps) sums the partials
  // ---- the last block sums the partials in block order (deterministic)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(arrivals, 1) == (int)gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float a = 0.f, m = 0.f;
  for (int b0 = rsub; b0 < (int)gridDim.x; b0 += 8 * lanes) {
    float va[8], vm[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int b = b0 + u * lanes;
      va[u] = b < (int)gridDim.x ? __ldcg(zpart + (int64_t)b * kUStack + q) : 0.f;
      vm[u] = b < (int)gridDim.x ? __ldcg(zpart + (int64_t)b * kUStack + kUQ + q) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { a += va[u]; m = fmaxf(m, vm[u]); }
  }
  __syncthreads();
  zs[rsub][q] = a; zm[rsub][q] = m;
  __syncthreads();
  if (threadIdx.x < kUQ && threadIdx.x < Q) {
    float t = 0.f, mm = 0.f;
    for (int r = 0; r < lanes; ++r) { t += zs[r][threadIdx.x]; mm = fmaxf(mm, zm[r][threadIdx.x]); }
    z[threadIdx.x] = t;
    kmax[threadIdx.x] = mm;
    // sum_i k_i / kmax <= (#rows with k_i >= tau kmax) + N tau: more than kListCap "effective rows" means the
    // query's significant-row list must overflow
    if (dense_flag && t > (float)kListCap * mm) atomicOr(dense_flag, 1);
  }
}

// z[q] = sum_b zpart[b][q], kmax[q] = max_b zpart[b][64+q]; one block per query row.
__global__ void __launch_bounds__(256)
k_umma_zreduce(const float* __restrict__ zpart, int nblocks, float* __restrict__ z, float* __restrict__ kmax,
               int* __restrict__ dense_flag) {
  __shared__ float red[33];
  __shared__ float mx[8];
  pdl_launch_dependents();
  pdl_wait();
  const int q = blockIdx.x;
  float t = 0.f, m = 0.f;
  for (int b = threadIdx.x; b < nblocks; b += 256) {
    t += zpart[(int64_t)b * kUStack + q];
    m = fmaxf(m, zpart[(int64_t)b * kUStack + kUQ + q]);
  }
  t = block_sum(t, red);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) mx[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    z[q] = t;
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, mx[w]);
    kmax[q] = m;
    // sum_i k_i / kmax <= (#rows with k_i >= tau kmax) + N tau: more than kListCap "effective rows" means the
    // query's significant-row list must overflow
    if (dense_flag && t > (float)kListCap * m) atomicOr(dense_flag, 1);
  }
}

// Block-sparse phase B.  With a small sigma the weights of a query span many orders of magnitude; a block of 64
// bank rows in which EVERY weight is below kSkipRel times its query's largest weight contributes less than
// N * kSkipRel (3e-6 at N = 3000) relative to the dominant term -- below the fp32 summation noise -- and phase B
// does not read it.  flags[rb] = 1 when row block rb holds at least one weight above that bound.
constexpr float kSkipRel = 1e-9f;

// One thread per (bank row, query row): mark row blocks and append (row, k) to the query's list when the weight is
// significant.  grid = Npad / 4, block = 256.
__global__ void __launch_bounds__(kWRows * kUQ)
k_umma_siglist(const __nv_bfloat16* __restrict__ P, const float* __restrict__ kmax, int N, int Q, SigLists L) {
  __shared__ int sg[kWRows][kUQ];
  if (__ldg(L.dense)) return;                          // flat regime: dense phase B, nothing to build
  const int q = threadIdx.x & (kUQ - 1), rsub = threadIdx.x >> 6;
  const int i = blockIdx.x * kWRows + rsub;
  float k = 0.f;
  bool sig = false;
  if (i < N && q < Q) {
    k = __bfloat162float(P[(int64_t)i * kUStack + q]) + __bfloat162float(P[(int64_t)i * kUStack + kUQ + q]);
    sig = k > 0.f && k >= kSkipRel * kmax[q];
  }
  sg[rsub][q] = sig ? 1 : 0;
  const int any = __syncthreads_or(sig ? 1 : 0);
  if (any && threadIdx.x == 0) L.flags[(blockIdx.x * kWRows) / kUK] = 1;   // kWRows divides 64: one block, one flag
  if (!sig) return;
  // A query whose weights are significant on every row of this block is in a flat regime: its list would
  // overflow anyway, so mark it overflowed with a plain store instead of hammering its counter; later threads
  // see the mark and stop too.  Peaked regimes (the case the lists exist for) add a handful of entries.
  if (sg[0][q] + sg[1][q] + sg[2][q] + sg[3][q] == kWRows) {
    if (rsub == 0 && *reinterpret_cast<volatile int*>(L.count + q) <= kListCap)
      *reinterpret_cast<volatile int*>(L.count + q) = kListCap + (1 << 20);
    return;
  }
  if (*reinterpret_cast<volatile int*>(L.count + q) <= kListCap) {
    const int pos = atomicAdd(L.count + q, 1);
    if (pos < kListCap) {
      L.rows[q * kListCap + pos] = i;
      L.ks[q * kListCap + pos] = k;
    }
  }
}

// Block-cooperative (every thread of the block must call it): true when every query has at most kListCap
// significant rows.  One parallel load per query row -- a serial loop of 64 dependent L2 loads costs ~25 us.
__device__ __forceinline__ bool siglist_all_short(const int* count, int Q) {
  bool ok = true;
  for (int q = threadIdx.x; q < Q; q += blockDim.x) ok = ok && (__ldcg(count + q) <= kListCap);
  return __syncthreads_and(ok ? 1 : 0) != 0;
}

// Listed accumulate: when every query has at most kListCap significant bank rows (peaked weights: small sigma),
// num[q] = sum over that query's list of k * (hi + lo)[row] on the CUDA cores -- a few MB instead of a pass over
// the bank.  grid (D / 1024, Q), block 256, thread = 4 consecutive d.  Exits at once when a list overflowed
// (k_umma_accum then does the block-sparse tensor-core pass).
struct AccumEpi;
__global__ void __launch_bounds__(256)
k_umma_listed_accum(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int64_t D, int Q,
                    SigLists L, float* __restrict__ num, const float* z, float eps, float scale, float gate_thr,
                    int flags, float* x0, float* neg_out, float* denom_out, int32_t* gate_out, float* mean_out,
                    float inv_qd) {
  __shared__ int srow[kListCap];
  __shared__ float sk[kListCap];
  __shared__ float red[33];
  if (__ldg(L.dense) || !siglist_all_short(L.count, Q)) return;
  const int q = blockIdx.y;
  const int n = L.count[q];
  // sort the list by row index so that the fp32 summation order does not depend on atomic arrival order
  if (threadIdx.x < n) {
    const int r = L.rows[q * kListCap + threadIdx.x];
    int rank = 0;
    for (int e = 0; e < n; ++e) rank += (L.rows[q * kListCap + e] < r) ? 1 : 0;
    srow[rank] = r;
    sk[rank] = L.ks[q * kListCap + threadIdx.x];
  }
  __syncthreads();
  const int64_t d = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  float msum = 0.f;
  if (d < D) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < n; ++e) {
      const int64_t o = (int64_t)srow[e] * D + d;
      const uint2 h = __ldg(reinterpret_cast<const uint2*>(hi + o));
      const uint2 l = lo ? __ldg(reinterpret_cast<const uint2*>(lo + o)) : make_uint2(0u, 0u);   // bf16 zero bits
      const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h);
      const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&l);
      const float k = sk[e];
      acc.x = fmaf(k, __bfloat162float(hb[0]) + __bfloat162float(lb[0]), acc.x);
      acc.y = fmaf(k, __bfloat162float(hb[1]) + __bfloat162float(lb[1]), acc.y);
      acc.z = fmaf(k, __bfloat162float(hb[2]) + __bfloat162float(lb[2]), acc.z);
      acc.w = fmaf(k, __bfloat162float(hb[3]) + __bfloat162float(lb[3]), acc.w);
    }
    const int64_t o = (int64_t)q * D + d;
    if (num) *reinterpret_cast<float4*>(num + o) = acc;
    if (z) {
      const float denom = z[q] + eps;
      const float4 nn = make_float4(acc.x / denom, acc.y / denom, acc.z / denom, acc.w / denom);
      if (neg_out) *reinterpret_cast<float4*>(neg_out + o) = nn;
      if (x0) {
        float4 x = *reinterpret_cast<const float4*>(x0 + o);
        x.x = fmaf(-scale, nn.x, x.x); x.y = fmaf(-scale, nn.y, x.y);
        x.z = fmaf(-scale, nn.z, x.z); x.w = fmaf(-scale, nn.w, x.w);
        *reinterpret_cast<float4*>(x0 + o) = x;
      }
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (denom_out) denom_out[q] = denom;
        if (gate_out) gate_out[q] = (!(flags & SDN_EPI_GATE) || denom > gate_thr) ? 1 : 0;
      }
      msum = fminf(fmaxf(nn.x, -1e10f), 1e10f) + fminf(fmaxf(nn.y, -1e10f), 1e10f) +
             fminf(fmaxf(nn.z, -1e10f), 1e10f) + fminf(fmaxf(nn.w, -1e10f), 1e10f);
    }
  }
  if (z && mean_out) {
    msum = block_sum(msum, red);
    if (threadIdx.x == 0) atomicAdd(mean_out, msum * inv_qd);
  }
}

// ------------------------------------------------------------------------------------------ phase B
constexpr int kBChunk = 64;        // row blocks per TMEM accumulation chain in phase B (512 MMAs)
constexpr int kMaxActive = 4096;   // row blocks a CTA can index in its active list (N <= 262144 per split)

// Optional correction fused into phase B's epilogue (one GPU, no bank-row split): x0 -= scale * num / (z + eps).
// Pointers address query row 0 of the pass; group g of the pass uses rows [64 g, 64 g + 64).
// experiment: per-CTA timestamps of phase B (AccumEpi::dbg & 64), read by umma_accum_trace_read
__device__ unsigned long long g_accum_trace[256 * 8];
__device__ __forceinline__ void accum_trace(int on, int ev) {
  if (on && blockIdx.y == 0 && blockIdx.x < 256) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_accum_trace[blockIdx.x * 8 + ev] = t;
  }
}

__device__ __forceinline__ void accum_trace_dep(int on, int ev, float dep) {
  if (on && blockIdx.y == 0 && blockIdx.x < 256) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) : "f"(dep) : "memory");
    g_accum_trace[blockIdx.x * 8 + ev] = t;
  }
}

struct AccumEpi {
  const float* z;       // null: no fused correction
  float eps, scale, gate_thr; int flags;
  float* x0; float* neg_out; float* denom_out; int32_t* gate_out; float* mean_out; float inv_qd;
  int reverse;          // walk the row blocks from the last to the first (the rows phase A read last are still in L2)
  // z from the weights kernel's per-block partials [nzpart][128] (group g at + g * zpart_stride), summed in block order
  // by this kernel's epilogue warps while the main loop runs -- instead of a k_umma_zreduce launch (6 us at cfg3).  The
  // CTAs of d-block 0 also write the sums to `z` (which then is an output).  Null: read z.
  const float* zpart; int nzpart; int64_t zpart_stride; float* z_out;
  int merged;           // tm_bank4 / tm_p3 are the merged-request maps
  int dbg_noshared;     // experiment (WRONG results): the weight tiles are loaded for the first stages only
  int dbg_nomma;        // experiment (WRONG results): stages are released without any MMA
  // Non-null: the sums go to a tile-contiguous scratch [D/128][64 G][128] and k_umma_untile applies the fused correction
  // (and writes the optional outputs) with row-contiguous accesses; null: this kernel updates x0 itself.
  float* tile_out;
  int dbg;              // experiment bits (WRONG results): 1 no epilogue loads/stores, 2 no z partial sums, 4 no TMEM loads
  int z_only;           // zpart given but no correction here (N-sharded banks: the merge kernel applies it): write z_out only
};

// grid (D / 128, n splits).  num[q][d] (+)= sum_i P[q][i] * (hi+lo)[i][d] over this split's bank rows, for the
// Q <= 64 G query rows of the pass; the weight planes of group g are rows [g * p_group_rows, +Npad) of tm_p.
template <bool CHUNKED, int G>
__global__ void __launch_bounds__(UCfg<G>::kThreads, 1)
k_umma_accum(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_hi,
             const __grid_constant__ CUtensorMap tm_lo, const __grid_constant__ CUtensorMap tm_bank4,
             const __grid_constant__ CUtensorMap tm_p3, float* __restrict__ num, int64_t D, int Q,
             int rblocks_total, int nsplit, int64_t split_stride, int use_lo, int p_group_rows, const int* rowflags,
             const int* __restrict__ list_count, const int* __restrict__ dense_flag, const AccumEpi epi) {
  using C = UCfg<G>;
  pdl_launch_dependents();
  const int tr = epi.dbg & 64;
  if (threadIdx.x == 0) accum_trace(tr, 0);
  if (dense_flag || list_count || rowflags) pdl_wait();          // the lists come from the kernels before this one
  if (dense_flag && __ldg(dense_flag)) {
    rowflags = nullptr;                                          // flat regime: no flags were built
  } else if (list_count && siglist_all_short(list_count, Q)) {
    return;                                                      // k_umma_listed_accum produced the result
  }
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint16_t act[kMaxActive];
  __shared__ int nact_s;
  __shared__ float z_half[G][2][kUQ], z_sum[G][kUQ], z_rinv[G][kUQ];
  const USmem sm = u_carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb0 = (int)((int64_t)blockIdx.y * rblocks_total / nsplit);
  const int rb1 = (int)((int64_t)(blockIdx.y + 1) * rblocks_total / nsplit);
  const int nrb = rb1 - rb0;
  // persistent over d-blocks: this CTA owns blocks blockIdx.x, blockIdx.x + gridDim.x, ... and alternates
  // between two TMEM accumulators so that the epilogue of one block overlaps the stream of the next
  const int dblocks = (int)(D / kUDBlock);
  const int ntasks = (dblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  // stage `it` <- row block i of this split, d-block task t
  // what: 1 = arm the barrier + the bank tiles, 2 = the weight tiles (made by the kernels before this one), 3 = both
  auto load_stage = [&](int it, int i, int t, int what = 3) {
    const int d0 = ((int)blockIdx.x + t * (int)gridDim.x) * kUDBlock;
    const int s = it % C::kStages;
    uint8_t* st = sm.tiles + (size_t)s * C::kStageBytes;
    const int rc = (rb0 + i) * kUK;                   // first bank row of this block
    // P tiles: per group two boxes of [64 rows][64 stacked q] (hi parts, lo parts), 8 KiB apart
    const bool skip_p = epi.dbg_noshared && it >= C::kStages;
    if ((what & 2) && !skip_p) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (epi.merged) {             // both halves of the weight tile: one 16 KiB request
          u_tma_3d(st + (size_t)g * kTileBytes, &tm_p3, 0, g * p_group_rows + rc, 0, &sm.full[s]);
        } else {
          u_tma_2d(st + (size_t)g * kTileBytes, &tm_p, 0, g * p_group_rows + rc, &sm.full[s]);
          u_tma_2d(st + (size_t)g * kTileBytes + 8192, &tm_p, 64, g * p_group_rows + rc, &sm.full[s]);
        }
      }
    }
    if (!(what & 1)) return;
    u_mbar_expect_tx(&sm.full[s], (uint32_t)((skip_p ? 0 : G) + 1 + (use_lo ? 1 : 0)) * kTileBytes);
    // bank^T tiles: two boxes of [64 rows][64 d] per plane, 8 KiB apart -- one 4-D request of 32 KiB for all four
    if (epi.merged && use_lo) {
      u_tma_4d(st + C::kHiOff, &tm_bank4, 0, rc, d0 / 64, 0, &sm.full[s]);
      return;
    }
    u_tma_2d(st + C::kHiOff, &tm_hi, d0, rc, &sm.full[s]);
    u_tma_2d(st + C::kHiOff + 8192, &tm_hi, d0 + 64, rc, &sm.full[s]);
    if (use_lo) {
      u_tma_2d(st + C::kLoOff, &tm_lo, d0, rc, &sm.full[s]);
      u_tma_2d(st + C::kLoOff + 8192, &tm_lo, d0 + 64, rc, &sm.full[s]);
    }
  };
  // dense pass: the first stages depend on nothing but this thread's own barriers, so their latency overlaps
  // the TMEM allocation and the start-up barrier
  const int npre = (rowflags == nullptr && nrb > 0) ? min(ntasks * nrb, C::kStages) : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) { u_mbar_init(&sm.full[s], 1); u_mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { u_mbar_init(&sm.acc_full[b], 1); u_mbar_init(&sm.acc_empty[b], 4 * G); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    u_prefetch_map(&tm_p); u_prefetch_map(&tm_hi); u_prefetch_map(&tm_lo); u_prefetch_map(&tm_bank4); u_prefetch_map(&tm_p3);
    // the bank tiles of the first stages do not depend on the weights: they stream while the weights kernels finish
    for (int it = 0; it < npre; ++it) load_stage(it, epi.reverse ? nrb - 1 - it % nrb : it % nrb, it / nrb, 1);
    pdl_wait();
    for (int it = 0; it < npre; ++it) load_stage(it, epi.reverse ? nrb - 1 - it % nrb : it % nrb, it / nrb, 2);
  }
  pdl_wait();
  if (warp == 1) u_tmem_alloc(sm.tmem_base, 2 * C::kAccCols);
  if (warp == 0) {
    // compact the list of row blocks that hold a non-negligible weight (dense when no flags are given)
    int cnt = 0;
    if (rowflags) {
      for (int base = 0; base < nrb; base += 32) {
        const int i = base + lane;
        const int f = (i < nrb) ? rowflags[rb0 + i] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, f != 0);
        if (f) act[cnt + __popc(m & ((1u << lane) - 1))] = (uint16_t)i;
        cnt += __popc(m);
      }
    } else {
      cnt = nrb;
    }
    if (lane == 0) nact_s = cnt;
  }
  u_fence_before();
  __syncthreads();
  u_fence_after();
  if (threadIdx.x == 0) accum_trace(tr, 1);
  const uint32_t tmem = *sm.tmem_base;
  const int nact = nact_s;
  const bool dense = rowflags == nullptr;
  const int nchunks = max(1, (nact + kBChunk - 1) / kBChunk);

  if (warp == 0) {
    if (lane == 0) {
      for (int it = npre; it < ntasks * nact; ++it) {
        const int t = it / nact, j = it - t * nact;
        if (it >= C::kStages) u_mbar_wait(&sm.empty[it % C::kStages], (uint32_t)(((it / C::kStages) + 1) & 1));
        const int jj = epi.reverse ? nact - 1 - j : j;
        load_stage(it, dense ? jj : (int)act[jj], t);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_full = u_idesc(kUDBlock, kUStack, 1, 1);   // hi * [P_hi | P_lo]
      constexpr uint32_t id_half = u_idesc(kUDBlock, kUQ, 1, 1);       // lo * P_hi
      // one stage = one block of 64 bank rows: 4 K-steps x G groups x (hi, lo) MMAs
      auto mma_stage = [&](int it, uint32_t acc, bool first) {
        const int s = it % C::kStages;
        u_mbar_wait(&sm.full[s], (uint32_t)((it / C::kStages) & 1));
        u_fence_after();
        const uint32_t base = u_smem(sm.tiles + (size_t)s * C::kStageBytes);
#pragma unroll
        for (int kk = 0; kk < (epi.dbg_nomma ? (first ? 1 : 0) : kUK / 16); ++kk) {
          const uint64_t ah = u_desc(base + C::kHiOff + kk * 2048, 8192, 1024);        // bank^T, MN-major
          const uint64_t al = u_desc(base + C::kLoOff + kk * 2048, 8192, 1024);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const uint64_t b = u_desc(base + (uint32_t)g * kTileBytes + kk * 2048, 8192, 1024);   // P_g^T, MN-major
            u_mma(acc + (uint32_t)g * kUStack, ah, b, id_full, (!first || kk > 0) ? 1u : 0u);
            if (use_lo) u_mma(acc + (uint32_t)g * kUStack, al, b, id_half, 1u);
          }
        }
        u_commit(&sm.empty[s]);
      };
      if constexpr (!CHUNKED) {
        for (int t = 0; t < ntasks && nact > 0; ++t) {
          const int buf = t & 1;
          if (t >= 2) {
            u_mbar_wait(&sm.acc_empty[buf], (uint32_t)(((t >> 1) + 1) & 1));
            u_fence_after();
          }
          const uint32_t acc = tmem + (uint32_t)buf * C::kAccCols;
          for (int i = 0; i < nact; ++i) mma_stage(t * nact + i, acc, i == 0);
          u_commit(&sm.acc_full[buf]);
        }
      } else {
        // unit u = (d-block task, chunk of kBChunk row blocks); consecutive units alternate TMEM accumulators so
        // that no fp32 accumulation chain in the tensor core is longer than kBChunk * 8 MMAs (its adds truncate)
        for (int u = 0; nact > 0 && u < ntasks * nchunks; ++u) {
          const int t = u / nchunks, c = u - t * nchunks;
          const int buf = u & 1;
          if (u >= 2) {
            u_mbar_wait(&sm.acc_empty[buf], (uint32_t)(((u >> 1) + 1) & 1));
            u_fence_after();
          }
          const uint32_t acc = tmem + (uint32_t)buf * C::kAccCols;
          const int j0 = c * kBChunk, j1 = min(nact, j0 + kBChunk);
          for (int i = j0; i < j1; ++i) mma_stage(t * nact + i, acc, i == j0);
          u_commit(&sm.acc_full[buf]);
        }
      }
    }
  } else {
    // epilogue: warps 2..5 serve query group 0, warps 6..9 group 1
    const int lq = warp & 3;
    const int g = (warp - 2) >> 2;
    const int Qg = min(kUQ, Q - g * kUQ);                         // query rows of this group (>= 1)
    const int64_t qoff = (int64_t)g * kUQ * D;
    // bank-row splits write their own partial [Q][D] (summed in split order by k_umma_splitsum): no atomics
    float* const numg = num ? num + qoff + (int64_t)blockIdx.y * split_stride : nullptr;
    // z of this group: from global memory (made by k_umma_zreduce), or summed here from the weights kernel's partials
    const float* zg = epi.z ? epi.z + g * kUQ : nullptr;
    if (epi.zpart && (epi.dbg & 2)) {
      if (threadIdx.x < kUQ + 64) z_sum[g][threadIdx.x & (kUQ - 1)] = 1.f;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      zg = z_sum[g];
    } else if (epi.zpart) {
      const int t = (warp - 2 - 4 * g) * 32 + lane;          // 0..127 within the group's four warps
      const int qz = t & (kUQ - 1), half = t >> 6;
      const float* zp = epi.zpart + (int64_t)g * epi.zpart_stride + qz;
      float a = 0.f;
      for (int b0 = half; b0 < epi.nzpart; b0 += 16) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = b0 + 2 * u < epi.nzpart ? __ldcg(zp + (int64_t)(b0 + 2 * u) * kUStack) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) a += v[u];
      }
      z_half[g][half][qz] = a;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if (t < kUQ) {
        const float zz = z_half[g][0][t] + z_half[g][1][t];
        z_sum[g][t] = zz;
        if (blockIdx.x == 0 && blockIdx.y == 0 && epi.z_out && t < Qg) epi.z_out[g * kUQ + t] = zz;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      zg = epi.z_only ? nullptr : z_sum[g];
    }
    float* const x0g = epi.x0 ? epi.x0 + qoff : nullptr;
    float* const negg = epi.neg_out ? epi.neg_out + qoff : nullptr;
    float* const denomg = epi.denom_out ? epi.denom_out + g * kUQ : nullptr;
    int32_t* const gateg = epi.gate_out ? epi.gate_out + g * kUQ : nullptr;
    const uint32_t tcol = (uint32_t)(g * kUStack);
    float msum = 0.f;
    if constexpr (!CHUNKED) {
    // The epilogue runs on ONE warp per scheduler with nothing to overlap it once the bank stream has ended: its
    // instruction count is kernel time.  A per-element IEEE division, clamps and 64-bit index arithmetic (75 SASS
    // instructions per element) cost 12 us of the 49.5 us kernel at cfg3 (profiles/r02_experiments.txt); so: 1 / (z + eps)
    // once per query row, the optional outputs decided once per chunk, denominators and gates written once.
    if (zg) {
      const int tq = (warp - 2 - 4 * g) * 32 + lane;
      if (tq < kUQ) z_rinv[g][tq] = 1.f / (zg[min(tq, Qg - 1)] + epi.eps);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if (blockIdx.x == 0 && blockIdx.y == 0 && lq == 0) {
        for (int q = lane; q < Qg; q += 32) {
          const float dn = zg[q] + epi.eps;
          if (denomg) denomg[q] = dn;
          if (gateg) gateg[q] = (!(epi.flags & SDN_EPI_GATE) || dn > epi.gate_thr) ? 1 : 0;
        }
      }
    }
    const bool extras = numg != nullptr || negg != nullptr || epi.mean_out != nullptr;
    // TMEM lane = d within the block, column = stacked query row
#pragma unroll 1
    for (int t = 0; t < ntasks; ++t) {
    const int buf = t & 1;
    const int64_t d = (int64_t)((int)blockIdx.x + t * (int)gridDim.x) * kUDBlock + lq * 32 + lane;
    if (nact > 0) {
      u_mbar_wait(&sm.acc_full[buf], (uint32_t)((t >> 1) & 1));
      u_fence_after();
    }
    if (warp == 2 && lane == 0) accum_trace(tr, 4);
    const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)buf * C::kAccCols + tcol;
    const unsigned row_bytes = (unsigned)(D * 4);           // D < 2^30 (umma_supported): one IMAD.WIDE per address
#pragma unroll 1
    for (int c = 0; c < kUQ / 32; ++c) {
      float a[32];
      if (nact > 0 && !(epi.dbg & 4)) {
        float b[32];
        u_tmem_ld32(tl + (uint32_t)(c * 32), a);            // hi*P_hi + lo*P_hi, queries [32c, 32c+32)
        u_tmem_ld32(tl + (uint32_t)(kUQ + c * 32), b);      // hi*P_lo
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] += b[j];
      } else {                                              // no row block of this split matters: the sum is zero
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = 0.f;
      }
      const int qn = Qg - c * 32;                           // query rows of this chunk that exist
      const int64_t o0 = (int64_t)(c * 32) * D + d;
      if (tr && c == 0 && warp == 2 && lane == 0) accum_trace_dep(tr, 2, a[31]);      // TMEM loads of chunk 0 done
      if (epi.dbg & 1) {                                    // experiment (WRONG results): no loads, no stores
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) v += a[j];
        if (v == 1.2345f && x0g) x0g[d] = v;
      } else if (zg) {
        const float* ri = z_rinv[g] + c * 32;
        if (epi.tile_out) {
          // 128 bytes per warp and query row, rows 512 bytes apart: the pattern of phase A's partial stores.  The same
          // stores straight into x0 (rows D * 4 bytes apart) take 10 us per CTA (tools/gpu_accum_trace.py).
          const int dblock = (int)blockIdx.x + t * (int)gridDim.x;
          float* const tp = epi.tile_out + (((int64_t)dblock * G + g) * kUQ + c * 32) * kUDBlock + lq * 32 + lane;
#pragma unroll
          for (int j = 0; j < 32; ++j) tp[j * kUDBlock] = a[j];
        } else if (x0g) {
          // All loads of the chunk first: the stores below may alias them as far as the compiler knows.  (Fetching the
          // whole column before the wait for the accumulator makes the second chunk's stores take 20 us instead of 10:
          // tools/gpu_accum_trace.py, profiles/r02_experiments.txt.)
          char* const pb = reinterpret_cast<char*>(x0g + o0);
          float xv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            xv[j] = j < qn ? __ldcg(reinterpret_cast<const float*>(pb + (size_t)j * row_bytes)) : 0.f;
          if (tr && c == 0 && warp == 2 && lane == 0) accum_trace_dep(tr, 3, xv[0] + xv[31]);   // x0 loads of chunk 0 landed
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < qn) *reinterpret_cast<float*>(pb + (size_t)j * row_bytes) = fmaf(-epi.scale, a[j] * ri[j], xv[j]);
          if (tr && c == 0 && warp == 2 && lane == 0) accum_trace(tr, 7);                        // stores of chunk 0 issued
        }
        if (extras && !epi.tile_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < qn) {
              const float n = a[j] * ri[j];
              if (numg) numg[o0 + (int64_t)j * D] = a[j];
              if (negg) negg[o0 + (int64_t)j * D] = n;
              msum += fminf(fmaxf(n, -1e10f), 1e10f);
            }
          }
        }
      } else if (epi.tile_out) {                            // plain sums (partial-sums calls), through the tile scratch as well
        const int dblock = (int)blockIdx.x + t * (int)gridDim.x;
        float* const tp = epi.tile_out + (((int64_t)dblock * G + g) * kUQ + c * 32) * kUDBlock + lq * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) tp[j * kUDBlock] = a[j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < qn) numg[o0 + (int64_t)j * D] = a[j];
      }
    }
    if (warp == 2 && lane == 0) accum_trace(tr, 5);
    u_fence_before();
    __syncwarp();
    if (lane == 0 && nact > 0)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(u_smem(&sm.acc_empty[buf])) : "memory");
    }   // tasks
    } else {
    // running fp32 sums over the chunks
    float sum[kUQ];
#pragma unroll
    for (int j = 0; j < kUQ; ++j) sum[j] = 0.f;
    const int nunits = nact > 0 ? ntasks * nchunks : ntasks;     // nothing to read when no row block matters
#pragma unroll 1
    for (int u = 0; u < nunits; ++u) {
      const int t = nact > 0 ? u / nchunks : u;
      const int c = nact > 0 ? u - t * nchunks : 0;
      const int buf = u & 1;
      if (nact > 0) {
        u_mbar_wait(&sm.acc_full[buf], (uint32_t)((u >> 1) & 1));
        u_fence_after();
        const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)buf * C::kAccCols + tcol;
#pragma unroll
        for (int cq = 0; cq < kUQ / 32; ++cq) {
          float a[32], b[32];
          u_tmem_ld32(tl + (uint32_t)(cq * 32), a);            // hi*P_hi + lo*P_hi, queries [32cq, 32cq+32)
          u_tmem_ld32(tl + (uint32_t)(kUQ + cq * 32), b);      // hi*P_lo
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[cq * 32 + j] += a[j] + b[j];
        }
        u_fence_before();
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(u_smem(&sm.acc_empty[buf])) : "memory");
      }
      if (c != nchunks - 1 && nact > 0) continue;
      // ---- last chunk of this d-block: write it out
      const int64_t d = (int64_t)((int)blockIdx.x + t * (int)gridDim.x) * kUDBlock + lq * 32 + lane;
      if (zg) {
#pragma unroll
        for (int cq = 0; cq < kUQ / 32; ++cq) {
          float xv[32], dn[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = min(cq * 32 + j, Qg - 1);
            dn[j] = zg[q] + epi.eps;
            xv[j] = x0g ? __ldcg(x0g + (int64_t)q * D + d) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = cq * 32 + j;
            if (q < Qg) {
              const float v = sum[cq * 32 + j];
              const int64_t o = (int64_t)q * D + d;
              const float n = v / dn[j];
              if (numg) numg[o] = v;
              if (negg) negg[o] = n;
              if (x0g) x0g[o] = fmaf(-epi.scale, n, xv[j]);
              msum += fminf(fmaxf(n, -1e10f), 1e10f);
              if (blockIdx.x == 0 && lq == 0 && lane == 0) {
                if (denomg) denomg[q] = dn[j];
                if (gateg) gateg[q] = (!(epi.flags & SDN_EPI_GATE) || dn[j] > epi.gate_thr) ? 1 : 0;
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < kUQ; ++j) {
          if (j < Qg) {
            float* o = numg + (int64_t)j * D + d;
            *o = sum[j];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kUQ; ++j) sum[j] = 0.f;
    }   // units
    }
    if (zg && epi.mean_out && !epi.tile_out) {
      msum = warp_sum(msum);
      if (lane == 0) atomicAdd(epi.mean_out, msum * epi.inv_qd);
    }
  }
  u_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) accum_trace(tr, 6);
  if (warp == 1) {
    u_fence_after();
    u_tmem_dealloc(tmem, 2 * C::kAccCols);
  }
}

// ------------------------------------------------------------------------------------------ phase B on every SM
// k_umma_accum runs D / 128 CTAs (128 of the 148 SMs at SD-1.4 shapes) and every CTA is paced at one 48 KiB stage per
// ~0.85 us whatever is switched off (stages, requests, MMAs: profiles/r02_experiments.txt), so idle SMs are lost
// bandwidth.  Here the (d-block, 64-row block) units are dealt in d-block-major order in equal contiguous ranges to
// min(#SMs, units) CTAs (one per SM: all co-resident).  A d-block whose row blocks span several CTAs (2-3 at SD-1.4 shapes)
// is finished by the CTA that holds its FIRST row block -- for that CTA it is the last segment of its range -- and the
// others (for which it is the first or the only segment) publish their partial in `scratch` and count themselves in on a
// flag per (d-block, query group).  The finisher therefore waits, at the very end of its work, for partials that were
// mostly written long before; it adds them in slot order (deterministic).  A first version chained the partials in CTA
// order (each CTA waiting for its predecessor's LAST segment): that serialises the whole grid, 514 us instead of 49.
// Dense accumulate, chains of <= 64 row blocks (N <= 4096: no chunked TMEM drain needed), no bank-row split.
struct BalArgs {
  float* scratch;       // [dblocks][slots][G * 64 query rows][128 d] partials of the non-finishing CTAs
  unsigned* flags;      // [dblocks * G] arrivals (zeroed by the query-prepare kernel of the pass)
  int rblocks;          // 64-row blocks of the bank
  int slots;            // most CTAs a d-block can span, minus the finisher
};

template <int G>
__global__ void __launch_bounds__(UCfg<G>::kThreads, 1)
k_umma_accum_bal(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_hi,
                 const __grid_constant__ CUtensorMap tm_lo, float* __restrict__ num, int64_t D, int Q, int use_lo,
                 int p_group_rows, const BalArgs bal, const AccumEpi epi) {
  using C = UCfg<G>;
  pdl_launch_dependents();
  extern __shared__ unsigned char smem_raw[];
  __shared__ float z_half[G][2][kUQ], z_sum[G][kUQ];
  const USmem sm = u_carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rblocks = bal.rblocks;
  const int dblocks = (int)(D / kUDBlock);
  const int units = dblocks * rblocks;
  const int ncta = (int)gridDim.x, cta = (int)blockIdx.x;
  const int per = units / ncta, rem = units % ncta;
  const int u0 = cta * per + min(cta, rem), u1 = u0 + per + (cta < rem ? 1 : 0);
  const int nun = u1 - u0;
  auto cta_of_unit = [&](int u) { return u < (per + 1) * rem ? u / (per + 1) : rem + (u - (per + 1) * rem) / max(per, 1); };

  auto load_stage = [&](int it, int what) {
    const int u = u0 + it, db = u / rblocks, rb = u - db * rblocks;
    const int d0 = db * kUDBlock, rc = rb * kUK;
    const int s = it % C::kStages;
    uint8_t* st = sm.tiles + (size_t)s * C::kStageBytes;
    if (what & 2) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        u_tma_2d(st + (size_t)g * kTileBytes, &tm_p, 0, g * p_group_rows + rc, &sm.full[s]);
        u_tma_2d(st + (size_t)g * kTileBytes + 8192, &tm_p, 64, g * p_group_rows + rc, &sm.full[s]);
      }
    }
    if (!(what & 1)) return;
    u_mbar_expect_tx(&sm.full[s], (uint32_t)(G + 1 + (use_lo ? 1 : 0)) * kTileBytes);
    u_tma_2d(st + C::kHiOff, &tm_hi, d0, rc, &sm.full[s]);
    u_tma_2d(st + C::kHiOff + 8192, &tm_hi, d0 + 64, rc, &sm.full[s]);
    if (use_lo) {
      u_tma_2d(st + C::kLoOff, &tm_lo, d0, rc, &sm.full[s]);
      u_tma_2d(st + C::kLoOff + 8192, &tm_lo, d0 + 64, rc, &sm.full[s]);
    }
  };
  const int npre = min(nun, C::kStages);
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) { u_mbar_init(&sm.full[s], 1); u_mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { u_mbar_init(&sm.acc_full[b], 1); u_mbar_init(&sm.acc_empty[b], 4 * G); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    u_prefetch_map(&tm_p); u_prefetch_map(&tm_hi); u_prefetch_map(&tm_lo);
    // the bank tiles of the first stages do not depend on the weights: they stream while the weights kernel finishes
    for (int it = 0; it < npre; ++it) load_stage(it, 1);
    pdl_wait();
    for (int it = 0; it < npre; ++it) load_stage(it, 2);
  }
  pdl_wait();
  if (warp == 1) u_tmem_alloc(sm.tmem_base, 2 * C::kAccCols);
  u_fence_before();
  __syncthreads();
  u_fence_after();
  const uint32_t tmem = *sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = npre; it < nun; ++it) {
        u_mbar_wait(&sm.empty[it % C::kStages], (uint32_t)(((it / C::kStages) + 1) & 1));
        load_stage(it, 3);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_full = u_idesc(kUDBlock, kUStack, 1, 1);   // hi * [P_hi | P_lo]
      constexpr uint32_t id_half = u_idesc(kUDBlock, kUQ, 1, 1);       // lo * P_hi
      int seg = 0;
      for (int it = 0; it < nun;) {
        const int db = (u0 + it) / rblocks;
        const int it_end = min(nun, (db + 1) * rblocks - u0);          // the units of this d-block that are mine
        const int buf = seg & 1;
        if (seg >= 2) {
          u_mbar_wait(&sm.acc_empty[buf], (uint32_t)(((seg >> 1) + 1) & 1));
          u_fence_after();
        }
        const uint32_t acc = tmem + (uint32_t)buf * C::kAccCols;
        for (int first = it; it < it_end; ++it) {
          const int s = it % C::kStages;
          u_mbar_wait(&sm.full[s], (uint32_t)((it / C::kStages) & 1));
          u_fence_after();
          const uint32_t base = u_smem(sm.tiles + (size_t)s * C::kStageBytes);
#pragma unroll
          for (int kk = 0; kk < kUK / 16; ++kk) {
            const uint64_t ah = u_desc(base + C::kHiOff + kk * 2048, 8192, 1024);        // bank^T, MN-major
            const uint64_t al = u_desc(base + C::kLoOff + kk * 2048, 8192, 1024);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const uint64_t b = u_desc(base + (uint32_t)g * kTileBytes + kk * 2048, 8192, 1024);   // P_g^T, MN-major
              u_mma(acc + (uint32_t)g * kUStack, ah, b, id_full, (it > first || kk > 0) ? 1u : 0u);
              if (use_lo) u_mma(acc + (uint32_t)g * kUStack, al, b, id_half, 1u);
            }
          }
          u_commit(&sm.empty[s]);
        }
        u_commit(&sm.acc_full[buf]);
        ++seg;
      }
    }
  } else {
    // epilogue: warps 2..5 serve query group 0, warps 6..9 group 1; TMEM lane = d within the block, column = stacked q
    const int lq = warp & 3;
    const int g = (warp - 2) >> 2;
    const int tg = (warp - 2 - 4 * g) * 32 + lane;               // 0..127 within the group's four warps
    const int Qg = min(kUQ, Q - g * kUQ);
    const int64_t qoff = (int64_t)g * kUQ * D;
    float* const numg = num ? num + qoff : nullptr;
    const float* zg = epi.z ? epi.z + g * kUQ : nullptr;
    if (epi.zpart) {
      const int qz = tg & (kUQ - 1), half = tg >> 6;
      const float* zp = epi.zpart + (int64_t)g * epi.zpart_stride + qz;
      float a = 0.f;
      for (int b0 = half; b0 < epi.nzpart; b0 += 16) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = b0 + 2 * u < epi.nzpart ? __ldcg(zp + (int64_t)(b0 + 2 * u) * kUStack) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) a += v[u];
      }
      z_half[g][half][qz] = a;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if (tg < kUQ) {
        const float zz = z_half[g][0][tg] + z_half[g][1][tg];
        z_sum[g][tg] = zz;
        if (cta == 0 && epi.z_out && tg < Qg) epi.z_out[g * kUQ + tg] = zz;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      zg = epi.z_only ? nullptr : z_sum[g];
    }
    float* const x0g = epi.x0 ? epi.x0 + qoff : nullptr;
    float* const negg = epi.neg_out ? epi.neg_out + qoff : nullptr;
    float* const denomg = epi.denom_out ? epi.denom_out + g * kUQ : nullptr;
    int32_t* const gateg = epi.gate_out ? epi.gate_out + g * kUQ : nullptr;
    const uint32_t tcol = (uint32_t)(g * kUStack);
    float msum = 0.f;
    int seg = 0;
#pragma unroll 1
    for (int it = 0; it < nun; ++seg) {
      const int db = (u0 + it) / rblocks;
      const int rb_first = (u0 + it) - db * rblocks;
      const int it_end = min(nun, (db + 1) * rblocks - u0);
      const int rb_last = (u0 + it_end - 1) - db * rblocks;
      it = it_end;
      // the CTA that holds the d-block's first row block finishes it (that segment is the LAST of its range, or the whole
      // d-block); every other CTA of the chain only publishes its partial (its FIRST or only segment)
      const bool finisher = rb_first == 0;
      const int c_first = finisher ? cta : cta_of_unit(db * rblocks);
      const int ncontrib = finisher ? cta_of_unit(db * rblocks + rblocks - 1) - cta : 0;
      (void)rb_last;
      const int buf = seg & 1;
      u_mbar_wait(&sm.acc_full[buf], (uint32_t)((seg >> 1) & 1));
      u_fence_after();
      const int dl = lq * 32 + lane;
      const int64_t d = (int64_t)db * kUDBlock + dl;
      const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)buf * C::kAccCols + tcol;
      // [d-block][slot][group][q][128 d]; slot = position in the chain - 1
      const int64_t slot_stride = (int64_t)G * kUQ * kUDBlock;
      float* const sc0 = bal.scratch + ((int64_t)db * bal.slots * G + g) * kUQ * kUDBlock + dl;
      float* const sc = sc0 + (int64_t)(finisher ? 0 : cta - c_first - 1) * slot_stride;
      unsigned* const fl = bal.flags + db * G + g;
      if (ncontrib > 0) {
        if (lane == 0) {
          uint32_t spin = 0;
          while (u_ld_acquire(fl) < (unsigned)ncontrib) {
            __nanosleep(64);
            if (++spin > (1u << 24)) __trap();     // a broken chain must not hang the GPU
          }
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int c = 0; c < kUQ / 32; ++c) {
        float a[32];
        {
          float b[32];
          u_tmem_ld32(tl + (uint32_t)(c * 32), a);            // hi*P_hi + lo*P_hi, queries [32c, 32c+32)
          u_tmem_ld32(tl + (uint32_t)(kUQ + c * 32), b);      // hi*P_lo
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] += b[j];
        }
        if (!finisher) {
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(sc + (int64_t)(c * 32 + j) * kUDBlock, a[j]);
          continue;
        }
#pragma unroll 1
        for (int s = 0; s < ncontrib; ++s) {                   // mine + slot 0 + slot 1 ...: a fixed order
          float pv[32];
          const float* ps = sc0 + (int64_t)s * slot_stride;
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = __ldcg(ps + (int64_t)(c * 32 + j) * kUDBlock);
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] += pv[j];
        }
        if (zg) {
          float xv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = min(c * 32 + j, Qg - 1);
            xv[j] = x0g ? __ldcg(x0g + (int64_t)q * D + d) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = c * 32 + j;
            if (q < Qg) {
              const float dn = zg[q] + epi.eps;
              const int64_t o = (int64_t)q * D + d;
              const float n = a[j] / dn;
              if (numg) numg[o] = a[j];
              if (negg) negg[o] = n;
              if (x0g) x0g[o] = fmaf(-epi.scale, n, xv[j]);
              msum += fminf(fmaxf(n, -1e10f), 1e10f);
              if (db == 0 && lq == 0 && lane == 0) {
                if (denomg) denomg[q] = dn;
                if (gateg) gateg[q] = (!(epi.flags & SDN_EPI_GATE) || dn > epi.gate_thr) ? 1 : 0;
              }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = c * 32 + j;
            if (q < Qg) numg[(int64_t)q * D + d] = a[j];
          }
        }
      }
      u_fence_before();
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(u_smem(&sm.acc_empty[buf])) : "memory");
      if (!finisher) {
        // publish: every thread's stores, then a gpu-scope fence, then the group's barrier, then ONE arrival
        __threadfence();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        if (tg == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(fl) : "memory");
      }
    }
    if (zg && epi.mean_out) {
      msum = warp_sum(msum);
      if (lane == 0) atomicAdd(epi.mean_out, msum * epi.inv_qd);
    }
  }
  u_fence_before();
  __syncthreads();
  if (warp == 1) {
    u_fence_after();
    u_tmem_dealloc(tmem, 2 * C::kAccCols);
  }
}

// x0[q][d] -= scale * num[q][d] / (z[q] + eps) with num read from phase B's tile-contiguous scratch [D/128][64 G][128]:
// every access of a warp is 512 contiguous bytes.  Same arithmetic as the in-kernel correction
// (fmaf(-scale, num * (1 / (z + eps)), x0)); the optional outputs (num, neg, mean of the clamped neg) are written here too.
__global__ void __launch_bounds__(256)
k_umma_untile(const float* __restrict__ tiles, float* __restrict__ x0, float* __restrict__ num_out,
              float* __restrict__ neg_out, const float* __restrict__ z, float eps, int Q, int64_t D, int G, float scale,
              float* __restrict__ mean_out, float inv_qd) {
  __shared__ float wsum[8];
  pdl_launch_dependents();
  pdl_wait();
  const int64_t e = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  float m = 0.f;
  if (e < (int64_t)Q * D) {
    const int q = (int)(e / D);
    const int64_t d = e - (int64_t)q * D;
    const int g = q / kUQ, ql = q % kUQ;
    const int64_t db = d / kUDBlock;
    const int dl = (int)(d % kUDBlock);
    const float4 a = __ldcg(reinterpret_cast<const float4*>(tiles + ((db * G + g) * kUQ + ql) * kUDBlock + dl));
    if (num_out) *reinterpret_cast<float4*>(num_out + e) = a;
    if (x0) {                           // null: a partial-sums call, only the sums go out (row-major)
      float4 x = __ldcg(reinterpret_cast<const float4*>(x0 + e));
      const float ri = 1.f / (__ldcg(z + q) + eps);
      const float4 n = make_float4(a.x * ri, a.y * ri, a.z * ri, a.w * ri);
      x.x = fmaf(-scale, n.x, x.x);
      x.y = fmaf(-scale, n.y, x.y);
      x.z = fmaf(-scale, n.z, x.z);
      x.w = fmaf(-scale, n.w, x.w);
      *reinterpret_cast<float4*>(x0 + e) = x;
      if (neg_out) *reinterpret_cast<float4*>(neg_out + e) = n;
      m = fminf(fmaxf(n.x, -1e10f), 1e10f) + fminf(fmaxf(n.y, -1e10f), 1e10f) + fminf(fmaxf(n.z, -1e10f), 1e10f) +
          fminf(fmaxf(n.w, -1e10f), 1e10f);
    }
  }
  if (mean_out) {                       // the reference's logging scalar: mean of the clamped projection
    m = warp_sum(m);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += wsum[w];
      atomicAdd(mean_out, t * inv_qd);
    }
  }
}

// num[q][d] = sum_s part[s][q][d] in split order (deterministic; the splits used to atomicAdd into num).  Exits like
// k_umma_accum does when the listed accumulate already produced the result.
__global__ void __launch_bounds__(256)
k_umma_splitsum(const float* __restrict__ part, int nsplit, int64_t split_stride, int64_t QD, int Q, float* __restrict__ num,
                const int* __restrict__ list_count, const int* __restrict__ dense_flag) {
  if (dense_flag && __ldg(dense_flag)) {
  } else if (list_count && siglist_all_short(list_count, Q)) {
    return;
  }
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i >= QD) return;
  float4 s4 = *reinterpret_cast<const float4*>(part + i);
  for (int sp = 1; sp < nsplit; ++sp) {
    const float4 p = *reinterpret_cast<const float4*>(part + (int64_t)sp * split_stride + i);
    s4.x += p.x; s4.y += p.y; s4.z += p.z; s4.w += p.w;
  }
  *reinterpret_cast<float4*>(num + i) = s4;
}

// ------------------------------------------------------------------------------------------ host side
std::atomic<bool> g_skip_negligible{true};

namespace {
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

bool load_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  return g_encode != nullptr;
}

// 2-D bf16 row-major tensor [rows][cols], box [box_rows][box_cols], 128-byte swizzle.
int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SDN_OK : SDN_E_PARAM;
}

// bf16 tensor of `rank` dimensions (innermost first; strides in BYTES for dimensions 1..rank-1, need not be sorted),
// 128-byte swizzle on the innermost dimension.
int make_map_nd(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* stride_bytes,
                const uint32_t* box) {
  cuuint64_t d[5]; cuuint64_t st[4]; cuuint32_t b[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = stride_bytes[i];
  const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SDN_OK : SDN_E_PARAM;
}

}  // namespace

size_t umma_accum_trace_read(void* host_out, size_t bytes) {
  const size_t n = std::min(bytes, sizeof(unsigned long long) * 256 * 8);
  if (cudaMemcpyFromSymbol(host_out, g_accum_trace, n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
  if (!load_encode()) return SDN_E_DEVICE;
  return make_map(map, base, rows, cols, box_rows, box_cols);
}

namespace {
int umma_ksplit(int64_t npad, int64_t D) {
  const int row_tiles = (int)(npad / kUBankTile);
  const int kblocks = (int)(D / kUK);
  // one task per SM if possible, at least 8 K-blocks per task (fill/drain), at most 64 partial buffers
  int ksplit = std::max(1, std::min(kblocks / 8, kNumSMs / row_tiles));   // never a second, nearly empty wave
  return std::min(ksplit, 64);
}

struct UmmaLayout {
  int G;               // query groups of 64 rows per pass (1 or 2)
  int64_t npad;        // bank rows padded to 128
  int ksplit;
  int nflags;          // row blocks of 64 bank rows
  int xsq_nparts;
  size_t off_x, off_s, off_p, off_z, off_f, off_q, off_c, off_n, total;
  int nsplit;          // bank-row splits of phase B (partials in the workspace, summed in order)
  bool bal;            // shape the balanced phase B (k_umma_accum_bal) takes: no row split, <= 64 row blocks
  int bal_slots;       // partial slots per d-block of the balanced phase B
  // per-group strides (elements)
  int64_t split_stride, zpart_stride;
};
UmmaLayout umma_layout(int64_t Q, int64_t N, int64_t D) {
  UmmaLayout L;
  L.G = Q > kUQ ? 2 : 1;
  L.npad = cdiv(N, 128) * 128;
  L.ksplit = umma_ksplit(L.npad, D);
  L.nflags = (int)(L.npad / kUK);
  L.xsq_nparts = (int)cdiv(D, 1024);
  L.split_stride = L.npad * kUStack;
  L.zpart_stride = (L.npad / kWRows) * kUStack + 64;
  const size_t G = (size_t)L.G;
  size_t o = 0;
  L.off_x = o; o += G * kUStack * D * 2;                     // X planes [128 G][D] (layout: k_umma_xprep)
  o = (o + 255) / 256 * 256;
  L.off_s = o; o += (size_t)L.ksplit * L.split_stride * 4;   // S^T [ksplit][Npad][128], one partial per K split
  o = (o + 255) / 256 * 256;
  L.off_p = o; o += G * kUStack * L.npad * 2;                // P planes [G][Npad][128]
  o = (o + 255) / 256 * 256;
  L.off_z = o; o += G * L.zpart_stride * 4;                  // per-block z sums | maxima, per group
  o = (o + 255) / 256 * 256;
  // row-block flags | kmax[G 64] | count[G 64] | dense (4 words) | rows[G 64][cap] | ks[G 64][cap]
  L.off_f = o; o += (size_t)L.nflags * 4 + G * kUQ * 4 + G * kUQ * 4 + 16 + G * kUQ * kListCap * 8 + 256;
  o = (o + 255) / 256 * 256;
  L.off_q = o; o += G * L.xsq_nparts * kUQ * 4;              // ||x||^2 partials of the fused query prepare
  o = (o + 255) / 256 * 256;
  // arrival counters of the weights step (fused: one per row tile + 1), then the chain flags of the balanced phase B
  L.off_c = o; o += (size_t)(L.npad / kUBankTile + 2 + (D / kUDBlock) * G) * 4;
  o = (o + 255) / 256 * 256;
  {
    const int dblocks = (int)(D / kUDBlock), rblocks = (int)(L.npad / kUK);
    L.nsplit = std::max(1, std::min(rblocks / 8, kNumSMs / std::max(1, dblocks)));
  }
  // phase-B split partials [nsplit][Q][D], or the hand-over buffer of the balanced phase B [D/128][64 G][128]
  L.bal = L.nsplit == 1 && L.npad / kUK <= kBChunk;
  {
    const int dblocks = (int)(D / kUDBlock), rblocks = (int)(L.npad / kUK);
    const int per = std::max(1, dblocks * rblocks / std::min(kNumSMs, dblocks * rblocks));
    L.bal_slots = std::max(1, (rblocks - 1 + per - 1) / per);
  }
  L.off_n = o;
  o += L.nsplit > 1 ? (size_t)L.nsplit * (size_t)Q * D * 4 : (L.bal ? (size_t)L.bal_slots * G * kUQ * (size_t)D * 4 : 0);
  L.total = (o + 255) / 256 * 256;
  return L;
}
}  // namespace

bool umma_supported(int64_t Q, int64_t N, int64_t D, const void* planes) {
  return planes != nullptr && Q >= 1 && N >= 1 && D >= kUDBlock && D % kUDBlock == 0 &&
         N < (1ll << 30) && D < (1ll << 30);      // a row of x0 in bytes fits 32 bits (epilogue addressing)
}

size_t umma_workspace_bytes(int64_t Q, int64_t N, int64_t D) {
  if (Q < 1 || D % kUDBlock) return 0;
  return umma_layout(Q, N, D).total;
}

static int umma_pass(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                     const float* xsq, int64_t Q, float inv2s2, int power, float alpha, float* num, float* z,
                     float* k_out, void* ws, size_t ws_bytes, cudaStream_t st, const AccumEpi* epi,
                     float* zero_word, bool bf16_bank);

// One two-phase pass over the bank serves up to 128 query rows (two groups of 64: the TMEM of an SM holds two
// double-buffered 128 x 128 accumulators); more rows take ceil(Q / 128) passes.
int umma_partial(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                 const float* xsq, int64_t Q, float inv2s2, int power, float alpha, float* num, float* z,
                 float* k_out, void* ws, size_t ws_bytes, cudaStream_t st, bool bf16_bank) {
  if (!umma_supported(Q, N, D, planes)) return SDN_E_UNSUPPORTED;
  for (int64_t q0 = 0; q0 < Q; q0 += 2 * kUQ) {
    const int64_t qn = std::min<int64_t>(2 * kUQ, Q - q0);
    const int rc = umma_pass(planes, sqnorm, N, D, xq + q0 * D, xsq ? xsq + q0 : nullptr, qn, inv2s2, power, alpha,
                             num ? num + q0 * D : nullptr, z + q0, k_out ? k_out + q0 * N : nullptr, ws, ws_bytes, st,
                             nullptr, nullptr, bf16_bank);
    if (rc) return rc;
  }
  return SDN_OK;
}

// conditioning() for batched queries on one GPU: query planes + ||x||^2 in one kernel, correction fused into
// phase B's epilogue.  z must be a valid [Q] buffer (scratch if the caller does not want it).
int umma_conditioning(const void* planes, const float* sqnorm, int64_t N, int64_t D, float* x0_inout, int64_t Q,
                      float inv2s2, int power, float alpha, float eps, float scale, float gate_thr, int flags,
                      float* num_out, float* z, float* neg_out, float* denom_out, int32_t* gate_out, float* mean_out,
                      float* k_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!umma_supported(Q, N, D, planes)) return SDN_E_UNSUPPORTED;
  {
    const int dblocks = (int)(D / kUDBlock), rblocks = (int)(cdiv(N, 128) * 128 / kUK);
    if (std::max(1, std::min(rblocks / 8, kNumSMs / dblocks)) != 1)
      return SDN_E_UNSUPPORTED;   // phase B would split the bank rows: the correction cannot be fused
  }
  for (int64_t q0 = 0; q0 < Q; q0 += 2 * kUQ) {
    const int64_t qn = std::min<int64_t>(2 * kUQ, Q - q0);
    AccumEpi epi{};
    epi.z = z + q0; epi.eps = eps; epi.scale = scale; epi.gate_thr = gate_thr; epi.flags = flags;
    epi.x0 = x0_inout + q0 * D; epi.neg_out = neg_out ? neg_out + q0 * D : nullptr;
    epi.denom_out = denom_out ? denom_out + q0 : nullptr; epi.gate_out = gate_out ? gate_out + q0 : nullptr;
    epi.mean_out = mean_out; epi.inv_qd = 1.f / (float)(Q * D);
    const int rc = umma_pass(planes, sqnorm, N, D, x0_inout + q0 * D, nullptr, qn, inv2s2, power, alpha,
                             num_out ? num_out + q0 * D : nullptr, z + q0, k_out ? k_out + q0 * N : nullptr, ws,
                             ws_bytes, st, &epi, q0 == 0 ? mean_out : nullptr, false);
    if (rc) return rc;
  }
  return SDN_OK;
}

namespace {
// Launch with (or without) the programmatic-stream-serialisation attribute: the kernel may start before the previous
// kernel of the stream has finished and orders itself with griddepcontrol.wait (pdl_wait above).
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("SDN_PDL"); return !(e && atoi(e) == 0); }();
  return on;
}

template <int G>
int configure_kernels() {
  SDN_CUDA_OK(cudaFuncSetAttribute(k_umma_dots<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUSmemBytes));
  SDN_CUDA_OK(cudaFuncSetAttribute(k_umma_accum<false, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUSmemBytes));
  SDN_CUDA_OK(cudaFuncSetAttribute(k_umma_accum<true, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUSmemBytes));
  SDN_CUDA_OK(cudaFuncSetAttribute(k_umma_accum_bal<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUSmemBytes));
  return SDN_OK;
}

struct AccumLaunch {
  CUtensorMap tm_p, tm_hi, tm_lo, tm_bank4, tm_p3;
  float* num; int64_t D; int64_t split_stride; int Q, rblocks, nsplit, use_lo, p_group_rows;
  const int* flags; const int* count; const int* dense;
  AccumEpi e;
  int gridx; bool chunked; bool pdl;
};
template <int G>
void launch_accum(const AccumLaunch& a, cudaStream_t st) {
  const dim3 grid(a.gridx, a.nsplit);
  if (a.chunked)
    launch_ex(k_umma_accum<true, G>, grid, dim3(UCfg<G>::kThreads), kUSmemBytes, st, a.pdl, a.tm_p, a.tm_hi, a.tm_lo,
              a.tm_bank4, a.tm_p3, a.num,
              a.D, a.Q, a.rblocks, a.nsplit, a.split_stride, a.use_lo, a.p_group_rows, a.flags, a.count, a.dense, a.e);
  else
    launch_ex(k_umma_accum<false, G>, grid, dim3(UCfg<G>::kThreads), kUSmemBytes, st, a.pdl, a.tm_p, a.tm_hi, a.tm_lo,
              a.tm_bank4, a.tm_p3, a.num,
              a.D, a.Q, a.rblocks, a.nsplit, a.split_stride, a.use_lo, a.p_group_rows, a.flags, a.count, a.dense, a.e);
}
}  // namespace

static int umma_pass(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                     const float* xsq, int64_t Q, float inv2s2, int power, float alpha, float* num, float* z,
                     float* k_out, void* ws, size_t ws_bytes, cudaStream_t st, const AccumEpi* epi,
                     float* zero_word, bool bf16_bank) {
  const UmmaLayout L = umma_layout(Q, N, D);
  const int G = L.G;
  if (!ws || ws_bytes < L.total) return SDN_E_WORKSPACE;
  if (!load_encode()) return SDN_E_DEVICE;
  static std::atomic<bool> configured[kMaxDevices];
  const int dev = device_slot();
  if (!configured[dev].load(std::memory_order_acquire)) {
    int rc;
    if ((rc = configure_kernels<1>())) return rc;
    if ((rc = configure_kernels<2>())) return rc;
    configured[dev].store(true, std::memory_order_release);
  }
  char* w = static_cast<char*>(ws);
  __nv_bfloat16* xpl = reinterpret_cast<__nv_bfloat16*>(w + L.off_x);
  float* S_T = reinterpret_cast<float*>(w + L.off_s);
  __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(w + L.off_p);
  const __nv_bfloat16* hi = static_cast<const __nv_bfloat16*>(planes);
  const __nv_bfloat16* lo = hi + N * D;

  // tensor maps depend only on (device, planes, workspace, N, D, G): a small per-process cache (several projectors /
  // devices in one process each keep their entry; least recently used is replaced)
  struct MapCache { int dev; const void* planes; const void* ws; int64_t N, D; int G; CUtensorMap m[9]; bool merged; uint64_t stamp; };
  static MapCache cache[8];
  static int cache_n = 0;
  static uint64_t cache_clock = 0;
  static std::mutex cache_mu;
  CUtensorMap tm_x, tm_hiA, tm_loA, tm_p, tm_hiB, tm_loB, tm_bankA, tm_bankB, tm_p3;
  bool merged = false;
  {
    std::lock_guard<std::mutex> lk(cache_mu);
    MapCache* hit = nullptr;
    for (int i = 0; i < cache_n && !hit; ++i)
      if (cache[i].dev == dev && cache[i].planes == planes && cache[i].ws == ws && cache[i].N == N && cache[i].D == D && cache[i].G == G)
        hit = &cache[i];
    if (!hit) {
      MapCache c{};
      int rc;
      if ((rc = make_map(&c.m[0], xpl, (uint64_t)G * kUStack, D, kUStack, kUK))) return rc;
      if ((rc = make_map(&c.m[1], hi, N, D, kUBankTile, kUK))) return rc;
      if ((rc = make_map(&c.m[2], lo, N, D, kUBankTile, kUK))) return rc;
      if ((rc = make_map(&c.m[3], P, (uint64_t)G * L.npad, kUStack, kUK, 64))) return rc;
      if ((rc = make_map(&c.m[4], hi, N, D, kUK, 64))) return rc;
      if ((rc = make_map(&c.m[5], lo, N, D, kUK, 64))) return rc;
      // merged requests: hi + lo tile of phase A in one 3-D box; both planes x both 64-wide halves of phase B's bank^T
      // tile in one 4-D box; both halves of a weight tile in one 3-D box.  (Older drivers may refuse the unsorted strides:
      // the 2-D maps above then do the job.)
      // Measured: accepted by the driver, bit-identical results, NO change in time (cfg3 phase A 41.5, phase B 49.7 us
      // either way) -- the request count is not what bounds these kernels.  Off unless SDN_UMMA_MERGED_TMA=1.
      static const bool want_merged = [] { const char* e = getenv("SDN_UMMA_MERGED_TMA"); return e && atoi(e) != 0; }();
      c.merged = want_merged;
      if (c.merged) {
        const uint64_t dA[3] = {(uint64_t)D, (uint64_t)N, 2}, sA[2] = {(uint64_t)D * 2, (uint64_t)N * D * 2};
        const uint32_t bA[3] = {kUK, kUBankTile, 2};
        const uint64_t dB[4] = {64, (uint64_t)N, (uint64_t)(D / 64), 2}, sB[3] = {(uint64_t)D * 2, 128, (uint64_t)N * D * 2};
        const uint32_t bB[4] = {64, kUK, 2, 2};
        const uint64_t dP[3] = {64, (uint64_t)G * L.npad, 2}, sP[2] = {(uint64_t)kUStack * 2, 128};
        const uint32_t bP[3] = {64, kUK, 2};
        const int r6 = make_map_nd(&c.m[6], hi, 3, dA, sA, bA), r7 = make_map_nd(&c.m[7], hi, 4, dB, sB, bB),
                  r8 = make_map_nd(&c.m[8], P, 3, dP, sP, bP);
        if (getenv("SDN_UMMA_DEBUG")) fprintf(stderr, "[sdn_umma] merged tensor maps: phase A %d, phase B bank %d, weights %d\n", r6, r7, r8);
        if (r6 || r7 || r8) c.merged = false;
      }
      c.dev = dev; c.planes = planes; c.ws = ws; c.N = N; c.D = D; c.G = G;
      int slot = cache_n;
      if (cache_n == 8) {
        slot = 0;
        for (int i = 1; i < 8; ++i) if (cache[i].stamp < cache[slot].stamp) slot = i;
      } else {
        ++cache_n;
      }
      cache[slot] = c;
      hit = &cache[slot];
    }
    hit->stamp = ++cache_clock;
    tm_x = hit->m[0]; tm_hiA = hit->m[1]; tm_loA = hit->m[2]; tm_p = hit->m[3]; tm_hiB = hit->m[4]; tm_loB = hit->m[5];
    merged = hit->merged;
    tm_bankA = merged ? hit->m[6] : tm_hiA; tm_bankB = merged ? hit->m[7] : tm_hiB; tm_p3 = merged ? hit->m[8] : tm_p;
  }

  // significant-row lists, shared by the groups of the pass (flags and the dense mark are unions over the groups;
  // counts, rows and weights are indexed by the query row within the pass)
  const int nflags = L.nflags;
  SigLists lists{};
  lists.flags = reinterpret_cast<int*>(w + L.off_f);
  float* kmax = reinterpret_cast<float*>(lists.flags + nflags);
  lists.count = reinterpret_cast<int*>(kmax + G * kUQ);
  lists.dense = lists.count + G * kUQ;
  lists.rows = lists.dense + 4;
  lists.ks = reinterpret_cast<float*>(lists.rows + G * kUQ * kListCap);
  // (SPELL weights are not exponentially peaked: no significant-row lists in that mode)
  const bool sparse = g_skip_negligible.load(std::memory_order_relaxed) && nflags <= kMaxActive && (num || epi) &&
                      power != kPowerSparse;
  float* zpart = reinterpret_cast<float*>(w + L.off_z);
  float* xsq_part = reinterpret_cast<float*>(w + L.off_q);
  auto group_rows = [&](int g) { return (int)std::min<int64_t>(kUQ, Q - (int64_t)g * kUQ); };

  // query planes (one launch per group of 64 rows); the first block also resets the pass's lists and counters
  const int row_tiles = (int)(L.npad / kUBankTile);
  int* const counters = reinterpret_cast<int*>(w + L.off_c);
  // SDN_UMMA_FUSED_WEIGHTS=1: the last K-split CTA of a row tile makes the tile's weights (dots_tail) instead of the
  // k_umma_weights + k_umma_zreduce launches.  Measured slower (cfg3: dots 41 -> 64 us against 12.7 + 4 us saved): one
  // SM ingests the tile's K-split partials at ~45 GB/s.  Kept for tiny shards / experiments, off by default.
  static const bool fuse_weights = [] { const char* e = getenv("SDN_UMMA_FUSED_WEIGHTS"); return e && atoi(e) != 0; }();
  int pid = g_prof.begin(xsq ? "k_umma_xprep" : "k_umma_qprep", st);
  for (int g = 0; g < G; ++g) {
    StartClear clr{};
    if (g == 0) {
      if (sparse) { clr.lists = lists; clr.lists.ncount = G * kUQ; clr.nflags = nflags; }
      clr.counters = counters; clr.ncounters = row_tiles + 2 + (int)(D / kUDBlock) * G;
    }
    const float* xg = xq + (int64_t)g * kUQ * D;
    __nv_bfloat16* pg = xpl + (int64_t)g * kUQ * D;      // hi part of the group's row 0; lo parts G * 64 rows below
    if (xsq) {
      k_umma_xprep<<<dim3((unsigned)cdiv(D, 1024), kUQ), 256, 0, st>>>(xg, group_rows(g), D, pg, G * kUQ, clr);
    } else {
      k_umma_qprep<<<dim3((unsigned)L.xsq_nparts, kUQ), 256, 0, st>>>(xg, group_rows(g), D, pg, G * kUQ,
                                                                     xsq_part + (int64_t)g * L.xsq_nparts * kUQ,
                                                                     g == 0 ? zero_word : nullptr, clr);
    }
    SDN_LAUNCHED();
  }
  g_prof.end(pid, st);

  // phase A: split K (= D) so that roughly every SM gets one task; the last arrival of a row tile makes its weights
  const int kblocks = (int)(D / kUK);
  DotsTail tail{};
  if (fuse_weights) {
    tail.enabled = 1; tail.tile_count = counters; tail.sqnorm = sqnorm; tail.xsq = xsq; tail.xsq_part = xsq_part;
    tail.xsq_nparts = L.xsq_nparts; tail.N = (int)N; tail.Q = (int)Q; tail.row_tiles = row_tiles;
    tail.inv2s2 = inv2s2; tail.alpha = alpha; tail.power = power;
    tail.P = P; tail.p_group_stride = L.npad * kUStack; tail.zpart = zpart; tail.zpart_stride = L.zpart_stride;
    tail.z = z; tail.kmax = kmax; tail.dense_flag = sparse ? lists.dense : nullptr; tail.k_out = k_out;
  }
  // L2 carry-over between the phases: the last `keep` MB of the bank that phase A streams stay in the L2 (evict_last)
  // and phase B reads them first
  static const int keep_mb = [] { const char* e = getenv("SDN_UMMA_L2KEEP_MB"); return e ? atoi(e) : 0; }();
  tail.keep_from_row = -1;
  // on by default: cfg3 phase A 43.6 -> 41.5 us, step 92.2 -> 90.1 us (SDN_UMMA_INTERLEAVE_K=0 restores contiguous ranges)
  static const int interleave_k = [] { const char* e = getenv("SDN_UMMA_INTERLEAVE_K"); return e ? atoi(e) : 1; }();
  tail.interleave_k = interleave_k;
  tail.merged = merged ? 1 : 0;
  static const int dbg_noshared = [] { const char* e = getenv("SDN_UMMA_DBG_NOSHARED"); return e ? atoi(e) : 0; }();
  tail.dbg_noshared = dbg_noshared & 1;
  tail.dbg_nomma = (dbg_noshared >> 2) & 1;
  tail.dbg_nostore = (dbg_noshared >> 12) & 1;
  const bool l2keep = keep_mb > 0 && (num || epi) && L.nsplit == 1;
  if (l2keep) {
    const int64_t keep_rows = std::min<int64_t>(L.npad, (int64_t)keep_mb * 1000000 / (D * 4));
    tail.keep_from_row = (int)((L.npad - keep_rows) / kUBankTile * kUBankTile);
  }
  pid = g_prof.begin("k_umma_dots", st);
  const bool pdl = pdl_enabled();
  if (G == 1)
    launch_ex(k_umma_dots<1>, dim3(row_tiles, L.ksplit), dim3(kUThreadsA), kUSmemBytes, st, pdl,
              tm_x, tm_hiA, tm_loA, tm_bankA, S_T, L.split_stride, kblocks, L.ksplit, bf16_bank ? 0 : 1, tail);
  else
    launch_ex(k_umma_dots<2>, dim3(row_tiles, L.ksplit), dim3(kUThreadsA), kUSmemBytes, st, pdl,
              tm_x, tm_hiA, tm_loA, tm_bankA, S_T, L.split_stride, kblocks, L.ksplit, bf16_bank ? 0 : 1, tail);
  g_prof.end(pid, st);
  SDN_LAUNCHED();

  // weights, z, lists: per group (unless the fused step is on)
  // single GPU, dense accumulate, correction fused into phase B: z is summed by phase B's idle epilogue warps from this
  // kernel's partials -- no k_umma_zreduce launch (SDN_UMMA_ZREDUCE=1 keeps it)
  static const bool keep_zreduce = [] { const char* e = getenv("SDN_UMMA_ZREDUCE"); return e && atoi(e) != 0; }();
  const bool fuse_z = (epi != nullptr || (num != nullptr && L.nsplit == 1)) && !sparse && !fuse_weights && !keep_zreduce;
  const int fuse_z_rpb = (L.npad >= 3072 ? 2 : 1) * kWMaxLanes;      // small banks: one row per thread (more blocks)
  pid = g_prof.begin(fuse_weights ? "k_umma_siglist" : "k_umma_weights", st);
  for (int g = 0; g < G && !fuse_weights; ++g) {
    // one bank row per thread and step: few fat blocks (many rows in a loop, last block sums z) measured slower -- 8
    // warps per SM do not hide the L2 latency of the partial loads (14.7 us at cfg3 against 5 + 4 us for this kernel +
    // k_umma_zreduce).  fuse_z: 1024-thread blocks of 32 rows, so that phase B's epilogue warps have few partials to sum
    const int rpb = fuse_z ? fuse_z_rpb : kWRows;
    launch_ex(k_umma_weights, dim3((unsigned)cdiv(L.npad, rpb)), dim3((fuse_z ? kWMaxLanes : kWRows) * kUQ), 0, st, pdl,
              (const float*)S_T, L.split_stride, L.ksplit, g * kUQ, G == 1 ? kUQ : 0, sqnorm, xsq ? xsq + g * kUQ : nullptr,
              (const float*)(xsq_part + (int64_t)g * L.xsq_nparts * kUQ), L.xsq_nparts, (int)N, group_rows(g), inv2s2, power,
              alpha, P + (int64_t)g * L.npad * kUStack, zpart + (int64_t)g * L.zpart_stride,
              k_out ? k_out + (int64_t)g * kUQ * N : nullptr, rpb, (int)L.npad, (int*)nullptr, z + g * kUQ, kmax + g * kUQ,
              sparse ? lists.dense : (int*)nullptr);
    SDN_LAUNCHED();
  }
  for (int g = 0; g < G && !fuse_weights && !fuse_z; ++g) {
    launch_ex(k_umma_zreduce, dim3((unsigned)group_rows(g)), dim3(256), 0, st, pdl,
              (const float*)(zpart + (int64_t)g * L.zpart_stride), (int)(L.npad / kWRows), z + g * kUQ, kmax + g * kUQ,
              sparse ? lists.dense : (int*)nullptr);
    SDN_LAUNCHED();
  }
  if (sparse) {
    for (int g = 0; g < G; ++g) {
      SigLists lg = lists;
      lg.count = lists.count + g * kUQ; lg.rows = lists.rows + g * kUQ * kListCap; lg.ks = lists.ks + g * kUQ * kListCap;
      k_umma_siglist<<<(unsigned)(L.npad / kWRows), kWRows * kUQ, 0, st>>>(P + (int64_t)g * L.npad * kUStack, kmax + g * kUQ,
                                                                          (int)N, group_rows(g), lg);
      SDN_LAUNCHED();
    }
  }
  g_prof.end(pid, st);
  if (!num && !epi) return SDN_OK;   // z only (empirical_beta): no phase B

  // phase B: one CTA per 128 d, bank rows split when that leaves SMs idle
  const int dblocks = (int)(D / kUDBlock);
  const int rblocks = (int)(L.npad / kUK);
  const int nsplit = L.nsplit;
  AccumEpi e{};
  if (epi) {
    if (nsplit != 1) return SDN_E_UNSUPPORTED;
    e = *epi;
  }
  e.reverse = l2keep ? 1 : 0;
  e.merged = merged ? 1 : 0;
  e.dbg_noshared = (dbg_noshared >> 1) & 1;
  e.dbg_nomma = (dbg_noshared >> 3) & 1;
  e.dbg = dbg_noshared >> 4;
  if (fuse_z) {
    e.zpart = zpart; e.nzpart = (int)cdiv(L.npad, fuse_z_rpb); e.zpart_stride = L.zpart_stride; e.z_out = z;
    e.z_only = epi ? 0 : 1;
  }
  float* const part = reinterpret_cast<float*>(w + L.off_n);
  if (sparse) {
    pid = g_prof.begin("k_umma_listed_accum", st);
    k_umma_listed_accum<<<dim3((unsigned)cdiv(D, 1024), (unsigned)Q), 256, 0, st>>>(
        hi, bf16_bank ? nullptr : lo, D, (int)Q, lists, num, epi ? e.z : nullptr, e.eps, e.scale, e.gate_thr, e.flags, e.x0,
        e.neg_out, e.denom_out, e.gate_out, e.mean_out, e.inv_qd);
    g_prof.end(pid, st);
    SDN_LAUNCHED();
  }
  pid = g_prof.begin("k_umma_accum", st);
  AccumLaunch al{};
  al.tm_p = tm_p; al.tm_hi = tm_hiB; al.tm_lo = tm_loB; al.tm_bank4 = tm_bankB; al.tm_p3 = tm_p3; al.num = nsplit > 1 ? part : num; al.D = D; al.Q = (int)Q; al.rblocks = rblocks;
  al.nsplit = nsplit; al.split_stride = nsplit > 1 ? Q * D : 0; al.use_lo = bf16_bank ? 0 : 1; al.p_group_rows = (int)L.npad;
  al.flags = sparse ? lists.flags : nullptr; al.count = sparse ? lists.count : nullptr;
  al.dense = sparse ? lists.dense : nullptr; al.e = e;
  al.gridx = nsplit == 1 ? std::min(dblocks, kNumSMs) : dblocks;
  al.pdl = pdl;
  // chains longer than kBChunk row blocks are split over the two TMEM accumulators (fp32 register drain)
  al.chunked = (rblocks + nsplit - 1) / nsplit > kBChunk;
  // phase B on every SM (SDN_UMMA_BALANCED=1; measured no gain: cfg3 51.6 us against 49.5 us -- the kernel is bounded by a
  // chip-wide rate, not by the 20 idle SMs) when the shape allows and nothing is skipped
  static const bool want_bal = [] { const char* e = getenv("SDN_UMMA_BALANCED"); return e && atoi(e) != 0; }();
  if (want_bal && L.bal && !sparse && dblocks * rblocks > dblocks) {
    BalArgs bal{};
    bal.scratch = part; bal.flags = reinterpret_cast<unsigned*>(counters + row_tiles + 2); bal.rblocks = rblocks;
    bal.slots = L.bal_slots;
    const int nb = std::min(kNumSMs, dblocks * rblocks);
    if (G == 1)
      launch_ex(k_umma_accum_bal<1>, dim3(nb), dim3(UCfg<1>::kThreads), kUSmemBytes, st, pdl, tm_p, tm_hiB, tm_loB, num, D, (int)Q,
                bf16_bank ? 0 : 1, (int)L.npad, bal, e);
    else
      launch_ex(k_umma_accum_bal<2>, dim3(nb), dim3(UCfg<2>::kThreads), kUSmemBytes, st, pdl, tm_p, tm_hiB, tm_loB, num, D, (int)Q,
                bf16_bank ? 0 : 1, (int)L.npad, bal, e);
  } else {
    // plain fused correction of a dense, unsplit, unchunked pass: through the tile scratch + k_umma_untile
    static const bool want_untile = [] { const char* e = getenv("SDN_UMMA_UNTILE"); return !(e && atoi(e) == 0); }();
    // partial-sums calls through the scratch too (SDN_UMMA_UNTILE_PARTIAL=1): measured neutral (N = 375: phase B 16.7 ->
    // 14.7 us, one more launch, step 38.9 us both ways), so the N-sharded chain stays as it was validated on 8 GPUs
    static const bool untile_partial = [] { const char* e = getenv("SDN_UMMA_UNTILE_PARTIAL"); return e && atoi(e) != 0; }();
    const bool tileable = want_untile && L.bal && !sparse && !al.chunked && nsplit == 1;
    if (tileable && epi && e.x0 && z && (e.z || e.zpart) && !e.z_only)
      al.e.tile_out = part;                       // fused correction
    else if (tileable && !epi && num && untile_partial)
      al.e.tile_out = part;                       // partial sums (three-call sequence, N-sharded banks): k_umma_untile copies them out
    if (G == 1) launch_accum<1>(al, st); else launch_accum<2>(al, st);
  }
  g_prof.end(pid, st);
  SDN_LAUNCHED();
  if (al.e.tile_out) {
    pid = g_prof.begin("k_umma_untile", st);
    launch_ex(k_umma_untile, dim3((unsigned)cdiv(Q * D / 4, 256)), dim3(256), 0, st, pdl, (const float*)part, epi ? e.x0 : nullptr, num,
              epi ? e.neg_out : nullptr, (const float*)z, e.eps, (int)Q, D, G, e.scale, epi ? e.mean_out : nullptr, e.inv_qd);
    g_prof.end(pid, st);
    SDN_LAUNCHED();
  }
  if (nsplit > 1) {
    pid = g_prof.begin("k_umma_splitsum", st);
    k_umma_splitsum<<<(unsigned)cdiv(Q * D, 1024), 256, 0, st>>>(part, nsplit, Q * D, Q * D, (int)Q, num,
                                                                  sparse ? lists.count : nullptr, sparse ? lists.dense : nullptr);
    g_prof.end(pid, st);
    SDN_LAUNCHED();
  }
  return SDN_OK;
}

}  // namespace sdn
