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
#include <mutex>

#include "sdn_internal.h"

namespace sdn {

constexpr int kUK = 64;          // bf16 elements per K block = one 128-byte swizzle row
constexpr int kUStack = 128;     // stacked query rows: [0,64) hi parts, [64,128) lo parts
constexpr int kUQ = 64;          // max query rows per launch
constexpr int kUBankTile = 128;  // bank rows per phase-A tile (MMA N)
constexpr int kUDBlock = 128;    // d per phase-B CTA (MMA M)
constexpr int kUStages = 4;
constexpr int kUThreads = 192;
constexpr uint32_t kTileBytes = 128 * 128;   // a [128 rows][64 bf16] tile = 16 KiB
constexpr uint32_t kStageBytes = 3 * kTileBytes;
constexpr size_t kUSmemBytes = kUStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t u_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void u_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(u_smem(bar)), "r"(count));
}
__device__ __forceinline__ void u_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(u_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void u_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(u_smem(bar)), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();   // a broken pipeline must not hang the GPU
  }
}
__device__ __forceinline__ void u_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(u_smem(dst)), "l"(map), "r"(u_smem(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void u_prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void u_tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(u_smem(smem_dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void u_tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void u_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void u_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void u_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrive when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void u_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u_smem(bar)) : "memory");
}
__device__ __forceinline__ void u_tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor (sm_100 format): 128-byte swizzle, version 1.
//   K-major  tile [rows][64 bf16]: 8-row groups 1024 B apart (SBO); LBO unused (1).
//   MN-major tile [k rows][64 bf16] x (MN blocks `lbo_bytes` apart): 8-row k groups 1024 B apart (SBO).
__device__ __forceinline__ uint64_t u_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t u_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct USmem {
  uint8_t* tiles;       // [stages][3][16 KiB], 1024-byte aligned
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
  s.full = reinterpret_cast<uint64_t*>(s.tiles + (size_t)kUStages * kStageBytes);
  s.empty = s.full + kUStages;
  s.acc_full = s.empty + kUStages;
  s.acc_empty = s.acc_full + 2;
  s.tmem_base = reinterpret_cast<uint32_t*>(s.acc_empty + 2);
  return s;
}

// ------------------------------------------------------------------------------------------ query planes
__global__ void __launch_bounds__(256)
k_umma_xprep(const float* __restrict__ xq, int Q, int64_t D, __nv_bfloat16* __restrict__ planes) {
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
  *reinterpret_cast<uint2*>(planes + (int64_t)(kUQ + q) * D + j) = *reinterpret_cast<const uint2*>(l);
}

// Same, plus ||x_q||^2 partials (one per 1024-element chunk, summed by k_umma_weights) so that the batched
// conditioning call needs no separate query-prepare launch.  grid (D/1024, 64).
__global__ void __launch_bounds__(256)
k_umma_qprep(const float* __restrict__ xq, int Q, int64_t D, __nv_bfloat16* __restrict__ planes,
             float* __restrict__ xsq_part, float* __restrict__ zero_word) {
  __shared__ float red[33];
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
    *reinterpret_cast<uint2*>(planes + (int64_t)(kUQ + q) * D + j) = *reinterpret_cast<const uint2*>(l);
  }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) xsq_part[(int64_t)blockIdx.x * kUQ + q] = ss;
}

// ------------------------------------------------------------------------------------------ phase A
// grid (row tiles, k splits).  S_T [ksplit][Npad][128] fp32: S_T[s][i][r] = sum over split s of X[r][d] * bank[i][d].
//
// The tensor core adds into its fp32 accumulator with truncation, and near a negative the distance is the
// small difference of large dot products, so long accumulation chains cost accuracy (measured: 43 K-blocks in
// one chain -> 3.6e-4 on the weights).  The MMA warp therefore alternates between two TMEM accumulators every
// kUChunk K-blocks and the epilogue warps drain the finished one into fp32 registers (round-to-nearest adds).
constexpr int kUChunk = 4;

__global__ void __launch_bounds__(kUThreads, 1)
k_umma_dots(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hi,
            const __grid_constant__ CUtensorMap tm_lo, float* __restrict__ S_T, int64_t split_stride,
            int kblocks_total, int ksplit, int use_lo) {
  extern __shared__ unsigned char smem_raw[];
  const USmem sm = u_carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kUBankTile;
  const int kb0 = (int)((int64_t)blockIdx.y * kblocks_total / ksplit);
  const int kb1 = (int)((int64_t)(blockIdx.y + 1) * kblocks_total / ksplit);
  const int nkb = kb1 - kb0;
  const int nchunks = (nkb + kUChunk - 1) / kUChunk;

  auto load_stage = [&](int i) {
    const int s = i % kUStages;
    uint8_t* st = sm.tiles + (size_t)s * kStageBytes;
    u_mbar_expect_tx(&sm.full[s], use_lo ? kStageBytes : 2 * kTileBytes);
    const int kc = (kb0 + i) * kUK;
    u_tma_2d(st, &tm_x, kc, 0, &sm.full[s]);
    u_tma_2d(st + kTileBytes, &tm_hi, kc, row0, &sm.full[s]);
    if (use_lo) u_tma_2d(st + 2 * kTileBytes, &tm_lo, kc, row0, &sm.full[s]);
  };
  const int npre = min(nkb, kUStages);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kUStages; ++s) { u_mbar_init(&sm.full[s], 1); u_mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { u_mbar_init(&sm.acc_full[b], 1); u_mbar_init(&sm.acc_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the first stages need nothing but this thread's own barriers: their HBM latency overlaps the TMEM
    // allocation and the start-up barrier (the kernel is fill/drain bound at N ~ 3000)
    u_prefetch_map(&tm_x); u_prefetch_map(&tm_hi); u_prefetch_map(&tm_lo);
    for (int i = 0; i < npre; ++i) load_stage(i);
  }
  if (warp == 1) u_tmem_alloc(sm.tmem_base, 2 * kUBankTile);
  u_fence_before();
  __syncthreads();
  u_fence_after();
  const uint32_t tmem = *sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = npre; i < nkb; ++i) {
        u_mbar_wait(&sm.empty[i % kUStages], (uint32_t)(((i / kUStages) + 1) & 1));
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
        const uint32_t acc = tmem + (uint32_t)(buf * kUBankTile);
        const int i1 = min(nkb, (c + 1) * kUChunk);
        for (int i = c * kUChunk; i < i1; ++i) {
          const int s = i % kUStages;
          u_mbar_wait(&sm.full[s], (uint32_t)((i / kUStages) & 1));
          u_fence_after();
          const uint32_t base = u_smem(sm.tiles + (size_t)s * kStageBytes);
#pragma unroll
          for (int kk = 0; kk < kUK / 16; ++kk) {
            const uint64_t a = u_desc(base + kk * 32, 16, 1024);
            const uint64_t bh = u_desc(base + kTileBytes + kk * 32, 16, 1024);
            const uint64_t bl = u_desc(base + 2 * kTileBytes + kk * 32, 16, 1024);
            u_mma(acc, a, bh, idesc, (i > c * kUChunk || kk > 0) ? 1u : 0u);
            if (use_lo) u_mma(acc, a, bl, idesc, 1u);
          }
          u_commit(&sm.empty[s]);
        }
        u_commit(&sm.acc_full[buf]);
      }
    }
  } else {
    // epilogue: warp w may touch TMEM lanes [32*(w%4), +32); lane index = stacked query row
    const int lq = warp & 3;
    float sum[kUBankTile];
#pragma unroll
    for (int j = 0; j < kUBankTile; ++j) sum[j] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      u_mbar_wait(&sm.acc_full[buf], (uint32_t)((c >> 1) & 1));
      u_fence_after();
      const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * kUBankTile);
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
#pragma unroll
    for (int j = 0; j < kUBankTile; ++j) dst[(int64_t)j * kUStack] = sum[j];
  }
  u_fence_before();
  __syncthreads();
  if (warp == 1) {
    u_fence_after();
    u_tmem_dealloc(tmem, 2 * kUBankTile);
  }
}

constexpr int kListCap = 32;     // significant rows per query the listed accumulate handles

struct SigLists {
  int* flags;        // [Npad/64]      row block holds a significant weight
  int* count;        // [64]           significant rows per query (may exceed kListCap)
  int* dense;        // [1]            set by k_umma_zreduce when some query's weights are flat (z / kmax > cap):
                     //                its list must overflow, so nobody builds lists or flags and phase B is dense
  int* rows;         // [64][kListCap] their bank row indices (unsorted, first kListCap arrivals)
  float* ks;         // [64][kListCap] their weights
};

// zeroes the lists (launched as part of k_umma_weights' grid: block 0 does it before anyone appends)
__device__ __forceinline__ void siglist_clear(const SigLists& L, int nflags) {
  for (int i = threadIdx.x; i < nflags; i += blockDim.x) L.flags[i] = 0;
  if (threadIdx.x < kUQ) L.count[threadIdx.x] = 0;
  if (threadIdx.x == 0) *L.dense = 0;
}

// ------------------------------------------------------------------------------------------ weights
// One thread per (bank row, query row): sum the K-split partials, k = exp(-dist / 2 sigma^2), write the weight
// planes P [Npad][128] bf16 (stacked query index contiguous: hi parts in columns [0,64), lo parts in [64,128)),
// which phase B reads as an MN-major B operand.  Everything is coalesced; block = 4 bank rows x 64 query rows.
constexpr int kWRows = 4;

__global__ void __launch_bounds__(kWRows * kUQ)
k_umma_weights(const float* __restrict__ S_T, int64_t split_stride, int ksplit, const float* __restrict__ sqnorm,
               const float* __restrict__ xsq, const float* __restrict__ xsq_part, int xsq_nparts, int N, int Q,
               float inv2s2, int power, float alpha, __nv_bfloat16* __restrict__ P, float* __restrict__ zpart,
               float* __restrict__ k_out, SigLists lists, int nflags) {
  __shared__ float zs[kWRows][kUQ];
  if (blockIdx.x == 0 && lists.flags) siglist_clear(lists, nflags);
  const int q = threadIdx.x & (kUQ - 1);
  const int rsub = threadIdx.x >> 6;
  const int i = blockIdx.x * kWRows + rsub;
  const float* s = S_T + (int64_t)i * kUStack;
  float dot = 0.f;
#pragma unroll 4
  for (int sp = 0; sp < ksplit; ++sp) {
    const float* p = s + (int64_t)sp * split_stride;
    dot += p[q] + p[kUQ + q];
  }
  float k = 0.f;
  if (i < N && q < Q) {
    float xs = 0.f;
    if (xsq) {
      xs = xsq[q];
    } else {
      for (int c = 0; c < xsq_nparts; ++c) xs += xsq_part[(int64_t)c * kUQ + q];
    }
    k = expf(-dist_from_dot(xs, sqnorm[i], dot, alpha, power) * inv2s2);
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(k);
  const __nv_bfloat16 l = __float2bfloat16_rn(k - __bfloat162float(h));
  P[(int64_t)i * kUStack + q] = h;
  P[(int64_t)i * kUStack + kUQ + q] = l;
  if (k_out && i < N && q < Q) k_out[(int64_t)q * N + i] = k;
  // z: per-block partials; k_umma_zreduce sums them in a fixed order (deterministic, no same-address atomics,
  // and no serial tail: a "last block reduces" variant spent 16 us in one SM)
  zs[rsub][q] = k;
  __syncthreads();
  if (threadIdx.x < kUQ) {
    float t = 0.f, m = 0.f;
#pragma unroll
    for (int r = 0; r < kWRows; ++r) {
      t += zs[r][threadIdx.x];
      m = fmaxf(m, zs[r][threadIdx.x]);
    }
    zpart[(int64_t)blockIdx.x * kUStack + threadIdx.x] = t;          // [block][0..63]  sums
    zpart[(int64_t)blockIdx.x * kUStack + kUQ + threadIdx.x] = m;    // [block][64..127] maxima
  }
}

// z[q] = sum_b zpart[b][q], kmax[q] = max_b zpart[b][64+q]; one block per query row.
__global__ void __launch_bounds__(256)
k_umma_zreduce(const float* __restrict__ zpart, int nblocks, float* __restrict__ z, float* __restrict__ kmax,
               int* __restrict__ dense_flag) {
  __shared__ float red[33];
  __shared__ float mx[8];
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
struct AccumEpi {
  const float* z;       // null: no fused correction
  float eps, scale, gate_thr; int flags;
  float* x0; float* neg_out; float* denom_out; int32_t* gate_out; float* mean_out; float inv_qd;
};

// grid (D / 128, n splits).  num[q][d] (+)= sum_i P[q][i] * (hi+lo)[i][d] over this split's bank rows.
template <bool CHUNKED>
__global__ void __launch_bounds__(kUThreads, 1)
k_umma_accum(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_hi,
             const __grid_constant__ CUtensorMap tm_lo, float* __restrict__ num, int64_t D, int Q,
             int rblocks_total, int nsplit, int use_atomic, int use_lo, const int* rowflags,
             const int* __restrict__ list_count, const int* __restrict__ dense_flag, const AccumEpi epi) {
  if (dense_flag && __ldg(dense_flag)) {
    rowflags = nullptr;                                          // flat regime: no flags were built
  } else if (list_count && siglist_all_short(list_count, Q)) {
    return;                                                      // k_umma_listed_accum produced the result
  }
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint16_t act[kMaxActive];
  __shared__ int nact_s;
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
  auto load_stage = [&](int it, int i, int t) {
    const int d0 = ((int)blockIdx.x + t * (int)gridDim.x) * kUDBlock;
    const int s = it % kUStages;
    uint8_t* st = sm.tiles + (size_t)s * kStageBytes;
    u_mbar_expect_tx(&sm.full[s], use_lo ? kStageBytes : 2 * kTileBytes);
    const int rc = (rb0 + i) * kUK;                   // first bank row of this block
    // P tile: two boxes of [64 rows][64 stacked q] (hi parts, lo parts), 8 KiB apart
    u_tma_2d(st, &tm_p, 0, rc, &sm.full[s]);
    u_tma_2d(st + 8192, &tm_p, 64, rc, &sm.full[s]);
    // bank^T tiles: two boxes of [64 rows][64 d] per plane, 8 KiB apart
    u_tma_2d(st + kTileBytes, &tm_hi, d0, rc, &sm.full[s]);
    u_tma_2d(st + kTileBytes + 8192, &tm_hi, d0 + 64, rc, &sm.full[s]);
    if (use_lo) {
      u_tma_2d(st + 2 * kTileBytes, &tm_lo, d0, rc, &sm.full[s]);
      u_tma_2d(st + 2 * kTileBytes + 8192, &tm_lo, d0 + 64, rc, &sm.full[s]);
    }
  };
  // dense pass: the first stages depend on nothing but this thread's own barriers, so their latency overlaps
  // the TMEM allocation and the start-up barrier
  const int npre = (rowflags == nullptr && nrb > 0) ? min(ntasks * nrb, kUStages) : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kUStages; ++s) { u_mbar_init(&sm.full[s], 1); u_mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { u_mbar_init(&sm.acc_full[b], 1); u_mbar_init(&sm.acc_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    u_prefetch_map(&tm_p); u_prefetch_map(&tm_hi); u_prefetch_map(&tm_lo);
    for (int it = 0; it < npre; ++it) load_stage(it, it % nrb, it / nrb);
  }
  if (warp == 1) u_tmem_alloc(sm.tmem_base, 2 * kUStack);
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
  const uint32_t tmem = *sm.tmem_base;
  const int nact = nact_s;
  const bool dense = rowflags == nullptr;
  const int nchunks = max(1, (nact + kBChunk - 1) / kBChunk);

  if (warp == 0) {
    if (lane == 0) {
      for (int it = npre; it < ntasks * nact; ++it) {
        const int t = it / nact, j = it - t * nact;
        if (it >= kUStages) u_mbar_wait(&sm.empty[it % kUStages], (uint32_t)(((it / kUStages) + 1) & 1));
        load_stage(it, dense ? j : (int)act[j], t);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_full = u_idesc(kUDBlock, kUStack, 1, 1);   // hi * [P_hi | P_lo]
      constexpr uint32_t id_half = u_idesc(kUDBlock, kUQ, 1, 1);       // lo * P_hi
      if constexpr (!CHUNKED) {
      for (int t = 0; t < ntasks && nact > 0; ++t) {
        const int buf = t & 1;
        if (t >= 2) {
          u_mbar_wait(&sm.acc_empty[buf], (uint32_t)(((t >> 1) + 1) & 1));
          u_fence_after();
        }
        const uint32_t acc = tmem + (uint32_t)(buf * kUStack);
        for (int i = 0; i < nact; ++i) {
          const int it = t * nact + i;
          const int s = it % kUStages;
          u_mbar_wait(&sm.full[s], (uint32_t)((it / kUStages) & 1));
          u_fence_after();
          const uint32_t base = u_smem(sm.tiles + (size_t)s * kStageBytes);
#pragma unroll
          for (int kk = 0; kk < kUK / 16; ++kk) {
            const uint64_t b = u_desc(base + kk * 2048, 8192, 1024);                    // P^T, MN-major
            const uint64_t ah = u_desc(base + kTileBytes + kk * 2048, 8192, 1024);      // bank^T, MN-major
            const uint64_t al = u_desc(base + 2 * kTileBytes + kk * 2048, 8192, 1024);
            u_mma(acc, ah, b, id_full, (i > 0 || kk > 0) ? 1u : 0u);
            if (use_lo) u_mma(acc, al, b, id_half, 1u);
          }
          u_commit(&sm.empty[s]);
        }
        u_commit(&sm.acc_full[buf]);
      }
      } else {
      // unit u = (d-block task, chunk of kBChunk row blocks); consecutive units alternate TMEM accumulators so that
      // no fp32 accumulation chain in the tensor core is longer than kBChunk * 8 MMAs (its adds truncate)
      for (int u = 0; nact > 0 && u < ntasks * nchunks; ++u) {
        const int t = u / nchunks, c = u - t * nchunks;
        const int buf = u & 1;
        if (u >= 2) {
          u_mbar_wait(&sm.acc_empty[buf], (uint32_t)(((u >> 1) + 1) & 1));
          u_fence_after();
        }
        const uint32_t acc = tmem + (uint32_t)(buf * kUStack);
        const int j0 = c * kBChunk, j1 = min(nact, j0 + kBChunk);
        for (int i = j0; i < j1; ++i) {
          const int it = t * nact + i;
          const int s = it % kUStages;
          u_mbar_wait(&sm.full[s], (uint32_t)((it / kUStages) & 1));
          u_fence_after();
          const uint32_t base = u_smem(sm.tiles + (size_t)s * kStageBytes);
#pragma unroll
          for (int kk = 0; kk < kUK / 16; ++kk) {
            const uint64_t b = u_desc(base + kk * 2048, 8192, 1024);                    // P^T, MN-major
            const uint64_t ah = u_desc(base + kTileBytes + kk * 2048, 8192, 1024);      // bank^T, MN-major
            const uint64_t al = u_desc(base + 2 * kTileBytes + kk * 2048, 8192, 1024);
            u_mma(acc, ah, b, id_full, (i > j0 || kk > 0) ? 1u : 0u);
            if (use_lo) u_mma(acc, al, b, id_half, 1u);
          }
          u_commit(&sm.empty[s]);
        }
        u_commit(&sm.acc_full[buf]);
      }
      }
    }
  } else {
    const int lq = warp & 3;
    float msum = 0.f;
    if constexpr (!CHUNKED) {
    // epilogue: TMEM lane = d within the block, column = stacked query row
#pragma unroll 1
    for (int t = 0; t < ntasks; ++t) {
    const int buf = t & 1;
    if (nact > 0) {
      u_mbar_wait(&sm.acc_full[buf], (uint32_t)((t >> 1) & 1));
      u_fence_after();
    }
    const int64_t d = (int64_t)((int)blockIdx.x + t * (int)gridDim.x) * kUDBlock + lq * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * kUStack);
#pragma unroll 1
    for (int c = 0; c < kUQ / 32; ++c) {
      float a[32], b[32];
      if (nact > 0) {
        u_tmem_ld32(tl + (uint32_t)(c * 32), a);            // hi*P_hi + lo*P_hi, queries [32c, 32c+32)
        u_tmem_ld32(tl + (uint32_t)(kUQ + c * 32), b);      // hi*P_lo
      } else {                                              // no row block of this split matters: the sum is zero
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = b[j] = 0.f;
      }
      if (epi.z) {
        // all loads of the chunk first: the stores below may alias them as far as the compiler knows, and one
        // load -> store round trip per query row costs ~25 us per launch
        float xv[32], dn[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int q = min(c * 32 + j, Q - 1);
          dn[j] = __ldg(epi.z + q) + epi.eps;
          xv[j] = epi.x0 ? __ldcg(epi.x0 + (int64_t)q * D + d) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int q = c * 32 + j;
          if (q < Q) {
            const float v = a[j] + b[j];
            const int64_t o = (int64_t)q * D + d;
            const float n = v / dn[j];
            if (num) num[o] = v;
            if (epi.neg_out) epi.neg_out[o] = n;
            if (epi.x0) epi.x0[o] = fmaf(-epi.scale, n, xv[j]);
            msum += fminf(fmaxf(n, -1e10f), 1e10f);
            if (blockIdx.x == 0 && lq == 0 && lane == 0) {
              if (epi.denom_out) epi.denom_out[q] = dn[j];
              if (epi.gate_out) epi.gate_out[q] = (!(epi.flags & SDN_EPI_GATE) || dn[j] > epi.gate_thr) ? 1 : 0;
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int q = c * 32 + j;
          if (q < Q) {
            float* o = num + (int64_t)q * D + d;
            if (use_atomic) atomicAdd(o, a[j] + b[j]); else *o = a[j] + b[j];
          }
        }
      }
    }
    u_fence_before();
    __syncwarp();
    if (lane == 0 && nact > 0)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(u_smem(&sm.acc_empty[buf])) : "memory");
    }   // tasks
    } else {
    // epilogue: TMEM lane = d within the block, column = stacked query row; running fp32 sums over the chunks
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
        const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * kUStack);
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
      if (epi.z) {
#pragma unroll
        for (int cq = 0; cq < kUQ / 32; ++cq) {
          float xv[32], dn[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = min(cq * 32 + j, Q - 1);
            dn[j] = __ldg(epi.z + q) + epi.eps;
            xv[j] = epi.x0 ? __ldcg(epi.x0 + (int64_t)q * D + d) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = cq * 32 + j;
            if (q < Q) {
              const float v = sum[cq * 32 + j];
              const int64_t o = (int64_t)q * D + d;
              const float n = v / dn[j];
              if (num) num[o] = v;
              if (epi.neg_out) epi.neg_out[o] = n;
              if (epi.x0) epi.x0[o] = fmaf(-epi.scale, n, xv[j]);
              msum += fminf(fmaxf(n, -1e10f), 1e10f);
              if (blockIdx.x == 0 && lq == 0 && lane == 0) {
                if (epi.denom_out) epi.denom_out[q] = dn[j];
                if (epi.gate_out) epi.gate_out[q] = (!(epi.flags & SDN_EPI_GATE) || dn[j] > epi.gate_thr) ? 1 : 0;
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < kUQ; ++j) {
          if (j < Q) {
            float* o = num + (int64_t)j * D + d;
            if (use_atomic) atomicAdd(o, sum[j]); else *o = sum[j];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kUQ; ++j) sum[j] = 0.f;
    }   // units
    }
    if (epi.z && epi.mean_out) {
      msum = warp_sum(msum);
      if (lane == 0) atomicAdd(epi.mean_out, msum * epi.inv_qd);
    }
  }
  u_fence_before();
  __syncthreads();
  if (warp == 1) {
    u_fence_after();
    u_tmem_dealloc(tmem, 2 * kUStack);
  }
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

int umma_ksplit(int64_t npad, int64_t D) {
  const int row_tiles = (int)(npad / kUBankTile);
  const int kblocks = (int)(D / kUK);
  // one task per SM if possible, at least 8 K-blocks per task (fill/drain), at most 64 partial buffers
  int ksplit = std::max(1, std::min(kblocks / 8, kNumSMs / row_tiles));   // never a second, nearly empty wave
  return std::min(ksplit, 64);
}

struct UmmaLayout {
  int64_t npad;        // bank rows padded to 128
  int ksplit;
  size_t off_x, off_s, off_p, off_z, off_f, off_q, total;
};
UmmaLayout umma_layout(int64_t N, int64_t D) {
  UmmaLayout L;
  L.npad = cdiv(N, 128) * 128;
  L.ksplit = umma_ksplit(L.npad, D);
  size_t o = 0;
  L.off_x = o; o += (size_t)kUStack * D * 2;                 // X planes
  o = (o + 255) / 256 * 256;
  L.off_s = o; o += (size_t)L.ksplit * L.npad * kUStack * 4; // S^T, one partial per K split
  o = (o + 255) / 256 * 256;
  L.off_p = o; o += (size_t)kUStack * L.npad * 2;            // P planes
  o = (o + 255) / 256 * 256;
  L.off_z = o; o += (size_t)(L.npad / kWRows) * kUStack * 4 + 256;   // per-block z sums | maxima
  o = (o + 255) / 256 * 256;
  L.off_f = o; o += (size_t)(L.npad / kUK) * 4 + kUQ * 4 + kUQ * 4 + 16 + (size_t)kUQ * kListCap * 8 + 256;
                                                       // row-block flags | kmax[64] | count[64] | rows | ks
  o = (o + 255) / 256 * 256;
  L.off_q = o; o += (size_t)cdiv(D, 1024) * kUQ * 4;          // ||x||^2 partials of the fused query prepare
  L.total = (o + 255) / 256 * 256;
  return L;
}
}  // namespace

bool umma_supported(int64_t Q, int64_t N, int64_t D, const void* planes) {
  return planes != nullptr && Q >= 1 && N >= 1 && D >= kUDBlock && D % kUDBlock == 0 &&
         N < (1ll << 31) && D < (1ll << 31);
}

size_t umma_workspace_bytes(int64_t Q, int64_t N, int64_t D) {
  if (Q < 1 || D % kUDBlock) return 0;
  return umma_layout(N, D).total;
}

static int umma_partial_64(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                           const float* xsq, int64_t Q, float inv2s2, int power, float alpha, float* num, float* z,
                           float* k_out, void* ws, size_t ws_bytes, cudaStream_t st, const AccumEpi* epi,
                           float* zero_word, bool bf16_bank);

// More than 64 query rows: one two-phase pass over the bank per group of 64 (the TMEM accumulator of phase B
// holds 128 d x 128 stacked query columns).
int umma_partial(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                 const float* xsq, int64_t Q, float inv2s2, int power, float alpha, float* num, float* z,
                 float* k_out, void* ws, size_t ws_bytes, cudaStream_t st, bool bf16_bank) {
  if (!umma_supported(Q, N, D, planes)) return SDN_E_UNSUPPORTED;
  for (int64_t q0 = 0; q0 < Q; q0 += kUQ) {
    const int64_t qn = std::min<int64_t>(kUQ, Q - q0);
    const int rc = umma_partial_64(planes, sqnorm, N, D, xq + q0 * D, xsq ? xsq + q0 : nullptr, qn, inv2s2, power, alpha,
                                   num ? num + q0 * D : nullptr, z + q0, k_out ? k_out + q0 * N : nullptr, ws,
                                   ws_bytes, st, nullptr, nullptr, bf16_bank);
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
  for (int64_t q0 = 0; q0 < Q; q0 += kUQ) {
    const int64_t qn = std::min<int64_t>(kUQ, Q - q0);
    AccumEpi epi{};
    epi.z = z + q0; epi.eps = eps; epi.scale = scale; epi.gate_thr = gate_thr; epi.flags = flags;
    epi.x0 = x0_inout + q0 * D; epi.neg_out = neg_out ? neg_out + q0 * D : nullptr;
    epi.denom_out = denom_out ? denom_out + q0 : nullptr; epi.gate_out = gate_out ? gate_out + q0 : nullptr;
    epi.mean_out = mean_out; epi.inv_qd = 1.f / (float)(Q * D);
    const int rc = umma_partial_64(planes, sqnorm, N, D, x0_inout + q0 * D, nullptr, qn, inv2s2, power, alpha,
                                   num_out ? num_out + q0 * D : nullptr, z + q0, k_out ? k_out + q0 * N : nullptr, ws,
                                   ws_bytes, st, &epi, q0 == 0 ? mean_out : nullptr, false);
    if (rc) return rc;
  }
  return SDN_OK;
}

static int umma_partial_64(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                           const float* xsq, int64_t Q, float inv2s2, int power, float alpha, float* num, float* z,
                           float* k_out, void* ws, size_t ws_bytes, cudaStream_t st, const AccumEpi* epi,
                           float* zero_word, bool bf16_bank) {
  const UmmaLayout L = umma_layout(N, D);
  if (!ws || ws_bytes < L.total) return SDN_E_WORKSPACE;
  if (!load_encode()) return SDN_E_DEVICE;
  static std::atomic<bool> configured[kMaxDevices];
  const int dev = device_slot();
  if (!configured[dev].load(std::memory_order_acquire)) {
    SDN_CUDA_OK(cudaFuncSetAttribute(k_umma_dots, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUSmemBytes));
    SDN_CUDA_OK(cudaFuncSetAttribute(k_umma_accum<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUSmemBytes));
    SDN_CUDA_OK(cudaFuncSetAttribute(k_umma_accum<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUSmemBytes));
    configured[dev].store(true, std::memory_order_release);
  }
  char* w = static_cast<char*>(ws);
  __nv_bfloat16* xpl = reinterpret_cast<__nv_bfloat16*>(w + L.off_x);
  float* S_T = reinterpret_cast<float*>(w + L.off_s);
  __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(w + L.off_p);
  const __nv_bfloat16* hi = static_cast<const __nv_bfloat16*>(planes);
  const __nv_bfloat16* lo = hi + N * D;

  // tensor maps depend only on (planes, workspace, N, D): re-encode when one of those changes
  struct MapCache { const void* planes; const void* ws; int64_t N, D; CUtensorMap m[6]; bool valid; };
  static MapCache cache{nullptr, nullptr, 0, 0, {}, false};
  static std::mutex cache_mu;
  CUtensorMap tm_x, tm_hiA, tm_loA, tm_p, tm_hiB, tm_loB;
  {
    std::lock_guard<std::mutex> lk(cache_mu);
    if (!(cache.valid && cache.planes == planes && cache.ws == ws && cache.N == N && cache.D == D)) {
      int rc;
      cache.valid = false;
      if ((rc = make_map(&cache.m[0], xpl, kUStack, D, kUStack, kUK))) return rc;
      if ((rc = make_map(&cache.m[1], hi, N, D, kUBankTile, kUK))) return rc;
      if ((rc = make_map(&cache.m[2], lo, N, D, kUBankTile, kUK))) return rc;
      if ((rc = make_map(&cache.m[3], P, L.npad, kUStack, kUK, 64))) return rc;
      if ((rc = make_map(&cache.m[4], hi, N, D, kUK, 64))) return rc;
      if ((rc = make_map(&cache.m[5], lo, N, D, kUK, 64))) return rc;
      cache.planes = planes; cache.ws = ws; cache.N = N; cache.D = D; cache.valid = true;
    }
    tm_x = cache.m[0]; tm_hiA = cache.m[1]; tm_loA = cache.m[2]; tm_p = cache.m[3]; tm_hiB = cache.m[4]; tm_loB = cache.m[5];
  }

  // query planes
  float* zpart = reinterpret_cast<float*>(w + L.off_z);
  float* xsq_part = reinterpret_cast<float*>(w + L.off_q);
  const int xsq_nparts = (int)cdiv(D, 1024);
  int pid = g_prof.begin(xsq ? "k_umma_xprep" : "k_umma_qprep", st);
  if (xsq) {
    k_umma_xprep<<<dim3((unsigned)cdiv(D, 1024), kUQ), 256, 0, st>>>(xq, (int)Q, D, xpl);
  } else {
    k_umma_qprep<<<dim3((unsigned)xsq_nparts, kUQ), 256, 0, st>>>(xq, (int)Q, D, xpl, xsq_part, zero_word);
  }
  g_prof.end(pid, st);
  SDN_LAUNCHED();

  // phase A: split K (= D) so that roughly every SM gets one task
  const int row_tiles = (int)(L.npad / kUBankTile);
  const int kblocks = (int)(D / kUK);
  const int64_t split_stride = L.npad * kUStack;
  pid = g_prof.begin("k_umma_dots", st);
  k_umma_dots<<<dim3(row_tiles, L.ksplit), kUThreads, kUSmemBytes, st>>>(tm_x, tm_hiA, tm_loA, S_T, split_stride,
                                                                        kblocks, L.ksplit, bf16_bank ? 0 : 1);
  g_prof.end(pid, st);
  SDN_LAUNCHED();

  // weights
  SigLists lists_w{};
  if (g_skip_negligible.load(std::memory_order_relaxed) && (L.npad / kUK) <= kMaxActive && (num || epi)) {
    lists_w.flags = reinterpret_cast<int*>(w + L.off_f);
    lists_w.count = reinterpret_cast<int*>(reinterpret_cast<float*>(lists_w.flags + L.npad / kUK) + kUQ);
    lists_w.dense = lists_w.count + kUQ;
  }
  pid = g_prof.begin("k_umma_weights", st);
  k_umma_weights<<<(unsigned)(L.npad / kWRows), kWRows * kUQ, 0, st>>>(S_T, split_stride, L.ksplit, sqnorm, xsq, xsq_part,
                                                                      xsq_nparts, (int)N, (int)Q, inv2s2, power, alpha,
                                                                      P, zpart, k_out, lists_w, (int)(L.npad / kUK));
  SDN_LAUNCHED();
  const int nflags = (int)(L.npad / kUK);
  SigLists lists{};
  lists.flags = reinterpret_cast<int*>(w + L.off_f);
  float* kmax = reinterpret_cast<float*>(lists.flags + nflags);
  lists.count = reinterpret_cast<int*>(kmax + kUQ);
  lists.dense = lists.count + kUQ;
  lists.rows = lists.dense + 4;
  lists.ks = reinterpret_cast<float*>(lists.rows + kUQ * kListCap);
  k_umma_zreduce<<<(unsigned)Q, 256, 0, st>>>(zpart, (int)(L.npad / kWRows), z, kmax, lists_w.flags ? lists.dense : nullptr);
  SDN_LAUNCHED();
  const bool sparse = g_skip_negligible.load(std::memory_order_relaxed) && nflags <= kMaxActive && (num || epi);
  if (sparse) {
    k_umma_siglist<<<(unsigned)(L.npad / kWRows), kWRows * kUQ, 0, st>>>(P, kmax, (int)N, (int)Q, lists);
    SDN_LAUNCHED();
  }
  g_prof.end(pid, st);
  if (!num && !epi) return SDN_OK;   // z only (empirical_beta): no phase B

  // phase B: one CTA per 128 d, bank rows split when that leaves SMs idle
  const int dblocks = (int)(D / kUDBlock);
  const int rblocks = (int)(L.npad / kUK);
  int nsplit = std::max(1, std::min(rblocks / 8, kNumSMs / dblocks));
  AccumEpi e{};
  if (epi) {
    if (nsplit != 1) return SDN_E_UNSUPPORTED;
    e = *epi;
  }
  if (nsplit > 1 && num) SDN_CUDA_OK(cudaMemsetAsync(num, 0, sizeof(float) * Q * D, st));
  if (sparse) {
    pid = g_prof.begin("k_umma_listed_accum", st);
    k_umma_listed_accum<<<dim3((unsigned)cdiv(D, 1024), (unsigned)Q), 256, 0, st>>>(
        hi, bf16_bank ? nullptr : lo, D, (int)Q, lists, num, epi ? e.z : nullptr, e.eps, e.scale, e.gate_thr, e.flags, e.x0,
        e.neg_out, e.denom_out, e.gate_out, e.mean_out, e.inv_qd);
    g_prof.end(pid, st);
    SDN_LAUNCHED();
  }
  pid = g_prof.begin("k_umma_accum", st);
  const int gridx = nsplit == 1 ? std::min(dblocks, kNumSMs) : dblocks;
  // chains longer than kBChunk row blocks are split over the two TMEM accumulators (fp32 register drain)
  if ((rblocks + nsplit - 1) / nsplit > kBChunk) {
    k_umma_accum<true><<<dim3(gridx, nsplit), kUThreads, kUSmemBytes, st>>>(tm_p, tm_hiB, tm_loB, num, D, (int)Q, rblocks,
                                                                     nsplit, nsplit > 1 ? 1 : 0, bf16_bank ? 0 : 1,
                                                                     sparse ? lists.flags : nullptr,
                                                                     sparse ? lists.count : nullptr,
                                                                     sparse ? lists.dense : nullptr, e);
  } else {
    k_umma_accum<false><<<dim3(gridx, nsplit), kUThreads, kUSmemBytes, st>>>(tm_p, tm_hiB, tm_loB, num, D, (int)Q, rblocks,
                                                                     nsplit, nsplit > 1 ? 1 : 0, bf16_bank ? 0 : 1,
                                                                     sparse ? lists.flags : nullptr,
                                                                     sparse ? lists.count : nullptr,
                                                                     sparse ? lists.dense : nullptr, e);
  }
  g_prof.end(pid, st);
  SDN_LAUNCHED();
  return SDN_OK;
}

}  // namespace sdn
