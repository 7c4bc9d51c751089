// Generic CUDA-core two-phase path: any Q, N, D (D % 4 == 0).
//
//   phase A  k_dots     S[q,i] = xq_q . n_i                       (bank pass 1)
//            k_weights  k_qi = exp(-dist_qi / 2 sigma^2), z_q = sum_i k_qi
//   phase B  k_accum    num[q,:] = sum_i k_qi n_i                 (bank pass 2)
//
// Replaces the [Q,N,D+1] broadcast of repellency_methods_fast.py:249-250.  This family is the
// always-correct path: the one-pass cluster kernel (sdn_stream.cu) and the tcgen05 kernels
// (sdn_umma.cu) take over for the shapes they are built for.
#include <algorithm>

#include "sdn_internal.h"

namespace sdn {

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

// One warp owns R bank rows and QT query rows; lanes stride D in float4 steps.
template <int QT, int R>
__global__ void __launch_bounds__(256)
k_dots(const float* __restrict__ bank, int64_t N, int64_t D, const float* __restrict__ x, int64_t Q,
       float* __restrict__ S) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * R;
  const int64_t q0 = (int64_t)blockIdx.y * QT;
  if (row0 >= N) return;
  const float* brow[R];
  const float* xrow[QT];
#pragma unroll
  for (int r = 0; r < R; ++r) brow[r] = bank + min(row0 + r, N - 1) * D;
#pragma unroll
  for (int t = 0; t < QT; ++t) xrow[t] = x + min(q0 + t, Q - 1) * D;
  float acc[R][QT];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int t = 0; t < QT; ++t) acc[r][t] = 0.f;

  for (int64_t j = lane * 4; j < D; j += 128) {
    float4 b[R];
#pragma unroll
    for (int r = 0; r < R; ++r) b[r] = ld_stream4(brow[r] + j);
#pragma unroll
    for (int t = 0; t < QT; ++t) {
      const float4 xv = ldg4(xrow[t] + j);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r][t] = dot4(b[r], xv, acc[r][t]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int t = 0; t < QT; ++t) {
      const float v = warp_sum(acc[r][t]);
      if (lane == 0 && row0 + r < N && q0 + t < Q) S[(q0 + t) * N + row0 + r] = v;
    }
}

__global__ void __launch_bounds__(256)
k_weights(float* __restrict__ S, const float* __restrict__ sqnorm, const float* __restrict__ xsq,
          int64_t N, float inv2s2, int power, float alpha, float* __restrict__ z) {
  __shared__ float red[33];
  const int64_t q = blockIdx.x;
  const float xs = xsq[q];
  float* row = S + q * N;
  float sum = 0.f;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const float d = dist_from_dot(xs, sqnorm[i], row[i], alpha, power);
    const float k = expf(-d * inv2s2);
    row[i] = k;
    sum += k;
  }
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) z[q] = sum;
}

__global__ void __launch_bounds__(256)
k_sparse_weights(float* __restrict__ S, const float* __restrict__ sqnorm, const float* __restrict__ xsq,
                 int64_t N, float radius, float* __restrict__ wsum) {
  __shared__ float red[33];
  const int64_t q = blockIdx.x;
  const float xs = xsq[q];
  float* row = S + q * N;
  float sum = 0.f;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const float d = dist_from_dot(xs, sqnorm[i], row[i], 1.f, 1);
    // fast.py:312-326: keep d < radius, weight relu(radius / d - 1)
    const float w = (d < radius) ? fmaxf(radius / d - 1.f, 0.f) : 0.f;
    row[i] = w;
    sum += w;
  }
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) wsum[q] = sum;
}

// Thread owns 4 consecutive d and QT query rows; block covers 512 d; blockIdx.y is the N split.
constexpr int kAccRows = 128;

template <int QT>
__global__ void __launch_bounds__(128)
k_accum(const float* __restrict__ bank, int64_t N, int64_t D, const float* __restrict__ k, int64_t Q,
        float* __restrict__ num, int nsplit) {
  __shared__ float ks[QT][kAccRows];
  const int64_t d = ((int64_t)blockIdx.x * 128 + threadIdx.x) * 4;
  const int64_t q0 = (int64_t)blockIdx.z * QT;
  const int64_t chunk = (N + nsplit - 1) / nsplit;
  const int64_t i0 = (int64_t)blockIdx.y * chunk;
  const int64_t i1 = min(N, i0 + chunk);
  float4 acc[QT];
#pragma unroll
  for (int t = 0; t < QT; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int64_t ib = i0; ib < i1; ib += kAccRows) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < QT * kAccRows; idx += 128) {
      const int t = idx / kAccRows, r = idx % kAccRows;
      const int64_t i = ib + r;
      ks[t][r] = (i < i1 && q0 + t < Q) ? k[(q0 + t) * N + i] : 0.f;
    }
    __syncthreads();
    if (d < D) {
      const int cnt = (int)min((int64_t)kAccRows, i1 - ib);
      for (int r = 0; r < cnt; r += 4) {
        float4 b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) b[u] = ld_stream4(bank + min(ib + r + u, N - 1) * D + d);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int t = 0; t < QT; ++t) {
            const float kv = ks[t][r + u];  // zero beyond cnt: clamped duplicate rows add nothing
            acc[t].x = fmaf(kv, b[u].x, acc[t].x);
            acc[t].y = fmaf(kv, b[u].y, acc[t].y);
            acc[t].z = fmaf(kv, b[u].z, acc[t].z);
            acc[t].w = fmaf(kv, b[u].w, acc[t].w);
          }
      }
    }
  }
  if (d < D) {
#pragma unroll
    for (int t = 0; t < QT; ++t) {
      if (q0 + t >= Q) break;
      float* o = num + (q0 + t) * D + d;
      if (nsplit == 1) {
        *reinterpret_cast<float4*>(o) = acc[t];
      } else {
        atomicAdd(o + 0, acc[t].x);
        atomicAdd(o + 1, acc[t].y);
        atomicAdd(o + 2, acc[t].z);
        atomicAdd(o + 3, acc[t].w);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_sparse_apply(const float* __restrict__ num, const float* __restrict__ wsum, int64_t D, float scale,
               const float* xq, float* x0, float* __restrict__ term_out) {
  const int64_t q = blockIdx.y;
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j >= D) return;
  const float ws = wsum[q];
  float4 x = *reinterpret_cast<const float4*>(x0 + q * D + j);
  const float4 xf = *reinterpret_cast<const float4*>(xq + q * D + j);   // the query the force is taken on
  const float4 s = *reinterpret_cast<const float4*>(num + q * D + j);
  // sum_i w_i (x - n_i) = (sum w) x - sum_i w_i n_i      (fast.py:321-328)
  float4 t = make_float4(fmaf(ws, xf.x, -s.x), fmaf(ws, xf.y, -s.y), fmaf(ws, xf.z, -s.z), fmaf(ws, xf.w, -s.w));
  x.x = fmaf(scale, t.x, x.x);
  x.y = fmaf(scale, t.y, x.y);
  x.z = fmaf(scale, t.z, x.z);
  x.w = fmaf(scale, t.w, x.w);
  *reinterpret_cast<float4*>(x0 + q * D + j) = x;
  if (term_out) *reinterpret_cast<float4*>(term_out + q * D + j) = t;
}

// ------------------------------------------------------------------------------------------

int generic_dots(const float* bank, int64_t N, int64_t D, const float* xq, int64_t Q, float* S,
                 cudaStream_t st) {
  constexpr int R = 4;
  const unsigned gx = (unsigned)cdiv(N, 8 * R);
  if (Q >= 8) {
    k_dots<8, R><<<dim3(gx, (unsigned)cdiv(Q, 8)), 256, 0, st>>>(bank, N, D, xq, Q, S);
  } else if (Q >= 4) {
    k_dots<4, R><<<dim3(gx, (unsigned)cdiv(Q, 4)), 256, 0, st>>>(bank, N, D, xq, Q, S);
  } else if (Q >= 2) {
    k_dots<2, R><<<dim3(gx, (unsigned)cdiv(Q, 2)), 256, 0, st>>>(bank, N, D, xq, Q, S);
  } else {
    k_dots<1, R><<<dim3(gx, 1), 256, 0, st>>>(bank, N, D, xq, Q, S);
  }
  SDN_LAUNCHED();
  return SDN_OK;
}

int generic_weights(float* S, const float* sqnorm, const float* xsq, int64_t Q, int64_t N,
                    float inv2s2, int power, float alpha, float* z, cudaStream_t st) {
  k_weights<<<(unsigned)Q, 256, 0, st>>>(S, sqnorm, xsq, N, inv2s2, power, alpha, z);
  SDN_LAUNCHED();
  return SDN_OK;
}

int sparse_weights(float* S, const float* sqnorm, const float* xsq, int64_t Q, int64_t N,
                   float radius, float* wsum, cudaStream_t st) {
  k_sparse_weights<<<(unsigned)Q, 256, 0, st>>>(S, sqnorm, xsq, N, radius, wsum);
  SDN_LAUNCHED();
  return SDN_OK;
}

int sparse_apply(const float* num, const float* wsum, int64_t Q, int64_t D, float scale, const float* xq,
                 float* x0_inout, float* term_out, cudaStream_t st) {
  k_sparse_apply<<<dim3((unsigned)cdiv(D, 1024), (unsigned)Q), 256, 0, st>>>(num, wsum, D, scale, xq, x0_inout,
                                                                              term_out);
  SDN_LAUNCHED();
  return SDN_OK;
}

int generic_accum(const float* bank, int64_t N, int64_t D, const float* k, int64_t Q, float* num,
                  cudaStream_t st) {
  const int64_t dblocks = cdiv(D, 512);
  const int QT = Q >= 8 ? 8 : (Q >= 4 ? 4 : (Q >= 2 ? 2 : 1));
  const int64_t qtiles = cdiv(Q, QT);
  // enough blocks for ~4 per SM, but at least 32 rows per split
  int64_t nsplit = cdiv(4 * kNumSMs, dblocks * qtiles);
  nsplit = std::max<int64_t>(1, std::min<int64_t>(nsplit, std::max<int64_t>(1, N / 32)));
  nsplit = std::min<int64_t>(nsplit, 65535);
  if (nsplit > 1) SDN_CUDA_OK(cudaMemsetAsync(num, 0, sizeof(float) * Q * D, st));
  const dim3 grid((unsigned)dblocks, (unsigned)nsplit, (unsigned)qtiles);
  switch (QT) {
    case 8: k_accum<8><<<grid, 128, 0, st>>>(bank, N, D, k, Q, num, (int)nsplit); break;
    case 4: k_accum<4><<<grid, 128, 0, st>>>(bank, N, D, k, Q, num, (int)nsplit); break;
    case 2: k_accum<2><<<grid, 128, 0, st>>>(bank, N, D, k, Q, num, (int)nsplit); break;
    default: k_accum<1><<<grid, 128, 0, st>>>(bank, N, D, k, Q, num, (int)nsplit); break;
  }
  SDN_LAUNCHED();
  return SDN_OK;
}

}  // namespace sdn
