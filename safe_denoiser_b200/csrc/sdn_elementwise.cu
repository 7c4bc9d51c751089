// Bank / query preparation and the fused elementwise epilogues (one HBM pass over Q*D each).
#include <algorithm>

#include "sdn_internal.h"

namespace sdn {

// ---------------------------------------------------------------- bank prepare
// One warp per bank row: ||n_i||^2 and (optionally) the bf16 hi/lo planes the tcgen05 path reads.
__global__ void __launch_bounds__(256)
k_bank_prepare(const float* __restrict__ bank, int64_t N, int64_t D, float* __restrict__ sqnorm,
               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= N) return;
  const float* row = bank + i * D;
  float s = 0.f;
  for (int64_t j = lane * 4; j < D; j += 128) {
    const float4 v = ld_stream4(row + j);
    s = fmaf(v.x, v.x, s);
    s = fmaf(v.y, v.y, s);
    s = fmaf(v.z, v.z, s);
    s = fmaf(v.w, v.w, s);
    if (hi) {
      const float f[4] = {v.x, v.y, v.z, v.w};
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        h[u] = __float2bfloat16_rn(f[u]);
        l[u] = __float2bfloat16_rn(f[u] - __bfloat162float(h[u]));
      }
      *reinterpret_cast<uint2*>(hi + i * D + j) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(lo + i * D + j) = *reinterpret_cast<const uint2*>(l);
    }
  }
  s = warp_sum(s);
  if (lane == 0) sqnorm[i] = s;
}

// Bank build (RepellencyMethod.project, fast.py:45-70, after the VAE): per-pixel channel L2-normalisation of the
// raw latents fused with ||n_i||^2 and the bf16 planes -- one pass instead of norm + div + the prepare pass.
// One thread per pixel, channels strided by HW.
__global__ void __launch_bounds__(256)
k_bank_build(const float* __restrict__ latents, int64_t HW, int C, float* __restrict__ bank,
             float* __restrict__ sqnorm, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float red[33];
  const int64_t n = blockIdx.y;
  const int64_t base = n * HW * C;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  float s = 0.f;
  if (p < HW) {
    float ss = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = latents[base + (int64_t)c * HW + p];
      ss = fmaf(v, v, ss);
    }
    const float nrm = sqrtf(ss);
    for (int c = 0; c < C; ++c) {
      const int64_t o = base + (int64_t)c * HW + p;
      const float u = latents[o] / nrm;
      bank[o] = u;
      if (hi) {
        const __nv_bfloat16 h = __float2bfloat16_rn(u);
        hi[o] = h;
        lo[o] = __float2bfloat16_rn(u - __bfloat162float(h));
      }
      s = fmaf(u, u, s);
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(sqnorm + n, s);
}

// ---------------------------------------------------------------- query prepare
__global__ void __launch_bounds__(256)
k_query_prepare(const float* __restrict__ x_in, const float* __restrict__ m, float c_x, float c_m,
                int64_t D, float* __restrict__ x0_out, float* __restrict__ xq_out,
                float* __restrict__ xsq) {
  __shared__ float red[33];
  const int64_t q = blockIdx.y;
  const int64_t base = q * D;
  float s = 0.f;
  const int64_t j0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  const int64_t stride = (int64_t)gridDim.x * 1024;
  for (int64_t j = j0; j < D; j += stride) {
    float4 v = *reinterpret_cast<const float4*>(x_in + base + j);
    if (m) {
      const float4 e = *reinterpret_cast<const float4*>(m + base + j);
      v.x = fmaf(c_m, e.x, c_x * v.x);
      v.y = fmaf(c_m, e.y, c_x * v.y);
      v.z = fmaf(c_m, e.z, c_x * v.z);
      v.w = fmaf(c_m, e.w, c_x * v.w);
    } else if (c_x != 1.f) {
      v.x *= c_x; v.y *= c_x; v.z *= c_x; v.w *= c_x;
    }
    if (x0_out) *reinterpret_cast<float4*>(x0_out + base + j) = v;
    if (xq_out && xq_out != x0_out) *reinterpret_cast<float4*>(xq_out + base + j) = v;
    s = fmaf(v.x, v.x, s);
    s = fmaf(v.y, v.y, s);
    s = fmaf(v.z, v.z, s);
    s = fmaf(v.w, v.w, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(xsq + q, s);
}

// SD3: one thread per pixel, channels strided by HW (fast_sdv3.py:239).
__global__ void __launch_bounds__(256)
k_query_prepare_norm(const float* __restrict__ x_in, const float* __restrict__ m, float c_x, float c_m,
                     int64_t HW, int C, float* __restrict__ x0_out, float* __restrict__ xq_out,
                     float* __restrict__ xsq) {
  __shared__ float red[33];
  const int64_t q = blockIdx.y;
  const int64_t base = q * HW * C;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  float s = 0.f;
  if (p < HW) {
    float ss = 0.f;
    for (int c = 0; c < C; ++c) {
      const int64_t o = base + (int64_t)c * HW + p;
      float v = c_x * x_in[o];
      if (m) v = fmaf(c_m, m[o], v);
      if (x0_out) x0_out[o] = v;
      ss = fmaf(v, v, ss);
    }
    const float nrm = sqrtf(ss);
    for (int c = 0; c < C; ++c) {
      const int64_t o = base + (int64_t)c * HW + p;
      float v = c_x * x_in[o];
      if (m) v = fmaf(c_m, m[o], v);
      const float u = v / nrm;
      xq_out[o] = u;
      s = fmaf(u, u, s);
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(xsq + q, s);
}

// ---------------------------------------------------------------- epilogues
enum { EPI_CORRECT = 0, EPI_DDPM = 1, EPI_DDIM = 2, EPI_FLOW = 3 };

struct EpiArgs {
  const float* num; const float* z;
  int64_t D; int64_t QD;
  float eps, scale, gate_thr; int flags;
  float* x0_inout;        // CORRECT
  const float* x_t; const float* m; const float* z1; const float* z2;
  float a0, a1, a2, a3, a4;  // mode-specific coefficients
  float* out; float* aux_out;  // latents / neg_out (CORRECT) ; x0c_out
  float* denom_out; int32_t* gate_out; float* mean_out;
};

__device__ __forceinline__ float clamp10(float v) { return fminf(fmaxf(v, -1e10f), 1e10f); }

template <int MODE>
__global__ void __launch_bounds__(256) k_epilogue(const EpiArgs a) {
  __shared__ float red[33];
  const int64_t q = blockIdx.y;
  const float denom = a.z[q] + a.eps;
  const bool gate = !(a.flags & SDN_EPI_GATE) || (denom > a.gate_thr);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (a.denom_out) a.denom_out[q] = denom;
    if (a.gate_out) a.gate_out[q] = gate ? 1 : 0;
  }
  const bool ret_neg = (a.flags & SDN_EPI_RETURN_NEG) != 0;
  float msum = 0.f;
  const int64_t j = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (j < a.D) {
    const int64_t o = q * a.D + j;
    const float4 nm = *reinterpret_cast<const float4*>(a.num + o);
    float neg[4] = {nm.x / denom, nm.y / denom, nm.z / denom, nm.w / denom};
    float res[4], x0c[4];
    if (MODE == EPI_CORRECT) {
      if (a.x0_inout) {
        const float4 x = *reinterpret_cast<const float4*>(a.x0_inout + o);
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) x0c[u] = fmaf(-a.scale, neg[u], xv[u]);
        *reinterpret_cast<float4*>(a.x0_inout + o) = make_float4(x0c[0], x0c[1], x0c[2], x0c[3]);
      }
      if (a.out) *reinterpret_cast<float4*>(a.out + o) = make_float4(neg[0], neg[1], neg[2], neg[3]);
    } else if (MODE == EPI_DDPM || MODE == EPI_DDIM) {
      // a0 = sqrt(abar), a1 = sqrt(1-abar); DDPM: a2 = c_x0, a3 = c_xt, a4 = sigma_noise
      //                                       DDIM: a2 = sqrt(abar_prev), a3 = sqrt(1-abar_prev)
      const float4 xt4 = *reinterpret_cast<const float4*>(a.x_t + o);
      const float4 e4 = *reinterpret_cast<const float4*>(a.m + o);
      const float xt[4] = {xt4.x, xt4.y, xt4.z, xt4.w};
      const float e[4] = {e4.x, e4.y, e4.z, e4.w};
      float n1[4] = {0.f, 0.f, 0.f, 0.f}, n2[4] = {0.f, 0.f, 0.f, 0.f};
      if (gate) {
        const float4 t = *reinterpret_cast<const float4*>(a.z1 + o);
        n1[0] = t.x; n1[1] = t.y; n1[2] = t.z; n1[3] = t.w;
      }
      if (MODE == EPI_DDPM && a.z2) {
        const float4 t = *reinterpret_cast<const float4*>(a.z2 + o);
        n2[0] = t.x; n2[1] = t.y; n2[2] = t.z; n2[3] = t.w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float x0 = (xt[u] - a.a1 * e[u]) / a.a0;
        x0c[u] = fmaf(-a.scale, neg[u], x0);
        const float src = ret_neg ? neg[u] : x0c[u];
        const float xt2 = gate ? fmaf(a.a0, src, a.a1 * n1[u]) : xt[u];
        const float x0b = (xt2 - a.a1 * e[u]) / a.a0;
        if (MODE == EPI_DDPM) res[u] = fmaf(a.a2, x0b, fmaf(a.a3, xt2, a.a4 * n2[u]));
        else res[u] = fmaf(a.a2, x0b, a.a3 * e[u]);
      }
      *reinterpret_cast<float4*>(a.out + o) = make_float4(res[0], res[1], res[2], res[3]);
      if (a.aux_out) *reinterpret_cast<float4*>(a.aux_out + o) = make_float4(x0c[0], x0c[1], x0c[2], x0c[3]);
    } else {  // EPI_FLOW: a0 = sigma, a1 = sigma_next
      const float4 x4 = *reinterpret_cast<const float4*>(a.x_t + o);
      const float4 v4 = *reinterpret_cast<const float4*>(a.m + o);
      const float4 t4 = *reinterpret_cast<const float4*>(a.z1 + o);
      const float x[4] = {x4.x, x4.y, x4.z, x4.w};
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      const float zn[4] = {t4.x, t4.y, t4.z, t4.w};
      const float s_n = sqrtf(a.a1), s_1n = sqrtf(1.f - a.a1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float x0 = fmaf(-a.a0, v[u], x[u]);
        const float x1 = fmaf(1.f - a.a0, v[u], x[u]);
        x0c[u] = fmaf(-a.scale, neg[u], x0);
        const float noise = fmaf(s_n, x1, s_1n * zn[u]);
        res[u] = fmaf(a.a1, noise - x0c[u], x0c[u]);
      }
      *reinterpret_cast<float4*>(a.out + o) = make_float4(res[0], res[1], res[2], res[3]);
      if (a.aux_out) *reinterpret_cast<float4*>(a.aux_out + o) = make_float4(x0c[0], x0c[1], x0c[2], x0c[3]);
    }
    msum = clamp10(neg[0]) + clamp10(neg[1]) + clamp10(neg[2]) + clamp10(neg[3]);
  }
  if (a.mean_out) {
    msum = block_sum(msum, red);
    if (threadIdx.x == 0) atomicAdd(a.mean_out, msum / (float)a.QD);
  }
}

template <int MODE>
static int launch_epilogue(const EpiArgs& a, int64_t Q, cudaStream_t st) {
  k_epilogue<MODE><<<dim3((unsigned)cdiv(a.D, 1024), (unsigned)Q), 256, 0, st>>>(a);
  SDN_LAUNCHED();
  return SDN_OK;
}

}  // namespace sdn

using namespace sdn;

extern "C" {

int sdn_bank_prepare(const float* bank, int64_t N, int64_t D, float* sqnorm_out, void* planes_out,
                     void* stream) {
  if (!bank || !sqnorm_out) return SDN_E_NULL;
  if (N <= 0 || D <= 0) return SDN_E_SHAPE;
  if (D % 4 != 0 || !aligned16(bank) || (planes_out && !aligned16(planes_out))) return SDN_E_ALIGN;
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(planes_out);
  __nv_bfloat16* lo = hi ? hi + N * D : nullptr;
  k_bank_prepare<<<(unsigned)cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(bank, N, D, sqnorm_out, hi, lo);
  SDN_LAUNCHED();
  return SDN_OK;
}

int sdn_bank_build(const float* latents, int64_t N, int32_t C, int64_t HW, float* bank_out, float* sqnorm_out,
                   void* planes_out, void* stream) {
  if (!latents || !bank_out || !sqnorm_out) return SDN_E_NULL;
  if (N <= 0 || C <= 0 || HW <= 0 || N > 65535) return SDN_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  SDN_CUDA_OK(cudaMemsetAsync(sqnorm_out, 0, sizeof(float) * N, st));
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(planes_out);
  __nv_bfloat16* lo = hi ? hi + N * C * HW : nullptr;
  k_bank_build<<<dim3((unsigned)cdiv(HW, 256), (unsigned)N), 256, 0, st>>>(latents, HW, C, bank_out, sqnorm_out, hi, lo);
  SDN_LAUNCHED();
  return SDN_OK;
}

int sdn_query_prepare(const float* x_in, const float* model_out, float c_x, float c_m, int64_t Q,
                      int64_t D, int32_t normalize_C, float* x0_out, float* xq_out, float* xsq_out,
                      void* stream) {
  if (!x_in || !xsq_out) return SDN_E_NULL;
  if (Q <= 0 || D <= 0 || Q > 65535) return SDN_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  SDN_CUDA_OK(cudaMemsetAsync(xsq_out, 0, sizeof(float) * Q, st));
  if (normalize_C > 0) {
    if (!xq_out) return SDN_E_NULL;
    if (D % normalize_C != 0) return SDN_E_SHAPE;
    const int64_t HW = D / normalize_C;
    k_query_prepare_norm<<<dim3((unsigned)cdiv(HW, 256), (unsigned)Q), 256, 0, st>>>(
        x_in, model_out, c_x, c_m, HW, normalize_C, x0_out, xq_out, xsq_out);
  } else {
    if (D % 4 != 0 || !aligned16(x_in) || (model_out && !aligned16(model_out)) ||
        (x0_out && !aligned16(x0_out)) || (xq_out && !aligned16(xq_out)))
      return SDN_E_ALIGN;
    const unsigned gx = (unsigned)std::min<int64_t>(cdiv(D, 1024), 64);
    k_query_prepare<<<dim3(gx, (unsigned)Q), 256, 0, st>>>(x_in, model_out, c_x, c_m, D, x0_out, xq_out,
                                                           xsq_out);
  }
  SDN_LAUNCHED();
  return SDN_OK;
}

static int check_epi(const float* num, const float* z, int64_t Q, int64_t D) {
  if (!num || !z) return SDN_E_NULL;
  if (Q <= 0 || D <= 0 || Q > 65535) return SDN_E_SHAPE;
  if (D % 4 != 0 || !aligned16(num)) return SDN_E_ALIGN;
  return SDN_OK;
}

int sdn_epilogue_correct(const float* num, const float* z, int64_t Q, int64_t D, float eps, float scale,
                         float gate_threshold, int32_t flags, float* x0_inout, float* neg_out,
                         float* denom_out, int32_t* gate_out, float* mean_out, void* stream) {
  int rc = check_epi(num, z, Q, D);
  if (rc) return rc;
  if (!x0_inout && !neg_out) return SDN_E_NULL;
  if ((x0_inout && !aligned16(x0_inout)) || (neg_out && !aligned16(neg_out))) return SDN_E_ALIGN;
  EpiArgs a{};
  a.num = num; a.z = z; a.D = D; a.QD = Q * D; a.eps = eps; a.scale = scale; a.gate_thr = gate_threshold;
  a.flags = flags; a.x0_inout = x0_inout; a.out = neg_out; a.denom_out = denom_out; a.gate_out = gate_out;
  a.mean_out = mean_out;
  return launch_epilogue<EPI_CORRECT>(a, Q, (cudaStream_t)stream);
}

int sdn_epilogue_ddpm(const float* num, const float* z, int64_t Q, int64_t D, float eps, float scale,
                      float gate_threshold, int32_t flags, const float* x_t, const float* eps_pred,
                      const float* z1, const float* z2, float sqrt_ab, float sqrt_1m_ab, float c_x0,
                      float c_xt, float sigma_noise, float* latents_out, float* x0c_out, float* denom_out,
                      int32_t* gate_out, float* mean_out, void* stream) {
  int rc = check_epi(num, z, Q, D);
  if (rc) return rc;
  if (!x_t || !eps_pred || !z1 || !latents_out) return SDN_E_NULL;
  if (!z2 && sigma_noise != 0.f) return SDN_E_NULL;
  if (!aligned16(x_t) || !aligned16(eps_pred) || !aligned16(z1) || (z2 && !aligned16(z2)) ||
      !aligned16(latents_out) || (x0c_out && !aligned16(x0c_out)))
    return SDN_E_ALIGN;
  EpiArgs a{};
  a.num = num; a.z = z; a.D = D; a.QD = Q * D; a.eps = eps; a.scale = scale; a.gate_thr = gate_threshold;
  a.flags = flags; a.x_t = x_t; a.m = eps_pred; a.z1 = z1; a.z2 = z2;
  a.a0 = sqrt_ab; a.a1 = sqrt_1m_ab; a.a2 = c_x0; a.a3 = c_xt; a.a4 = sigma_noise;
  a.out = latents_out; a.aux_out = x0c_out; a.denom_out = denom_out; a.gate_out = gate_out;
  a.mean_out = mean_out;
  return launch_epilogue<EPI_DDPM>(a, Q, (cudaStream_t)stream);
}

int sdn_epilogue_ddim(const float* num, const float* z, int64_t Q, int64_t D, float eps, float scale,
                      float gate_threshold, int32_t flags, const float* x_t, const float* eps_pred,
                      const float* z1, float sqrt_ab, float sqrt_1m_ab, float sqrt_ab_prev,
                      float sqrt_1m_ab_prev, float* latents_out, float* x0c_out, float* denom_out,
                      int32_t* gate_out, float* mean_out, void* stream) {
  int rc = check_epi(num, z, Q, D);
  if (rc) return rc;
  if (!x_t || !eps_pred || !z1 || !latents_out) return SDN_E_NULL;
  if (!aligned16(x_t) || !aligned16(eps_pred) || !aligned16(z1) || !aligned16(latents_out) ||
      (x0c_out && !aligned16(x0c_out)))
    return SDN_E_ALIGN;
  EpiArgs a{};
  a.num = num; a.z = z; a.D = D; a.QD = Q * D; a.eps = eps; a.scale = scale; a.gate_thr = gate_threshold;
  a.flags = flags; a.x_t = x_t; a.m = eps_pred; a.z1 = z1;
  a.a0 = sqrt_ab; a.a1 = sqrt_1m_ab; a.a2 = sqrt_ab_prev; a.a3 = sqrt_1m_ab_prev;
  a.out = latents_out; a.aux_out = x0c_out; a.denom_out = denom_out; a.gate_out = gate_out;
  a.mean_out = mean_out;
  return launch_epilogue<EPI_DDIM>(a, Q, (cudaStream_t)stream);
}

int sdn_epilogue_flow(const float* num, const float* z, int64_t Q, int64_t D, float eps, float scale,
                      const float* x, const float* v, const float* zn, float sigma, float sigma_next,
                      float* latents_out, float* x0c_out, float* denom_out, float* mean_out,
                      void* stream) {
  int rc = check_epi(num, z, Q, D);
  if (rc) return rc;
  if (!x || !v || !zn || !latents_out) return SDN_E_NULL;
  if (!aligned16(x) || !aligned16(v) || !aligned16(zn) || !aligned16(latents_out) ||
      (x0c_out && !aligned16(x0c_out)))
    return SDN_E_ALIGN;
  EpiArgs a{};
  a.num = num; a.z = z; a.D = D; a.QD = Q * D; a.eps = eps; a.scale = scale; a.gate_thr = 0.f;
  a.flags = 0; a.x_t = x; a.m = v; a.z1 = zn; a.a0 = sigma; a.a1 = sigma_next;
  a.out = latents_out; a.aux_out = x0c_out; a.denom_out = denom_out; a.mean_out = mean_out;
  return launch_epilogue<EPI_FLOW>(a, Q, (cudaStream_t)stream);
}

}  // extern "C"
