// Internal launch functions shared between translation units (not part of the C ABI).
#pragma once
#include <cuda.h>

#include "sdn_common.cuh"

namespace sdn {

// ---- generic CUDA-core two-phase path (sdn_generic.cu) ----
// S[q,i] = xq_q . n_i
int generic_dots(const float* bank, int64_t N, int64_t D, const float* xq, int64_t Q, float* S,
                 cudaStream_t st);
// in place: S[q,i] -> k_qi ; z[q] = sum_i k_qi   (deterministic, one block per row)
int generic_weights(float* S, const float* sqnorm, const float* xsq, int64_t Q, int64_t N,
                    float inv_two_sigma_sq, int power, float alpha, float* z, cudaStream_t st);
// num[q,:] = sum_i k[q,i] n_i
int generic_accum(const float* bank, int64_t N, int64_t D, const float* k, int64_t Q, float* num,
                  cudaStream_t st);
// in place: S[q,i] -> relu(radius / d_qi - 1) ; wsum[q] = sum_i of that   (SPELL)
int sparse_weights(float* S, const float* sqnorm, const float* xsq, int64_t Q, int64_t N,
                   float radius, float* wsum, cudaStream_t st);
int sparse_apply(const float* num, const float* wsum, int64_t Q, int64_t D, float scale, const float* xq,
                 float* x0_inout, float* term_out, cudaStream_t st);

// ---- one-pass cluster path for GEMV-shaped calls (sdn_stream.cu) ----
bool stream_supported(int64_t Q, int64_t N, int64_t D);
size_t stream_workspace_bytes(int64_t Q, int64_t N, int64_t D);
int stream_partial(const float* bank, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                   const float* xsq, int64_t Q, float inv_two_sigma_sq, int power, float alpha,
                   float* num, float* z, float* k_out, void* ws, size_t ws_bytes, cudaStream_t st);

// one-pass projection + correction in two launches (||x||^2 in-kernel, reduce fused with the epilogue)
int stream_conditioning(const float* bank, const float* sqnorm, int64_t N, int64_t D, float* x0_inout,
                        const float* xq, int64_t Q, float inv_two_sigma_sq, int power, float alpha, float eps,
                        float scale, float gate_thr, int flags, float* num_out, float* z_out, float* neg_out,
                        float* denom_out, int32_t* gate_out, float* mean_out, float* k_out, void* ws,
                        size_t ws_bytes, cudaStream_t st);

extern std::atomic<bool> g_skip_negligible;   // SDN_OPT_SKIP_NEGLIGIBLE

// ---- tcgen05 two-phase path for batched calls (sdn_umma.cu) ----
bool umma_supported(int64_t Q, int64_t N, int64_t D, const void* planes);
size_t umma_workspace_bytes(int64_t Q, int64_t N, int64_t D);
int umma_partial(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                 const float* xsq, int64_t Q, float inv_two_sigma_sq, int power, float alpha,
                 float* num, float* z, float* k_out, void* ws, size_t ws_bytes, cudaStream_t st,
                 bool bf16_bank = false);

int umma_conditioning(const void* planes, const float* sqnorm, int64_t N, int64_t D, float* x0_inout, int64_t Q,
                      float inv_two_sigma_sq, int power, float alpha, float eps, float scale, float gate_thr,
                      int flags, float* num_out, float* z, float* neg_out, float* denom_out, int32_t* gate_out,
                      float* mean_out, float* k_out, void* ws, size_t ws_bytes, cudaStream_t st);

// 2-D bf16 row-major tensor [rows][cols] -> TMA descriptor with box [box_rows][box_cols], 128-byte swizzle.
int tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);

// ---- one-pass tcgen05 path for batched calls (sdn_flash.cu): the bank is read once ----
struct FlashEpi {
  // fused != 0: the correction of conditioning() runs in the kernel's epilogue (otherwise num_out / z_out only)
  int fused;
  float eps, scale, gate_thr; int flags;
  float* x0; float* neg_out; float* denom_out; int32_t* gate_out; float* mean_out; float inv_qd;
  int zero_mean;        // CTA 0 zeroes mean_out at start
};
bool flash_shape_ok(int64_t Q, int64_t N, int64_t D);
bool flash_supported(int64_t Q, int64_t N, int64_t D, const void* planes);
int flash_run(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq, int64_t Q,
              float inv_two_sigma_sq, int power, float alpha, float* num_out, float* z_out, float* k_out,
              const FlashEpi* epi, cudaStream_t st);
int flash_diag_read(uint32_t* out, int n);
size_t flash_trace_read(void* host_out, size_t bytes);
size_t umma_accum_trace_read(void* host_out, size_t bytes);   // [256 CTAs][8 events] globaltimer ns of the last traced phase B

}  // namespace sdn
