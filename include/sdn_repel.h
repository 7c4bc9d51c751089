/*
 * sdn_repel.h -- C ABI of the B200-native repellency projection.
 *
 * The reference (MingyuKim87/Safe_Denoiser) is pure Python and has no FFI; the
 * boundary it exposes for this path is the Python object returned by
 * get_repellency_method() (repellency/repellency_methods_fast.py:19-22) and its
 * .conditioning(x_0_hat) method (:120-132).  Every entry point below replaces
 * one slice of what that method does with torch ops; the reference lines are
 * cited per function.  The modules under safe_denoiser_b200/repellency/ are the host-side
 * mirror of the reference interface and bind these symbols through ctypes
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers and sizes, no torch types, nothing thrown;
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - no hidden allocation: scratch is caller-provided, its size comes from
 *     the matching *_workspace_bytes() call.  Two exceptions: the *_host convenience
 *     calls own a cached device staging area, and SDN_PATH_FLASH owns ~37 MB of
 *     per-device synchronisation memory (exchange rings of its persistent grid,
 *     allocated and zeroed at its first call on a device -- so make that first call
 *     outside a CUDA-graph capture; calls of that path on one device must be stream-ordered
 *     with each other, its grid fills the GPU anyway);
 *   - return value: 0 ok; <0 invalid argument (SDN_E_*); >0 a cudaError_t.
 *   - all tensors are contiguous row-major fp32 unless stated.
 *
 * Shapes: Q query rows, N negatives in (this shard of) the bank, D = C*H*W.
 */
#ifndef SDN_REPEL_H_
#define SDN_REPEL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDN_ABI_VERSION 1

enum {
  SDN_OK = 0,
  SDN_E_NULL = -1,        /* required pointer is NULL */
  SDN_E_SHAPE = -2,       /* non-positive or inconsistent size */
  SDN_E_ALIGN = -3,       /* pointer / D not aligned as required (16 B, D % 4 == 0) */
  SDN_E_PARAM = -4,       /* unsupported dist_power / mode / flag */
  SDN_E_WORKSPACE = -5,   /* workspace too small */
  SDN_E_UNSUPPORTED = -6, /* shape outside what the selected kernel family handles */
  SDN_E_DEVICE = -7       /* not an sm_100 device */
};

/* Kernel family selector for sdn_repel_partial (SDN_PATH_AUTO picks by shape). */
enum {
  SDN_PATH_AUTO = 0,
  SDN_PATH_GENERIC = 1,   /* CUDA-core two-phase kernels, any shape */
  SDN_PATH_STREAM = 2,    /* one-pass cluster kernel, GEMV-shaped (small Q) */
  SDN_PATH_UMMA = 3,      /* tcgen05 / TMEM / TMA two-phase kernels, batched Q */
  SDN_PATH_UMMA_BF16 = 4, /* same kernels reading ONLY the bf16 hi plane of the bank: half the bytes per pass,
                             outside the 1e-3 parity tolerance near a negative; explicit opt-in, never AUTO */
  SDN_PATH_FLASH = 5      /* ONE-HBM-pass tcgen05 kernel, batched Q: both contractions in one persistent grid of D/128
                             co-resident CTAs, a few tiles apart, so that the second read of a bank tile hits the L2;
                             the dot products are reduced across the CTAs through L2 inside the kernel.
                             D % 1024 == 0, 8192 <= D <= 16384 (SD-1.4 latents), 64 query rows per pass.  Explicit opt-in
                             (or SDN_PREFER_FLASH=1 in the environment): it reads the bank once but is not yet faster
                             than SDN_PATH_UMMA */
};

/* Epilogue flags (bit-or). */
enum {
  SDN_EPI_GATE = 1,          /* apply the beta gate: row q is re-noised only if Z_q + eps > gate_threshold
                                (threshold.py:181-185); without it every row is treated as negated */
  SDN_EPI_RETURN_NEG = 2     /* the tensor handed on to the re-noise is the negative mean, not the
                                corrected x0 (threshold.py:190-193) */
};

int sdn_abi_version(void);
const char* sdn_error_string(int code);
/* Number of kernels this library has launched since load (all entry points); bench.py reports it. */
uint64_t sdn_launch_count(void);

/* Library-wide options.
 * SDN_OPT_SKIP_NEGLIGIBLE (default 1): the tcgen05 accumulate pass does not read blocks of 64 bank rows in which
 *   every weight k_qi is below 1e-9 x the largest weight of its query (their total contribution is < N * 1e-9 of
 *   the dominant term, below fp32 summation noise).  Exact to rounding, data dependent; 0 forces the dense pass
 *   (bench.py's default, so that its numbers are the worst case). */
enum { SDN_OPT_SKIP_NEGLIGIBLE = 1 };
int sdn_set_option(int32_t key, int32_t value);

/* Per-kernel timing of the LAST sdn_repel_partial call (CUDA events on its stream).  Off by default.
 * sdn_profile_read(i, ...) returns 1 and fills the i-th kernel's name and duration in ms (it synchronises on
 * that kernel's end event), 0 when i is out of range.  bench.py uses it for the roofline of the dominant kernel. */
void sdn_profile_enable(int32_t on);
int32_t sdn_profile_read(int32_t index, char* name_out, int32_t name_cap, float* ms_out);

/* After a kernel of the one-pass path trapped on a bounded wait (CUDA error "unspecified launch failure"), the
 * first words of its host-mapped diagnostic record: {code, CTA, thread, tile, extra}.  Returns 0 when there is none. */
int32_t sdn_debug_read(uint32_t* words_out, int32_t n);
/* With SDN_FLASH_TRACE=1 in the environment at the first one-pass call, the kernel records %globaltimer at the hand-offs
 * of its pipeline: uint64 [128 CTAs][64 tiles][16 events] (1 MiB) of the last launch, copied to host_out.  Returns
 * the bytes written, 0 when tracing is off or host_out is too small.  (tools/gpu_flash_trace.py prints the stage latencies.) */
size_t sdn_debug_trace_read(void* host_out, size_t bytes);
/* Timestamps of the last traced phase-B launch (SDN_UMMA_DBG_NOSHARED bit 10): [256 CTAs][8 events] uint64 ns. */
size_t sdn_debug_accum_trace_read(void* host_out, size_t bytes);

/* ---- bank -------------------------------------------------------------------------------
 * Derived data of the proj_ref tensor, computed once at load (fast.py:109-111 loads the tensor;
 * torch.cdist recomputes ||n_i||^2 on every call, fast.py:249).
 *   sqnorm_out [N]        fp32 ||n_i||^2
 *   planes_out [2][N][D]  bf16 hi plane then lo plane (hi = bf16(n), lo = bf16(n - hi)); may be NULL
 */
int sdn_bank_prepare(const float* bank, int64_t N, int64_t D,
                     float* sqnorm_out, void* planes_out, void* stream);

/* Bank build from raw VAE latents [N][C][HW] (RepellencyMethod.project, fast.py:45-70): divides every pixel by its
 * L2 norm over the C channels (so that ||n_i||^2 = HW) and produces bank_out [N][C*HW] fp32 -- the tensor the
 * proj_ref cache stores -- together with sqnorm_out and (optionally) the planes, in one pass.  bank_out may alias
 * latents. */
int sdn_bank_build(const float* latents, int64_t N, int32_t C, int64_t HW,
                   float* bank_out, float* sqnorm_out, void* planes_out, void* stream);

/* ---- query ------------------------------------------------------------------------------
 * x0 = c_x * x_in + c_m * model_out   (model_out may be NULL -> x0 = c_x * x_in)
 *   DDPM eps-prediction (diffusers DDPMScheduler.step -> pred_original_sample, called at
 *   ...threshold_time.py:554): c_x = 1/sqrt(abar_t), c_m = -sqrt(1-abar_t)/sqrt(abar_t)
 *   SD3 flow (safe_denoiser_pipeline.py:1146): c_x = 1, c_m = -sigma
 * xq = x0, or x0 divided by its per-pixel L2 norm over the C channels when normalize_C > 0
 *   (fast_sdv3.py:239; D = normalize_C * HW).
 * x0_out [Q,D] (may alias x_in; may be NULL if not wanted), xq_out [Q,D] (may be NULL when
 * normalize_C == 0 and x0_out is given: then xq == x0), xsq_out [Q] = ||xq||^2.
 */
int sdn_query_prepare(const float* x_in, const float* model_out, float c_x, float c_m,
                      int64_t Q, int64_t D, int32_t normalize_C,
                      float* x0_out, float* xq_out, float* xsq_out, void* stream);

/* ---- projection -------------------------------------------------------------------------
 * Un-normalised sums over the rows of this bank (shard):
 *   dist_qi = sqrt(max(||xq||^2 + a^2 ||n_i||^2 - 2 a xq.n_i, 0))      (dist_power 1, the reference,
 *             fast.py:249: torch.cdist, un-squared)   or the squared form (dist_power 2, fast.py:413)
 *   k_qi    = exp(-dist_qi * inv_two_sigma_sq)                           (raw exp, no max-shift, Q3)
 *   num_out [Q,D] = sum_i k_qi n_i ,  z_out [Q] = sum_i k_qi             (fast.py:250; the ones column)
 * k_out [Q,N] receives k_qi when non-NULL (tests / weights parity).  num_out may be NULL when only
 * z_out is wanted (empirical_beta, threshold.py:351-384).
 * N-sharding: run this per shard, all-reduce(sum) num_out and z_out, then call an epilogue.
 */
size_t sdn_repel_workspace_bytes(int64_t Q, int64_t N, int64_t D, int32_t path);
/* Which kernel family sdn_repel_partial would run for this shape (SDN_PATH_*), given whether planes exist.
 * xsq may be NULL in sdn_repel_partial when the answer is SDN_PATH_STREAM or SDN_PATH_UMMA(_BF16): those kernels
 * compute ||xq||^2 themselves, which saves the sdn_query_prepare launch for a plain query. */
int32_t sdn_repel_path(int64_t Q, int64_t N, int64_t D, int32_t has_planes, int32_t path);

int sdn_repel_partial(const float* bank, const float* sqnorm, const void* planes,
                      int64_t N, int64_t D,
                      const float* xq, const float* xsq, int64_t Q,
                      float inv_two_sigma_sq, int32_t dist_power, float bank_alpha,
                      float* num_out, float* z_out, float* k_out,
                      void* workspace, size_t workspace_bytes, int32_t path, void* stream);

/* ---- epilogues (one pass over Q*D) ------------------------------------------------------
 * Common part (fast.py:253-257, :131; threshold.py:181-187):
 *   denom_q = z_q + eps ; neg = num / denom_q ; gate_q = !(flags&GATE) || denom_q > gate_threshold
 *   x0c = x0 - scale * neg
 * Outputs that may be NULL are skipped.  mean_out[0] += sum(clamp(neg, +-1e10)) / (Q*D)  (the
 * reference's logging scalar, fast.py:258); the caller zeroes it.
 */

/* conditioning(): x0_inout <- x0c in place (Q8); neg_out [Q,D]; denom_out [Q]; gate_out [Q] int32.
 * x0_inout may be NULL when only neg_out is wanted (empirical_denoiser alone, fast.py:223-262). */
int sdn_epilogue_correct(const float* num, const float* z, int64_t Q, int64_t D,
                         float eps, float scale, float gate_threshold, int32_t flags,
                         float* x0_inout, float* neg_out, float* denom_out, int32_t* gate_out,
                         float* mean_out, void* stream);

/* SD-1.4 in-window step (...threshold_time.py:554-576 with diffusers DDPMScheduler):
 *   x0 = (x_t - s1 eps)/sa ; src = RETURN_NEG ? neg : x0c
 *   x_t' = gate ? sa*src + s1*z1 : x_t ; x0'' = (x_t' - s1 eps)/sa
 *   latents_out = c_x0*x0'' + c_xt*x_t' + sigma_noise*z2
 * with sa = sqrt(abar_t), s1 = sqrt(1-abar_t).  z2 may be NULL when sigma_noise == 0.
 * x0c_out [Q,D] optional (the corrected x0 the reference would have returned).
 */
int sdn_epilogue_ddpm(const float* num, const float* z, int64_t Q, int64_t D,
                      float eps, float scale, float gate_threshold, int32_t flags,
                      const float* x_t, const float* eps_pred, const float* z1, const float* z2,
                      float sqrt_ab, float sqrt_1m_ab, float c_x0, float c_xt, float sigma_noise,
                      float* latents_out, float* x0c_out, float* denom_out, int32_t* gate_out,
                      float* mean_out, void* stream);

/* DDIM (eta = 0) variant: latents_out = sqrt_ab_prev * x0'' + sqrt_1m_ab_prev * eps. */
int sdn_epilogue_ddim(const float* num, const float* z, int64_t Q, int64_t D,
                      float eps, float scale, float gate_threshold, int32_t flags,
                      const float* x_t, const float* eps_pred, const float* z1,
                      float sqrt_ab, float sqrt_1m_ab, float sqrt_ab_prev, float sqrt_1m_ab_prev,
                      float* latents_out, float* x0c_out, float* denom_out, int32_t* gate_out,
                      float* mean_out, void* stream);

/* SD3 flow-matching step (safe_denoiser_pipeline.py:1142-1161):
 *   x0 = x - sigma v ; x1 = x + (1-sigma) v ; noise = sqrt(sigma_next) x1 + sqrt(1-sigma_next) zn
 *   latents_out = x0c + sigma_next (noise - x0c)
 */
int sdn_epilogue_flow(const float* num, const float* z, int64_t Q, int64_t D,
                      float eps, float scale,
                      const float* x, const float* v, const float* zn,
                      float sigma, float sigma_next,
                      float* latents_out, float* x0c_out, float* denom_out,
                      float* mean_out, void* stream);

/* ---- conditioning() with the fewest launches (one GPU) ------------------------------------------------------
 * = sdn_query_prepare + sdn_repel_partial + sdn_epilogue_correct for a query that needs no eps->x0 conversion and
 * no channel normalisation (fast.py:120-132, threshold.py:171-193).
 *   Q <= 8 : the one-pass kernel computes ||x||^2 itself and the per-cluster reduction applies the correction
 *            (2 launches);
 *   Q > 8  : (needs `planes` and z_out) query planes + ||x||^2 in one kernel, phase A, weights, phase B (which also sums
 *            z), and a row-contiguous kernel that applies the correction from phase B's tile scratch (5 launches
 *            instead of 8; chunked / split / block-sparse passes correct in phase B's epilogue: 4).  One pass over
 *            the bank serves up to 128 query rows (two groups
 *            of 64 sharing every bank tile; the small element-wise kernels run once per group), more rows take
 *            ceil(Q / 128) passes.
 * x0_inout [Q,D] is corrected in place; num_out / neg_out / k_out are optional extra outputs; mean_out is zeroed
 * by the call.  `path`: SDN_PATH_AUTO, or SDN_PATH_STREAM / SDN_PATH_UMMA / SDN_PATH_FLASH to force one family
 * (SDN_PATH_FLASH: ONE launch -- query planes, ||x||^2, both contractions, the cross-CTA reduction and the correction
 * in the same persistent kernel; the bank is read from HBM once).
 * Returns SDN_E_UNSUPPORTED for shapes the selected fused path does not take (callers then use the three-call
 * sequence).  Workspace: sdn_repel_workspace_bytes(Q, N, D, SDN_PATH_AUTO). */
int sdn_conditioning_fused(const float* bank, const float* sqnorm, const void* planes, int64_t N, int64_t D,
                           float* x0_inout, int64_t Q, float inv_two_sigma_sq, int32_t dist_power, float bank_alpha,
                           float eps, float scale, float gate_threshold, int32_t flags,
                           float* num_out, float* z_out, float* neg_out, float* denom_out, int32_t* gate_out,
                           float* mean_out, float* k_out, void* workspace, size_t workspace_bytes, int32_t path,
                           void* stream);

/* ---- N-sharded banks: merge + correction in one kernel over NVLink peer memory ------------------------------
 * Replaces "all-reduce(num|z) then sdn_epilogue_correct" when the bank is sharded by rows over `world` GPUs of one
 * NVSwitch domain (the reference has no multi-GPU path; this is the exchange step of SURVEY 8e).
 * peer_packed[p] / peer_out[p] / peer_sig[p] are HOST arrays of `world` device pointers: rank p's partial buffer
 * [Q*D | Q] fp32 (written by sdn_repel_partial), result buffer [Q*D] fp32 and flag words (uint32 [16], zeroed once),
 * all peer-mapped in this process (CUDA IPC / symmetric memory).  Rank `rank` reduces D-slice `rank` over all peers
 * in rank order, applies x0' = x0 - scale * num / (sum_p z_p + eps) and stores the slice into every rank's result
 * buffer.  `epoch_word` is a local device uint32 initialised to 1 on every rank; the kernel reads the epoch of the
 * launch from it and advances it, so the call can be replayed from a CUDA graph.  `counter` is a local zeroed uint32 [4]:
 * word 0 is scratch, word 1 a sticky error flag the kernel raises when a peer did not show up within
 * SDN_SHARD_TIMEOUT_S seconds (default 60; a late host only delays the others, nothing traps) -- the result of that
 * step is then undefined and the caller should stop using the link.  Every rank must launch the call. */
int sdn_shard_merge_correct(const void* const* peer_packed, void* const* peer_out, void* const* peer_sig,
                            int32_t rank, int32_t world, void* epoch_word, int64_t Q, int64_t D,
                            float eps, float scale, float gate_threshold, int32_t flags,
                            const float* x0_local, float* denom_out, int32_t* gate_out, void* counter,
                            void* stream);

/* ---- SPELL baseline (fast.py:306-340, threshold.py:415-454) ------------------------------
 * dist_out [Q,N] = ||x_q - n_i|| from the same expansion; the force is
 *   term_q = sum_i relu(radius/d_qi - 1) (x_q - n_i) ; x0_inout += scale * term.
 * wsum_out [Q] = sum_i relu(radius/d_qi - 1)  (is_negation = wsum != 0, threshold.py:447-450).
 * xq (may be NULL = x0_inout) is the query the distances and the force are taken on; fast_sdv3.py:332 takes them on
 * the channel-normalised query while the update still lands on the un-normalised x0.  xsq = ||xq||^2.
 * Q <= 8 on a shape the one-pass cluster kernel supports, and workspace >= sdn_repel_workspace_bytes(Q, N, D,
 * SDN_PATH_STREAM) + Q*D*4 (+ 512): the bank is read ONCE (SPELL weight as a functor of k_stream); otherwise the
 * generic two-pass kernels run (workspace >= Q*N*4 + Q*D*4 + 512).
 */
int sdn_sparse_repel(const float* bank, const float* sqnorm, int64_t N, int64_t D,
                     float* x0_inout, const float* xq, const float* xsq, int64_t Q,
                     float radius, float scale,
                     float* term_out, float* wsum_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* The same in two steps, for N-sharded banks: un-normalised sums over this shard's rows
 *   num_out [Q,D] = sum_i w_i n_i ,  wsum_out [Q] = sum_i w_i ,  w_i = relu(radius/d_qi - 1)
 * (all-reduce both over the ranks), then term = wsum * xq - num ; x0_inout += scale * term (term_out optional). */
int sdn_sparse_partial(const float* bank, const float* sqnorm, int64_t N, int64_t D,
                       const float* xq, const float* xsq, int64_t Q, float radius,
                       float* num_out, float* wsum_out, void* workspace, size_t workspace_bytes, void* stream);
int sdn_sparse_apply(const float* num, const float* wsum, int64_t Q, int64_t D, float scale,
                     const float* xq, float* x0_inout, float* term_out, void* stream);
/* sdn_sparse_partial for batched queries on the tcgen05 kernels (the bf16 hi/lo planes of sdn_bank_prepare instead of
 * the fp32 bank): the SPELL weight replaces the Gaussian one in the weights step between the two contractions.
 * workspace >= sdn_repel_workspace_bytes(Q, N, D, SDN_PATH_UMMA).  Follow with sdn_sparse_apply. */
int sdn_sparse_partial_planes(const void* planes, const float* sqnorm, int64_t N, int64_t D,
                              const float* xq, const float* xsq, int64_t Q, float radius,
                              float* num_out, float* wsum_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- host-buffer convenience: the e2e path ------------------------------------------------
 * One conditioning() call with HOST tensors: H2D of x0_host [Q,D], projection over a device-resident
 * prepared bank, correction, D2H of the corrected x0 (in place) and denom_host [Q].  Synchronous.
 * normalize_C as in sdn_query_prepare.  The staging buffers are cached inside the library and
 * released by sdn_host_release().
 */
int sdn_conditioning_host(const float* bank, const float* sqnorm, const void* planes,
                          int64_t N, int64_t D,
                          float* x0_host, int64_t Q, int32_t normalize_C,
                          float inv_two_sigma_sq, int32_t dist_power, float bank_alpha,
                          float eps, float scale,
                          float* denom_host, int32_t path, void* stream);
void sdn_host_release(void);

/* ---- host-buffer calls in flight: the serving / throughput form of the e2e path -----------
 * The reference runs one conditioning() per sampler step and prompt (models/textuals_visual/
 * modified_safree_diffusion_pipeline_threshold_time.py:550-576); prompts are independent, so a server keeps several
 * calls in flight.  A pipe owns `slots` (1..8) device staging areas with one stream each for a fixed (Q, N, D).
 * submit: enqueue H2D of x0_in_host [Q,D] -> fused conditioning over the device-resident prepared bank -> D2H of the
 * corrected query into x0_out_host [Q,D] (may alias x0_in_host) and of denom_host [Q]; returns without waiting.
 * wait: block until that slot's results are in the host buffers.  The bank must be complete before the first submit
 * (there is no caller stream to order against).  Host buffers should be pinned (pageable memory makes the copies
 * synchronous).  SDN_E_UNSUPPORTED from submit: the shape has no fused sequence, use sdn_conditioning_host.
 */
int sdn_host_pipe_create(int64_t Q, int64_t N, int64_t D, int32_t slots, void** pipe_out);
int sdn_host_pipe_submit(void* pipe, int32_t slot, const float* bank, const float* sqnorm, const void* planes,
                         const float* x0_in_host, float* x0_out_host, float* denom_host,
                         float inv_two_sigma_sq, int32_t dist_power, float bank_alpha, float eps, float scale);
int sdn_host_pipe_wait(void* pipe, int32_t slot);
void sdn_host_pipe_destroy(void* pipe);

#ifdef __cplusplus
}
#endif
#endif /* SDN_REPEL_H_ */
