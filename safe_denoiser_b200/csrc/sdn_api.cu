// extern "C" entry points: argument checking, kernel-family dispatch, host-buffer convenience.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "sdn_internal.h"

namespace sdn {
std::atomic<uint64_t> g_launches{0};
Prof g_prof;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t generic_workspace_bytes(int64_t Q, int64_t N) { return align_up(sizeof(float) * Q * N, 256); }

// Batched calls go to the tensor cores.  The one-pass CUDA-core kernel is FMA-bound from Q = 5 on (its Q = 5..8
// instance runs at ~25 % of the HBM roofline), so for a bank of >= 1024 rows the two-phase tcgen05 path wins there too.
static bool batched(int64_t Q, int64_t N) { return Q > 8 || (Q > 4 && N >= 1024); }

static int pick_path(int32_t path, int64_t Q, int64_t N, int64_t D, const void* planes) {
  if (path == SDN_PATH_AUTO) {
    // SDN_PATH_FLASH (one HBM pass) is opt-in: measured on B200 it is still slower than the two-phase kernels (its
    // weights exchange costs ~2 us per 64-row tile), except with SDN_PREFER_FLASH=1
    static const bool prefer_flash = [] { const char* e = getenv("SDN_PREFER_FLASH"); return e && atoi(e) != 0; }();
    if (prefer_flash && flash_supported(Q, N, D, planes) && batched(Q, N)) return SDN_PATH_FLASH;
    if (umma_supported(Q, N, D, planes) && batched(Q, N)) return SDN_PATH_UMMA;
    if (stream_supported(Q, N, D)) return SDN_PATH_STREAM;
    return SDN_PATH_GENERIC;
  }
  return path;
}
}  // namespace sdn

using namespace sdn;

extern "C" {

int sdn_abi_version(void) { return SDN_ABI_VERSION; }

uint64_t sdn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* sdn_error_string(int code) {
  switch (code) {
    case SDN_OK: return "ok";
    case SDN_E_NULL: return "required pointer is NULL";
    case SDN_E_SHAPE: return "non-positive or inconsistent size";
    case SDN_E_ALIGN: return "pointer not 16-byte aligned or D not a multiple of 4";
    case SDN_E_PARAM: return "unsupported parameter value";
    case SDN_E_WORKSPACE: return "workspace too small";
    case SDN_E_UNSUPPORTED: return "shape not supported by the selected kernel family";
    case SDN_E_DEVICE: return "device is not sm_100";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}

int sdn_set_option(int32_t key, int32_t value) {
  switch (key) {
    case SDN_OPT_SKIP_NEGLIGIBLE: g_skip_negligible.store(value != 0); return SDN_OK;
    default: return SDN_E_PARAM;
  }
}

void sdn_profile_enable(int32_t on) { g_prof.enabled = on != 0; g_prof.reset(); }

int32_t sdn_profile_read(int32_t index, char* name_out, int32_t name_cap, float* ms_out) {
  if (index < 0 || index >= g_prof.n) return 0;
  const ProfSlot& s = g_prof.slot[index];
  if (cudaEventSynchronize(s.e1) != cudaSuccess) return 0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, s.e0, s.e1) != cudaSuccess) return 0;
  if (ms_out) *ms_out = ms;
  if (name_out && name_cap > 0) {
    strncpy(name_out, s.name, (size_t)name_cap - 1);
    name_out[name_cap - 1] = 0;
  }
  return 1;
}

int32_t sdn_debug_read(uint32_t* words_out, int32_t n) {
  if (!words_out || n <= 0) return 0;
  return flash_diag_read(words_out, n);
}

size_t sdn_debug_trace_read(void* host_out, size_t bytes) { return flash_trace_read(host_out, bytes); }
size_t sdn_debug_accum_trace_read(void* host_out, size_t bytes) { return umma_accum_trace_read(host_out, bytes); }

int32_t sdn_repel_path(int64_t Q, int64_t N, int64_t D, int32_t has_planes, int32_t path) {
  if (Q <= 0 || N <= 0 || D <= 0) return SDN_E_SHAPE;
  static const char dummy = 0;
  return pick_path(path, Q, N, D, has_planes ? &dummy : nullptr);
}

size_t sdn_repel_workspace_bytes(int64_t Q, int64_t N, int64_t D, int32_t path) {
  if (Q <= 0 || N <= 0 || D <= 0) return 0;
  size_t need = generic_workspace_bytes(Q, N);
  if (path == SDN_PATH_AUTO || path == SDN_PATH_STREAM) need = std::max(need, stream_workspace_bytes(Q, N, D));
  if (path == SDN_PATH_AUTO || path == SDN_PATH_UMMA || path == SDN_PATH_UMMA_BF16)
    need = std::max(need, umma_workspace_bytes(Q, N, D));
  return need;
}

int sdn_repel_partial(const float* bank, const float* sqnorm, const void* planes, int64_t N, int64_t D,
                      const float* xq, const float* xsq, int64_t Q, float inv_two_sigma_sq,
                      int32_t dist_power, float bank_alpha, float* num_out, float* z_out, float* k_out,
                      void* workspace, size_t workspace_bytes, int32_t path, void* stream) {
  if (!sqnorm || !xq || !z_out) return SDN_E_NULL;  // num_out NULL: z only (empirical_beta)
  if (!bank && !planes) return SDN_E_NULL;
  if (Q <= 0 || N <= 0 || D <= 0 || Q > 65535) return SDN_E_SHAPE;
  if (dist_power != 1 && dist_power != 2) return SDN_E_PARAM;
  if (D % 4 != 0 || (bank && !aligned16(bank)) || !aligned16(xq) || (num_out && !aligned16(num_out)))
    return SDN_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  g_prof.reset();
  int chosen = pick_path(path, Q, N, D, planes);
  if (!num_out && chosen == SDN_PATH_STREAM) chosen = SDN_PATH_GENERIC;   // z only: phase A kernels only
  if (!num_out && chosen == SDN_PATH_FLASH) chosen = SDN_PATH_UMMA;
  switch (chosen) {
    case SDN_PATH_FLASH: {
      if (!flash_supported(Q, N, D, planes)) return SDN_E_UNSUPPORTED;
      const int rc = flash_run(planes, sqnorm, N, D, xq, Q, inv_two_sigma_sq, dist_power, bank_alpha, num_out, z_out,
                               k_out, nullptr, st);
      // the grid cannot be co-resident on this device (MIG slice, another context): the two-phase kernels can
      if (rc != SDN_E_UNSUPPORTED || path == SDN_PATH_FLASH || !umma_supported(Q, N, D, planes)) return rc;
      return umma_partial(planes, sqnorm, N, D, xq, xsq, Q, inv_two_sigma_sq, dist_power, bank_alpha,
                          num_out, z_out, k_out, workspace, workspace_bytes, st, false);
    }
    case SDN_PATH_STREAM:
      if (!bank) return SDN_E_NULL;
      if (!stream_supported(Q, N, D)) return SDN_E_UNSUPPORTED;
      return stream_partial(bank, sqnorm, N, D, xq, xsq, Q, inv_two_sigma_sq, dist_power, bank_alpha,
                            num_out, z_out, k_out, workspace, workspace_bytes, st);
    case SDN_PATH_UMMA:
    case SDN_PATH_UMMA_BF16:
      if (!umma_supported(Q, N, D, planes)) return SDN_E_UNSUPPORTED;
      return umma_partial(planes, sqnorm, N, D, xq, xsq, Q, inv_two_sigma_sq, dist_power, bank_alpha,
                          num_out, z_out, k_out, workspace, workspace_bytes, st, chosen == SDN_PATH_UMMA_BF16);
    case SDN_PATH_GENERIC: break;
    default: return SDN_E_PARAM;
  }
  if (!bank || !xsq) return SDN_E_NULL;   // the generic kernels need ||xq||^2 from sdn_query_prepare
  float* S = k_out;
  if (!S) {
    if (!workspace || workspace_bytes < generic_workspace_bytes(Q, N)) return SDN_E_WORKSPACE;
    S = static_cast<float*>(workspace);
  }
  int pid = g_prof.begin("k_dots", st);
  int rc = generic_dots(bank, N, D, xq, Q, S, st);
  g_prof.end(pid, st);
  if (rc) return rc;
  pid = g_prof.begin("k_weights", st);
  rc = generic_weights(S, sqnorm, xsq, Q, N, inv_two_sigma_sq, dist_power, bank_alpha, z_out, st);
  g_prof.end(pid, st);
  if (rc) return rc;
  if (!num_out) return SDN_OK;
  pid = g_prof.begin("k_accum", st);
  rc = generic_accum(bank, N, D, S, Q, num_out, st);
  g_prof.end(pid, st);
  return rc;
}

int sdn_conditioning_fused(const float* bank, const float* sqnorm, const void* planes, int64_t N, int64_t D,
                           float* x0_inout, int64_t Q, float inv_two_sigma_sq, int32_t dist_power, float bank_alpha,
                           float eps, float scale, float gate_threshold, int32_t flags, float* num_out, float* z_out,
                           float* neg_out, float* denom_out, int32_t* gate_out, float* mean_out, float* k_out,
                           void* workspace, size_t workspace_bytes, int32_t path, void* stream) {
  if (!sqnorm || !x0_inout || (!bank && !planes)) return SDN_E_NULL;
  if (path != SDN_PATH_AUTO && path != SDN_PATH_STREAM && path != SDN_PATH_UMMA && path != SDN_PATH_FLASH) return SDN_E_PARAM;
  if (Q <= 0 || N <= 0 || D <= 0) return SDN_E_SHAPE;
  if (dist_power != 1 && dist_power != 2) return SDN_E_PARAM;
  if (D % 4 != 0 || (bank && !aligned16(bank)) || !aligned16(x0_inout) || (num_out && !aligned16(num_out)) ||
      (neg_out && !aligned16(neg_out)))
    return SDN_E_ALIGN;
  g_prof.reset();
  cudaStream_t st = (cudaStream_t)stream;
  const bool want_batched = path == SDN_PATH_AUTO ? batched(Q, N) : path != SDN_PATH_STREAM;
  if (path == SDN_PATH_FLASH && !flash_supported(Q, N, D, planes)) return SDN_E_UNSUPPORTED;
  if (want_batched && pick_path(path, Q, N, D, planes) == SDN_PATH_FLASH) {
    FlashEpi e{};
    e.fused = 1; e.eps = eps; e.scale = scale; e.gate_thr = gate_threshold; e.flags = flags;
    e.x0 = x0_inout; e.neg_out = neg_out; e.denom_out = denom_out; e.gate_out = gate_out; e.mean_out = mean_out;
    e.inv_qd = 1.f / (float)(Q * D);
    const int rc = flash_run(planes, sqnorm, N, D, x0_inout, Q, inv_two_sigma_sq, dist_power, bank_alpha, num_out, z_out,
                             k_out, &e, st);
    if (rc != SDN_E_UNSUPPORTED || path == SDN_PATH_FLASH) return rc;
  }
  if (want_batched && planes && z_out && umma_supported(Q, N, D, planes)) {
    if (!workspace || workspace_bytes < umma_workspace_bytes(Q, N, D)) return SDN_E_WORKSPACE;
    return umma_conditioning(planes, sqnorm, N, D, x0_inout, Q, inv_two_sigma_sq, dist_power, bank_alpha, eps, scale,
                             gate_threshold, flags, num_out, z_out, neg_out, denom_out, gate_out, mean_out, k_out,
                             workspace, workspace_bytes, st);
  }
  if (!bank || !stream_supported(Q, N, D)) return SDN_E_UNSUPPORTED;
  return stream_conditioning(bank, sqnorm, N, D, x0_inout, x0_inout, Q, inv_two_sigma_sq, dist_power, bank_alpha, eps,
                             scale, gate_threshold, flags, num_out, z_out, neg_out, denom_out, gate_out, mean_out,
                             k_out, workspace, workspace_bytes, st);
}

int sdn_sparse_repel(const float* bank, const float* sqnorm, int64_t N, int64_t D, float* x0_inout,
                     const float* xq, const float* xsq, int64_t Q, float radius, float scale, float* term_out,
                     float* wsum_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!bank || !sqnorm || !x0_inout || !xsq || !wsum_out || !workspace) return SDN_E_NULL;
  if (Q <= 0 || N <= 0 || D <= 0 || Q > 65535) return SDN_E_SHAPE;
  if (D % 4 != 0 || !aligned16(bank) || !aligned16(x0_inout) || (term_out && !aligned16(term_out)))
    return SDN_E_ALIGN;
  const size_t s_bytes = generic_workspace_bytes(Q, N);
  const size_t need = s_bytes + align_up(sizeof(float) * Q * D, 256);
  if (workspace_bytes < need) return SDN_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* S = static_cast<float*>(workspace);
  float* num = reinterpret_cast<float*>(static_cast<char*>(workspace) + s_bytes);
  const float* query = xq ? xq : x0_inout;
  if (!aligned16(query)) return SDN_E_ALIGN;
  // Q <= 8 on a shape the one-pass cluster kernel takes: the SPELL weight is one more functor of k_stream -- the bank
  // is read once (the generic kernels below read it twice and keep a [Q,N] scratch)
  {
    const size_t sws = align_up(stream_workspace_bytes(Q, N, D), 256);
    if (stream_supported(Q, N, D) && workspace_bytes >= sws + align_up(sizeof(float) * Q * D, 256)) {
      float* num2 = reinterpret_cast<float*>(static_cast<char*>(workspace) + sws);
      const int rc2 = stream_partial(bank, sqnorm, N, D, query, xsq, Q, radius, kPowerSparse, 1.f, num2, wsum_out, nullptr,
                                     workspace, sws, st);
      if (rc2) return rc2;
      return sparse_apply(num2, wsum_out, Q, D, scale, query, x0_inout, term_out, st);
    }
  }
  int rc = generic_dots(bank, N, D, query, Q, S, st);
  if (rc) return rc;
  rc = sparse_weights(S, sqnorm, xsq, Q, N, radius, wsum_out, st);
  if (rc) return rc;
  rc = generic_accum(bank, N, D, S, Q, num, st);
  if (rc) return rc;
  return sparse_apply(num, wsum_out, Q, D, scale, query, x0_inout, term_out, st);
}

int sdn_sparse_partial(const float* bank, const float* sqnorm, int64_t N, int64_t D, const float* xq, const float* xsq,
                       int64_t Q, float radius, float* num_out, float* wsum_out, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (!bank || !sqnorm || !xq || !xsq || !num_out || !wsum_out || !workspace) return SDN_E_NULL;
  if (Q <= 0 || N <= 0 || D <= 0 || Q > 65535) return SDN_E_SHAPE;
  if (D % 4 != 0 || !aligned16(bank) || !aligned16(xq) || !aligned16(num_out)) return SDN_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  if (stream_supported(Q, N, D) && workspace_bytes >= stream_workspace_bytes(Q, N, D))
    return stream_partial(bank, sqnorm, N, D, xq, xsq, Q, radius, kPowerSparse, 1.f, num_out, wsum_out, nullptr, workspace,
                          workspace_bytes, st);
  if (workspace_bytes < generic_workspace_bytes(Q, N)) return SDN_E_WORKSPACE;
  float* S = static_cast<float*>(workspace);
  int rc = generic_dots(bank, N, D, xq, Q, S, st);
  if (rc) return rc;
  rc = sparse_weights(S, sqnorm, xsq, Q, N, radius, wsum_out, st);
  if (rc) return rc;
  return generic_accum(bank, N, D, S, Q, num_out, st);
}

int sdn_sparse_partial_planes(const void* planes, const float* sqnorm, int64_t N, int64_t D, const float* xq,
                              const float* xsq, int64_t Q, float radius, float* num_out, float* wsum_out,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (!planes || !sqnorm || !xq || !xsq || !num_out || !wsum_out || !workspace) return SDN_E_NULL;
  if (Q <= 0 || N <= 0 || D <= 0 || Q > 65535) return SDN_E_SHAPE;
  if (D % 4 != 0 || !aligned16(xq) || !aligned16(num_out)) return SDN_E_ALIGN;
  if (!umma_supported(Q, N, D, planes)) return SDN_E_UNSUPPORTED;
  if (workspace_bytes < umma_workspace_bytes(Q, N, D)) return SDN_E_WORKSPACE;
  return umma_partial(planes, sqnorm, N, D, xq, xsq, Q, radius, kPowerSparse, 1.f, num_out, wsum_out, nullptr, workspace,
                      workspace_bytes, (cudaStream_t)stream, false);
}

int sdn_sparse_apply(const float* num, const float* wsum, int64_t Q, int64_t D, float scale, const float* xq,
                     float* x0_inout, float* term_out, void* stream) {
  if (!num || !wsum || !xq || !x0_inout) return SDN_E_NULL;
  if (Q <= 0 || D <= 0) return SDN_E_SHAPE;
  if (D % 4 != 0 || !aligned16(num) || !aligned16(xq) || !aligned16(x0_inout) || (term_out && !aligned16(term_out)))
    return SDN_E_ALIGN;
  return sparse_apply(num, wsum, Q, D, scale, xq, x0_inout, term_out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ host-buffer path (e2e)
namespace {
// One replayable CUDA graph per (host buffers, bank, shape, scalars): H2D copy + the fused conditioning sequence + the
// two D2H copies as ONE launch.  A sampler calls with the same pinned buffers every step, so from the third call on the
// host pays one cudaGraphLaunch instead of two copy calls and four kernel launches, and the GPU runs the nodes without
// launch gaps (small queries only, see the call site).  First call of a key: eager (one-time set-up of the
// kernels must not happen under capture); second call: capture + instantiate; SDN_HOST_GRAPH=0 switches it off.
struct HostGraphKey {
  int device; const void* x0_host; const void* denom_host; const void* bank; const void* sqnorm; const void* planes;
  int64_t N, D, Q; float inv2s2; int power; float alpha, eps, scale;
  bool operator==(const HostGraphKey& o) const {
    return device == o.device && x0_host == o.x0_host && denom_host == o.denom_host && bank == o.bank && sqnorm == o.sqnorm &&
           planes == o.planes && N == o.N && D == o.D && Q == o.Q && inv2s2 == o.inv2s2 && power == o.power && alpha == o.alpha &&
           eps == o.eps && scale == o.scale;
  }
};
struct HostGraph { HostGraphKey key{}; int seen = 0; cudaGraphExec_t exec = nullptr; uint64_t stamp = 0; };

struct HostCache {
  std::mutex mu;
  void* dev = nullptr;
  size_t dev_bytes = 0;
  int device = -1;      // the device `dev` was allocated on
  HostGraph graphs[4];
  uint64_t clock = 0;
  cudaStream_t gstream = nullptr;   // capture / replay stream (the caller's may be the legacy default stream)
  int gstream_device = -1;
  void drop_graphs() {
    for (HostGraph& g : graphs) {
      if (g.exec) cudaGraphExecDestroy(g.exec);
      g = HostGraph{};
    }
  }
} g_host;

bool pinned_host(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
}  // namespace

void sdn_host_release(void) {
  std::lock_guard<std::mutex> lk(g_host.mu);
  g_host.drop_graphs();
  if (g_host.gstream) cudaStreamDestroy(g_host.gstream);
  g_host.gstream = nullptr;
  g_host.gstream_device = -1;
  if (g_host.dev) cudaFree(g_host.dev);
  g_host.dev = nullptr;
  g_host.dev_bytes = 0;
}

int sdn_conditioning_host(const float* bank, const float* sqnorm, const void* planes, int64_t N,
                          int64_t D, float* x0_host, int64_t Q, int32_t normalize_C,
                          float inv_two_sigma_sq, int32_t dist_power, float bank_alpha, float eps,
                          float scale, float* denom_host, int32_t path, void* stream) {
  if (!x0_host || !denom_host) return SDN_E_NULL;
  if (Q <= 0 || N <= 0 || D <= 0) return SDN_E_SHAPE;
  std::lock_guard<std::mutex> lk(g_host.mu);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t qd = align_up(sizeof(float) * Q * D, 256);
  const size_t qv = align_up(sizeof(float) * Q, 256);
  const size_t ws = align_up(sdn_repel_workspace_bytes(Q, N, D, path), 256);
  // layout: x0 | xq | num | xsq | z | denom | workspace
  const size_t need = 3 * qd + 3 * qv + ws;
  int cur_dev = 0;
  SDN_CUDA_OK(cudaGetDevice(&cur_dev));
  if (need > g_host.dev_bytes || cur_dev != g_host.device) {
    g_host.drop_graphs();             // they address the old staging buffers
    if (g_host.dev) SDN_CUDA_OK(cudaFree(g_host.dev));
    g_host.dev = nullptr;
    g_host.dev_bytes = 0;
    SDN_CUDA_OK(cudaMalloc(&g_host.dev, need));
    g_host.dev_bytes = need;
    g_host.device = cur_dev;
  }
  char* p = static_cast<char*>(g_host.dev);
  float* x0 = reinterpret_cast<float*>(p);
  float* xq = reinterpret_cast<float*>(p + qd);
  float* num = reinterpret_cast<float*>(p + 2 * qd);
  float* xsq = reinterpret_cast<float*>(p + 3 * qd);
  float* z = reinterpret_cast<float*>(p + 3 * qd + qv);
  float* denom = reinterpret_cast<float*>(p + 3 * qd + 2 * qv);
  void* wsp = p + 3 * qd + 3 * qv;

  // ---- replay path: the whole call as one CUDA graph (plain query, AUTO path, pinned host buffers, no profiling)
  static const bool host_graph = [] { const char* e = getenv("SDN_HOST_GRAPH"); return !(e && atoi(e) == 0); }();
  // Only for small queries (<= 512 KiB: the Q <= 8 shapes the reference runs).  Measured: cfg1 (64 KiB) 12.7 k -> 17.1 k
  // calls/s; cfg3 (4 MiB) 190.7 k -> 184.1 k proj/s -- with eager launches the host enqueues the kernels WHILE the 86 us
  // H2D copy runs, a graph pays its launch latency before the copy starts.
  if (host_graph && normalize_C == 0 && path == SDN_PATH_AUTO && !g_prof.enabled &&
      sizeof(float) * (size_t)Q * (size_t)D <= (512u << 10) && pinned_host(x0_host) && pinned_host(denom_host)) {
    const HostGraphKey key{cur_dev, x0_host, denom_host, bank, sqnorm, planes, N, D, Q, inv_two_sigma_sq, dist_power,
                           bank_alpha, eps, scale};
    HostGraph* slot = nullptr;
    for (HostGraph& g : g_host.graphs)
      if (g.stamp != 0 && g.key == key) slot = &g;
    if (!slot) {                      // new key: least recently used slot, this call runs eagerly below
      slot = &g_host.graphs[0];
      for (HostGraph& g : g_host.graphs)
        if (g.stamp < slot->stamp) slot = &g;
      if (slot->exec) cudaGraphExecDestroy(slot->exec);
      *slot = HostGraph{};
      slot->key = key;
    }
    slot->stamp = ++g_host.clock;
    if (slot->seen >= 0) ++slot->seen;          // < 0: this key cannot be captured, stay eager
    if (slot->seen >= 2) {
      if (!g_host.gstream || g_host.gstream_device != cur_dev) {
        if (g_host.gstream) cudaStreamDestroy(g_host.gstream);
        SDN_CUDA_OK(cudaStreamCreateWithFlags(&g_host.gstream, cudaStreamNonBlocking));
        g_host.gstream_device = cur_dev;
      }
      cudaStream_t gs = g_host.gstream;
      if (!slot->exec) {
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          ok = cudaMemcpyAsync(x0, x0_host, sizeof(float) * Q * D, cudaMemcpyHostToDevice, gs) == cudaSuccess;
          ok = ok && sdn_conditioning_fused(bank, sqnorm, planes, N, D, x0, Q, inv_two_sigma_sq, dist_power, bank_alpha, eps,
                                            scale, 0.f, 0, nullptr, z, nullptr, denom, nullptr, nullptr, nullptr, wsp, ws,
                                            SDN_PATH_AUTO, gs) == SDN_OK;
          ok = ok && cudaMemcpyAsync(x0_host, x0, sizeof(float) * Q * D, cudaMemcpyDeviceToHost, gs) == cudaSuccess;
          ok = ok && cudaMemcpyAsync(denom_host, denom, sizeof(float) * Q, cudaMemcpyDeviceToHost, gs) == cudaSuccess;
          const bool ended = cudaStreamEndCapture(gs, &graph) == cudaSuccess;
          ok = ok && ended && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&slot->exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {                    // not capturable here (e.g. a shape the fused sequence does not take): stay eager
          cudaGetLastError();
          slot->exec = nullptr;
          slot->seen = -1;
        }
      }
      if (slot->exec) {
        // order after the caller's stream, then run and wait: the call is synchronous
        SDN_CUDA_OK(cudaStreamSynchronize(st));
        SDN_CUDA_OK(cudaGraphLaunch(slot->exec, gs));
        SDN_CUDA_OK(cudaStreamSynchronize(gs));
        return SDN_OK;
      }
    }
  }

  SDN_CUDA_OK(cudaMemcpyAsync(x0, x0_host, sizeof(float) * Q * D, cudaMemcpyHostToDevice, st));
  int rc = SDN_E_UNSUPPORTED;
  if (normalize_C == 0 && path == SDN_PATH_AUTO)      // plain query: the few-launch sequence when the shape allows
    rc = sdn_conditioning_fused(bank, sqnorm, planes, N, D, x0, Q, inv_two_sigma_sq, dist_power, bank_alpha, eps, scale,
                                0.f, 0, nullptr, z, nullptr, denom, nullptr, nullptr, nullptr, wsp, ws, SDN_PATH_AUTO, stream);
  if (rc == SDN_E_UNSUPPORTED) {
    rc = sdn_query_prepare(x0, nullptr, 1.f, 0.f, Q, D, normalize_C, nullptr, normalize_C > 0 ? xq : nullptr, xsq,
                           stream);
    if (rc) return rc;
    const float* query = normalize_C > 0 ? xq : x0;
    rc = sdn_repel_partial(bank, sqnorm, planes, N, D, query, xsq, Q, inv_two_sigma_sq, dist_power, bank_alpha, num, z,
                           nullptr, wsp, ws, path, stream);
    if (rc) return rc;
    rc = sdn_epilogue_correct(num, z, Q, D, eps, scale, 0.f, 0, x0, nullptr, denom, nullptr, nullptr, stream);
  }
  if (rc) return rc;
  SDN_CUDA_OK(cudaMemcpyAsync(x0_host, x0, sizeof(float) * Q * D, cudaMemcpyDeviceToHost, st));
  SDN_CUDA_OK(cudaMemcpyAsync(denom_host, denom, sizeof(float) * Q, cudaMemcpyDeviceToHost, st));
  SDN_CUDA_OK(cudaStreamSynchronize(st));
  return SDN_OK;
}

// ------------------------------------------------------------------ host-buffer calls in flight (throughput)
// A serving loop has independent requests: while one's corrected query travels back over PCIe the next one's query can
// travel in and a third can be on the SMs.  A pipe owns `slots` staging areas, each with its own stream; submit enqueues
// H2D copy -> fused conditioning -> D2H copies on the slot's stream and returns, wait blocks until that slot's results
// are in the host buffers.  Work on one slot is stream-ordered, so re-submitting a slot without waiting is safe for the
// device buffers (the caller's host output is overwritten, of course).
namespace {
struct HostPipe {
  int device = 0;
  int64_t Q = 0, N = 0, D = 0;
  int nslots = 0;
  size_t qd = 0, qv = 0, ws = 0, slot_bytes = 0;
  char* dev = nullptr;
  cudaStream_t streams[8] = {};
};
}  // namespace

int sdn_host_pipe_create(int64_t Q, int64_t N, int64_t D, int32_t slots, void** pipe_out) {
  if (!pipe_out) return SDN_E_NULL;
  if (Q <= 0 || N <= 0 || D <= 0 || slots < 1 || slots > 8) return SDN_E_SHAPE;
  HostPipe* hp = new HostPipe();
  SDN_CUDA_OK(cudaGetDevice(&hp->device));
  hp->Q = Q; hp->N = N; hp->D = D; hp->nslots = slots;
  hp->qd = align_up(sizeof(float) * Q * D, 256);
  hp->qv = align_up(sizeof(float) * Q, 256);
  hp->ws = align_up(sdn_repel_workspace_bytes(Q, N, D, SDN_PATH_AUTO), 256);
  hp->slot_bytes = hp->qd + 2 * hp->qv + hp->ws;      // x0 | z | denom | workspace
  cudaError_t e = cudaMalloc(&hp->dev, hp->slot_bytes * slots);
  for (int i = 0; e == cudaSuccess && i < slots; ++i) e = cudaStreamCreateWithFlags(&hp->streams[i], cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    sdn_host_pipe_destroy(hp);
    return (int)e;
  }
  *pipe_out = hp;
  return SDN_OK;
}

int sdn_host_pipe_submit(void* pipe, int32_t slot, const float* bank, const float* sqnorm, const void* planes,
                         const float* x0_in_host, float* x0_out_host, float* denom_host, float inv_two_sigma_sq,
                         int32_t dist_power, float bank_alpha, float eps, float scale) {
  HostPipe* hp = static_cast<HostPipe*>(pipe);
  if (!hp || !x0_in_host || !x0_out_host || !denom_host) return SDN_E_NULL;
  if (slot < 0 || slot >= hp->nslots) return SDN_E_PARAM;
  int cur_dev = 0;
  SDN_CUDA_OK(cudaGetDevice(&cur_dev));
  if (cur_dev != hp->device) return SDN_E_DEVICE;
  char* p = hp->dev + (size_t)slot * hp->slot_bytes;
  float* x0 = reinterpret_cast<float*>(p);
  float* z = reinterpret_cast<float*>(p + hp->qd);
  float* denom = reinterpret_cast<float*>(p + hp->qd + hp->qv);
  void* wsp = p + hp->qd + 2 * hp->qv;
  cudaStream_t st = hp->streams[slot];
  const int64_t Q = hp->Q, D = hp->D;
  SDN_CUDA_OK(cudaMemcpyAsync(x0, x0_in_host, sizeof(float) * Q * D, cudaMemcpyHostToDevice, st));
  const int rc = sdn_conditioning_fused(bank, sqnorm, planes, hp->N, D, x0, Q, inv_two_sigma_sq, dist_power, bank_alpha, eps,
                                        scale, 0.f, 0, nullptr, z, nullptr, denom, nullptr, nullptr, nullptr, wsp, hp->ws,
                                        SDN_PATH_AUTO, st);
  if (rc) return rc;          // SDN_E_UNSUPPORTED: a shape without a fused sequence (use sdn_conditioning_host)
  SDN_CUDA_OK(cudaMemcpyAsync(x0_out_host, x0, sizeof(float) * Q * D, cudaMemcpyDeviceToHost, st));
  SDN_CUDA_OK(cudaMemcpyAsync(denom_host, denom, sizeof(float) * Q, cudaMemcpyDeviceToHost, st));
  return SDN_OK;
}

int sdn_host_pipe_wait(void* pipe, int32_t slot) {
  HostPipe* hp = static_cast<HostPipe*>(pipe);
  if (!hp) return SDN_E_NULL;
  if (slot < 0 || slot >= hp->nslots) return SDN_E_PARAM;
  SDN_CUDA_OK(cudaStreamSynchronize(hp->streams[slot]));
  return SDN_OK;
}

void sdn_host_pipe_destroy(void* pipe) {
  HostPipe* hp = static_cast<HostPipe*>(pipe);
  if (!hp) return;
  for (int i = 0; i < 8; ++i)
    if (hp->streams[i]) { cudaStreamSynchronize(hp->streams[i]); cudaStreamDestroy(hp->streams[i]); }
  if (hp->dev) cudaFree(hp->dev);
  delete hp;
}

}  // extern "C"
