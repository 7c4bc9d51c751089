// One-pass cluster kernel for GEMV-shaped calls (small Q).  Filled in below the generic path.
#include "sdn_internal.h"

namespace sdn {
bool stream_supported(int64_t, int64_t, int64_t) { return false; }
size_t stream_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
int stream_partial(const float*, const float*, int64_t, int64_t, const float*, const float*, int64_t, float,
                   int, float, float*, float*, float*, void*, size_t, cudaStream_t) {
  return SDN_E_UNSUPPORTED;
}
}  // namespace sdn
