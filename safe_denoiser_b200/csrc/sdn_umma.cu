// tcgen05 / TMEM / TMA two-phase kernels for batched calls.  Filled in below the generic path.
#include "sdn_internal.h"

namespace sdn {
bool umma_supported(int64_t, int64_t, int64_t, const void*) { return false; }
size_t umma_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
int umma_partial(const void*, const float*, int64_t, int64_t, const float*, const float*, int64_t, float,
                 int, float, float*, float*, float*, void*, size_t, cudaStream_t) {
  return SDN_E_UNSUPPORTED;
}
}  // namespace sdn
