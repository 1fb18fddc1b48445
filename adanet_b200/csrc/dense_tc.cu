// placeholder until the tcgen05 kernels land (next commit): nothing is supported,
// so ADN_PATH_AUTO resolves to the SIMT fp32 path for every shape.
#include "dense_tc.cuh"
namespace adn {
namespace tc {
int init() { return ADN_OK; }
bool fwd_supported(int64_t, int64_t, int64_t) { return false; }
bool bwd_supported(int64_t, int64_t, int64_t) { return false; }
int64_t dense_bwd_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
int dense_fwd(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int, cudaStream_t) {
  return fail(ADN_ERR_UNSUPPORTED, "tcgen05 dense_fwd not built");
}
int dense_bwd(const float*, const float*, const float*, float*, float*, float*, int64_t, int64_t, int64_t, int,
              void*, int64_t, cudaStream_t) {
  return fail(ADN_ERR_UNSUPPORTED, "tcgen05 dense_bwd not built");
}
}  // namespace tc
}  // namespace adn
