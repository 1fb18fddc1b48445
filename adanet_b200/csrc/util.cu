// Device-side step bookkeeping (graph-capturable): loss trace + counters.
// Replaces the per-spec step variables / hooks of adanet/core/iteration.py:150-205,961-996.
#include "common.cuh"

namespace adn {
struct RecordParams {
  const float* src[16];
  int n;
};
__global__ void record_scalars_kernel(const __grid_constant__ RecordParams r, float* trace, int64_t stride,
                                      const int64_t* step, int64_t capacity) {
  const int i = threadIdx.x;
  if (i < r.n) trace[(size_t)(*step % capacity) * stride + i] = *r.src[i];
}
__global__ void counter_add_kernel(int64_t* c, int64_t d) { *c += d; }
}  // namespace adn

using namespace adn;

extern "C" int adn_record_scalars(const float* const* src_host, int n, float* trace, int64_t stride,
                                  const int64_t* step_dev, int64_t capacity, void* stream) {
  if (!src_host || !trace || !step_dev || n < 1 || n > 16 || stride < n || capacity < 1)
    return fail(ADN_ERR_INVALID, "adn_record_scalars: bad argument");
  RecordParams r{};
  for (int i = 0; i < n; ++i) {
    if (!src_host[i]) return fail(ADN_ERR_INVALID, "adn_record_scalars: src %d null", i);
    r.src[i] = src_host[i];
  }
  r.n = n;
  record_scalars_kernel<<<1, 32, 0, as_stream(stream)>>>(r, trace, stride, step_dev, capacity);
  ADN_CHECK_LAUNCH("record_scalars");
  return ADN_OK;
}

extern "C" int adn_counter_add(int64_t* counter_dev, int64_t delta, void* stream) {
  if (!counter_dev) return fail(ADN_ERR_INVALID, "adn_counter_add: null pointer");
  counter_add_kernel<<<1, 1, 0, as_stream(stream)>>>(counter_dev, delta);
  ADN_CHECK_LAUNCH("counter_add");
  return ADN_OK;
}
