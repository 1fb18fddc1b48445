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
// dw[i] += coef * sign(w[i])   (gradient of coef * ||w||_1; TF: d|w|/dw = sign(w))
__global__ void __launch_bounds__(256) l1_grad_add_kernel(float* dw, const float* w, int64_t n, float coef) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = w[i];
    dw[i] += coef * ((v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f));
  }
}
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

extern "C" int adn_l1_grad_add(float* dw, const float* w, int64_t n, float coef, void* stream) {
  if (!dw || !w || n <= 0) return fail(ADN_ERR_INVALID, "adn_l1_grad_add: bad argument");
  const int blocks = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  l1_grad_add_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dw, w, n, coef);
  ADN_CHECK_LAUNCH("l1_grad_add");
  return ADN_OK;
}
