// K5: TF1 optimizer update rules, one launch for all parameter tensors of a
// candidate (multi-tensor apply).  HBM-bound elementwise work: float4 where
// the tensors are 16B-aligned, grid sized to the work.
//
// Replaces the apply half of optimizer.minimize at
//   adanet/examples/simple_dnn.py:110 and adanet/ensemble/weighted.py:616.
// Update rules are TensorFlow 1.x's (not vendored by the reference):
//   GradientDescent  v -= lr*g
//   Momentum         acc = m*acc + g ; v -= lr*acc
//   RMSProp          ms = rho*ms + (1-rho) g^2 ; mom = mu*mom + lr*g/sqrt(ms+eps) ; v -= mom   (ms init 1)
//   Adam             lr_t = lr*sqrt(1-b2^t)/(1-b1^t) ; m += (1-b1)(g-m) ; v += (1-b2)(g^2-v) ;
//                    var -= lr_t*m/(sqrt(v)+eps)
//   Momentum+cosine  Momentum with lr_t = tf.train.cosine_decay(lr, step, decay_steps, alpha) read from the
//                    device step counter (customizing_adanet.ipynb SimpleCNNBuilder.build_subnetwork_train_op)
#include "common.cuh"
#include "plane_fmt.cuh"

namespace adn {

static constexpr int kMaxTensors = 32;
static constexpr int kChunk = 4096;  // elements per CTA

struct OptParams {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* s0[kMaxTensors];
  float* s1[kMaxTensors];
  int chunk_start[kMaxTensors + 1];  // prefix sum of chunks per tensor
  int64_t size[kMaxTensors];
  int n;
  int kind;
  float h0, h1, h2, h3;
  const int64_t* step_dev;
  // optional split planes of 2-D parameters, refreshed with the update (plane_fmt.cuh layout)
  void* plane_hi[kMaxTensors];    // hi plane base (nullable per tensor)
  void* plane_lo[kMaxTensors];
  int cols[kMaxTensors];
  int fmt;
  unsigned int* ovf;
};

__device__ __forceinline__ void store_planes(const OptParams& o, int t, int64_t j, float v) {
  if (!o.plane_hi[t]) return;
  const int cols = o.cols[t];
  const int64_t rows = o.size[t] / cols;
  const int64_t r = j / cols;
  const int c = (int)(j - r * cols);
  pl::plane_store(pl::PlaneView{o.plane_hi[t], o.plane_lo[t], nullptr, rows, o.fmt}, r, c, v, o.ovf);
}

__device__ __forceinline__ void apply_one(int kind, float& p, float g, float& s0, float& s1, float h0, float h1,
                                          float h2, float h3, float lr_t) {
  switch (kind) {
    case ADN_OPT_SGD:
      p -= h0 * g;
      break;
    case ADN_OPT_MOMENTUM:
      s0 = h1 * s0 + g;
      p -= h0 * s0;
      break;
    case ADN_OPT_MOMENTUM_COSINE:
      s0 = h1 * s0 + g;
      p -= lr_t * s0;
      break;
    case ADN_OPT_RMSPROP:
      s0 = h1 * s0 + (1.f - h1) * g * g;
      s1 = h2 * s1 + h0 * g / sqrtf(s0 + h3);
      p -= s1;
      break;
    default:  // ADAM
      s0 += (1.f - h1) * (g - s0);
      s1 += (1.f - h2) * (g * g - s1);
      p -= lr_t * s0 / (sqrtf(s1) + h3);
      break;
  }
}

__global__ void __launch_bounds__(256) opt_step_kernel(const __grid_constant__ OptParams o) {
  // locate tensor for this CTA (n <= 32: linear scan)
  int t = 0;
  while (t + 1 < o.n && (int)blockIdx.x >= o.chunk_start[t + 1]) ++t;
  const int64_t off = (int64_t)(blockIdx.x - o.chunk_start[t]) * kChunk;
  const int64_t end = min(o.size[t], off + kChunk);
  float* p = o.p[t];
  const float* g = o.g[t];
  float* s0 = o.s0[t];
  float* s1 = o.s1[t];
  float lr_t = o.h0;
  if (o.kind == ADN_OPT_ADAM) {
    const float tt = (float)(*o.step_dev + 1);
    lr_t = o.h0 * sqrtf(1.f - powf(o.h2, tt)) / (1.f - powf(o.h1, tt));
  } else if (o.kind == ADN_OPT_MOMENTUM_COSINE) {
    // tf.train.cosine_decay [TF]: step clipped to decay_steps, fp32 arithmetic
    const float st = fminf((float)(*o.step_dev), o.h2);
    const float cosine = 0.5f * (1.f + cosf(3.14159265358979323846f * (st / o.h2)));
    lr_t = o.h0 * ((1.f - o.h3) * cosine + o.h3);
  }
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)s0 | (uintptr_t)s1) & 15) == 0;
  if (vec) {
    for (int64_t i = off + threadIdx.x * 4; i < end; i += 256 * 4) {
      if (i + 3 < end) {
        float4 pv = *reinterpret_cast<float4*>(p + i);
        float4 gv = __ldg(reinterpret_cast<const float4*>(g + i));
        float4 a = s0 ? *reinterpret_cast<float4*>(s0 + i) : make_float4(0, 0, 0, 0);
        float4 b = s1 ? *reinterpret_cast<float4*>(s1 + i) : make_float4(0, 0, 0, 0);
        apply_one(o.kind, pv.x, gv.x, a.x, b.x, o.h0, o.h1, o.h2, o.h3, lr_t);
        apply_one(o.kind, pv.y, gv.y, a.y, b.y, o.h0, o.h1, o.h2, o.h3, lr_t);
        apply_one(o.kind, pv.z, gv.z, a.z, b.z, o.h0, o.h1, o.h2, o.h3, lr_t);
        apply_one(o.kind, pv.w, gv.w, a.w, b.w, o.h0, o.h1, o.h2, o.h3, lr_t);
        *reinterpret_cast<float4*>(p + i) = pv;
        store_planes(o, t, i, pv.x); store_planes(o, t, i + 1, pv.y);
        store_planes(o, t, i + 2, pv.z); store_planes(o, t, i + 3, pv.w);
        if (s0) *reinterpret_cast<float4*>(s0 + i) = a;
        if (s1) *reinterpret_cast<float4*>(s1 + i) = b;
      } else {
        for (int64_t j = i; j < end; ++j) {
          float pv = p[j], a = s0 ? s0[j] : 0.f, b = s1 ? s1[j] : 0.f;
          apply_one(o.kind, pv, g[j], a, b, o.h0, o.h1, o.h2, o.h3, lr_t);
          p[j] = pv;
          store_planes(o, t, j, pv);
          if (s0) s0[j] = a;
          if (s1) s1[j] = b;
        }
      }
    }
  } else {
    for (int64_t j = off + threadIdx.x; j < end; j += 256) {
      float pv = p[j], a = s0 ? s0[j] : 0.f, b = s1 ? s1[j] : 0.f;
      apply_one(o.kind, pv, g[j], a, b, o.h0, o.h1, o.h2, o.h3, lr_t);
      p[j] = pv;
      store_planes(o, t, j, pv);
      if (s0) s0[j] = a;
      if (s1) s1[j] = b;
    }
  }
}

__global__ void step_increment_kernel(int64_t* step) { *step += 1; }

}  // namespace adn

using namespace adn;


extern "C" int adn_opt_step(int kind, float* const* params_host, const float* const* grads_host,
                            float* const* slot0_host, float* const* slot1_host, const int64_t* sizes_host,
                            int n_tensors, const float* hyper_host, int64_t* step_dev, void* stream) {
  return adn_opt_step_p(kind, params_host, grads_host, slot0_host, slot1_host, sizes_host, n_tensors, hyper_host,
                        step_dev, nullptr, nullptr, stream);
}

extern "C" int adn_opt_step_p(int kind, float* const* params_host, const float* const* grads_host,
                              float* const* slot0_host, float* const* slot1_host, const int64_t* sizes_host,
                              int n_tensors, const float* hyper_host, int64_t* step_dev,
                              void* const* planes_host, const int64_t* cols_host, void* stream) {
  if (kind < ADN_OPT_SGD || kind > ADN_OPT_MOMENTUM_COSINE) return fail(ADN_ERR_INVALID, "adn_opt_step: bad kind %d", kind);
  if (n_tensors < 1 || n_tensors > kMaxTensors)
    return fail(ADN_ERR_UNSUPPORTED, "adn_opt_step: n_tensors %d not in [1,%d]", n_tensors, kMaxTensors);
  if (!params_host || !grads_host || !sizes_host || !hyper_host)
    return fail(ADN_ERR_INVALID, "adn_opt_step: null pointer");
  const int need_slots = kind == ADN_OPT_SGD ? 0 : ((kind == ADN_OPT_MOMENTUM || kind == ADN_OPT_MOMENTUM_COSINE) ? 1 : 2);
  if (need_slots >= 1 && !slot0_host) return fail(ADN_ERR_INVALID, "adn_opt_step: slot0 required");
  if (need_slots >= 2 && !slot1_host) return fail(ADN_ERR_INVALID, "adn_opt_step: slot1 required");
  if ((kind == ADN_OPT_ADAM || kind == ADN_OPT_MOMENTUM_COSINE) && !step_dev)
    return fail(ADN_ERR_INVALID, "adn_opt_step: Adam / cosine-decay Momentum need step_dev");
  if (kind == ADN_OPT_MOMENTUM_COSINE && !(hyper_host[2] > 0.f))
    return fail(ADN_ERR_INVALID, "adn_opt_step: cosine decay needs decay_steps > 0");
  OptParams o{};
  int chunks = 0;
  for (int t = 0; t < n_tensors; ++t) {
    if (!params_host[t] || !grads_host[t] || sizes_host[t] <= 0)
      return fail(ADN_ERR_INVALID, "adn_opt_step: tensor %d null or empty", t);
    o.p[t] = params_host[t];
    o.g[t] = grads_host[t];
    o.s0[t] = need_slots >= 1 ? slot0_host[t] : nullptr;
    o.s1[t] = need_slots >= 2 ? slot1_host[t] : nullptr;
    if ((need_slots >= 1 && !o.s0[t]) || (need_slots >= 2 && !o.s1[t]))
      return fail(ADN_ERR_INVALID, "adn_opt_step: slot for tensor %d is null", t);
    o.size[t] = sizes_host[t];
    o.plane_hi[t] = nullptr;
    if (planes_host && planes_host[t]) {
      if (!cols_host || cols_host[t] <= 0 || sizes_host[t] % cols_host[t] != 0 || cols_host[t] > INT32_MAX)
        return fail(ADN_ERR_INVALID, "adn_opt_step_p: tensor %d: cols must divide its size", t);
      const pl::PlaneView v = pl::plane_view(pl::format(), planes_host[t], sizes_host[t] / cols_host[t], cols_host[t]);
      o.plane_hi[t] = v.hi;
      o.plane_lo[t] = v.lo;
      o.cols[t] = (int)cols_host[t];
    }
    o.chunk_start[t] = chunks;
    chunks += (int)ceil_div(sizes_host[t], kChunk);
  }
  o.chunk_start[n_tensors] = chunks;
  o.n = n_tensors;
  o.kind = kind;
  o.h0 = hyper_host[0];
  o.h1 = kind >= ADN_OPT_MOMENTUM ? hyper_host[1] : 0.f;
  o.h2 = kind >= ADN_OPT_RMSPROP ? hyper_host[2] : 0.f;
  o.h3 = kind >= ADN_OPT_RMSPROP ? hyper_host[3] : 0.f;   // MOMENTUM_COSINE (4): {lr, momentum, decay_steps, alpha}
  o.step_dev = step_dev;
  o.fmt = pl::format();
  o.ovf = pl::overflow_flag();
  opt_step_kernel<<<chunks, 256, 0, as_stream(stream)>>>(o);
  ADN_CHECK_LAUNCH("opt_step");
  if (step_dev) {
    step_increment_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev);
    ADN_CHECK_LAUNCH("step_increment");
  }
  return ADN_OK;
}
