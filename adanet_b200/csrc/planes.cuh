// Plane-native tcgen05 dense pipeline (see planes.cu for the design).
#pragma once
#include "common.cuh"

namespace adn {
namespace pl {

int init();
// floats in ONE plane (hi or lo) of a [rows, cols] tensor; a plane tensor is hi followed by lo
int64_t plane_floats(int64_t rows, int64_t cols);
int64_t bits_words(int64_t rows, int64_t cols);   // uint32 words of the sign-bit block that follows the two planes
int64_t planes_bytes(int64_t rows, int64_t cols);
int64_t dense_bwd_workspace_bytes(int64_t batch, int64_t in, int64_t out);

int split(const float* src, int64_t rows, int64_t cols, float* planes, cudaStream_t st);
int merge(const float* planes, int64_t rows, int64_t cols, float* dst, cudaStream_t st);
// one dense layer of one subnetwork; groups = the same layer wave of several subnetworks in one launch
struct FwdOp {
  const float* xp;      // planes [batch, in]
  const float* wp;      // planes [in, out]
  const float* bias;    // [out] or null
  float* yp;            // planes [batch, out] (exactly one of yp / y)
  float* y;             // dense  [batch, out]
  int64_t in, out;
  int act;
};
struct BwdOp {
  const float* xp;      // planes [batch, in]
  const float* wp;      // planes [in, out] (needed when dx is requested)
  const float* dzp;     // planes [batch, out]
  float* dxp;           // planes [batch, in] or null
  float* dx;            // dense  [batch, in] or null (at most one of dxp / dx)
  float* dx_colsum;     // [in] or null
  float* dw;            // dense [in, out] or null
  int64_t in, out;
  int x_relu_mask;
  void* ws;             // dense_bwd_workspace_bytes(batch, in, out), one per op
  int64_t ws_bytes;
};
int dense_fwd_group(const FwdOp* ops, int n, int64_t batch, cudaStream_t st);
int dense_bwd_group(const BwdOp* ops, int n, int64_t batch, cudaStream_t st);
// exactly one of yp (planes out) / y (dense fp32 out) is non-null
int dense_fwd(const float* xp, const float* wp, const float* bias, float* yp, float* y, int64_t batch, int64_t in,
              int64_t out, int act, cudaStream_t st);
// dw nullable; at most one of dxp (planes) / dx (dense); dx_colsum[in] = column sums of dx (nullable)
int dense_bwd(const float* xp, const float* wp, const float* dzp, float* dxp, float* dx, float* dx_colsum, float* dw,
              int64_t batch, int64_t in, int64_t out, int x_relu_mask, void* ws, int64_t ws_bytes, cudaStream_t st);

}  // namespace pl
}  // namespace adn
