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
// exactly one of yp (planes out) / y (dense fp32 out) is non-null
int dense_fwd(const float* xp, const float* wp, const float* bias, float* yp, float* y, int64_t batch, int64_t in,
              int64_t out, int act, cudaStream_t st);
// dw nullable; at most one of dxp (planes) / dx (dense); dx_colsum[in] = column sums of dx (nullable)
int dense_bwd(const float* xp, const float* wp, const float* dzp, float* dxp, float* dx, float* dx_colsum, float* dw,
              int64_t batch, int64_t in, int64_t out, int x_relu_mask, void* ws, int64_t ws_bytes, cudaStream_t st);

}  // namespace pl
}  // namespace adn
