// Local normalisation of arc weights (weight_fns.hat_normalize / log_softmax_normalize,
// /root/reference/last_torch/weight_fns.py:99-136) as one fused row-wise kernel each way:
// the epilogue a locally normalised model applies to the (blank [M], lexical [M, V]) output
// of its weight function.  One warp per row (1 + V values), a row is read once in the
// forward pass (two passes over registers / L1 for the log-sum-exp) and the gradient is
// recomputed from the INPUTS in the backward pass, so nothing but the outputs is stored.
//
//   LT_NORM_HAT          blank' = blank - softplus(blank)                      (log sigmoid)
//                        lex'_y = log_softmax(lex)_y - softplus(blank)
//   LT_NORM_LOG_SOFTMAX  (blank', lex') = log_softmax(blank ++ lex)
// HBM-bound: 2 * M * (V + 1) * 4 bytes forward, 4 * M * (V + 1) * 4 backward.
#include "common.cuh"
#include "params.cuh"

namespace lt {
namespace {

__device__ __forceinline__ float softplus(float x) {
  // log(1 + e^x) without overflow: max(x, 0) + log1p(e^-|x|)
  return fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x)));
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// log-sum-exp of the lexical row (and optionally the blank weight) of one row
__device__ __forceinline__ float row_lse(const float* __restrict__ lx, int V, int lane, bool with_blank,
                                         float b) {
  float m = with_blank ? b : neg_inf();
  for (int y = lane; y < V; y += 32) m = fmaxf(m, lx[y]);
  m = warp_max(m);
  const float ms = msafe(m);
  float s = 0.f;
  for (int y = lane; y < V; y += 32) s += __expf(lx[y] - ms);
  s = warp_sum(s);
  if (with_blank) s += __expf(b - ms);
  return ms + __logf(s);
}

__global__ void normalize_fwd_kernel(int mode, const float* __restrict__ blank,
                                     const float* __restrict__ lexical, long long M, int V,
                                     float* __restrict__ ob, float* __restrict__ ol) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
    const float* lx = lexical + (size_t)r * V;
    float* out = ol + (size_t)r * V;
    const float b = blank[r];
    if (mode == LT_NORM_HAT) {
      const float z = softplus(b);
      const float lse = row_lse(lx, V, lane, false, 0.f);
      for (int y = lane; y < V; y += 32) out[y] = lx[y] - lse - z;
      if (lane == 0) ob[r] = -softplus(-b);     // = b - softplus(b) without the cancellation
    } else {
      const float lse = row_lse(lx, V, lane, true, b);
      for (int y = lane; y < V; y += 32) out[y] = lx[y] - lse;
      if (lane == 0) ob[r] = b - lse;
    }
  }
}

// gradients w.r.t. the INPUTS given the cotangents (gb, gl) of the normalised outputs
__global__ void normalize_bwd_kernel(int mode, const float* __restrict__ blank,
                                     const float* __restrict__ lexical,
                                     const float* __restrict__ gb, const float* __restrict__ gl,
                                     long long M, int V, float* __restrict__ dblank,
                                     float* __restrict__ dlex) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < M; r += warps) {
    const float* lx = lexical + (size_t)r * V;
    const float* g = gl + (size_t)r * V;
    float* d = dlex + (size_t)r * V;
    const float b = blank[r], gbr = gb[r];
    float gsum = 0.f;
    for (int y = lane; y < V; y += 32) gsum += g[y];
    gsum = warp_sum(gsum);
    if (mode == LT_NORM_HAT) {
      // y_b = b - sp(b); y_l = l - lse(l) - sp(b);  sp'(b) = sigmoid(b)
      const float lse = row_lse(lx, V, lane, false, 0.f);
      const float sig = 1.f / (1.f + __expf(-b));
      for (int y = lane; y < V; y += 32) d[y] = g[y] - __expf(lx[y] - lse) * gsum;
      if (lane == 0) dblank[r] = gbr * (1.f - sig) - sig * gsum;
    } else {
      const float lse = row_lse(lx, V, lane, true, b);
      const float tot = gsum + gbr;
      for (int y = lane; y < V; y += 32) d[y] = g[y] - __expf(lx[y] - lse) * tot;
      if (lane == 0) dblank[r] = gbr - __expf(b - lse) * tot;
    }
  }
}

}  // namespace
}  // namespace lt

using namespace lt;

extern "C" {

int lt_local_normalize_forward(int mode, const float* blank, const float* lexical, int64_t M,
                               int V, float* out_blank, float* out_lexical, void* stream) {
  LT_CHECK_ARG(mode == LT_NORM_HAT || mode == LT_NORM_LOG_SOFTMAX,
               "lt_local_normalize_forward: unknown mode %d", mode);
  LT_CHECK_ARG(M >= 0 && V > 0, "lt_local_normalize_forward: bad sizes M=%lld V=%d", (long long)M, V);
  if (M == 0) return LT_OK;
  LT_CHECK_ARG(blank && lexical && out_blank && out_lexical, "lt_local_normalize_forward: NULL pointer");
  const long long blocks = (M + 7) / 8;
  const unsigned grid = (unsigned)(blocks < 148 * 32 ? blocks : 148 * 32);
  normalize_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mode, blank, lexical, M, V, out_blank,
                                                              out_lexical);
  LT_LAUNCHED();
  return LT_OK;
}

int lt_local_normalize_backward(int mode, const float* blank, const float* lexical,
                                const float* grad_out_blank, const float* grad_out_lexical,
                                int64_t M, int V, float* grad_blank, float* grad_lexical,
                                void* stream) {
  LT_CHECK_ARG(mode == LT_NORM_HAT || mode == LT_NORM_LOG_SOFTMAX,
               "lt_local_normalize_backward: unknown mode %d", mode);
  LT_CHECK_ARG(M >= 0 && V > 0, "lt_local_normalize_backward: bad sizes M=%lld V=%d", (long long)M, V);
  if (M == 0) return LT_OK;
  LT_CHECK_ARG(blank && lexical && grad_out_blank && grad_out_lexical && grad_blank && grad_lexical,
               "lt_local_normalize_backward: NULL pointer");
  const long long blocks = (M + 7) / 8;
  const unsigned grid = (unsigned)(blocks < 148 * 32 ? blocks : 148 * 32);
  normalize_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      mode, blank, lexical, grad_out_blank, grad_out_lexical, M, V, grad_blank, grad_lexical);
  LT_LAUNCHED();
  return LT_OK;
}

}  // extern "C"
