// Semiring (+) on arbitrary tensors: Log / MaxTropical `plus` and `sum` with
// the reference's gradient rules
// (/root/reference/last_torch/semirings.py:202-220, :222-303, :330-401).
// Real is plain + / sum and stays a torch expression on the host side.
#include "common.cuh"
#include "params.cuh"

namespace lt {

template <int SR>
__global__ void plus_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = Sr<SR>::plus(a[i], b[i]);
}

// Log: grad = g * exp(x - out) (== g * e / z of semirings.py:264-269), 0 when
// out == -inf (z == 0 rule).  MaxTropical: ties go to `a` (semirings.py:363).
template <int SR>
__global__ void plus_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                const float* __restrict__ g, float* __restrict__ ga,
                                float* __restrict__ gb, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float x = a[i], y = b[i], gi = g[i];
    if constexpr (SR == LT_LOG) {
      const float c = fmaxf(x, y);
      const float cs = is_finite(c) ? c : 0.f;
      const float ea = fast_exp(x - cs), eb = fast_exp(y - cs);
      float z = ea + eb;
      z = (z != 0.f) ? z : 1.f;
      const float s = gi / z;
      ga[i] = s * ea;
      gb[i] = s * eb;
    } else {
      const bool ca = x >= y;
      ga[i] = ca ? gi : 0.f;
      gb[i] = ca ? 0.f : gi;
    }
  }
}

// a viewed as [outer, R, inner]; one thread per (o, i) when inner > 1 (coalesced
// across i), one warp per o when inner == 1 (coalesced across r).
template <int SR>
__global__ void sum_fwd_strided_kernel(const float* __restrict__ a, int64_t outer, int64_t R,
                                       int64_t inner, float* __restrict__ out,
                                       int32_t* __restrict__ argmax) {
  const int64_t total = outer * inner;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = idx / inner, i = idx % inner;
    const float* base = a + o * R * inner + i;
    if constexpr (SR == LT_LOG) {
      float m = neg_inf();
      for (int64_t r = 0; r < R; ++r) m = fmaxf(m, base[r * inner]);
      const float ms = msafe(m);
      float s = 0.f;
      for (int64_t r = 0; r < R; ++r) s += fast_exp(base[r * inner] - ms);
      out[idx] = ms + fast_log(s);
    } else {
      float m = base[0]; int32_t am = 0;
      for (int64_t r = 1; r < R; ++r) {
        const float v = base[r * inner];
        if (v > m) { m = v; am = (int32_t)r; }
      }
      out[idx] = m;
      if (argmax) argmax[idx] = am;
    }
  }
}

template <int SR>
__global__ void sum_fwd_rows_kernel(const float* __restrict__ a, int64_t outer, int64_t R,
                                    float* __restrict__ out, int32_t* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t o = warp; o < outer; o += nwarp) {
    const float* base = a + o * R;
    if constexpr (SR == LT_LOG) {
      float m = neg_inf();
      for (int64_t r = lane; r < R; r += 32) m = fmaxf(m, base[r]);
      for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
      const float ms = msafe(m);
      float s = 0.f;
      for (int64_t r = lane; r < R; r += 32) s += fast_exp(base[r] - ms);
      for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      if (lane == 0) out[o] = ms + fast_log(s);
    } else {
      float m = neg_inf(); int32_t am = 0x7fffffff;
      for (int64_t r = lane; r < R; r += 32) {
        const float v = base[r];
        if (v > m || (v == m && (int32_t)r < am)) { m = v; am = (int32_t)r; }
      }
      for (int s = 16; s > 0; s >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, s);
        const int32_t oa = __shfl_xor_sync(0xffffffffu, am, s);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
      }
      if (am == 0x7fffffff) am = 0;
      if (lane == 0) { out[o] = m; if (argmax) argmax[o] = am; }
    }
  }
}

// Log: grad_a = g * exp(a - out) (semirings.py:296-300 with the z != 0 guard);
// MaxTropical: one-hot at the first arg-max (semirings.py:389-398).
template <int SR>
__global__ void sum_bwd_kernel(const float* __restrict__ a, const float* __restrict__ out,
                               const int32_t* __restrict__ argmax, const float* __restrict__ g,
                               int64_t outer, int64_t R, int64_t inner,
                               float* __restrict__ ga) {
  const int64_t total = outer * R * inner;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % inner;
    const int64_t r = (idx / inner) % R;
    const int64_t o = idx / (inner * R);
    const int64_t oi = o * inner + i;
    if constexpr (SR == LT_LOG) {
      const float z = out[oi];
      ga[idx] = (z == neg_inf()) ? 0.f : g[oi] * fast_exp(a[idx] - z);
    } else {
      ga[idx] = (argmax[oi] == (int32_t)r) ? g[oi] : 0.f;
    }
  }
}

static inline int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

int semiring_plus_forward_launch(int sr, const float* a, const float* b, float* out, int64_t n,
                                 cudaStream_t stream) {
  if (n == 0) return LT_OK;
  if (sr == LT_LOG) plus_fwd_kernel<LT_LOG><<<grid_for(n, 256), 256, 0, stream>>>(a, b, out, n);
  else if (sr == LT_MAXTROPICAL) plus_fwd_kernel<LT_MAXTROPICAL><<<grid_for(n, 256), 256, 0, stream>>>(a, b, out, n);
  else if (sr == LT_REAL) plus_fwd_kernel<LT_REAL><<<grid_for(n, 256), 256, 0, stream>>>(a, b, out, n);
  else { set_error("lt_semiring_plus_forward: unknown semiring %d", sr); return LT_ERR_INVALID_ARGUMENT; }
  LT_LAUNCHED();
  return LT_OK;
}

int semiring_plus_backward_launch(int sr, const float* a, const float* b, const float* g,
                                  float* ga, float* gb, int64_t n, cudaStream_t stream) {
  if (n == 0) return LT_OK;
  if (sr == LT_LOG) plus_bwd_kernel<LT_LOG><<<grid_for(n, 256), 256, 0, stream>>>(a, b, g, ga, gb, n);
  else if (sr == LT_MAXTROPICAL) plus_bwd_kernel<LT_MAXTROPICAL><<<grid_for(n, 256), 256, 0, stream>>>(a, b, g, ga, gb, n);
  else { set_error("lt_semiring_plus_backward: semiring must be Log or MaxTropical, got %d", sr); return LT_ERR_INVALID_ARGUMENT; }
  LT_LAUNCHED();
  return LT_OK;
}

int semiring_sum_forward_launch(int sr, const float* a, int64_t outer, int64_t R, int64_t inner,
                                float* out, int32_t* argmax, cudaStream_t stream) {
  if (outer * inner == 0) return LT_OK;
  if (R <= 0) { set_error("lt_semiring_sum_forward: empty reduction axis"); return LT_ERR_INVALID_ARGUMENT; }
  if (sr != LT_LOG && sr != LT_MAXTROPICAL) {
    set_error("lt_semiring_sum_forward: semiring must be Log or MaxTropical, got %d", sr);
    return LT_ERR_INVALID_ARGUMENT;
  }
  if (inner == 1) {
    const int grid = grid_for(outer * 32, 256);
    if (sr == LT_LOG) sum_fwd_rows_kernel<LT_LOG><<<grid, 256, 0, stream>>>(a, outer, R, out, argmax);
    else sum_fwd_rows_kernel<LT_MAXTROPICAL><<<grid, 256, 0, stream>>>(a, outer, R, out, argmax);
  } else {
    const int grid = grid_for(outer * inner, 256);
    if (sr == LT_LOG) sum_fwd_strided_kernel<LT_LOG><<<grid, 256, 0, stream>>>(a, outer, R, inner, out, argmax);
    else sum_fwd_strided_kernel<LT_MAXTROPICAL><<<grid, 256, 0, stream>>>(a, outer, R, inner, out, argmax);
  }
  LT_LAUNCHED();
  return LT_OK;
}

int semiring_sum_backward_launch(int sr, const float* a, const float* out, const int32_t* argmax,
                                 const float* g, int64_t outer, int64_t R, int64_t inner,
                                 float* ga, cudaStream_t stream) {
  const int64_t n = outer * R * inner;
  if (n == 0) return LT_OK;
  if (sr == LT_LOG) sum_bwd_kernel<LT_LOG><<<grid_for(n, 256), 256, 0, stream>>>(a, out, argmax, g, outer, R, inner, ga);
  else if (sr == LT_MAXTROPICAL) {
    if (!argmax) { set_error("lt_semiring_sum_backward: MaxTropical needs argmax"); return LT_ERR_INVALID_ARGUMENT; }
    sum_bwd_kernel<LT_MAXTROPICAL><<<grid_for(n, 256), 256, 0, stream>>>(a, out, argmax, g, outer, R, inner, ga);
  } else { set_error("lt_semiring_sum_backward: semiring must be Log or MaxTropical, got %d", sr); return LT_ERR_INVALID_ARGUMENT; }
  LT_LAUNCHED();
  return LT_OK;
}

// alphas[b,t,c] += alpha_norm[b,t] (* ln 2 when the offsets are in log2 units): turns the renormalised alpha~ of
// lt_lattice_forward_norm into the alphas RecognitionLattice._forward returns (lattices.py:496).
__global__ void alphas_denormalize_kernel(float* __restrict__ alphas,
                                          const int32_t* __restrict__ alpha_norm, int T, int C,
                                          int64_t rows) {
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int64_t b = r / T, t = r - b * T;
    const int32_t* an = alpha_norm + b * (T + 3);
    const double off = (double)an[t] * (an[T + 2] == 0 ? 0.6931471805599453 : 1.0);
    float* row = alphas + r * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) row[c] = (float)((double)row[c] + off);
  }
}

int alphas_denormalize_launch(float* alphas, const int32_t* alpha_norm, int B, int T, int C,
                              cudaStream_t stream) {
  const int64_t rows = (int64_t)B * T;
  if (rows == 0 || C == 0) return LT_OK;
  const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 64);
  const int grid = (int)(rows < 148 * 16 ? rows : 148 * 16);
  alphas_denormalize_kernel<<<grid, threads, 0, stream>>>(alphas, alpha_norm, T, C, rows);
  LT_LAUNCHED();
  return LT_OK;
}

}  // namespace lt
