// Generic 64x64x16 fp32 tile GEMM on CUDA cores with functor-defined operands.
// Used by the JointWeightFn kernels for the small/odd shapes the tcgen05 path
// does not take, and as the fp32 cross-check of that path.
#pragma once
#include "common.cuh"

namespace lt {

// C[m, n] = sum_{k in [k_lo, k_hi)} A(m, k) * B(k, n); epilogue(m, n, acc).
// grid: (ceil(N/64), ceil(M/64), ksplit); block: 256 threads, 4x4 outputs each.
template <typename LoadA, typename LoadB, typename Epilogue>
__global__ void __launch_bounds__(256)
tile_gemm_kernel(int64_t M, int N, int64_t K, int64_t kchunk, LoadA load_a, LoadB load_b,
                 Epilogue epilogue) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float sa[BK][BM + 4];
  __shared__ float sb[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int64_t k_lo = (int64_t)blockIdx.z * kchunk;
  const int64_t k_hi = min(K, k_lo + kchunk);
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int64_t k0 = k_lo; k0 < k_hi; k0 += BK) {
    for (int i = tid; i < BM * BK; i += 256) {
      const int kk = i % BK, mm = i / BK;
      const int64_t m = m0 + mm, k = k0 + kk;
      sa[kk][mm] = (m < M && k < k_hi) ? load_a(m, k) : 0.f;
    }
    for (int i = tid; i < BN * BK; i += 256) {
      const int nn = i % BN, kk = i / BN;
      const int n = n0 + nn;
      const int64_t k = k0 + kk;
      sb[kk][nn] = (n < N && k < k_hi) ? load_b(k, n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sb[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t m = m0 + ty * 4 + i;
      const int n = n0 + tx * 4 + j;
      if (m < M && n < N) epilogue(m, n, acc[i][j]);
    }
}

}  // namespace lt
