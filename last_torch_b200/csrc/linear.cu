// The two bias-free input projections of JointWeightFn (weight_fns.py:208-211:
// context_projection [C, E] -> [C, H], blank_projection [N, D] -> [N, H]) and their weight
// gradients, fp32 on CUDA cores.
//
// These are skinny products -- the frame projection is [32000 x 80] . [80 x 512] at the headline
// shape, its weight gradient a 32000-long reduction into [512 x 80] -- that a general sgemm serves
// badly (0.35 + 0.39 ms per step in the round-1 launch list).  fp32 FMAs keep them exact to the
// reference's own arithmetic; the tensor-core bf16x3 split buys nothing at 2.6 GFLOP.
//
//   lt_linear_forward : y[M, N] = x[M, K] . w[N, K]^T          (nn.Linear without bias)
//   lt_linear_wgrad   : gw[N, K] = gy[M, N]^T . x[M, K]        (fixed-order two-pass reduction)
// The input gradient gx = gy . w is lt_linear_forward(gy, w^T).
#include <cuda.h>
#include <stdint.h>

#include "common.cuh"
#include "params.cuh"

namespace lt {
namespace {

constexpr int kLinThreads = 256;
constexpr int kLinTile = 128;        // output tile edge
constexpr int kLinKC = 32;           // reduction chunk

// y tile [128 x 128]: 16 x 16 threads, 8 x 8 outputs each; x and w chunks are staged k-major
// ([kc][128 + 4]) so that a thread's eight rows / columns are two 128-bit reads.
__global__ void __launch_bounds__(kLinThreads, 2)
linear_forward_kernel(const float* __restrict__ x, const float* __restrict__ w,
                      float* __restrict__ y, int64_t M, int K, int N) {
  constexpr int LD = kLinTile + 4;
  __shared__ __align__(16) float xs[kLinKC][LD];
  __shared__ __align__(16) float ws[kLinKC][LD];
  const int tid = threadIdx.x;
  const int tm = tid >> 4, tn = tid & 15;
  const int64_t m0 = (int64_t)blockIdx.x * kLinTile;
  const int n0 = blockIdx.y * kLinTile;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kLinKC) {
    // 128 rows x 32 k of each operand: thread -> (row, 4 consecutive k)
    for (int idx = tid; idx < kLinTile * (kLinKC / 4); idx += kLinThreads) {
      const int r = idx >> 3, kq = (idx & 7) * 4;
      float vx[4] = {0.f, 0.f, 0.f, 0.f}, vw[4] = {0.f, 0.f, 0.f, 0.f};
      const int64_t m = m0 + r;
      const int n = n0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        if (k < K) {
          if (m < M) vx[j] = x[m * K + k];
          if (n < N) vw[j] = w[(int64_t)n * K + k];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { xs[kq + j][r] = vx[j]; ws[kq + j][r] = vw[j]; }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kLinKC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&xs[kk][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&xs[kk][tm * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&ws[kk][tn * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&ws[kk][tn * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + tm * 8 + i;
    if (m >= M) continue;
    float* row = y + m * N + n0 + tn * 8;
    if (n0 + tn * 8 + 8 <= N && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
      *reinterpret_cast<float4*>(row) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(row + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n0 + tn * 8 + j < N) row[j] = acc[i][j];
    }
  }
}

// Partial weight gradients: CTA (s, nb, kb) reduces the rows m = s*128, (s+S)*128, ... of
// gy[:, nb*128 ..] against x[:, kb*32 ..] into part[s][n][k]; 16 x 16 threads, 8 n x 2 k outputs.
__global__ void __launch_bounds__(kLinThreads, 2)
linear_wgrad_partial_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                            float* __restrict__ part, int64_t M, int K, int N) {
  constexpr int LDG = kLinTile + 4;
  __shared__ __align__(16) float gs[kLinKC][LDG];        // [m chunk of 32][128 n]
  __shared__ __align__(16) float xs[kLinKC][kLinKC + 2]; // [m chunk of 32][32 k]
  const int tid = threadIdx.x;
  const int tn = tid >> 4, tk = tid & 15;
  const int S = gridDim.x;
  const int n0 = blockIdx.y * kLinTile, k0 = blockIdx.z * kLinKC;
  float acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.f;
  for (int64_t mb = (int64_t)blockIdx.x * kLinKC; mb < M; mb += (int64_t)S * kLinKC) {
    for (int idx = tid; idx < kLinKC * (kLinTile / 4); idx += kLinThreads) {
      const int r = idx >> 5, nq = (idx & 31) * 4;
      const int64_t m = mb + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        const float* src = gy + m * N + n0 + nq;
        if (n0 + nq + 4 <= N && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          v = *reinterpret_cast<const float4*>(src);
        } else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < 4; ++j) if (n0 + nq + j < N) t[j] = src[j];
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      *reinterpret_cast<float4*>(&gs[r][nq]) = v;
    }
    for (int idx = tid; idx < kLinKC * kLinKC; idx += kLinThreads) {
      const int r = idx >> 5, k = idx & 31;
      const int64_t m = mb + r;
      xs[r][k] = (m < M && k0 + k < K) ? x[m * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < kLinKC; ++r) {
      const float4 g0 = *reinterpret_cast<const float4*>(&gs[r][tn * 8]);
      const float4 g1 = *reinterpret_cast<const float4*>(&gs[r][tn * 8 + 4]);
      const float2 xv = *reinterpret_cast<const float2*>(&xs[r][tk * 2]);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(g[i], xv.x, acc[i][0]);
        acc[i][1] = fmaf(g[i], xv.y, acc[i][1]);
      }
    }
    __syncthreads();
  }
  float* out = part + (size_t)blockIdx.x * N * K;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + tn * 8 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = k0 + tk * 2 + j;
      if (k < K) out[(size_t)n * K + k] = acc[i][j];
    }
  }
}

__global__ void linear_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw,
                                           int S, int64_t NK) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < NK;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < S; ++p) s += part[(size_t)p * NK + i];      // fixed order
    gw[i] = s;
  }
}

int wgrad_slabs(int64_t M) {
  int64_t blocks = (M + kLinKC - 1) / kLinKC;
  return (int)(blocks < 64 ? (blocks < 1 ? 1 : blocks) : 64);
}

}  // namespace
}  // namespace lt

using namespace lt;

extern "C" int lt_linear_forward(const float* x, const float* w, float* y, int64_t M, int K, int N,
                                 void* stream) {
  LT_CHECK_ARG(M >= 0 && K > 0 && N > 0, "lt_linear_forward: bad sizes M=%lld K=%d N=%d",
               (long long)M, K, N);
  if (M == 0) return LT_OK;
  LT_CHECK_ARG(x && w && y, "lt_linear_forward: NULL pointer");
  dim3 grid((unsigned)((M + kLinTile - 1) / kLinTile), (unsigned)((N + kLinTile - 1) / kLinTile));
  linear_forward_kernel<<<grid, kLinThreads, 0, (cudaStream_t)stream>>>(x, w, y, M, K, N);
  LT_LAUNCHED();
  return LT_OK;
}

extern "C" int64_t lt_linear_wgrad_workspace_bytes(int64_t M, int K, int N) {
  return (int64_t)wgrad_slabs(M) * N * K * (int64_t)sizeof(float);
}

extern "C" int lt_linear_wgrad(const float* gy, const float* x, float* gw, int64_t M, int K, int N,
                               void* workspace, void* stream) {
  LT_CHECK_ARG(M >= 0 && K > 0 && N > 0, "lt_linear_wgrad: bad sizes M=%lld K=%d N=%d",
               (long long)M, K, N);
  LT_CHECK_ARG(gw && (M == 0 || (gy && x && workspace)), "lt_linear_wgrad: NULL pointer");
  if (M == 0) {
    LT_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)N * K, (cudaStream_t)stream));
    return LT_OK;
  }
  const int S = wgrad_slabs(M);
  dim3 grid((unsigned)S, (unsigned)((N + kLinTile - 1) / kLinTile),
            (unsigned)((K + kLinKC - 1) / kLinKC));
  float* part = reinterpret_cast<float*>(workspace);
  linear_wgrad_partial_kernel<<<grid, kLinThreads, 0, (cudaStream_t)stream>>>(gy, x, part, M, K, N);
  LT_LAUNCHED();
  const int64_t nk = (int64_t)N * K;
  linear_wgrad_reduce_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      part, gw, S, nk);
  LT_LAUNCHED();
  return LT_OK;
}
