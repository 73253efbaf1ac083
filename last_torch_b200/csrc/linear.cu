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
#include "umma.cuh"

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


// ------------------------------------------------------------------------------------------
// Tensor-core variants (tcgen05, bf16x3 operand split, fp32 accumulation in tensor memory) for
// the shapes of a production-sized joint network (D = E = H = 512 at the headline: 16.8 GFLOP
// per projection, where the fp32 FMA kernels -- the library's or the ones above -- need 0.35 ms).
// Same building blocks as the joint kernels (umma.cuh): operands are split into bf16 hi / lo
// while they are staged into the SWIZZLE_128B shared-memory layout, three MMAs per product
// (hi*hi, hi*lo, lo*hi: 2^-17 relative), one 128 x <=256 fp32 accumulator per CTA, two CTAs per SM
// so that one stages while the other multiplies.
constexpr int kLinTcThreads = 256;   // warps 0-3 and 4-7 share the TMEM lane quadrants in the epilogue

__device__ __forceinline__ void lin_mbar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void lin_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LTL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LTL_DONE;\n"
      "bra LTL_WAIT;\n"
      "LTL_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// rows [0, nrows) x 64 K-elements of a row-major fp32 matrix -> bf16 hi / lo tiles, K-major
// SWIZZLE_128B; rows at or beyond `valid` are written as zeros.  Four 32-byte pieces per thread
// are in flight at a time.
__device__ __forceinline__ void lin_fill_kmajor(const float* __restrict__ src, int64_t ld, int nrows,
                                                int64_t valid, unsigned char* hi, unsigned char* lo,
                                                int tid, int nthreads) {
  for (int idx0 = tid; idx0 < nrows * 8; idx0 += 4 * nthreads) {
    float4 v[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = idx0 + j * nthreads;
      const int row = idx >> 3, chunk = idx & 7;
      if (idx < nrows * 8 && row < valid) {
        const float4* q = reinterpret_cast<const float4*>(src + (int64_t)row * ld + chunk * 8);
        v[j][0] = q[0]; v[j][1] = q[1];
      } else {
        v[j][0] = v[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = idx0 + j * nthreads;
      if (idx >= nrows * 8) break;
      const int row = idx >> 3, chunk = idx & 7;
      const float x[8] = {v[j][0].x, v[j][0].y, v[j][0].z, v[j][0].w,
                          v[j][1].x, v[j][1].y, v[j][1].z, v[j][1].w};
      uint4 h, l;
      umma::split_pack8(x, h, l);
      const uint32_t off = umma::swizzled_offset(row, chunk);
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = l;
    }
  }
}

// y[m0 .. m0+128, n0 .. n0+nt] = x[.., K] . w[n0 .., K]^T ; K % 64 == 0, nt % 16 == 0, nt <= 256
__global__ void __launch_bounds__(kLinTcThreads, 2)
linear_forward_tc_kernel(const float* __restrict__ x, const float* __restrict__ w,
                         float* __restrict__ y, int64_t M, int K, int N) {
  extern __shared__ __align__(1024) unsigned char lsm_raw[];
  unsigned char* sm = lsm_raw + ((1024u - (smem_u32(lsm_raw) & 1023u)) & 1023u);
  unsigned char* a_hi = sm;                          // 128 x 128 B
  unsigned char* a_lo = a_hi + 128 * 128;
  unsigned char* b_hi = a_lo + 128 * 128;            // 256 x 128 B
  unsigned char* b_lo = b_hi + 256 * 128;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  const int n0 = blockIdx.y * 256;
  const int nt = min(256, N - n0);
  if (tid == 0) {
    lin_mbar_init(smem_u32(&mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) umma::tmem_alloc(smem_u32(&tmem_base), 256);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = umma::make_idesc_bf16(128, nt);
  const int nchunks = K / 64;
  for (int kc = 0; kc < nchunks; ++kc) {
    lin_fill_kmajor(x + m0 * K + kc * 64, K, 128, M - m0, a_hi, a_lo, tid, kLinTcThreads);
    lin_fill_kmajor(w + (int64_t)n0 * K + kc * 64, K, nt, nt, b_hi, b_lo, tid, kLinTcThreads);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> async proxy
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_thread_sync();
#pragma unroll
      for (int k = 0; k < 4; ++k) {           // 4 x (K = 16 bf16 = 32 bytes) per 128-byte row
        const uint64_t dah = umma::make_smem_desc_sw128(smem_u32(a_hi) + k * 32);
        const uint64_t dal = umma::make_smem_desc_sw128(smem_u32(a_lo) + k * 32);
        const uint64_t dbh = umma::make_smem_desc_sw128(smem_u32(b_hi) + k * 32);
        const uint64_t dbl = umma::make_smem_desc_sw128(smem_u32(b_lo) + k * 32);
        umma::mma_bf16(tmem, dah, dbh, idesc, (kc | k) > 0);
        umma::mma_bf16(tmem, dah, dbl, idesc, 1);
        umma::mma_bf16(tmem, dal, dbh, idesc, 1);
      }
      umma::commit(smem_u32(&mbar));          // arrives when the MMAs above have completed
    }
    lin_mbar_wait(smem_u32(&mbar), kc & 1);   // the operand tiles may be overwritten after this
  }
  umma::fence_after_thread_sync();
  // epilogue: warp w (and w + 4 for the other half of the columns) owns TMEM lanes (= rows)
  // 32 (w % 4) .. + 31, 32 columns at a time
  const int q = warp & 3, half = warp >> 2;
  const int64_t row = m0 + q * 32 + (tid & 31);
  for (int c0 = half * 32; c0 < nt; c0 += 64) {
    float v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
    if (row < M) {
      float* dst = y + row * N + n0 + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (c0 + i < nt)
          *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// Weight gradient: part[s][n0 .. n0+128, k0 .. k0+kt] = sum over the rows m of split s of
// gy[m, n0 ..]^T x[m, k0 ..]: both operands are contiguous along the OUTPUT dimensions (MN-major),
// the reduction runs over M in chunks of 32 rows.  kt % 64 == 0, kt <= 256.
__global__ void __launch_bounds__(kLinTcThreads, 2)
linear_wgrad_tc_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                       float* __restrict__ part, int64_t M, int K, int N, int64_t rows_per_split) {
  extern __shared__ __align__(1024) unsigned char wsm_raw[];
  unsigned char* sm = wsm_raw + ((1024u - (smem_u32(wsm_raw) & 1023u)) & 1023u);
  unsigned char* a_hi = sm;                        // [128 x 32] bf16 = 8 KB
  unsigned char* a_lo = a_hi + 128 * 32 * 2;
  unsigned char* b_hi = a_lo + 128 * 32 * 2;      // [256 x 32] bf16 = 16 KB
  unsigned char* b_lo = b_hi + 256 * 32 * 2;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n0 = blockIdx.x * 128, k0 = blockIdx.y * 256;
  const int kt = min(256, K - k0);
  const int64_t mb = (int64_t)blockIdx.z * rows_per_split;
  const int64_t me = mb + rows_per_split < M ? mb + rows_per_split : M;
  if (tid == 0) {
    lin_mbar_init(smem_u32(&mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) umma::tmem_alloc(smem_u32(&tmem_base), 256);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = umma::make_idesc_bf16_mn(128, kt);
  const uint32_t a_sbo = (128 / 64) * 1024, b_sbo = (uint32_t)(kt / 64) * 1024, lbo = 1024;
  int it = 0;
  for (int64_t m = mb; m < me; m += 32, ++it) {
    // A: gy rows m .. m+31 (the reduction index), 128 contiguous output features each
    for (int idx = tid; idx < 32 * 16; idx += kLinTcThreads) {
      const int r = idx >> 4, mn = (idx & 15) * 8;
      float v[8];
      if (m + r < me) {
        const float4* q = reinterpret_cast<const float4*>(gy + (m + r) * N + n0 + mn);
        const float4 q0 = q[0], q1 = q[1];
        v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w;
        v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      uint4 h, l;
      umma::split_pack8(v, h, l);
      const uint32_t off = umma::mn_major_chunk_offset(128, mn, r);
      *reinterpret_cast<uint4*>(a_hi + off) = h;
      *reinterpret_cast<uint4*>(a_lo + off) = l;
    }
    // B: x rows m .. m+31, kt contiguous input features each
    for (int idx = tid; idx < 32 * (kt / 8); idx += kLinTcThreads) {
      const int r = idx / (kt / 8), mn = (idx % (kt / 8)) * 8;
      float v[8];
      if (m + r < me) {
        const float4* q = reinterpret_cast<const float4*>(x + (m + r) * K + k0 + mn);
        const float4 q0 = q[0], q1 = q[1];
        v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w;
        v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      uint4 h, l;
      umma::split_pack8(v, h, l);
      const uint32_t off = umma::mn_major_chunk_offset(kt, mn, r);
      *reinterpret_cast<uint4*>(b_hi + off) = h;
      *reinterpret_cast<uint4*>(b_lo + off) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_thread_sync();
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {            // K = 16 per instruction = 2 atoms deep
        const uint32_t aoff = 2 * ks * a_sbo, boff = 2 * ks * b_sbo;
        const uint64_t dah = umma::make_smem_desc_mn_sw128(smem_u32(a_hi) + aoff, lbo, a_sbo);
        const uint64_t dal = umma::make_smem_desc_mn_sw128(smem_u32(a_lo) + aoff, lbo, a_sbo);
        const uint64_t dbh = umma::make_smem_desc_mn_sw128(smem_u32(b_hi) + boff, lbo, b_sbo);
        const uint64_t dbl = umma::make_smem_desc_mn_sw128(smem_u32(b_lo) + boff, lbo, b_sbo);
        umma::mma_bf16(tmem, dah, dbh, idesc, (it | ks) > 0);
        umma::mma_bf16(tmem, dah, dbl, idesc, 1);
        umma::mma_bf16(tmem, dal, dbh, idesc, 1);
      }
      umma::commit(smem_u32(&mbar));
    }
    lin_mbar_wait(smem_u32(&mbar), it & 1);
  }
  umma::fence_after_thread_sync();
  float* out = part + (size_t)blockIdx.z * N * K;
  const int q = warp & 3, half = warp >> 2;
  const int row = n0 + q * 32 + (tid & 31);
  for (int c0 = half * 32; c0 < kt; c0 += 64) {
    float v[32];
    if (it > 0) {
      umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;     // a split without rows
    }
    if (row < N) {
      float* dst = out + (size_t)row * K + k0 + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (c0 + i < kt)
          *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

bool linear_tc_forward_ok(const void* x, const void* w, const void* y, int64_t M, int K, int N) {
  return M >= 128 && K % 64 == 0 && N % 16 == 0 &&
         ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) |
           reinterpret_cast<uintptr_t>(y)) & 15) == 0;
}
// gw [N, K] = gy[M, N]^T x[M, K]
bool linear_tc_wgrad_ok(const void* gy, const void* x, int64_t M, int K, int N) {
  return M >= 1024 && N % 128 == 0 && K % 64 == 0 &&
         ((reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
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
  if (!option(OPT_LINEAR_SIMT) && linear_tc_forward_ok(x, w, y, M, K, N)) {
    const size_t smem = 2 * 128 * 128 + 2 * 256 * 128 + 1024;
    LT_CUDA(cudaFuncSetAttribute(linear_forward_tc_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 tgrid((unsigned)((M + 127) / 128), (unsigned)((N + 255) / 256));
    linear_forward_tc_kernel<<<tgrid, kLinTcThreads, smem, (cudaStream_t)stream>>>(x, w, y, M, K, N);
    LT_LAUNCHED();
    return LT_OK;
  }
  dim3 grid((unsigned)((M + kLinTile - 1) / kLinTile), (unsigned)((N + kLinTile - 1) / kLinTile));
  linear_forward_kernel<<<grid, kLinThreads, 0, (cudaStream_t)stream>>>(x, w, y, M, K, N);
  LT_LAUNCHED();
  return LT_OK;
}

// 1 when lt_linear_forward / lt_linear_wgrad take the tcgen05 kernels for [M, K] x [N, K]
// (16-byte aligned buffers assumed) AND the product is large enough for them to beat an fp32 FMA
// GEMM: measured 0.109 vs 0.33 ms (forward) and 0.147 vs 0.38 ms (weight gradient) at
// 32000 x 512 x 512, but 0.040 vs 0.015 ms at 257 x 512 x 512.
extern "C" int lt_linear_tensor_core(int64_t M, int K, int N) {
  const void* aligned = reinterpret_cast<const void*>(uintptr_t(256));
  if (option(OPT_LINEAR_SIMT) || M < 4096) return 0;
  return (linear_tc_forward_ok(aligned, aligned, aligned, M, K, N) &&
          linear_tc_wgrad_ok(aligned, aligned, M, K, N)) ? 1 : 0;
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
  float* part = reinterpret_cast<float*>(workspace);
  if (!option(OPT_LINEAR_SIMT) && linear_tc_wgrad_ok(gy, x, M, K, N)) {
    // as many splits of the reduction as fill the chip twice over, at most the workspace's
    // slabs; every split is a multiple of 32 rows
    const int tiles = (N / 128) * ((K + 255) / 256);
    int S = (2 * 148 + tiles - 1) / tiles;
    const int smax = wgrad_slabs(M);
    S = S < 1 ? 1 : (S > smax ? smax : S);
    int64_t rows = (M + S - 1) / S;
    rows = (rows + 31) / 32 * 32;
    S = (int)((M + rows - 1) / rows);
    const size_t smem = 2 * 128 * 64 + 2 * 256 * 64 + 1024;
    LT_CUDA(cudaFuncSetAttribute(linear_wgrad_tc_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 tgrid((unsigned)(N / 128), (unsigned)((K + 255) / 256), (unsigned)S);
    linear_wgrad_tc_kernel<<<tgrid, kLinTcThreads, smem, (cudaStream_t)stream>>>(gy, x, part, M, K, N, rows);
    LT_LAUNCHED();
    const int64_t nk = (int64_t)N * K;
    linear_wgrad_reduce_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        part, gw, S, nk);
    LT_LAUNCHED();
    return LT_OK;
  }
  const int S = wgrad_slabs(M);
  dim3 grid((unsigned)S, (unsigned)((N + kLinTile - 1) / kLinTile),
            (unsigned)((K + kLinKC - 1) / kLinKC));
  linear_wgrad_partial_kernel<<<grid, kLinThreads, 0, (cudaStream_t)stream>>>(gy, x, part, M, K, N);
  LT_LAUNCHED();
  const int64_t nk = (int64_t)N * K;
  linear_wgrad_reduce_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      part, gw, S, nk);
  LT_LAUNCHED();
  return LT_OK;
}
