// K4 on 5th-generation tensor cores (tcgen05 + TMEM), hand-written.
//
// Precision: the reference computes the vocabulary projection in fp32
// (weight_fns.py:220-227) and parity is 1e-5, which plain bf16 / tf32 inputs
// cannot hold.  Both operands are therefore split  x = hi + lo  into two bf16
// values (|x - hi - lo| <= 2^-17 |x|) and the product is accumulated in fp32 in
// TMEM as  Ah*Bh + Ah*Bl + Al*Bh  (three tcgen05.mma per K step; the dropped
// Al*Bl term is 2^-18 relative).
//
// This file currently holds the validated building block (ltx_umma_probe, used
// by tests/test_gpu_umma.py): operands written by threads into the K-major
// SWIZZLE_128B shared-memory layout, tcgen05.mma issued by one thread,
// completion via tcgen05.commit -> mbarrier, accumulator read back with
// tcgen05.ld.  The pipelined JointWeightFn kernels are built from it.
#include <cuda.h>
#include <type_traits>
#include <stdlib.h>

#include "common.cuh"
#include "params.cuh"
#include "umma.cuh"

namespace lt {
namespace {

__device__ __forceinline__ void mbar_init1(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LTJ_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1" LT_MBAR_HINT ";\n"
      "@p bra LTJ_DONE;\n"
      "bra LTJ_WAIT;\n"
      "LTJ_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// Writes rows [0, nrows) x 64 K-elements of a row-major fp32 matrix (leading
// dimension ld) as bf16 hi / lo tiles in the K-major SWIZZLE_128B layout.
__device__ __forceinline__ void fill_tile_split(const float* __restrict__ src, int ld, int nrows,
                                                unsigned char* hi, unsigned char* lo, int tid,
                                                int nthreads) {
  for (int idx = tid; idx < nrows * 8; idx += nthreads) {
    const int row = idx >> 3, chunk = idx & 7;
    const float4 x0 = *reinterpret_cast<const float4*>(src + (size_t)row * ld + chunk * 8);
    const float4 x1 = *reinterpret_cast<const float4*>(src + (size_t)row * ld + chunk * 8 + 4);
    const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) umma::split_bf16(x[i], h[i], l[i]);
    const uint32_t off = umma::swizzled_offset(row, chunk);
    *reinterpret_cast<uint4*>(hi + off) =
        make_uint4(umma::pack_bf16(h[0], h[1]), umma::pack_bf16(h[2], h[3]),
                   umma::pack_bf16(h[4], h[5]), umma::pack_bf16(h[6], h[7]));
    *reinterpret_cast<uint4*>(lo + off) =
        make_uint4(umma::pack_bf16(l[0], l[1]), umma::pack_bf16(l[2], l[3]),
                   umma::pack_bf16(l[4], l[5]), umma::pack_bf16(l[6], l[7]));
  }
}

// D[128, N] = A[128, K] * B[N, K]^T  (fp32 in/out, bf16x3 on tcgen05).  One CTA.
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                  int N, int K, int terms) {
  extern __shared__ __align__(1024) unsigned char psmem_raw[];
  // SWIZZLE_128B atoms need a 1024-byte aligned base: align by hand
  unsigned char* psmem = psmem_raw + ((1024u - (smem_u32(psmem_raw) & 1023u)) & 1023u);
  unsigned char* a_hi = psmem;                       // 128 x 128 B
  unsigned char* a_lo = a_hi + 128 * 128;
  unsigned char* b_hi = a_lo + 128 * 128;            // N x 128 B
  unsigned char* b_lo = b_hi + 256 * 128;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init1(smem_u32(&mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) umma::tmem_alloc(smem_u32(&tmem_base), 256);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = umma::make_idesc_bf16(128, N);
  const int nchunks = K / 64;
  for (int kc = 0; kc < nchunks; ++kc) {
    fill_tile_split(A + kc * 64, K, 128, a_hi, a_lo, tid, blockDim.x);
    fill_tile_split(B + kc * 64, K, N, b_hi, b_lo, tid, blockDim.x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> async proxy
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_thread_sync();
      for (int k = 0; k < 4; ++k) {           // 4 x (K = 16 bf16 = 32 bytes) per 128-byte row
        const uint64_t dah = umma::make_smem_desc_sw128(smem_u32(a_hi) + k * 32);
        const uint64_t dal = umma::make_smem_desc_sw128(smem_u32(a_lo) + k * 32);
        const uint64_t dbh = umma::make_smem_desc_sw128(smem_u32(b_hi) + k * 32);
        const uint64_t dbl = umma::make_smem_desc_sw128(smem_u32(b_lo) + k * 32);
        umma::mma_bf16(tmem, dah, dbh, idesc, (kc | k) > 0);
        if (terms >= 2) umma::mma_bf16(tmem, dah, dbl, idesc, 1);
        if (terms >= 3) umma::mma_bf16(tmem, dal, dbh, idesc, 1);
      }
      umma::commit(smem_u32(&mbar));          // arrives when the MMAs above have completed
    }
    mbar_wait_parity(smem_u32(&mbar), kc & 1);   // operands may be overwritten after this
  }
  umma::fence_after_thread_sync();
  // epilogue: warp w owns TMEM lanes (= rows) 32w .. 32w+31
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}


// MN-major probe: D[128, N] = At^T * Bt with At [K, 128], Bt [K, N] fp32 row-major
// (both operands contiguous along M / N).  `swap` exchanges LBO and SBO in the
// descriptors (used once to pin the descriptor semantics on hardware).
__global__ void __launch_bounds__(128, 1)
umma_probe_mn_kernel(const float* __restrict__ At, const float* __restrict__ Bt,
                     float* __restrict__ D, int N, int K, int swap) {
  extern __shared__ __align__(1024) unsigned char qsmem_raw[];
  unsigned char* q = qsmem_raw + ((1024u - (smem_u32(qsmem_raw) & 1023u)) & 1023u);
  unsigned char* a_hi = q;                        // [128 x 32] bf16 = 8 KB
  unsigned char* a_lo = a_hi + 128 * 32 * 2;
  unsigned char* b_hi = a_lo + 128 * 32 * 2;      // [256 x 32] bf16 = 16 KB
  unsigned char* b_lo = b_hi + 256 * 32 * 2;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init1(smem_u32(&mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) umma::tmem_alloc(smem_u32(&tmem_base), 256);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = umma::make_idesc_bf16_mn(128, N);
  for (int kc = 0; kc < K / 32; ++kc) {
    for (int idx = tid; idx < 32 * (128 / 8); idx += blockDim.x) {      // A chunks
      const int k = idx / 16, mn = (idx % 16) * 8;
      __nv_bfloat16 h[8], l[8];
      for (int i = 0; i < 8; ++i) umma::split_bf16(At[(size_t)(kc * 32 + k) * 128 + mn + i], h[i], l[i]);
      const uint32_t off = umma::mn_major_chunk_offset(128, mn, k);
      *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(umma::pack_bf16(h[0], h[1]), umma::pack_bf16(h[2], h[3]), umma::pack_bf16(h[4], h[5]), umma::pack_bf16(h[6], h[7]));
      *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(umma::pack_bf16(l[0], l[1]), umma::pack_bf16(l[2], l[3]), umma::pack_bf16(l[4], l[5]), umma::pack_bf16(l[6], l[7]));
    }
    for (int idx = tid; idx < 32 * (N / 8); idx += blockDim.x) {        // B chunks
      const int k = idx / (N / 8), mn = (idx % (N / 8)) * 8;
      __nv_bfloat16 h[8], l[8];
      for (int i = 0; i < 8; ++i) umma::split_bf16(Bt[(size_t)(kc * 32 + k) * N + mn + i], h[i], l[i]);
      const uint32_t off = umma::mn_major_chunk_offset(N, mn, k);
      *reinterpret_cast<uint4*>(b_hi + off) = make_uint4(umma::pack_bf16(h[0], h[1]), umma::pack_bf16(h[2], h[3]), umma::pack_bf16(h[4], h[5]), umma::pack_bf16(h[6], h[7]));
      *reinterpret_cast<uint4*>(b_lo + off) = make_uint4(umma::pack_bf16(l[0], l[1]), umma::pack_bf16(l[2], l[3]), umma::pack_bf16(l[4], l[5]), umma::pack_bf16(l[6], l[7]));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_thread_sync();
      const uint32_t a_sbo = (128 / 64) * 1024, b_sbo = (N / 64) * 1024, lbo = 1024;
      for (int ks = 0; ks < 2; ++ks) {            // K = 16 per instruction = 2 atoms deep
        const uint32_t aoff = 2 * ks * a_sbo, boff = 2 * ks * b_sbo;
        const uint64_t dah = umma::make_smem_desc_mn_sw128(smem_u32(a_hi) + aoff, swap ? a_sbo : lbo, swap ? lbo : a_sbo);
        const uint64_t dal = umma::make_smem_desc_mn_sw128(smem_u32(a_lo) + aoff, swap ? a_sbo : lbo, swap ? lbo : a_sbo);
        const uint64_t dbh = umma::make_smem_desc_mn_sw128(smem_u32(b_hi) + boff, swap ? b_sbo : lbo, swap ? lbo : b_sbo);
        const uint64_t dbl = umma::make_smem_desc_mn_sw128(smem_u32(b_lo) + boff, swap ? b_sbo : lbo, swap ? lbo : b_sbo);
        umma::mma_bf16(tmem, dah, dbh, idesc, (kc | ks) > 0);
        umma::mma_bf16(tmem, dah, dbl, idesc, 1);
        umma::mma_bf16(tmem, dal, dbh, idesc, 1);
      }
      umma::commit(smem_u32(&mbar));
    }
    mbar_wait_parity(smem_u32(&mbar), kc & 1);
  }
  umma::fence_after_thread_sync();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// ===========================================================================
// JointWeightFn forward on tcgen05 (weight_fns.py:208-227, whole-utterance form)
//   lexical[m, :] = tanh(pc[c] + pf[n]) . W_vocab^T + b_vocab      m = n*C + c
//   blank[m]      = tanh(pc[c] + pf[n]) . w_blank   + b_blank
// Persistent CTAs (one per SM), tile = 128 rows x V columns, K = H in chunks of 64.
// Warp roles (448 threads):
//   warp 0      TMA producer: W_hi / W_lo chunk [V x 64] bf16 (SWIZZLE_128B tensor maps)
//   warp 1      MMA issuer (one thread): 3 tcgen05.mma per 16-wide K step (bf16x3 split)
//   warps 2-5   epilogue: TMEM -> registers -> + bias -> coalesced row stores
//   warps 6-13  A producers: tanh(pc + pf) on the fly, hi/lo split, swizzled smem
//               stores; the same threads accumulate the blank mat-vec in registers
// Pipelines: smem full/empty (2 stages), TMEM full/empty (2 accumulators of 256 columns),
// so the epilogue of tile i overlaps the MMAs of tile i+1.  The [M, H] joint
// activations never touch HBM.
// ===========================================================================
constexpr int kJThreads = 448;
constexpr int kJStages = 2;
constexpr int kJProducers = 256;
// forward kernel: 16 producer warps (2 row passes each) -> more overlapping tanh chains
constexpr int kFProdWarps = 8;
constexpr int kFThreads = (6 + kFProdWarps) * 32;
constexpr int kFProducers = kFProdWarps * 32;
constexpr int kFPasses = 128 / (kFProdWarps * 4);

struct JointTcParams {
  const float* pc;       // [C, H]  e^(2 proj_ctx)   (joint_exp_table_kernel)
  const float* pf;       // [N, H]  e^(2 proj_frame)
  const float* w_blank;  // [H]
  const float* b_vocab;  // [V]
  const float* b_blank;   // device scalar
  long long M;           // N * C
  int C, H, V;
  float* blank;          // [M]
  float* lexical;        // [M, V]
};

__device__ __forceinline__ void mbar_init_n(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                       uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// tanh(x) = 1 - 2 / (1 + e^(2x)): two MUFU ops; absolute error ~1e-7
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (2.f * kLog2e)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return fmaf(-2.f, r, 1.f);
}


// Epilogue store: the accumulator arrives with lane = row (TMEM lane) and 32 columns
// per thread.  Writing it directly would touch 32 different 128-byte lines per store
// instruction; instead the 32 x 32 block goes through a padded per-warp shared tile
// (conflict-free both ways) and every store instruction writes ONE full line of one row.
__device__ __forceinline__ void store_block_coalesced(const float (&v)[32], float* tr, int lane,
                                                      float* out_row0, size_t row_stride,
                                                      int rows_valid) {
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = v[j];
  __syncwarp();
#pragma unroll 8
  for (int r = 0; r < 32; ++r)
    if (r < rows_valid) out_row0[(size_t)r * row_stride + lane] = tr[r * 33 + lane];
}

// Exponential tables of the two projections: out = e^(2x) = 2^(2 log2(e) x), exponent clamped
// to +-126 so that every entry is a normal, non-zero fp32 number.  tanh(pc + pf) then costs one
// reciprocal: 1 - 2 / (1 + E_c E_f); the product may overflow or flush to zero, which IS the
// saturated tanh (common.cuh).  Evaluated in double precision (the table is O((C + N) H), the
// joint is O(N C H)): each entry is correctly rounded, so the product carries ~1.5 ulp -- no worse
// than rounding the fp32 argument 2 log2(e) (pc + pf) of an ex2.  The clamp only matters for
// |x| > 43.6, where fp32 tanh has long saturated (|x| > 9.1) unless the OTHER projection cancels
// it to within 9: both beyond 34 with opposite signs.
__global__ void joint_exp_table_kernel(const float* __restrict__ x, float* __restrict__ out,
                                       long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    double a = (double)x[i] * 2.8853900817779268;
    a = a < -126.0 ? -126.0 : (a > 126.0 ? 126.0 : a);
    out[i] = (float)exp2(a);
  }
}

__global__ void split_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                     __nv_bfloat16* __restrict__ lo, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    umma::split_bf16(w[i], hi[i], lo[i]);
}

__global__ void __launch_bounds__(kFThreads, 1)
joint_forward_tc_kernel(const __grid_constant__ CUtensorMap map_hi,
                        const __grid_constant__ CUtensorMap map_lo,
                        const __grid_constant__ CUtensorMap map_out, const JointTcParams p) {
  extern __shared__ __align__(1024) unsigned char jsmem_raw[];
  unsigned char* base = jsmem_raw + ((1024u - (smem_u32(jsmem_raw) & 1023u)) & 1023u);
  const int V = p.V, H = p.H;
  const uint32_t a_bytes = 128 * 128;                 // one 128 x 64 bf16 tile
  const uint32_t b_bytes = (uint32_t)V * 128;         // one V x 64 bf16 tile
  const uint32_t stage_bytes = 2 * a_bytes + 2 * 256 * 128;
  // stage layout: A_hi | A_lo | B_hi | B_lo
  // epilogue staging: one [32 rows x 128 B] SWIZZLE_128B tile per epilogue warp (1024-aligned)
  unsigned char* s_out = base + kJStages * stage_bytes;                    // 4 x 4096 B
  float* s_wb = reinterpret_cast<float*>(s_out + 4 * 4096);                // [H], permuted
  float* s_bias = s_wb + H;                                                // [V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint64_t* full = bars;                    // [stages]  producers + TMA -> MMA
  uint64_t* empty = bars + kJStages;        // [stages]  MMA (commit) -> producers, TMA
  uint64_t* tfull = bars + 2 * kJStages;    // [2]       MMA (commit) -> epilogue
  uint64_t* tempty = tfull + 2;             // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = H / 64;
  const long long num_tiles = (p.M + 127) / 128;

  // w_blank permuted inside every 64-wide chunk so that the 8 producer lanes of a row read two
  // contiguous 128-byte runs (elements 0-3 of each lane, then elements 4-7) with one wavefront
  // each instead of 16-byte pieces at a 32-byte stride
  for (int i = tid; i < H; i += kFThreads) {
    const int w = i & 63, lane8 = w >> 3, e = w & 7;
    s_wb[(i & ~63) + (e < 4 ? lane8 * 4 + e : 32 + lane8 * 4 + (e - 4))] = p.w_blank[i];
  }
  for (int i = tid; i < V; i += kFThreads) s_bias[i] = p.b_vocab[i];
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    for (int s = 0; s < kJStages; ++s) {
      mbar_init_n(smem_u32(&full[s]), kFProducers + 1);
      mbar_init_n(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init_n(smem_u32(&tfull[a]), 1);
      mbar_init_n(smem_u32(&tempty[a]), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
  }
  if (warp == 1) umma::tmem_alloc(smem_u32(tmem_slot), 512);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------- TMA producer (B)
    if (lane == 0) {
      uint32_t g = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < nchunks; ++kc, ++g) {
          const int s = g % kJStages;
          mbar_wait_parity(smem_u32(&empty[s]), ((g / kJStages) & 1) ^ 1);
          const uint32_t bar = smem_u32(&full[s]);
          const uint32_t dst = smem_u32(base) + s * stage_bytes + 2 * a_bytes;
          mbar_expect_tx(bar, 2 * b_bytes);
          tma_2d(dst, &map_hi, kc * 64, 0, bar);
          tma_2d(dst + 256 * 128, &map_lo, kc * 64, 0, bar);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_bf16(128, V);
      uint32_t g = 0, it = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1;
        mbar_wait_parity(smem_u32(&tempty[acc]), ((it >> 1) & 1) ^ 1);
        umma::fence_after_thread_sync();
        const uint32_t d = tmem + acc * 256;
        for (int kc = 0; kc < nchunks; ++kc, ++g) {
          const int s = g % kJStages;
          mbar_wait_parity(smem_u32(&full[s]), (g / kJStages) & 1);
          umma::fence_after_thread_sync();
          const uint32_t sa = smem_u32(base) + s * stage_bytes;
          const uint32_t sb = sa + 2 * a_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dah = umma::make_smem_desc_sw128(sa + k * 32);
            const uint64_t dal = umma::make_smem_desc_sw128(sa + a_bytes + k * 32);
            const uint64_t dbh = umma::make_smem_desc_sw128(sb + k * 32);
            const uint64_t dbl = umma::make_smem_desc_sw128(sb + 256 * 128 + k * 32);
            umma::mma_bf16(d, dah, dbh, idesc, (kc | k) > 0);
            umma::mma_bf16(d, dah, dbl, idesc, 1);
            umma::mma_bf16(d, dal, dbh, idesc, 1);
          }
          umma::commit(smem_u32(&empty[s]));           // stage reusable once these MMAs retire
        }
        umma::commit(smem_u32(&tfull[acc]));           // accumulator complete
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;                          // TMEM lane quadrant of this warp
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      mbar_wait_parity(smem_u32(&tfull[acc]), (it >> 1) & 1);
      umma::fence_after_thread_sync();
      const long long m0 = tile * 128 + quad * 32;
      // The accumulator arrives with lane = row and 32 consecutive columns per thread.  Each
      // thread writes its 128 bytes into a [32 x 128 B] SWIZZLE_128B staging tile (conflict-
      // free 128-bit stores) and ONE bulk tensor store per block moves it to global memory --
      // full lines, rows past M clipped by the hardware, and no load / store instruction of
      // the output goes through the LSU (it is the L1 data pipe that bounds this kernel).
      unsigned char* stage_tile = s_out + quad * 4096;
      for (int c0 = 0; c0 < V; c0 += 32) {
        float v[32];
        umma::tmem_ld32(tmem + acc * 256 + ((uint32_t)(quad * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + j);
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
        // the previous block's bulk store must have finished READING the staging tile
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4*>(stage_tile + umma::swizzled_offset(lane, k)) =
              make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && m0 < p.M) {
          asm volatile(
              "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                  &map_out),
              "r"(c0), "r"((int)m0), "r"(smem_u32(stage_tile))
              : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      umma::fence_before_thread_sync();
      mbar_arrive(smem_u32(&tempty[acc]));
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // -------------------------------------------------------------- A producers
    // 8 lanes cover one row's 64-wide K chunk (256 contiguous bytes of pc / pf, one
    // 16-byte bf16 chunk per lane), 4 rows per warp, 4 passes over the 128 rows: global
    // loads are fully coalesced and the 8 lanes of a row write 8 distinct swizzled chunks.
    const int pw = warp - 6;                            // 0 .. kFProdWarps-1
    const int ch = lane & 7, rsub = lane >> 3;
    uint32_t g = 0;
    // 256-bit loads: 8 lanes read 256 contiguous bytes of a row with ONE wavefront pair (two
    // 128-bit loads per lane touch every line twice -- the L1 data pipe, shared with the
    // tensor core's operand reads, is what bounds this kernel).  With C >= 128 a tile of 128
    // consecutive joint rows touches at most two frames: pf is then loaded once per chunk
    // for both of them instead of once per pass.
    auto run = [&](auto two_frames_tag) {
      constexpr bool TWO = decltype(two_frames_tag)::value;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const float* pc_row[kFPasses];
        const float* pf_row[kFPasses];
        bool valid[kFPasses], second[kFPasses];
        float bacc[kFPasses];
        const long long n_first = tile * 128 / p.C;
        const long long n_last = min(tile * 128 + 127, p.M - 1) / p.C;
#pragma unroll
        for (int q = 0; q < kFPasses; ++q) {
          bacc[q] = 0.f;
          const long long m = tile * 128 + q * (kFProdWarps * 4) + pw * 4 + rsub;
          valid[q] = m < p.M;
          const long long n = valid[q] ? m / p.C : 0;
          const int c = valid[q] ? (int)(m - n * p.C) : 0;
          second[q] = n != n_first;
          pc_row[q] = p.pc + (size_t)c * H + ch * 8;
          pf_row[q] = p.pf + (size_t)n * H + ch * 8;
        }
        const float* pf_a = p.pf + (size_t)n_first * H + ch * 8;
        const float* pf_b = p.pf + (size_t)n_last * H + ch * 8;
        for (int kc = 0; kc < nchunks; ++kc, ++g) {
          const int s = g % kJStages;
          uint4 hi[kFPasses], lo[kFPasses];
          float a[kFPasses][8], fa[8], fb[8];
#pragma unroll
          for (int q = 0; q < kFPasses; ++q) ldg_cached8(pc_row[q] + kc * 64, a[q]);
          if (TWO) {
            ldg_cached8(pf_a + kc * 64, fa);
            ldg_cached8(pf_b + kc * 64, fb);
          }
#pragma unroll
          for (int q = 0; q < kFPasses; ++q) {
            float t[8];
            if (TWO) {
#pragma unroll
              for (int e = 0; e < 8; ++e) t[e] = second[q] ? fb[e] : fa[e];
            } else {
              ldg_cached8(pf_row[q] + kc * 64, t);
            }
            const float4 w0 = *reinterpret_cast<const float4*>(s_wb + kc * 64 + ch * 4);
            const float4 w1 = *reinterpret_cast<const float4*>(s_wb + kc * 64 + 32 + ch * 4);
            const float wb[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              t[e] = valid[q] ? tanh_from_exp(a[q][e], t[e]) : 0.f;
              bacc[q] = fmaf(t[e], wb[e], bacc[q]);
            }
            umma::split_pack8(t, hi[q], lo[q]);
          }
          mbar_wait_parity(smem_u32(&empty[s]), ((g / kJStages) & 1) ^ 1);
          unsigned char* a_hi = base + s * stage_bytes;
          unsigned char* a_lo = a_hi + a_bytes;
#pragma unroll
          for (int q = 0; q < kFPasses; ++q) {
            const uint32_t off = umma::swizzled_offset(q * (kFProdWarps * 4) + pw * 4 + rsub, ch);
            *reinterpret_cast<uint4*>(a_hi + off) = hi[q];
            *reinterpret_cast<uint4*>(a_lo + off) = lo[q];
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(smem_u32(&full[s]));
        }
#pragma unroll
        for (int q = 0; q < kFPasses; ++q) {
          float b = bacc[q];
          b += __shfl_xor_sync(0xffffffffu, b, 1);
          b += __shfl_xor_sync(0xffffffffu, b, 2);
          b += __shfl_xor_sync(0xffffffffu, b, 4);
          if (valid[q] && ch == 0)
            p.blank[tile * 128 + q * (kFProdWarps * 4) + pw * 4 + rsub] = b + __ldg(p.b_blank);
        }
      }
    };
    if (p.C >= 128) run(std::true_type{}); else run(std::false_type{});
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}


// ===========================================================================
// JointWeightFn backward, part 1 (dgrad) on tcgen05:
//   Gp[m, j] = (sum_v G[m, v] * W_vocab[v, j] + gb[m] * w_blank[j]) * (1 - h[m, j]^2)
// with h = tanh(pc[c] + pf[n]) recomputed in the epilogue.  Same pipeline as the
// forward kernel; a work unit is (128-row tile, block of NH <= 256 hidden columns).
//   A operand: G rows (fp32 from HBM) split hi/lo on the fly, K = V in chunks of 64
//   B operand: W_vocab^T [H, V] bf16 hi/lo (pre-split, TMA, SWIZZLE_128B)
// Gp [M, H] is written to a caller-provided workspace and reduced into
// grad_proj_ctx (sum over frames) / grad_proj_frame (sum over context states) by
// joint_reduce_kernel (one streaming pass).
// ===========================================================================
struct JointDgradParams {
  const float* pc;       // [C, H]
  const float* pf;       // [N, H]
  const float* w_blank;  // [H]
  const float* gl;       // [M, V]  grad_lexical
  const float* gb;       // [M]     grad_blank
  long long M;
  int C, H, V, NH;       // NH: hidden columns per unit (<= 256, multiple of 32)
  float* gp;             // [M, H]
};

__global__ void transpose_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ lo, int V, int H) {
  // w [V, H] -> hi/lo [H, V]
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < V * H; i += gridDim.x * blockDim.x) {
    const int j = i / V, v = i % V;
    umma::split_bf16(w[(size_t)v * H + j], hi[i], lo[i]);
  }
}

__global__ void __launch_bounds__(kJThreads, 1)
joint_dgrad_tc_kernel(const __grid_constant__ CUtensorMap map_hi,
                      const __grid_constant__ CUtensorMap map_lo, const JointDgradParams p) {
  extern __shared__ __align__(1024) unsigned char dsmem_raw[];
  unsigned char* base = dsmem_raw + ((1024u - (smem_u32(dsmem_raw) & 1023u)) & 1023u);
  const int V = p.V, H = p.H, NH = p.NH;
  const uint32_t a_bytes = 128 * 128;
  const uint32_t b_bytes = (uint32_t)NH * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * 256 * 128;
  float* s_wb = reinterpret_cast<float*>(base + kJStages * stage_bytes);   // [H]
  float* s_tr = s_wb + H;                                                  // 4 x [32][33]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_tr + 4 * 32 * 33);
  uint64_t* full = bars;
  uint64_t* empty = bars + kJStages;
  uint64_t* tfull = bars + 2 * kJStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = V / 64;
  const int nblk = H / NH;
  const long long num_units = ((p.M + 127) / 128) * nblk;

  for (int i = tid; i < H; i += kJThreads) s_wb[i] = p.w_blank[i];
  if (tid == 0) {
    for (int s = 0; s < kJStages; ++s) {
      mbar_init_n(smem_u32(&full[s]), kJProducers + 1);
      mbar_init_n(smem_u32(&empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init_n(smem_u32(&tfull[a]), 1);
      mbar_init_n(smem_u32(&tempty[a]), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
  }
  if (warp == 1) umma::tmem_alloc(smem_u32(tmem_slot), 512);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (long long unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int blk = (int)(unit % nblk);
        for (int kc = 0; kc < nchunks; ++kc, ++g) {
          const int s = g % kJStages;
          mbar_wait_parity(smem_u32(&empty[s]), ((g / kJStages) & 1) ^ 1);
          const uint32_t bar = smem_u32(&full[s]);
          const uint32_t dst = smem_u32(base) + s * stage_bytes + 2 * a_bytes;
          mbar_expect_tx(bar, 2 * b_bytes);
          tma_2d(dst, &map_hi, kc * 64, blk * NH, bar);
          tma_2d(dst + 256 * 128, &map_lo, kc * 64, blk * NH, bar);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_bf16(128, NH);
      uint32_t g = 0, it = 0;
      for (long long unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const uint32_t acc = it & 1;
        mbar_wait_parity(smem_u32(&tempty[acc]), ((it >> 1) & 1) ^ 1);
        umma::fence_after_thread_sync();
        const uint32_t d = tmem + acc * 256;
        for (int kc = 0; kc < nchunks; ++kc, ++g) {
          const int s = g % kJStages;
          mbar_wait_parity(smem_u32(&full[s]), (g / kJStages) & 1);
          umma::fence_after_thread_sync();
          const uint32_t sa = smem_u32(base) + s * stage_bytes;
          const uint32_t sb = sa + 2 * a_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dah = umma::make_smem_desc_sw128(sa + k * 32);
            const uint64_t dal = umma::make_smem_desc_sw128(sa + a_bytes + k * 32);
            const uint64_t dbh = umma::make_smem_desc_sw128(sb + k * 32);
            const uint64_t dbl = umma::make_smem_desc_sw128(sb + 256 * 128 + k * 32);
            umma::mma_bf16(d, dah, dbh, idesc, (kc | k) > 0);
            umma::mma_bf16(d, dah, dbl, idesc, 1);
            umma::mma_bf16(d, dal, dbh, idesc, 1);
          }
          umma::commit(smem_u32(&empty[s]));
        }
        umma::commit(smem_u32(&tfull[acc]));
      }
    }
  } else if (warp < 6) {
    // epilogue: thread = joint row (TMEM lane).  Only the rank-1 blank term is added here;
    // the tanh' factor is applied by the reduction kernel, which streams this buffer anyway
    // and has the warps to hide the pc / pf load latency (with it in this epilogue, one
    // warp per TMEM quadrant serialised load -> tanh -> store and stalled the MMA pipe).
    const int quad = warp & 3;
    uint32_t it = 0;
    for (long long unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const long long tile = unit / nblk;
      const int blk = (int)(unit % nblk);
      const uint32_t acc = it & 1;
      const long long m0 = tile * 128 + quad * 32;
      const long long m = m0 + lane;
      const int rows_valid = (int)max(0ll, min(32ll, p.M - m0));
      const float gbm = m < p.M ? p.gb[m] : 0.f;
      const float* wb = s_wb + blk * NH;
      float* out = p.gp + (size_t)m0 * H + blk * NH;
      mbar_wait_parity(smem_u32(&tfull[acc]), (it >> 1) & 1);
      umma::fence_after_thread_sync();
      for (int c0 = 0; c0 < NH; c0 += 32) {
        float v[32];
        umma::tmem_ld32(tmem + acc * 256 + ((uint32_t)(quad * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(gbm, wb[c0 + j], v[j]);
        store_block_coalesced(v, s_tr + quad * (32 * 33), lane, out + c0, (size_t)H, rows_valid);
      }
      umma::fence_before_thread_sync();
      mbar_arrive(smem_u32(&tempty[acc]));
    }
  } else {
    // A producers: grad_lexical rows, hi/lo split, swizzled K-major stores.  8 lanes per
    // row (256 contiguous bytes per K chunk), 4 rows per warp, 4 passes: coalesced loads.
    // The loads of chunk i+1 are issued BEFORE chunk i is converted and stored, so the HBM
    // latency of the gradient stream is hidden behind the conversion work.
    const int pw = warp - 6;
    const int ch = lane & 7, rsub = lane >> 3;
    auto issue = [&](long long unit, int kc, float4 (&x)[4][2]) {
      const long long tile = unit / nblk;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long m = tile * 128 + q * 32 + pw * 4 + rsub;
        x[q][0] = x[q][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M) {
          const float* src = p.gl + (size_t)m * V + kc * 64 + ch * 8;
          x[q][0] = ldg_stream4(src);
          x[q][1] = ldg_stream4(src + 4);
        }
      }
    };
    float4 cur[4][2], nxt[4][2];
    long long unit = blockIdx.x;
    int kc = 0;
    if (unit < num_units) issue(unit, 0, cur);
    uint32_t g = 0;
    while (unit < num_units) {
      long long nunit = unit;
      int nkc = kc + 1;
      if (nkc == nchunks) { nkc = 0; nunit += gridDim.x; }
      if (nunit < num_units) issue(nunit, nkc, nxt);
      const int s = g % kJStages;
      uint4 hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float x[8] = {cur[q][0].x, cur[q][0].y, cur[q][0].z, cur[q][0].w,
                            cur[q][1].x, cur[q][1].y, cur[q][1].z, cur[q][1].w};
        umma::split_pack8(x, hi[q], lo[q]);
      }
      mbar_wait_parity(smem_u32(&empty[s]), ((g / kJStages) & 1) ^ 1);
      unsigned char* a_hi = base + s * stage_bytes;
      unsigned char* a_lo = a_hi + a_bytes;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = umma::swizzled_offset(q * 32 + pw * 4 + rsub, ch);
        *reinterpret_cast<uint4*>(a_hi + off) = hi[q];
        *reinterpret_cast<uint4*>(a_lo + off) = lo[q];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(smem_u32(&full[s]));
#pragma unroll
      for (int q = 0; q < 4; ++q) { cur[q][0] = nxt[q][0]; cur[q][1] = nxt[q][1]; }
      unit = nunit; kc = nkc; ++g;
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

// Wide variant for H % 128 == 0: a warp owns a row (512 contiguous bytes, float4 per
// lane), 16 warps sweep the C rows of a frame with 4 independent loads in flight per
// thread; the [C, 128] column block of grad_proj_ctx is accumulated in shared memory.
__global__ void __launch_bounds__(512)
joint_reduce128_kernel(const float* __restrict__ gp, const float* __restrict__ pc,
                       const float* __restrict__ pf, long long N, int C, int H,
                       long long frames_per_block, float* __restrict__ g_pc,
                       float* __restrict__ g_pf) {
  extern __shared__ __align__(16) float4 r4[];    // [C][32] accumulators, then [16][32] partials
  float4* part = r4 + (size_t)C * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int jcol = blockIdx.x * 128 + lane * 4;
  const long long n_lo = (long long)blockIdx.y * frames_per_block;
  const long long n_hi = min(N, n_lo + frames_per_block);
  for (int i = threadIdx.x; i < C * 32; i += 512) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (long long n = n_lo; n < n_hi; ++n) {
    const float* base = gp + ((size_t)n * C) * H + jcol;
    const float4 fr = __ldg(reinterpret_cast<const float4*>(pf + (size_t)n * H + jcol));
    float4 pf = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = warp; c < C; c += 64) {
      float4 x[4], a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cc = c + 16 * u;
        x[u] = a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cc < C) {
          x[u] = ldg_stream4(base + (size_t)cc * H);
          a[u] = __ldg(reinterpret_cast<const float4*>(pc + (size_t)cc * H + jcol));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cc = c + 16 * u;
        if (cc < C) {
          // d tanh: Gp = (G.W + gb*wb) * (1 - tanh(pc + pf)^2)
          const float h0 = tanh_fast(a[u].x + fr.x), h1 = tanh_fast(a[u].y + fr.y);
          const float h2 = tanh_fast(a[u].z + fr.z), h3 = tanh_fast(a[u].w + fr.w);
          x[u].x *= fmaf(-h0, h0, 1.f); x[u].y *= fmaf(-h1, h1, 1.f);
          x[u].z *= fmaf(-h2, h2, 1.f); x[u].w *= fmaf(-h3, h3, 1.f);
          pf.x += x[u].x; pf.y += x[u].y; pf.z += x[u].z; pf.w += x[u].w;
          float4 a = r4[cc * 32 + lane];          // (cc, lane) has exactly one owner thread
          a.x += x[u].x; a.y += x[u].y; a.z += x[u].z; a.w += x[u].w;
          r4[cc * 32 + lane] = a;
        }
      }
    }
    part[warp * 32 + lane] = pf;
    __syncthreads();
    if (warp == 0) {
      float4 s = part[lane];
#pragma unroll
      for (int w = 1; w < 16; ++w) {
        const float4 o = part[w * 32 + lane];
        s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
      }
      float4* out = reinterpret_cast<float4*>(g_pf + (size_t)n * H + jcol);
      float4 cur = *out;                          // single owner of (n, jcol..jcol+3)
      cur.x += s.x; cur.y += s.y; cur.z += s.z; cur.w += s.w;
      *out = cur;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < C * 32; i += 512) {
    const int c = i >> 5, l = i & 31;
    const float4 a = r4[i];
    float* dst = g_pc + (size_t)c * H + blockIdx.x * 128 + l * 4;
    atomicAdd(dst, a.x); atomicAdd(dst + 1, a.y); atomicAdd(dst + 2, a.z); atomicAdd(dst + 3, a.w);
  }
}

// Gp [N, C, H] -> grad_proj_frame [N, H] (sum over c, written) and grad_proj_ctx
// [C, H] (sum over n, accumulated with atomics once per CTA).  grid = (H / jw, nblocks);
// block = 512 threads = (512 / jw) row groups x jw columns; smem = C * jw floats.
__global__ void __launch_bounds__(512)
joint_reduce_kernel(const float* __restrict__ gp, const float* __restrict__ pc,
                    const float* __restrict__ pf, long long N, int C, int H, int jw,
                    long long frames_per_block, float* __restrict__ g_pc,
                    float* __restrict__ g_pf) {
  extern __shared__ float racc[];                 // [C][jw] then [RG][jw] scratch
  const int RG = 512 / jw;
  float* pf_part = racc + (size_t)C * jw;
  const int jj = threadIdx.x % jw, rg = threadIdx.x / jw;
  const int j = blockIdx.x * jw + jj;
  const long long n_lo = (long long)blockIdx.y * frames_per_block;
  const long long n_hi = min(N, n_lo + frames_per_block);
  for (int i = threadIdx.x; i < C * jw; i += 512) racc[i] = 0.f;
  __syncthreads();
  for (long long n = n_lo; n < n_hi; ++n) {
    const float* row = gp + ((size_t)n * C) * H + j;
    const float fr = pf[(size_t)n * H + j];
    float acc = 0.f;
    for (int c = rg; c < C; c += RG) {
      const float h = tanh_fast(pc[(size_t)c * H + j] + fr);
      const float x = ldg_stream(row + (size_t)c * H) * fmaf(-h, h, 1.f);
      acc += x;
      racc[c * jw + jj] += x;                     // (c, jj) is owned by exactly one thread
    }
    pf_part[rg * jw + jj] = acc;
    __syncthreads();
    if (rg == 0) {
      float s = 0.f;
      for (int r = 0; r < RG; ++r) s += pf_part[r * jw + jj];
      g_pf[(size_t)n * H + j] += s;               // single owner of (n, j)
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < C * jw; i += 512) {
    const int c = i / jw, q = i % jw;
    atomicAdd(g_pc + (size_t)c * H + blockIdx.x * jw + q, racc[i]);
  }
}


// ===========================================================================
// JointWeightFn backward, part 2 (wgrad) on tcgen05:
//   grad_W_vocab[v, j] = sum_m G[m, v] * h[m, j]        h = tanh(pc[c] + pf[n])
// The reduction runs over ALL M = N*C joint rows, so both operands are contiguous
// along M / N (G rows along v, h rows along j): MN-major SWIZZLE_128B tiles, written
// by the producer warps (G split hi/lo; h recomputed, split hi/lo), K = 32 rows per
// stage.  A CTA owns one block of NJ <= 256 hidden columns and a contiguous range of
// rows; its [V, NJ] fp32 accumulator lives in TMEM for the whole range (V / 128
// accumulators of 128 lanes x NJ columns) and is added to global memory once.
// The same producer threads accumulate grad_b_vocab, grad_w_blank and grad_b_blank.
// ===========================================================================
constexpr int kWThreads = 288;          // warp 0: MMA issuer; warps 1-8: producers (8 warps x ~190
                                        // registers: room for the one-stage-ahead gradient prefetch)
constexpr int kWProducers = 256;
constexpr int kWStages = 6;
constexpr int kWK = 16;                 // joint rows per stage

struct JointWgradParams {
  const float* pc;       // [C, H]  e^(2 proj_ctx)   (joint_exp_table_kernel)
  const float* pf;       // [N, H]  e^(2 proj_frame)
  const float* gl;       // [M, V]
  const float* gb;       // [M]
  long long M, rows_per_cta;
  int C, H, V, NJ;
  float* gwv;            // [V, H]
  float* gwb;            // [H]
  float* gbv;            // [V]
  float* gbb;            // [1]
};

// AI / BI: 16-byte operand chunks per producer thread and stage (V / 128, NJ / 128)
// SPLIT: grad_lexical rows are [V bf16 hi | V bf16 lo] (lt_lattice_backward with
// LT_FLAG_GRAD_SPLIT): the A operand is copied, not converted.
template <int AI, int BI, bool SPLIT>
__global__ void __launch_bounds__(kWThreads, 1)
joint_wgrad_tc_kernel(const JointWgradParams p) {
  extern __shared__ __align__(1024) unsigned char wsmem_raw[];
  unsigned char* base = wsmem_raw + ((1024u - (smem_u32(wsmem_raw) & 1023u)) & 1023u);
  const int V = p.V, H = p.H, NJ = p.NJ;
  const uint32_t op_bytes = 256 * kWK * 2;            // one [256 x 32] bf16 operand tile (max)
  const uint32_t stage_bytes = 4 * op_bytes;          // A_hi | A_lo | B_hi | B_lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kWStages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWStages;
  uint64_t* done = bars + 2 * kWStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nj = H / NJ;
  const int jblk = blockIdx.x % nj;
  const long long m_lo = (long long)(blockIdx.x / nj) * p.rows_per_cta;
  const long long m_hi = min(p.M, m_lo + p.rows_per_cta);
  if (m_lo >= m_hi) return;                           // uniform for the CTA
  const int nchunks = (int)((m_hi - m_lo + kWK - 1) / kWK);

  if (tid == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init_n(smem_u32(&full[s]), kWProducers);
      mbar_init_n(smem_u32(&empty[s]), 1);
    }
    mbar_init_n(smem_u32(done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) umma::tmem_alloc(smem_u32(tmem_slot), 512);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;
  const int nvb = V / 128;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_bf16_mn(128, NJ);
      const uint32_t a_sbo = (uint32_t)(V / 64) * 1024, b_sbo = (uint32_t)(NJ / 64) * 1024;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int s = ch % kWStages;
        mbar_wait_parity(smem_u32(&full[s]), (ch / kWStages) & 1);
        umma::fence_after_thread_sync();
        const uint32_t sa = smem_u32(base) + s * stage_bytes;
        const uint32_t sb = sa + 2 * op_bytes;
#pragma unroll
        for (int ks = 0; ks < kWK / 16; ++ks) {
          const uint64_t dbh = umma::make_smem_desc_mn_sw128(sb + 2 * ks * b_sbo, 1024, b_sbo);
          const uint64_t dbl = umma::make_smem_desc_mn_sw128(sb + op_bytes + 2 * ks * b_sbo, 1024, b_sbo);
          for (int vb = 0; vb < nvb; ++vb) {
            const uint32_t aoff = 2 * ks * a_sbo + vb * 2 * 1024;
            const uint64_t dah = umma::make_smem_desc_mn_sw128(sa + aoff, 1024, a_sbo);
            const uint64_t dal = umma::make_smem_desc_mn_sw128(sa + op_bytes + aoff, 1024, a_sbo);
            const uint32_t d = tmem + vb * 256;
            umma::mma_bf16(d, dah, dbh, idesc, (ch | ks) > 0);
            umma::mma_bf16(d, dah, dbl, idesc, 1);
            umma::mma_bf16(d, dal, dbh, idesc, 1);
          }
        }
        umma::commit(smem_u32(&empty[s]));
      }
      umma::commit(smem_u32(done));
    }
  } else {
    const int pidx = tid - 32;                         // 0 .. 255
    // V = 128 * AI and NJ = 128 * BI (the launcher picks the instantiation), so every per-row
    // stride below is a compile-time constant: a thread's operand rows are 16 / AI (16 / BI)
    // joint rows apart = 2048 floats of grad_lexical, and 4096 bytes apart in the MN-major tile.
    constexpr int kAStep = kWK / AI, kBStep = kWK / BI;
    const int a_vch = pidx % (16 * AI), a_k0 = pidx / (16 * AI);
    const int b_jch = pidx % (16 * BI), b_k0 = pidx / (16 * BI);
    const int j0 = jblk * NJ + b_jch * 8;
    const uint32_t a_off0 = umma::mn_major_chunk_offset(128 * AI, a_vch * 8, a_k0);
    const uint32_t b_off0 = umma::mn_major_chunk_offset(128 * BI, b_jch * 8, b_k0);
    const int C = p.C;
    float bv_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float wb_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float bb_acc = 0.f;
    const float* ga = p.gl + (size_t)(m_lo + a_k0) * (128 * AI) + a_vch * 8;   // stage 0, row 0
    const float* pcj = p.pc + j0;
    const float* pfj = p.pf + j0;
    const int nfull = (int)((m_hi - m_lo) / kWK);      // stages whose kWK rows are all live

    // Everything a stage consumes from global memory is loaded ONE STAGE AHEAD into registers
    // with 256-bit loads (grad_lexical rows = the HBM stream, pc rows and grad_blank = L2
    // hits).  pf[n, j0 .. j0+7] is the same for the C consecutive joint rows of a frame: it
    // lives in registers -- pfa for the frame of the stage's first row, pfb for the next frame
    // (C >= 32 > kWK: a stage touches at most two frames) -- and is reloaded when the first
    // row crosses into a new frame.  Full stages (all but possibly the last one of the last
    // CTA) take a path with no bounds checks at all.
    // (fn, fc): frame and context state of the first row of the stage being PREFETCHED;
    // (cn, cc): the same for the stage being converted.
    const long long nframes = p.M / C;
    long long fn = m_lo / C;
    int fc = (int)(m_lo - fn * C);
    long long cn = fn;
    int cc = fc;
    float pfa[8], pfb[8];
    auto load_pf = [&](long long n, float (&v)[8]) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (n < nframes) ldg_cached8(pfj + (size_t)n * H, v);
    };
    load_pf(cn, pfa);
    load_pf(cn + 1, pfb);
    struct Pre {                       // one stage's worth of prefetched operands
      float ax[AI][8];                 // grad_lexical (SPLIT: words 0-3 = hi pairs, 4-7 = lo pairs)
      float bp[BI][8];                 // pc
      float gbm[BI];                   // grad_blank
    };
    auto prefetch = [&](int ch, Pre& q, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      const long long mrow0 = m_lo + (long long)ch * kWK;
      const float* src = ga + (size_t)ch * (kWK * 128 * AI);
#pragma unroll
      for (int i = 0; i < AI; ++i) {
        if (!FULL) {
#pragma unroll
          for (int e = 0; e < 8; ++e) q.ax[i][e] = 0.f;
        }
        if (FULL || (ch < nchunks && mrow0 + a_k0 + i * kAStep < m_hi)) {
          if (SPLIT) {
            // this thread's 8 elements: 16 bytes of the hi half, 16 bytes of the lo half
            const unsigned char* row = reinterpret_cast<const unsigned char*>(src + i * 2048) -
                                       a_vch * 32 + a_vch * 16;
            const float4 h = ldg_stream4(reinterpret_cast<const float*>(row));
            const float4 l = ldg_stream4(reinterpret_cast<const float*>(row + V * 2));
            q.ax[i][0] = h.x; q.ax[i][1] = h.y; q.ax[i][2] = h.z; q.ax[i][3] = h.w;
            q.ax[i][4] = l.x; q.ax[i][5] = l.y; q.ax[i][6] = l.z; q.ax[i][7] = l.w;
          } else {
            ldg_stream8(src + i * 2048, q.ax[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < BI; ++i) {
        int c = fc + b_k0 + i * kBStep;
        if (c >= C) c -= C;
        if (!FULL) {
#pragma unroll
          for (int e = 0; e < 8; ++e) q.bp[i][e] = 0.f;
          q.gbm[i] = 0.f;
        }
        const long long m = mrow0 + b_k0 + i * kBStep;
        if (FULL || (ch < nchunks && m < m_hi)) {
          ldg_cached8(pcj + (size_t)c * H, q.bp[i]);
          q.gbm[i] = __ldg(p.gb + m);
        }
      }
      fc += kWK;
      if (fc >= C) { fc -= C; ++fn; }
    };
    Pre bufa, bufb;                    // ping-pong: no register moves between stages
    if (nfull > 0) prefetch(0, bufa, std::true_type{}); else prefetch(0, bufa, std::false_type{});

    auto stage = [&](int ch, Pre& cur, Pre& nxt, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      const int s = ch % kWStages;
      if (ch + 1 < nfull) prefetch(ch + 1, nxt, std::true_type{});
      else prefetch(ch + 1, nxt, std::false_type{});
      mbar_wait_parity(smem_u32(&empty[s]), ((ch / kWStages) & 1) ^ 1);
      unsigned char* a_hi = base + s * stage_bytes + a_off0;
      unsigned char* b_hi = base + s * stage_bytes + 2 * op_bytes + b_off0;
#pragma unroll
      for (int i = 0; i < AI; ++i) {                   // A = G^T chunk
        uint4 h4, l4;
        if (SPLIT) {
          h4 = make_uint4(__float_as_uint(cur.ax[i][0]), __float_as_uint(cur.ax[i][1]),
                          __float_as_uint(cur.ax[i][2]), __float_as_uint(cur.ax[i][3]));
          l4 = make_uint4(__float_as_uint(cur.ax[i][4]), __float_as_uint(cur.ax[i][5]),
                          __float_as_uint(cur.ax[i][6]), __float_as_uint(cur.ax[i][7]));
          const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {           // value = hi + lo (bf16 -> fp32 by a shift)
            bv_acc[2 * e] += __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
            bv_acc[2 * e + 1] += __uint_as_float(hw[e] & 0xffff0000u) +
                                 __uint_as_float(lw[e] & 0xffff0000u);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) bv_acc[e] += cur.ax[i][e];
          umma::split_pack8(cur.ax[i], h4, l4);
        }
        *reinterpret_cast<uint4*>(a_hi + i * 4096) = h4;
        *reinterpret_cast<uint4*>(a_hi + op_bytes + i * 4096) = l4;
      }
#pragma unroll
      for (int i = 0; i < BI; ++i) {                   // B = h^T chunk (recomputed)
        const bool live = FULL || m_lo + (long long)ch * kWK + b_k0 + i * kBStep < m_hi;
        const bool wrapped = cc + b_k0 + i * kBStep >= C;       // warp-uniform
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          t[e] = live ? tanh_from_exp(cur.bp[i][e], wrapped ? pfb[e] : pfa[e]) : 0.f;
          wb_acc[e] = fmaf(cur.gbm[i], t[e], wb_acc[e]);
        }
        if (b_jch == 0) bb_acc += cur.gbm[i];
        uint4 h4, l4;
        umma::split_pack8(t, h4, l4);
        *reinterpret_cast<uint4*>(b_hi + i * 4096) = h4;
        *reinterpret_cast<uint4*>(b_hi + op_bytes + i * 4096) = l4;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(smem_u32(&full[s]));
      cc += kWK;
      if (cc >= C) {                                   // the next stage starts in a new frame
        cc -= C; ++cn;
#pragma unroll
        for (int e = 0; e < 8; ++e) pfa[e] = pfb[e];
        load_pf(cn + 1, pfb);
      }
    };
    int ch = 0;
#pragma unroll 1
    for (; ch + 1 < nfull; ch += 2) {
      stage(ch, bufa, bufb, std::true_type{});
      stage(ch + 1, bufb, bufa, std::true_type{});
    }
    // at most one full stage left, then at most one partial stage (ch is even here)
    if (ch < nfull) { stage(ch, bufa, bufb, std::true_type{}); ++ch; }
    if (ch < nchunks) {
      if (ch & 1) stage(ch, bufb, bufa, std::false_type{});
      else stage(ch, bufa, bufb, std::false_type{});
    }
    // side sums: bias / blank-projection gradients
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(p.gwb + j0 + e, wb_acc[e]);
    if (jblk == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(p.gbv + a_vch * 8 + e, bv_acc[e]);
      if (b_jch == 0) atomicAdd(p.gbb, bb_acc);
    }
  }
  // epilogue: warps 0-3 (TMEM lane quadrant = warp) add the accumulators to global memory
  __syncthreads();
  if (warp < 4) {
    mbar_wait_parity(smem_u32(done), 0);
    umma::fence_after_thread_sync();
    for (int vb = 0; vb < nvb; ++vb) {
      float* out = p.gwv + (size_t)(vb * 128 + warp * 32 + lane) * H + jblk * NJ;
      for (int c0 = 0; c0 < NJ; c0 += 32) {
        float v[32];
        umma::tmem_ld32(tmem + vb * 256 + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(out + c0 + i, v[i]);
      }
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace lt

namespace lt {

typedef CUresult (*EncodeTiledFnJ)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnJ joint_encode_fn() {
  static EncodeTiledFnJ fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
          cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFnJ>(sym);
  return fn;
}

bool joint_tc_supported(int64_t N, int C, int H, int V, const void* pc, const void* pf,
                        const void* lexical) {
  if (option(OPT_JOINT_SIMT)) return false;
  if (V % 32 != 0 || V < 32 || V > 256) return false;
  if (H % 64 != 0 || H > 4096) return false;
  if (N * (int64_t)C < 1) return false;
  auto al = [](const void* q, int a) { return reinterpret_cast<uintptr_t>(q) % a == 0; };
  return al(pc, 32) && al(pf, 32) && al(lexical, 16);      // pc / pf: 256-bit loads
}

// workspace layout: [bf16 hi | lo of W_vocab (or its transpose)] [E_c [C,H] | E_f [N,H]] ...
int64_t joint_split_bytes(int H, int V) { return (((int64_t)V * H * 4 + 255) / 256) * 256; }
int64_t joint_table_bytes(int64_t N, int C, int H) {
  return (((N + C) * (int64_t)H * 4 + 255) / 256) * 256;
}
int joint_split_weights_launch(const float* w, __nv_bfloat16* hi, __nv_bfloat16* lo, int n,
                               cudaStream_t stream) {
  split_weights_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, hi, lo, n);
  LT_LAUNCHED();
  return LT_OK;
}
int joint_exp_table_launch(const float* x, float* out, long long n, cudaStream_t stream) {
  joint_exp_table_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 4096), 256, 0, stream>>>(
      x, out, n);
  LT_LAUNCHED();
  return LT_OK;
}
// e^(2 pc) -> ec [C,H], e^(2 pf) -> ef [N,H]
int joint_exp_tables_launch(const float* pc, const float* pf, int64_t N, int C, int H,
                                   float* ec, float* ef, cudaStream_t stream) {
  const long long nc = (long long)C * H, nf = (long long)N * H;
  joint_exp_table_kernel<<<(unsigned)std::min<long long>((nc + 255) / 256, 4096), 256, 0, stream>>>(
      pc, ec, nc);
  LT_LAUNCHED();
  joint_exp_table_kernel<<<(unsigned)std::min<long long>((nf + 255) / 256, 4096), 256, 0, stream>>>(
      pf, ef, nf);
  LT_LAUNCHED();
  return LT_OK;
}

// lexical / blank of all M = N*C joint rows on tcgen05 (bf16x3 split, fp32 accumulate).
int joint_forward_tc_launch(const float* pc, const float* pf, const float* wb, const float* bb,
                            const float* wv, const float* bv, int64_t N, int C, int H, int V,
                            float* blank, float* lexical, void* workspace, cudaStream_t stream) {
  EncodeTiledFnJ encode = joint_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled is unavailable in this driver"); return LT_ERR_CUDA; }
  // caller-provided scratch for the bf16 hi / lo split of W_vocab (V*H*4 bytes)
  __nv_bfloat16* whi = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* wlo = whi + (size_t)V * H;
  split_weights_kernel<<<(V * H + 255) / 256, 256, 0, stream>>>(wv, whi, wlo, V * H);
  LT_LAUNCHED();
  CUtensorMap map_hi, map_lo;
  cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)V};
  cuuint64_t strides[1] = {(cuuint64_t)H * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)V};
  cuuint32_t estr[2] = {1, 1};
  for (int i = 0; i < 2; ++i) {
    CUresult r = encode(i == 0 ? &map_hi : &map_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        i == 0 ? whi : wlo, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (W_vocab) failed with %d", (int)r);
      return LT_ERR_CUDA;
    }
  }
  float* ec = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) +
                                       joint_split_bytes(H, V));
  float* ef = ec + (size_t)C * H;
  if (int rc = joint_exp_tables_launch(pc, pf, N, C, H, ec, ef, stream)) return rc;
  JointTcParams p = {};
  p.pc = ec; p.pf = ef; p.w_blank = wb; p.b_vocab = bv; p.b_blank = bb;
  p.M = (long long)N * C; p.C = C; p.H = H; p.V = V; p.blank = blank; p.lexical = lexical;
  // output map: lexical [M, V] fp32, box = one epilogue block (32 rows x 32 columns = 128 B rows)
  CUtensorMap map_out;
  {
    cuuint64_t odims[2] = {(cuuint64_t)V, (cuuint64_t)p.M};
    cuuint64_t ostrides[1] = {(cuuint64_t)V * 4};
    cuuint32_t obox[2] = {32, 32};
    cuuint32_t oestr[2] = {1, 1};
    CUresult r = encode(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, lexical, odims, ostrides,
                        obox, oestr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (lexical output) failed with %d", (int)r);
      return LT_ERR_CUDA;
    }
  }
  const size_t smem = (size_t)kJStages * (2 * 128 * 128 + 2 * 256 * 128) + 4 * 4096 +
                      sizeof(float) * (H + 256) + 16 * 8 + 16 + 1024;
  int dev = 0, sms = 0;
  LT_CUDA(cudaGetDevice(&dev));
  LT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tiles = (p.M + 127) / 128;
  const int grid = (int)(tiles < sms ? tiles : sms);
  LT_CUDA(cudaFuncSetAttribute(joint_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  joint_forward_tc_kernel<<<grid, kFThreads, smem, stream>>>(map_hi, map_lo, map_out, p);
  LT_LAUNCHED();
  return LT_OK;
}


bool joint_dgrad_tc_supported(int64_t N, int C, int H, int V, const void* gl, const void* pc,
                              const void* pf) {
  if (option(OPT_JOINT_SIMT)) return false;
  if (V % 64 != 0 || V < 64 || V > 4096) return false;
  if (H % 32 != 0 || H > 4096 || (H > 256 && H % 256 != 0)) return false;
  if (N * (int64_t)C < 1) return false;
  auto al = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  return al(gl) && al(pc) && al(pf);
}

static bool dgrad2_shape_ok(int H, int V) {
  return !option(OPT_JOINT_SIMT) && !option(OPT_JOINT_DGRAD_V1) && V % 64 == 0 && V >= 64 &&
         V <= 256 && H % 128 == 0 && H <= 4096;
}

// fp32 rows -> rows of [V bf16 hi | V bf16 lo] (the form the lattice backward kernel can emit
// directly): lt_joint_split_rows
__global__ void joint_split_rows_kernel(const float* __restrict__ g, unsigned char* __restrict__ out,
                                        long long M, int V) {
  const long long items = M * (V / 8);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / (V / 8);
    const int v0 = (int)(i % (V / 8)) * 8;
    float x[8];
    ldg_stream8(g + m * V + v0, x);
    uint4 hi, lo;
    umma::split_pack8(x, hi, lo);
    unsigned char* row = out + (size_t)m * V * 4;
    *reinterpret_cast<uint4*>(row + v0 * 2) = hi;
    *reinterpret_cast<uint4*>(row + V * 2 + v0 * 2) = lo;
  }
}
int joint_split_rows_launch(const float* g, void* out, int64_t M, int V, cudaStream_t stream) {
  joint_split_rows_kernel<<<4096, 256, 0, stream>>>(g, reinterpret_cast<unsigned char*>(out), M, V);
  LT_LAUNCHED();
  return LT_OK;
}
// split-format grad_lexical: the fused dgrad AND the tensor-core wgrad must both apply
bool joint_backward_split_supported(int64_t N, int C, int H, int V) {
  return dgrad2_shape_ok(H, V) && (V == 128 || V == 256) && H % 128 == 0 && H <= 4096 &&
         (H <= 256 || H % 256 == 0) && N >= 1 && C >= 32 && !option(OPT_JOINT_WGRAD_SIMT);
}

int64_t joint_backward_workspace_bytes(int64_t N, int C, int H, int V) {
  // bf16 hi / lo of W_vocab^T; the first-generation dgrad also needs the [M, H] buffer
  const int64_t split = joint_split_bytes(H, V) + joint_table_bytes(N, C, H);
  if (dgrad2_shape_ok(H, V))
    return split;
  return split + N * (int64_t)C * H * 4;
}

// dgrad on tcgen05 + streaming reduction into grad_proj_ctx / grad_proj_frame.
int joint_dgrad_tc_launch(const float* pc, const float* pf, const float* wb, const float* wv,
                          const float* gb, const float* gl, int split, int64_t N, int C, int H,
                          int V, float* gpc, float* gpf, void* workspace, cudaStream_t stream) {
  EncodeTiledFnJ encode = joint_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled is unavailable in this driver"); return LT_ERR_CUDA; }
  __nv_bfloat16* whi = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* wlo = whi + (size_t)H * V;
  float* ec = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) +
                                       joint_split_bytes(H, V));
  float* ef = ec + (size_t)C * H;
  float* gp = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ec) +
                                       joint_table_bytes(N, C, H));
  if (int rc = joint_exp_tables_launch(pc, pf, N, C, H, ec, ef, stream)) return rc;
  transpose_split_kernel<<<(V * H + 255) / 256, 256, 0, stream>>>(wv, whi, wlo, V, H);
  LT_LAUNCHED();
  if (joint_dgrad2_supported(N, C, H, V, gl, pc, pf)) {
    // second generation: fused tanh' + both reductions, no [M, H] round trip (joint_dgrad2.cu)
    CUtensorMap m_hi, m_lo;
    cuuint64_t dims2[2] = {(cuuint64_t)V, (cuuint64_t)H};
    cuuint64_t strides2[1] = {(cuuint64_t)V * 2};
    cuuint32_t box2[2] = {64, 128};
    cuuint32_t estr2[2] = {1, 1};
    for (int i = 0; i < 2; ++i) {
      CUresult r = encode(i == 0 ? &m_hi : &m_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                          i == 0 ? whi : wlo, dims2, strides2, box2, estr2,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (W_vocab^T, 128-row box) failed with %d", (int)r);
        return LT_ERR_CUDA;
      }
    }
    CUtensorMap m_g = m_hi;
    if (split) {
      // rows of [V bf16 hi | V bf16 lo]: {2V, C, N} bf16, one [128 frames x 1 state x 64] box
      cuuint64_t gd[3] = {(cuuint64_t)2 * V, (cuuint64_t)C, (cuuint64_t)N};
      cuuint64_t gs[2] = {(cuuint64_t)V * 4, (cuuint64_t)C * V * 4};
      cuuint32_t gbox[3] = {64, 1, (cuuint32_t)(joint_dgrad2_pair(H, V)
                                                    ? 64 : 128 / joint_dgrad2_multicast(H, V))};
      cuuint32_t ge[3] = {1, 1, 1};
      CUresult r = encode(&m_g, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<float*>(gl), gd,
                          gs, gbox, ge, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (split grad_lexical) failed with %d", (int)r);
        return LT_ERR_CUDA;
      }
    }
    return joint_dgrad2_launch(m_hi, m_lo, m_g, split, ec, ef, wb, gb, gl, N, C, H, V, gpc, gpf,
                               stream);
  }
  if (split) { set_error("split grad_lexical needs the fused dgrad kernel"); return LT_ERR_UNSUPPORTED; }
  const int NH = H > 256 ? 256 : H;
  CUtensorMap map_hi, map_lo;
  cuuint64_t dims[2] = {(cuuint64_t)V, (cuuint64_t)H};
  cuuint64_t strides[1] = {(cuuint64_t)V * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)NH};
  cuuint32_t estr[2] = {1, 1};
  for (int i = 0; i < 2; ++i) {
    CUresult r = encode(i == 0 ? &map_hi : &map_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        i == 0 ? whi : wlo, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (W_vocab^T) failed with %d", (int)r);
      return LT_ERR_CUDA;
    }
  }
  JointDgradParams p = {};
  p.pc = pc; p.pf = pf; p.w_blank = wb; p.gl = gl; p.gb = gb;
  p.M = (long long)N * C; p.C = C; p.H = H; p.V = V; p.NH = NH; p.gp = gp;
  const size_t smem = (size_t)kJStages * (2 * 128 * 128 + 2 * 256 * 128) +
                      sizeof(float) * (H + 4 * 32 * 33) + 16 * 8 + 16 + 1024;
  int dev = 0, sms = 0;
  LT_CUDA(cudaGetDevice(&dev));
  LT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long units = ((p.M + 127) / 128) * (H / NH);
  const int grid = (int)(units < sms ? units : sms);
  LT_CUDA(cudaFuncSetAttribute(joint_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  joint_dgrad_tc_kernel<<<grid, kJThreads, smem, stream>>>(map_hi, map_lo, p);
  LT_LAUNCHED();
  // reduction: pick the widest column block whose [C, jw] accumulator fits in shared memory
  if (H % 128 == 0 && ((size_t)C * 32 + 16 * 32) * sizeof(float4) <= 200 * 1024 &&
      reinterpret_cast<uintptr_t>(gpf) % 16 == 0) {
    const int jblocks = H / 128;
    long long nblocks = (2 * sms + jblocks - 1) / jblocks;
    if (nblocks > N) nblocks = N;
    if (nblocks < 1) nblocks = 1;
    const long long fpb = (N + nblocks - 1) / nblocks;
    nblocks = (N + fpb - 1) / fpb;
    const size_t rsmem = ((size_t)C * 32 + 16 * 32) * sizeof(float4);
    LT_CUDA(cudaFuncSetAttribute(joint_reduce128_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
    joint_reduce128_kernel<<<dim3(jblocks, (unsigned)nblocks), 512, rsmem, stream>>>(
        gp, pc, pf, (long long)N, C, H, fpb, gpc, gpf);
    LT_LAUNCHED();
    return LT_OK;
  }
  int jw = 128;
  while (jw > 8 && ((size_t)C * jw * 4 + 512 * 4 > 200 * 1024 || H % jw != 0)) jw >>= 1;
  if (H % jw != 0 || (size_t)C * jw * 4 + 512 * 4 > 200 * 1024) {
    set_error("joint backward: %d context states do not fit the reduction kernel", C);
    return LT_ERR_UNSUPPORTED;
  }
  const int jblocks = H / jw;
  long long nblocks = (2 * sms + jblocks - 1) / jblocks;
  if (nblocks > N) nblocks = N;
  if (nblocks < 1) nblocks = 1;
  const long long fpb = (N + nblocks - 1) / nblocks;
  nblocks = (N + fpb - 1) / fpb;
  const size_t rsmem = ((size_t)C * jw + 512) * sizeof(float);
  LT_CUDA(cudaFuncSetAttribute(joint_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)rsmem));
  joint_reduce_kernel<<<dim3(jblocks, (unsigned)nblocks), 512, rsmem, stream>>>(
      gp, pc, pf, (long long)N, C, H, jw, fpb, gpc, gpf);
  LT_LAUNCHED();
  return LT_OK;
}


bool joint_wgrad_tc_supported(int64_t N, int C, int H, int V, const void* gl, const void* pc,
                              const void* pf) {
  if (option(OPT_JOINT_SIMT) || option(OPT_JOINT_WGRAD_SIMT)) return false;
  if (V != 128 && V != 256) return false;
  if (H % 128 != 0 || H > 4096 || (H > 256 && H % 256 != 0)) return false;
  if (N < 1 || C < 32) return false;      // a stage of kWK rows touches at most two frames
  auto al = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 32 == 0; };   // 256-bit loads
  return al(gl) && al(pc) && al(pf);
}

int joint_wgrad_tc_launch(const float* pc, const float* pf, const float* gb, const float* gl,
                          int split, int64_t N, int C, int H, int V, float* gwb, float* gbb,
                          float* gwv, float* gbv, cudaStream_t stream) {
  JointWgradParams p = {};
  p.pc = pc; p.pf = pf; p.gl = gl; p.gb = gb;
  p.M = (long long)N * C; p.C = C; p.H = H; p.V = V; p.NJ = H > 256 ? 256 : H;
  p.gwv = gwv; p.gwb = gwb; p.gbv = gbv; p.gbb = gbb;
  int dev = 0, sms = 0;
  LT_CUDA(cudaGetDevice(&dev));
  LT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int nj = H / p.NJ;
  long long ranges = sms / nj;
  if (ranges < 1) ranges = 1;
  long long rows = (p.M + ranges - 1) / ranges;
  rows = (rows + kWK - 1) / kWK * kWK;
  ranges = (p.M + rows - 1) / rows;
  p.rows_per_cta = rows;
  const size_t smem = (size_t)kWStages * 4 * 256 * kWK * 2 + 16 * 8 + 16 + 1024;
  const unsigned grid = (unsigned)(ranges * nj);
#define LT_WGRAD1(AI, BI, SP)                                                                  \
  do {                                                                                         \
    LT_CUDA(cudaFuncSetAttribute(joint_wgrad_tc_kernel<AI, BI, SP>,                            \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    joint_wgrad_tc_kernel<AI, BI, SP><<<grid, kWThreads, smem, stream>>>(p);                   \
  } while (0)
#define LT_WGRAD(AI, BI)                                                                       \
  do {                                                                                         \
    if (split) LT_WGRAD1(AI, BI, true); else LT_WGRAD1(AI, BI, false);                         \
  } while (0)
  const int ai = V / 128, bi = p.NJ / 128;
  if (ai == 2 && bi == 2) LT_WGRAD(2, 2);
  else if (ai == 2 && bi == 1) LT_WGRAD(2, 1);
  else if (ai == 1 && bi == 2) LT_WGRAD(1, 2);
  else LT_WGRAD(1, 1);
#undef LT_WGRAD
#undef LT_WGRAD1
  LT_LAUNCHED();
  return LT_OK;
}

}  // namespace lt

extern "C" int ltx_umma_probe_mn(const float* At, const float* Bt, float* D, int N, int K, int swap,
                                 void* stream) {
  using namespace lt;
  LT_CHECK_ARG(N % 64 == 0 && N >= 64 && N <= 256 && K % 32 == 0 && K > 0,
               "ltx_umma_probe_mn: need N %% 64 == 0, N <= 256, K %% 32 == 0 (N=%d K=%d)", N, K);
  const size_t smem = 2 * 128 * 64 + 2 * 256 * 64 + 1024;
  LT_CUDA(cudaFuncSetAttribute(umma_probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  umma_probe_mn_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(At, Bt, D, N, K, swap);
  LT_LAUNCHED();
  return LT_OK;
}

// Diagnostic entry point (not part of include/last_lattice.h): D = A * B^T with
// A [128,K], B [N,K] fp32 device pointers, N % 16 == 0, N <= 256, K % 64 == 0.
extern "C" int ltx_umma_probe(const float* A, const float* B, float* D, int N, int K, int terms,
                              void* stream) {
  using namespace lt;
  LT_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && K % 64 == 0 && K > 0,
               "ltx_umma_probe: need N %% 16 == 0, N <= 256, K %% 64 == 0 (N=%d K=%d)", N, K);
  const size_t smem = 2 * 128 * 128 + 2 * 256 * 128 + 1024;
  LT_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, terms);
  LT_LAUNCHED();
  return LT_OK;
}
