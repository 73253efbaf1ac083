// JointWeightFn forward (weight_fns.py:208-227, whole-utterance form) with the generated operand
// in TENSOR MEMORY.
//
//   lexical[m, :] = tanh(pc[c] + pf[n]) . W_vocab^T + b_vocab      m = n*C + c
//   blank[m]      = tanh(pc[c] + pf[n]) . w_blank   + b_blank
//
// joint_forward_tc_kernel (joint_tc.cu) writes the tanh tile into shared memory and the tensor
// core reads it back: its L1 data pipe carries the producers' shared stores, their global loads
// AND the tensor core's reads of both operands, and saturates (ncu, profiles/r02_ncu_summary.csv:
// 46.6 % tensor-core operand wavefronts + 59 % LSU wavefronts of the cycles).  Here the A operand
// never exists in shared memory: a producer thread owns ONE joint row (= one TMEM lane), computes
// tanh for a run of hidden units in registers and writes the bf16 hi / lo pairs straight into TMEM
// (tcgen05.st); tcgen05.mma takes A from TMEM and only W_vocab (TMA, SWIZZLE_128B) from shared
// memory.  Per 16-deep K step the tensor core reads 24 KB instead of 36 KB of shared memory and
// the producers issue no shared stores at all.
//
// Tensor memory (512 columns): three accumulators of 128 columns, each one HALF of a tile's
// vocabulary columns ([128 rows x V/2]), used in rotation -- tile i uses (2i) % 3 and (2i+1) % 3,
// so one half of the next tile always has a free accumulator and the other waits only for the
// first half of the epilogue -- plus two A stages of 64 columns (64 hidden units: 32 columns hi,
// 32 columns lo).
//
// The e^(2 proj_ctx) table is stored as [H/4][C][4]: the 32 lanes of a producer warp (32
// consecutive context states, one hidden-unit quad) read 512 contiguous bytes per instruction.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "params.cuh"
#include "umma.cuh"

namespace lt {
namespace {

__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LTS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1" LT_MBAR_HINT ";\n"
      "@p bra LTS_DONE;\n"
      "bra LTS_WAIT;\n"
      "LTS_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                               uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      ".multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst), "l"(map), "r"(c0), "r"(c1),
      "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ldg_nc4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Probe (tests/test_gpu_umma.py): D[128, N] = A[128, K] * B[N, K]^T with A written to TMEM by
// tcgen05.st, bf16x3 split.  Pins the TMEM operand layout on hardware.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_ts_kernel(const float* __restrict__ A, const float* __restrict__ B,
                     float* __restrict__ D, int N, int K) {
  extern __shared__ __align__(1024) unsigned char tsmem_raw[];
  unsigned char* sm = tsmem_raw + ((1024u - (smem_u32(tsmem_raw) & 1023u)) & 1023u);
  unsigned char* b_hi = sm;                        // N x 128 B
  unsigned char* b_lo = b_hi + 256 * 128;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    bar_init(smem_u32(&mbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) umma::tmem_alloc(smem_u32(&tmem_base), 512);
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = umma::make_idesc_bf16(128, N);
  const uint32_t a_col = 256;
  for (int kc = 0; kc < K / 64; ++kc) {
    // A: thread = row = TMEM lane; 64 K elements -> 32 columns hi + 32 columns lo
    const float* arow = A + (size_t)tid * K + kc * 64;
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        umma::split_pack2(arow[part * 32 + 2 * j], arow[part * 32 + 2 * j + 1], hi[j], lo[j]);
      const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16) + a_col + part * 16;
      umma::tmem_st16(t, hi);
      umma::tmem_st16(t + 32, lo);
    }
    umma::tmem_st_wait();
    // B: swizzled shared tiles
    for (int idx = tid; idx < N * 8; idx += blockDim.x) {
      const int row = idx >> 3, chunk = idx & 7;
      const float* src = B + (size_t)row * K + kc * 64 + chunk * 8;
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = src[i];
      uint4 h, l;
      umma::split_pack8(x, h, l);
      const uint32_t off = umma::swizzled_offset(row, chunk);
      *reinterpret_cast<uint4*>(b_hi + off) = h;
      *reinterpret_cast<uint4*>(b_lo + off) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    umma::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_thread_sync();
      for (int k = 0; k < 4; ++k) {
        const uint32_t ah = tmem + a_col + k * 8, al = ah + 32;
        const uint64_t dbh = umma::make_smem_desc_sw128(smem_u32(b_hi) + k * 32);
        const uint64_t dbl = umma::make_smem_desc_sw128(smem_u32(b_lo) + k * 32);
        umma::mma_bf16_ts(tmem, ah, dbh, idesc, (kc | k) > 0);
        umma::mma_bf16_ts(tmem, ah, dbl, idesc, 1);
        umma::mma_bf16_ts(tmem, al, dbh, idesc, 1);
      }
      umma::commit(smem_u32(&mbar));
    }
    bar_wait(smem_u32(&mbar), kc & 1);
    umma::fence_after_thread_sync();
  }
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// Forward kernel.  Warp roles:
//   warp 0        TMA producer: W_hi / W_lo chunk [V x 64] bf16 (3 stages of shared memory); in a
//                 cluster of CS CTAs every CTA loads V / CS rows and MULTICASTS them to all
//   warp 1        MMA issuer (one thread): per 16-deep K step and per half, Ah*Bh + Ah*Bl + Al*Bh
//   warps 2-5     epilogue: TMEM -> registers -> + bias -> swizzled staging tile -> TMA store
//   warps 6..     A producers: PW warps per TMEM lane quadrant; thread = (row, 64 / PW hidden
//                 units of the chunk); the same threads accumulate the blank mat-vec
//
// What bounds the shared-memory-operand kernel is the L2 -> SM fabric (~6300 B / clock for the
// whole chip): per 128-row tile it pulls W_vocab hi + lo (512 KB at V = 256, H = 512), one e^(2 pc)
// row per joint row (256 KB) and writes 128 KB of logits.  Two changes cut that to 456 KB:
//   * tile = 32 context states x 4 frames (quadrant q of the tile = frame 4 fg + q): the four
//     quadrant warps read the SAME 32 e^(2 pc) rows (one L2 fetch, three L1 hits), 64 KB per tile;
//     the C % 32 left-over states are covered by tiles of one state x 128 frames;
//   * W_vocab is multicast inside a cluster of two CTAs: 256 KB per tile.
// ---------------------------------------------------------------------------------------------
constexpr int kBStages = 3;
constexpr int kAStages = 2;
constexpr uint32_t kACol = 384;
constexpr uint32_t kBStageBytes = 2 * 256 * 128;

struct JointTsParams {
  const float* ec4;      // [H/4][C][4]  e^(2 proj_ctx)
  const float* ef;       // [N, H]       e^(2 proj_frame)
  const float* w_blank;  // [H]
  const float* b_vocab;  // [V]
  const float* b_blank;  // device scalar
  long long N;           // frames
  int C, H, V;
  int Cb;                // full 32-state blocks: C / 32
  long long FB;          // 128-frame blocks: ceil(N / 128)
  long long n_main;      // ceil(N / 4) * Cb tiles of 32 states x 4 frames ...
  long long num_tiles;   // ... then (C % 32) * FB tiles of one state x 128 frames
  long long rounds;      // tiles per CTA (every CTA of a cluster runs the same number)
  float* blank;          // [N, C]
  float* lexical;        // [N, C, V]
};

// e^(2x) table of proj_ctx in the [H/4][C][4] layout (see joint_exp_table_kernel, joint_tc.cu)
__global__ void joint_exp_table4_kernel(const float* __restrict__ x, float* __restrict__ out, int C,
                                        int H) {
  const long long n = (long long)C * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / H), h = (int)(i % H);
    double a = (double)x[i] * 2.8853900817779268;
    a = a < -126.0 ? -126.0 : (a > 126.0 ? 126.0 : a);
    out[((size_t)(h >> 2) * C + c) * 4 + (h & 3)] = (float)exp2(a);
  }
}

// joint row of (tile, TMEM lane quadrant, lane)
struct TsRow {
  long long n;
  int c;
  bool valid;
};
__device__ __forceinline__ TsRow ts_row(const JointTsParams& p, long long tile, int quad, int lane) {
  TsRow r;
  if (tile < p.n_main) {
    const long long fg = tile / p.Cb;
    r.n = fg * 4 + quad;
    r.c = (int)(tile - fg * p.Cb) * 32 + lane;
  } else {
    const long long t = tile - p.n_main;
    const long long cr = t / p.FB;
    r.n = (t - cr * p.FB) * 128 + quad * 32 + lane;
    r.c = p.Cb * 32 + (int)cr;
  }
  r.valid = tile < p.num_tiles && r.n < p.N;
  if (!r.valid) { r.n = 0; r.c = 0; }
  return r;
}

template <int PW, int CS>
__global__ void __launch_bounds__((6 + 4 * PW) * 32, 1)
joint_forward_ts_kernel(const __grid_constant__ CUtensorMap map_hi,
                        const __grid_constant__ CUtensorMap map_lo,
                        const __grid_constant__ CUtensorMap map_main,
                        const __grid_constant__ CUtensorMap map_rem, const JointTsParams p) {
  constexpr int kThreads = (6 + 4 * PW) * 32;
  constexpr int KPT = 64 / PW;            // hidden units per thread per chunk
  constexpr int QPT = KPT / 4;            // quads
  extern __shared__ __align__(1024) unsigned char fsmem_raw[];
  unsigned char* base = fsmem_raw + ((1024u - (smem_u32(fsmem_raw) & 1023u)) & 1023u);
  const int V = p.V, H = p.H, Vh = V / 2;
  const uint32_t b_bytes = (uint32_t)V * 128;
  unsigned char* s_out = base + kBStages * kBStageBytes;                   // 4 x 4096 B
  float* s_wb = reinterpret_cast<float*>(s_out + 4 * 4096);                // [H]
  float* s_bias = s_wb + H;                                                // [256]
  float* s_part = s_bias + 256;                                            // [2][PW-1][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 2 * (PW - 1) * 128);
  uint64_t* full_b = bars;                       // [kBStages]  TMA -> MMA
  uint64_t* empty_b = full_b + kBStages;         // [kBStages]  MMA (commit, all CTAs) -> TMA
  uint64_t* full_a = empty_b + kBStages;         // [kAStages]  producers -> MMA
  uint64_t* empty_a = full_a + kAStages;         // [kAStages]  MMA (commit) -> producers
  uint64_t* tfull = empty_a + kAStages;          // [3]         MMA (commit) -> epilogue
  uint64_t* tempty = tfull + 3;                  // [3]         epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = H / 64;
  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
  const long long cluster_id = blockIdx.x / CS, num_clusters = gridDim.x / CS;
  // tile of round j: ((j * num_clusters + cluster_id) * CS + rank)
  const long long tile0 = cluster_id * CS + rank, tile_step = num_clusters * CS;

  for (int i = tid; i < H; i += kThreads) s_wb[i] = p.w_blank[i];
  for (int i = tid; i < V; i += kThreads) s_bias[i] = p.b_vocab[i];
  if (tid == 0) {
    for (int s = 0; s < kBStages; ++s) {
      bar_init(smem_u32(&full_b[s]), 1);
      bar_init(smem_u32(&empty_b[s]), CS);
    }
    for (int s = 0; s < kAStages; ++s) {
      bar_init(smem_u32(&full_a[s]), 4 * PW);       // one arrival per producer warp
      bar_init(smem_u32(&empty_a[s]), 1);
    }
    for (int a = 0; a < 3; ++a) {
      bar_init(smem_u32(&tfull[a]), 1);
      bar_init(smem_u32(&tempty[a]), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_main) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_rem) : "memory");
  }
  if (warp == 1) umma::tmem_alloc(smem_u32(tmem_slot), 512);
  umma::fence_before_thread_sync();
  __syncthreads();
  if (CS > 1) cluster_sync_all();               // peers' barriers exist before anything remote
  umma::fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;
  const uint16_t all_mask = (uint16_t)((1u << CS) - 1);

  if (warp == 0) {
    // ---------------------------------------------------------- TMA producer (B)
    if (lane == 0) {
      const uint32_t share = b_bytes / CS;          // bytes of one of hi / lo this CTA loads
      const int row0 = (int)rank * (V / CS);
      uint32_t g = 0;
      for (long long j = 0; j < p.rounds; ++j) {
        for (int kc = 0; kc < nchunks; ++kc, ++g) {
          const uint32_t s = g % kBStages;
          bar_wait(smem_u32(&empty_b[s]), ((g / kBStages) & 1) ^ 1);
          const uint32_t bar = smem_u32(&full_b[s]);
          const uint32_t dst = smem_u32(base) + s * kBStageBytes + rank * share;
          bar_expect_tx(bar, 2 * b_bytes);
          if (CS > 1) {
            tma_load_2d_mc(dst, &map_hi, kc * 64, row0, bar, all_mask);
            tma_load_2d_mc(dst + 256 * 128, &map_lo, kc * 64, row0, bar, all_mask);
          } else {
            tma_load_2d(dst, &map_hi, kc * 64, 0, bar);
            tma_load_2d(dst + 256 * 128, &map_lo, kc * 64, 0, bar);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_bf16(128, Vh);
      uint32_t g = 0;
      for (uint32_t it = 0; it < (uint32_t)p.rounds; ++it) {
        const uint32_t u0 = 2 * it, u1 = u0 + 1;
        const uint32_t a0 = u0 % 3, a1 = u1 % 3;
        // a0 has been free since the tile before last; a1 is the accumulator the epilogue is
        // draining right now (the previous tile's first half): wait for it only after this
        // tile's first chunk has been issued for a0
        bar_wait(smem_u32(&tempty[a0]), ((u0 / 3) & 1) ^ 1);
        umma::fence_after_thread_sync();
        const uint32_t d0 = tmem + a0 * 128, d1 = tmem + a1 * 128;
        for (int kc = 0; kc < nchunks; ++kc, ++g) {
          const uint32_t sb = g % kBStages, sa = g % kAStages;
          bar_wait(smem_u32(&full_b[sb]), (g / kBStages) & 1);
          bar_wait(smem_u32(&full_a[sa]), (g / kAStages) & 1);
          umma::fence_after_thread_sync();
          const uint32_t bh = smem_u32(base) + sb * kBStageBytes;
          const uint32_t bl = bh + 256 * 128;
          const uint32_t half = (uint32_t)Vh * 128;
          if (kc == 0) {
            // first chunk: all of a0, then (once the epilogue has let go of it) all of a1
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t ah = tmem + kACol + sa * 64 + k * 8, al = ah + 32;
              const uint64_t dbh0 = umma::make_smem_desc_sw128(bh + k * 32);
              const uint64_t dbl0 = umma::make_smem_desc_sw128(bl + k * 32);
              umma::mma_bf16_ts(d0, ah, dbh0, idesc, k > 0);
              umma::mma_bf16_ts(d0, ah, dbl0, idesc, 1);
              umma::mma_bf16_ts(d0, al, dbh0, idesc, 1);
            }
            bar_wait(smem_u32(&tempty[a1]), ((u1 / 3) & 1) ^ 1);
            umma::fence_after_thread_sync();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t ah = tmem + kACol + sa * 64 + k * 8, al = ah + 32;
              const uint64_t dbh1 = umma::make_smem_desc_sw128(bh + half + k * 32);
              const uint64_t dbl1 = umma::make_smem_desc_sw128(bl + half + k * 32);
              umma::mma_bf16_ts(d1, ah, dbh1, idesc, k > 0);
              umma::mma_bf16_ts(d1, ah, dbl1, idesc, 1);
              umma::mma_bf16_ts(d1, al, dbh1, idesc, 1);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t ah = tmem + kACol + sa * 64 + k * 8, al = ah + 32;
              const uint64_t dbh0 = umma::make_smem_desc_sw128(bh + k * 32);
              const uint64_t dbl0 = umma::make_smem_desc_sw128(bl + k * 32);
              const uint64_t dbh1 = umma::make_smem_desc_sw128(bh + half + k * 32);
              const uint64_t dbl1 = umma::make_smem_desc_sw128(bl + half + k * 32);
              umma::mma_bf16_ts(d0, ah, dbh0, idesc, 1);
              umma::mma_bf16_ts(d0, ah, dbl0, idesc, 1);
              umma::mma_bf16_ts(d0, al, dbh0, idesc, 1);
              umma::mma_bf16_ts(d1, ah, dbh1, idesc, 1);
              umma::mma_bf16_ts(d1, ah, dbl1, idesc, 1);
              umma::mma_bf16_ts(d1, al, dbh1, idesc, 1);
            }
          }
          if (CS > 1) umma::commit_mc(smem_u32(&empty_b[sb]), all_mask);
          else umma::commit(smem_u32(&empty_b[sb]));
          umma::commit(smem_u32(&empty_a[sa]));
        }
        umma::commit(smem_u32(&tfull[a0]));
        umma::commit(smem_u32(&tfull[a1]));
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;
    unsigned char* stage_tile = s_out + quad * 4096;
    for (uint32_t it = 0; it < (uint32_t)p.rounds; ++it) {
      const long long tile = tile0 + it * tile_step;
      // output block of this quadrant: 32 consecutive states of one frame (main tiles) or one
      // state of 32 consecutive frames (left-over tiles); frames past N are clipped by the TMA
      const bool main_tile = tile < p.n_main;
      int oc, on;
      bool store;
      if (main_tile) {
        const long long fg = tile / p.Cb;
        oc = (int)(tile - fg * p.Cb) * 32;
        on = (int)(fg * 4 + quad);
        store = on < p.N;
      } else {
        const long long t = tile - p.n_main;
        const long long cr = t / p.FB;
        oc = p.Cb * 32 + (int)cr;
        on = (int)((t - cr * p.FB) * 128 + quad * 32);
        store = tile < p.num_tiles && on < p.N;
      }
      for (uint32_t h = 0; h < 2; ++h) {
        const uint32_t u = 2 * it + h, a = u % 3;
        bar_wait(smem_u32(&tfull[a]), (u / 3) & 1);
        umma::fence_after_thread_sync();
        for (int c0 = 0; c0 < Vh; c0 += 32) {
          float v[32];
          umma::tmem_ld32(tmem + a * 128 + ((uint32_t)(quad * 32) << 16) + c0, v);
          const int col = h * Vh + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + col + j);
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(stage_tile + umma::swizzled_offset(lane, k)) =
                make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && store) {
            const CUtensorMap* map = main_tile ? &map_main : &map_rem;
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::
                    "l"(map), "r"(col), "r"(oc), "r"(on), "r"(smem_u32(stage_tile))
                : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        umma::fence_before_thread_sync();
        bar_arrive(smem_u32(&tempty[a]));
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // -------------------------------------------------------------- A producers
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may touch
    const int sub = (warp - 6) >> 2;              // which KPT-wide run of the chunk
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16) + kACol + sub * (KPT / 2);
    const size_t quad_stride = (size_t)p.C * 4;   // floats between consecutive hidden-unit quads
    const float bb = __ldg(p.b_blank);
    const float* wb = s_wb + sub * KPT;
    auto ec_of = [&](const TsRow& r) {
      return p.ec4 + (size_t)r.c * 4 + (size_t)(sub * QPT) * quad_stride;
    };
    auto ef_of = [&](const TsRow& r) { return p.ef + (size_t)r.n * H + sub * KPT; };
    TsRow cur = ts_row(p, tile0, quad, lane);
    const float* ec = ec_of(cur);
    const float* ef = ef_of(cur);
    // the operands of a chunk are loaded while the previous one is being computed: e / f always
    // hold the NEXT chunk's table entries (no extra registers, the L2 latency is off the chain)
    float4 e[QPT], f[QPT];
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
      e[i] = ldg_nc4(ec + (size_t)i * quad_stride);
      f[i] = ldg_nc4(ef + i * 4);
    }
    uint32_t g = 0;
    for (uint32_t it = 0; it < (uint32_t)p.rounds; ++it) {
      const bool has_next = it + 1 < (uint32_t)p.rounds;
      const TsRow nxt = ts_row(p, tile0 + (it + 1) * tile_step, quad, lane);
      const float* ec_n = ec_of(nxt);
      const float* ef_n = ef_of(nxt);
      float bacc = 0.f;
      for (int kc = 0; kc < nchunks; ++kc, ++g) {
        const uint32_t sa = g % kAStages;
        const bool last = kc + 1 == nchunks;
        const float* ec_l = last ? ec_n : ec + (size_t)(kc + 1) * 16 * quad_stride;
        const float* ef_l = last ? ef_n : ef + (kc + 1) * 64;
        const bool reload = !last || has_next;
        uint32_t hi[KPT / 2], lo[KPT / 2];
#pragma unroll
        for (int i = 0; i < QPT; ++i) {
          const float4 w = *reinterpret_cast<const float4*>(wb + kc * 64 + i * 4);
          float t0 = tanh_from_exp(e[i].x, f[i].x), t1 = tanh_from_exp(e[i].y, f[i].y);
          float t2 = tanh_from_exp(e[i].z, f[i].z), t3 = tanh_from_exp(e[i].w, f[i].w);
          if (reload) {
            e[i] = ldg_nc4(ec_l + (size_t)i * quad_stride);
            f[i] = ldg_nc4(ef_l + i * 4);
          }
          if (!cur.valid) t0 = t1 = t2 = t3 = 0.f;
          bacc = fmaf(t0, w.x, bacc);
          bacc = fmaf(t1, w.y, bacc);
          bacc = fmaf(t2, w.z, bacc);
          bacc = fmaf(t3, w.w, bacc);
          umma::split_pack2(t0, t1, hi[2 * i], lo[2 * i]);
          umma::split_pack2(t2, t3, hi[2 * i + 1], lo[2 * i + 1]);
        }
        bar_wait(smem_u32(&empty_a[sa]), ((g / kAStages) & 1) ^ 1);
        umma::fence_after_thread_sync();
        if constexpr (KPT == 32) {
          umma::tmem_st16(t_lane + sa * 64, hi);
          umma::tmem_st16(t_lane + sa * 64 + 32, lo);
        } else {
          umma::tmem_st8(t_lane + sa * 64, hi);
          umma::tmem_st8(t_lane + sa * 64 + 32, lo);
        }
        umma::tmem_st_wait();
        umma::fence_before_thread_sync();
        __syncwarp();
        if (lane == 0) bar_arrive(smem_u32(&full_a[sa]));
      }
      // blank: the PW partial sums of a row meet in shared memory (double-buffered by tile parity)
      float* part = s_part + (it & 1) * (PW - 1) * 128;
      if (sub > 0) part[(sub - 1) * 128 + row] = bacc;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "r"(32 * PW) : "memory");
      if (sub == 0 && cur.valid) {
#pragma unroll
        for (int s = 0; s < PW - 1; ++s) bacc += part[s * 128 + row];
        p.blank[cur.n * p.C + cur.c] = bacc + bb;
      }
      cur = nxt;
      ec = ec_n;
      ef = ef_n;
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (CS > 1) cluster_sync_all();               // no CTA leaves while a peer may still write to it
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);
EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
          cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeFn>(sym);
  return fn;
}

template <int CS>
int launch_ts(const CUtensorMap& map_hi, const CUtensorMap& map_lo, const CUtensorMap& map_main,
              const CUtensorMap& map_rem, JointTsParams p, int sms, cudaStream_t stream) {
  constexpr int PW = 4;
  constexpr int kThreads = (6 + 4 * PW) * 32;
  const size_t smem = (size_t)kBStages * kBStageBytes + 4 * 4096 +
                      sizeof(float) * (p.H + 256 + 2 * (PW - 1) * 128) + 16 * 8 + 16 + 1024;
  auto kernel = joint_forward_ts_kernel<PW, CS>;
  LT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  long long clusters = sms / CS;
  if (CS > 1) {
    cfg.gridDim = dim3((unsigned)(clusters * CS));
    int max_clusters = 0;
    LT_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
    if (max_clusters < 1) { set_error("joint forward: no cluster of %d CTAs fits", CS); return LT_ERR_CUDA; }
    clusters = std::min<long long>(clusters, max_clusters);
  }
  clusters = std::min<long long>(clusters, (p.num_tiles + CS - 1) / CS);
  p.rounds = (p.num_tiles + clusters * CS - 1) / (clusters * CS);
  cfg.gridDim = dim3((unsigned)(clusters * CS));
  LT_CUDA(cudaLaunchKernelEx(&cfg, kernel, map_hi, map_lo, map_main, map_rem, p));
  LT_LAUNCHED();
  return LT_OK;
}

}  // namespace

bool joint_forward_ts_supported(int64_t N, int C, int H, int V, const void* lexical) {
  if (option(OPT_JOINT_SIMT) || option(OPT_JOINT_FWD_SS)) return false;
  if (V % 64 != 0 || V < 64 || V > 256) return false;
  if (H % 64 != 0 || H > 1024) return false;
  if (N < 1 || N > 0x7fffffff || C < 1) return false;
  return reinterpret_cast<uintptr_t>(lexical) % 16 == 0;
}

// workspace: [W_vocab bf16 hi | lo] [e^(2 pc) as [H/4][C][4] | e^(2 pf) [N,H]]  (same size as the
// shared-memory-operand kernel's)
int joint_forward_ts_launch(const float* pc, const float* pf, const float* wb, const float* bb,
                            const float* wv, const float* bv, int64_t N, int C, int H, int V,
                            float* blank, float* lexical, void* workspace, cudaStream_t stream) {
  EncodeFn encode = encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled is unavailable in this driver"); return LT_ERR_CUDA; }
  const int cs = option(OPT_JOINT_FWD_CLUSTER) == 1 ? 1 : 2;
  __nv_bfloat16* whi = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* wlo = whi + (size_t)V * H;
  if (int rc = joint_split_weights_launch(wv, whi, wlo, V * H, stream)) return rc;
  CUtensorMap map_hi, map_lo, map_main, map_rem;
  {
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)V};
    cuuint64_t strides[1] = {(cuuint64_t)H * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)(V / cs)};      // every CTA of a cluster loads V / cs rows
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
  }
  {
    // lexical [N, C, V]: one epilogue block is 32 states of one frame or one state of 32 frames
    cuuint64_t odims[3] = {(cuuint64_t)V, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t ostrides[2] = {(cuuint64_t)V * 4, (cuuint64_t)C * V * 4};
    cuuint32_t oestr[3] = {1, 1, 1};
    for (int i = 0; i < 2; ++i) {
      cuuint32_t obox[3] = {32, i == 0 ? 32u : 1u, i == 0 ? 1u : 32u};
      CUresult r = encode(i == 0 ? &map_main : &map_rem, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                          lexical, odims, ostrides, obox, oestr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (lexical output) failed with %d", (int)r);
        return LT_ERR_CUDA;
      }
    }
  }
  float* ec4 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) +
                                        joint_split_bytes(H, V));
  float* ef = ec4 + (size_t)C * H;
  const long long nc = (long long)C * H;
  joint_exp_table4_kernel<<<(unsigned)std::min<long long>((nc + 255) / 256, 4096), 256, 0, stream>>>(
      pc, ec4, C, H);
  LT_LAUNCHED();
  if (int rc = joint_exp_table_launch(pf, ef, (long long)N * H, stream)) return rc;
  JointTsParams p = {};
  p.ec4 = ec4; p.ef = ef; p.w_blank = wb; p.b_vocab = bv; p.b_blank = bb;
  p.N = N; p.C = C; p.H = H; p.V = V; p.blank = blank; p.lexical = lexical;
  p.Cb = C / 32;
  p.FB = (N + 127) / 128;
  p.n_main = ((N + 3) / 4) * p.Cb;
  p.num_tiles = p.n_main + (long long)(C % 32) * p.FB;
  int dev = 0, sms = 0;
  LT_CUDA(cudaGetDevice(&dev));
  LT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return cs == 2 ? launch_ts<2>(map_hi, map_lo, map_main, map_rem, p, sms, stream)
                 : launch_ts<1>(map_hi, map_lo, map_main, map_rem, p, sms, stream);
}

}  // namespace lt

// Diagnostic entry point (not part of include/last_lattice.h): D = A * B^T with A [128,K] written
// to tensor memory, B [N,K]; N % 16 == 0, N <= 256, K % 64 == 0.
extern "C" int ltx_umma_probe_ts(const float* A, const float* B, float* D, int N, int K,
                                 void* stream) {
  using namespace lt;
  LT_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && K % 64 == 0 && K > 0,
               "ltx_umma_probe_ts: need N %% 16 == 0, N <= 256, K %% 64 == 0 (N=%d K=%d)", N, K);
  const size_t smem = 2 * 256 * 128 + 1024;
  LT_CUDA(cudaFuncSetAttribute(umma_probe_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  umma_probe_ts_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K);
  LT_LAUNCHED();
  return LT_OK;
}
