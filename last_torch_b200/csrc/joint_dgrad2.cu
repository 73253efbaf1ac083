// JointWeightFn backward, data gradient, second generation: ONE tcgen05 kernel that goes
// from grad_lexical / grad_blank to grad_proj_ctx and grad_proj_frame without writing the
// [M, H] pre-activation gradient to memory (the first generation wrote 16.8 GB of it and
// streamed it back through a reduction kernel).
//
//   Gp[m, j] = (sum_v G[m, v] W_vocab[v, j] + gb[m] w_blank[j]) * (1 - tanh^2(pc[c, j] + pf[n, j]))
//            = (...) * 4 r (1 - r),  r = 1 / (1 + E_c[c, j] E_f[n, j])   (exponential tables)
//   grad_proj_ctx[c, j]   = sum_n Gp[(n, c), j]        grad_proj_frame[n, j] = sum_c Gp[(n, c), j]
//
// The product is computed TRANSPOSED, D[j, i] = sum_v W^T[j, v] G[m_i, v], so that a TMEM lane
// (= an epilogue thread) is a hidden unit j and the TMEM columns are joint rows:
//   * a CTA owns one block of 128 hidden units for its whole life; the A operand, W_vocab^T
//     [128 x V] as bf16 hi / lo (bf16x3 split, see joint_tc.cu), is loaded ONCE by TMA and stays
//     resident in shared memory (128 KB at V = 256);
//   * a work item is (block of 128 frames n0.., this hidden block); it is swept as C tiles, tile
//     c = the 128 joint rows {(n0 + i, c)}: the B operand, their grad_lexical rows split hi / lo
//     on the fly by 16 producer warps into a 3-stage K-major SWIZZLE_128B ring;
//   * with that tiling the sum over frames (grad_proj_ctx[c, j]) is a serial sum over the
//     columns of one tile in the epilogue thread, and the sum over context states
//     (grad_proj_frame[n0 + i, j]) accumulates over the C tiles of the item IN TENSOR MEMORY:
//     128 TMEM columns hold the running sums, 128 more hold pf[n0 + i, j] for the item, so the
//     epilogue never touches shared or global memory per element (tcgen05.ld / tcgen05.st);
//   * two accumulators (2 x 128 TMEM columns): the epilogue of tile c overlaps the MMAs of c+1;
//   * the (frame block, c) tiles of the whole problem form ONE sequence that is cut into equal
//     contiguous ranges, one per CTA of a hidden block (no tail imbalance); a frame block that
//     straddles two ranges is flushed by both owners, so grad_proj_frame is added atomically;
//   * the epilogue reads TMEM in 8-column groups, the loads of group g+1 in flight while group
//     g is being computed.
// Two input forms of grad_lexical (template SPLIT): fp32 rows, split hi / lo on the fly by 16
// producer warps (6.6 ms at B=32, T=1000, C=257, H=512, V=256; the first version took 11.8 ms
// with 266 producer instructions per 16-element chunk), or "split rows" written by the lattice
// backward kernel, loaded by TMA with no producer warps at all (4.9 ms, tensor pipe 65 %).
// What bounds the latter: with N = 128 every SS-mode tcgen05.mma reads 8 KB of shared memory per
// 64 issue cycles -- the whole 128 B/clk port whenever the tensor core is active -- and every
// gradient row is pulled from L2 by the four CTAs (hidden blocks) that need it (6.9 TB/s).
// DESIGN.md section 6 lists what was tried against that (CTA pairs, TMA multicast, more epilogue
// warps, register-resident sums) and did not help.
// TMEM map (512 columns): [0,128) D0 | [128,256) D1 | [256,384) pf | [384,512) grad_proj_frame.
//
// Reference: the autograd of weight_fns.py:208-227 (tanh joint + two Linear layers).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "params.cuh"
#include "umma.cuh"

namespace lt {
namespace {

__device__ __align__(32) float lt_zero_line[64];   // what rows past the last frame read

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
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LTD_WAIT%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1" LT_MBAR_HINT ";\n"
      "@p bra LTD_DONE%=;\n"
      "bra LTD_WAIT%=;\n"
      "LTD_DONE%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                       uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                       uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// ---- TMA multicast: the hidden blocks of one tile range form a cluster; every CTA loads 1/mc
// of a B tile and multicasts it to all of them (the same shared-memory offset and the same
// mbarrier offset in every destination CTA), so a gradient row leaves L2 once, not mc times.
__device__ __forceinline__ void tma_3d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                          int c2, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      ".multicast::cluster [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dst), "l"(map), "r"(c0),
      "r"(c1), "r"(c2), "r"(bar), "h"(mask)
      : "memory");
}
// arrive (count 1) on the barrier at this offset in every CTA of `mask` once the MMAs retire
__device__ __forceinline__ void commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

// ---- CTA pair (cta_group::2) plumbing, as in joint_fwd2.cu: rank 0 of the cluster is the leader
// and the only MMA issuer; "leader" barriers are addressed through the shared::cluster window
// with the peer bit (bit 24) cleared.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void mbar_wait_parity_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LTDC_WAIT%=:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1" LT_MBAR_HINT ";\n"
      "@p bra LTDC_DONE%=;\n"
      "bra LTDC_WAIT%=;\n"
      "LTDC_DONE%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_leader(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::
                   "r"(bar & kPeerBitMask), "r"(bytes) : "memory");
}
// 2-SM TMA loads: data lands in THIS CTA's shared memory, the bytes complete on the leader's barrier
__device__ __forceinline__ void tma_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                           uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1),
      "r"(bar & kPeerBitMask)
      : "memory");
}
__device__ __forceinline__ void tma_3d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                           int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2),
      "r"(bar & kPeerBitMask)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::
                   "r"(smem_result_addr), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// tanh(x) = 1 - 2 / (1 + e^(2x)): two MUFU ops; absolute error ~1e-7
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (2.f * kLog2e)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return fmaf(-2.f, r, 1.f);
}

// 16 consecutive fp32 TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

constexpr int kDProdWarps = 16;
constexpr int kDThreads = (6 + kDProdWarps) * 32;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue,
                                                    // warps 6.. producers
constexpr int kDProducers = kDProdWarps * 32;
constexpr int kDSplitThreads = 6 * 32;    // split-row variant: TMA, MMA, 4 epilogue warps, no producers
                                          // (8 epilogue warps, two per TMEM lane quadrant, measured
                                          // slower: 5.15 vs 4.92 ms)
constexpr int kDPasses = 128 / (kDProdWarps * 4);   // row passes per producer thread and chunk
constexpr int kDStages = 3;
constexpr int kDPairStages = 5;          // CTA-pair variant: 16 KB stages (64 rows per CTA)
constexpr int kTile = 128;              // joint rows per tile = frames per work item
constexpr int kJB = 128;                // hidden units per CTA
// grad_blank ring (tiles): the producers run kDStages chunks ahead of the MMAs, which are at most
// one tile (two accumulators) ahead of the epilogue -- two tiles with V / 64 = 4 chunks per tile
// (3 slots, all that fits beside the 128 KB resident operand), up to four tiles with one chunk
// per tile (8 slots).  A slot's mbarrier must never run two phases ahead of its reader.
constexpr int kGbRingMax = 8;
__host__ __device__ constexpr int gb_ring_slots(int V) { return V >= 256 ? 3 : kGbRingMax; }

struct Dgrad2Params {
  const float* pc;       // [C, H]  e^(2 proj_ctx)   (joint_exp_table_kernel)
  const float* pf;       // [N, H]  e^(2 proj_frame)
  const float* w_blank;  // [H]
  const float* gl;       // [N*C, V]
  const float* gb;       // [N*C]
  long long N;
  int C, H, V;
  float* gpc;            // [C, H]  += (atomics)
  float* gpf;            // [N, H]  += (atomics: a frame block can straddle two CTAs)
  int mc;                // split rows: CTAs per cluster that share every B tile by TMA multicast
};

// SPLIT: grad_lexical arrives as rows of [V bf16 hi | V bf16 lo] (same bytes as fp32; written by
// the lattice backward kernel, see lt_lattice_backward / LT_FLAG_GRAD_SPLIT).  The B operand is
// then a plain TMA load -- one [128 frames x 1 state x 64] box of map_g per half -- and the 16
// producer warps disappear: warp 0 issues the loads and ferries the grad_blank slices.
// PAIR (split rows only): a cluster of two CTAs shares every tile of 128 joint rows.  CTA r
// owns hidden block 2 * jp + r (its own resident A, accumulators, epilogue) and loads HALF of
// the tile's rows (64); one tcgen05.mma.cta_group::2 with M = 256 multiplies both hidden blocks
// by the whole tile, each SM reading only its half of B from shared memory.  At N = 128 the
// single-CTA kernel's operand reads alone are one 128-byte wavefront per clock; the pair reads
// 96 bytes per clock and writes half the TMA bytes.
template <bool SPLIT, bool PAIR>
__global__ void __launch_bounds__(SPLIT ? kDSplitThreads : kDThreads, 1)
joint_dgrad2_kernel(const __grid_constant__ CUtensorMap map_hi,
                    const __grid_constant__ CUtensorMap map_lo,
                    const __grid_constant__ CUtensorMap map_g, const Dgrad2Params p) {
  extern __shared__ __align__(1024) unsigned char d2smem_raw[];
  unsigned char* base = d2smem_raw + ((1024u - (smem_u32(d2smem_raw) & 1023u)) & 1023u);
  const int V = p.V, H = p.H, C = p.C;
  const int nk = V / 64;                              // K chunks
  const uint32_t a_chunk = 2 * kJB * 128;             // hi | lo of one [128 j x 64 v] chunk
  static_assert(!PAIR || SPLIT, "the CTA-pair variant reads split rows");
  constexpr int kBRows = PAIR ? kTile / 2 : kTile;    // tile rows in this CTA's shared memory
  constexpr int kStages = PAIR ? kDPairStages : kDStages;
  const uint32_t b_stage = 2 * kBRows * 128;          // hi | lo of one [rows x 64 v] chunk
  unsigned char* a_res = base;                        // nk chunks, resident
  unsigned char* b_ring = a_res + (size_t)nk * a_chunk;
  const uint32_t gbring = PAIR ? kGbRingMax : gb_ring_slots(V);
  float* s_gb = reinterpret_cast<float*>(b_ring + kStages * b_stage);       // [gbring][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_gb + gbring * kTile);
  uint64_t* full = bars;                    // [stages]  producers -> MMA
  uint64_t* empty = full + kStages;         // [stages]  MMA (commit) -> producers
  uint64_t* tfull = empty + kStages;        // [2]       MMA (commit) -> epilogue
  uint64_t* tempty = tfull + 2;             // [2]       epilogue -> MMA
  uint64_t* gbfull = tempty + 2;            // [gbring]  producers -> epilogue (grad_blank slice)
  uint64_t* afull = gbfull + kGbRingMax;       // [1]       TMA -> MMA (resident A)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;   // 0 = leader
  const int mc = (SPLIT && !PAIR) ? p.mc : 1;           // multicast cluster (= hidden blocks)
  const uint32_t mrank = mc > 1 ? cluster_ctarank() : 0;
  const uint16_t mmask = (uint16_t)((1u << mc) - 1u);
  const int njb = H / kJB;
  // PAIR: cluster id = blockIdx.x / 2 enumerates (group, pair of hidden blocks)
  const int unit = PAIR ? blockIdx.x >> 1 : blockIdx.x;
  const int nunits = PAIR ? njb / 2 : njb;
  const int jb = PAIR ? (unit % nunits) * 2 + (int)rank : unit % nunits;
  const int group = unit / nunits, ngroups = (PAIR ? gridDim.x >> 1 : gridDim.x) / nunits;
  const long long nblocks = (p.N + kTile - 1) / kTile;
  // this CTA's contiguous range of the (frame block, c) tile sequence
  const long long ttot = nblocks * C;
  const long long t_lo = ttot * group / ngroups, t_hi = ttot * (group + 1) / ngroups;
  const long long nb_lo = t_lo / C;
  const int c_lo = (int)(t_lo - nb_lo * C);

  if (tid == 0) {
    // PAIR: full / tempty / afull are the LEADER's (arrivals from both CTAs)
    for (int s = 0; s < kStages; ++s) {
      mbar_init_n(smem_u32(&full[s]), PAIR ? 2 : (SPLIT ? 1 : kDProducers));
      mbar_init_n(smem_u32(&empty[s]), mc);       // multicast: every CTA's MMAs free the stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init_n(smem_u32(&tfull[a]), 1);
      mbar_init_n(smem_u32(&tempty[a]), PAIR ? 8 : 128);      // PAIR: one arrival per warp
    }
    for (int r = 0; r < kGbRingMax; ++r) mbar_init_n(smem_u32(&gbfull[r]), SPLIT ? 32 : kTile);
    mbar_init_n(smem_u32(afull), PAIR ? 2 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    if (SPLIT) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2cta(smem_u32(tmem_slot), 512);
    else umma::tmem_alloc(smem_u32(tmem_slot), 512);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (PAIR || mc > 1) cluster_sync_all();       // barriers initialised cluster-wide
  umma::fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------ resident A: W_vocab^T block, once
    if (lane == 0 && t_lo < t_hi) {
      const uint32_t bar = smem_u32(afull);
      if (PAIR) mbar_expect_tx_leader(bar, (uint32_t)nk * a_chunk);
      else mbar_expect_tx(bar, (uint32_t)nk * a_chunk);
      for (int kc = 0; kc < nk; ++kc) {
        const uint32_t dst = smem_u32(a_res) + kc * a_chunk;
        if (PAIR) {
          tma_2d_2sm(dst, &map_hi, kc * 64, jb * kJB, bar);
          tma_2d_2sm(dst + kJB * 128, &map_lo, kc * 64, jb * kJB, bar);
        } else {
          tma_2d(dst, &map_hi, kc * 64, jb * kJB, bar);
          tma_2d(dst + kJB * 128, &map_lo, kc * 64, jb * kJB, bar);
        }
      }
    }
    if constexpr (SPLIT) {
      // ---------------------------------------------- B loader: TMA boxes of the split rows
      // lane l ferries grad_blank of rows l, l+32, l+64, l+96 of every tile, loaded one tile
      // ahead; the ring slot is written when the tile's first chunk gets its stage (same
      // lead over the epilogue as the register producers of the fp32 variant)
      auto load_gb = [&](long long nb, int c, float (&v)[4]) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const long long n = nb * kTile + r * 32 + lane;
          v[r] = (nb < nblocks && n < p.N) ? ldg_stream(p.gb + (size_t)n * C + c) : 0.f;
        }
      };
      long long nb = nb_lo;
      int c = c_lo;
      float gbv[4];
      if (t_lo < t_hi) load_gb(nb, c, gbv);
      uint32_t s = 0, par = 1, ringpos = 0;
      for (long long t = t_lo; t < t_hi; ++t) {
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait_parity(smem_u32(&empty[s]), par);
          if (lane == 0) {
            const uint32_t bar = smem_u32(&full[s]);
            const uint32_t dst = smem_u32(b_ring) + s * b_stage;
            if (PAIR) {      // this CTA's 64 rows of the tile, counted on the leader's barrier
              const int n0 = (int)(nb * kTile) + (int)rank * kBRows;
              mbar_expect_tx_leader(bar, b_stage);
              tma_3d_2sm(dst, &map_g, kc * 64, c, n0, bar);
              tma_3d_2sm(dst + kBRows * 128, &map_g, V + kc * 64, c, n0, bar);
            } else if (mc > 1) {
              // my 128 / mc rows of the tile, to the same place in every CTA of the cluster;
              // this CTA's barrier collects the bytes of all mc slices
              const int rows = kTile / mc;
              const int n0 = (int)(nb * kTile) + (int)mrank * rows;
              const uint32_t off = mrank * rows * 128;
              mbar_expect_tx(bar, b_stage);
              tma_3d_mc(dst + off, &map_g, kc * 64, c, n0, bar, mmask);
              tma_3d_mc(dst + kTile * 128 + off, &map_g, V + kc * 64, c, n0, bar, mmask);
            } else {
              mbar_expect_tx(bar, b_stage);
              tma_3d(dst, &map_g, kc * 64, c, (int)(nb * kTile), bar);
              tma_3d(dst + kTile * 128, &map_g, V + kc * 64, c, (int)(nb * kTile), bar);
            }
          }
          if (kc == 0) {
#pragma unroll
            for (int r = 0; r < 4; ++r) s_gb[ringpos * kTile + r * 32 + lane] = gbv[r];
            mbar_arrive(smem_u32(&gbfull[ringpos]));
            if (++ringpos == gbring) ringpos = 0;
          }
          if (++s == kStages) { s = 0; par ^= 1; }
        }
        if (++c == C) { c = 0; ++nb; }
        if (t + 1 < t_hi) load_gb(nb, c, gbv);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0 && t_lo < t_hi && rank == 0) {
      const uint32_t idesc = umma::make_idesc_bf16(PAIR ? 2 * kJB : kJB, kTile);
      auto wait = [&](uint32_t bar, uint32_t parity) {     // PAIR: arrivals come from both CTAs
        if (PAIR) mbar_wait_parity_cluster(bar, parity); else mbar_wait_parity(bar, parity);
      };
      wait(smem_u32(afull), 0);
      uint32_t g = 0;
      const uint32_t ntiles = (uint32_t)(t_hi - t_lo);
      for (uint32_t it = 0; it < ntiles; ++it) {
        const uint32_t acc = it & 1;
        wait(smem_u32(&tempty[acc]), ((it >> 1) & 1) ^ 1);
        umma::fence_after_thread_sync();
        const uint32_t d = tmem + acc * kTile;
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const uint32_t s = g % kStages;
          wait(smem_u32(&full[s]), (g / kStages) & 1);
          umma::fence_after_thread_sync();
          const uint32_t sa = smem_u32(a_res) + kc * a_chunk;
          const uint32_t sb = smem_u32(b_ring) + s * b_stage;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dah = umma::make_smem_desc_sw128(sa + k * 32);
            const uint64_t dal = umma::make_smem_desc_sw128(sa + kJB * 128 + k * 32);
            const uint64_t dbh = umma::make_smem_desc_sw128(sb + k * 32);
            const uint64_t dbl = umma::make_smem_desc_sw128(sb + kBRows * 128 + k * 32);
            if (PAIR) {
              mma_bf16_2cta(d, dah, dbh, idesc, (kc | k) > 0);
              mma_bf16_2cta(d, dah, dbl, idesc, 1);
              mma_bf16_2cta(d, dal, dbh, idesc, 1);
            } else {
              umma::mma_bf16(d, dah, dbh, idesc, (kc | k) > 0);
              umma::mma_bf16(d, dah, dbl, idesc, 1);
              umma::mma_bf16(d, dal, dbh, idesc, 1);
            }
          }
          if (PAIR) commit_2cta(smem_u32(&empty[s]));
          else if (mc > 1) commit_mc(smem_u32(&empty[s]), mmask);
          else umma::commit(smem_u32(&empty[s]));
        }
        if (PAIR) commit_2cta(smem_u32(&tfull[acc])); else umma::commit(smem_u32(&tfull[acc]));
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue
    constexpr int kCols = kTile;
    constexpr int col0 = 0;
    const int quad = warp & 3;                          // TMEM lane quadrant of this warp
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int jg = jb * kJB + quad * 32 + lane;         // hidden unit of this thread
    const float wbj = p.w_blank[jg];
    const uint32_t t_pf = tmem + 2 * kTile + lane_base, t_acc = tmem + 3 * kTile + lane_base;
    uint32_t it = 0;
    long long t = t_lo, nb = nb_lo;
    int c = c_lo;
    while (t < t_hi) {
      const long long n0 = nb * kTile;
      const int nvalid = (int)min((long long)kTile, p.N - n0);
      const int c_end = (int)min((long long)C, c + (t_hi - t));
      // item prologue: pf[n0 + i, jg] -> TMEM, running sums := 0
      for (int c0 = col0; c0 < col0 + kCols; c0 += 16) {
        float v[16], z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          v[i] = (c0 + i < nvalid) ? __ldg(p.pf + (size_t)(n0 + c0 + i) * H + jg) : 0.f;
          z[i] = 0.f;
        }
        tmem_st16(t_pf + c0, v);
        tmem_st16(t_acc + c0, z);
      }
      tmem_wait_st();
      float pcj = __ldg(p.pc + (size_t)c * H + jg);
      for (; c < c_end; ++c, ++t, ++it) {
        const uint32_t acc = it & 1, ring = it % gbring;
        const float pc_cur = pcj;
        if (c + 1 < c_end) pcj = __ldg(p.pc + (size_t)(c + 1) * H + jg);
        mbar_wait_parity(smem_u32(&gbfull[ring]), (it / gbring) & 1);
        mbar_wait_parity(smem_u32(&tfull[acc]), (it >> 1) & 1);
        umma::fence_after_thread_sync();
        const float* gbr = s_gb + ring * kTile;
        const uint32_t t_d = tmem + acc * kTile + lane_base;
        float csum = 0.f;
        // two register sets of GW columns: the loads of one are in flight while the other is
        // being computed (tcgen05.wait::ld after the compute covers them).  GW = 16 in the
        // split-row variant (192 threads: registers to spare, half as many waits per tile),
        // 8 where 22 warps share the register file.
        constexpr int GW = SPLIT ? 16 : 8;
        float d0[GW], f0[GW], a0[GW], d1[GW], f1[GW], a1[GW];
        auto ld = [&](uint32_t taddr, float (&v)[GW]) {
          if constexpr (GW == 16) tmem_ld16(taddr, v); else tmem_ld8(taddr, v);
        };
        // tanh' = 1 - tanh^2 = 4 r (1 - r) with r = 1 / (1 + E_c E_f): one MUFU op per element;
        // the factor 4 is applied once, when csum / the running sums are flushed
        auto compute = [&](int c0, const float (&d)[GW], const float (&f)[GW], float (&a)[GW]) {
#pragma unroll
          for (int i = 0; i < GW; ++i) {
            const float x = fmaf(gbr[c0 + i], wbj, d[i]);
            const float r = rcp_1p(pc_cur * f[i]);
            const float gp = x * fmaf(-r, r, r);
            csum += gp;
            a[i] += gp;
          }
          if constexpr (GW == 16) tmem_st16(t_acc + c0, a); else tmem_st8(t_acc + c0, a);
        };
        ld(t_d + col0, d0); ld(t_pf + col0, f0); ld(t_acc + col0, a0);
#pragma unroll 1
        for (int c0 = col0; c0 < col0 + kCols; c0 += 2 * GW) {
          tmem_wait_ld();
          ld(t_d + c0 + GW, d1); ld(t_pf + c0 + GW, f1); ld(t_acc + c0 + GW, a1);
          compute(c0, d0, f0, a0);
          tmem_wait_ld();
          if (c0 + 2 * GW < col0 + kCols) {
            ld(t_d + c0 + 2 * GW, d0); ld(t_pf + c0 + 2 * GW, f0); ld(t_acc + c0 + 2 * GW, a0);
          }
          compute(c0 + GW, d1, f1, a1);
        }
        tmem_wait_st();
        umma::fence_before_thread_sync();
        if (PAIR) {
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(smem_u32(&tempty[acc]));
        } else {
          mbar_arrive(smem_u32(&tempty[acc]));
        }
        atomicAdd(p.gpc + (size_t)c * H + jg, 4.f * csum);
      }
      // item epilogue: grad_proj_frame[n0 + i, jg] += running sums
      for (int c0 = col0; c0 < col0 + kCols; c0 += 16) {
        float a[16];
        tmem_ld16(t_acc + c0, a);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < nvalid) atomicAdd(p.gpf + (size_t)(n0 + c0 + i) * H + jg, 4.f * a[i]);
      }
      c = 0;
      ++nb;
    }
  } else if constexpr (!SPLIT) {
    // ---------------------------------------------------------------- B producers
    // 8 lanes per joint row (one 256-bit load each: 256 contiguous bytes of grad_lexical per row
    // and K chunk), 4 rows per warp, kDPasses passes over the 128 rows of the tile.
    //  * Along the CTA's tile sequence the chunks of a row are CONTIGUOUS in memory
    //    ((c, kc) -> offset (c * V/64 + kc) * 64 floats), so a thread keeps one pointer per row
    //    and adds 64 floats per chunk; only a new frame block recomputes it.  Rows past N point
    //    at a zero line with step 0.
    //  * Registers hold one chunk ahead of the conversion (two buffers, ping-pong); the HBM
    //    latency is covered by L2 prefetches kPrefTiles tiles ahead of the loads.
    const int pw = warp - 6;
    const int ch = lane & 7, rsub = lane >> 3;
    const uint32_t gtot = (uint32_t)(t_hi - t_lo) * (uint32_t)nk;
    constexpr int kPrefTiles = 2;
    const float* pp[kDPasses];
    const float* gbp[kDPasses];
    int pstep[kDPasses];                      // 64 floats per chunk, 0 for rows past N
    long long pnb = nb_lo;
    uint32_t left = 0, pkc = 0;               // chunks left in the frame block / K chunk, both at
                                              // the LOAD position
    auto place = [&](long long nb, int c) {
#pragma unroll
      for (int r = 0; r < kDPasses; ++r) {
        const long long n = nb * kTile + r * (kDProdWarps * 4) + pw * 4 + rsub;
        const bool valid = n < p.N;
        pp[r] = valid ? p.gl + ((size_t)n * C + c) * V + ch * 8 : lt_zero_line + ch * 8;
        gbp[r] = valid ? p.gb + (size_t)n * C + c : lt_zero_line;
        pstep[r] = valid ? 64 : 0;
      }
      left = (uint32_t)(C - c) * (uint32_t)nk;
    };
    place(nb_lo, c_lo);
    struct Buf { float x[kDPasses][8]; float gb[kDPasses]; };
    auto load = [&](Buf& b) {
#pragma unroll
      for (int r = 0; r < kDPasses; ++r) ldg_stream8(pp[r], b.x[r]);
      if (pkc == 0 && ch == 0) {
#pragma unroll
        for (int r = 0; r < kDPasses; ++r) b.gb[r] = ldg_stream(gbp[r]);
      }
      if (left > (uint32_t)(kPrefTiles * nk) && (ch & 3) == 0) {
#pragma unroll
        for (int r = 0; r < kDPasses; ++r)
          if (pstep[r])
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp[r] + kPrefTiles * V));
      }
#pragma unroll
      for (int r = 0; r < kDPasses; ++r) pp[r] += pstep[r];
      if (++pkc == (uint32_t)nk) {
        pkc = 0;
#pragma unroll
        for (int r = 0; r < kDPasses; ++r) gbp[r] += pstep[r] >> 6;
      }
      if (--left == 0) { ++pnb; place(pnb, 0); }
    };
    uint32_t s = 0, par = 1;                  // ring stage / parity to wait for on empty[s]
    uint32_t ckc = 0, ringpos = 0;            // K chunk and grad_blank ring slot, CONVERT position
    const uint32_t row0_off = umma::swizzled_offset(pw * 4 + rsub, ch);   // row r: + r * 64 rows
    auto convert = [&](const Buf& b) {
      uint4 hi[kDPasses], lo[kDPasses];
#pragma unroll
      for (int r = 0; r < kDPasses; ++r) umma::split_pack8(b.x[r], hi[r], lo[r]);
      if (ckc == 0 && ch == 0) {              // grad_blank slice of this tile for the epilogue
#pragma unroll
        for (int r = 0; r < kDPasses; ++r)
          s_gb[ringpos * kTile + r * (kDProdWarps * 4) + pw * 4 + rsub] = b.gb[r];
#pragma unroll
        for (int r = 0; r < kDPasses; ++r) mbar_arrive(smem_u32(&gbfull[ringpos]));
      }
      mbar_wait_parity(smem_u32(&empty[s]), par);
      unsigned char* b_hi = b_ring + s * b_stage + row0_off;
#pragma unroll
      for (int r = 0; r < kDPasses; ++r) {
        // rows r * 64 apart: same swizzle phase (64 % 8 == 0), 64 * 128 bytes further
        *reinterpret_cast<uint4*>(b_hi + r * (kDProdWarps * 4) * 128) = hi[r];
        *reinterpret_cast<uint4*>(b_hi + kTile * 128 + r * (kDProdWarps * 4) * 128) = lo[r];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(smem_u32(&full[s]));
      if (++ckc == (uint32_t)nk) { ckc = 0; if (++ringpos == gbring) ringpos = 0; }
      if (++s == kDStages) { s = 0; par ^= 1; }
    };
    Buf bufa, bufb;
    if (gtot > 0) load(bufa);
    uint32_t g = 0;
#pragma unroll 1
    for (; g + 1 < gtot; g += 2) {
      load(bufb);
      convert(bufa);
      if (g + 2 < gtot) load(bufa);
      convert(bufb);
    }
    if (g < gtot) convert(bufa);
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  if (PAIR || mc > 1) cluster_sync_all();       // no CTA leaves while peers still signal it
  if (warp == 1) {
    if (PAIR) tmem_dealloc_2cta(tmem, 512); else umma::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

// split rows: CTAs per cluster sharing each B tile by TMA multicast = number of hidden blocks
// (the CTAs of a group walk the same tile sequence), when that is a legal cluster size.
// Opt-in: it cuts the L2 -> SM operand traffic by the cluster size (6.9 TB/s without it) but
// couples the three-stage rings of four CTAs -- a stage is refilled only when ALL of them have
// consumed it -- and measured 9.3 ms against 4.9 ms.
int joint_dgrad2_multicast(int H, int V) {
  (void)V;
  if (!option(OPT_JOINT_DGRAD_MULTICAST) || option(OPT_JOINT_DGRAD_PAIR)) return 1;
  const int njb = H / kJB;
  return (njb == 2 || njb == 4 || njb == 8) ? njb : 1;
}

// CTA-pair variant of the split-row kernel: pairs of hidden blocks (opt-in while it is measured)
bool joint_dgrad2_pair(int H, int V) {
  return option(OPT_JOINT_DGRAD_PAIR) && H % (2 * kJB) == 0 && V % 64 == 0;
}

bool joint_dgrad2_supported(int64_t N, int C, int H, int V, const void* gl, const void* pc,
                            const void* pf) {
  if (option(OPT_JOINT_SIMT) || option(OPT_JOINT_DGRAD_V1)) return false;
  if (V % 64 != 0 || V < 64 || V > 256) return false;
  if (H % 128 != 0 || H > 4096) return false;
  if (N < 1 || C < 1) return false;
  auto al = [](const void* q, int a) { return reinterpret_cast<uintptr_t>(q) % a == 0; };
  return al(gl, 32) && al(pc, 16) && al(pf, 16);      // grad_lexical: 256-bit loads
}

// whi / wlo: W_vocab^T [H, V] as bf16 hi / lo (transpose_split_kernel), map_* their tensor maps
// with box [64 x 128].  split != 0: gl holds rows of [V bf16 hi | V bf16 lo] and map_g is its
// 3-D tensor map {2V, C, N} with box {64, 1, 128}; otherwise gl is fp32 and map_g is unused.
int joint_dgrad2_launch(const CUtensorMap& map_hi, const CUtensorMap& map_lo,
                        const CUtensorMap& map_g, int split, const float* pc,
                        const float* pf, const float* wb, const float* gb, const float* gl,
                        int64_t N, int C, int H, int V, float* gpc, float* gpf,
                        cudaStream_t stream) {
  Dgrad2Params p = {};
  p.pc = pc; p.pf = pf; p.w_blank = wb; p.gl = gl; p.gb = gb;
  p.N = N; p.C = C; p.H = H; p.V = V; p.gpc = gpc; p.gpf = gpf; p.mc = 1;
  const int nk = V / 64;
  const bool pair = split && joint_dgrad2_pair(H, V);
  const size_t smem = (size_t)nk * 2 * kJB * 128 +
                      (pair ? (size_t)kDPairStages * kTile * 128
                            : (size_t)kDStages * 2 * kTile * 128) +
                      sizeof(float) * (pair ? kGbRingMax : gb_ring_slots(V)) * kTile + 8 * 32 +
                      16 + 1024;
  int dev = 0, sms = 0;
  LT_CUDA(cudaGetDevice(&dev));
  LT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int njb = H / kJB;
  int groups = sms / njb;
  if (groups < 1) groups = 1;
  const long long nblocks = (N + kTile - 1) / kTile;
  if (groups > nblocks) groups = (int)nblocks;
  if (pair) {
    LT_CUDA(cudaFuncSetAttribute(joint_dgrad2_kernel<true, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(groups * njb));      // clusters of two CTAs = hidden-block pairs
    cfg.blockDim = dim3(kDSplitThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    LT_CUDA(cudaLaunchKernelEx(&cfg, joint_dgrad2_kernel<true, true>, map_hi, map_lo, map_g, p));
  } else if (split) {
    LT_CUDA(cudaFuncSetAttribute(joint_dgrad2_kernel<true, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.mc = joint_dgrad2_multicast(H, V);
    if (p.mc > 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(groups * njb));    // cluster = the njb hidden blocks of a group
      cfg.blockDim = dim3(kDSplitThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)p.mc;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      LT_CUDA(cudaLaunchKernelEx(&cfg, joint_dgrad2_kernel<true, false>, map_hi, map_lo, map_g, p));
    } else {
      joint_dgrad2_kernel<true, false><<<groups * njb, kDSplitThreads, smem, stream>>>(
          map_hi, map_lo, map_g, p);
    }
  } else {
    LT_CUDA(cudaFuncSetAttribute(joint_dgrad2_kernel<false, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    joint_dgrad2_kernel<false, false><<<groups * njb, kDThreads, smem, stream>>>(map_hi, map_lo,
                                                                                 map_g, p);
  }
  LT_LAUNCHED();
  return LT_OK;
}

}  // namespace lt
