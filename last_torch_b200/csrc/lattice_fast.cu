// K1/K2 fast path for the headline shape: FullNGram context_size 1 (bigram,
// C = V + 1 states), FrameDependent alignment, V in {64, 128, 192, 256}.
//
// B200-first design (see DESIGN.md "Fast path"):
//   * one thread-block CLUSTER of S = V/64 CTAs per utterance, persistent over
//     all T frames; alpha / beta live in shared memory (a full replica per CTA);
//   * arc weights never depend on alpha, so they are streamed by TMA
//     (cp.async.bulk[.tensor]) into a shared-memory ring several frames AHEAD of
//     the recursion, completion tracked by mbarriers -- the HBM stream is
//     decoupled from the sequential dependency chain;
//   * forward: CTA r owns 64 destination columns of the [V+1, V] frame tile
//     (2-D TMA box [V rows x 64 cols]); column log-sum-exp is two passes over
//     REGISTERS (max, then one ex2 per arc), combined across warps in smem;
//   * backward: CTA r owns 64 source rows (one contiguous 64*V*4-byte bulk
//     copy); 8 lanes per row, conflict-free 128-bit smem reads, ONE ex2 per arc
//     shared by the row log-sum-exp (beta) and the arc posterior, gradients
//     written straight from registers with coalesced 128-bit streaming stores;
//   * the new 64-entry slice of alpha / beta is all-gathered with DSMEM stores
//     to every CTA of the cluster, followed by one cluster barrier per frame.
//
// Reference semantics: lattices.py:436-462 + alignments.py:294-297 (forward),
// alignments.py:300-318 + lattices.py:775-779 (backward), contexts.py:207-256.
#include <cuda.h>

#include "common.cuh"
#include "params.cuh"

namespace lt {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kColsPerCta = 64;

// ----------------------------------------------------------------- PTX helpers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LT_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LT_WAIT_DONE;\n"
      "bra LT_WAIT_LOOP;\n"
      "LT_WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// one MUFU.EX2 (max rel. error 2^-22; results below 2^-126 flush to +0)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// All-gather of the recursion state without a cluster barrier: a DSMEM store
// that also completes 4 bytes of the destination CTA's mbarrier transaction.
// The receiver arms the barrier with expect_tx(C * 4) once per frame and waits
// on it; no memory fence is involved (the mbarrier orders the data).
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::
                   "r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void xchg_store(float* base, int idx, float v, uint64_t* bar,
                                           uint32_t nrank) {
  const uint32_t a = smem_u32(base + idx), bb = smem_u32(bar);
  for (uint32_t r = 0; r < nrank; ++r)
    st_async_f32(map_shared_rank(a, r), v, map_shared_rank(bb, r));
}

// The Log-semiring fast kernels keep alpha / beta in LOG2 units on chip:
//   y = fma(w, log2(e), alpha2)  is one FFMA whose rounding error is the fp32
//   representation error of the sum itself (same as the reference's a + w), and
//   the max-shifted exponent y - m is then an exact difference fed to ex2.
// (m, s) pair of a running log2-sum-exp2: value = msafe(m) + log2(s).
__device__ __forceinline__ void lse2_merge(float& m, float& s, float om, float os) {
  const float mn = fmaxf(m, om);
  const float mns = msafe(mn);
  const float sa = (m == neg_inf()) ? 0.f : ex2(msafe(m) - mns);
  const float sb = (om == neg_inf()) ? 0.f : ex2(msafe(om) - mns);
  s = s * sa + os * sb;
  m = mn;
}
// log2(2^a + 2^b) with the non-finite-max rule of semirings.py:250-251
__device__ __forceinline__ float log2_add_exp2(float a, float b) {
  const float c = fmaxf(a, b);
  const float cs = msafe(c);
  return cs + __log2f(ex2(a - cs) + ex2(b - cs));
}
template <int SR> __device__ __forceinline__ float to_dom(float x) {
  return SR == LT_LOG ? x * kLog2e : x;
}
template <int SR> __device__ __forceinline__ float from_dom(float x) {
  return SR == LT_LOG ? x * kLn2 : x;
}
// weight (x) destination value: Log works in log2 units (w * log2e + v), Real is w * v
template <int SR> __device__ __forceinline__ float arc(float w, float v) {
  return SR == LT_LOG ? fmaf(w, kLog2e, v) : w * v;
}

__device__ __forceinline__ void bcast_f32(float* base, int idx, float v, uint32_t nrank) {
  const uint32_t a = smem_u32(base + idx);
  for (uint32_t r = 0; r < nrank; ++r) st_shared_cluster_f32(map_shared_rank(a, r), v);
}

struct FastFwdParams {
  int V, B, T, stages;
  const float* blank;
  const float* lexical;
  const int32_t* num_frames;
  const float* alpha_init;
  float* dist;
  float* alphas;
  float* alpha_final;
  int16_t* backptr;
};

// ============================================================== forward (K1) ==
// VD = V / 64 = cluster size.  Tile stage: [V rows][64 cols] fp32.
template <int SR, int VD>
__global__ void __launch_bounds__(kThreads, 1)
lattice_forward_fast(const __grid_constant__ CUtensorMap tmap, const FastFwdParams p) {
  using S = Sr<SR>;
  constexpr int V = 64 * VD;
  constexpr int C = V + 1;
  constexpr int CP = (C + 3) & ~3;
  constexpr int RPT = 2 * VD;                   // rows per thread (V / 32 row groups)
  constexpr uint32_t kStageBytes = V * kColsPerCta * 4;
  extern __shared__ __align__(128) unsigned char fsmem[];
  const int NS = p.stages;
  float* tiles = reinterpret_cast<float*>(fsmem);
  float* alpha_buf = reinterpret_cast<float*>(fsmem + (size_t)NS * kStageBytes);
  float* part_m = alpha_buf + 2 * CP;
  float* part_s = part_m + kWarps * kColsPerCta;
  uint64_t* bars = reinterpret_cast<uint64_t*>(part_s + kWarps * kColsPerCta);

  const uint32_t nrank = VD;
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / VD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cg = tid & 15;                      // column group: columns 4cg .. 4cg+3
  const int rg = tid >> 4;                      // row group: rows rg*RPT .. +RPT-1
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;
  const int col0 = rank * kColsPerCta;          // first tile column owned by this CTA

  uint64_t* xbar = bars + NS;       // xbar[i]: "alpha buffer i has received all C entries"
  if (tid == 0) {
    prefetch_tensormap(&tmap);
    for (int s = 0; s < NS + 2; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = tid; c < CP; c += kThreads) {
    float v = S::zero();
    if (c < C) v = p.alpha_init ? p.alpha_init[(size_t)b * C + c] : (c == 0 ? S::one() : S::zero());
    alpha_buf[c] = to_dom<SR>(v);
    alpha_buf[CP + c] = S::zero();
  }
  __syncthreads();
  cluster_sync_all();

  // prologue: fill the ring
  if (tid == 0) {
    for (int s = 0; s < NS && s < nf; ++s) {
      const uint32_t bar = smem_u32(&bars[s]);
      mbar_arrive_expect_tx(bar, kStageBytes);
      tma_load_2d(smem_u32(tiles) + s * kStageBytes, &tmap, col0, (int)((bt0 + s) * C), bar);
    }
  }

  // finalizer threads: tid < 64 own destination q = 1 + col0 + tid; thread 64 of
  // rank 0 owns state 0 (no incoming lexical arc, contexts.py:217-218).
  // Tree finalizer: 4 threads per destination column (tid < 256) each merge 4 of the 16
  // per-warp partials, two shuffle merges combine them, the part == 0 thread publishes.
  const bool in_tree = tid < 4 * kColsPerCta;
  const int fcol = tid >> 2, fpart = tid & 3;
  const bool is_fin = in_tree && fpart == 0;
  const bool is_q0 = (rank == 0 && tid == 4 * kColsPerCta);
  const int q = is_fin ? 1 + col0 + fcol : 0;
  float nblank = 0.f, ntail = 0.f;              // prefetched blank[t][q], lexical[t][V][col]
  if (nf > 0) {
    if (is_fin) {
      nblank = ldg_stream(p.blank + bt0 * C + q);
      ntail = ldg_stream(p.lexical + bt0 * (size_t)C * V + (size_t)V * V + col0 + fcol);
    } else if (is_q0) {
      nblank = ldg_stream(p.blank + bt0 * C);
    }
  }

  for (int t = 0; t < nf; ++t) {
    const int stage = t % NS;
    const uint32_t parity = (t / NS) & 1;
    float* cur = alpha_buf + (t & 1) * CP;
    float* nxt = alpha_buf + ((t + 1) & 1) * CP;
    // alpha_t (t > 0) is complete once every CTA's st.async stores have landed
    if (t > 0) mbar_wait(smem_u32(&xbar[t & 1]), ((t - 1) >> 1) & 1);
    if (tid == 0) mbar_arrive_expect_tx(smem_u32(&xbar[(t + 1) & 1]), C * 4);
    const float cblank = nblank, ctail = ntail;
    if (t + 1 < nf) {
      if (is_fin) {
        nblank = ldg_stream(p.blank + (bt0 + t + 1) * C + q);
        ntail = ldg_stream(p.lexical + (bt0 + t + 1) * (size_t)C * V + (size_t)V * V + col0 + fcol);
      } else if (is_q0) {
        nblank = ldg_stream(p.blank + (bt0 + t + 1) * C);
      }
    }
    if (p.alphas && (is_fin || is_q0)) p.alphas[(bt0 + t) * C + q] = from_dom<SR>(cur[q]);

    mbar_wait(smem_u32(&bars[stage]), parity);
    const float* tile = tiles + (size_t)stage * (kStageBytes / 4);
    const int r0 = rg * RPT;
    float a[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) a[i] = cur[r0 + i];
    float4 x[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
      x[i] = *reinterpret_cast<const float4*>(tile + (size_t)(r0 + i) * kColsPerCta + cg * 4);

    float pm[4], ps[4];
    if constexpr (SR == LT_LOG) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        x[i].x = fmaf(x[i].x, kLog2e, a[i]); x[i].y = fmaf(x[i].y, kLog2e, a[i]);
        x[i].z = fmaf(x[i].z, kLog2e, a[i]); x[i].w = fmaf(x[i].w, kLog2e, a[i]);
      }
      pm[0] = x[0].x; pm[1] = x[0].y; pm[2] = x[0].z; pm[3] = x[0].w;
#pragma unroll
      for (int i = 1; i < RPT; ++i) {
        pm[0] = fmaxf(pm[0], x[i].x); pm[1] = fmaxf(pm[1], x[i].y);
        pm[2] = fmaxf(pm[2], x[i].z); pm[3] = fmaxf(pm[3], x[i].w);
      }
      float ms[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { ms[j] = msafe(pm[j]); ps[j] = 0.f; }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        ps[0] += ex2(x[i].x - ms[0]);
        ps[1] += ex2(x[i].y - ms[1]);
        ps[2] += ex2(x[i].z - ms[2]);
        ps[3] += ex2(x[i].w - ms[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float om = __shfl_xor_sync(0xffffffffu, pm[j], 16);
        const float os = __shfl_xor_sync(0xffffffffu, ps[j], 16);
        lse2_merge(pm[j], ps[j], om, os);
      }
    } else if constexpr (SR == LT_MAXTROPICAL) {
      // (max, first arg-max row): rows ascend inside a thread, ties keep the lower row
#pragma unroll
      for (int j = 0; j < 4; ++j) { pm[j] = neg_inf(); ps[j] = __int_as_float(r0); }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float v0 = x[i].x + a[i], v1 = x[i].y + a[i], v2 = x[i].z + a[i], v3 = x[i].w + a[i];
        if (v0 > pm[0]) { pm[0] = v0; ps[0] = __int_as_float(r0 + i); }
        if (v1 > pm[1]) { pm[1] = v1; ps[1] = __int_as_float(r0 + i); }
        if (v2 > pm[2]) { pm[2] = v2; ps[2] = __int_as_float(r0 + i); }
        if (v3 > pm[3]) { pm[3] = v3; ps[3] = __int_as_float(r0 + i); }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float om = __shfl_xor_sync(0xffffffffu, pm[j], 16);
        const int oa = __shfl_xor_sync(0xffffffffu, __float_as_int(ps[j]), 16);
        const int ma = __float_as_int(ps[j]);
        if (om > pm[j] || (om == pm[j] && oa < ma)) { pm[j] = om; ps[j] = __int_as_float(oa); }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { pm[j] = 0.f; ps[j] = 0.f; }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        pm[0] = fmaf(a[i], x[i].x, pm[0]); pm[1] = fmaf(a[i], x[i].y, pm[1]);
        pm[2] = fmaf(a[i], x[i].z, pm[2]); pm[3] = fmaf(a[i], x[i].w, pm[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) pm[j] += __shfl_xor_sync(0xffffffffu, pm[j], 16);
    }
    if (lane < 16) {
      *reinterpret_cast<float4*>(part_m + warp * kColsPerCta + cg * 4) =
          make_float4(pm[0], pm[1], pm[2], pm[3]);
      if constexpr (SR != LT_REAL)
        *reinterpret_cast<float4*>(part_s + warp * kColsPerCta + cg * 4) =
            make_float4(ps[0], ps[1], ps[2], ps[3]);
    }
    __syncthreads();   // partials visible; every thread is done with this tile stage

    if (tid == 0 && t + NS < nf) {
      const uint32_t bar = smem_u32(&bars[stage]);
      mbar_arrive_expect_tx(bar, kStageBytes);
      tma_load_2d(smem_u32(tiles) + stage * kStageBytes, &tmap, col0, (int)((bt0 + t + NS) * C), bar);
    }

    if (in_tree) {                      // warps 0-7, warp-uniform
      const int w0 = fpart * 4;
      float m = part_m[w0 * kColsPerCta + fcol];
      float s = (SR == LT_REAL) ? 0.f : part_s[w0 * kColsPerCta + fcol];
#pragma unroll
      for (int w = 1; w < 4; ++w) {
        const float om = part_m[(w0 + w) * kColsPerCta + fcol];
        if constexpr (SR == LT_LOG) {
          lse2_merge(m, s, om, part_s[(w0 + w) * kColsPerCta + fcol]);
        } else if constexpr (SR == LT_MAXTROPICAL) {
          const int oa = __float_as_int(part_s[(w0 + w) * kColsPerCta + fcol]);
          if (om > m || (om == m && oa < __float_as_int(s))) { m = om; s = __int_as_float(oa); }
        } else {
          m += om;
        }
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const float os = __shfl_xor_sync(0xffffffffu, s, o);
        if constexpr (SR == LT_LOG) {
          lse2_merge(m, s, om, os);
        } else if constexpr (SR == LT_MAXTROPICAL) {
          if (om > m || (om == m && __float_as_int(os) < __float_as_int(s))) { m = om; s = os; }
        } else {
          m += om;
        }
      }
      if (is_fin) {
        const float ab = S::times(cur[q], to_dom<SR>(cblank));
        const float xt = S::times(cur[V], to_dom<SR>(ctail));   // source row V (not in the TMA box)
        float v;
        if constexpr (SR == LT_LOG) {
          lse2_merge(m, s, xt, xt == neg_inf() ? 0.f : 1.f);
          v = log2_add_exp2(ab, msafe(m) + __log2f(s));
        } else if constexpr (SR == LT_MAXTROPICAL) {
          int am = __float_as_int(s);
          if (xt > m) { m = xt; am = V; }
          const bool take_blank = ab >= m;               // semirings.py:363
          v = take_blank ? ab : m;
          if (p.backptr) p.backptr[(bt0 + t) * C + q] = take_blank ? (int16_t)-1 : (int16_t)am;
        } else {
          v = ab + (m + xt);
        }
        xchg_store(nxt, q, v, &xbar[(t + 1) & 1], nrank);
      }
    } else if (is_q0) {
      const float v = S::times(cur[0], to_dom<SR>(cblank));
      if constexpr (SR == LT_MAXTROPICAL) { if (p.backptr) p.backptr[(bt0 + t) * C] = (int16_t)-1; }
      xchg_store(nxt, 0, v, &xbar[(t + 1) & 1], nrank);
    }
  }
  float* cur = alpha_buf + (nf & 1) * CP;
  if (nf > 0) mbar_wait(smem_u32(&xbar[nf & 1]), ((nf - 1) >> 1) & 1);

  // padding frames keep alpha (lattices.py:460-461) and are still recorded (:462)
  if (is_fin || is_q0) {
    if (p.alphas)
      for (int t = nf; t < p.T; ++t) p.alphas[(bt0 + t) * C + q] = from_dom<SR>(cur[q]);
    if (p.alpha_final) p.alpha_final[(size_t)b * C + q] = from_dom<SR>(cur[q]);
  }
  if (rank == 0) {       // dist = (+)_c alpha_T[c]  (lattices.py:496)
    float* red = part_m;
    if constexpr (SR == LT_LOG) {
      float m = neg_inf();
      for (int c = tid; c < C; c += kThreads) m = fmaxf(m, cur[c]);
      red[tid] = m;
      __syncthreads();
      for (int s = kThreads >> 1; s > 0; s >>= 1) {
        if (tid < s) red[tid] = fmaxf(red[tid], red[tid + s]);
        __syncthreads();
      }
      const float ms = msafe(red[0]);
      __syncthreads();
      float s = 0.f;
      for (int c = tid; c < C; c += kThreads) s += ex2(cur[c] - ms);      // log2 domain
      red[tid] = s;
      __syncthreads();
      for (int st = kThreads >> 1; st > 0; st >>= 1) {
        if (tid < st) red[tid] += red[tid + st];
        __syncthreads();
      }
      if (tid == 0) p.dist[b] = (ms + __log2f(red[0])) * kLn2;
    } else {
      float m = (SR == LT_REAL) ? 0.f : neg_inf();
      for (int c = tid; c < C; c += kThreads) m = S::plus(m, cur[c]);
      red[tid] = m;
      __syncthreads();
      for (int s = kThreads >> 1; s > 0; s >>= 1) {
        if (tid < s) red[tid] = S::plus(red[tid], red[tid + s]);
        __syncthreads();
      }
      if (tid == 0) p.dist[b] = red[0];
    }
  }
  cluster_sync_all();
}

// ============================================================= backward (K2) ==
struct FastBwdParams {
  int V, B, T, stages;
  const float* blank;
  const float* lexical;
  const int32_t* num_frames;
  const float* alphas;
  const float* dist;
  const float* grad_dist;
  float* grad_blank;
  float* grad_lexical;
  float* beta_final;
};

// One row (source state) of the frame handled by `NL` cooperating lanes:
// x = lex + beta'[next]; Log: m, s, posterior = e * rs; Real: dot product.
template <int SR, int VD>
__global__ void __launch_bounds__(kThreads, 1)
lattice_backward_fast(const FastBwdParams p) {
  using S = Sr<SR>;
  constexpr int V = 64 * VD;
  constexpr int C = V + 1;
  constexpr int CH = 2 * VD;                    // float4 chunks per lane (8 lanes per row)
  constexpr int kRows = 64;                     // rows per CTA (+ tail row V on the last rank)
  constexpr uint32_t kSlabBytes = kRows * V * 4;
  constexpr uint32_t kStageBytes = kSlabBytes + V * 4;
  constexpr int BP = ((C + 3 + 3) & ~3) + 4;    // beta buffer: entry q at index 3 + q
  extern __shared__ __align__(128) unsigned char bsmem[];
  const int NS = p.stages;
  float* tiles = reinterpret_cast<float*>(bsmem);
  float* beta_buf = reinterpret_cast<float*>(bsmem + (size_t)NS * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(beta_buf + 2 * BP);

  const uint32_t nrank = VD;
  const uint32_t rank = cluster_ctarank();
  const bool last_rank = rank == nrank - 1;
  const int b = blockIdx.x / VD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane >> 3, sl = lane & 7;     // row within the warp, lane within the row
  const int row = warp * 4 + sub;               // local row 0..63
  const int prow = rank * kRows + row;          // source state
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;
  const float logz = p.dist[b];
  const float logz2 = logz * kLog2e;            // Log: everything on chip is in log2 units
  const float gscale = p.grad_dist ? p.grad_dist[b] : 1.f;
  const bool scale_ok = (SR != LT_LOG) || is_finite(logz);
  const uint32_t stage_tx = last_rank ? kStageBytes : kSlabBytes;

  uint64_t* xbar = bars + NS;       // xbar[i]: "beta buffer i has received all C entries"
  if (tid == 0) {
    for (int s = 0; s < NS + 2; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = tid; c < 2 * BP; c += kThreads) beta_buf[c] = S::one();   // lattices.py:789-790
  __syncthreads();
  cluster_sync_all();

  auto issue = [&](int it) {       // iteration `it` processes frame t = nf - 1 - it
    const int t = nf - 1 - it;
    const int s = it % NS;
    const uint32_t bar = smem_u32(&bars[s]);
    const float* src = p.lexical + (bt0 + t) * (size_t)C * V + (size_t)rank * kRows * V;
    const uint32_t dst = smem_u32(tiles) + s * kStageBytes;
    mbar_arrive_expect_tx(bar, stage_tx);
    bulk_load_1d(dst, src, kSlabBytes, bar);
    if (last_rank)
      bulk_load_1d(dst + kSlabBytes, p.lexical + (bt0 + t) * (size_t)C * V + (size_t)V * V, V * 4, bar);
  };
  if (tid == 0)
    for (int it = 0; it < NS && it < nf; ++it) issue(it);

  // padding frames: zero gradients (lattices.py:775-779)
  for (int t = nf; t < p.T; ++t) {
    float4* gl = reinterpret_cast<float4*>(p.grad_lexical + (bt0 + t) * (size_t)C * V +
                                           (size_t)rank * kRows * V);
    for (int i = tid; i < kRows * V / 4; i += kThreads) stg_stream4(reinterpret_cast<float*>(gl + i), make_float4(0, 0, 0, 0));
    if (tid < kRows) p.grad_blank[(bt0 + t) * C + rank * kRows + tid] = 0.f;
    if (last_rank) {
      float* tail = p.grad_lexical + (bt0 + t) * (size_t)C * V + (size_t)V * V;
      for (int i = tid; i < V; i += kThreads) tail[i] = 0.f;
      if (tid == 0) p.grad_blank[(bt0 + t) * C + V] = 0.f;
    }
  }

  // row owners (lane sl == 0) prefetch alpha_t[p], blank_t[p] one frame ahead;
  // warp 0 lane 0 of the last rank also owns the tail row V.
  const bool owner = sl == 0;
  const bool tail_owner = last_rank && warp == 0 && lane == 0;
  float n_alpha = 0.f, n_blank = 0.f, n_talpha = 0.f, n_tblank = 0.f;
  if (nf > 0) {
    const size_t o = (bt0 + nf - 1) * C;
    if (owner) { n_alpha = p.alphas[o + prow]; n_blank = ldg_stream(p.blank + o + prow); }
    if (tail_owner) { n_talpha = p.alphas[o + V]; n_tblank = ldg_stream(p.blank + o + V); }
  }

  for (int it = 0; it < nf; ++it) {
    const int t = nf - 1 - it;
    const int stage = it % NS;
    const uint32_t parity = (it / NS) & 1;
    float* beta = beta_buf + (it & 1) * BP;          // beta_{t+1}; entry q at beta[3 + q]
    float* nxt = beta_buf + ((it + 1) & 1) * BP;
    if (it > 0) {
      // every row of the previous frame has been reduced cluster-wide: beta is
      // complete and the tile stage of iteration it-1 is free for the next TMA
      mbar_wait(smem_u32(&xbar[it & 1]), ((it - 1) >> 1) & 1);
      if (tid == 0 && it - 1 + NS < nf) issue(it - 1 + NS);
    }
    if (tid == 0) mbar_arrive_expect_tx(smem_u32(&xbar[(it + 1) & 1]), C * 4);
    const float c_alpha = n_alpha, c_blank = n_blank, c_talpha = n_talpha, c_tblank = n_tblank;
    if (t > 0) {
      const size_t o = (bt0 + t - 1) * C;
      if (owner) { n_alpha = p.alphas[o + prow]; n_blank = ldg_stream(p.blank + o + prow); }
      if (tail_owner) { n_talpha = p.alphas[o + V]; n_tblank = ldg_stream(p.blank + o + V); }
    }
    mbar_wait(smem_u32(&bars[stage]), parity);
    const float* tile = tiles + (size_t)stage * (kStageBytes / 4);
    const float* bnext = beta + 4;                     // bnext[y] = beta'[1 + y]
    float* gl = p.grad_lexical + (bt0 + t) * (size_t)C * V;
    float* gb = p.grad_blank + (bt0 + t) * C;

    {
      const float* trow = tile + (size_t)row * V;
      float4 x[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int c4 = (sl + 8 * i) * 4;
        const float4 w = *reinterpret_cast<const float4*>(trow + c4);
        const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
        x[i] = make_float4(arc<SR>(w.x, bn.x), arc<SR>(w.y, bn.y), arc<SR>(w.z, bn.z),
                           arc<SR>(w.w, bn.w));
      }
      const float alpha_p = to_dom<SR>(__shfl_sync(0xffffffffu, c_alpha, lane & ~7));
      float* grow = gl + (size_t)prow * V;
      float rowsum;
      if constexpr (SR == LT_LOG) {
        float m = neg_inf();
#pragma unroll
        for (int i = 0; i < CH; ++i) m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        const float ms = msafe(m);
        const float rs = scale_ok ? gscale * ex2(alpha_p + ms - logz2) : 0.f;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          float4 e;
          e.x = ex2(x[i].x - ms); e.y = ex2(x[i].y - ms);
          e.z = ex2(x[i].z - ms); e.w = ex2(x[i].w - ms);
          s += (e.x + e.y) + (e.z + e.w);
          stg_stream4(grow + (sl + 8 * i) * 4, make_float4(e.x * rs, e.y * rs, e.z * rs, e.w * rs));
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        rowsum = ms + __log2f(s);
      } else {
        float s = 0.f;
        const float ga = gscale * alpha_p;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int c4 = (sl + 8 * i) * 4;
          const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
          s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
          stg_stream4(grow + c4, make_float4(ga * bn.x, ga * bn.y, ga * bn.z, ga * bn.w));
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        rowsum = s;
      }
      if (owner) {
        const float bp = beta[3 + prow];
        const float bb = arc<SR>(c_blank, bp);
        if constexpr (SR == LT_LOG) gb[prow] = scale_ok ? gscale * ex2(alpha_p + bb - logz2) : 0.f;
        else gb[prow] = gscale * c_alpha * bp;
        xchg_store(nxt, 3 + prow, SR == LT_LOG ? log2_add_exp2(bb, rowsum) : bb + rowsum,
                   &xbar[(it + 1) & 1], nrank);
      }
    }
    if (last_rank && warp == 0) {          // tail row: source state V, all 32 lanes
      const float* trow = tile + (size_t)kRows * V;
      const float alpha_p = to_dom<SR>(__shfl_sync(0xffffffffu, c_talpha, 0));
      float* grow = gl + (size_t)V * V;
      float rowsum;
      if constexpr (SR == LT_LOG) {
        float m = neg_inf();
        for (int c4 = lane * 4; c4 < V; c4 += 128) {
          const float4 w = *reinterpret_cast<const float4*>(trow + c4);
          const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
          m = fmaxf(m, fmaxf(fmaxf(arc<SR>(w.x, bn.x), arc<SR>(w.y, bn.y)),
                             fmaxf(arc<SR>(w.z, bn.z), arc<SR>(w.w, bn.w))));
        }
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float ms = msafe(m);
        const float rs = scale_ok ? gscale * ex2(alpha_p + ms - logz2) : 0.f;
        float s = 0.f;
        for (int c4 = lane * 4; c4 < V; c4 += 128) {
          const float4 w = *reinterpret_cast<const float4*>(trow + c4);
          const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
          float4 e;
          e.x = ex2(arc<SR>(w.x, bn.x) - ms); e.y = ex2(arc<SR>(w.y, bn.y) - ms);
          e.z = ex2(arc<SR>(w.z, bn.z) - ms); e.w = ex2(arc<SR>(w.w, bn.w) - ms);
          s += (e.x + e.y) + (e.z + e.w);
          stg_stream4(grow + c4, make_float4(e.x * rs, e.y * rs, e.z * rs, e.w * rs));
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        rowsum = ms + __log2f(s);
      } else {
        float s = 0.f;
        const float ga = gscale * alpha_p;
        for (int c4 = lane * 4; c4 < V; c4 += 128) {
          const float4 w = *reinterpret_cast<const float4*>(trow + c4);
          const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
          s += (w.x * bn.x + w.y * bn.y) + (w.z * bn.z + w.w * bn.w);
          stg_stream4(grow + c4, make_float4(ga * bn.x, ga * bn.y, ga * bn.z, ga * bn.w));
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        rowsum = s;
      }
      if (lane == 0) {
        const float bp = beta[3 + V];
        const float bb = arc<SR>(c_tblank, bp);
        if constexpr (SR == LT_LOG) gb[V] = scale_ok ? gscale * ex2(alpha_p + bb - logz2) : 0.f;
        else gb[V] = gscale * c_talpha * bp;
        xchg_store(nxt, 3 + V, SR == LT_LOG ? log2_add_exp2(bb, rowsum) : bb + rowsum,
                   &xbar[(it + 1) & 1], nrank);
      }
    }
  }
  float* beta = beta_buf + (nf & 1) * BP;
  if (nf > 0) mbar_wait(smem_u32(&xbar[nf & 1]), ((nf - 1) >> 1) & 1);
  if (p.beta_final) {
    if (tid < kRows)
      p.beta_final[(size_t)b * C + rank * kRows + tid] = from_dom<SR>(beta[3 + rank * kRows + tid]);
    if (last_rank && tid == 0) p.beta_final[(size_t)b * C + V] = from_dom<SR>(beta[3 + V]);
  }
  cluster_sync_all();
}

// ------------------------------------------------------------------- host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
          cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

template <typename KernelT, typename... Args>
static int launch_fast(KernelT kernel, int grid, size_t smem, int cluster, cudaStream_t stream,
                       Args... args) {
  LT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LT_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  note_launch();
  return LT_OK;
}

constexpr size_t kSmemBudget = 227 * 1024;

}  // namespace

bool lattice_fast_supported(const NGram& g, int k, unsigned flags, const void* lexical) {
  if (flags & LT_FLAG_FORCE_GENERIC) return false;
  if ((flags >> LT_FLAG_CLUSTER_SHIFT) & 0xf) return false;   // explicit cluster size => generic
  if (k >= 1 || g.n != 1) return false;
  if (g.V % 64 != 0 || g.V > 256) return false;
  if (reinterpret_cast<uintptr_t>(lexical) % 16 != 0) return false;
  return true;
}

int lattice_forward_fast_launch(int semiring, const NGram& g, const FwdParams& base,
                                cudaStream_t stream) {
  const int V = g.V, C = g.C, VD = V / 64;
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled is unavailable in this driver"); return LT_ERR_CUDA; }
  CUtensorMap tmap;
  const cuuint64_t rows = (cuuint64_t)base.B * base.T * C;
  cuuint64_t dims[2] = {(cuuint64_t)V, rows};
  cuuint64_t strides[1] = {(cuuint64_t)V * 4};
  cuuint32_t box[2] = {(cuuint32_t)kColsPerCta, (cuuint32_t)V};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base.lexical),
                      dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return LT_ERR_CUDA; }
  const size_t stage = (size_t)V * kColsPerCta * 4;
  const size_t fixed = sizeof(float) * (2 * ((C + 3) & ~3) + 2 * kWarps * kColsPerCta) + 8 * 16 + 256;
  int stages = (int)((kSmemBudget - fixed) / stage);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("fast forward: not enough shared memory"); return LT_ERR_UNSUPPORTED; }
  const size_t smem = stage * stages + fixed;
  FastFwdParams p = {};
  p.V = V; p.B = base.B; p.T = base.T; p.stages = stages;
  p.blank = base.blank; p.lexical = base.lexical; p.num_frames = base.num_frames;
  p.alpha_init = base.alpha_init; p.dist = base.dist; p.alphas = base.alphas;
  p.alpha_final = base.alpha_final; p.backptr = base.backptr;
  const int grid = base.B * VD;
#define LT_FWD(SR)                                                                             \
  switch (VD) {                                                                                \
    case 1: return launch_fast(lattice_forward_fast<SR, 1>, grid, smem, 1, stream, tmap, p);   \
    case 2: return launch_fast(lattice_forward_fast<SR, 2>, grid, smem, 2, stream, tmap, p);   \
    case 3: return launch_fast(lattice_forward_fast<SR, 3>, grid, smem, 3, stream, tmap, p);   \
    default: return launch_fast(lattice_forward_fast<SR, 4>, grid, smem, 4, stream, tmap, p);  \
  }
  if (semiring == LT_LOG) { LT_FWD(LT_LOG) }
  if (semiring == LT_MAXTROPICAL) { LT_FWD(LT_MAXTROPICAL) }
  LT_FWD(LT_REAL)
#undef LT_FWD
}

int lattice_backward_fast_launch(int semiring, const NGram& g, const BwdParams& base,
                                 cudaStream_t stream) {
  const int V = g.V, C = g.C, VD = V / 64;
  const size_t stage = (size_t)64 * V * 4 + (size_t)V * 4;
  const size_t fixed = sizeof(float) * 2 * ((((C + 6) & ~3) + 4)) + 8 * 16 + 256;
  int stages = (int)((kSmemBudget - fixed) / stage);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("fast backward: not enough shared memory"); return LT_ERR_UNSUPPORTED; }
  const size_t smem = stage * stages + fixed;
  FastBwdParams p = {};
  p.V = V; p.B = base.B; p.T = base.T; p.stages = stages;
  p.blank = base.blank; p.lexical = base.lexical; p.num_frames = base.num_frames;
  p.alphas = base.alphas; p.dist = base.dist; p.grad_dist = base.grad_dist;
  p.grad_blank = base.grad_blank; p.grad_lexical = base.grad_lexical; p.beta_final = base.beta_final;
  const int grid = base.B * VD;
#define LT_BWD(SR)                                                                         \
  switch (VD) {                                                                            \
    case 1: return launch_fast(lattice_backward_fast<SR, 1>, grid, smem, 1, stream, p);    \
    case 2: return launch_fast(lattice_backward_fast<SR, 2>, grid, smem, 2, stream, p);    \
    case 3: return launch_fast(lattice_backward_fast<SR, 3>, grid, smem, 3, stream, p);    \
    default: return launch_fast(lattice_backward_fast<SR, 4>, grid, smem, 4, stream, p);   \
  }
  if (semiring == LT_LOG) { LT_BWD(LT_LOG) }
  LT_BWD(LT_REAL)
#undef LT_BWD
}

}  // namespace lt
