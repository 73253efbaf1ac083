// K1 for higher-order contexts (FullNGram context_size >= 2, e.g. BASELINE.json
// configs[2]: vocab 64, 4161 states) and both alignment lattices: "one thread per
// destination column".
//
// FullNGram.forward_reduce (contexts.py:207-230) on a frame's [C, V] weights is, for the
// N = V^n full-order destinations, a COLUMN reduction of the row-major [K = V+1, N] matrix
// that starts Alow*V floats into the frame (see NGram in common.cuh); element (kk, j) comes
// from source state Alow + j/V + kk*(N/V).  Here
//   * a cluster of CL CTAs owns one utterance, CTA r owns NCOL = N/CL columns and every
//     consumer thread owns CPT = NCOL/256 adjacent columns, so the (+)-reduction over the K
//     rows is a serial loop in registers: no shuffles, no partials, no block barrier;
//     a warp covers <= V adjacent columns, so the source value of a row is one broadcast
//     shared-memory read per warp;
//   * the matrix is streamed by a dedicated producer warp with 3-D TMA boxes
//     [256 cols x R rows x 1 frame] into a ring (full / empty mbarriers), several chunks
//     ahead of the recursion; FrameLabelDependent(k) streams the frame k times (the
//     re-reads hit the 126 MB L2) because every level needs the previous level complete;
//   * alpha (and the FrameLabelDependent level vectors) live in shared memory as a full
//     replica per CTA in two buffers that alternate per level; new values are all-gathered
//     with st.async + mbarrier complete_tx (no cluster barrier in the loop);
//   * two CTAs per SM (<= 113 KB of shared memory, 288 threads each), so the exchange
//     latency of one utterance hides behind the arithmetic of another.
// The A = sum_{i<n} V^i low-order states have at most one incoming arc each and are
// handled by spare lanes of rank 0.
//
// Renormalised state (Log, `alpha_norm` given; same contract as lattice_fast2.cu):
// alpha_t = alpha~_t + off_t with exact integer offsets in log2 units, off_{t+1} = off_t + D_t,
// the shift applied where alpha_{t+1} is published.  No CTA holds all of alpha, so the maximum
// the shift follows is exchanged: at the first level of a frame every warp sends the maximum of
// the alpha~_t entries it owns to every CTA with the level's own st.async exchange (64 floats per
// CTA, nothing on the dependency chain) and every warp reduces the 8 CL values after the next
// wait.  FrameLabelDependent has the maximum of alpha~_t in hand before the frame's last level:
// D_t = floor(max alpha~_t).  FrameDependent (one level per frame) sees it one frame late and
// predicts: the shift cancels the growth last observed in full and half of the level it expects
// (see the loop), which keeps max |alpha~| within about one frame of growth; plain feedback of the
// late maximum would ring, half of it settles at two frames of growth.
//
// Reference semantics: lattices.py:436-462, alignments.py:294-297 (FrameDependent),
// alignments.py:362-376 (FrameLabelDependent), contexts.py:207-230; MaxTropical ties:
// semirings.py:363 (blank >= lexical), :382 (first arg-max = lowest row block / fewest
// expansions).
#include <cuda.h>

#include "common.cuh"
#include "fast_ptx.cuh"
#include "params.cuh"

namespace lt {

namespace {

using namespace fastptx;

constexpr int kConsumers = 256;
constexpr int kColsThreads = kConsumers + 32;   // + one producer warp
constexpr int kBoxCols = 256;                   // TMA box width (boxDim <= 256)
constexpr int kMaxR = 16;                       // rows per ring stage (<=)
constexpr int kLowPerThread = 2;                // low-order destinations per rank-0 thread
constexpr int kSrcStride = 32;                  // floats per row block of the source buffer (>= W)

__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap* map, int c0,
                                                 int c1, int c2, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// On-chip number domain: the Log semiring is carried in LOG2 units (one bare MUFU.EX2 per arc,
// w * log2(e) folded into an FFMA), Real / MaxTropical are carried as they are.
template <int SR> struct Dom {
  __device__ static float in(float x) { return SR == LT_LOG ? x * kLog2e : x; }
  __device__ static float out(float x) { return SR == LT_LOG ? x * kLn2 : x; }
  // on-chip value (x) natural-unit arc weight
  __device__ static float times(float a, float w) {
    if constexpr (SR == LT_LOG) return fmaf(w, kLog2e, a);
    else if constexpr (SR == LT_MAXTROPICAL) return a + w;
    else return a * w;
  }
  __device__ static float plus(float a, float b) {
    if constexpr (SR == LT_LOG) return log2_add_exp2(a, b);
    else return Sr<SR>::plus(a, b);
  }
};
// (+)-accumulator in the on-chip domain: Acc<SR> for Real / MaxTropical, a log2-domain
// (max, sum) pair for Log.
template <int SR> struct DAcc : Acc<SR> {};
template <> struct DAcc<LT_LOG> {
  float m, s;
  __device__ void init() { m = neg_inf(); s = 0.f; }
  __device__ void add(float x, int) {
    const float mn = fmaxf(m, x);
    const float ms = msafe(mn);
    const float sc = (m == neg_inf()) ? 0.f : ex2(msafe(m) - ms);
    s = s * sc + ex2(x - ms);
    m = mn;
  }
  template <int N> __device__ void add_chunk(const float (&x)[N], float cm) {
    const float mn = fmaxf(m, cm);
    const float ms = msafe(mn);
    const float sc = (m == neg_inf()) ? 0.f : ex2(msafe(m) - ms);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) acc += ex2(x[i] - ms);
    s = s * sc + acc;
    m = mn;
  }
  __device__ void merge(const DAcc& o) { lse2_merge(m, s, o.m, o.s); }
  __device__ float value() const { return msafe(m) + __log2f(s); }
  __device__ int arg() const { return 0; }
};

struct ColsParams {
  NGram g;
  int k;            // max_expansions or -1
  int B, T;
  int R, nchunks, stages, cl, ncol, sbuf;
  const float* blank;
  const float* lexical;
  const int32_t* num_frames;
  const float* alpha_init;
  float* dist;
  float* alphas;
  float* alpha_final;
  float* levels;
  int16_t* backptr;
  uint8_t* termptr;
  int32_t* alpha_norm;   // Log only: [B, T+3] offsets of the renormalised state (fast2.cuh), or nullptr
};

// Epilogue of one destination for one level (shared by column and low-order destinations).
// `r`, `arg`: the (+)-reduction over the incoming lexical arcs and its arg-max row block.
// Returns the value to publish into the next state buffer.
template <int SR, bool FLD>
__device__ __forceinline__ float finish_dest(const ColsParams& p, size_t bt, int q, int level,
                                             float r, int arg, float src_q, float blank_q,
                                             DAcc<SR>& term) {
  using S = Sr<SR>;
  const int C = p.g.C;
  if constexpr (!FLD) {
    const float a = Dom<SR>::times(src_q, blank_q);
    if constexpr (SR == LT_MAXTROPICAL) {
      const bool take_blank = a >= r;                        // semirings.py:363
      if (p.backptr) p.backptr[bt * C + q] = take_blank ? (int16_t)-1 : (int16_t)arg;
      return take_blank ? a : r;
    } else {
      return Dom<SR>::plus(a, r);
    }
  } else {
    if (level == 0) { term.init(); term.add(Dom<SR>::times(src_q, blank_q), 0); }   // term_0 = alpha (x) blank
    if (p.levels) p.levels[(bt * p.k + level) * C + q] = Dom<SR>::out(r);
    if constexpr (SR == LT_MAXTROPICAL) {
      if (p.backptr) p.backptr[(bt * p.k + level) * C + q] = (int16_t)arg;
    }
    term.add(Dom<SR>::times(r, blank_q), level + 1);         // strict '>' keeps fewer expansions
    if (level + 1 < p.k) return r;                           // last_{level+1}
    if constexpr (SR == LT_MAXTROPICAL) {
      if (p.termptr) p.termptr[bt * C + q] = (uint8_t)term.arg();
    }
    return term.value();                                     // alpha_{t+1}
  }
}

template <int SR, int CPT, bool FLD, bool NORM>
__global__ void __launch_bounds__(kColsThreads, 2)
lattice_forward_cols(const __grid_constant__ CUtensorMap tmap, const ColsParams p) {
  using S = Sr<SR>;
  extern __shared__ __align__(128) unsigned char csmem[];
  const NGram& g = p.g;
  const int C = g.C, V = g.V, K = g.K, A = g.A, Alow = g.Alow;
  const int NS = p.stages, R = p.R, nchunks = p.nchunks;
  const int NCOL = p.ncol;
  const int W = NCOL / V;                 // source groups (runs of V columns) per CTA
  const int SB = p.sbuf;                  // floats per state buffer: K rows of kSrcStride tail
                                          // sources (W used) + Alow low ones
  const int LOW0 = K * kSrcStride;        // first low-order slot
  const uint32_t stage_bytes = (uint32_t)R * NCOL * 4;
  const int nlev = FLD ? p.k : 1;

  float* tiles = reinterpret_cast<float*>(csmem);
  float* buf = reinterpret_cast<float*>(csmem + (size_t)NS * stage_bytes);    // [2][SB]
  float* dpart = buf + 2 * SB;                                                // [2 * 8] dist partials
  float* wmx = dpart + 16;                                                    // [2][8 ranks][8 warps] maxima
  uint64_t* full = reinterpret_cast<uint64_t*>(wmx + 128);
  uint64_t* empty = full + NS;
  uint64_t* xbar = empty + NS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t CL = p.cl;
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / CL;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;

  auto init_value = [&](int st) -> float {
    return Dom<SR>::in(p.alpha_init ? p.alpha_init[(size_t)b * C + st]
                                    : (st == 0 ? S::one() : S::zero()));
  };

  if (tid == 0) {
    prefetch_tensormap(&tmap);
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), kConsumers / 32);
    }
    mbar_init(smem_u32(&xbar[0]), 1);
    mbar_init(smem_u32(&xbar[1]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  // The sources this CTA reads: state Alow + (rank*W + i) + kk*pstride sits at slot
  // kk*kSrcStride + i; rank 0 also keeps the states below Alow (sources of the single-arc
  // destinations).
  for (int idx = tid; idx < SB; idx += kColsThreads) {
    float v = S::zero();
    if (idx < LOW0) {
      if (idx % kSrcStride < W)
        v = init_value(Alow + (int)rank * W + idx % kSrcStride + (idx / kSrcStride) * g.pstride);
    } else if (rank == 0 && idx - LOW0 < Alow) {
      v = init_value(idx - LOW0);
    }
    buf[idx] = v;
    buf[SB + idx] = S::zero();
  }
  __syncthreads();
  cluster_sync_all();

  const long long total_levels = (long long)nf * nlev;
  static_assert(!NORM || SR == LT_LOG, "renormalisation is a Log-semiring feature");
  constexpr bool norm = NORM;             // compile-time: the plain kernels carry none of it
  int32_t* an = norm ? p.alpha_norm + (size_t)b * (p.T + 3) : nullptr;
  int off = 0;                            // off_t (consumers)

  if (warp == kConsumers / 32) {
    // ------------------------------------------------------------ producer warp
    if (lane == 0) {
      const long long total = total_levels * nchunks;
      const int nbox = NCOL / kBoxCols;
      // FrameLabelDependent streams a frame once per level: keep it in L2 until the last one
      const uint64_t keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
      int stage = 0;
      uint32_t use = 0;            // how many times the ring has wrapped
      for (long long gi = 0; gi < total; ++gi) {
        const long long lev = gi / nchunks;
        const int c = (int)(gi - lev * nchunks);
        const int t = (int)(lev / nlev);
        if (use > 0) mbar_wait(smem_u32(&empty[stage]), (use - 1) & 1);
        const uint32_t bar = smem_u32(&full[stage]);
        mbar_arrive_expect_tx(bar, stage_bytes);
        const int level = (int)(lev - (long long)t * nlev);
        const uint64_t pol = (level + 1 < nlev) ? keep : drop;
        for (int bx = 0; bx < nbox; ++bx)
          tma_load_3d_hint(smem_u32(tiles) + stage * stage_bytes + bx * (R * kBoxCols * 4), &tmap,
                           (int)rank * NCOL + bx * kBoxCols, c * R, (int)(bt0 + t), bar, pol);
        if (++stage == NS) { stage = 0; ++use; }
      }
    }
  } else {
    // ---------------------------------------------------------------- consumers
    // destinations of this thread: columns jc0 .. jc0+CPT-1 of the [K, N] matrix (states
    // A + jc) and, on rank 0, up to kLowPerThread of the A low-order states.
    constexpr int ND = CPT + kLowPerThread;
    const int jl = tid * CPT;                       // column inside the CTA slice
    const int jc0 = (int)rank * NCOL + jl;
    const int box_off = (jl / kBoxCols) * (R * kBoxCols) + (jl % kBoxCols);
    const int li = jl / V;                          // source group of this thread's columns
                                                    // (CPT divides V: they share it)
    int dq[ND];                                     // destination state, -1 = none
    uint32_t raddr[ND], rbar[ND];                   // where its value goes (buffer 0 / xbar[0])
    float areg[ND];                                 // alpha_t[q], kept by the owner thread
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      int q;
      if (d < CPT) q = A + jc0 + d;
      else { q = tid + (d - CPT) * kConsumers; if (rank != 0 || q >= A) q = -1; }
      dq[d] = q;
      raddr[d] = 0; rbar[d] = 0; areg[d] = S::zero();
      if (q >= 0) {
        areg[d] = init_value(q);
        // the one CTA that reads state q as a source, and the slot it expects it in
        int rdst = 0, slot = LOW0 + q;
        if (q >= Alow) {
          const int pp = q - Alow, ig = pp % g.pstride, kk = pp / g.pstride;
          rdst = ig / W;
          slot = kk * kSrcStride + (ig - rdst * W);
        }
        raddr[d] = map_shared_rank(smem_u32(buf + slot), rdst);
        rbar[d] = map_shared_rank(smem_u32(&xbar[0]), rdst);
      }
    }
    const uint32_t expect = (uint32_t)(W * K + (rank == 0 ? Alow : 0)) * 4;   // bytes per level
    DAcc<SR> term[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) term[d].init();

    // per-frame operands that do not depend on alpha, prefetched one frame ahead
    float nbl[ND], nlx[kLowPerThread];
    auto prefetch = [&](int t) {
      const float* bl = p.blank + (bt0 + t) * C;
      const float* lx = p.lexical + (bt0 + t) * (size_t)C * V;
#pragma unroll
      for (int d = 0; d < ND; ++d) nbl[d] = dq[d] >= 0 ? ldg_stream(bl + dq[d]) : 0.f;
#pragma unroll
      for (int i = 0; i < kLowPerThread; ++i)
        nlx[i] = dq[CPT + i] >= g.off ? ldg_stream(lx + (dq[CPT + i] - g.off)) : 0.f;
    };
    if (nf > 0) prefetch(0);
    float cbl[ND], clx[kLowPerThread];

    int stage = 0;
    uint32_t use = 0;
    float pend = 0.f;                     // D_t, applied where alpha_{t+1} is published
    [[maybe_unused]] float m_prev = 0.f, d_prev = 0.f, d_prev2 = 0.f;   // M_{t-2}, D_{t-1}, D_{t-2} (FrameDependent)
    for (long long lev = 0; lev < total_levels; ++lev) {
      const int t = (int)(lev / nlev);
      const int level = (int)(lev - (long long)t * nlev);
      const size_t bt = bt0 + t;
      const float* src = buf + (lev & 1) * SB;
      const uint32_t dpar = (uint32_t)((lev + 1) & 1);
      if (lev > 0) mbar_wait(smem_u32(&xbar[lev & 1]), (uint32_t)(((lev - 1) >> 1) & 1));
      if (tid == 0)
        mbar_arrive_expect_tx(smem_u32(&xbar[dpar]), expect + ((norm && level == 0) ? CL * 32u : 0u));
      if (norm) {
        if (lev > 0 && (nlev == 1 || level == 1)) {
          // the warp maxima of alpha~ sent during the previous (first) level have landed
          const float* wm = wmx + (lev & 1) * 64;
          float m = lane < (int)CL * 8 ? wm[lane] : neg_inf();
          if (lane + 32 < (int)CL * 8) m = fmaxf(m, wm[lane + 32]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          if constexpr (FLD) {
            pend = norm_shift(m);                       // m = max alpha~_t: no lag (k >= 2), or ...
            if (nlev == 1) pend = norm_shift(0.5f * m); // ... k = 1: one frame late, damped
          } else {
            // m = M_{t-1}, one frame late.  M_{t+1} = M_t + G_t - D_t with G the growth of the
            // maximum; predict M_t = M_{t-1} + G^ - D_{t-1} and G_t = G^ from the last observed
            // growth G^ = M_{t-1} - M_{t-2} + D_{t-2}, cancel the growth fully and the predicted
            // level by half (the half damps the prediction error instead of echoing it)
            float raw = 0.5f * m;
            if (t >= 3 && is_finite(m) && is_finite(m_prev)) {
              const float ghat = m - m_prev + d_prev2;
              raw = fmaf(0.5f, m - d_prev + ghat, ghat);
            }
            pend = norm_shift(raw);
            m_prev = m; d_prev2 = d_prev; d_prev = pend;
          }
        }
        if (level == 0) {
          if (rank == 0 && tid == 0) an[t] = off;
          float m = neg_inf();
#pragma unroll
          for (int d = 0; d < ND; ++d)
            if (dq[d] >= 0) m = fmaxf(m, areg[d]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          if (lane < (int)CL)
            st_async_f32(map_shared_rank(smem_u32(wmx + dpar * 64 + rank * 8 + warp), lane), m,
                         map_shared_rank(smem_u32(&xbar[dpar]), lane));
        }
      }
      if (level == 0) {
#pragma unroll
        for (int d = 0; d < ND; ++d) cbl[d] = nbl[d];
#pragma unroll
        for (int i = 0; i < kLowPerThread; ++i) clx[i] = nlx[i];
        if (t + 1 < nf) prefetch(t + 1);
        if (p.alphas) {
#pragma unroll
          for (int d = 0; d < ND; ++d)
            if (dq[d] >= 0) p.alphas[bt * C + dq[d]] = Dom<SR>::out(areg[d]);
        }
      }

      // ---- column reduction over the K rows, chunk by chunk
      DAcc<SR> acc[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) acc[c].init();
      const float* srow = src + li;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(smem_u32(&full[stage]), use & 1);
        const float* tile = tiles + (size_t)stage * (stage_bytes / 4) + box_off;
        const int kk0 = ch * R;
        const int rows = min(R, K - kk0);
        const float* sp = srow + kk0 * kSrcStride;
        auto load_w = [&](int r, float (&w)[CPT]) {
          if constexpr (CPT == 2) {
            const float2 v2 = *reinterpret_cast<const float2*>(tile + r * kBoxCols);
            w[0] = v2.x; w[1] = v2.y;
          } else {
            w[0] = tile[r * kBoxCols];
          }
        };
        if constexpr (SR == LT_LOG) {
          float x[CPT][kMaxR];
          float cm[CPT];
#pragma unroll
          for (int c = 0; c < CPT; ++c) cm[c] = neg_inf();
#pragma unroll
          for (int r = 0; r < kMaxR; ++r) {
            if (r < rows) {
              float w[CPT];
              load_w(r, w);
              const float sv = sp[r * kSrcStride];
#pragma unroll
              for (int c = 0; c < CPT; ++c) {
                x[c][r] = fmaf(w[c], kLog2e, sv);
                cm[c] = fmaxf(cm[c], x[c][r]);
              }
            } else {
#pragma unroll
              for (int c = 0; c < CPT; ++c) x[c][r] = neg_inf();
            }
          }
#pragma unroll
          for (int c = 0; c < CPT; ++c) acc[c].add_chunk(x[c], cm[c]);
        } else if constexpr (SR == LT_MAXTROPICAL) {
          // (max, first arg-max row): the row index inside the chunk is an immediate of the
          // unrolled loop (one select per candidate, no index arithmetic); it is rebased to the
          // frame's row numbering once per chunk.  Rows ascend, strict '>' keeps the first.
          float bm[CPT];
          int br[CPT];
#pragma unroll
          for (int c = 0; c < CPT; ++c) { bm[c] = acc[c].m; br[c] = -1; }
#pragma unroll
          for (int r = 0; r < kMaxR; ++r) {
            if (r < rows) {
              float w[CPT];
              load_w(r, w);
              const float sv = sp[r * kSrcStride];
#pragma unroll
              for (int c = 0; c < CPT; ++c) {
                const float v = sv + w[c];
                const bool better = v > bm[c];
                bm[c] = better ? v : bm[c];
                br[c] = better ? r : br[c];
              }
            }
          }
#pragma unroll
          for (int c = 0; c < CPT; ++c)
            if (br[c] >= 0) { acc[c].m = bm[c]; acc[c].a = kk0 + br[c]; }
        } else {
          int r = 0;
          for (; r + 4 <= rows; r += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float w[CPT];
              load_w(r + u, w);
              const float sv = sp[(r + u) * kSrcStride];
#pragma unroll
              for (int c = 0; c < CPT; ++c) acc[c].add(Dom<SR>::times(sv, w[c]), kk0 + r + u);
            }
          }
          for (; r < rows; ++r) {
            float w[CPT];
            load_w(r, w);
            const float sv = sp[r * kSrcStride];
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[c].add(Dom<SR>::times(sv, w[c]), kk0 + r);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[stage]));
        if (++stage == NS) { stage = 0; ++use; }
      }

      // ---- epilogue: every new value goes to the one CTA that reads it as a source
      const uint32_t boff = dpar * (uint32_t)SB * 4, xoff = dpar * 8;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const int q = dq[d];
        if (q < 0) continue;
        float r;
        int arg = 0;
        if (d < CPT) {
          r = acc[d < CPT ? d : 0].value();
          arg = acc[d < CPT ? d : 0].arg();
        } else {
          r = S::zero();                                                   // state 0: no incoming arc
          if (q >= g.off) r = Dom<SR>::times(src[LOW0 + (q - g.off) / V], clx[d >= CPT ? d - CPT : 0]);
        }
        float v = finish_dest<SR, FLD>(p, bt, q, level, r, arg, areg[d], cbl[d], term[d]);
        if (norm && level + 1 == nlev) v -= pend;           // alpha~_{t+1}: -inf stays -inf
        if (!FLD || level + 1 == nlev) areg[d] = v;
        st_async_f32(raddr[d] + boff, v, rbar[d] + xoff);
      }
      if (norm && level + 1 == nlev) off += (int)pend;
    }

    if (total_levels > 0)
      mbar_wait(smem_u32(&xbar[total_levels & 1]), (uint32_t)(((total_levels - 1) >> 1) & 1));

    // padding frames keep alpha (lattices.py:460-461) and are still recorded (:462);
    // dist = (+)_c alpha_T[c] (lattices.py:496): per-thread, per-warp, per-CTA, per-cluster
    DAcc<SR> part; part.init();
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const int q = dq[d];
      if (q < 0) continue;
      if (p.alphas)
        for (int t = nf; t < p.T; ++t) p.alphas[(bt0 + t) * C + q] = Dom<SR>::out(areg[d]);
      if (p.alpha_final)
        p.alpha_final[(size_t)b * C + q] =
            norm ? (float)(((double)areg[d] + (double)off) * 0.6931471805599453) : Dom<SR>::out(areg[d]);
      DAcc<SR> one; one.init(); one.add(areg[d], q);
      part.merge(one);
    }
    for (int o = 16; o > 0; o >>= 1) {
      DAcc<SR> other = part;
      if constexpr (SR == LT_LOG) {
        other.m = __shfl_xor_sync(0xffffffffu, part.m, o);
        other.s = __shfl_xor_sync(0xffffffffu, part.s, o);
      } else if constexpr (SR == LT_MAXTROPICAL) {
        other.m = __shfl_xor_sync(0xffffffffu, part.m, o);
        other.a = __shfl_xor_sync(0xffffffffu, part.a, o);
      } else {
        other.s = __shfl_xor_sync(0xffffffffu, part.s, o);
      }
      part.merge(other);
    }
    if (norm && rank == 0)
      for (int t = nf + tid; t <= p.T; t += kConsumers) an[t] = off;
    // warp partials -> rank 0's dpart[rank * ... ] is too small for 8 warps x 8 ranks; reduce
    // inside the CTA through the (now idle) state buffer first
    float* red = buf;                    // nobody reads the state buffers any more
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
    if (lane == 0) {
      if constexpr (SR == LT_LOG) { red[2 * warp] = part.m; red[2 * warp + 1] = part.s; }
      else if constexpr (SR == LT_MAXTROPICAL) { red[2 * warp] = part.m; red[2 * warp + 1] = __int_as_float(part.a); }
      else { red[2 * warp] = part.s; red[2 * warp + 1] = 0.f; }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
    if (tid == 0) {
      for (int w = 1; w < kConsumers / 32; ++w) {
        DAcc<SR> other = part;
        if constexpr (SR == LT_LOG) { other.m = red[2 * w]; other.s = red[2 * w + 1]; }
        else if constexpr (SR == LT_MAXTROPICAL) { other.m = red[2 * w]; other.a = __float_as_int(red[2 * w + 1]); }
        else { other.s = red[2 * w]; }
        part.merge(other);
      }
      float v0, v1;
      if constexpr (SR == LT_LOG) { v0 = part.m; v1 = part.s; }
      else if constexpr (SR == LT_MAXTROPICAL) { v0 = part.m; v1 = 0.f; }
      else { v0 = part.s; v1 = 0.f; }
      st_shared_cluster_f32(map_shared_rank(smem_u32(dpart + 2 * rank), 0), v0);
      st_shared_cluster_f32(map_shared_rank(smem_u32(dpart + 2 * rank + 1), 0), v1);
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (rank == 0 && tid == 0) {
    DAcc<SR> tot; tot.init();
    for (uint32_t r = 0; r < CL; ++r) {
      DAcc<SR> other = tot;
      if constexpr (SR == LT_LOG) { other.m = dpart[2 * r]; other.s = dpart[2 * r + 1]; }
      else if constexpr (SR == LT_MAXTROPICAL) { other.m = dpart[2 * r]; other.a = 0; }
      else { other.s = dpart[2 * r]; }
      tot.merge(other);
    }
    if (norm) {
      // logZ = (off_T + r) ln 2, rounded once from double; the backward reads the pair
      const float r = tot.value();
      an[p.T + 1] = __float_as_int(r);
      an[p.T + 2] = 0;                   // offsets are in log2 units
      p.dist[b] = (float)(((double)r + (double)off) * 0.6931471805599453);
    } else {
      p.dist[b] = Dom<SR>::out(tot.value());
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn3() {
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

template <typename KernelT>
static int launch_cols(KernelT kernel, int grid, size_t smem, int cluster, cudaStream_t stream,
                       const CUtensorMap& tmap, const ColsParams& p) {
  LT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kColsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LT_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmap, p));
  note_launch();
  return LT_OK;
}

// cluster size / columns per thread for a geometry, or false if the path does not apply
static size_t cols_fixed_bytes(const NGram& g, int ncol) {
  const int sbuf = (g.K * kSrcStride + g.Alow + 3) & ~3;
  return sizeof(float) * (2 * (size_t)sbuf + 16 + 128) + 8 * (2 * 8 + 2) + 256;
}

static bool cols_geometry(const NGram& g, int* cl, int* cpt) {
  if (g.n < 2) return false;
  if (g.V % 4 != 0) return false;                       // 16-byte TMA strides and base
  if (g.A > kLowPerThread * kConsumers) return false;
  if (g.N % kBoxCols != 0) return false;
  // prefer the widest split (more SMs per utterance) that still gives every thread a column
  for (int c = 8; c >= 1; c >>= 1) {
    if (g.N % c != 0) continue;
    const int ncol = g.N / c;
    if (ncol % kBoxCols != 0 || ncol % g.V != 0 || ncol / g.V > kSrcStride) continue;
    const int per = ncol / kConsumers;
    if (per == 1 || per == 2) { *cl = c; *cpt = per; return true; }
  }
  return false;
}

}  // namespace

bool lattice_cols_supported(const NGram& g, int k, unsigned flags, const void* lexical) {
  if (flags & LT_FLAG_FORCE_GENERIC) return false;
  if ((flags >> LT_FLAG_CLUSTER_SHIFT) & 0xf) return false;
  if (reinterpret_cast<uintptr_t>(lexical) % 16 != 0) return false;
  int cl, cpt;
  if (!cols_geometry(g, &cl, &cpt)) return false;
  const size_t fixed = cols_fixed_bytes(g, g.N / cl);
  const size_t min_stage = (size_t)4 * (g.N / cl) * 4;    // at least 4 rows per stage, 2 stages
  return fixed + 2 * min_stage <= 112 * 1024;
}

int lattice_forward_cols_launch(int semiring, const NGram& g, int k, const FwdParams& base,
                                cudaStream_t stream) {
  int cl = 1, cpt = 1;
  if (!cols_geometry(g, &cl, &cpt)) { set_error("cols path: unsupported geometry"); return LT_ERR_UNSUPPORTED; }
  EncodeTiledFn encode = get_encode_fn3();
  if (!encode) { set_error("cuTensorMapEncodeTiled is unavailable in this driver"); return LT_ERR_CUDA; }
  const int ncol = g.N / cl;
  const size_t budget = 112 * 1024;
  const size_t fixed = cols_fixed_bytes(g, ncol);
  // rows per stage: as few chunks as possible with stages of <= 32 KB and >= 2 stages
  int rmax = (int)(32 * 1024 / ((size_t)ncol * 4));
  if (rmax > kMaxR) rmax = kMaxR;
  while (rmax > 1 && fixed + 2 * (size_t)rmax * ncol * 4 > budget) --rmax;
  if (rmax < 1) { set_error("cols path: not enough shared memory"); return LT_ERR_UNSUPPORTED; }
  const int nchunks = (g.K + rmax - 1) / rmax;
  const int R = (g.K + nchunks - 1) / nchunks;
  const size_t stage = (size_t)R * ncol * 4;
  int stages = (int)((budget - fixed) / stage);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("cols path: not enough shared memory"); return LT_ERR_UNSUPPORTED; }
  const size_t smem = stage * stages + fixed;

  CUtensorMap tmap;
  cuuint64_t dims[3] = {(cuuint64_t)g.N, (cuuint64_t)g.K, (cuuint64_t)base.B * base.T};
  cuuint64_t strides[2] = {(cuuint64_t)g.N * 4, (cuuint64_t)g.C * g.V * 4};
  cuuint32_t box[3] = {(cuuint32_t)kBoxCols, (cuuint32_t)R, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  float* tail = const_cast<float*>(base.lexical) + (size_t)g.Alow * g.V;
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, tail, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (3-D) failed with %d", (int)r); return LT_ERR_CUDA; }

  ColsParams p = {};
  p.g = g; p.k = k; p.B = base.B; p.T = base.T;
  p.R = R; p.nchunks = nchunks; p.stages = stages; p.cl = cl; p.ncol = ncol;
  p.sbuf = (g.K * kSrcStride + g.Alow + 3) & ~3;
  p.blank = base.blank; p.lexical = base.lexical; p.num_frames = base.num_frames;
  p.alpha_init = base.alpha_init; p.dist = base.dist; p.alphas = base.alphas;
  p.alpha_final = base.alpha_final; p.levels = base.levels; p.backptr = base.backptr;
  p.termptr = base.termptr;
  p.alpha_norm = semiring == LT_LOG ? base.alpha_norm : nullptr;
  const int grid = base.B * cl;
  const bool fld = k >= 1;
#define LT_COLS3(SR, CPTV, NORM)                                                             \
  return fld ? launch_cols(lattice_forward_cols<SR, CPTV, true, NORM>, grid, smem, cl, stream, \
                           tmap, p)                                                            \
             : launch_cols(lattice_forward_cols<SR, CPTV, false, NORM>, grid, smem, cl, stream, \
                           tmap, p);
#define LT_COLS2(SR, NORM)               \
  switch (cpt) {                         \
    case 1: LT_COLS3(SR, 1, NORM)        \
    default: LT_COLS3(SR, 2, NORM)       \
  }
  if (semiring == LT_LOG && p.alpha_norm) { LT_COLS2(LT_LOG, true) }
  if (semiring == LT_LOG) { LT_COLS2(LT_LOG, false) }
  if (semiring == LT_MAXTROPICAL) { LT_COLS2(LT_MAXTROPICAL, false) }
  LT_COLS2(LT_REAL, false)
#undef LT_COLS2
#undef LT_COLS3
}

}  // namespace lt
