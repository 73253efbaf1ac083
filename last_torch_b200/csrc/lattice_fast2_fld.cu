// K1/K2 fast path for FrameLabelDependent(k) lattices over bigram contexts (FullNGram
// context_size 1, V in {64, 128, 192, 256}, k <= 3): the kernels of lattice_fast2.cu with the
// frame's K levels run against the SAME resident tile.
//
// alignments.py:362-376 (forward): inside frame t the lattice has alignment states 0 .. k (labels
// emitted in this frame).  last_0 = alpha_t; last_{j+1} = ContextForward(last_j, lexical_t)
// (contexts.py:207-230: the column reduction of lattice_fast2.cu WITHOUT a blank term);
// alpha_{t+1}[q] = (+)_{j=0..k} last_j[q] (x) blank_t[q].  Every level needs the previous one
// complete, so a frame costs k column reductions and k all-gathers (last_1 .. last_{k-1} and
// alpha_{t+1}; last_k is only used by the CTA that produced it) -- but ONE pass over HBM: the
// tile a CTA owns ([V rows x 32 columns], 32 KB at V = 256) stays in its shared-memory stage for
// all k levels, and the TMA ring keeps running NS frames ahead.
//
// alignments.py:378-418 (backward): nb_k = blank (x) beta_{t+1}; for j = k-1 .. 0:
// nb_j[p] = blank[p] (x) beta_{t+1}[p] (+) (+)_y lexical[p, y] (x) nb_{j+1}[next(p, y)], beta_t = nb_0.
// Arc posteriors of the k levels share one weight, so the lexical gradient is their SUM: it is
// accumulated in registers across the levels and written once (fp32 or split rows); the blank
// gradient sums the k + 1 blank arcs of a state.  The level vectors last_1 .. last_k come from
// the forward kernel (`levels [B, T, k, C]`, same renormalised scale as alpha~_t).
//
// NORM (Log): as in lattice_fast2.cu -- the integer offset follows floor(max alpha~_t) once per
// frame; the level vectors and the in-frame nb_j live in the scale of alpha~_t / beta~_{t+1}, and
// every posterior exponent is  last_j[p] + w + nb_{j+1}[q] - (r + d_t).
//
// MaxTropical: back-pointers per level (`backptr [B, T, k, C]`: arg-max source row of the level's
// reduction) and `termptr [B, T, C]`: the number of expansions j of the winning blank term
// (ties: fewest expansions, semirings.py:382), consumed by lt_viterbi_backtrace.
#include <cuda.h>

#include "common.cuh"
#include "fast_ptx.cuh"
#include "params.cuh"
#include "umma.cuh"
#include "fast2.cuh"

namespace lt {

namespace {

using namespace fastptx;

// ============================================================== forward (K1) ==
template <int SR, int V, int K, bool NORM>
__global__ void __launch_bounds__(kGroupThreads, 2)
lattice_forward_fld2(const __grid_constant__ CUtensorMap tmap, const Fast2FwdParams p) {
  static_assert(!NORM || SR == LT_LOG, "renormalisation is a Log-semiring feature");
  using S = Sr<SR>;
  constexpr int CL = V / kCols;                 // cluster size
  constexpr int C = V + 1;
  constexpr int CP = (C + 3) & ~3;
  constexpr int RPT = V / 32;                   // rows per thread
  constexpr uint32_t kStageBytes = V * kCols * 4;
  constexpr int kPart = kGroupWarps * kCols;    // one partial array
  extern __shared__ __align__(128) unsigned char smem2[];
  const int NS = p.stages;

  const int gt = threadIdx.x;
  const int lane = gt & 31, warp = gt >> 5;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / CL;
  __shared__ int b_slot;
  const int b = utterance_of_cluster(cluster_id, p.num_frames, p.B, p.T, &b_slot);
  const bool active = b < p.B;

  float* tiles = reinterpret_cast<float*>(smem2);
  float* small = reinterpret_cast<float*>(smem2 + (size_t)NS * kStageBytes);
  float* alpha_buf = small;                     // [2][CP]  alpha_t / alpha_{t+1}
  float* lvl_buf = alpha_buf + 2 * CP;          // [2][CP]  last_j, alternating per level
  float* part_m = lvl_buf + 2 * CP;             // [2][warps][32]
  float* part_s = part_m + 2 * kPart;           // [2][warps][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(part_s + 2 * kPart);   // NS full + 2 + 2 exchange
  uint64_t* xa = bars + NS;                     // alpha_{t+1} complete
  uint64_t* xl = xa + 2;                        // last_{j+1} complete
  float* wmax = reinterpret_cast<float*>(bars + 16);   // NORM: [2][8] per-warp maxima of alpha~_t

  const int cg = gt & 7;                        // column group: columns 4cg .. 4cg+3
  const int rg = gt >> 3;                       // row group: rows rg*RPT .. +RPT-1
  const int r0 = rg * RPT;
  const int nf = active ? max(0, min(p.num_frames[b], p.T)) : 0;
  const size_t bt0 = (size_t)(active ? b : 0) * p.T;
  const int col0 = rank * kCols;

  if (gt == 0) {
    prefetch_tensormap(&tmap);
    for (int s = 0; s < NS + 4; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = gt; c < CP; c += kGroupThreads) {
    float v = S::zero();
    if (c < C && active)
      v = p.alpha_init ? p.alpha_init[(size_t)b * C + c] : (c == 0 ? S::one() : S::zero());
    alpha_buf[c] = to_dom<SR>(v);
    alpha_buf[CP + c] = S::zero();
    lvl_buf[c] = S::zero();                     // entry 0 stays zero: state 0 has no incoming
    lvl_buf[CP + c] = S::zero();                // lexical arc (contexts.py:217-218)
  }
  __syncthreads();
  cluster_sync_all();

  if (gt == 0) {
    for (int s = 0; s < NS && s < nf; ++s) {
      const uint32_t bar = smem_u32(&bars[s]);
      mbar_arrive_expect_tx(bar, kStageBytes);
      tma_load_2d(smem_u32(tiles) + s * kStageBytes, &tmap, col0, (int)((bt0 + s) * C), bar);
    }
  }

  // finalisers: warp 0, lane j owns destination q = 1 + col0 + j; thread 32 on rank 0 owns state 0
  const bool is_fin = warp == 0;
  const bool is_q0 = (rank == 0 && gt == 32);
  const int q = is_fin ? 1 + col0 + lane : 0;
  float nblank = 0.f, ntail = 0.f;              // prefetched blank[t][q], lexical[t][V][col]
  if (nf > 0) {
    if (is_fin) {
      nblank = ldg_stream(p.blank + bt0 * C + q);
      ntail = ldg_stream(p.lexical + bt0 * (size_t)C * V + (size_t)V * V + col0 + lane);
    } else if (is_q0) {
      nblank = ldg_stream(p.blank + bt0 * C);
    }
  }

  float4 x[RPT];                                // the thread's slice of the tile, raw weights
  int stage = 0;
  uint32_t parity = 0;
  int off = 0;                                  // NORM: alpha_t = alpha~_t + off (log2 units)
  const float* tile = tiles;
  if (nf > 0) {
    mbar_wait(smem_u32(&bars[0]), 0);
#pragma unroll
    for (int i = 0; i < RPT; ++i)
      x[i] = *reinterpret_cast<const float4*>(tile + (size_t)(r0 + i) * kCols + cg * 4);
  }

  for (int t = 0; t < nf; ++t) {
    float* cur = alpha_buf + (t & 1) * CP;
    float* nxt = alpha_buf + ((t + 1) & 1) * CP;
    // alpha_t (t > 0) is complete once every CTA's st.async stores have landed
    if (t > 0) mbar_wait(smem_u32(&xa[t & 1]), ((t - 1) >> 1) & 1);
    if (gt == 0) mbar_arrive_expect_tx(smem_u32(&xa[(t + 1) & 1]), C * 4);
    const float cblank = to_dom<SR>(nblank), ctail = to_dom<SR>(ntail);
    if (t + 1 < nf) {
      if (is_fin) {
        nblank = ldg_stream(p.blank + (bt0 + t + 1) * C + q);
        ntail = ldg_stream(p.lexical + (bt0 + t + 1) * (size_t)C * V + (size_t)V * V + col0 + lane);
      } else if (is_q0) {
        nblank = ldg_stream(p.blank + (bt0 + t + 1) * C);
      }
    }
    if (p.alphas && (is_fin || is_q0)) p.alphas[(bt0 + t) * C + q] = from_dom<SR>(cur[q]);
    float shift = 0.f;
    // the k + 1 blank terms of a destination (Log), their running best (MaxTropical) or sum (Real)
    float term[K + 1];
    int best = 0;
    if (is_fin || is_q0) term[0] = S::times(cur[q], cblank);

#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int e = t * K + j;                  // level step
      float* pm_buf = part_m + (e & 1) * kPart;
      float* ps_buf = part_s + (e & 1) * kPart;
      const float* src = j == 0 ? cur : lvl_buf + ((j - 1) & 1) * CP;
      if (j > 0) {
        const int le = t * (K - 1) + (j - 1);
        mbar_wait(smem_u32(&xl[le & 1]), (le >> 1) & 1);
      }
      if (j + 1 < K && gt == 0) {
        const int le = t * (K - 1) + j;         // states 1 .. V arrive, entry 0 stays zero
        mbar_arrive_expect_tx(smem_u32(&xl[le & 1]), V * 4);
      }
      float a[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) a[i] = src[r0 + i];
      if constexpr (NORM) {
        if (j == 0) {       // d_t = floor(max_{c < V} alpha~_t[c]), see lattice_fast2.cu
          float am = a[0];
#pragma unroll
          for (int i = 1; i < RPT; ++i) am = fmaxf(am, a[i]);
          am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 8));
          am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 16));
          if (lane == 0) wmax[(t & 1) * 8 + warp] = am;
          if (is_q0) p.alpha_norm[(size_t)b * (p.T + 3) + t] = off;
        }
      }

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
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          pm[c] = fmaxf(pm[c], __shfl_xor_sync(0xffffffffu, pm[c], 8));
          pm[c] = fmaxf(pm[c], __shfl_xor_sync(0xffffffffu, pm[c], 16));
        }
      } else if constexpr (SR == LT_MAXTROPICAL) {
        // (max, first arg-max row): rows ascend inside a thread, ties keep the lower row
#pragma unroll
        for (int c = 0; c < 4; ++c) { pm[c] = neg_inf(); ps[c] = __int_as_float(r0); }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const float v0 = x[i].x + a[i], v1 = x[i].y + a[i], v2 = x[i].z + a[i], v3 = x[i].w + a[i];
          if (v0 > pm[0]) { pm[0] = v0; ps[0] = __int_as_float(r0 + i); }
          if (v1 > pm[1]) { pm[1] = v1; ps[1] = __int_as_float(r0 + i); }
          if (v2 > pm[2]) { pm[2] = v2; ps[2] = __int_as_float(r0 + i); }
          if (v3 > pm[3]) { pm[3] = v3; ps[3] = __int_as_float(r0 + i); }
        }
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float om = __shfl_xor_sync(0xffffffffu, pm[c], o);
            const int oa = __shfl_xor_sync(0xffffffffu, __float_as_int(ps[c]), o);
            const int ma = __float_as_int(ps[c]);
            if (om > pm[c] || (om == pm[c] && oa < ma)) { pm[c] = om; ps[c] = __int_as_float(oa); }
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) { pm[c] = 0.f; ps[c] = 0.f; }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          pm[0] = fmaf(a[i], x[i].x, pm[0]); pm[1] = fmaf(a[i], x[i].y, pm[1]);
          pm[2] = fmaf(a[i], x[i].z, pm[2]); pm[3] = fmaf(a[i], x[i].w, pm[3]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          pm[c] += __shfl_xor_sync(0xffffffffu, pm[c], 8);
          pm[c] += __shfl_xor_sync(0xffffffffu, pm[c], 16);
        }
      }
      if (lane < 8) {
        *reinterpret_cast<float4*>(pm_buf + warp * kCols + cg * 4) = make_float4(pm[0], pm[1], pm[2], pm[3]);
        if constexpr (SR == LT_MAXTROPICAL)
          *reinterpret_cast<float4*>(ps_buf + warp * kCols + cg * 4) = make_float4(ps[0], ps[1], ps[2], ps[3]);
      }
      group_sync(0);   // #1: partial maxima visible; every thread holds its tile slice in registers
      if constexpr (NORM) {
        if (j == 0 && warp < 2) {               // finalisers (warp 0) and the state-0 warp
          const float4 w0 = *reinterpret_cast<const float4*>(wmax + (t & 1) * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wmax + (t & 1) * 8 + 4);
          shift = norm_shift(fmaxf(fmaxf(fmaxf(w0.x, w0.y), fmaxf(w0.z, w0.w)),
                                   fmaxf(fmaxf(w1.x, w1.y), fmaxf(w1.z, w1.w))));
          off += (int)shift;
        }
      }
      // after the LAST level every thread is done with the stage: refill it NS frames ahead
      if (j == K - 1 && gt == 0 && t + NS < nf) {
        const uint32_t bar = smem_u32(&bars[stage]);
        mbar_arrive_expect_tx(bar, kStageBytes);
        tma_load_2d(smem_u32(tiles) + stage * kStageBytes, &tmap, col0, (int)((bt0 + t + NS) * C), bar);
      }

      if constexpr (SR == LT_LOG) {
        float4 mx = *reinterpret_cast<const float4*>(pm_buf + cg * 4);
#pragma unroll
        for (int w = 1; w < kGroupWarps; ++w) {
          const float4 o = *reinterpret_cast<const float4*>(pm_buf + w * kCols + cg * 4);
          mx.x = fmaxf(mx.x, o.x); mx.y = fmaxf(mx.y, o.y); mx.z = fmaxf(mx.z, o.z); mx.w = fmaxf(mx.w, o.w);
        }
        const float m0 = msafe(mx.x), m1 = msafe(mx.y), m2 = msafe(mx.z), m3 = msafe(mx.w);
        ps[0] = ps[1] = ps[2] = ps[3] = 0.f;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          ps[0] += ex2(x[i].x - m0);
          ps[1] += ex2(x[i].y - m1);
          ps[2] += ex2(x[i].z - m2);
          ps[3] += ex2(x[i].w - m3);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ps[c] += __shfl_xor_sync(0xffffffffu, ps[c], 8);
          ps[c] += __shfl_xor_sync(0xffffffffu, ps[c], 16);
        }
        if (lane < 8)
          *reinterpret_cast<float4*>(ps_buf + warp * kCols + cg * 4) = make_float4(ps[0], ps[1], ps[2], ps[3]);
        group_sync(0);   // #2: partial sums visible
      }

      // the raw tile slice of the next step: the same stage for the next level, the next stage
      // (advance the ring) for the next frame
      const bool next_frame = j == K - 1;
      if (next_frame) {
        if (++stage == NS) { stage = 0; parity ^= 1; }
        tile = tiles + (size_t)stage * (kStageBytes / 4);
      }
      auto pull = [&]() {
        if (next_frame) {
          if (t + 1 >= nf) return;
          mbar_wait(smem_u32(&bars[stage]), parity);
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i)
          x[i] = *reinterpret_cast<const float4*>(tile + (size_t)(r0 + i) * kCols + cg * 4);
      };
      if (!is_fin) pull();

      if (is_fin) {
        float m = pm_buf[lane];
        float s = (SR == LT_REAL) ? 0.f : ps_buf[lane];
#pragma unroll
        for (int w = 1; w < kGroupWarps; ++w) {
          const float om = pm_buf[w * kCols + lane];
          if constexpr (SR == LT_LOG) {
            m = fmaxf(m, om);
            s += ps_buf[w * kCols + lane];
          } else if constexpr (SR == LT_MAXTROPICAL) {
            const int oa = __float_as_int(ps_buf[w * kCols + lane]);
            if (om > m || (om == m && oa < __float_as_int(s))) { m = om; s = __int_as_float(oa); }
          } else {
            m += om;
          }
        }
        const float xt = S::times(src[V], ctail);       // source row V (not in the TMA box)
        float r;                                        // last_{j+1}[q]
        if constexpr (SR == LT_LOG) {
          lse2_merge(m, s, xt, xt == neg_inf() ? 0.f : 1.f);
          r = msafe(m) + __log2f(s);
          if (m == neg_inf()) r = neg_inf();
        } else if constexpr (SR == LT_MAXTROPICAL) {
          int am = __float_as_int(s);
          if (xt > m) { m = xt; am = V; }
          r = m;
          if (p.backptr) p.backptr[((bt0 + t) * K + j) * C + q] = (int16_t)am;
        } else {
          r = m + xt;
        }
        if (p.levels) p.levels[((bt0 + t) * K + j) * C + q] = from_dom<SR>(r);
        const float tj = S::times(r, cblank);
        if constexpr (SR == LT_LOG) {
          term[j + 1] = tj;
        } else if constexpr (SR == LT_MAXTROPICAL) {
          if (tj > term[0]) { term[0] = tj; best = j + 1; }   // ties: fewest expansions
        } else {
          term[0] += tj;
        }
        if (j + 1 < K) {
          const int le = t * (K - 1) + j;
          xchg_store(lvl_buf + (j & 1) * CP, q, r, &xl[le & 1], CL);
        } else {
          float v;
          if constexpr (SR == LT_LOG) {
            float tm = term[0];
#pragma unroll
            for (int i = 1; i <= K; ++i) tm = fmaxf(tm, term[i]);
            const float ts = msafe(tm);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i <= K; ++i) sum += ex2(term[i] - ts);
            v = (tm == neg_inf()) ? neg_inf() : ts + __log2f(sum) - shift;
          } else {
            v = term[0];
          }
          if constexpr (SR == LT_MAXTROPICAL) {
            if (p.termptr) p.termptr[(bt0 + t) * C + q] = (uint8_t)best;
          }
          xchg_store(nxt, q, v, &xa[(t + 1) & 1], CL);
        }
        pull();
      } else if (is_q0) {
        // state 0: no incoming lexical arc, last_{j+1}[0] = zero
        if (p.levels) p.levels[((bt0 + t) * K + j) * C] = S::zero();
        if constexpr (SR == LT_MAXTROPICAL) {
          if (p.backptr) p.backptr[((bt0 + t) * K + j) * C] = (int16_t)0;
        }
        if (j + 1 == K) {
          const float v = term[0] - shift;
          if constexpr (SR == LT_MAXTROPICAL) { if (p.termptr) p.termptr[(bt0 + t) * C] = (uint8_t)0; }
          xchg_store(nxt, 0, v, &xa[(t + 1) & 1], CL);
        }
      }
    }
  }
  float* cur = alpha_buf + (nf & 1) * CP;
  if (nf > 0) mbar_wait(smem_u32(&xa[nf & 1]), ((nf - 1) >> 1) & 1);

  // padding frames keep alpha (lattices.py:460-461) and are still recorded (:462)
  if (active && (is_fin || is_q0)) {
    if (p.alphas)
      for (int t = nf; t < p.T; ++t) p.alphas[(bt0 + t) * C + q] = from_dom<SR>(cur[q]);
    if (p.alpha_final)
      p.alpha_final[(size_t)b * C + q] =
          NORM ? (float)(((double)cur[q] + (double)off) * 0.6931471805599453) : from_dom<SR>(cur[q]);
  }
  if (active && rank == 0 && warp == 1) {       // dist = (+)_c alpha_T[c]  (lattices.py:496)
    if constexpr (SR == LT_LOG) {
      float m = neg_inf();
      for (int c = lane; c < C; c += 32) m = fmaxf(m, cur[c]);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float ms = msafe(m);
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s += ex2(cur[c] - ms);            // log2 domain
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if constexpr (NORM) {
        const float r = ms + __log2f(s);
        int32_t* an = p.alpha_norm + (size_t)b * (p.T + 3);
        for (int t = nf + lane; t <= p.T; t += 32) an[t] = off;
        if (lane == 0) {
          an[p.T + 1] = __float_as_int(r);
          an[p.T + 2] = 0;               // offsets are in log2 units
          p.dist[b] = (float)(((double)r + (double)off) * 0.6931471805599453);
        }
      } else {
        if (lane == 0) p.dist[b] = (ms + __log2f(s)) * kLn2;
      }
    } else {
      float m = (SR == LT_REAL) ? 0.f : neg_inf();
      for (int c = lane; c < C; c += 32) m = S::plus(m, cur[c]);
      for (int o = 16; o > 0; o >>= 1) m = S::plus(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) p.dist[b] = m;
    }
  }
  cluster_sync_all();
}

// ============================================================= backward (K2) ==
template <int SR, int V, int K, bool SPLIT, bool NORM>
__global__ void __launch_bounds__(kGroupThreads, 2)
lattice_backward_fld2(const Fast2BwdParams p) {
  static_assert(!NORM || SR == LT_LOG, "renormalisation is a Log-semiring feature");
  static_assert(SR != LT_MAXTROPICAL, "MaxTropical gradients come from the back-trace");
  constexpr int CL = V / kCols;
  constexpr int C = V + 1;
  constexpr int CH = V / 32;                    // float4 chunks per lane (8 lanes per row)
  constexpr int kRows = kCols;                  // rows per CTA (+ tail row V on the last rank)
  constexpr int TCH = (V + 127) / 128;          // float4 chunks per lane of the tail row
  constexpr uint32_t kSlabBytes = kRows * V * 4;
  constexpr uint32_t kStageBytes = kSlabBytes + V * 4;
  constexpr int BP = ((C + 3 + 3) & ~3) + 4;    // vectors over states: entry q at index 3 + q
  extern __shared__ __align__(128) unsigned char smem2[];
  const int NS = p.stages;

  const int gt = threadIdx.x;
  const int lane = gt & 31, warp = gt >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool last_rank = rank == CL - 1;
  const int cluster_id = blockIdx.x / CL;
  __shared__ int b_slot;
  const int b = utterance_of_cluster(cluster_id, p.num_frames, p.B, p.T, &b_slot);
  const bool active = b < p.B;

  float* tiles = reinterpret_cast<float*>(smem2);
  float* small = reinterpret_cast<float*>(smem2 + (size_t)NS * kStageBytes);
  float* beta_buf = small;                      // [2][BP]  beta_{t+1} / beta_t
  float* bb_buf = beta_buf + 2 * BP;            // [BP]     blank (x) beta_{t+1} = nb_k
  float* nb_buf = bb_buf + BP;                  // [2][BP]  nb_j, alternating per level
  uint64_t* bars = reinterpret_cast<uint64_t*>(nb_buf + 2 * BP);
  uint64_t* xbar = bars + NS;                   // beta_t complete
  uint64_t* xn = xbar + 2;                      // nb_j complete

  const int sub = lane >> 3, sl = lane & 7;     // row within the warp, lane within the row
  const int row = warp * 4 + sub;               // local row 0..31
  const int prow = rank * kRows + row;          // source state
  const int nf = active ? max(0, min(p.num_frames[b], p.T)) : 0;
  const size_t bt0 = (size_t)(active ? b : 0) * p.T;
  const float logz = active ? p.dist[b] : 0.f;
  const int32_t* an = NORM ? p.alpha_norm + (size_t)(active ? b : 0) * (p.T + 3) : nullptr;
  const float logz2 = NORM ? __int_as_float(an[p.T + 1]) : logz * kLog2e;
  const float gscale = (active && p.grad_dist) ? p.grad_dist[b] : 1.f;
  const bool scale_ok = (SR != LT_LOG) || is_finite(logz);
  constexpr bool split = SPLIT;
  const uint32_t stage_tx = last_rank ? kStageBytes : kSlabBytes;

  if (gt == 0) {
    for (int s = 0; s < NS + 4; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = gt; c < 5 * BP; c += kGroupThreads) beta_buf[c] = Sr<SR>::one();   // lattices.py:789-790
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
  if (gt == 0)
    for (int it = 0; it < NS && it < nf; ++it) issue(it);

  // padding frames: zero gradients (lattices.py:775-779)
  if (active) {
    for (int t = nf; t < p.T; ++t) {
      float4* gl = reinterpret_cast<float4*>(p.grad_lexical + (bt0 + t) * (size_t)C * V +
                                             (size_t)rank * kRows * V);
      for (int i = gt; i < kRows * V / 4; i += kGroupThreads)
        stg_stream4(reinterpret_cast<float*>(gl + i), make_float4(0, 0, 0, 0));
      if (gt < kRows) p.grad_blank[(bt0 + t) * C + rank * kRows + gt] = 0.f;
      if (last_rank) {
        float* tail = p.grad_lexical + (bt0 + t) * (size_t)C * V + (size_t)V * V;
        for (int i = gt; i < V; i += kGroupThreads) tail[i] = 0.f;
        if (gt == 0) p.grad_blank[(bt0 + t) * C + V] = 0.f;
      }
    }
  }

  // Row owners (lane sl == 0) prefetch, one frame ahead, the k + 1 source values of their state
  // (alpha_t[p], last_1[p] .. last_k[p]); warp 0 lane 0 of the last rank also owns the tail row V.
  // Every thread prefetches one entry of blank_t (thread 0 also entry V): nb_k needs all of it.
  const bool owner = sl == 0;
  const bool tail_owner = last_rank && warp == 0 && lane == 0;
  float n_src[K + 1], n_tsrc[K + 1];
  float n_bl = 0.f, n_bl2 = 0.f;
  int n_off = 0, c_off1 = 0;                    // NORM: off_t (prefetched), off_{t+1}
#pragma unroll
  for (int i = 0; i <= K; ++i) n_src[i] = n_tsrc[i] = 0.f;
  auto prefetch = [&](int t) {
    const size_t o = (bt0 + t) * C;
    const float* lev = p.levels + (bt0 + t) * (size_t)K * C;
    if (owner) {
      n_src[0] = p.alphas[o + prow];
#pragma unroll
      for (int i = 1; i <= K; ++i) n_src[i] = lev[(size_t)(i - 1) * C + prow];
    }
    if (tail_owner) {
      n_tsrc[0] = p.alphas[o + V];
#pragma unroll
      for (int i = 1; i <= K; ++i) n_tsrc[i] = lev[(size_t)(i - 1) * C + V];
    }
    if (gt < V) n_bl = ldg_stream(p.blank + o + gt);
    if (gt == 0) n_bl2 = ldg_stream(p.blank + o + V);
  };
  if (nf > 0) {
    prefetch(nf - 1);
    if constexpr (NORM) { n_off = an[nf - 1]; c_off1 = an[nf]; }
  }

  int stage = 0;
  uint32_t parity = 0;
  for (int it = 0; it < nf; ++it) {
    const int t = nf - 1 - it;
    float* beta = beta_buf + (it & 1) * BP;          // beta_{t+1}; entry q at beta[3 + q]
    float* nxt = beta_buf + ((it + 1) & 1) * BP;
    if (it > 0) {
      // every row of the previous frame has been reduced cluster-wide: beta is complete and
      // the tile stage of iteration it-1 is free for the next TMA
      mbar_wait(smem_u32(&xbar[it & 1]), ((it - 1) >> 1) & 1);
      if (gt == 0 && it - 1 + NS < nf) issue(it - 1 + NS);
    }
    if (gt == 0) mbar_arrive_expect_tx(smem_u32(&xbar[(it + 1) & 1]), C * 4);
    float c_src[K + 1], c_tsrc[K + 1];
#pragma unroll
    for (int i = 0; i <= K; ++i) { c_src[i] = to_dom<SR>(n_src[i]); c_tsrc[i] = to_dom<SR>(n_tsrc[i]); }
    const float c_bl = n_bl, c_bl2 = n_bl2;
    const float shift = NORM ? (float)(c_off1 - n_off) : 0.f;
    const float zref = logz2 + shift;
    if constexpr (NORM) c_off1 = n_off;
    if (t > 0) {
      prefetch(t - 1);
      if constexpr (NORM) n_off = an[t - 1];
    }
    // nb_k = blank (x) beta_{t+1}, all states, by every CTA (alignments.py:405)
    if (gt < V) bb_buf[3 + gt] = arc<SR>(c_bl, beta[3 + gt]);
    if (gt == 0) bb_buf[3 + V] = arc<SR>(c_bl2, beta[3 + V]);
    mbar_wait(smem_u32(&bars[stage]), parity);
    const float* tile = tiles + (size_t)stage * (kStageBytes / 4);
    if (++stage == NS) { stage = 0; parity ^= 1; }
    group_sync(0);                                    // nb_k visible
    float* gl = p.grad_lexical + (bt0 + t) * (size_t)C * V;
    float* gb = p.grad_blank + (bt0 + t) * C;

    float4 acc[CH];                                   // lexical gradient of this thread's arcs
    float4 tacc[TCH];                                 // ... of the tail row (last rank, warp 0)
#pragma unroll
    for (int j = K - 1; j >= 0; --j) {
      // destination vector nb_{j+1}
      const float* nbv;
      if (j == K - 1) {
        nbv = bb_buf;
      } else {
        const int ne = it * (K - 1) + (K - 2 - j);
        nbv = nb_buf + ((K - 2 - j) & 1) * BP;
        mbar_wait(smem_u32(&xn[ne & 1]), (ne >> 1) & 1);
      }
      if (j > 0 && gt == 0) {
        const int ne = it * (K - 1) + (K - 1 - j);
        mbar_arrive_expect_tx(smem_u32(&xn[ne & 1]), C * 4);
      }
      const float* bnext = nbv + 4;                   // bnext[y] = nb_{j+1}[1 + y]
      // where nb_j goes: the level buffers, or beta_t (j == 0)
      float* dst = j > 0 ? nb_buf + ((K - 1 - j) & 1) * BP : nxt;
      uint64_t* dbar = j > 0 ? &xn[(it * (K - 1) + (K - 1 - j)) & 1] : &xbar[(it + 1) & 1];
      const float dshift = j == 0 ? shift : 0.f;
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
        const float src_p = __shfl_sync(0xffffffffu, c_src[j], lane & ~7);
        float rowsum;
        if constexpr (SR == LT_LOG) {
          float m = neg_inf();
#pragma unroll
          for (int i = 0; i < CH; ++i) m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
          const float ms = msafe(m);
          const float rs = scale_ok ? gscale * ex2(src_p + ms - zref) : 0.f;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            float4 e;
            e.x = ex2(x[i].x - ms); e.y = ex2(x[i].y - ms);
            e.z = ex2(x[i].z - ms); e.w = ex2(x[i].w - ms);
            s += (e.x + e.y) + (e.z + e.w);
            if (j == K - 1) acc[i] = make_float4(e.x * rs, e.y * rs, e.z * rs, e.w * rs);
            else acc[i] = make_float4(fmaf(e.x, rs, acc[i].x), fmaf(e.y, rs, acc[i].y),
                                      fmaf(e.z, rs, acc[i].z), fmaf(e.w, rs, acc[i].w));
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          rowsum = (m == neg_inf()) ? neg_inf() : ms + __log2f(s);
        } else {
          float s = 0.f;
          const float ga = gscale * src_p;
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int c4 = (sl + 8 * i) * 4;
            const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
            s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
            if (j == K - 1) acc[i] = make_float4(ga * bn.x, ga * bn.y, ga * bn.z, ga * bn.w);
            else acc[i] = make_float4(fmaf(ga, bn.x, acc[i].x), fmaf(ga, bn.y, acc[i].y),
                                      fmaf(ga, bn.z, acc[i].z), fmaf(ga, bn.w, acc[i].w));
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          rowsum = s;
        }
        if (j == 0) {
          float* grow = gl + (size_t)prow * V;
#pragma unroll
          for (int i = 0; i < CH; ++i) store_grad4(grow, (sl + 8 * i) * 4, V, split, acc[i]);
        }
        if (owner) {
          const float bbp = bb_buf[3 + prow];
          if (j == 0) {
            // blank marginals of the k + 1 alignment states (alignments.py:398-403)
            float g = 0.f;
#pragma unroll
            for (int i = 0; i <= K; ++i) {
              if constexpr (SR == LT_LOG) g += scale_ok ? gscale * ex2(c_src[i] + bbp - zref) : 0.f;
              else g += gscale * c_src[i] * beta[3 + prow];
            }
            gb[prow] = g;
          }
          // (the all-gather stays on the owner lane here: spreading it over the row group, as the
          // FrameDependent kernels do, measured 8 % slower in this kernel -- 6.20 -> 6.68 ms)
          xchg_store(dst, 3 + prow,
                     SR == LT_LOG ? log2_add_exp2(bbp, rowsum) - dshift : bbp + rowsum, dbar, CL);
        }
      }
      if (last_rank && warp == 0) {          // tail row: source state V, all 32 lanes
        const float* trow = tile + (size_t)kRows * V;
        const float src_p = __shfl_sync(0xffffffffu, c_tsrc[j], 0);
        float rowsum;
        if constexpr (SR == LT_LOG) {
          float4 x[TCH];
          float m = neg_inf();
#pragma unroll
          for (int i = 0; i < TCH; ++i) {
            const int c4 = lane * 4 + i * 128;
            if (c4 < V) {
              const float4 w = *reinterpret_cast<const float4*>(trow + c4);
              const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
              x[i] = make_float4(arc<SR>(w.x, bn.x), arc<SR>(w.y, bn.y), arc<SR>(w.z, bn.z),
                                 arc<SR>(w.w, bn.w));
              m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
            }
          }
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          const float ms = msafe(m);
          const float rs = scale_ok ? gscale * ex2(src_p + ms - zref) : 0.f;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < TCH; ++i) {
            const int c4 = lane * 4 + i * 128;
            if (c4 < V) {
              float4 e;
              e.x = ex2(x[i].x - ms); e.y = ex2(x[i].y - ms);
              e.z = ex2(x[i].z - ms); e.w = ex2(x[i].w - ms);
              s += (e.x + e.y) + (e.z + e.w);
              if (j == K - 1) tacc[i] = make_float4(e.x * rs, e.y * rs, e.z * rs, e.w * rs);
              else tacc[i] = make_float4(fmaf(e.x, rs, tacc[i].x), fmaf(e.y, rs, tacc[i].y),
                                         fmaf(e.z, rs, tacc[i].z), fmaf(e.w, rs, tacc[i].w));
            }
          }
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          rowsum = (m == neg_inf()) ? neg_inf() : ms + __log2f(s);
        } else {
          float s = 0.f;
          const float ga = gscale * src_p;
#pragma unroll
          for (int i = 0; i < TCH; ++i) {
            const int c4 = lane * 4 + i * 128;
            if (c4 < V) {
              const float4 w = *reinterpret_cast<const float4*>(trow + c4);
              const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
              s += (w.x * bn.x + w.y * bn.y) + (w.z * bn.z + w.w * bn.w);
              if (j == K - 1) tacc[i] = make_float4(ga * bn.x, ga * bn.y, ga * bn.z, ga * bn.w);
              else tacc[i] = make_float4(fmaf(ga, bn.x, tacc[i].x), fmaf(ga, bn.y, tacc[i].y),
                                         fmaf(ga, bn.z, tacc[i].z), fmaf(ga, bn.w, tacc[i].w));
            }
          }
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          rowsum = s;
        }
        if (j == 0) {
          float* grow = gl + (size_t)V * V;
#pragma unroll
          for (int i = 0; i < TCH; ++i) {
            const int c4 = lane * 4 + i * 128;
            if (c4 < V) store_grad4(grow, c4, V, split, tacc[i]);
          }
        }
        if (lane == 0) {
          const float bbp = bb_buf[3 + V];
          if (j == 0) {
            float g = 0.f;
#pragma unroll
            for (int i = 0; i <= K; ++i) {
              if constexpr (SR == LT_LOG) g += scale_ok ? gscale * ex2(c_tsrc[i] + bbp - zref) : 0.f;
              else g += gscale * c_tsrc[i] * beta[3 + V];
            }
            gb[V] = g;
          }
          xchg_store(dst, 3 + V,
                     SR == LT_LOG ? log2_add_exp2(bbp, rowsum) - dshift : bbp + rowsum, dbar, CL);
        }
      }
    }
  }
  float* beta = beta_buf + (nf & 1) * BP;
  if (nf > 0) mbar_wait(smem_u32(&xbar[nf & 1]), ((nf - 1) >> 1) & 1);
  if (active && p.beta_final) {
    const double boff = NORM ? (double)(an[p.T] - an[0]) : 0.0;
    auto out = [&](float v) {
      return NORM ? (float)(((double)v + boff) * 0.6931471805599453) : from_dom<SR>(v);
    };
    if (gt < kRows)
      p.beta_final[(size_t)b * C + rank * kRows + gt] = out(beta[3 + rank * kRows + gt]);
    if (last_rank && gt == 0) p.beta_final[(size_t)b * C + V] = out(beta[3 + V]);
  }
  cluster_sync_all();
}

}  // namespace

bool lattice_fast2_fld_supported(const NGram& g, int k, unsigned flags, const void* lexical) {
  if (option(OPT_FLD_GENERIC)) return false;
  if (flags & (LT_FLAG_FORCE_GENERIC | LT_FLAG_LEVEL_WEIGHTS)) return false;
  if ((flags >> LT_FLAG_CLUSTER_SHIFT) & 0xf) return false;   // explicit cluster size => generic
  if (k < 1 || k > 3 || g.n != 1) return false;
  if (g.V % 64 != 0 || g.V > 256) return false;
  if (lexical && reinterpret_cast<uintptr_t>(lexical) % 16 != 0) return false;
  return true;
}

int lattice_forward_fld2_launch(int semiring, const NGram& g, const FwdParams& base,
                                unsigned flags, cudaStream_t stream) {
  const int V = g.V, C = g.C, CL = V / kCols, K = base.k;
  EncodeTiledFn encode = get_encode_fn2();
  if (!encode) { set_error("cuTensorMapEncodeTiled is unavailable in this driver"); return LT_ERR_CUDA; }
  CUtensorMap tmap;
  const cuuint64_t rows = (cuuint64_t)base.B * base.T * C;
  cuuint64_t dims[2] = {(cuuint64_t)V, rows};
  cuuint64_t strides[1] = {(cuuint64_t)V * 4};
  cuuint32_t box[2] = {(cuuint32_t)kCols, (cuuint32_t)V};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base.lexical),
                      dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)r); return LT_ERR_CUDA; }
  const size_t stage = (size_t)V * kCols * 4;
  const int CP = (C + 3) & ~3;
  const size_t small = sizeof(float) * (4 * CP + 4 * kGroupWarps * kCols + 2 * 16 + 16);
  int stages = (int)((kSharedBudget - small - 256) / stage);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("fast forward: not enough shared memory"); return LT_ERR_UNSUPPORTED; }
  const size_t smem = stage * stages + small;
  Fast2FwdParams p = {};
  p.B = base.B; p.T = base.T; p.stages = stages;
  p.blank = base.blank; p.lexical = base.lexical; p.num_frames = base.num_frames;
  p.alpha_init = base.alpha_init; p.dist = base.dist; p.alphas = base.alphas;
  p.alpha_final = base.alpha_final; p.backptr = base.backptr;
  p.alpha_norm = semiring == LT_LOG ? base.alpha_norm : nullptr;
  p.levels = base.levels; p.termptr = base.termptr;
  const int grid = base.B * CL;
#define LT_FLDF(SR, VV, KK, NORM) \
  return launch_fast2(lattice_forward_fld2<SR, VV, KK, NORM>, grid, kGroupThreads, smem, CL, stream, tmap, p);
#define LT_FLDFK(SR, VV, NORM)             \
  switch (K) {                             \
    case 1: LT_FLDF(SR, VV, 1, NORM)       \
    case 2: LT_FLDF(SR, VV, 2, NORM)       \
    default: LT_FLDF(SR, VV, 3, NORM)      \
  }
#define LT_FLDFV(SR, NORM)                 \
  switch (V) {                             \
    case 64: LT_FLDFK(SR, 64, NORM)        \
    case 128: LT_FLDFK(SR, 128, NORM)      \
    case 192: LT_FLDFK(SR, 192, NORM)      \
    default: LT_FLDFK(SR, 256, NORM)       \
  }
  if (semiring == LT_LOG && p.alpha_norm) { LT_FLDFV(LT_LOG, true) }
  if (semiring == LT_LOG) { LT_FLDFV(LT_LOG, false) }
  if (semiring == LT_MAXTROPICAL) { LT_FLDFV(LT_MAXTROPICAL, false) }
  LT_FLDFV(LT_REAL, false)
#undef LT_FLDFV
#undef LT_FLDFK
#undef LT_FLDF
}

int lattice_backward_fld2_launch(int semiring, const NGram& g, const BwdParams& base,
                                 unsigned flags, cudaStream_t stream) {
  const int V = g.V, C = g.C, CL = V / kCols, K = base.k;
  const size_t stage = (size_t)kCols * V * 4 + (size_t)V * 4;
  const int BP = ((C + 6) & ~3) + 4;
  const size_t small = sizeof(float) * (5 * BP + 2 * 16);
  int stages = (int)((kSharedBudget - small - 256) / stage);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("fast backward: not enough shared memory"); return LT_ERR_UNSUPPORTED; }
  const size_t smem = stage * stages + small;
  Fast2BwdParams p = {};
  p.B = base.B; p.T = base.T; p.stages = stages;
  p.blank = base.blank; p.lexical = base.lexical; p.num_frames = base.num_frames;
  p.alphas = base.alphas; p.dist = base.dist; p.grad_dist = base.grad_dist;
  p.grad_blank = base.grad_blank; p.grad_lexical = base.grad_lexical; p.beta_final = base.beta_final;
  p.split = (flags & LT_FLAG_GRAD_SPLIT) ? 1 : 0;
  p.alpha_norm = semiring == LT_LOG ? base.alpha_norm : nullptr;
  p.levels = base.levels;
  const int grid = base.B * CL;
#define LT_FLDB(SR, VV, KK, SPLIT, NORM) \
  return launch_fast2(lattice_backward_fld2<SR, VV, KK, SPLIT, NORM>, grid, kGroupThreads, smem, CL, stream, p);
#define LT_FLDBS(SR, VV, KK, NORM)                    \
  if (p.split) { LT_FLDB(SR, VV, KK, true, NORM) }    \
  LT_FLDB(SR, VV, KK, false, NORM)
#define LT_FLDBK(SR, VV, NORM)             \
  switch (K) {                             \
    case 1: LT_FLDBS(SR, VV, 1, NORM)      \
    case 2: LT_FLDBS(SR, VV, 2, NORM)      \
    default: LT_FLDBS(SR, VV, 3, NORM)     \
  }
#define LT_FLDBV(SR, NORM)                 \
  switch (V) {                             \
    case 64: LT_FLDBK(SR, 64, NORM)        \
    case 128: LT_FLDBK(SR, 128, NORM)      \
    case 192: LT_FLDBK(SR, 192, NORM)      \
    default: LT_FLDBK(SR, 256, NORM)       \
  }
  if (semiring == LT_LOG && p.alpha_norm) { LT_FLDBV(LT_LOG, true) }
  if (semiring == LT_LOG) { LT_FLDBV(LT_LOG, false) }
  LT_FLDBV(LT_REAL, false)
#undef LT_FLDBV
#undef LT_FLDBK
#undef LT_FLDBS
#undef LT_FLDB
}

}  // namespace lt
