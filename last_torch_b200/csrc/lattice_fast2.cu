// K1/K2 fast path (FullNGram context_size 1, FrameDependent, V in {64, 128, 192, 256}),
// organised so that the sequential dependency chain of one utterance (reduce -> finalise ->
// DSMEM all-gather -> wake-up) is hidden behind the arithmetic of another one on the same SM:
//   * a cluster of CL = V/32 CTAs owns one utterance; a CTA is one 256-thread group working on
//     32 destination columns (forward) / 32 source rows (backward); it uses at most 113 KB of
//     shared memory and 128 registers per thread, so TWO CTAs (of different utterances) share
//     an SM and drift freely: while one sits in its finaliser / exchange phase the other one
//     keeps the FMA/MUFU pipes busy;
//   * each CTA has its own TMA ring (3 stages of 32 KB at V = 256) fed by one elected thread,
//     several frames ahead of the recursion;
//   * forward, Log: the exact column maximum is established FIRST (per-thread max,
//     two shuffles, one 1 KB exchange through shared memory), then every arc costs
//     exactly one FFMA + one FADD + one MUFU.EX2 + one FADD; no (max, sum) pair
//     merges anywhere, the finaliser adds eight partial sums;
//   * the next frame's tile is pulled into registers BEFORE the group blocks on the
//     alpha exchange, so the shared-memory reads are off the critical path;
//   * state exchange: st.async + mbarrier complete_tx, no cluster barrier and no fence
//     inside the loop;
//   * NORM (Log semiring, `alpha_norm` given): alpha is kept RENORMALISED on chip and in the
//     `alphas` buffer -- alpha_t = alpha~_t + off_t with an exact integer offset (log2 units)
//     off_{t+1} = off_t + floor(max_c alpha~_t[c]) -- so that every sum the recursion rounds
//     (w + alpha, alpha + w + beta - logZ) has magnitude O(10) instead of O(logZ): at T = 1000
//     (logZ ~ 5.5e3, one fp32 ulp = 4.9e-4) the arc posteriors keep ~1e-6 relative accuracy
//     where the plain fp32 recursion (and the fp32 reference) has 1e-4..1e-3.  The backward
//     kernel needs no normaliser of its own: beta~_t = beta_t - (off_T - off_t) follows the
//     SAME per-frame shifts d_t = off_{t+1} - off_t, and the posterior exponent becomes
//     alpha~_t[p] + w + beta~_{t+1}[q] - d_t - r with logZ = off_T + r.
//
// Reference semantics: lattices.py:436-462 + alignments.py:294-297 (forward),
// alignments.py:300-318 + lattices.py:775-779 (backward), contexts.py:207-256.
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
template <int SR, int V, bool NORM>
__global__ void __launch_bounds__(kGroupThreads, 2)
lattice_forward_fast2(const __grid_constant__ CUtensorMap tmap, const Fast2FwdParams p) {
  static_assert(!NORM || SR == LT_LOG, "renormalisation is a Log-semiring feature");
  constexpr int G = 1;
  using S = Sr<SR>;
  constexpr int CL = V / kCols;                 // cluster size
  constexpr int C = V + 1;
  constexpr int CP = (C + 3) & ~3;
  constexpr int RPT = V / 32;                   // rows per thread
  constexpr uint32_t kStageBytes = V * kCols * 4;
  constexpr int kPart = kGroupWarps * kCols;    // one partial array
  extern __shared__ __align__(128) unsigned char smem2[];
  const int NS = p.stages;

  const int tid = threadIdx.x;
  const int grp = G == 1 ? 0 : tid >> 8;       // warp-uniform
  const int gt = tid & (kGroupThreads - 1);
  const int lane = gt & 31, warp = gt >> 5;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x / CL;
  __shared__ int b_slot;
  const int b = utterance_of_cluster(cluster_id, p.num_frames, p.B, p.T, &b_slot);
  const bool active = b < p.B;

  // shared-memory carve-up: [group][stage] tiles, then per-group small state
  float* tiles = reinterpret_cast<float*>(smem2) + (size_t)grp * NS * (kStageBytes / 4);
  float* small = reinterpret_cast<float*>(smem2 + (size_t)G * NS * kStageBytes);
  constexpr int kSmallFloats = 2 * CP + 4 * kPart + 2 * 16 + 16;  // alpha x2, pmax x2, psum x2, bars, wmax x2
  small += (size_t)grp * kSmallFloats;
  float* alpha_buf = small;
  float* part_m = alpha_buf + 2 * CP;           // [2][warps][32]
  float* part_s = part_m + 2 * kPart;           // [2][warps][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(part_s + 2 * kPart);   // NS full + 2 exchange
  uint64_t* xbar = bars + NS;
  float* wmax = reinterpret_cast<float*>(bars + 16);   // NORM: [2][8] per-warp maxima of alpha~_t

  const int cg = gt & 7;                        // column group: columns 4cg .. 4cg+3
  const int rg = gt >> 3;                       // row group: rows rg*RPT .. +RPT-1
  const int r0 = rg * RPT;
  const int nf = active ? max(0, min(p.num_frames[b], p.T)) : 0;
  const size_t bt0 = (size_t)(active ? b : 0) * p.T;
  const int col0 = rank * kCols;

  if (gt == 0) {
    if (grp == 0) prefetch_tensormap(&tmap);
    for (int s = 0; s < NS + 2; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = gt; c < CP; c += kGroupThreads) {
    float v = S::zero();
    if (c < C && active)
      v = p.alpha_init ? p.alpha_init[(size_t)b * C + c] : (c == 0 ? S::one() : S::zero());
    alpha_buf[c] = to_dom<SR>(v);
    alpha_buf[CP + c] = S::zero();
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

  // finalisers: warp 0, lane j owns destination q = 1 + col0 + j; thread 32 of the
  // group on rank 0 owns state 0 (no incoming lexical arc, contexts.py:217-218).
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

  float4 x[RPT];
  int stage = 0;
  uint32_t parity = 0;
  int off = 0;                                  // NORM: alpha_t = alpha~_t + off (log2 units)
  if (nf > 0) {
    mbar_wait(smem_u32(&bars[0]), 0);
#pragma unroll
    for (int i = 0; i < RPT; ++i)
      x[i] = *reinterpret_cast<const float4*>(tiles + (size_t)(r0 + i) * kCols + cg * 4);
  }

  for (int t = 0; t < nf; ++t) {
    float* cur = alpha_buf + (t & 1) * CP;
    float* nxt = alpha_buf + ((t + 1) & 1) * CP;
    float* pm_buf = part_m + (t & 1) * kPart;
    float* ps_buf = part_s + (t & 1) * kPart;
    // alpha_t (t > 0) is complete once every CTA's st.async stores have landed
    if (t > 0) mbar_wait(smem_u32(&xbar[t & 1]), ((t - 1) >> 1) & 1);
    if (gt == 0) mbar_arrive_expect_tx(smem_u32(&xbar[(t + 1) & 1]), C * 4);
    const float cblank = nblank, ctail = ntail;
    if (t + 1 < nf) {
      if (is_fin) {
        nblank = ldg_stream(p.blank + (bt0 + t + 1) * C + q);
        ntail = ldg_stream(p.lexical + (bt0 + t + 1) * (size_t)C * V + (size_t)V * V + col0 + lane);
      } else if (is_q0) {
        nblank = ldg_stream(p.blank + (bt0 + t + 1) * C);
      }
    }
    if (p.alphas && (is_fin || is_q0)) p.alphas[(bt0 + t) * C + q] = from_dom<SR>(cur[q]);
    float a[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) a[i] = cur[r0 + i];
    if constexpr (NORM) {
      // d_t = floor(max_{c < V} alpha~_t[c]): every CTA derives it from its own (identical)
      // replica, so no exchange is needed; an integer keeps `off` exact and the subtraction
      // below exact.  The maximum is taken over the values every thread has just loaded (the
      // row groups of a warp cover 32 states, the 8 warps states 0 .. V-1; the shift need not
      // see state V) and merged through the shared-memory hand-over that sync #1 orders anyway:
      // no extra pass over alpha and nothing added to the frame's dependency chain.
      float am = a[0];
#pragma unroll
      for (int i = 1; i < RPT; ++i) am = fmaxf(am, a[i]);
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 8));
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 16));
      if (lane == 0) wmax[(t & 1) * 8 + warp] = am;
      if (is_q0) p.alpha_norm[(size_t)b * (p.T + 3) + t] = off;
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
      for (int j = 0; j < 4; ++j) {
        pm[j] = fmaxf(pm[j], __shfl_xor_sync(0xffffffffu, pm[j], 8));
        pm[j] = fmaxf(pm[j], __shfl_xor_sync(0xffffffffu, pm[j], 16));
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
      for (int o = 8; o <= 16; o <<= 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float om = __shfl_xor_sync(0xffffffffu, pm[j], o);
          const int oa = __shfl_xor_sync(0xffffffffu, __float_as_int(ps[j]), o);
          const int ma = __float_as_int(ps[j]);
          if (om > pm[j] || (om == pm[j] && oa < ma)) { pm[j] = om; ps[j] = __int_as_float(oa); }
        }
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
      for (int j = 0; j < 4; ++j) {
        pm[j] += __shfl_xor_sync(0xffffffffu, pm[j], 8);
        pm[j] += __shfl_xor_sync(0xffffffffu, pm[j], 16);
      }
    }
    if (lane < 8) {
      *reinterpret_cast<float4*>(pm_buf + warp * kCols + cg * 4) = make_float4(pm[0], pm[1], pm[2], pm[3]);
      if constexpr (SR == LT_MAXTROPICAL)
        *reinterpret_cast<float4*>(ps_buf + warp * kCols + cg * 4) = make_float4(ps[0], ps[1], ps[2], ps[3]);
    }
    group_sync(grp);   // #1: partial maxima visible; every thread holds its tile slice in registers
    float shift = 0.f;
    if constexpr (NORM) {
      if (warp < 2) {                         // finalisers (warp 0) and the state-0 / logZ warp
        const float4 w0 = *reinterpret_cast<const float4*>(wmax + (t & 1) * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wmax + (t & 1) * 8 + 4);
        shift = norm_shift(fmaxf(fmaxf(fmaxf(w0.x, w0.y), fmaxf(w0.z, w0.w)),
                                 fmaxf(fmaxf(w1.x, w1.y), fmaxf(w1.z, w1.w))));
        off += (int)shift;
      }
    }

    if (gt == 0 && t + NS < nf) {               // the stage is free: refill it NS frames ahead
      const uint32_t bar = smem_u32(&bars[stage]);
      mbar_arrive_expect_tx(bar, kStageBytes);
      tma_load_2d(smem_u32(tiles) + stage * kStageBytes, &tmap, col0, (int)((bt0 + t + NS) * C), bar);
    }

    if constexpr (SR == LT_LOG) {
      // exact column maximum, then ONE ex2 per arc
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
      for (int j = 0; j < 4; ++j) {
        ps[j] += __shfl_xor_sync(0xffffffffu, ps[j], 8);
        ps[j] += __shfl_xor_sync(0xffffffffu, ps[j], 16);
      }
      if (lane < 8)
        *reinterpret_cast<float4*>(ps_buf + warp * kCols + cg * 4) = make_float4(ps[0], ps[1], ps[2], ps[3]);
      group_sync(grp);   // #2: partial sums visible
    }

    // advance the ring; non-finaliser warps pull the next tile into registers now
    if (++stage == NS) { stage = 0; parity ^= 1; }
    if (!is_fin && t + 1 < nf) {
      mbar_wait(smem_u32(&bars[stage]), parity);
      const float* tile = tiles + (size_t)stage * (kStageBytes / 4);
#pragma unroll
      for (int i = 0; i < RPT; ++i)
        x[i] = *reinterpret_cast<const float4*>(tile + (size_t)(r0 + i) * kCols + cg * 4);
    }

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
      const float ab = S::times(cur[q], to_dom<SR>(cblank));
      const float xt = S::times(cur[V], to_dom<SR>(ctail));   // source row V (not in the TMA box)
      float v;
      if constexpr (SR == LT_LOG) {
        lse2_merge(m, s, xt, xt == neg_inf() ? 0.f : 1.f);
        v = log2_add_exp2(ab, msafe(m) + __log2f(s)) - shift;
      } else if constexpr (SR == LT_MAXTROPICAL) {
        int am = __float_as_int(s);
        if (xt > m) { m = xt; am = V; }
        const bool take_blank = ab >= m;               // semirings.py:363
        v = take_blank ? ab : m;
        if (p.backptr) p.backptr[(bt0 + t) * C + q] = take_blank ? (int16_t)-1 : (int16_t)am;
      } else {
        v = ab + (m + xt);
      }
      xchg_store(nxt, q, v, &xbar[(t + 1) & 1], CL);
      if (t + 1 < nf) {
        mbar_wait(smem_u32(&bars[stage]), parity);
        const float* tile = tiles + (size_t)stage * (kStageBytes / 4);
#pragma unroll
        for (int i = 0; i < RPT; ++i)
          x[i] = *reinterpret_cast<const float4*>(tile + (size_t)(r0 + i) * kCols + cg * 4);
      }
    } else if (is_q0) {
      const float v = S::times(cur[0], to_dom<SR>(cblank)) - shift;
      if constexpr (SR == LT_MAXTROPICAL) { if (p.backptr) p.backptr[(bt0 + t) * C] = (int16_t)-1; }
      xchg_store(nxt, 0, v, &xbar[(t + 1) & 1], CL);
    }
  }
  float* cur = alpha_buf + (nf & 1) * CP;
  if (nf > 0) mbar_wait(smem_u32(&xbar[nf & 1]), ((nf - 1) >> 1) & 1);

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
        // logZ = (off_T + r) ln 2, rounded once from double; the pair (off_T, r) is what the
        // backward kernel uses (dist[b] alone has lost r's low bits at |logZ| ~ 1e3)
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
// posterior mass e of an arc times its value v; pruned arcs (e == 0) may carry v = -inf
__device__ __forceinline__ float pv(float e, float v) { return e > 0.f ? e * v : 0.f; }

template <int SR, int V, bool SPLIT, bool NORM, bool EXPECT = false>
__global__ void __launch_bounds__(kGroupThreads, 2)
lattice_backward_fast2(const Fast2BwdParams p) {
  static_assert(!NORM || SR == LT_LOG, "renormalisation is a Log-semiring feature");
  static_assert(!EXPECT || (SR == LT_LOG && !SPLIT), "expectations: Log semiring, no gradient");
  constexpr int G = 1;
  using S = Sr<SR>;
  constexpr int CL = V / kCols;
  constexpr int C = V + 1;
  constexpr int CH = V / 32;                    // float4 chunks per lane (8 lanes per row)
  constexpr int kRows = kCols;                  // rows per CTA (+ tail row V on the last rank)
  constexpr uint32_t kSlabBytes = kRows * V * 4;
  constexpr uint32_t kStageBytes = kSlabBytes + V * 4;
  constexpr int BP = ((C + 3 + 3) & ~3) + 4;    // beta buffer: entry q at index 3 + q
  extern __shared__ __align__(128) unsigned char smem2[];
  const int NS = p.stages;

  const int tid = threadIdx.x;
  const int grp = G == 1 ? 0 : tid >> 8;
  const int gt = tid & (kGroupThreads - 1);
  const int lane = gt & 31, warp = gt >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool last_rank = rank == CL - 1;
  const int cluster_id = blockIdx.x / CL;
  __shared__ int b_slot;
  const int b = utterance_of_cluster(cluster_id, p.num_frames, p.B, p.T, &b_slot);
  const bool active = b < p.B;

  float* tiles = reinterpret_cast<float*>(smem2) + (size_t)grp * NS * (kStageBytes / 4);
  float* small = reinterpret_cast<float*>(smem2 + (size_t)G * NS * kStageBytes);
  constexpr int kSmallFloats = 2 * BP + 2 * 16;
  small += (size_t)grp * kSmallFloats;
  float* beta_buf = small;
  uint64_t* bars = reinterpret_cast<uint64_t*>(beta_buf + 2 * BP);
  uint64_t* xbar = bars + NS;

  const int sub = lane >> 3, sl = lane & 7;     // row within the warp, lane within the row
  const int row = warp * 4 + sub;               // local row 0..31
  const int prow = rank * kRows + row;          // source state
  const int nf = active ? max(0, min(p.num_frames[b], p.T)) : 0;
  const size_t bt0 = (size_t)(active ? b : 0) * p.T;
  const float logz = active ? p.dist[b] : 0.f;
  // Log: everything on chip is in log2 units.  NORM: logZ = off_T + r, and the offsets cancel
  // against those of alpha~ / beta~ up to the per-frame shift d_t (see the file comment).
  const int32_t* an = NORM ? p.alpha_norm + (size_t)(active ? b : 0) * (p.T + 3) : nullptr;
  const float logz2 = NORM ? __int_as_float(an[p.T + 1]) : logz * kLog2e;
  const float gscale = (active && p.grad_dist) ? p.grad_dist[b] : 1.f;
  const bool scale_ok = (SR != LT_LOG) || is_finite(logz);
  constexpr bool split = SPLIT;      // compile-time: the fp32 kernel is unchanged
  const uint32_t stage_tx = last_rank ? kStageBytes : kSlabBytes;

  if (gt == 0) {
    for (int s = 0; s < NS + 2; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = gt; c < 2 * BP; c += kGroupThreads) beta_buf[c] = S::one();   // lattices.py:789-790
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

  double eacc = 0.0;                            // EXPECT: this thread's sum of posterior * value
  // padding frames: zero gradients (lattices.py:775-779)
  if (active && !EXPECT) {
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

  // row owners (lane sl == 0) prefetch alpha_t[p], blank_t[p] one frame ahead;
  // warp 0 lane 0 of the last rank also owns the tail row V.
  const bool owner = sl == 0;
  const bool tail_owner = last_rank && warp == 0 && lane == 0;
  float n_alpha = 0.f, n_blank = 0.f, n_talpha = 0.f, n_tblank = 0.f;
  int n_off = 0, c_off1 = 0;                    // NORM: off_t (prefetched), off_{t+1}
  if (nf > 0) {
    const size_t o = (bt0 + nf - 1) * C;
    if (owner) { n_alpha = p.alphas[o + prow]; n_blank = ldg_stream(p.blank + o + prow); }
    if (tail_owner) { n_talpha = p.alphas[o + V]; n_tblank = ldg_stream(p.blank + o + V); }
    if constexpr (NORM) { n_off = an[nf - 1]; c_off1 = an[nf]; }
  }

  int stage = 0;
  uint32_t parity = 0;
  for (int it = 0; it < nf; ++it) {
    const int t = nf - 1 - it;
    float* beta = beta_buf + (it & 1) * BP;          // beta_{t+1}; entry q at beta[3 + q]
    float* nxt = beta_buf + ((it + 1) & 1) * BP;
    if (it > 0) {
      // every row of the previous frame has been reduced cluster-wide: beta is
      // complete and the tile stage of iteration it-1 is free for the next TMA
      mbar_wait(smem_u32(&xbar[it & 1]), ((it - 1) >> 1) & 1);
      if (gt == 0 && it - 1 + NS < nf) issue(it - 1 + NS);
    }
    if (gt == 0) mbar_arrive_expect_tx(smem_u32(&xbar[(it + 1) & 1]), C * 4);
    const float c_alpha = n_alpha, c_blank = n_blank, c_talpha = n_talpha, c_tblank = n_tblank;
    // NORM: d_t = off_{t+1} - off_t; `zref` replaces logZ in every posterior exponent
    const float shift = NORM ? (float)(c_off1 - n_off) : 0.f;
    const float zref = logz2 + shift;
    if constexpr (NORM) c_off1 = n_off;
    if (t > 0) {
      const size_t o = (bt0 + t - 1) * C;
      if (owner) { n_alpha = p.alphas[o + prow]; n_blank = ldg_stream(p.blank + o + prow); }
      if (tail_owner) { n_talpha = p.alphas[o + V]; n_tblank = ldg_stream(p.blank + o + V); }
      if constexpr (NORM) n_off = an[t - 1];
    }
    mbar_wait(smem_u32(&bars[stage]), parity);
    const float* tile = tiles + (size_t)stage * (kStageBytes / 4);
    if (++stage == NS) { stage = 0; parity ^= 1; }
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
        const float rs = scale_ok ? gscale * ex2(alpha_p + ms - zref) : 0.f;
        float s = 0.f, fa = 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          float4 e;
          e.x = ex2(x[i].x - ms); e.y = ex2(x[i].y - ms);
          e.z = ex2(x[i].z - ms); e.w = ex2(x[i].w - ms);
          s += (e.x + e.y) + (e.z + e.w);
          if constexpr (EXPECT) {
            const int c4 = (sl + 8 * i) * 4;
            const float4 vv = p.value_lexical
                ? ldg_stream4(p.value_lexical + (bt0 + t) * (size_t)C * V + (size_t)prow * V + c4)
                : *reinterpret_cast<const float4*>(trow + c4);
            fa += (pv(e.x, vv.x) + pv(e.y, vv.y)) + (pv(e.z, vv.z) + pv(e.w, vv.w));
          } else {
            store_grad4(grow, (sl + 8 * i) * 4, V, split, make_float4(e.x * rs, e.y * rs, e.z * rs, e.w * rs));
          }
        }
        if constexpr (EXPECT) { if (rs > 0.f) eacc += (double)(fa * rs); }
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
          store_grad4(grow, c4, V, split, make_float4(ga * bn.x, ga * bn.y, ga * bn.z, ga * bn.w));
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        rowsum = s;
      }
      float bnew = 0.f;
      if (owner) {
        const float bp = beta[3 + prow];
        const float bb = arc<SR>(c_blank, bp);
        if constexpr (EXPECT) {
          const float post = scale_ok ? ex2(alpha_p + bb - zref) : 0.f;
          eacc += (double)pv(post, p.value_blank ? p.value_blank[(bt0 + t) * C + prow] : c_blank);
        } else if constexpr (SR == LT_LOG) gb[prow] = scale_ok ? gscale * ex2(alpha_p + bb - zref) : 0.f;
        else gb[prow] = gscale * c_alpha * bp;
        bnew = SR == LT_LOG ? log2_add_exp2(bb, rowsum) - shift : bb + rowsum;
      }
      // lane sl of the row group sends the row's new beta to rank sl
      xchg_store_group8(nxt, 3 + prow, bnew, &xbar[(it + 1) & 1], CL, lane, true);
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
        const float rs = scale_ok ? gscale * ex2(alpha_p + ms - zref) : 0.f;
        float s = 0.f, fa = 0.f;
        for (int c4 = lane * 4; c4 < V; c4 += 128) {
          const float4 w = *reinterpret_cast<const float4*>(trow + c4);
          const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
          float4 e;
          e.x = ex2(arc<SR>(w.x, bn.x) - ms); e.y = ex2(arc<SR>(w.y, bn.y) - ms);
          e.z = ex2(arc<SR>(w.z, bn.z) - ms); e.w = ex2(arc<SR>(w.w, bn.w) - ms);
          s += (e.x + e.y) + (e.z + e.w);
          if constexpr (EXPECT) {
            const float4 vv = p.value_lexical
                ? ldg_stream4(p.value_lexical + (bt0 + t) * (size_t)C * V + (size_t)V * V + c4) : w;
            fa += (pv(e.x, vv.x) + pv(e.y, vv.y)) + (pv(e.z, vv.z) + pv(e.w, vv.w));
          } else {
            store_grad4(grow, c4, V, split, make_float4(e.x * rs, e.y * rs, e.z * rs, e.w * rs));
          }
        }
        if constexpr (EXPECT) { if (rs > 0.f) eacc += (double)(fa * rs); }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        rowsum = ms + __log2f(s);
      } else {
        float s = 0.f;
        const float ga = gscale * alpha_p;
        for (int c4 = lane * 4; c4 < V; c4 += 128) {
          const float4 w = *reinterpret_cast<const float4*>(trow + c4);
          const float4 bn = *reinterpret_cast<const float4*>(bnext + c4);
          s += (w.x * bn.x + w.y * bn.y) + (w.z * bn.z + w.w * bn.w);
          store_grad4(grow, c4, V, split, make_float4(ga * bn.x, ga * bn.y, ga * bn.z, ga * bn.w));
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        rowsum = s;
      }
      if (lane == 0) {
        const float bp = beta[3 + V];
        const float bb = arc<SR>(c_tblank, bp);
        if constexpr (EXPECT) {
          const float post = scale_ok ? ex2(alpha_p + bb - zref) : 0.f;
          eacc += (double)pv(post, p.value_blank ? p.value_blank[(bt0 + t) * C + V] : c_tblank);
        } else if constexpr (SR == LT_LOG) gb[V] = scale_ok ? gscale * ex2(alpha_p + bb - zref) : 0.f;
        else gb[V] = gscale * c_talpha * bp;
        xchg_store(nxt, 3 + V, SR == LT_LOG ? log2_add_exp2(bb, rowsum) - shift : bb + rowsum,
                   &xbar[(it + 1) & 1], CL);
      }
    }
  }
  float* beta = beta_buf + (nf & 1) * BP;
  if (nf > 0) mbar_wait(smem_u32(&xbar[nf & 1]), ((nf - 1) >> 1) & 1);
  if constexpr (EXPECT) {
    __shared__ double esum[kGroupWarps];
    for (int o = 16; o > 0; o >>= 1) eacc += __shfl_xor_sync(0xffffffffu, eacc, o);
    if (lane == 0) esum[warp] = eacc;
    __syncthreads();
    if (gt == 0 && active) {
      double tot = 0.0;
      for (int w = 0; w < kGroupWarps; ++w) tot += esum[w];
      p.expect_part[(size_t)b * CL + rank] = tot;
    }
  }
  if (active && p.beta_final) {
    // NORM: beta_0 = beta~_0 + (off_T - off_0)
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

bool lattice_fast2_supported(const NGram& g, int k, unsigned flags, const void* lexical) {
  if (flags & LT_FLAG_FORCE_GENERIC) return false;
  if ((flags >> LT_FLAG_CLUSTER_SHIFT) & 0xf) return false;   // explicit cluster size => generic
  if (k >= 1 || g.n != 1) return false;
  if (g.V % 64 != 0 || g.V > 256) return false;
  if (lexical && reinterpret_cast<uintptr_t>(lexical) % 16 != 0) return false;
  return true;
}

int lattice_forward_fast2_launch(int semiring, const NGram& g, const FwdParams& base,
                                 unsigned flags, cudaStream_t stream) {
  const int V = g.V, C = g.C, CL = V / kCols;
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
  const size_t small = sizeof(float) * (2 * CP + 4 * kGroupWarps * kCols + 2 * 16 + 16);
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
  const int grid = base.B * CL;
#define LT_FWD2V(SR, VV, NORM) \
  return launch_fast2(lattice_forward_fast2<SR, VV, NORM>, grid, kGroupThreads, smem, CL, stream, tmap, p);
#define LT_FWD2(SR, NORM)                  \
  switch (V) {                             \
    case 64: LT_FWD2V(SR, 64, NORM)        \
    case 128: LT_FWD2V(SR, 128, NORM)      \
    case 192: LT_FWD2V(SR, 192, NORM)      \
    default: LT_FWD2V(SR, 256, NORM)       \
  }
  if (semiring == LT_LOG && p.alpha_norm) { LT_FWD2(LT_LOG, true) }
  if (semiring == LT_LOG) { LT_FWD2(LT_LOG, false) }
  if (semiring == LT_MAXTROPICAL) { LT_FWD2(LT_MAXTROPICAL, false) }
  LT_FWD2(LT_REAL, false)
#undef LT_FWD2
#undef LT_FWD2V
}

int lattice_backward_fast2_launch(int semiring, const NGram& g, const BwdParams& base,
                                  unsigned flags, cudaStream_t stream) {
  const int V = g.V, C = g.C, CL = V / kCols;
  const size_t stage = (size_t)kCols * V * 4 + (size_t)V * 4;
  const int BP = ((C + 6) & ~3) + 4;
  const size_t small = sizeof(float) * (2 * BP + 2 * 16);
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
  const int grid = base.B * CL;
#define LT_BWD2K(SR, VV, SPLIT, NORM) \
  return launch_fast2(lattice_backward_fast2<SR, VV, SPLIT, NORM>, grid, kGroupThreads, smem, CL, stream, p);
#define LT_BWD2V(SR, VV, NORM)                        \
  if (p.split) { LT_BWD2K(SR, VV, true, NORM) }       \
  LT_BWD2K(SR, VV, false, NORM)
#define LT_BWD2(SR, NORM)                  \
  switch (V) {                             \
    case 64: LT_BWD2V(SR, 64, NORM)        \
    case 128: LT_BWD2V(SR, 128, NORM)      \
    case 192: LT_BWD2V(SR, 192, NORM)      \
    default: LT_BWD2V(SR, 256, NORM)       \
  }
  if (base.expect_part) {       // lt_lattice_expectation: Log only (checked by the caller)
    p.value_blank = base.value_blank; p.value_lexical = base.value_lexical;
    p.expect_part = base.expect_part;
#define LT_EXP2(VV, NORM) \
  return launch_fast2(lattice_backward_fast2<LT_LOG, VV, false, NORM, true>, grid, kGroupThreads, smem, CL, stream, p);
#define LT_EXP2V(NORM)                     \
  switch (V) {                             \
    case 64: LT_EXP2(64, NORM)             \
    case 128: LT_EXP2(128, NORM)           \
    case 192: LT_EXP2(192, NORM)           \
    default: LT_EXP2(256, NORM)            \
  }
    if (p.alpha_norm) { LT_EXP2V(true) }
    LT_EXP2V(false)
#undef LT_EXP2V
#undef LT_EXP2
  }
  if (semiring == LT_LOG && p.alpha_norm) { LT_BWD2(LT_LOG, true) }
  if (semiring == LT_LOG) { LT_BWD2(LT_LOG, false) }
  LT_BWD2(LT_REAL, false)
#undef LT_BWD2
#undef LT_BWD2V
#undef LT_BWD2K
}

}  // namespace lt
