// K2 for higher-order contexts (FullNGram context_size >= 2, FrameDependent, Log / Real):
// the backward (beta) recursion + arc posteriors as "one 8-lane group per source row",
// the companion of the thread-per-column forward in lattice_cols.cu.
//
// For a fixed source state p the V destinations next(p, .) are contiguous
// (contexts.py:190-205): A + ((p - Alow) * V mod N) + y for the full-order part, 1 + p*V + y
// for the low-order rows.  So a row of the frame is a contiguous stream of lexical[p, :]
// against a contiguous, 16-byte aligned window of beta'.
//   * cluster of CL CTAs per utterance, CTA r owns a contiguous range of source rows; a
//     producer warp streams them with bulk (TMA) copies of 64 rows into a ring, several
//     chunks ahead of the recursion (full / empty mbarriers);
//   * 8 lanes per row, conflict-free LDS.128 of the slab and of the beta window; ONE
//     exponential per arc serves both the row log-sum-exp (beta) and the arc posterior,
//     which is written straight from registers with coalesced streaming stores;
//   * beta lives in shared memory as a full replica per CTA (every CTA's windows wrap over all
//     N full-order states), double-buffered, all-gathered with st.async + mbarrier
//     complete_tx: no cluster barrier in the loop; two CTAs of 288 threads per SM.
// Log arithmetic in log2 units as in lattice_fast2.cu.  With the renormalised pair of the
// thread-per-column forward (`alpha_norm`: alphas hold alpha~_t, logZ = off_T + r) beta is kept as
// beta~_t = beta_t - (off_T - off_t): it follows the forward's own shifts d_t = off_{t+1} - off_t
// (subtracted where beta_t is published) and every posterior exponent becomes
// alpha~ + w + beta~ - (r + d_t), all O(10) instead of O(logZ).
//
// Reference semantics: alignments.py:300-318, lattices.py:775-779, contexts.py:232-256.
#include <cuda.h>

#include "common.cuh"
#include "fast_ptx.cuh"
#include "params.cuh"

namespace lt {

namespace {

using namespace fastptx;

constexpr int kRConsumers = 256;
constexpr int kRThreads = kRConsumers + 32;
constexpr int kRChunk = 64;                   // rows per ring stage
constexpr int kRPad = 3;                      // beta entry q at index kRPad + q (windows 16-B aligned)

__device__ __forceinline__ void mbar_arrive_r(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct RowsParams {
  NGram g;
  int B, T;
  int k;                                      // max_expansions (FrameLabelDependent) or -1
  int stages, cl, rpc;                        // ring depth, cluster size, rows per CTA
  const float* levels;                        // [B,T,k,C] (FrameLabelDependent)
  const float* blank;
  const float* lexical;
  const int32_t* num_frames;
  const float* alphas;
  const float* dist;
  const float* grad_dist;
  float* grad_blank;
  float* grad_lexical;
  float* beta_final;
  const int32_t* alpha_norm;                  // Log: offsets written by the renormalised forward, or nullptr
};

template <int SR, int CHL>      // CHL = V / 32 float4 chunks per lane
__global__ void __launch_bounds__(kRThreads, 2)
lattice_backward_rows(const RowsParams p) {
  using S = Sr<SR>;
  extern __shared__ __align__(128) unsigned char rsmem[];
  const NGram& g = p.g;
  const int C = g.C, V = g.V;
  const int NS = p.stages;
  const uint32_t stage_bytes = (uint32_t)kRChunk * V * 4;
  const int BP = (kRPad + C + 3 + 3) & ~3;

  float* tiles = reinterpret_cast<float*>(rsmem);
  float* beta_buf = reinterpret_cast<float*>(rsmem + (size_t)NS * stage_bytes);   // [2][BP]
  uint64_t* full = reinterpret_cast<uint64_t*>(beta_buf + 2 * BP);
  uint64_t* empty = full + NS;
  uint64_t* xbar = empty + NS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t CL = p.cl;
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / CL;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;
  const int p_lo = min(C, (int)rank * p.rpc), p_hi = min(C, p_lo + p.rpc);
  const int nrows = p_hi - p_lo;
  const int nchunks = (nrows + kRChunk - 1) / kRChunk;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), kRConsumers / 32);
    }
    mbar_init(smem_u32(&xbar[0]), 1);
    mbar_init(smem_u32(&xbar[1]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = tid; c < 2 * BP; c += kRThreads) beta_buf[c] = to_dom<SR>(S::one());   // lattices.py:789-790
  __syncthreads();
  cluster_sync_all();

  if (warp == kRConsumers / 32) {
    // ------------------------------------------------------------ producer warp
    if (lane == 0 && nrows > 0) {
      int stage = 0;
      uint32_t use = 0;
      for (int it = 0; it < nf; ++it) {
        const int t = nf - 1 - it;
        const float* frame = p.lexical + (bt0 + t) * (size_t)C * V;
        for (int ch = 0; ch < nchunks; ++ch) {
          const int r0 = p_lo + ch * kRChunk;
          const uint32_t bytes = (uint32_t)min(kRChunk, p_hi - r0) * V * 4;
          if (use > 0) mbar_wait(smem_u32(&empty[stage]), (use - 1) & 1);
          const uint32_t bar = smem_u32(&full[stage]);
          mbar_arrive_expect_tx(bar, bytes);
          bulk_load_1d(smem_u32(tiles) + stage * stage_bytes, frame + (size_t)r0 * V, bytes, bar);
          if (++stage == NS) { stage = 0; ++use; }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- consumers
    const int sub = lane >> 3, sl = lane & 7;       // row within the warp pass, lane within the row
    const int rloc = warp * 4 + sub;                // 0 .. 31: row inside a 32-row pass
    const float logz = p.dist[b];
    const bool norm = SR == LT_LOG && p.alpha_norm != nullptr;
    const int32_t* an = norm ? p.alpha_norm + (size_t)b * (p.T + 3) : nullptr;
    const float lz_res = norm ? __int_as_float(an[p.T + 1]) : ((SR == LT_LOG) ? logz * kLog2e : logz);
    int an_hi = norm ? an[nf] : 0, an_lo = (norm && nf > 0) ? an[nf - 1] : 0;
    const float gscale = p.grad_dist ? p.grad_dist[b] : 1.f;
    const bool scale_ok = (SR != LT_LOG) || is_finite(logz);
    const bool owner = sl == 0;

    // padding frames: zero gradients (lattices.py:775-779)
    for (int t = nf; t < p.T; ++t) {
      float4* gl = reinterpret_cast<float4*>(p.grad_lexical + ((bt0 + t) * (size_t)C + p_lo) * V);
      for (int i = tid; i < nrows * V / 4; i += kRConsumers)
        stg_stream4(reinterpret_cast<float*>(gl + i), make_float4(0, 0, 0, 0));
      for (int i = tid; i < nrows; i += kRConsumers) p.grad_blank[(bt0 + t) * C + p_lo + i] = 0.f;
    }

    // Per-thread row bookkeeping.  A lane group visits rows row0, row0 + 32, row0 + 64, ... of
    // every frame; the beta window of a full-order row advances by 32*V (mod N) per visit, so
    // the closed form of row_window (an integer modulo) is evaluated once per kernel.
    const int row0 = p_lo + rloc;
    long long wr0l = ((long long)(row0 - g.Alow) * V) % g.N;
    if (wr0l < 0) wr0l += g.N;
    const int wr0 = (int)wr0l;
    const int wstep = (32 * V) % g.N;

    // alpha_t[p], blank_t[p] of the two rows this owner lane handles in the NEXT chunk
    float n_alpha[2] = {0.f, 0.f}, n_blank[2] = {0.f, 0.f};
    auto prefetch = [&](int it, int ch) {
      if (!owner || it >= nf) return;
      const size_t o = (bt0 + (nf - 1 - it)) * C + row0 + ch * kRChunk;
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        if (row0 + ch * kRChunk + ps * 32 < p_hi) {
          n_alpha[ps] = p.alphas[o + ps * 32];
          n_blank[ps] = ldg_stream(p.blank + o + ps * 32);
        }
      }
    };
    if (nrows > 0) prefetch(0, 0);

    const uint32_t expect = (uint32_t)C * 4;
    int stage = 0;
    uint32_t use = 0;
    for (int it = 0; it < nf; ++it) {
      const int t = nf - 1 - it;
      float* beta = beta_buf + (it & 1) * BP;          // beta_{t+1}; entry q at beta[kRPad + q]
      float* nxt = beta_buf + ((it + 1) & 1) * BP;
      // d_t = off_{t+1} - off_t; the offset of the next frame is fetched a frame ahead
      const float dsh = (float)(an_hi - an_lo);
      const float logz2 = lz_res + dsh;
      an_hi = an_lo;
      if (norm && t > 0) an_lo = an[t - 1];
      if (it > 0) mbar_wait(smem_u32(&xbar[it & 1]), ((it - 1) >> 1) & 1);
      if (tid == 0) mbar_arrive_expect_tx(smem_u32(&xbar[(it + 1) & 1]), expect);
      float* gb = p.grad_blank + (bt0 + t) * C;
      float* grow = p.grad_lexical + ((bt0 + t) * (size_t)C + row0) * V + sl * 4;
      int prow = row0, wrel = wr0;

      for (int ch = 0; ch < nchunks; ++ch) {
        const float c_alpha[2] = {n_alpha[0], n_alpha[1]}, c_blank[2] = {n_blank[0], n_blank[1]};
        if (ch + 1 < nchunks) prefetch(it, ch + 1); else prefetch(it + 1, 0);
        mbar_wait(smem_u32(&full[stage]), use & 1);
        const float* tile = tiles + (size_t)stage * (stage_bytes / 4);
#pragma unroll
        for (int ps = 0; ps < 2; ++ps, prow += 32, grow += (size_t)32 * V) {
          const bool live = prow < p_hi;
          // a warp covers 4 consecutive rows; lanes of dead rows still take part in shuffles
          const float* trow = tile + (live ? (ps * 32 + rloc) * V : 0) + sl * 4;
          const int win = !live ? g.A : (prow < g.Alow ? g.off + prow * V : g.A + wrel);
          const float* bwin = beta + kRPad + win + sl * 4;
          wrel += wstep;
          if (wrel >= g.N) wrel -= g.N;
          float4 x[CHL];
#pragma unroll
          for (int i = 0; i < CHL; ++i) {
            const float4 w = *reinterpret_cast<const float4*>(trow + 32 * i);
            const float4 bn = *reinterpret_cast<const float4*>(bwin + 32 * i);
            x[i] = make_float4(arc<SR>(w.x, bn.x), arc<SR>(w.y, bn.y), arc<SR>(w.z, bn.z),
                               arc<SR>(w.w, bn.w));
          }
          const float alpha_raw = __shfl_sync(0xffffffffu, c_alpha[ps], lane & ~7);
          const float alpha_p = to_dom<SR>(alpha_raw);
          float rowsum;
          if constexpr (SR == LT_LOG) {
            float m = neg_inf();
#pragma unroll
            for (int i = 0; i < CHL; ++i) m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            const float ms = msafe(m);
            const float rs = scale_ok ? gscale * ex2(alpha_p + ms - logz2) : 0.f;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < CHL; ++i) {
              float4 e;
              e.x = ex2(x[i].x - ms); e.y = ex2(x[i].y - ms);
              e.z = ex2(x[i].z - ms); e.w = ex2(x[i].w - ms);
              s += (e.x + e.y) + (e.z + e.w);
              if (live) stg_stream4(grow + 32 * i, make_float4(e.x * rs, e.y * rs, e.z * rs, e.w * rs));
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            rowsum = ms + __log2f(s);
          } else {
            float s = 0.f;
            const float ga = gscale * alpha_p;
#pragma unroll
            for (int i = 0; i < CHL; ++i) {
              const float4 bn = *reinterpret_cast<const float4*>(bwin + 32 * i);
              s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
              if (live) stg_stream4(grow + 32 * i, make_float4(ga * bn.x, ga * bn.y, ga * bn.z, ga * bn.w));
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            rowsum = s;
          }
          float bnew = 0.f;
          if (owner && live) {
            const float bp = beta[kRPad + prow];
            const float bb = arc<SR>(c_blank[ps], bp);
            if constexpr (SR == LT_LOG) gb[prow] = scale_ok ? gscale * ex2(alpha_p + bb - logz2) : 0.f;
            else gb[prow] = gscale * alpha_raw * bp;
            bnew = SR == LT_LOG ? log2_add_exp2(bb, rowsum) - dsh : bb + rowsum;
          }
          // lane sl of the row group sends the row's new beta to rank sl
          xchg_store_group8(nxt, kRPad + prow, bnew, &xbar[(it + 1) & 1], CL, lane, live);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_r(smem_u32(&empty[stage]));
        if (++stage == NS) { stage = 0; ++use; }
      }
    }
    float* beta = beta_buf + (nf & 1) * BP;
    if (nf > 0) mbar_wait(smem_u32(&xbar[nf & 1]), ((nf - 1) >> 1) & 1);
    if (p.beta_final)
      for (int i = tid; i < nrows; i += kRConsumers)
        p.beta_final[(size_t)b * C + p_lo + i] =
            norm ? (float)(((double)beta[kRPad + p_lo + i] + (double)an[p.T]) * 0.6931471805599453)
                 : from_dom<SR>(beta[kRPad + p_lo + i]);
  }
  __syncthreads();
  cluster_sync_all();
}

// ---------------------------------------------------------------------------------------
// FrameLabelDependent(k) variant (alignments.py:378-418): k chained row passes per frame.
//   nb_k = blank (x) beta';  for j = k-1 .. 0:
//     rowsum_j[p] = (+)_y lex[p,y] (x) nb_{j+1}[next(p,y)]
//     grad_lex[p,y] (+)= g * exp(src_j[p] + lex[p,y] + nb_{j+1}[next] - logZ),  src_0 = alpha, src_j = last_j
//     nb_j[p] = blank[p] (x) beta'[p] (+) rowsum_j[p]          (all-gathered, one exchange per pass)
//   beta_t = nb_0;  grad_blank[q] = g * sum_{i<=k} exp(src_i[q] + blank[q] + beta'[q] - logZ)
// Three state vectors (beta', and a pair for the nb_j); the frame is streamed k times through the
// same ring (the re-reads hit L2).  Exchange e uses barrier e & 1, phase (e >> 1) & 1.
// k = 2 (RECOMP): the first pass only produces nb_1 and the second one writes the gradient ONCE,
// recomputing the level-1 posterior of every arc from nb_2 (still in its buffer: beta_t is
// gathered into beta's own buffer, whose entries a CTA only reads for its own rows) -- one more
// exponential per arc instead of writing, re-reading and re-writing [B,T,C,V] through HBM.
// k > 2: passes after the first accumulate into grad_lexical.
template <int SR, int CHL>
__global__ void __launch_bounds__(kRThreads, 2)
lattice_backward_rows_fld(const RowsParams p) {
  using S = Sr<SR>;
  extern __shared__ __align__(128) unsigned char rsmem[];
  const NGram& g = p.g;
  const int C = g.C, V = g.V, K = p.k;
  const int NS = p.stages;
  const uint32_t stage_bytes = (uint32_t)kRChunk * V * 4;
  const int BP = (kRPad + C + 3 + 3) & ~3;

  float* tiles = reinterpret_cast<float*>(rsmem);
  float* vec = reinterpret_cast<float*>(rsmem + (size_t)NS * stage_bytes);        // [3][BP]
  uint64_t* full = reinterpret_cast<uint64_t*>(vec + 3 * BP);
  uint64_t* empty = full + NS;
  uint64_t* xbar = empty + NS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t CL = p.cl;
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / CL;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;
  const int p_lo = min(C, (int)rank * p.rpc), p_hi = min(C, p_lo + p.rpc);
  const int nrows = p_hi - p_lo;
  const int nchunks = (nrows + kRChunk - 1) / kRChunk;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), kRConsumers / 32);
    }
    mbar_init(smem_u32(&xbar[0]), 1);
    mbar_init(smem_u32(&xbar[1]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = tid; c < 3 * BP; c += kRThreads) vec[c] = to_dom<SR>(S::one());   // lattices.py:789-790
  __syncthreads();
  cluster_sync_all();

  if (warp == kRConsumers / 32) {
    // ------------------------------------------------------------ producer warp
    if (lane == 0 && nrows > 0) {
      int stage = 0;
      uint32_t use = 0;
      for (int it = 0; it < nf; ++it) {
        const int t = nf - 1 - it;
        const float* frame = p.lexical + (bt0 + t) * (size_t)C * V;
        for (int j = 0; j < K; ++j) {
          for (int ch = 0; ch < nchunks; ++ch) {
            const int r0 = p_lo + ch * kRChunk;
            const uint32_t bytes = (uint32_t)min(kRChunk, p_hi - r0) * V * 4;
            if (use > 0) mbar_wait(smem_u32(&empty[stage]), (use - 1) & 1);
            const uint32_t bar = smem_u32(&full[stage]);
            mbar_arrive_expect_tx(bar, bytes);
            bulk_load_1d(smem_u32(tiles) + stage * stage_bytes, frame + (size_t)r0 * V, bytes, bar);
            if (++stage == NS) { stage = 0; ++use; }
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- consumers
    const int sub = lane >> 3, sl = lane & 7;
    const int rloc = warp * 4 + sub;
    const float logz = p.dist[b];
    const bool norm = SR == LT_LOG && p.alpha_norm != nullptr;
    const int32_t* an = norm ? p.alpha_norm + (size_t)b * (p.T + 3) : nullptr;
    const float lz_res = norm ? __int_as_float(an[p.T + 1]) : ((SR == LT_LOG) ? logz * kLog2e : logz);
    int an_hi = norm ? an[nf] : 0, an_lo = (norm && nf > 0) ? an[nf - 1] : 0;
    const float gscale = p.grad_dist ? p.grad_dist[b] : 1.f;
    const bool scale_ok = (SR != LT_LOG) || is_finite(logz);
    const bool owner = sl == 0;

    for (int t = nf; t < p.T; ++t) {                 // padding frames: zero gradients
      float4* gl = reinterpret_cast<float4*>(p.grad_lexical + ((bt0 + t) * (size_t)C + p_lo) * V);
      for (int i = tid; i < nrows * V / 4; i += kRConsumers)
        stg_stream4(reinterpret_cast<float*>(gl + i), make_float4(0, 0, 0, 0));
      for (int i = tid; i < nrows; i += kRConsumers) p.grad_blank[(bt0 + t) * C + p_lo + i] = 0.f;
    }

    const int row0 = p_lo + rloc;
    long long wr0l = ((long long)(row0 - g.Alow) * V) % g.N;
    if (wr0l < 0) wr0l += g.N;
    const int wr0 = (int)wr0l;
    const int wstep = (32 * V) % g.N;

    // src_j[p] and blank[p] of the two rows this owner lane handles in the NEXT chunk; the
    // position runs over (frame, pass j = K-1 .. 0, chunk)
    struct Pos { int it, j, ch; };
    auto advance = [&](Pos q) {
      if (++q.ch == nchunks) { q.ch = 0; if (--q.j < 0) { q.j = K - 1; ++q.it; } }
      return q;
    };
    const bool recomp = K == 2;
    float n_src[2] = {0.f, 0.f}, n_blank[2] = {0.f, 0.f}, n_lv1[2] = {0.f, 0.f};
    auto prefetch = [&](const Pos& q) {
      if (!owner || q.it >= nf) return;
      const size_t bt = bt0 + (nf - 1 - q.it);
      const float* srcv = q.j == 0 ? p.alphas + bt * C : p.levels + (bt * K + (q.j - 1)) * C;
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        const int row = row0 + q.ch * kRChunk + ps * 32;
        if (row < p_hi) {
          n_src[ps] = srcv[row];
          n_blank[ps] = ldg_stream(p.blank + bt * C + row);
          if (recomp && q.j == 0) n_lv1[ps] = p.levels[(bt * K) * C + row];
        }
      }
    };
    Pos pos = {0, K - 1, 0};
    if (nrows > 0) prefetch(pos);

    const uint32_t expect = (uint32_t)C * 4;
    float* bp = vec;                 // beta_{t+1}
    float* f1 = vec + BP;
    float* f2 = vec + 2 * BP;
    long long e = 0;                 // exchange counter
    int stage = 0;
    uint32_t use = 0;
    for (int it = 0; it < nf; ++it) {
      const int t = nf - 1 - it;
      const size_t bt = bt0 + t;
      const float dsh = (float)(an_hi - an_lo);        // d_t = off_{t+1} - off_t
      const float logz2 = lz_res + dsh;
      an_hi = an_lo;
      if (norm && t > 0) an_lo = an[t - 1];
      if (e > 0) mbar_wait(smem_u32(&xbar[(e - 1) & 1]), (uint32_t)(((e - 1) >> 1) & 1));
      // nb_K = blank (x) beta' for ALL states (every CTA keeps a full replica)
      {
        const float* bl = p.blank + bt * C;
        for (int c = tid; c < C; c += kRConsumers) f1[kRPad + c] = arc<SR>(ldg_stream(bl + c), bp[kRPad + c]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kRConsumers) : "memory");
      // blank marginals of the rows this CTA owns (alignments.py:398-403)
      for (int q = p_lo + tid; q < p_hi; q += kRConsumers) {
        float acc = 0.f;
        if constexpr (SR == LT_LOG) {
          if (scale_ok) {
            const float base = f1[kRPad + q] - logz2;
            acc = ex2(fmaf(p.alphas[bt * C + q], kLog2e, base));
            for (int i = 0; i < K; ++i) acc += ex2(fmaf(p.levels[(bt * K + i) * C + q], kLog2e, base));
            acc *= gscale;
          }
        } else {
          acc = p.alphas[bt * C + q];
          for (int i = 0; i < K; ++i) acc += p.levels[(bt * K + i) * C + q];
          acc *= gscale * f1[kRPad + q];
        }
        p.grad_blank[bt * C + q] = acc;
      }
      float* src = f1;
      float* dst = f2;
      for (int j = K - 1; j >= 0; --j) {
        if (tid == 0) mbar_arrive_expect_tx(smem_u32(&xbar[e & 1]), expect);
        float* grow = p.grad_lexical + (bt * (size_t)C + row0) * V + sl * 4;
        int prow = row0, wrel = wr0;
        const bool accumulate = !recomp && j != K - 1;
        const bool store = !recomp || j == 0;        // RECOMP: the gradient is written in pass 0
        const bool redo1 = recomp && j == 0;         // ... including the level-1 posteriors
        if (redo1) dst = bp;                         // beta_t goes straight into beta's buffer
        for (int ch = 0; ch < nchunks; ++ch) {
          const float c_src[2] = {n_src[0], n_src[1]}, c_blank[2] = {n_blank[0], n_blank[1]};
          const float c_lv1[2] = {n_lv1[0], n_lv1[1]};
          pos = advance(pos);
          prefetch(pos);
          mbar_wait(smem_u32(&full[stage]), use & 1);
          const float* tile = tiles + (size_t)stage * (stage_bytes / 4);
#pragma unroll
          for (int ps = 0; ps < 2; ++ps, prow += 32, grow += (size_t)32 * V) {
            const bool live = prow < p_hi;
            const float* trow = tile + (live ? (ps * 32 + rloc) * V : 0) + sl * 4;
            const int win = !live ? g.A : (prow < g.Alow ? g.off + prow * V : g.A + wrel);
            const float* bwin = src + kRPad + win + sl * 4;
            wrel += wstep;
            if (wrel >= g.N) wrel -= g.N;
            float4 x[CHL];
#pragma unroll
            for (int i = 0; i < CHL; ++i) {
              const float4 w = *reinterpret_cast<const float4*>(trow + 32 * i);
              const float4 bn = *reinterpret_cast<const float4*>(bwin + 32 * i);
              x[i] = make_float4(arc<SR>(w.x, bn.x), arc<SR>(w.y, bn.y), arc<SR>(w.z, bn.z),
                                 arc<SR>(w.w, bn.w));
            }
            const float src_raw = __shfl_sync(0xffffffffu, c_src[ps], lane & ~7);
            const float src_p = to_dom<SR>(src_raw);
            const float lv1_p = to_dom<SR>(__shfl_sync(0xffffffffu, c_lv1[ps], lane & ~7));
            const float* bbwin = f1 + kRPad + win + sl * 4;      // nb_K window (RECOMP)
            float rowsum;
            if constexpr (SR == LT_LOG) {
              float m = neg_inf();
#pragma unroll
              for (int i = 0; i < CHL; ++i) m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
              m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
              m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
              m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
              const float ms = msafe(m);
              const float rs = scale_ok ? gscale * ex2(src_p + ms - logz2) : 0.f;
              float s = 0.f;
#pragma unroll
              for (int i = 0; i < CHL; ++i) {
                float4 ev;
                ev.x = ex2(x[i].x - ms); ev.y = ex2(x[i].y - ms);
                ev.z = ex2(x[i].z - ms); ev.w = ex2(x[i].w - ms);
                s += (ev.x + ev.y) + (ev.z + ev.w);
                if (live && store) {
                  float4 gv = make_float4(ev.x * rs, ev.y * rs, ev.z * rs, ev.w * rs);
                  if (accumulate) {
                    const float4 old = *reinterpret_cast<const float4*>(grow + 32 * i);
                    gv.x += old.x; gv.y += old.y; gv.z += old.z; gv.w += old.w;
                  }
                  if (redo1 && scale_ok) {
                    // level 1: exp(last_1[p] + w + nb_2[q] - logZ), the arc's weight re-read
                    const float4 w = *reinterpret_cast<const float4*>(trow + 32 * i);
                    const float4 b2 = *reinterpret_cast<const float4*>(bbwin + 32 * i);
                    const float o = lv1_p - logz2;
                    gv.x = fmaf(gscale, ex2(arc<SR>(w.x, b2.x) + o), gv.x);
                    gv.y = fmaf(gscale, ex2(arc<SR>(w.y, b2.y) + o), gv.y);
                    gv.z = fmaf(gscale, ex2(arc<SR>(w.z, b2.z) + o), gv.z);
                    gv.w = fmaf(gscale, ex2(arc<SR>(w.w, b2.w) + o), gv.w);
                  }
                  stg_stream4(grow + 32 * i, gv);
                }
              }
              s += __shfl_xor_sync(0xffffffffu, s, 1);
              s += __shfl_xor_sync(0xffffffffu, s, 2);
              s += __shfl_xor_sync(0xffffffffu, s, 4);
              rowsum = ms + __log2f(s);
            } else {
              float s = 0.f;
              const float ga = gscale * src_p;
#pragma unroll
              for (int i = 0; i < CHL; ++i) {
                const float4 bn = *reinterpret_cast<const float4*>(bwin + 32 * i);
                s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
                if (live && store) {
                  float4 gv = make_float4(ga * bn.x, ga * bn.y, ga * bn.z, ga * bn.w);
                  if (accumulate) {
                    const float4 old = *reinterpret_cast<const float4*>(grow + 32 * i);
                    gv.x += old.x; gv.y += old.y; gv.z += old.z; gv.w += old.w;
                  }
                  if (redo1) {
                    const float4 b2 = *reinterpret_cast<const float4*>(bbwin + 32 * i);
                    const float g1 = gscale * lv1_p;
                    gv.x = fmaf(g1, b2.x, gv.x); gv.y = fmaf(g1, b2.y, gv.y);
                    gv.z = fmaf(g1, b2.z, gv.z); gv.w = fmaf(g1, b2.w, gv.w);
                  }
                  stg_stream4(grow + 32 * i, gv);
                }
              }
              s += __shfl_xor_sync(0xffffffffu, s, 1);
              s += __shfl_xor_sync(0xffffffffu, s, 2);
              s += __shfl_xor_sync(0xffffffffu, s, 4);
              rowsum = s;
            }
            float bnew = 0.f;
            if (owner && live) {
              const float bb = arc<SR>(c_blank[ps], bp[kRPad + prow]);       // blank (x) beta'
              // beta~_t = nb_0 - d_t: the shift is applied to the last pass only
              bnew = SR == LT_LOG ? log2_add_exp2(bb, rowsum) - (j == 0 ? dsh : 0.f) : bb + rowsum;
            }
            xchg_store_group8(dst, kRPad + prow, bnew, &xbar[e & 1], CL, lane, live);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_r(smem_u32(&empty[stage]));
          if (++stage == NS) { stage = 0; ++use; }
        }
        if (j > 0) {                                   // the next pass reads what this one produced
          mbar_wait(smem_u32(&xbar[e & 1]), (uint32_t)((e >> 1) & 1));
          float* tmp = src; src = dst; dst = tmp;
        }
        ++e;
      }
      // beta_t now (being) gathered in dst; the old beta' and the other vector become scratch
      // (RECOMP: dst is beta's own buffer, nothing rotates)
      if (!recomp) {
        float* old = bp;
        bp = dst;
        f1 = old;
        f2 = src;
      }
    }
    if (e > 0) mbar_wait(smem_u32(&xbar[(e - 1) & 1]), (uint32_t)(((e - 1) >> 1) & 1));
    if (p.beta_final)
      for (int i = tid; i < nrows; i += kRConsumers)
        p.beta_final[(size_t)b * C + p_lo + i] =
            norm ? (float)(((double)bp[kRPad + p_lo + i] + (double)an[p.T]) * 0.6931471805599453)
                 : from_dom<SR>(bp[kRPad + p_lo + i]);
  }
  __syncthreads();
  cluster_sync_all();
}

template <typename KernelT>
static int launch_rows(KernelT kernel, int grid, size_t smem, int cluster, cudaStream_t stream,
                       const RowsParams& p) {
  LT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kRThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LT_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  note_launch();
  return LT_OK;
}

static size_t rows_fixed_bytes(const NGram& g, bool fld) {
  const size_t BP = (kRPad + g.C + 3 + 3) & ~3;
  return sizeof(float) * (fld ? 3 : 2) * BP + 8 * (2 * 8 + 2) + 256;
}

}  // namespace

bool lattice_rows_supported(const NGram& g, int k, unsigned flags, const void* lexical,
                            const void* grad_lexical) {
  if (flags & LT_FLAG_FORCE_GENERIC) return false;
  if ((flags >> LT_FLAG_CLUSTER_SHIFT) & 0xf) return false;
  if (g.n < 2) return false;
  if (g.V != 32 && g.V != 64 && g.V != 96 && g.V != 128 && g.V != 256) return false;
  if (reinterpret_cast<uintptr_t>(lexical) % 16 != 0) return false;
  if (reinterpret_cast<uintptr_t>(grad_lexical) % 16 != 0) return false;
  return rows_fixed_bytes(g, k >= 1) + 2 * (size_t)kRChunk * g.V * 4 <= 112 * 1024;
}

int lattice_backward_rows_launch(int semiring, const NGram& g, const BwdParams& base, int sm_count,
                                 cudaStream_t stream) {
  // widest cluster that still gives every CTA a few chunks of rows
  int cl = 8;
  while (cl > 1 && g.C / cl < 2 * kRChunk) cl >>= 1;
  const size_t budget = 112 * 1024;
  const bool fld = base.k >= 1;
  const size_t fixed = rows_fixed_bytes(g, fld);
  const size_t stage = (size_t)kRChunk * g.V * 4;
  int stages = (int)((budget - fixed) / stage);
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("rows path: not enough shared memory"); return LT_ERR_UNSUPPORTED; }
  const size_t smem = stage * stages + fixed;
  RowsParams p = {};
  p.g = g; p.B = base.B; p.T = base.T; p.stages = stages; p.cl = cl;
  p.k = base.k; p.levels = base.levels;
  p.rpc = (g.C + cl - 1) / cl;
  p.blank = base.blank; p.lexical = base.lexical; p.num_frames = base.num_frames;
  p.alphas = base.alphas; p.dist = base.dist; p.grad_dist = base.grad_dist;
  p.grad_blank = base.grad_blank; p.grad_lexical = base.grad_lexical; p.beta_final = base.beta_final;
  p.alpha_norm = semiring == LT_LOG ? base.alpha_norm : nullptr;
  const int grid = base.B * cl;
  (void)sm_count;
#define LT_ROWS1(SR, N)                                                                     \
  return fld ? launch_rows(lattice_backward_rows_fld<SR, N>, grid, smem, cl, stream, p)     \
             : launch_rows(lattice_backward_rows<SR, N>, grid, smem, cl, stream, p);
#define LT_ROWS(SR)                  \
  switch (g.V / 32) {                \
    case 1: LT_ROWS1(SR, 1)          \
    case 2: LT_ROWS1(SR, 2)          \
    case 3: LT_ROWS1(SR, 3)          \
    case 4: LT_ROWS1(SR, 4)          \
    default: LT_ROWS1(SR, 8)         \
  }
  if (semiring == LT_LOG) { LT_ROWS(LT_LOG) }
  LT_ROWS(LT_REAL)
#undef LT_ROWS
#undef LT_ROWS1
}

}  // namespace lt
