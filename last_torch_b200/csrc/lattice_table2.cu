// Table-driven lattice kernels, second generation: a CLUSTER per utterance.
//
// contexts.NextStateTable (/root/reference/last_torch/contexts.py:266-324) with the
// FrameDependent alignment lattice (alignments.py:286-318); same results as lattice_table.cu
// (which stays the path for FrameLabelDependent and for shapes outside the limits below).
//
//   * a cluster of CL CTAs owns one utterance; CTA r owns the R = ceil(C / CL) SOURCE rows
//     [r R, r R + R) of every frame's [C, V] lexical weights: one contiguous slab, streamed
//     by one `cp.async.bulk` per frame into a ring of NS stages (full mbarriers), NS frames
//     ahead of the recursion -- HBM sees only long sequential reads, each byte once;
//   * forward: destination-major PULL restricted to the local slab.  The global CSR of
//     incoming arcs is ascending in p V + y inside a destination, so the arcs of destination
//     q that start in this CTA's rows are a contiguous piece of q's segment; the pieces are
//     found once by binary search and kept in shared memory as 16-bit slab offsets in ELL
//     order (a warp reads consecutive halfwords; arcs beyond the average in-degree + 1 stay in
//     global memory).  One thread per destination reduces its local arcs (no atomics in any
//     semiring, first arg-max in flat-arc order) and sends the partial VALUE -- one float --
//     to every CTA of the cluster with st.async; every CTA merges the CL partials of every
//     destination in rank order (= ascending arc order), so all CTAs hold the full new alpha
//     and the CTA that holds a MaxTropical winner writes its back-pointer;
//   * backward: source-major, one warp per local row, beta'[table[p, y]] a shared-memory
//     gather (the slab of the table is kept in shared memory as 16-bit states), ONE
//     exponential per arc for the row log-sum-exp and the arc posterior; the new beta of a
//     row goes to every CTA's buffer with st.async;
//   * no cluster barrier inside either loop: the st.async stores complete the transaction
//     count of the receiver's mbarrier (as in lattice_fast2.cu); partial / beta buffers and
//     their mbarriers alternate per frame, and a slab stage is refilled as soon as the frame
//     that used it is known to be finished.
#include <stdlib.h>

#include "common.cuh"
#include "fast_ptx.cuh"
#include "params.cuh"
#include "table_params.cuh"

namespace lt {

namespace {

using namespace fastptx;

constexpr int kT2BwdThreads = 288;  // backward: 36 rows per pass, two CTAs per SM
constexpr int kT2MaxQ = 4;          // destinations per thread in the forward kernel
constexpr int kT2MaxStages = 4;
constexpr int kT2MaxThreads = 512;  // forward: one thread per destination where C allows
constexpr int kT2MaxCluster = 8;

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int lo, int hi,
                                               int key) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
constexpr float kClampLow = -3.0e38f;

// 16 bytes to a peer CTA, completing 16 bytes of its mbarrier transaction count: a quarter of
// the remote transactions (and of the serialised complete_tx updates) of four scalar stores
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::
          "r"(remote_addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)),
      "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(remote_bar)
      : "memory");
}

// (+) over the arcs of one destination, four at a time without predicates.  fetch(i) = slab
// offset of the i-th local arc; its source row is offset / V (one IMAD.HI with a magic
// constant).  The lists are padded to a multiple of four with the offset ONE PAST the slab: that
// float of every ring stage holds the semiring zero (the bulk copies never touch it), so a
// padded arc contributes nothing whatever alpha it is paired with.  Log works in LOG2 units
// (src is alpha * log2 e): one FFMA per arc, a clamped running maximum (never -inf, so no
// special cases) and bare ex2; value = m + log2(s).
template <int SR, int NMAX, typename F>
__device__ __forceinline__ void reduce_piece(F fetch, int ne, const float* __restrict__ src,
                                             const float* __restrict__ slab, uint32_t magic,
                                             int base, float& m, float& s, int& arg) {
  using S = Sr<SR>;
#pragma unroll
  for (int i = 0; i < NMAX; i += 4) {
    if (i >= ne) break;
    float x[4];
    uint32_t off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      off[u] = fetch(i + u);
      const float w = slab[off[u]], a = src[__umulhi(off[u], magic)];
      x[u] = SR == LT_LOG ? fmaf(w, kLog2e, a) : S::times(a, w);
    }
    if constexpr (SR == LT_LOG) {
      const float mn = fmaxf(m, fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])));
      s = fmaf(s, ex2(m - mn), (ex2(x[0] - mn) + ex2(x[1] - mn)) + (ex2(x[2] - mn) + ex2(x[3] - mn)));
      m = mn;
    } else if constexpr (SR == LT_MAXTROPICAL) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (x[u] > m) { m = x[u]; arg = base + (int)off[u]; }   // ascending: first max
    } else {
      s += (x[0] + x[1]) + (x[2] + x[3]);
    }
  }
}

// ============================================================== forward ==
// Partial of one destination over the local slab, as ONE float: the (+)-value itself
// (Log: msafe(m) + log s).  The winning arc of a MaxTropical partial stays in a register of
// the thread that found it: every CTA merges the same CL values in rank order, so each CTA
// knows whether it holds the winner and writes the back-pointer itself.
template <int SR>
__global__ void __launch_bounds__(kT2MaxThreads)
table_forward2_kernel(const TableParams p, const int R, const int NS, const uint32_t magic,
                      const int ell) {
  using S = Sr<SR>;
  extern __shared__ __align__(128) unsigned char t2sm[];
  const int C = p.C, V = p.V, Cp = (C + 3) & ~3;
  const int tid = threadIdx.x, nth = blockDim.x;
  const uint32_t rank = cluster_ctarank(), CL = cluster_nctarank();
  const int b = blockIdx.x / CL;
  const int row0 = rank * R, nrows = min(R, C - row0);
  const int base = row0 * V, lim = (row0 + nrows) * V;
  const uint32_t slab_bytes = (uint32_t)nrows * V * 4;
  const size_t stage_floats = (size_t)R * V + 4;            // + the padding arc's weight

  float* slabs = reinterpret_cast<float*>(t2sm);
  unsigned char* ptr = t2sm + (size_t)NS * stage_floats * 4;
  float* alpha = reinterpret_cast<float*>(ptr); ptr += (size_t)2 * (Cp + 4) * 4;   // slots >= C: 0
  float* part = reinterpret_cast<float*>(ptr); ptr += (size_t)2 * kT2MaxCluster * Cp * 4;   // [2][CL][Cp]
  int* seg_g = reinterpret_cast<int*>(ptr); ptr += (size_t)Cp * 4;     // global start of the piece
  int* seg_n = reinterpret_cast<int*>(ptr); ptr += (size_t)Cp * 4;     // arcs in the piece
  uint64_t* bars = reinterpret_cast<uint64_t*>(ptr); ptr += (kT2MaxStages + 2) * 8;
  uint64_t* xbar = bars + kT2MaxStages;
  uint16_t* arcs = reinterpret_cast<uint16_t*>(ptr);        // ELL: arcs[i * Cp + q], i < ell; nrows * V = padding

  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(smem_u32(&bars[s]), 1);
    mbar_init(smem_u32(&xbar[0]), 1);
    mbar_init(smem_u32(&xbar[1]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  // local piece of every destination's arc segment; its first `ell` arcs as 16-bit slab
  // offsets in ELL order (a warp reads consecutive halfwords), the rest stays in global memory
  for (int q = tid; q < C; q += nth) {
    const int lo = p.in_offsets[q], hi = p.in_offsets[q + 1];
    const int a0 = lower_bound_i32(p.in_arcs, lo, hi, base);
    const int a1 = lower_bound_i32(p.in_arcs, a0, hi, lim);
    seg_g[q] = a0;
    seg_n[q] = a1 - a0;
    const int ne = min(a1 - a0, ell);
    for (int i = 0; i < ne; ++i) arcs[(size_t)i * Cp + q] = (uint16_t)(p.in_arcs[a0 + i] - base);
    for (int i = ne; i < ell; ++i) arcs[(size_t)i * Cp + q] = (uint16_t)(nrows * V);
  }
  if (tid < NS) slabs[(size_t)tid * stage_floats + (size_t)nrows * V] = S::zero();   // natural units
  for (int c = C + tid; c < Cp + 4; c += nth) { alpha[c] = 0.f; alpha[(Cp + 4) + c] = 0.f; }
  for (int c = tid; c < C; c += nth)
    alpha[c] = to_dom<SR>(p.alpha_init ? p.alpha_init[(size_t)b * C + c]
                                       : (c == 0 ? S::one() : S::zero()));
  __syncthreads();
  cluster_sync_all();

  auto issue = [&](int t) {
    const int s = t % NS;
    const uint32_t bar = smem_u32(&bars[s]);
    mbar_arrive_expect_tx(bar, slab_bytes);
    bulk_load_1d(smem_u32(slabs + (size_t)s * stage_floats),
                 p.lexical + (bt0 + t) * (size_t)C * V + base, slab_bytes, bar);
  };
  if (tid == 0)
    for (int t = 0; t < NS && t < nf; ++t) issue(t);

  float* cur = alpha;
  float* nxt = alpha + (Cp + 4);
  bool mine[kT2MaxQ];                  // this CTA writes the outputs of destination q
#pragma unroll
  for (int j = 0; j < kT2MaxQ; ++j) mine[j] = (uint32_t)(tid + j * nth) % CL == rank;
  int stage = 0;
  uint32_t parity = 0;
  for (int t = 0; t < nf; ++t) {
    float* pbuf = part + (size_t)(t & 1) * kT2MaxCluster * Cp;
    uint64_t* xb = &xbar[t & 1];
    if (tid == 0) {
      // every thread passed the barrier that ended frame t-1: its stage is free
      if (t > 0 && t - 1 + NS < nf) issue(t - 1 + NS);
      mbar_arrive_expect_tx(smem_u32(xb), (uint32_t)Cp * CL * 4);
    }
    float bl[kT2MaxQ];
    int warc[kT2MaxQ];
#pragma unroll
    for (int j = 0; j < kT2MaxQ; ++j) {
      const int q = tid + j * nth;
      bl[j] = q < C ? ldg_stream(p.blank + (bt0 + t) * C + q) : 0.f;
      warc[j] = 0;
    }
    mbar_wait(smem_u32(&bars[stage]), parity);
    const float* slab = slabs + (size_t)stage * stage_floats;
    if (++stage == NS) { stage = 0; parity ^= 1; }
#pragma unroll
    for (int j = 0; j < kT2MaxQ; ++j) {
      const int q = tid + j * nth;
      if (q - (tid & 31) >= C) break;                  // warp-uniform
      float pv = S::zero();
      if (q < C) {
      if (p.alphas && mine[j]) p.alphas[(bt0 + t) * C + q] = from_dom<SR>(cur[q]);
      const int n = seg_n[q];
      const int ne = min(n, ell);
      float m = SR == LT_LOG ? kClampLow : neg_inf(), sum = 0.f;
      int arg = 0;
      {
        const uint16_t* al = arcs + q;
        for (int i0 = 0; i0 < ne; i0 += 32)
          reduce_piece<SR, 32>([&](int i) { return (uint32_t)al[(size_t)(i0 + i) * Cp]; },
                               ne - i0, cur + row0, slab, magic, base, m, sum, arg);
      }
      if (n > ell) {                                   // overflow of a high in-degree state
        const int32_t* ga = p.in_arcs + seg_g[q];
        for (int i = ell; i < n; ++i) {
          const uint32_t a = (uint32_t)(ga[i] - base);
          const float w = slab[a], av = cur[row0 + a / (uint32_t)V];
          if constexpr (SR == LT_LOG) {
            const float x = fmaf(w, kLog2e, av), mn = fmaxf(m, x);
            sum = fmaf(sum, ex2(m - mn), ex2(x - mn));
            m = mn;
          } else if constexpr (SR == LT_MAXTROPICAL) {
            const float x = av + w;
            if (x > m) { m = x; arg = base + (int)a; }
          } else {
            sum = fmaf(av, w, sum);
          }
        }
      }
      warc[j] = arg;
      pv = SR == LT_LOG ? m + __log2f(sum) : (SR == LT_MAXTROPICAL ? m : sum);
      }
      // all-to-all, four destinations per store: lane k of a quad sends the quad's 16 bytes to
      // ranks k, k + 4: slot [my rank][q0..q0+3] of their buffers
      const int lane = tid & 31, qb = lane & ~3;
      float4 quad;
      quad.x = __shfl_sync(0xffffffffu, pv, qb);
      quad.y = __shfl_sync(0xffffffffu, pv, qb + 1);
      quad.z = __shfl_sync(0xffffffffu, pv, qb + 2);
      quad.w = __shfl_sync(0xffffffffu, pv, qb + 3);
      const int q0 = q - (lane & 3);
      if (q0 < C) {
        const uint32_t dst = smem_u32(&pbuf[(size_t)rank * Cp + q0]), bb = smem_u32(xb);
        for (uint32_t r = lane & 3; r < CL; r += 4)
          st_async_v4(map_shared_rank(dst, r), quad, map_shared_rank(bb, r));
      }
    }
    mbar_wait(smem_u32(xb), (t >> 1) & 1);     // the partials of every CTA have landed

#pragma unroll
    for (int j = 0; j < kT2MaxQ; ++j) {
      const int q = tid + j * nth;
      if (q >= C) break;
      float pv[kT2MaxCluster];
#pragma unroll
      for (int r = 0; r < kT2MaxCluster; ++r)
        pv[r] = (uint32_t)r < CL ? pbuf[(size_t)r * Cp + q] : S::zero();
      float tot;
      int win = 0;
      if constexpr (SR == LT_LOG) {
        float m = pv[0];
#pragma unroll
        for (int r = 1; r < kT2MaxCluster; ++r) m = fmaxf(m, pv[r]);
        const float ms = fmaxf(m, kClampLow);
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < kT2MaxCluster; ++r) sum += ex2(pv[r] - ms);
        tot = ms + __log2f(sum);
      } else if constexpr (SR == LT_MAXTROPICAL) {
        tot = pv[0];
#pragma unroll
        for (int r = 1; r < kT2MaxCluster; ++r)
          if (pv[r] > tot) { tot = pv[r]; win = r; }   // strict: the lowest rank = lowest arc wins ties
      } else {
        tot = pv[0];
#pragma unroll
        for (int r = 1; r < kT2MaxCluster; ++r) tot += pv[r];
      }
      float blv = bl[j];
      asm volatile("" : "+f"(blv));    // first use of the load stays after the exchange wait
      const float a0 = S::times(cur[q], to_dom<SR>(blv));
      float v;
      if constexpr (SR == LT_MAXTROPICAL) {
        const bool take_blank = a0 >= tot;           // semirings.py:363
        v = take_blank ? a0 : tot;
        if (p.backarc) {
          if (take_blank) { if (mine[j]) p.backarc[(bt0 + t) * C + q] = -1; }
          else if ((uint32_t)win == rank) p.backarc[(bt0 + t) * C + q] = warc[j];
        }
      } else if constexpr (SR == LT_LOG) {
        v = log2_add_exp2(a0, tot);
      } else {
        v = a0 + tot;
      }
      nxt[q] = v;
    }
    __syncthreads();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }

  for (int c = tid; c < C; c += nth) {
    if ((uint32_t)c % CL != rank) continue;
    if (p.alphas)
      for (int t = nf; t < p.T; ++t) p.alphas[(bt0 + t) * C + c] = from_dom<SR>(cur[c]);
    if (p.alpha_final) p.alpha_final[(size_t)b * C + c] = from_dom<SR>(cur[c]);
  }
  if (rank == 0 && tid < 32) {         // dist = (+)_c alpha_T[c]  (lattices.py:496)
    Acc<SR> acc;
    acc.init();
    for (int c = tid; c < C; c += 32) acc.add(from_dom<SR>(cur[c]), c);
    for (int o = 16; o > 0; o >>= 1) {
      Acc<SR> other;
      if constexpr (SR == LT_LOG) {
        other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
        other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
      } else if constexpr (SR == LT_MAXTROPICAL) {
        other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
        other.a = __shfl_xor_sync(0xffffffffu, acc.a, o);
      } else {
        other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
      }
      acc.merge(other);
    }
    if (tid == 0) p.dist[b] = acc.value();
  }
  cluster_sync_all();      // nobody leaves while stores into its buffers may be in flight
}

// ============================================================= backward ==
// 8 lanes per row, 4 rows per warp, blockDim / 8 rows per pass (as lattice_fast2.cu's K2; the
// block is sized so that a slab of up to 36 rows is ONE pass): a lane holds
// up to 32 arcs of its row in registers, so the row maximum and the row sum are 3-step
// shuffles and every arc costs one FFMA + one FADD + one ex2.  Log works in LOG2 units on chip.
template <int SR>
__global__ void __launch_bounds__(kT2BwdThreads, 2)
table_backward2_kernel(const TableParams p, const int R, const int NS) {
  using S = Sr<SR>;
  extern __shared__ __align__(128) unsigned char t2sm[];
  const int C = p.C, V = p.V, Cp = (C + 3) & ~3;
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int rpp = nth >> 3;                                 // rows per pass: 8 lanes per row
  const int sub = lane >> 3, sl = lane & 7;                 // row within the warp, lane within the row
  const uint32_t rank = cluster_ctarank(), CL = cluster_nctarank();
  const int b = blockIdx.x / CL;
  const int row0 = rank * R, nrows = min(R, C - row0);
  const int base = row0 * V;
  const uint32_t slab_bytes = (uint32_t)nrows * V * 4;
  const size_t stage_floats = (size_t)R * V;

  float* slabs = reinterpret_cast<float*>(t2sm);
  unsigned char* ptr = t2sm + (size_t)NS * stage_floats * 4;
  float* beta_buf = reinterpret_cast<float*>(ptr); ptr += (size_t)2 * Cp * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ptr); ptr += (kT2MaxStages + 2) * 8;
  uint64_t* xbar = bars + kT2MaxStages;
  uint16_t* tbl = reinterpret_cast<uint16_t*>(ptr);         // [nrows * V] next states of the slab

  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;
  const float logz = p.dist_in[b];
  const float logz_d = to_dom<SR>(logz);
  const float gscale = p.grad_dist ? p.grad_dist[b] : 1.f;
  const bool scale_ok = (SR != LT_LOG) || is_finite(logz);

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(smem_u32(&bars[s]), 1);
    mbar_init(smem_u32(&xbar[0]), 1);
    mbar_init(smem_u32(&xbar[1]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = tid; c < 2 * Cp; c += nth) beta_buf[c] = to_dom<SR>(S::one());   // lattices.py:789-790
  for (int i = tid; i < nrows * V; i += nth) tbl[i] = (uint16_t)p.table[base + i];
  __syncthreads();
  cluster_sync_all();

  auto issue = [&](int it) {
    const int t = nf - 1 - it;
    const int s = it % NS;
    const uint32_t bar = smem_u32(&bars[s]);
    mbar_arrive_expect_tx(bar, slab_bytes);
    bulk_load_1d(smem_u32(slabs + (size_t)s * stage_floats),
                 p.lexical + (bt0 + t) * (size_t)C * V + base, slab_bytes, bar);
  };
  if (tid == 0)
    for (int it = 0; it < NS && it < nf; ++it) issue(it);

  // padding frames: zero gradients of this CTA's rows (lattices.py:775-779)
  for (int t = nf; t < p.T; ++t) {
    float* gl = p.grad_lexical + (bt0 + t) * (size_t)C * V + base;
    for (int i = tid * 4; i < nrows * V; i += nth * 4)
      stg_stream4(gl + i, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int r = tid; r < nrows; r += nth) p.grad_blank[(bt0 + t) * C + row0 + r] = 0.f;
  }

  // the row owner (lane sl == 0) fetches alpha_t[p], blank_t[p] of its rows in the first two
  // passes one frame ahead
  const int slot = warp * 4 + sub;
  const bool own0 = sl == 0 && slot < nrows, own1 = sl == 0 && slot + rpp < nrows;
  float n_alpha0 = 0.f, n_alpha1 = 0.f, n_blank0 = 0.f, n_blank1 = 0.f;
  auto prefetch = [&](int t) {
    const size_t o = (bt0 + t) * C + row0 + slot;
    if (own0) { n_alpha0 = p.alphas_in[o]; n_blank0 = ldg_stream(p.blank + o); }
    if (own1) { n_alpha1 = p.alphas_in[o + rpp]; n_blank1 = ldg_stream(p.blank + o + rpp); }
  };
  if (nf > 0) prefetch(nf - 1);

  int stage = 0;
  uint32_t parity = 0;
  for (int it = 0; it < nf; ++it) {
    const int t = nf - 1 - it;
    if (it > 0) {
      // every row of the previous frame has arrived from every CTA: beta is complete and
      // the slab stage of iteration it-1 is free
      mbar_wait(smem_u32(&xbar[it & 1]), ((it - 1) >> 1) & 1);
      if (tid == 0 && it - 1 + NS < nf) issue(it - 1 + NS);
    }
    if (tid == 0) mbar_arrive_expect_tx(smem_u32(&xbar[(it + 1) & 1]), (uint32_t)C * 4);
    const float c_alpha0 = n_alpha0, c_alpha1 = n_alpha1, c_blank0 = n_blank0, c_blank1 = n_blank1;
    if (t > 0) prefetch(t - 1);
    const float* beta = beta_buf + (size_t)(it & 1) * Cp;         // beta_{t+1}
    float* nxt = beta_buf + (size_t)((it + 1) & 1) * Cp;
    const float* alpha = p.alphas_in + (bt0 + t) * C;
    const float* blank = p.blank + (bt0 + t) * C;
    float* gb = p.grad_blank + (bt0 + t) * C;
    float* gl = p.grad_lexical + (bt0 + t) * (size_t)C * V;
    mbar_wait(smem_u32(&bars[stage]), parity);
    const float* slab = slabs + (size_t)stage * stage_floats;
    if (++stage == NS) { stage = 0; parity ^= 1; }

    for (int pass = 0; pass * rpp < nrows; ++pass) {
      if (pass * rpp + warp * 4 >= nrows) break;       // warp-uniform: no live row left for this warp
      const int lr = pass * rpp + slot;
      const bool live = lr < nrows;                  // uniform over the 8 lanes of a row
      const int lrc = live ? lr : 0;
      const int prow = row0 + lrc;
      float a_own, k_own;
      if (pass < 2) { a_own = pass ? c_alpha1 : c_alpha0; k_own = pass ? c_blank1 : c_blank0; }
      else { a_own = alpha[prow]; k_own = ldg_stream(blank + prow); }
      const float alpha_p = to_dom<SR>(__shfl_sync(0xffffffffu, a_own, lane & ~7));
      const float* row = slab + (size_t)lrc * V;
      const uint16_t* trow = tbl + (size_t)lrc * V;
      float* grow = gl + (size_t)prow * V;
      float rowv;
      if (V <= 256) {
        float4 x[8];
        float4 bv[SR == LT_REAL ? 8 : 1];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c4 = (sl + 8 * i) * 4;
          if (c4 < V) {
            const float4 w = *reinterpret_cast<const float4*>(row + c4);
            const uint2 tq = *reinterpret_cast<const uint2*>(trow + c4);
            const float4 bn = make_float4(beta[tq.x & 0xffffu], beta[tq.x >> 16],
                                          beta[tq.y & 0xffffu], beta[tq.y >> 16]);
            x[i] = make_float4(arc<SR>(w.x, bn.x), arc<SR>(w.y, bn.y), arc<SR>(w.z, bn.z),
                               arc<SR>(w.w, bn.w));
            if constexpr (SR == LT_REAL) bv[i] = bn;
          } else {
            const float z = SR == LT_LOG ? neg_inf() : 0.f;
            x[i] = make_float4(z, z, z, z);
            if constexpr (SR == LT_REAL) bv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if constexpr (SR == LT_LOG) {
          float m = neg_inf();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
          const float ms = msafe(m);
          const float rs = scale_ok ? gscale * ex2(alpha_p + ms - logz_d) : 0.f;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c4 = (sl + 8 * i) * 4;
            float4 e;
            e.x = ex2(x[i].x - ms); e.y = ex2(x[i].y - ms);
            e.z = ex2(x[i].z - ms); e.w = ex2(x[i].w - ms);
            s += (e.x + e.y) + (e.z + e.w);
            if (live && c4 < V)
              stg_stream4(grow + c4, make_float4(e.x * rs, e.y * rs, e.z * rs, e.w * rs));
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          rowv = ms + __log2f(s);
        } else {
          const float ga = gscale * alpha_p;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c4 = (sl + 8 * i) * 4;
            s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
            if (live && c4 < V)
              stg_stream4(grow + c4, make_float4(ga * bv[i].x, ga * bv[i].y, ga * bv[i].z, ga * bv[i].w));
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          rowv = s;
        }
      } else {
        // wide vocabularies: two passes over the row in shared memory
        if constexpr (SR == LT_LOG) {
          float m = neg_inf();
          for (int y = sl; y < V; y += 8) m = fmaxf(m, arc<SR>(row[y], beta[trow[y]]));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
          m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
          const float ms = msafe(m);
          const float rs = scale_ok ? gscale * ex2(alpha_p + ms - logz_d) : 0.f;
          float s = 0.f;
          for (int y = sl; y < V; y += 8) {
            const float e = ex2(arc<SR>(row[y], beta[trow[y]]) - ms);
            s += e;
            if (live) __stcs(grow + y, e * rs);
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          rowv = ms + __log2f(s);
        } else {
          const float ga = gscale * alpha_p;
          float s = 0.f;
          for (int y = sl; y < V; y += 8) {
            const float bvv = beta[trow[y]];
            s += row[y] * bvv;
            if (live) __stcs(grow + y, ga * bvv);
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          rowv = s;
        }
      }
      float bnew = 0.f;
      if (sl == 0 && live) {
        const float bp = beta[prow];
        const float bb = arc<SR>(k_own, bp);
        if constexpr (SR == LT_LOG) gb[prow] = scale_ok ? gscale * ex2(alpha_p + bb - logz_d) : 0.f;
        else gb[prow] = gscale * a_own * bp;
        bnew = SR == LT_LOG ? log2_add_exp2(bb, rowv) : bb + rowv;
      }
      // lane sl of the row group sends the row's new beta to rank sl (cluster size <= 8)
      xchg_store_group8(nxt, prow, bnew, &xbar[(it + 1) & 1], CL, lane, live);
    }
  }
  if (nf > 0) mbar_wait(smem_u32(&xbar[nf & 1]), ((nf - 1) >> 1) & 1);
  cluster_sync_all();      // nobody leaves while stores into its buffers may be in flight
}

// ------------------------------------------------------------------ host ----
struct T2Geom {
  int CL, R, NS, threads, ell;
  size_t smem;
  uint32_t magic;
};

// ELL width of the forward kernel's arc lists: the average in-degree of a slab, plus one
static int t2_ell(const TableParams& p, int R) {
  return ((int)(((size_t)R * p.V + p.C - 1) / p.C) + 1 + 3) & ~3;      // lists are read four at a time
}

static size_t t2_fixed_bytes(const TableParams& p, int R, bool backward) {
  const size_t Cp = (p.C + 3) & ~3;
  const size_t bars = (kT2MaxStages + 2) * 8;
  if (backward) return 2 * Cp * 4 + bars + (((size_t)R * p.V * 2 + 15) & ~(size_t)15);
  return 2 * (Cp + 4) * 4 + 2 * kT2MaxCluster * Cp * 4 + 2 * Cp * 4 + bars +
         (((size_t)t2_ell(p, R) * Cp * 2 + 15) & ~(size_t)15);
}

static bool t2_geometry(const TableParams& p, bool backward, T2Geom* g) {
  if (option(OPT_TABLE_V1)) return false;
  if (p.k >= 1 || p.T <= 0) return false;
  if (p.V % 4 != 0) return false;                                   // 16-byte bulk copies
  if (reinterpret_cast<uintptr_t>(p.lexical) % 16 != 0) return false;
  if (backward && reinterpret_cast<uintptr_t>(p.grad_lexical) % 16 != 0) return false;
  if (!backward && p.C > kT2MaxThreads * kT2MaxQ) return false;
  if (p.C > 65535) return false;                                    // 16-bit next states
  const int forced = option(OPT_TABLE_CLUSTER);
  for (int cl = 1; cl <= 8; cl <<= 1) {
    const int R = (p.C + cl - 1) / cl;
    if ((cl - 1) * R >= p.C) continue;                              // a rank without rows
    const size_t slab = (size_t)R * p.V * 4 + (backward ? 0 : 16);   // forward: + the padding arc's weight
    if (forced ? cl != forced : (slab > 40 * 1024 && cl < 8)) continue;
    if (!backward && (size_t)R * p.V + 1 > 65535) return false;     // 16-bit slab offsets
    const size_t fixed = t2_fixed_bytes(p, R, backward);
    size_t budget = 113 * 1024;                                     // two CTAs per SM
    if (fixed + 2 * slab > budget) budget = 227 * 1024;
    if (fixed + 2 * slab > budget) return false;
    int ns = (int)((budget - fixed) / slab);
    if (ns > kT2MaxStages) ns = kT2MaxStages;
    g->CL = cl; g->R = R; g->NS = ns;
    g->ell = t2_ell(p, R);
    g->threads = ((R + 3) / 4) * 32;       // backward: 8 lanes per row, one pass where R <= 36
    if (g->threads < 128) g->threads = 128;
    if (g->threads > kT2BwdThreads) g->threads = kT2BwdThreads;
    if (!backward) {                       // one thread per destination where C allows
      g->threads = (p.C + 31) / 32 * 32;
      if (g->threads < 128) g->threads = 128;
      if (g->threads > kT2MaxThreads) g->threads = kT2MaxThreads;
    }
    g->smem = fixed + (size_t)ns * slab;
    g->magic = (uint32_t)(((1ull << 32) + p.V - 1) / p.V);          // a / V for a < 2^16
    return true;
  }
  return false;
}

template <typename KernelT, typename... Args>
static int launch_t2(KernelT kernel, const T2Geom& g, int B, cudaStream_t stream, Args... args) {
  LT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)B * g.CL);
  cfg.blockDim = dim3(g.threads);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LT_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  note_launch();
  return LT_OK;
}

}  // namespace

int table2_cluster_size(int C, int V, int k, bool backward) {
  TableParams p = {};
  p.C = C; p.V = V; p.k = k; p.B = 1; p.T = 1;
  T2Geom g;
  return t2_geometry(p, backward, &g) ? g.CL : 0;     // null pointers count as aligned
}

bool table2_forward_supported(const TableParams& p) {
  T2Geom g;
  return t2_geometry(p, false, &g);
}

bool table2_backward_supported(const TableParams& p) {
  T2Geom g;
  return t2_geometry(p, true, &g);
}

int table2_forward_launch(int semiring, const TableParams& p, cudaStream_t stream) {
  T2Geom g;
  if (!t2_geometry(p, false, &g)) { set_error("table cluster path: unsupported shape"); return LT_ERR_UNSUPPORTED; }
#define LT_T2F(SR) \
  return launch_t2(table_forward2_kernel<SR>, g, p.B, stream, p, g.R, g.NS, g.magic, g.ell)
  if (semiring == LT_LOG) LT_T2F(LT_LOG);
  if (semiring == LT_MAXTROPICAL) LT_T2F(LT_MAXTROPICAL);
  LT_T2F(LT_REAL);
#undef LT_T2F
}

int table2_backward_launch(int semiring, const TableParams& p, cudaStream_t stream) {
  T2Geom g;
  if (!t2_geometry(p, true, &g)) { set_error("table cluster path: unsupported shape"); return LT_ERR_UNSUPPORTED; }
  if (semiring == LT_LOG)
    return launch_t2(table_backward2_kernel<LT_LOG>, g, p.B, stream, p, g.R, g.NS);
  return launch_t2(table_backward2_kernel<LT_REAL>, g, p.B, stream, p, g.R, g.NS);
}

}  // namespace lt
