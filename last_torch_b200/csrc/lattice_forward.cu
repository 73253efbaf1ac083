// K1 (generic): persistent forward (alpha) recursion over the recognition
// lattice.  Replaces the Python T-loop of RecognitionLattice._forward
// (/root/reference/last_torch/lattices.py:379-496, :856-892).
//
// One thread-block CLUSTER per utterance.  Every CTA of the cluster keeps a
// full copy of alpha_t in shared memory for all T frames and owns a contiguous
// slice of DESTINATION states: it streams exactly the arc weights that lead
// into its slice (coalesced along the destination index), reduces them with the
// semiring (+), and all-gathers the new slice into every peer's shared memory
// through DSMEM, followed by one cluster barrier per recursion level.
//
// The destination-major formulation follows FullNGram.forward_reduce
// (contexts.py:207-230): see NGram in common.cuh.
#include "common.cuh"
#include "params.cuh"

namespace lt {


// Reduce all arcs into the destination slice [q_lo, q_hi) from source vector
// `src` (shared memory, full C entries).  Result for destination q is left in
// pm[q - q_lo] (value; Log: the running maximum m) and ps[q - q_lo] (MaxTropical: arg-max;
// Log: the sum s relative to msafe(m), i.e. value = msafe(m) + log(s) -- kept as a pair so that
// the caller can merge further terms and round the new alpha ONCE).
template <int SR>
__device__ __forceinline__ void reduce_into_slice(
    const NGram& g, const float* __restrict__ lex, const float* __restrict__ src,
    int q_lo, int q_hi, float* pm, float* ps, int ppad) {
  using S = Sr<SR>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int D = q_hi - q_lo;
  const int nchunk = (D + 31) >> 5;
  int KG = nwarps / nchunk;
  if (KG < 1) KG = 1;
  if (KG > g.K) KG = g.K;
  const int kper = (g.K + KG - 1) / KG;
  const int ntask = nchunk * KG;
  const int lowV = g.Alow * g.V;

  for (int task = warp; task < ntask; task += nwarps) {
    const int chunk = task % nchunk, kg = task / nchunk;
    const int d = (chunk << 5) + lane;
    const int q = q_lo + d;
    Acc<SR> acc; acc.init();
    if (d < D) {
      const int qa = q - g.off;       // flat index if this is a single-arc dest
      if (g.n > 0 && q == 0) {
        // state 0 has no incoming lexical arc (contexts.py:217-218)
      } else if (qa < lowV) {
        if (kg == 0) {
          const int p = qa / g.V;
          if constexpr (SR == LT_LOG) acc.add_d((double)src[p] + (double)ldg_stream(lex + qa));
          else acc.add(S::times(src[p], ldg_stream(lex + qa)), 0);
        }
      } else {
        const int j = q - g.A;
        const float* col = lex + lowV + j;
        const int p0 = g.Alow + (g.pstride ? j / g.V : 0);
        const int k_lo = kg * kper;
        const int k_hi = min(g.K, k_lo + kper);
        int kk = k_lo;
        for (; kk + 8 <= k_hi; kk += 8) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            x[i] = ldg_stream(col + (size_t)(kk + i) * g.N);
          if constexpr (SR == LT_LOG) {
            double xd[8];
            float cm = neg_inf();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              xd[i] = (double)src[p0 + (kk + i) * g.pstride] + (double)x[i];
              cm = fmaxf(cm, (float)xd[i]);
            }
            acc.add_chunk_d(xd, cm);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              x[i] = S::times(src[p0 + (kk + i) * g.pstride], x[i]);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc.add(x[i], kk + i);
          }
        }
        for (; kk < k_hi; ++kk) {
          if constexpr (SR == LT_LOG)
            acc.add_d((double)src[p0 + kk * g.pstride] + (double)ldg_stream(col + (size_t)kk * g.N));
          else
            acc.add(S::times(src[p0 + kk * g.pstride],
                             ldg_stream(col + (size_t)kk * g.N)), kk);
        }
      }
      const int slot = kg * (nchunk << 5) + d;
      if constexpr (SR == LT_LOG) { pm[slot] = acc.m; ps[slot] = acc.s; }
      else if constexpr (SR == LT_MAXTROPICAL) { pm[slot] = acc.m; ps[slot] = __int_as_float(acc.a); }
      else { pm[slot] = acc.s; }
    }
  }
  __syncthreads();
  // combine the KG partials of each destination into slot d of group 0
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    Acc<SR> acc;
    const int stride = nchunk << 5;
    if constexpr (SR == LT_LOG) { acc.m = pm[d]; acc.s = ps[d]; }
    else if constexpr (SR == LT_MAXTROPICAL) { acc.m = pm[d]; acc.a = __float_as_int(ps[d]); }
    else { acc.s = pm[d]; }
    for (int kg = 1; kg < KG; ++kg) {
      Acc<SR> o;
      const int slot = kg * stride + d;
      if constexpr (SR == LT_LOG) { o.m = pm[slot]; o.s = ps[slot]; }
      else if constexpr (SR == LT_MAXTROPICAL) { o.m = pm[slot]; o.a = __float_as_int(ps[slot]); }
      else { o.s = pm[slot]; }
      acc.merge(o);
    }
    if constexpr (SR == LT_LOG) { pm[d] = acc.m; ps[d] = acc.s; }
    else pm[d] = acc.value();
    if constexpr (SR == LT_MAXTROPICAL) ps[d] = __int_as_float(acc.arg());
  }
  __syncthreads();
}

// Store v into element idx of the shared array `base` of EVERY CTA in the cluster.
__device__ __forceinline__ void bcast_store(float* base, int idx, float v, uint32_t nrank) {
  const uint32_t a = smem_u32(base + idx);
  for (uint32_t r = 0; r < nrank; ++r) st_shared_cluster_f32(map_shared_rank(a, r), v);
}

template <int SR, bool FLD>
__global__ void __launch_bounds__(512)
lattice_forward_generic(const FwdParams p) {
  using S = Sr<SR>;
  extern __shared__ __align__(16) float smem[];
  const NGram& g = p.g;
  const int C = g.C;
  const int Cp = (C + 3) & ~3;
  const uint32_t nrank = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / nrank;
  const int tid = threadIdx.x, nth = blockDim.x;

  float* alpha0 = smem;              // ping
  float* alpha1 = alpha0 + Cp;       // pong
  float* last0 = alpha1 + Cp;        // FLD level ping
  float* last1 = last0 + (FLD ? Cp : 0);
  float* pm = last1 + (FLD ? Cp : 0);
  float* ps = pm + p.ppad;
  float* am = ps + p.ppad;           // FLD: running (+) of the terminated terms
  float* as = am + (FLD ? p.dslice : 0);

  const int q_lo = min(C, (int)rank * p.dslice);
  const int q_hi = min(C, q_lo + p.dslice);
  const int D = q_hi - q_lo;

  int nf = p.num_frames[b];
  nf = max(0, min(nf, p.T));

  for (int c = tid; c < C; c += nth)
    alpha0[c] = p.alpha_init ? p.alpha_init[(size_t)b * C + c]
                             : (c == 0 ? S::one() : S::zero());
  __syncthreads();
  cluster_sync_all();   // peers must not write into our buffers before init is done

  float* cur = alpha0;
  float* nxt = alpha1;
  const size_t bt0 = (size_t)b * p.T;
  // Renormalised recursion state (lt_lattice_forward_norm, Log only): alpha_t = alpha~_t + off_t
  // with an exact integer offset (natural-log units here) that follows floor(max_c alpha~_t);
  // every CTA derives the shift from its own identical replica.
  const bool norm = SR == LT_LOG && p.alpha_norm != nullptr;
  int32_t* an = norm ? p.alpha_norm + (size_t)b * (p.T + 3) : nullptr;
  __shared__ float shift_s;
  int off = 0;
  float shift = 0.f;

  const int LW = p.wlevels > 1 ? p.wlevels : 1;             // weight sets per frame
  const size_t lvb = LW > 1 ? (size_t)C : 0;                  // stride between the sets
  const size_t lvl = LW > 1 ? (size_t)C * g.V : 0;
  for (int t = 0; t < nf; ++t) {
    const float* blank = p.blank + (bt0 + t) * C * LW;        // level 0
    const float* lex = p.lexical + (bt0 + t) * (size_t)C * g.V * LW;
    if (p.alphas) {
      float* out = p.alphas + (bt0 + t) * C;
      for (int d = tid; d < D; d += nth) out[q_lo + d] = cur[q_lo + d];
    }
    if (norm) {
      if (tid < 32) {
        float m = neg_inf();
        for (int c = tid; c < C; c += 32) m = fmaxf(m, cur[c]);
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (tid == 0) {
          shift_s = norm_shift(m);
          if (rank == 0) an[t] = off;
        }
      }
      __syncthreads();
      shift = shift_s;
      off += (int)shift;
    }
    if constexpr (!FLD) {
      reduce_into_slice<SR>(g, lex, cur, q_lo, q_hi, pm, ps, p.ppad);
      for (int d = tid; d < D; d += nth) {
        const int q = q_lo + d;
        const float a = S::times(cur[q], blank[q]);
        const float r = pm[d];
        float v;
        if constexpr (SR == LT_LOG) {
          // merge the blank term into the (m, s) pair of the reduction and round once, after
          // the shift: (msafe(m) - shift) + log(s) in double
          Acc<LT_LOG> acc; acc.m = pm[d]; acc.s = ps[d];
          acc.add_d((double)cur[q] + (double)blank[q]);
          v = log_value_shifted(acc.m, acc.s, shift);
        } else if constexpr (SR == LT_MAXTROPICAL) {
          const bool take_blank = a >= r;      // semirings.py:363
          v = take_blank ? a : r;
          if (p.backptr)
            p.backptr[(bt0 + t) * C + q] = take_blank ? (int16_t)-1 : (int16_t)__float_as_int(ps[d]);
        } else {
          v = S::plus(a, r);
        }
        bcast_store(nxt, q, v, nrank);
      }
      cluster_sync_all();
    } else {
      // term_0 = alpha (x) blank
      for (int d = tid; d < D; d += nth) {
        const int q = q_lo + d;
        const float a = S::times(cur[q], blank[q]);
        if constexpr (SR == LT_LOG) {
          Acc<LT_LOG> acc; acc.init(); acc.add(a, 0);   // handles a == -inf
          am[d] = acc.m; as[d] = acc.s;
        }
        else if constexpr (SR == LT_MAXTROPICAL) { am[d] = a; as[d] = __int_as_float(0); }
        else { am[d] = a; }
      }
      const float* src = cur;
      float* lv = last0;
      for (int i = 0; i < p.k; ++i) {
        reduce_into_slice<SR>(g, lex + i * lvl, src, q_lo, q_hi, pm, ps, p.ppad);
        const bool need_bcast = (i + 1 < p.k);
        for (int d = tid; d < D; d += nth) {
          const int q = q_lo + d;
          float r = pm[d];
          if constexpr (SR == LT_LOG) r = log_value_shifted(pm[d], ps[d], 0.f);
          if (p.levels) p.levels[((bt0 + t) * p.k + i) * C + q] = r;
          if constexpr (SR == LT_MAXTROPICAL) {
            if (p.backptr) p.backptr[((bt0 + t) * p.k + i) * C + q] = (int16_t)__float_as_int(ps[d]);
          }
          const float term = S::times(r, blank[(i + 1) * lvb + q]);
          if constexpr (SR == LT_LOG) {
            Acc<LT_LOG> acc; acc.m = am[d]; acc.s = as[d];
            // am holds the running max, as the running sum relative to msafe(am)
            acc.add(term, 0);
            am[d] = acc.m; as[d] = acc.s;
          } else if constexpr (SR == LT_MAXTROPICAL) {
            if (term > am[d]) { am[d] = term; as[d] = __int_as_float(i + 1); }
          } else {
            am[d] += term;
          }
          if (need_bcast) bcast_store(lv, q, r, nrank);
        }
        if (need_bcast) {
          cluster_sync_all();
          src = lv;
          lv = (lv == last0) ? last1 : last0;
        } else {
          __syncthreads();
        }
      }
      for (int d = tid; d < D; d += nth) {
        const int q = q_lo + d;
        float v;
        if constexpr (SR == LT_LOG) { v = log_value_shifted(am[d], as[d], shift); }
        else { v = am[d]; }
        if constexpr (SR == LT_MAXTROPICAL) {
          if (p.termptr) p.termptr[(bt0 + t) * C + q] = (uint8_t)__float_as_int(as[d]);
        }
        bcast_store(nxt, q, v, nrank);
      }
      cluster_sync_all();
    }
    float* tmp = cur; cur = nxt; nxt = tmp;
  }

  // padding frames keep alpha (lattices.py:460-461) but are still recorded (:462)
  if (p.alphas) {
    for (int t = nf; t < p.T; ++t) {
      float* out = p.alphas + (bt0 + t) * C;
      for (int d = tid; d < D; d += nth) out[q_lo + d] = cur[q_lo + d];
    }
  }
  if (p.alpha_final)
    for (int d = tid; d < D; d += nth)
      p.alpha_final[(size_t)b * C + q_lo + d] =
          norm ? (float)((double)cur[q_lo + d] + (double)off) : cur[q_lo + d];

  // dist = (+)_c alpha_T[c]  (lattices.py:496), by CTA 0 of the cluster.
  if (rank == 0) {
    Acc<SR> acc; acc.init();
    if constexpr (SR == LT_LOG) {
      // two passes keep the exact global max like torch's logsumexp
      float m = neg_inf();
      for (int c = tid; c < C; c += nth) m = fmaxf(m, cur[c]);
      pm[tid] = m;
      __syncthreads();
      for (int s = nth >> 1; s > 0; s >>= 1) {
        if (tid < s) pm[tid] = fmaxf(pm[tid], pm[tid + s]);
        __syncthreads();
      }
      const float ms = msafe(pm[0]);
      __syncthreads();
      float s = 0.f;
      for (int c = tid; c < C; c += nth) s += fast_exp(cur[c] - ms);
      pm[tid] = s;
      __syncthreads();
      for (int st = nth >> 1; st > 0; st >>= 1) {
        if (tid < st) pm[tid] += pm[tid + st];
        __syncthreads();
      }
      if (tid == 0) {
        const double rd = (double)ms + log((double)pm[0]);
        const float r = (float)rd;
        p.dist[b] = norm ? (float)(rd + (double)off) : r;
        if (norm) {
          for (int t = nf; t <= p.T; ++t) an[t] = off;
          an[p.T + 1] = __float_as_int(r);
          an[p.T + 2] = 1;               // offsets are in natural-log units
        }
      }
    } else {
      for (int c = tid; c < C; c += nth) acc.add(cur[c], c);
      if constexpr (SR == LT_MAXTROPICAL) { pm[tid] = acc.m; ps[tid] = __int_as_float(acc.a); }
      else { pm[tid] = acc.s; }
      __syncthreads();
      for (int s = nth >> 1; s > 0; s >>= 1) {
        if (tid < s) {
          if constexpr (SR == LT_MAXTROPICAL) {
            Acc<SR> x, y;
            x.m = pm[tid]; x.a = __float_as_int(ps[tid]);
            y.m = pm[tid + s]; y.a = __float_as_int(ps[tid + s]);
            x.merge(y);
            pm[tid] = x.m; ps[tid] = __int_as_float(x.a);
          } else {
            pm[tid] += pm[tid + s];
          }
        }
        __syncthreads();
      }
      if (tid == 0) p.dist[b] = pm[0];
    }
  }
  // nobody may exit while a peer can still write into its shared memory
  cluster_sync_all();
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
int pick_cluster_size(const NGram& g, int B, unsigned flags, int sm_count) {
  int forced = (flags >> LT_FLAG_CLUSTER_SHIFT) & 0xf;
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
  // enough arcs per frame to be worth splitting, and enough SMs to host it
  const long long arcs = (long long)g.C * g.V;
  int s = 1;
  while (s < 8 && (long long)B * s * 2 <= sm_count && arcs / (s * 2) >= 4096) s *= 2;
  return s;
}

template <typename KernelT>
static int launch_cluster(KernelT kernel, int grid, int block, size_t smem, int cluster,
                          cudaStream_t stream, const FwdParams& p) {
  LT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
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

int lattice_forward_generic_launch(int semiring, const NGram& g, int k, const FwdParams& base,
                                   unsigned flags, int sm_count, cudaStream_t stream) {
  FwdParams p = base;
  const bool fld = k >= 1;
  const int cluster = pick_cluster_size(g, p.B, flags, sm_count);
  p.dslice = (g.C + cluster - 1) / cluster;
  const long long work = (long long)p.dslice * g.K;
  const int block = work >= 8192 ? 512 : (work >= 1024 ? 256 : 128);
  const int nwarps = block / 32;
  p.ppad = max(round_up(p.dslice, 32), 32 * nwarps);
  if (p.ppad < block) p.ppad = block;
  const int Cp = (g.C + 3) & ~3;
  size_t smem = sizeof(float) * ((size_t)Cp * (fld ? 4 : 2) + 2 * (size_t)p.ppad +
                                 (fld ? 2 * (size_t)p.dslice : 0));
  if (smem > 227 * 1024) {
    set_error("lt_lattice_forward: %d context states need %zu bytes of shared memory per CTA "
              "(limit 232448); context too large for the on-chip alpha design", g.C, smem);
    return LT_ERR_UNSUPPORTED;
  }
  const int grid = p.B * cluster;
#define LT_LAUNCH(SR)                                                                     \
  return fld ? launch_cluster(lattice_forward_generic<SR, true>, grid, block, smem, cluster, stream, p) \
             : launch_cluster(lattice_forward_generic<SR, false>, grid, block, smem, cluster, stream, p)
  switch (semiring) {
    case LT_REAL: LT_LAUNCH(LT_REAL);
    case LT_LOG: LT_LAUNCH(LT_LOG);
    case LT_MAXTROPICAL: LT_LAUNCH(LT_MAXTROPICAL);
  }
#undef LT_LAUNCH
  set_error("lt_lattice_forward: unknown semiring %d", semiring);
  return LT_ERR_INVALID_ARGUMENT;
}

}  // namespace lt
