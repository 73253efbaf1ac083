// Lattice kernels for an ARBITRARY context DFA given as a next-state table
// (contexts.NextStateTable, /root/reference/last_torch/contexts.py:266-324): the same
// recursions as lattice_forward.cu / lattice_backward.cu / viterbi.cu, with table-driven
// indexing instead of the closed-form FullNGram geometry.
//
//   forward : destination-major "pull" over a CSR of incoming arcs (arcs grouped by
//             destination state, ascending flat index p*V+y inside a group, built once by
//             the host), so that (+) needs no atomics in any semiring;
//   backward: source-major, one warp per source row: lexical[p, :] is a coalesced stream,
//             beta'[table[p, y]] a shared-memory gather; ONE exponential per arc serves
//             both the row log-sum-exp (beta) and the arc posterior;
//   Viterbi : back-pointers are the winning incoming ARC (flat index, int32; -1 = blank).
// One CTA per utterance, alpha / beta (and the FrameLabelDependent level vectors) in shared
// memory for all T frames.  The reference implements forward_reduce for the Real semiring
// only; here every semiring follows the documented contract
// out[q] = (+)_{p -y-> q} w[p, y] (contexts.py:74-90), first arg-max in flat-arc order.
#include "common.cuh"
#include "params.cuh"
#include "table_params.cuh"

namespace lt {


namespace {

// (+) over the incoming arcs of destination q: value in acc, arg = winning flat arc
template <int SR>
__device__ __forceinline__ void pull_dest(const TableParams& p, const float* __restrict__ lex,
                                          const float* __restrict__ src, int q, Acc<SR>& acc) {
  using S = Sr<SR>;
  acc.init();
  const int lo = p.in_offsets[q], hi = p.in_offsets[q + 1];
  for (int i = lo; i < hi; ++i) {
    const int arc = p.in_arcs[i];
    acc.add(S::times(src[arc / p.V], ldg_stream(lex + arc)), arc);
  }
}

template <int SR, bool FLD>
__global__ void __launch_bounds__(512)
table_forward_kernel(const TableParams p) {
  using S = Sr<SR>;
  extern __shared__ __align__(16) float tsm[];
  const int C = p.C, V = p.V;
  const int Cp = (C + 3) & ~3;
  float* cur = tsm;
  float* nxt = cur + Cp;
  float* lv0 = nxt + Cp;                       // FLD level ping / pong
  float* lv1 = lv0 + (FLD ? Cp : 0);
  float* am = lv1 + (FLD ? Cp : 0);            // FLD: running (+) of the terminated terms
  float* as = am + (FLD ? Cp : 0);
  const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;
  const int nlev = FLD ? p.k : 1;

  for (int c = tid; c < C; c += nth)
    cur[c] = p.alpha_init ? p.alpha_init[(size_t)b * C + c] : (c == 0 ? S::one() : S::zero());
  __syncthreads();

  for (int t = 0; t < nf; ++t) {
    const float* blank = p.blank + (bt0 + t) * C;
    const float* lex = p.lexical + (bt0 + t) * (size_t)C * V;
    if (p.alphas)
      for (int c = tid; c < C; c += nth) p.alphas[(bt0 + t) * C + c] = cur[c];
    if constexpr (!FLD) {
      for (int q = tid; q < C; q += nth) {
        Acc<SR> acc;
        pull_dest<SR>(p, lex, cur, q, acc);
        const float a = S::times(cur[q], blank[q]);
        float v;
        if constexpr (SR == LT_MAXTROPICAL) {
          const float r = acc.value();
          const bool take_blank = a >= r;          // semirings.py:363
          v = take_blank ? a : r;
          if (p.backarc) p.backarc[(bt0 + t) * C + q] = take_blank ? -1 : acc.arg();
        } else {
          v = S::plus(a, acc.value());
        }
        nxt[q] = v;
      }
      __syncthreads();
    } else {
      // term_0 = alpha (x) blank; last_{i+1} = reduce(last_i (x) lex); alpha' = (+)_i last_i (x) blank
      for (int q = tid; q < C; q += nth) {
        Acc<SR> term; term.init(); term.add(S::times(cur[q], blank[q]), 0);
        if constexpr (SR == LT_LOG) { am[q] = term.m; as[q] = term.s; }
        else if constexpr (SR == LT_MAXTROPICAL) { am[q] = term.m; as[q] = __int_as_float(term.a); }
        else { am[q] = term.s; }
      }
      const float* src = cur;
      float* lv = lv0;
      for (int i = 0; i < nlev; ++i) {
        for (int q = tid; q < C; q += nth) {
          Acc<SR> acc;
          pull_dest<SR>(p, lex, src, q, acc);
          const float r = acc.value();
          if (p.levels) p.levels[((bt0 + t) * p.k + i) * C + q] = r;
          if constexpr (SR == LT_MAXTROPICAL) {
            if (p.backarc) p.backarc[((bt0 + t) * p.k + i) * C + q] = acc.arg();
          }
          Acc<SR> term;
          if constexpr (SR == LT_LOG) { term.m = am[q]; term.s = as[q]; }
          else if constexpr (SR == LT_MAXTROPICAL) { term.m = am[q]; term.a = __float_as_int(as[q]); }
          else { term.s = am[q]; }
          term.add(S::times(r, blank[q]), i + 1);  // strict '>' keeps fewer expansions
          if constexpr (SR == LT_LOG) { am[q] = term.m; as[q] = term.s; }
          else if constexpr (SR == LT_MAXTROPICAL) { am[q] = term.m; as[q] = __int_as_float(term.a); }
          else { am[q] = term.s; }
          if (i + 1 < nlev) {
            lv[q] = r;
          } else {
            nxt[q] = term.value();
            if constexpr (SR == LT_MAXTROPICAL) {
              if (p.termptr) p.termptr[(bt0 + t) * C + q] = (uint8_t)term.arg();
            }
          }
        }
        __syncthreads();
        src = lv;
        lv = (lv == lv0) ? lv1 : lv0;
      }
    }
    float* tmp = cur; cur = nxt; nxt = tmp;
  }

  for (int c = tid; c < C; c += nth) {
    if (p.alphas)
      for (int t = nf; t < p.T; ++t) p.alphas[(bt0 + t) * C + c] = cur[c];
    if (p.alpha_final) p.alpha_final[(size_t)b * C + c] = cur[c];
  }
  if (tid == 0) {                      // dist = (+)_c alpha_T[c]  (lattices.py:496)
    if constexpr (SR == LT_LOG) {
      float m = neg_inf();
      for (int c = 0; c < C; ++c) m = fmaxf(m, cur[c]);
      const float ms = msafe(m);
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += fast_exp(cur[c] - ms);
      p.dist[b] = ms + fast_log(s);
    } else {
      float m = (SR == LT_REAL) ? 0.f : neg_inf();
      for (int c = 0; c < C; ++c) m = S::plus(m, cur[c]);
      p.dist[b] = m;
    }
  }
}

// One warp per source row p: lsum = (+)_y lex[p,y] (x) nb[table[p,y]];
//   Log : grad_lex[p,y] (+)= g * exp(src_alpha[p] + lex + nb - logZ)
//   Real: grad_lex[p,y] (+)= g * src_alpha[p] * nb
template <int SR>
__device__ __forceinline__ void table_rows_backward(
    const TableParams& p, const float* __restrict__ lex, float* __restrict__ glex,
    const float* __restrict__ nb, const float* __restrict__ src_alpha, float logz, float gscale,
    bool scale_ok, bool accumulate, float* row_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int C = p.C, V = p.V;
  for (int r = warp; r < C; r += nwarps) {
    const float* row = lex + (size_t)r * V;
    const int32_t* trow = p.table + (size_t)r * V;
    float* grow = glex + (size_t)r * V;
    const float a = src_alpha[r];
    if constexpr (SR == LT_LOG) {
      float m = neg_inf();
      for (int y = lane; y < V; y += 32) m = fmaxf(m, ldg_stream(row + y) + nb[trow[y]]);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float ms = msafe(m);
      const float rs = scale_ok ? gscale * fast_exp(a + ms - logz) : 0.f;
      float s = 0.f;
      for (int y = lane; y < V; y += 32) {
        const float e = fast_exp(ldg_stream(row + y) + nb[trow[y]] - ms);
        s += e;
        const float gv = e * rs;
        grow[y] = accumulate ? grow[y] + gv : gv;
      }
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) row_out[r] = ms + fast_log(s);
    } else {
      float s = 0.f;
      const float ga = gscale * a;
      for (int y = lane; y < V; y += 32) {
        const float bv = nb[trow[y]];
        s += ldg_stream(row + y) * bv;
        const float gv = ga * bv;
        grow[y] = accumulate ? grow[y] + gv : gv;
      }
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) row_out[r] = s;
    }
  }
}

template <int SR, bool FLD>
__global__ void __launch_bounds__(512)
table_backward_kernel(const TableParams p) {
  using S = Sr<SR>;
  extern __shared__ __align__(16) float tsm[];
  const int C = p.C, V = p.V;
  const int Cp = (C + 3) & ~3;
  float* buf0 = tsm;
  float* buf1 = buf0 + Cp;
  float* buf2 = buf1 + Cp;                     // FLD only
  float* row_out = buf2 + (FLD ? Cp : 0);      // [C]
  const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;
  const float logz = p.dist_in[b];
  const float gscale = p.grad_dist ? p.grad_dist[b] : 1.f;
  const bool scale_ok = (SR != LT_LOG) || is_finite(logz);

  for (int t = nf; t < p.T; ++t) {             // padding frames: zero gradients
    float* gb = p.grad_blank + (bt0 + t) * C;
    float* gl = p.grad_lexical + (bt0 + t) * (size_t)C * V;
    for (int c = tid; c < C; c += nth) gb[c] = 0.f;
    for (size_t i = tid; i < (size_t)C * V; i += nth) gl[i] = 0.f;
  }
  float* beta = buf0;
  float* nxt = buf1;
  float* spare = buf2;
  for (int c = tid; c < C; c += nth) beta[c] = S::one();   // lattices.py:789-790
  __syncthreads();

  for (int t = nf - 1; t >= 0; --t) {
    const float* blank = p.blank + (bt0 + t) * C;
    const float* lex = p.lexical + (bt0 + t) * (size_t)C * V;
    const float* alpha = p.alphas_in + (bt0 + t) * C;
    float* gb = p.grad_blank + (bt0 + t) * C;
    float* gl = p.grad_lexical + (bt0 + t) * (size_t)C * V;
    if constexpr (!FLD) {
      table_rows_backward<SR>(p, lex, gl, beta, alpha, logz, gscale, scale_ok, false, row_out);
      __syncthreads();
      for (int q = tid; q < C; q += nth) {
        const float bb = S::times(blank[q], beta[q]);
        if constexpr (SR == LT_LOG) gb[q] = scale_ok ? gscale * fast_exp(alpha[q] + bb - logz) : 0.f;
        else gb[q] = gscale * alpha[q] * beta[q];
        nxt[q] = S::plus(bb, row_out[q]);
      }
      __syncthreads();
      float* tmp = beta; beta = nxt; nxt = tmp;
    } else {
      const int k = p.k;
      const float* lev = p.levels_in + (bt0 + t) * (size_t)k * C;   // last_1 .. last_k
      float* nb = nxt;
      for (int c = tid; c < C; c += nth) nb[c] = S::times(blank[c], beta[c]);   // alignments.py:405
      for (int q = tid; q < C; q += nth) {                                      // blank marginals
        float acc = 0.f;
        if constexpr (SR == LT_LOG) {
          if (scale_ok) {
            const float base = blank[q] + beta[q] - logz;
            acc = fast_exp(alpha[q] + base);
            for (int i = 0; i < k; ++i) acc += fast_exp(lev[(size_t)i * C + q] + base);
            acc *= gscale;
          }
        } else {
          acc = alpha[q];
          for (int i = 0; i < k; ++i) acc += lev[(size_t)i * C + q];
          acc *= gscale * beta[q];
        }
        gb[q] = acc;
      }
      __syncthreads();
      float* out = spare;
      for (int j = k - 1; j >= 0; --j) {
        const float* src_alpha = (j == 0) ? alpha : lev + (size_t)(j - 1) * C;
        table_rows_backward<SR>(p, lex, gl, nb, src_alpha, logz, gscale, scale_ok, j != k - 1, row_out);
        __syncthreads();
        for (int q = tid; q < C; q += nth)
          out[q] = S::plus(S::times(blank[q], beta[q]), row_out[q]);            // alignments.py:414-415
        __syncthreads();
        float* tmp = nb; nb = out; out = tmp;
      }
      float* old = beta;
      beta = nb; nxt = old; spare = out;
    }
  }
}

__global__ void table_viterbi_kernel(const TableParams p, const float* alpha_final,
                                     int32_t* labels, int32_t* path_states, float* grad_blank,
                                     float* grad_lexical) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int C = p.C, V = p.V;
  const bool fld = p.k >= 1;
  const int nlev = fld ? p.k : 1, nlab = fld ? p.k + 1 : 1;
  const int nf = max(0, min(p.num_frames[b], p.T));
  float m = neg_inf(); int q = 0;
  for (int c = 0; c < C; ++c) {              // first arg-max of alpha_T (lattices.py:496)
    const float v = alpha_final[(size_t)b * C + c];
    if (v > m) { m = v; q = c; }
  }
  for (size_t i = 0; i < (size_t)p.T * nlab; ++i) labels[(size_t)b * p.T * nlab + i] = 0;
  if (path_states)
    for (int t = nf; t <= p.T; ++t) path_states[(size_t)b * (p.T + 1) + t] = q;
  const float gscale = p.grad_dist ? p.grad_dist[b] : 1.f;
  for (int t = nf - 1; t >= 0; --t) {
    const size_t bt = (size_t)b * p.T + t;
    const int32_t* bp = p.backarc + bt * (size_t)nlev * C;
    if (!fld) {
      const int arc = bp[q];
      if (arc < 0) {
        if (grad_blank) atomicAdd(grad_blank + bt * C + q, gscale);
      } else {
        if (grad_lexical) atomicAdd(grad_lexical + bt * (size_t)C * V + arc, gscale);
        labels[bt] = arc % V + 1;
        q = arc / V;
      }
    } else {
      const int nexp = p.termptr[bt * C + q];
      if (grad_blank) atomicAdd(grad_blank + bt * C + q, gscale);
      for (int i = nexp - 1; i >= 0; --i) {
        const int arc = bp[(size_t)i * C + q];
        if (grad_lexical) atomicAdd(grad_lexical + bt * (size_t)C * V + arc, gscale);
        labels[bt * nlab + i] = arc % V + 1;
        q = arc / V;
      }
    }
    if (path_states) path_states[(size_t)b * (p.T + 1) + t] = q;
  }
}

// forward_reduce on arbitrary leading dims: out[o, q] = (+)_{arcs into q} w[o, arc]
template <int SR>
__global__ void table_reduce_forward_kernel(const float* __restrict__ w,
                                            const int32_t* __restrict__ in_offsets,
                                            const int32_t* __restrict__ in_arcs, long long outer,
                                            int C, int V, float* __restrict__ out,
                                            int32_t* __restrict__ argarc) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= outer * C) return;
  const long long o = idx / C;
  const int q = (int)(idx - o * C);
  const float* wo = w + (size_t)o * C * V;
  Acc<SR> acc; acc.init();
  if constexpr (SR == LT_LOG) {            // exact max first, like torch.logsumexp
    float m = neg_inf();
    for (int i = in_offsets[q]; i < in_offsets[q + 1]; ++i) m = fmaxf(m, wo[in_arcs[i]]);
    const float ms = msafe(m);
    float s = 0.f;
    for (int i = in_offsets[q]; i < in_offsets[q + 1]; ++i) s += fast_exp(wo[in_arcs[i]] - ms);
    out[idx] = in_offsets[q] == in_offsets[q + 1] ? neg_inf() : ms + fast_log(s);
  } else {
    for (int i = in_offsets[q]; i < in_offsets[q + 1]; ++i) acc.add(wo[in_arcs[i]], in_arcs[i]);
    out[idx] = acc.value();
    if constexpr (SR == LT_MAXTROPICAL) {
      if (argarc) argarc[idx] = in_offsets[q] == in_offsets[q + 1] ? -1 : acc.arg();
    }
  }
}

template <int SR>
__global__ void table_reduce_backward_kernel(const float* __restrict__ w,
                                             const float* __restrict__ out,
                                             const int32_t* __restrict__ argarc,
                                             const float* __restrict__ gout,
                                             const int32_t* __restrict__ table, long long outer,
                                             int C, int V, float* __restrict__ gw) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)C * V;
  if (idx >= outer * per) return;
  const long long o = idx / per;
  const int arc = (int)(idx - o * per);
  const int q = table[arc];
  const float g = gout[o * C + q];
  if constexpr (SR == LT_REAL) {
    gw[idx] = g;
  } else if constexpr (SR == LT_LOG) {      // safe gradient: 0 when the sum is -inf
    const float z = out[o * C + q];
    gw[idx] = is_finite(z) ? g * fast_exp(w[idx] - z) : 0.f;
  } else {
    gw[idx] = argarc[o * C + q] == arc ? g : 0.f;
  }
}

}  // namespace

static size_t table_smem(int C, bool fld, bool backward) {
  const size_t Cp = (C + 3) & ~3;
  if (backward) return sizeof(float) * (Cp * (fld ? 3 : 2) + Cp);
  return sizeof(float) * Cp * (fld ? 6 : 2);
}

int table_lattice_forward_launch(int semiring, const TableParams& p, cudaStream_t stream) {
  if (table2_forward_supported(p)) return table2_forward_launch(semiring, p, stream);
  const bool fld = p.k >= 1;
  const size_t smem = table_smem(p.C, fld, false);
  if (smem > 227 * 1024) {
    set_error("lt_table_lattice_forward: %d context states need %zu bytes of shared memory", p.C, smem);
    return LT_ERR_UNSUPPORTED;
  }
#define LT_TF(SR)                                                                             \
  do {                                                                                        \
    if (fld) {                                                                                \
      LT_CUDA(cudaFuncSetAttribute(table_forward_kernel<SR, true>,                            \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      table_forward_kernel<SR, true><<<p.B, 512, smem, stream>>>(p);                          \
    } else {                                                                                  \
      LT_CUDA(cudaFuncSetAttribute(table_forward_kernel<SR, false>,                           \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      table_forward_kernel<SR, false><<<p.B, 512, smem, stream>>>(p);                         \
    }                                                                                         \
  } while (0)
  if (semiring == LT_LOG) LT_TF(LT_LOG);
  else if (semiring == LT_MAXTROPICAL) LT_TF(LT_MAXTROPICAL);
  else LT_TF(LT_REAL);
#undef LT_TF
  LT_LAUNCHED();
  return LT_OK;
}

int table_lattice_backward_launch(int semiring, const TableParams& p, cudaStream_t stream) {
  if (table2_backward_supported(p)) return table2_backward_launch(semiring, p, stream);
  const bool fld = p.k >= 1;
  const size_t smem = table_smem(p.C, fld, true);
  if (smem > 227 * 1024) {
    set_error("lt_table_lattice_backward: %d context states need %zu bytes of shared memory", p.C, smem);
    return LT_ERR_UNSUPPORTED;
  }
#define LT_TB(SR)                                                                             \
  do {                                                                                        \
    if (fld) {                                                                                \
      LT_CUDA(cudaFuncSetAttribute(table_backward_kernel<SR, true>,                           \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      table_backward_kernel<SR, true><<<p.B, 512, smem, stream>>>(p);                         \
    } else {                                                                                  \
      LT_CUDA(cudaFuncSetAttribute(table_backward_kernel<SR, false>,                          \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      table_backward_kernel<SR, false><<<p.B, 512, smem, stream>>>(p);                        \
    }                                                                                         \
  } while (0)
  if (semiring == LT_LOG) LT_TB(LT_LOG);
  else LT_TB(LT_REAL);
#undef LT_TB
  LT_LAUNCHED();
  return LT_OK;
}

int table_viterbi_launch(const TableParams& p, const float* alpha_final, int32_t* labels,
                         int32_t* path_states, float* grad_blank, float* grad_lexical,
                         cudaStream_t stream) {
  if (p.B == 0) return LT_OK;
  table_viterbi_kernel<<<(p.B + 31) / 32, 32, 0, stream>>>(p, alpha_final, labels, path_states,
                                                           grad_blank, grad_lexical);
  LT_LAUNCHED();
  return LT_OK;
}

int table_reduce_forward_launch(int semiring, const float* w, const int32_t* in_offsets,
                                const int32_t* in_arcs, int64_t outer, int C, int V, float* out,
                                int32_t* argarc, cudaStream_t stream) {
  const long long n = (long long)outer * C;
  if (n == 0) return LT_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (semiring == LT_LOG)
    table_reduce_forward_kernel<LT_LOG><<<grid, 256, 0, stream>>>(w, in_offsets, in_arcs, outer, C, V, out, argarc);
  else if (semiring == LT_MAXTROPICAL)
    table_reduce_forward_kernel<LT_MAXTROPICAL><<<grid, 256, 0, stream>>>(w, in_offsets, in_arcs, outer, C, V, out, argarc);
  else
    table_reduce_forward_kernel<LT_REAL><<<grid, 256, 0, stream>>>(w, in_offsets, in_arcs, outer, C, V, out, argarc);
  LT_LAUNCHED();
  return LT_OK;
}

int table_reduce_backward_launch(int semiring, const float* w, const float* out,
                                 const int32_t* argarc, const float* gout, const int32_t* table,
                                 int64_t outer, int C, int V, float* gw, cudaStream_t stream) {
  const long long n = (long long)outer * C * V;
  if (n == 0) return LT_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (semiring == LT_LOG)
    table_reduce_backward_kernel<LT_LOG><<<grid, 256, 0, stream>>>(w, out, argarc, gout, table, outer, C, V, gw);
  else if (semiring == LT_MAXTROPICAL)
    table_reduce_backward_kernel<LT_MAXTROPICAL><<<grid, 256, 0, stream>>>(w, out, argarc, gout, table, outer, C, V, gw);
  else
    table_reduce_backward_kernel<LT_REAL><<<grid, 256, 0, stream>>>(w, out, argarc, gout, table, outer, C, V, gw);
  LT_LAUNCHED();
  return LT_OK;
}

}  // namespace lt

// ------------------------------------------------------------------ C ABI ----
using namespace lt;

extern "C" {

static int check_table_args(const char* fn, int semiring, int C, int V, int k, int B, int T) {
  LT_CHECK_ARG(semiring == LT_REAL || semiring == LT_LOG || semiring == LT_MAXTROPICAL,
               "%s: unknown semiring %d", fn, semiring);
  LT_CHECK_ARG(C > 0 && V > 0, "%s: next_state_table should have a non-zero size, got [%d, %d]", fn, C, V);
  LT_CHECK_ARG((long long)C * V < (1ll << 31), "%s: %d x %d arcs exceed the int32 arc index", fn, C, V);
  LT_CHECK_ARG(k == LT_FRAME_DEPENDENT || (k >= 1 && k <= 254),
               "%s: max_expansions must be LT_FRAME_DEPENDENT or in [1, 254], got %d", fn, k);
  LT_CHECK_ARG(B >= 0 && T >= 0, "%s: negative batch (%d) or frame count (%d)", fn, B, T);
  return LT_OK;
}

int lt_table_lattice_cluster(int C, int V, int max_expansions, int backward) {
  if (C <= 0 || V <= 0) return 0;
  return table2_cluster_size(C, V, max_expansions, backward != 0);
}

int lt_table_lattice_forward(int semiring, int max_expansions, const int32_t* table,
                             const int32_t* in_offsets, const int32_t* in_arcs, int C, int V,
                             const float* blank, const float* lexical, const int32_t* num_frames,
                             int B, int T, const float* alpha_init, float* dist, float* alphas,
                             float* alpha_final, float* levels, int32_t* backarc,
                             uint8_t* termptr, void* stream) {
  int rc = check_table_args("lt_table_lattice_forward", semiring, C, V, max_expansions, B, T);
  if (rc) return rc;
  LT_CHECK_ARG(table && in_offsets && in_arcs && dist && num_frames,
               "lt_table_lattice_forward: NULL pointer");
  LT_CHECK_ARG(T == 0 || (blank && lexical), "lt_table_lattice_forward: blank/lexical must not be NULL");
  if (B == 0) return LT_OK;
  TableParams p = {};
  p.C = C; p.V = V; p.k = max_expansions; p.B = B; p.T = T;
  p.table = table; p.in_offsets = in_offsets; p.in_arcs = in_arcs;
  p.blank = blank; p.lexical = lexical; p.num_frames = num_frames; p.alpha_init = alpha_init;
  p.dist = dist; p.alphas = alphas; p.alpha_final = alpha_final;
  p.levels = max_expansions >= 1 ? levels : nullptr;
  p.backarc = semiring == LT_MAXTROPICAL ? backarc : nullptr;
  p.termptr = (semiring == LT_MAXTROPICAL && max_expansions >= 1) ? termptr : nullptr;
  return table_lattice_forward_launch(semiring, p, (cudaStream_t)stream);
}

int lt_table_lattice_backward(int semiring, int max_expansions, const int32_t* table, int C,
                              int V, const float* blank, const float* lexical,
                              const int32_t* num_frames, int B, int T, const float* alphas,
                              const float* levels, const float* dist, const float* grad_dist,
                              float* grad_blank, float* grad_lexical, void* stream) {
  int rc = check_table_args("lt_table_lattice_backward", semiring, C, V, max_expansions, B, T);
  if (rc) return rc;
  LT_CHECK_ARG(semiring != LT_MAXTROPICAL,
               "lt_table_lattice_backward: MaxTropical gradients come from lt_table_viterbi_backtrace");
  if (B == 0 || T == 0) return LT_OK;
  LT_CHECK_ARG(table && blank && lexical && num_frames && alphas && dist && grad_blank && grad_lexical,
               "lt_table_lattice_backward: NULL pointer");
  LT_CHECK_ARG(max_expansions < 1 || levels,
               "lt_table_lattice_backward: FrameLabelDependent needs the `levels` buffer of the forward");
  TableParams p = {};
  p.C = C; p.V = V; p.k = max_expansions; p.B = B; p.T = T;
  p.table = table; p.blank = blank; p.lexical = lexical; p.num_frames = num_frames;
  p.alphas_in = alphas; p.levels_in = levels; p.dist_in = dist; p.grad_dist = grad_dist;
  p.grad_blank = grad_blank; p.grad_lexical = grad_lexical;
  return table_lattice_backward_launch(semiring, p, (cudaStream_t)stream);
}

int lt_table_viterbi_backtrace(int max_expansions, int C, int V, const int32_t* backarc,
                               const uint8_t* termptr, const float* alpha_final,
                               const int32_t* num_frames, int B, int T, int32_t* labels,
                               int32_t* path_states, const float* grad_dist, float* grad_blank,
                               float* grad_lexical, void* stream) {
  int rc = check_table_args("lt_table_viterbi_backtrace", LT_MAXTROPICAL, C, V, max_expansions, B, T);
  if (rc) return rc;
  if (B == 0) return LT_OK;
  LT_CHECK_ARG(alpha_final && num_frames && labels && (T == 0 || backarc),
               "lt_table_viterbi_backtrace: NULL pointer");
  LT_CHECK_ARG(max_expansions < 1 || termptr || T == 0,
               "lt_table_viterbi_backtrace: FrameLabelDependent needs termptr");
  TableParams p = {};
  p.C = C; p.V = V; p.k = max_expansions; p.B = B; p.T = T;
  p.num_frames = num_frames; p.backarc = const_cast<int32_t*>(backarc);
  p.termptr = const_cast<uint8_t*>(termptr); p.grad_dist = grad_dist;
  return table_viterbi_launch(p, alpha_final, labels, path_states, grad_blank, grad_lexical,
                              (cudaStream_t)stream);
}

int lt_table_reduce_forward(int semiring, const float* w, const int32_t* in_offsets,
                            const int32_t* in_arcs, int64_t outer, int C, int V, float* out,
                            int32_t* argarc, void* stream) {
  int rc = check_table_args("lt_table_reduce_forward", semiring, C, V, LT_FRAME_DEPENDENT, 0, 0);
  if (rc) return rc;
  LT_CHECK_ARG(outer >= 0, "lt_table_reduce_forward: negative outer size");
  if (outer == 0) return LT_OK;
  LT_CHECK_ARG(w && in_offsets && in_arcs && out, "lt_table_reduce_forward: NULL pointer");
  LT_CHECK_ARG(semiring != LT_MAXTROPICAL || argarc,
               "lt_table_reduce_forward: MaxTropical needs the argarc output");
  return table_reduce_forward_launch(semiring, w, in_offsets, in_arcs, outer, C, V, out, argarc,
                                     (cudaStream_t)stream);
}

int lt_table_reduce_backward(int semiring, const float* w, const float* out,
                             const int32_t* argarc, const float* grad_out, const int32_t* table,
                             int64_t outer, int C, int V, float* grad_w, void* stream) {
  int rc = check_table_args("lt_table_reduce_backward", semiring, C, V, LT_FRAME_DEPENDENT, 0, 0);
  if (rc) return rc;
  if (outer <= 0) return LT_OK;
  LT_CHECK_ARG(w && out && grad_out && table && grad_w, "lt_table_reduce_backward: NULL pointer");
  LT_CHECK_ARG(semiring != LT_MAXTROPICAL || argarc,
               "lt_table_reduce_backward: MaxTropical needs argarc");
  return table_reduce_backward_launch(semiring, w, out, argarc, grad_out, table, outer, C, V,
                                      grad_w, (cudaStream_t)stream);
}

}  // extern "C"
