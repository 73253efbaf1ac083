// K2 (generic): backward (beta) recursion that writes arc posteriors directly
// as the weight gradient.  Implements the intent of
// RecognitionLattice._backward (/root/reference/last_torch/lattices.py:686-799)
// with alignment.backward (alignments.py:300-318 FrameDependent, :378-418
// FrameLabelDependent) and FullNGram.backward_broadcast (contexts.py:232-256),
// and replaces autograd through the unrolled forward loop.
//
// One thread-block cluster per utterance; every CTA keeps a full copy of
// beta_{t+1} in shared memory and owns a contiguous slice of SOURCE states
// (rows).  For a fixed source p the V destinations next(p, y) are contiguous,
// so a row is a coalesced stream of lexical[p, :] against a contiguous window
// of beta.  Row results (beta_t[p]) are all-gathered through DSMEM.
//
// Log:  grad_lex[p,y] = g * exp(alpha[p] + lex[p,y] + beta'[next] - logZ)
//       computed as  e = exp(x - m_p)  (the same exponential the row
//       log-sum-exp needs) times the per-row scalar exp(alpha[p] + m_p - logZ),
//       i.e. ONE exponential per arc for both beta and the posterior.
// Real: grad_lex[p,y] = g * alpha[p] * beta'[next].
#include "common.cuh"
#include "params.cuh"

namespace lt {


__device__ __forceinline__ void bcast_store_b(float* base, int idx, float v, uint32_t nrank) {
  const uint32_t a = smem_u32(base + idx);
  for (uint32_t r = 0; r < nrank; ++r) st_shared_cluster_f32(map_shared_rank(a, r), v);
}

template <int LPR>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = LPR >> 1; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Process the rows [p_lo, p_hi) of one frame against the destination vector
// `nb` (shared, full C).  For every row:
//   Log : lsum = logsumexp_y(lex[p,y] + nb[next(p,y)]);
//         grad_lex[p,y] (+)= exp(lex + nb - m) * exp(src_alpha[p] + m - logZ) * g
//   Real: lsum = sum_y lex[p,y] * nb[next(p,y)];
//         grad_lex[p,y] (+)= g * src_alpha[p] * nb[next(p,y)]
// row_out[p - p_lo] = lsum.  `accumulate` adds into grad_lex instead of storing.
template <int SR, int LPR>
__device__ __forceinline__ void rows_backward(
    const NGram& g, const float* __restrict__ lex, float* __restrict__ glex,
    const float* __restrict__ nb, const float* __restrict__ src_alpha,
    float logz, float gscale, bool scale_ok, int p_lo, int p_hi, bool accumulate,
    double* row_out) {
  constexpr int RPW = 32 / LPR;              // rows per warp pass
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int sub = lane / LPR, l = lane % LPR;
  const int V = g.V;
  const bool single = V <= 8 * LPR;
  for (int r0 = p_lo + warp * RPW; r0 < p_hi; r0 += nwarps * RPW) {
    const int p = r0 + sub;
    const bool active = p < p_hi;
    const int pc = active ? p : p_lo;
    const float* row = lex + (size_t)pc * V;
    float* grow = glex + (size_t)pc * V;
    const float* dst = nb + ngram_next(g, pc, 0);
    const int ds = g.pstride ? 1 : 0;   // n == 0: every arc leads to state 0
    const float a = src_alpha[pc];
    if constexpr (SR == LT_LOG) {
      // The sums w + beta'[next] are exact in double; the row maximum is only a shift (float),
      // and every exponent (w + beta') - ms is rounded ONCE, after the subtraction: an arc near
      // the maximum carries no rounding error of the sum (ulp(|w + beta'|) ~ 2e-6 at |w| ~ 30
      // when formed in float).  These kernels are latency-bound; B200 has full-rate-ish FP64.
      double x[8];
      float m = neg_inf();
      if (single) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int y = l + i * LPR;
          x[i] = (y < V) ? (double)ldg_stream(row + y) + (double)dst[y * ds] : (double)neg_inf();
          m = fmaxf(m, (float)x[i]);
        }
      } else {
        for (int y = l; y < V; y += LPR)
          m = fmaxf(m, (float)((double)ldg_stream(row + y) + (double)dst[y * ds]));
      }
      m = group_max<LPR>(m);
      const float ms = msafe(m);
      // per-row posterior scale; 0 when the lattice is unreachable (logZ = -inf)
      const float rs = scale_ok
          ? gscale * fast_exp((float)((double)a + (double)ms - (double)logz)) : 0.f;
      float s = 0.f;
      if (single) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int y = l + i * LPR;
          const float e = fast_exp((float)(x[i] - (double)ms));
          s += e;
          if (active && y < V) {
            const float gv = e * rs;
            grow[y] = accumulate ? grow[y] + gv : gv;
          }
        }
      } else {
        for (int y = l; y < V; y += LPR) {
          const float e = fast_exp(
              (float)((double)ldg_stream(row + y) + (double)dst[y * ds] - (double)ms));
          s += e;
          if (active) {
            const float gv = e * rs;
            grow[y] = accumulate ? grow[y] + gv : gv;
          }
        }
      }
      s = group_sum<LPR>(s);
      if (active && l == 0) row_out[p - p_lo] = (double)ms + (double)fast_log(s);
    } else {  // Real
      float s = 0.f;
      const float ga = gscale * a;
      for (int y = l; y < V; y += LPR) {
        const float bv = dst[y * ds];
        s += ldg_stream(row + y) * bv;
        if (active) {
          const float gv = ga * bv;
          grow[y] = accumulate ? grow[y] + gv : gv;
        }
      }
      s = group_sum<LPR>(s);
      if (active && l == 0) row_out[p - p_lo] = (double)s;
    }
  }
}

// log(exp(a) + exp(b)) - shift for exact double arguments, rounded once (the non-finite-max rule
// of semirings.py:250-251).
__device__ __forceinline__ float logaddexp_shifted_d(double a, double b, float shift) {
  const double c = a > b ? a : b;
  const double cs = (c == c && c - c == 0.0) ? c : 0.0;          // finite ? c : 0
  const float z = fast_exp((float)(a - cs)) + fast_exp((float)(b - cs));
  return (float)((cs - (double)shift) + (double)fast_log(z));
}

template <int SR, bool FLD, int LPR>
__global__ void __launch_bounds__(512)
lattice_backward_generic(const BwdParams p) {
  using S = Sr<SR>;
  extern __shared__ __align__(16) float smem[];
  const NGram& g = p.g;
  const int C = g.C;
  const int Cp = (C + 3) & ~3;
  const uint32_t nrank = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int b = blockIdx.x / nrank;
  const int tid = threadIdx.x, nth = blockDim.x;

  float* buf0 = smem;
  float* buf1 = buf0 + Cp;
  float* buf2 = buf1 + Cp;                      // FLD only
  double* row_out = reinterpret_cast<double*>(buf2 + (FLD ? Cp : 0));   // [dslice], 8-byte aligned

  const int p_lo = min(C, (int)rank * p.dslice);
  const int p_hi = min(C, p_lo + p.dslice);
  const int D = p_hi - p_lo;
  const int V = g.V;

  int nf = p.num_frames[b];
  nf = max(0, min(nf, p.T));
  const size_t bt0 = (size_t)b * p.T;
  const float logz_plain = p.dist[b];
  const float gscale = p.grad_dist ? p.grad_dist[b] : 1.f;
  const bool scale_ok = (SR != LT_LOG) || is_finite(logz_plain);
  // Renormalised pair (lt_lattice_backward_norm): alphas hold alpha~_t, logZ = off_T + r;
  // beta~_t = beta_t - (off_T - off_t) follows the forward's shifts d_t = off_{t+1} - off_t and
  // every posterior exponent becomes alpha~ + w + beta~ - (r + d_t).
  const bool norm = SR == LT_LOG && p.alpha_norm != nullptr;
  const int32_t* an = norm ? p.alpha_norm + (size_t)b * (p.T + 3) : nullptr;
  const float logz_res = norm ? __int_as_float(an[p.T + 1]) : logz_plain;

  const int LW = p.wlevels > 1 ? p.wlevels : 1;             // weight / gradient sets per frame
  const size_t lvb = LW > 1 ? (size_t)C : 0;
  const size_t lvl = LW > 1 ? (size_t)C * V : 0;
  // padding frames: zero gradients (lattices.py:775-779)
  for (int t = nf; t < p.T; ++t) {
    for (int lv = 0; lv < LW; ++lv) {
      float* gb = p.grad_blank + ((bt0 + t) * LW + lv) * C;
      float* gl = p.grad_lexical + ((bt0 + t) * LW + lv) * (size_t)C * V;
      for (int d = tid; d < D; d += nth) gb[p_lo + d] = 0.f;
      const size_t n = (size_t)D * V;
      float* base = gl + (size_t)p_lo * V;
      for (size_t i = tid; i < n; i += nth) base[i] = 0.f;
    }
  }

  float* beta = buf0;      // beta_{t+1}
  float* nxt = buf1;
  float* spare = buf2;
  for (int c = tid; c < C; c += nth) beta[c] = S::one();   // lattices.py:789-790
  __syncthreads();
  cluster_sync_all();

  for (int t = nf - 1; t >= 0; --t) {
    const float shift = norm ? (float)(an[t + 1] - an[t]) : 0.f;
    const float logz = logz_res + shift;
    const float* blank = p.blank + (bt0 + t) * C * LW;
    const float* lex = p.lexical + (bt0 + t) * (size_t)C * V * LW;
    const float* alpha = p.alphas + (bt0 + t) * C;
    float* gb = p.grad_blank + (bt0 + t) * C * LW;
    float* gl = p.grad_lexical + (bt0 + t) * (size_t)C * V * LW;
    if constexpr (!FLD) {
      rows_backward<SR, LPR>(g, lex, gl, beta, alpha, logz, gscale, scale_ok, p_lo, p_hi, false, row_out);
      __syncthreads();
      for (int d = tid; d < D; d += nth) {
        const int q = p_lo + d;
        float v;
        if constexpr (SR == LT_LOG) {
          const double bb = (double)blank[q] + (double)beta[q];
          v = logaddexp_shifted_d(bb, row_out[d], shift);
          gb[q] = scale_ok
              ? gscale * fast_exp((float)((double)alpha[q] + bb - (double)logz)) : 0.f;
        } else {
          v = blank[q] * beta[q] + (float)row_out[d];
          gb[q] = gscale * alpha[q] * beta[q];
        }
        bcast_store_b(nxt, q, v, nrank);
      }
      cluster_sync_all();
      float* tmp = beta; beta = nxt; nxt = tmp;
    } else {
      const int k = p.k;
      const float* lev = p.levels + (bt0 + t) * (size_t)k * C;   // last_1..last_k
      // nb_k = blank (x) beta'  (alignments.py:405), computed redundantly by every CTA
      float* nb = nxt;
      for (int c = tid; c < C; c += nth) nb[c] = S::times(blank[k * lvb + c], beta[c]);
      // blank marginals (alignments.py:398-403)
      for (int d = tid; d < D; d += nth) {
        const int q = p_lo + d;
        float acc = 0.f;
        if constexpr (SR == LT_LOG) {
          if (scale_ok) {
            // level i: exp(lexical_alphas[i] + blank[i] + beta' - logZ)  (alignments.py:398-403)
            for (int i = 0; i <= k; ++i) {
              const double src = i == 0 ? (double)alpha[q] : (double)lev[(size_t)(i - 1) * C + q];
              const float e = gscale * fast_exp((float)(
                  src + (double)blank[i * lvb + q] + (double)beta[q] - (double)logz));
              if (LW > 1) gb[i * lvb + q] = e; else acc += e;
            }
          } else if (LW > 1) {
            for (int i = 0; i <= k; ++i) gb[i * lvb + q] = 0.f;
          }
        } else {
          for (int i = 0; i <= k; ++i) {
            const float e = gscale * beta[q] * (i == 0 ? alpha[q] : lev[(size_t)(i - 1) * C + q]);
            if (LW > 1) gb[i * lvb + q] = e; else acc += e;
          }
        }
        if (LW == 1) gb[q] = acc;
      }
      if (LW > 1) {       // lexical[k] is never used (alignments.py:417): its gradient is zero
        float* base = gl + k * lvl + (size_t)p_lo * V;
        for (size_t i = tid; i < (size_t)D * V; i += nth) base[i] = 0.f;
      }
      __syncthreads();
      float* out = spare;
      for (int j = k - 1; j >= 0; --j) {
        const float* src_alpha = (j == 0) ? alpha : lev + (size_t)(j - 1) * C;
        rows_backward<SR, LPR>(g, lex + j * lvl, gl + j * lvl, nb, src_alpha, logz, gscale,
                               scale_ok, p_lo, p_hi, LW == 1 && j != k - 1, row_out);
        __syncthreads();
        for (int d = tid; d < D; d += nth) {
          const int q = p_lo + d;
          float v;                                 // alignments.py:414-415
          if constexpr (SR == LT_LOG)              // nb_0 is beta_t: move it to the frame of off_t
            v = logaddexp_shifted_d((double)blank[j * lvb + q] + (double)beta[q], row_out[d],
                                    j == 0 ? shift : 0.f);
          else
            v = blank[j * lvb + q] * beta[q] + (float)row_out[d];
          bcast_store_b(out, q, v, nrank);
        }
        cluster_sync_all();
        float* tmp = nb; nb = out; out = tmp;
      }
      // nb now holds beta_t; the old beta' and `out` become the scratch pair
      float* old = beta;
      beta = nb; nxt = old; spare = out;
    }
  }
  if (p.beta_final)
    for (int d = tid; d < D; d += nth)
      p.beta_final[(size_t)b * C + p_lo + d] =
          norm ? (float)((double)beta[p_lo + d] + (double)(an[p.T] - an[0])) : beta[p_lo + d];
  cluster_sync_all();
}

// ---------------------------------------------------------------------------

template <typename KernelT>
static int launch_cluster_b(KernelT kernel, int grid, int block, size_t smem, int cluster,
                            cudaStream_t stream, const BwdParams& p) {
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

template <int SR, bool FLD>
static int dispatch_lpr(int lpr, int grid, int block, size_t smem, int cluster,
                        cudaStream_t stream, const BwdParams& p) {
  switch (lpr) {
    case 1: return launch_cluster_b(lattice_backward_generic<SR, FLD, 1>, grid, block, smem, cluster, stream, p);
    case 2: return launch_cluster_b(lattice_backward_generic<SR, FLD, 2>, grid, block, smem, cluster, stream, p);
    case 4: return launch_cluster_b(lattice_backward_generic<SR, FLD, 4>, grid, block, smem, cluster, stream, p);
    case 8: return launch_cluster_b(lattice_backward_generic<SR, FLD, 8>, grid, block, smem, cluster, stream, p);
    case 16: return launch_cluster_b(lattice_backward_generic<SR, FLD, 16>, grid, block, smem, cluster, stream, p);
    default: return launch_cluster_b(lattice_backward_generic<SR, FLD, 32>, grid, block, smem, cluster, stream, p);
  }
}

int lattice_backward_generic_launch(int semiring, const NGram& g, int k, const BwdParams& base,
                                    unsigned flags, int sm_count, cudaStream_t stream) {
  BwdParams p = base;
  const bool fld = k >= 1;
  const int cluster = pick_cluster_size(g, p.B, flags, sm_count);
  p.dslice = (g.C + cluster - 1) / cluster;
  int lpr = 1;
  while (lpr < 32 && lpr * 8 < g.V) lpr *= 2;
  p.lpr = lpr;
  const long long work = (long long)p.dslice * g.V;
  const int block = work >= 8192 ? 512 : (work >= 1024 ? 256 : 128);
  const int Cp = (g.C + 3) & ~3;
  size_t smem = sizeof(float) * ((size_t)Cp * (fld ? 3 : 2) + 2 * (size_t)p.dslice);   // row_out: double
  if (smem > 227 * 1024) {
    set_error("lt_lattice_backward: %d context states need %zu bytes of shared memory per CTA "
              "(limit 232448)", g.C, smem);
    return LT_ERR_UNSUPPORTED;
  }
  const int grid = p.B * cluster;
  if (semiring == LT_LOG)
    return fld ? dispatch_lpr<LT_LOG, true>(lpr, grid, block, smem, cluster, stream, p)
               : dispatch_lpr<LT_LOG, false>(lpr, grid, block, smem, cluster, stream, p);
  if (semiring == LT_REAL)
    return fld ? dispatch_lpr<LT_REAL, true>(lpr, grid, block, smem, cluster, stream, p)
               : dispatch_lpr<LT_REAL, false>(lpr, grid, block, smem, cluster, stream, p);
  set_error("lt_lattice_backward: semiring must be LT_LOG or LT_REAL (MaxTropical gradients come "
            "from lt_viterbi_backtrace), got %d", semiring);
  return LT_ERR_INVALID_ARGUMENT;
}

}  // namespace lt
