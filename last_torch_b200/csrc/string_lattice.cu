// K3: numerator lattice = recognition lattice intersected with the reference
// label string (RecognitionLattice._string_forward,
// /root/reference/last_torch/lattices.py:250-377).
//
//  * string_gather / string_scatter_add: weight_step_scan + gather_weight
//    (lattices.py:300-342, :830-845) and its transpose.
//  * string_forward: shortest_distance_step_scan (lattices.py:347-377) with
//    alignment.string_forward (alignments.py:327-329 / :427-432).
//  * string_backward: chain forward-backward giving d numerator / d weights
//    (the reference relies on autograd here, broken as shipped: SURVEY D1/D2).
//
// The chain has only U+1 states, so one CTA per utterance with one thread per
// chain state; alpha/beta live in shared memory for all T frames.
#include "common.cuh"
#include "params.cuh"
#include "umma.cuh"

namespace lt {

// ------------------------------------------------------------------ gather --
__global__ void string_gather_kernel(int V, int C, const float* __restrict__ blank,
                                     const float* __restrict__ lexical,
                                     const int32_t* __restrict__ states,
                                     const int32_t* __restrict__ labels, int T, int U1,
                                     float* __restrict__ blank_w, float* __restrict__ lexical_w) {
  const size_t bt = blockIdx.x;           // b * T + t
  const int b = (int)(bt / T);
  const float* bl = blank + bt * C;
  const float* lx = lexical + bt * (size_t)C * V;
  for (int u = threadIdx.x; u < U1; u += blockDim.x) {
    // states / labels from lt_walk_states are in range; anything else is clamped (memory safety)
    const int s = min(max(states[(size_t)b * U1 + u], 0), C - 1);
    const int y = min(max(labels[(size_t)b * U1 + u] - 1, 0), V - 1);
    blank_w[bt * U1 + u] = bl[s];
    lexical_w[bt * U1 + u] = lx[(size_t)s * V + y];
  }
}

__global__ void string_scatter_kernel(int V, int C, const float* __restrict__ gbw,
                                      const float* __restrict__ glw,
                                      const int32_t* __restrict__ states,
                                      const int32_t* __restrict__ labels, int T, int U1,
                                      float scale, const float* __restrict__ utt_scale,
                                      float* __restrict__ gblank,
                                      float* __restrict__ glex) {
  const size_t bt = blockIdx.x;
  const int b = (int)(bt / T);
  if (utt_scale) scale *= utt_scale[b];
  if (scale == 0.f) return;
  float* bl = gblank + bt * C;
  float* lx = glex + bt * (size_t)C * V;
  for (int u = threadIdx.x; u < U1; u += blockDim.x) {
    const int s = min(max(states[(size_t)b * U1 + u], 0), C - 1);
    const int y = min(max(labels[(size_t)b * U1 + u] - 1, 0), V - 1);
    const float a = gbw[bt * U1 + u], c = glw[bt * U1 + u];
    if (a != 0.f) atomicAdd(bl + s, scale * a);
    if (c != 0.f) atomicAdd(lx + (size_t)s * V + y, scale * c);
  }
}

// The same for a split-row grad_lexical (rows of [V bf16 hi | V bf16 lo], see
// lt_joint_backward): there is no atomic on a (hi, lo) pair that lives in two places, so the
// block first merges the label positions that hit the same arc -- position u owns its target if
// no earlier position has it, and adds the contributions of all later duplicates -- and every
// owner then does ONE plain read-modify-write (value = hi + lo, re-split).  Blocks handle
// different frames, owners different arcs: no two threads touch the same 16-bit word.
__global__ void string_scatter_split_kernel(int V, int C, const float* __restrict__ gbw,
                                            const float* __restrict__ glw,
                                            const int32_t* __restrict__ states,
                                            const int32_t* __restrict__ labels, int T, int U1,
                                            float scale, const float* __restrict__ utt_scale,
                                            float* __restrict__ gblank,
                                            unsigned char* __restrict__ glex) {
  extern __shared__ int32_t s_key[];               // [U1] arc index s * V + y
  float* s_val = reinterpret_cast<float*>(s_key + U1);
  const size_t bt = blockIdx.x;
  const int b = (int)(bt / T);
  if (utt_scale) scale *= utt_scale[b];
  if (scale == 0.f) return;
  float* bl = gblank + bt * C;
  unsigned char* lx = glex + bt * (size_t)C * V * 4;
  for (int u = threadIdx.x; u < U1; u += blockDim.x) {
    const int s = min(max(states[(size_t)b * U1 + u], 0), C - 1);
    const int y = min(max(labels[(size_t)b * U1 + u] - 1, 0), V - 1);
    s_key[u] = s * V + y;
    s_val[u] = scale * glw[bt * U1 + u];
    const float a = gbw[bt * U1 + u];
    if (a != 0.f) atomicAdd(bl + s, scale * a);
  }
  __syncthreads();
  for (int u = threadIdx.x; u < U1; u += blockDim.x) {
    const int key = s_key[u];
    bool owner = true;
    for (int w = 0; w < u; ++w) owner = owner && s_key[w] != key;
    if (!owner) continue;
    float add = s_val[u];
    for (int w = u + 1; w < U1; ++w) add += s_key[w] == key ? s_val[w] : 0.f;
    if (add == 0.f) continue;
    const int s = key / V, y = key - s * V;
    unsigned short* hi = reinterpret_cast<unsigned short*>(lx + (size_t)s * V * 4) + y;
    unsigned short* lo = hi + V;
    const float v = __uint_as_float((uint32_t)*hi << 16) + __uint_as_float((uint32_t)*lo << 16) + add;
    __nv_bfloat16 h, l;
    umma::split_bf16(v, h, l);
    *hi = __bfloat16_as_ushort(h);
    *lo = __bfloat16_as_ushort(l);
  }
}

// ContextDependency.walk_states for FullNGram (contexts.py:109-146 with
// next_state :190-205) plus the "next label" row of lattices.py:336-338 / :314-315.
// One thread per utterance: U sequential integer steps.
// Positions u >= num_labels[b] (when given) are padding and read as epsilon whatever they hold:
// they cannot influence alpha[num_labels].  A label outside [0, V] before that is an error of the
// caller (the reference fails in one_hot, lattices.py:322): it is counted in *bad and read as
// epsilon, so every state / label this kernel emits is in range and the gather / scatter kernels
// never leave the [C, V] frame.
__global__ void walk_states_kernel(NGram g, const int32_t* __restrict__ labels,
                                   const int32_t* __restrict__ num_labels, int B, int U,
                                   int32_t* __restrict__ states, int32_t* __restrict__ next_labels,
                                   int32_t* __restrict__ bad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int U1 = U + 1;
  const int nl = num_labels ? num_labels[b] : U;
  int s = 0, nbad = 0;
  states[(size_t)b * U1] = 0;
  for (int u = 0; u < U; ++u) {
    int y = labels[(size_t)b * U + u];
    if (u >= nl) y = 0;
    if (y < 0 || y > g.V) { ++nbad; y = 0; }
    if (y != 0) s = ngram_next(g, s, y - 1);          // epsilon (0) stays in place
    states[(size_t)b * U1 + u + 1] = s;
    next_labels[(size_t)b * U1 + u] = y < 1 ? 1 : y;  // label 0 is read as label 1
  }
  next_labels[(size_t)b * U1 + U] = 1;
  if (nbad && bad) atomicAdd(bad, nbad);
}

int walk_states_launch(const NGram& g, const int32_t* labels, const int32_t* num_labels, int B,
                       int U, int32_t* states, int32_t* next_labels, int32_t* bad,
                       cudaStream_t stream) {
  if (B == 0) return LT_OK;
  walk_states_kernel<<<(B + 127) / 128, 128, 0, stream>>>(g, labels, num_labels, B, U, states,
                                                          next_labels, bad);
  LT_LAUNCHED();
  return LT_OK;
}

// One thread that sleeps: put at the head of a side stream so that the kernels behind it start
// a few tens of microseconds after a kernel on another stream that became launchable at the same
// moment (see lt_stream_delay in include/last_lattice.h).
__global__ void stream_delay_kernel(unsigned ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(1000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < ns);
}

int stream_delay_launch(unsigned ns, cudaStream_t stream) {
  if (ns == 0) return LT_OK;
  stream_delay_kernel<<<1, 1, 0, stream>>>(ns > 1000000u ? 1000000u : ns);
  LT_LAUNCHED();
  return LT_OK;
}

// ------------------------------------------------------------------ forward --

// last_i[u] = alpha[u-i] (x) lex[u-i] (x) ... (x) lex[u-1], associated in the
// reference's order (alignments.py:429-430); semiring zero if u < i.
template <int SR>
__device__ __forceinline__ float chain_last(const float* alpha, const float* lx, int u, int i) {
  using S = Sr<SR>;
  if (u < i) return S::zero();
  float v = alpha[u - i];
  for (int j = i; j >= 1; --j) v = S::times(v, lx[u - j]);
  return v;
}

template <int SR, bool FLD>
__global__ void string_forward_kernel(const StrParams p) {
  using S = Sr<SR>;
  extern __shared__ float smem[];
  const int b = blockIdx.x, U1 = p.U1;
  float* a0 = smem;
  float* a1 = smem + U1;
  int nf = max(0, min(p.num_frames[b], p.T));
  for (int u = threadIdx.x; u < U1; u += blockDim.x) a0[u] = (u == 0) ? S::one() : S::zero();
  __syncthreads();
  float* cur = a0; float* nxt = a1;
  for (int t = 0; t < p.T; ++t) {
    const size_t off = ((size_t)b * p.T + t) * U1;
    if (p.alphas)
      for (int u = threadIdx.x; u < U1; u += blockDim.x) p.alphas[off + u] = cur[u];
    if (t >= nf) continue;       // uniform per block
    const float* bl = p.blank_w + off;
    const float* lx = p.lexical_w + off;
    for (int u = threadIdx.x; u < U1; u += blockDim.x) {
      if constexpr (!FLD) {
        const float a = S::times(cur[u], bl[u]);
        const float l = (u > 0) ? S::times(cur[u - 1], lx[u - 1]) : S::zero();
        if constexpr (SR == LT_MAXTROPICAL) {
          const bool tb = a >= l;
          nxt[u] = tb ? a : l;
          if (p.backptr) p.backptr[off + u] = tb ? 0 : 1;
        } else {
          nxt[u] = S::plus(a, l);
        }
      } else {
        Acc<SR> acc; acc.init();
        const float bu = bl[u];
        for (int i = 0; i <= p.k; ++i)
          acc.add(S::times(chain_last<SR>(cur, lx, u, i), bu), i);
        nxt[u] = acc.value();
        if constexpr (SR == LT_MAXTROPICAL) { if (p.backptr) p.backptr[off + u] = (uint8_t)acc.arg(); }
      }
    }
    __syncthreads();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  if (threadIdx.x == 0) {
    const int nl = p.num_labels[b];
    p.dist[b] = (nl >= 0 && nl < U1) ? cur[nl] : S::zero();   // lattices.py:375-377
  }
}

// ----------------------------------------------------------------- backward --
template <int SR, bool FLD>
__global__ void string_backward_kernel(const StrParams p) {
  using S = Sr<SR>;
  extern __shared__ float smem[];
  const int b = blockIdx.x, U1 = p.U1;
  float* b0 = smem;
  float* b1 = smem + U1;
  float* n0 = smem + 2 * U1;    // FLD level buffers
  float* n1 = smem + 3 * U1;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const int nl = p.num_labels[b];
  const float z = p.dist_in[b];
  const float g = p.grad_dist ? p.grad_dist[b] : 1.f;
  const bool reachable = (nl >= 0 && nl < U1) &&
                         (SR == LT_REAL ? true : is_finite(z));
  const size_t base = (size_t)b * p.T * U1;
  // zero everything first (padding frames, unreachable utterances, MaxTropical)
  const int t_zero_from = (reachable && SR != LT_MAXTROPICAL) ? nf : 0;
  for (size_t i = (size_t)t_zero_from * U1 + threadIdx.x; i < (size_t)p.T * U1; i += blockDim.x) {
    p.grad_blank_w[base + i] = 0.f;
    p.grad_lexical_w[base + i] = 0.f;
  }
  if (!reachable) return;

  if constexpr (SR == LT_MAXTROPICAL) {
    __syncthreads();
    if (threadIdx.x == 0) {
      int u = nl;
      for (int t = nf - 1; t >= 0; --t) {
        const size_t off = base + (size_t)t * U1;
        const int i = p.backptr_in[off + u];
        // FrameLabelDependent: i lexical arcs, then the frame-closing blank arc.
        // FrameDependent: exactly one arc per frame, blank (0) or lexical (1).
        if (FLD || i == 0) p.grad_blank_w[off + u] += g;
        for (int j = 1; j <= i; ++j) p.grad_lexical_w[off + u - j] += g;
        u -= i;
      }
    }
    return;
  } else {
    for (int u = threadIdx.x; u < U1; u += blockDim.x) b0[u] = (u == nl) ? S::one() : S::zero();
    __syncthreads();
    float* beta = b0; float* nxt = b1;
    for (int t = nf - 1; t >= 0; --t) {
      const size_t off = base + (size_t)t * U1;
      const float* bl = p.blank_w + off;
      const float* lx = p.lexical_w + off;
      const float* al = p.alphas_in + off;
      if constexpr (!FLD) {
        for (int u = threadIdx.x; u < U1; u += blockDim.x) {
          const float bn = (u + 1 < U1) ? beta[u + 1] : S::zero();
          const float bb = S::times(bl[u], beta[u]);
          const float lb = S::times(lx[u], bn);
          if constexpr (SR == LT_LOG) {
            p.grad_blank_w[off + u] = g * fast_exp(al[u] + bb - z);
            p.grad_lexical_w[off + u] = g * fast_exp(al[u] + lb - z);
          } else {
            p.grad_blank_w[off + u] = g * al[u] * beta[u];
            p.grad_lexical_w[off + u] = g * al[u] * bn;
          }
          nxt[u] = S::plus(bb, lb);
        }
        __syncthreads();
      } else {
        const int k = p.k;
        float* nb = n0; float* out = n1;
        for (int u = threadIdx.x; u < U1; u += blockDim.x) {
          const float bb = S::times(bl[u], beta[u]);
          nb[u] = bb;
          float acc = 0.f;
          for (int i = 0; i <= k; ++i) {
            const float li = chain_last<SR>(al, lx, u, i);
            if constexpr (SR == LT_LOG) acc += fast_exp(li + bb - z);
            else acc += li;
          }
          if constexpr (SR == LT_LOG) p.grad_blank_w[off + u] = g * acc;
          else p.grad_blank_w[off + u] = g * acc * beta[u];
        }
        __syncthreads();
        for (int j = k - 1; j >= 0; --j) {
          for (int u = threadIdx.x; u < U1; u += blockDim.x) {
            const float bn = (u + 1 < U1) ? nb[u + 1] : S::zero();
            const float lb = S::times(lx[u], bn);
            const float lj = chain_last<SR>(al, lx, u, j);
            float gv;
            if constexpr (SR == LT_LOG) gv = g * fast_exp(lj + lb - z);
            else gv = g * lj * bn;
            if (j == k - 1) p.grad_lexical_w[off + u] = gv;
            else p.grad_lexical_w[off + u] += gv;
            out[u] = S::plus(S::times(bl[u], beta[u]), lb);
          }
          __syncthreads();
          float* tmp = nb; nb = out; out = tmp;
        }
        for (int u = threadIdx.x; u < U1; u += blockDim.x) nxt[u] = nb[u];
        __syncthreads();
      }
      float* tmp = beta; beta = nxt; nxt = tmp;
    }
  }
}

// ---------------------------------------------------------------------------
// Log chain with a DOUBLE state and the (integer part, fraction) storage of
// lt_string_forward_norm, for the shapes the register kernels below do not cover
// (FrameLabelDependent, U + 1 > 1024).  The state and every sum alpha + w / w + beta are double,
// so nothing is lost to the numerator's magnitude (|alpha| ~ 1e3 at T = 1000); the
// transcendental part of a logaddexp only ever sees the O(1) DIFFERENCE of its arguments and is
// evaluated in float with the accurate expf / log1pf (a few 1e-8 per step, unbiased; the bare
// MUFU forms accumulate ~1e-5 over k + 1 terms x 300 frames).  One thread
// per chain state, the frame's weights prefetched into registers one frame ahead, one
// __syncthreads per frame (state and the frame's lexical weights are double-buffered).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double dneg_inf() { return __longlong_as_double(0xfff0000000000000ll); }
__device__ __forceinline__ double dlogaddexp(double a, double b) {
  const double c = fmax(a, b);
  if (!(c > dneg_inf())) return c;                  // both are the semiring zero
  const float d = (float)(fmin(a, b) - c);          // <= 0, O(1) where it matters
  return c + (double)log1pf(expf(d));
}
__device__ __forceinline__ void ext_split(double v, int32_t& e, float& f) {
  if (!(v > dneg_inf())) { e = 0; f = neg_inf(); return; }
  double k = floor(v);
  k = fmin(fmax(k, -16777216.0), 16777216.0);
  e = (int32_t)k;
  f = (float)(v - k);
}
__device__ __forceinline__ double ext_join(int32_t e, float f) {
  return f > neg_inf() ? (double)e + (double)f : dneg_inf();
}
// last_i[u] = alpha[u-i] + lex[u-i] + ... + lex[u-1]  (alignments.py:429-430)
__device__ __forceinline__ double chain_last_d(const double* alpha, const float* lx, int u, int i) {
  if (u < i) return dneg_inf();
  double v = alpha[u - i];
  for (int j = i; j >= 1; --j) v += (double)lx[u - j];
  return v;
}

// shared memory: double a[2][U1]; float lxs[2][U1]
template <bool FLD>
__global__ void string_forward_ext_kernel(const StrParams p) {
  extern __shared__ double dsm[];
  const int b = blockIdx.x, U1 = p.U1, nth = blockDim.x;
  double* a0 = dsm;
  double* a1 = dsm + U1;
  float* l0 = reinterpret_cast<float*>(dsm + 2 * U1);
  float* l1 = l0 + U1;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t base = (size_t)b * p.T * U1;
  const bool single = U1 <= nth;                    // one chain state per thread: prefetch
  const int u1 = threadIdx.x;
  for (int u = threadIdx.x; u < U1; u += nth) a0[u] = (u == 0) ? 0.0 : dneg_inf();
  float pb = 0.f, pl = 0.f;
  if (single && u1 < U1 && nf > 0) { pb = ldg_stream(p.blank_w + base + u1); pl = ldg_stream(p.lexical_w + base + u1); }
  double* cur = a0; double* nxt = a1;
  float* lc = l0; float* ln = l1;
  for (int t = 0; t < nf; ++t) {
    const size_t off = base + (size_t)t * U1;
    const float wb = pb, wl = pl;
    if (single) {
      if (u1 < U1) {
        lc[u1] = wl;
        if (t + 1 < nf) { pb = ldg_stream(p.blank_w + off + U1 + u1); pl = ldg_stream(p.lexical_w + off + U1 + u1); }
      }
    } else {
      for (int u = threadIdx.x; u < U1; u += nth) lc[u] = p.lexical_w[off + u];
    }
    __syncthreads();                                // alpha_t and lexical_t visible
    for (int u = threadIdx.x; u < U1; u += nth) {
      int32_t e; float f;
      ext_split(cur[u], e, f);
      if (p.alphas) p.alphas[off + u] = f;
      p.alpha_exp[off + u] = e;
      const double bu = (double)(single ? wb : p.blank_w[off + u]);
      if constexpr (!FLD) {
        const double l = (u > 0) ? cur[u - 1] + (double)lc[u - 1] : dneg_inf();
        nxt[u] = dlogaddexp(cur[u] + bu, l);
      } else {
        double acc = cur[u];
        for (int i = 1; i <= p.k; ++i) acc = dlogaddexp(acc, chain_last_d(cur, lc, u, i));
        nxt[u] = acc + bu;
      }
    }
    double* tmp = cur; cur = nxt; nxt = tmp;
    float* lt = lc; lc = ln; ln = lt;
  }
  __syncthreads();
  for (int t = nf; t < p.T; ++t) {                  // padding frames keep alpha
    const size_t off = base + (size_t)t * U1;
    for (int u = threadIdx.x; u < U1; u += nth) {
      int32_t e; float f;
      ext_split(cur[u], e, f);
      if (p.alphas) p.alphas[off + u] = f;
      p.alpha_exp[off + u] = e;
    }
  }
  if (threadIdx.x == 0) {
    const int nl = p.num_labels[b];
    const double v = (nl >= 0 && nl < U1) ? cur[nl] : dneg_inf();   // lattices.py:375-377
    int32_t e; float f;
    ext_split(v, e, f);
    p.dist[b] = (float)v;
    p.dist_norm[2 * b] = e;
    p.dist_norm[2 * b + 1] = __float_as_int(f);
  }
}

// shared memory: double beta[2][U1], nb[2][U1], ad[U1]; float lxs[U1], bls[U1]
template <bool FLD>
__global__ void string_backward_ext_kernel(const StrParams p) {
  extern __shared__ double dsm[];
  const int b = blockIdx.x, U1 = p.U1, nth = blockDim.x;
  double* b0 = dsm;
  double* b1 = dsm + U1;
  double* n0 = dsm + 2 * U1;    // FLD level buffers
  double* n1 = dsm + 3 * U1;
  double* ad = dsm + 4 * U1;    // alpha_t
  float* lxs = reinterpret_cast<float*>(dsm + 5 * U1);
  float* bls = lxs + U1;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const int nl = p.num_labels[b];
  const float zf = __int_as_float(p.dist_norm_in[2 * b + 1]);
  const double z = ext_join(p.dist_norm_in[2 * b], zf);
  const float g = p.grad_dist ? p.grad_dist[b] : 1.f;
  const bool reachable = (nl >= 0 && nl < U1) && is_finite(zf);
  const size_t base = (size_t)b * p.T * U1;
  for (size_t i = (size_t)(reachable ? nf : 0) * U1 + threadIdx.x; i < (size_t)p.T * U1; i += nth) {
    p.grad_blank_w[base + i] = 0.f;
    p.grad_lexical_w[base + i] = 0.f;
  }
  if (!reachable) return;
  const bool single = U1 <= nth;
  const int u1 = threadIdx.x;
  for (int u = threadIdx.x; u < U1; u += nth) b0[u] = (u == nl) ? 0.0 : dneg_inf();
  float pb = 0.f, pl = 0.f, pa = 0.f;
  int32_t pe = 0;
  auto prefetch = [&](int t) {
    const size_t o = base + (size_t)t * U1 + u1;
    pb = ldg_stream(p.blank_w + o); pl = ldg_stream(p.lexical_w + o);
    pa = p.alphas_in[o]; pe = p.alpha_exp_in[o];
  };
  if (single && u1 < U1 && nf > 0) prefetch(nf - 1);
  double* beta = b0; double* nxt = b1;
  // posterior = exp(exponent): the exponent is formed in double, the exponential in float
  auto post = [&](double x) { return g * expf((float)x); };
  for (int t = nf - 1; t >= 0; --t) {
    const size_t off = base + (size_t)t * U1;
    if (single) {
      if (u1 < U1) {
        ad[u1] = ext_join(pe, pa); lxs[u1] = pl; bls[u1] = pb;
        if (t > 0) prefetch(t - 1);
      }
    } else {
      for (int u = threadIdx.x; u < U1; u += nth) {
        ad[u] = ext_join(p.alpha_exp_in[off + u], p.alphas_in[off + u]);
        lxs[u] = p.lexical_w[off + u];
        bls[u] = p.blank_w[off + u];
      }
    }
    __syncthreads();                                // beta_{t+1}, alpha_t, weights visible
    if constexpr (!FLD) {
      for (int u = threadIdx.x; u < U1; u += nth) {
        const double bn = (u + 1 < U1) ? beta[u + 1] : dneg_inf();
        const double bb = (double)bls[u] + beta[u];
        const double lb = (double)lxs[u] + bn;
        p.grad_blank_w[off + u] = post(ad[u] + bb - z);
        p.grad_lexical_w[off + u] = post(ad[u] + lb - z);
        nxt[u] = dlogaddexp(bb, lb);
      }
      __syncthreads();
    } else {
      const int k = p.k;
      double* nb = n0; double* out = n1;
      for (int u = threadIdx.x; u < U1; u += nth) {
        const double bb = (double)bls[u] + beta[u];
        nb[u] = bb;
        float acc = 0.f;
        for (int i = 0; i <= k; ++i) acc += post(chain_last_d(ad, lxs, u, i) + bb - z);
        p.grad_blank_w[off + u] = acc;
      }
      __syncthreads();
      for (int j = k - 1; j >= 0; --j) {
        for (int u = threadIdx.x; u < U1; u += nth) {
          const double bn = (u + 1 < U1) ? nb[u + 1] : dneg_inf();
          const double lb = (double)lxs[u] + bn;
          const float gv = post(chain_last_d(ad, lxs, u, j) + lb - z);
          if (j == k - 1) p.grad_lexical_w[off + u] = gv;
          else p.grad_lexical_w[off + u] += gv;
          out[u] = dlogaddexp((double)bls[u] + beta[u], lb);
        }
        __syncthreads();
        double* tmp = nb; nb = out; out = tmp;
      }
      for (int u = threadIdx.x; u < U1; u += nth) nxt[u] = nb[u];
      __syncthreads();
    }
    double* tmp = beta; beta = nxt; nxt = tmp;
  }
}

// ------------------------------------------------------------------- launch --
static int block_for(int U1) {
  int b = 32;
  while (b < U1 && b < 1024) b *= 2;
  return b;
}

int string_gather_launch(int V, int C, const float* blank, const float* lexical,
                         const int32_t* states, const int32_t* labels, int B, int T, int U1,
                         float* blank_w, float* lexical_w, cudaStream_t stream) {
  if ((size_t)B * T == 0 || U1 == 0) return LT_OK;
  string_gather_kernel<<<(unsigned)((size_t)B * T), min(block_for(U1), 256), 0, stream>>>(
      V, C, blank, lexical, states, labels, T, U1, blank_w, lexical_w);
  LT_LAUNCHED();
  return LT_OK;
}

int string_scatter_launch(int V, int C, const float* gbw, const float* glw,
                          const int32_t* states, const int32_t* labels, int B, int T, int U1,
                          float scale, const float* utt_scale, float* gblank, float* glex,
                          int split, cudaStream_t stream) {
  if ((size_t)B * T == 0 || U1 == 0) return LT_OK;
  if (split) {
    string_scatter_split_kernel<<<(unsigned)((size_t)B * T), min(block_for(U1), 256),
                                  (size_t)U1 * 8, stream>>>(
        V, C, gbw, glw, states, labels, T, U1, scale, utt_scale, gblank,
        reinterpret_cast<unsigned char*>(glex));
    LT_LAUNCHED();
    return LT_OK;
  }
  string_scatter_kernel<<<(unsigned)((size_t)B * T), min(block_for(U1), 256), 0, stream>>>(
      V, C, gbw, glw, states, labels, T, U1, scale, utt_scale, gblank, glex);
  LT_LAUNCHED();
  return LT_OK;
}

// ---------------------------------------------------------------------------
// FrameDependent fast variants (U+1 <= 1024): one thread per chain state keeps
// its alpha / beta in a REGISTER for all T frames; the per-frame weights are
// loaded kChunk frames ahead into registers (the chain is sequential in T, the
// weights are not), and the only per-frame communication is one neighbour
// value through a ping-pong shared array + one __syncthreads.
// ---------------------------------------------------------------------------
constexpr int kChunk = 8;

// EXT (Log only): every chain value is carried as (e, f) = integer part + fraction in [0, 1)
// (f = -inf, e = 0 for the semiring zero), i.e. alpha = e + f with e exact.  The numerator of a
// T = 1000 utterance reaches |alpha| ~ 1e3 where one fp32 ulp is 6e-5; with the integer part
// split off, every sum that is rounded (a + w, logaddexp of two neighbours, alpha + w + beta - z)
// has magnitude O(1), and the label-lattice posteriors keep ~1e-6 relative accuracy.
//   (e, f) <- logaddexp((ea, fa), (eb, fb)): both terms are re-based on the larger exponent of
//   the finite ones, so the smaller term only loses bits it could not contribute anyway.
// The chain is T sequential steps of one logaddexp each: its latency is the kernel's run time,
// so it uses the bare MUFU ops (ex2 / lg2, 2^-22 relative each).  With the integer parts split
// off both arguments are O(1), hence the absolute error per step is ~3e-7.
__device__ __forceinline__ float chain_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * kLog2e));
  return y;
}
__device__ __forceinline__ float chain_logaddexp(float a, float b) {
  const float c = fmaxf(a, b);
  const float cs = is_finite(c) ? c : 0.f;
  return cs + __log2f(chain_exp(a - cs) + chain_exp(b - cs)) * kLn2;
}
__device__ __forceinline__ void ext_logaddexp(int ea, float fa, int eb, float fb, int& e, float& f) {
  const bool oka = fa > neg_inf(), okb = fb > neg_inf();
  const int base = oka ? (okb ? max(ea, eb) : ea) : (okb ? eb : 0);
  const float xa = fa + (float)(ea - base), xb = fb + (float)(eb - base);
  const float r = chain_logaddexp(xa, xb);
  const float k = norm_shift(r);          // floor(r), 0 for -inf, clamped for absurd magnitudes
  e = base + (int)k;
  f = r - k;
}

// MAXT: block-size bound.  Label strings are short (U+1 <= 256 covers ASR utterances), and the
// kernels keep kChunk frames of weights in registers: bounding the block at 256 threads gives
// the compiler 255 registers per thread instead of 64 (the 1024-thread build spills).
template <int SR, bool EXT, int MAXT>
__global__ void __launch_bounds__(MAXT)
string_forward_fd_fast(const StrParams p) {
  using S = Sr<SR>;
  static_assert(!EXT || SR == LT_LOG, "the (e, f) representation is a Log-semiring feature");
  __shared__ float mv[2][MAXT + 1];
  __shared__ int me[EXT ? 2 : 1][EXT ? MAXT + 1 : 1];
  const int b = blockIdx.x, U1 = p.U1, u = threadIdx.x;
  const bool act = u < U1;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t base = (size_t)b * p.T * U1;
  float a = (u == 0) ? S::one() : S::zero();
  int ae = 0;                                          // EXT: alpha = ae + a
  float cb[kChunk], cl[kChunk], nb[kChunk], nl[kChunk];
  auto load = [&](int t0, float (&bb)[kChunk], float (&ll)[kChunk]) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const int t = t0 + i;
      const bool ok = act && t < nf;
      bb[i] = ok ? ldg_stream(p.blank_w + base + (size_t)t * U1 + u) : S::one();
      ll[i] = ok ? ldg_stream(p.lexical_w + base + (size_t)t * U1 + u) : S::zero();
    }
  };
  if (u == 0) { mv[0][0] = S::zero(); mv[1][0] = S::zero(); }   // "u - 1" of state 0
  if constexpr (EXT) { if (u == 0) { me[0][0] = 0; me[1][0] = 0; } }
  load(0, nb, nl);
  int par = 0;
  for (int t0 = 0; t0 < nf; t0 += kChunk) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i) { cb[i] = nb[i]; cl[i] = nl[i]; }
    load(t0 + kChunk, nb, nl);
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const int t = t0 + i;
      if (t < nf) {                                   // uniform over the block
        if (act) {
          if (p.alphas) p.alphas[base + (size_t)t * U1 + u] = a;
          mv[par][u + 1] = S::times(a, cl[i]);
          if constexpr (EXT) {
            p.alpha_exp[base + (size_t)t * U1 + u] = ae;
            me[par][u + 1] = ae;
          }
        }
        __syncthreads();
        if (act) {
          const float stay = S::times(a, cb[i]);
          const float move = mv[par][u];
          if constexpr (EXT) {
            ext_logaddexp(ae, stay, me[par][u], move, ae, a);
          } else if constexpr (SR == LT_MAXTROPICAL) {
            const bool tb = stay >= move;
            a = tb ? stay : move;
            if (p.backptr) p.backptr[base + (size_t)t * U1 + u] = tb ? 0 : 1;
          } else {
            a = S::plus(stay, move);
          }
        }
        par ^= 1;
      }
    }
  }
  if (act && p.alphas)
    for (int t = nf; t < p.T; ++t) {
      p.alphas[base + (size_t)t * U1 + u] = a;
      if constexpr (EXT) p.alpha_exp[base + (size_t)t * U1 + u] = ae;
    }
  const int nl_b = p.num_labels[b];
  if (act && u == nl_b) {                                       // lattices.py:375-377
    if constexpr (EXT) {
      p.dist[b] = (float)((double)ae + (double)a);
      p.dist_norm[2 * b] = ae;
      p.dist_norm[2 * b + 1] = __float_as_int(a);
    } else {
      p.dist[b] = a;
    }
  }
  if (u == 0 && (nl_b < 0 || nl_b >= U1)) {
    p.dist[b] = S::zero();
    if constexpr (EXT) { p.dist_norm[2 * b] = 0; p.dist_norm[2 * b + 1] = __float_as_int(S::zero()); }
  }
}

template <int SR, bool EXT, int MAXT>   // LT_LOG or LT_REAL
__global__ void __launch_bounds__(MAXT)
string_backward_fd_fast(const StrParams p) {
  using S = Sr<SR>;
  static_assert(!EXT || SR == LT_LOG, "the (e, f) representation is a Log-semiring feature");
  __shared__ float sh[2][MAXT + 1];
  __shared__ int she[EXT ? 2 : 1][EXT ? MAXT + 1 : 1];
  const int b = blockIdx.x, U1 = p.U1, u = threadIdx.x;
  const bool act = u < U1;
  const int nf = max(0, min(p.num_frames[b], p.T));
  const int nl_b = p.num_labels[b];
  const float z = EXT ? __int_as_float(p.dist_norm_in[2 * b + 1]) : p.dist_in[b];
  const int ze = EXT ? p.dist_norm_in[2 * b] : 0;       // EXT: numerator = ze + z
  const float g = p.grad_dist ? p.grad_dist[b] : 1.f;
  const bool reachable = (nl_b >= 0 && nl_b < U1) && (SR == LT_REAL ? true : is_finite(z));
  const size_t base = (size_t)b * p.T * U1;
  if (act) {
    for (int t = reachable ? nf : 0; t < p.T; ++t) {
      p.grad_blank_w[base + (size_t)t * U1 + u] = 0.f;
      p.grad_lexical_w[base + (size_t)t * U1 + u] = 0.f;
    }
  }
  if (!reachable) return;
  float beta = (u == nl_b) ? S::one() : S::zero();
  int be = 0;                                           // EXT: beta = be + beta
  float cb[kChunk], cl[kChunk], ca[kChunk], nb[kChunk], nl[kChunk], na[kChunk];
  int cae[EXT ? kChunk : 1], nae[EXT ? kChunk : 1];
  auto load = [&](int i0, float (&bb)[kChunk], float (&ll)[kChunk], float (&aa)[kChunk],
                  int (&ee)[EXT ? kChunk : 1]) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const int t = nf - 1 - (i0 + i);
      const bool ok = act && t >= 0;
      const size_t o = base + (size_t)(ok ? t : 0) * U1 + (act ? u : 0);
      bb[i] = ok ? ldg_stream(p.blank_w + o) : S::one();
      ll[i] = ok ? ldg_stream(p.lexical_w + o) : S::zero();
      aa[i] = ok ? p.alphas_in[o] : S::zero();
      if constexpr (EXT) ee[i] = ok ? p.alpha_exp_in[o] : 0;
    }
  };
  if (u == 0) { sh[0][U1] = S::zero(); sh[1][U1] = S::zero(); }   // "u + 1" of the last state
  if constexpr (EXT) { if (u == 0) { she[0][U1] = 0; she[1][U1] = 0; } }
  load(0, nb, nl, na, nae);
  int par = 0;
  for (int i0 = 0; i0 < nf; i0 += kChunk) {
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      cb[i] = nb[i]; cl[i] = nl[i]; ca[i] = na[i];
      if constexpr (EXT) cae[i] = nae[i];
    }
    load(i0 + kChunk, nb, nl, na, nae);
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
      const int t = nf - 1 - (i0 + i);
      if (t >= 0) {                                   // uniform over the block
        if (act) sh[par][u] = beta;
        if constexpr (EXT) { if (act) she[par][u] = be; }
        __syncthreads();
        if (act) {
          const float bn = sh[par][u + 1];
          const float bb = S::times(cb[i], beta);
          const float lb = S::times(cl[i], bn);
          const size_t o = base + (size_t)t * U1 + u;
          if constexpr (EXT) {
            // exponent = (alpha_f + w + beta_f - z_f) + (alpha_e + beta_e - z_e): the integer
            // parts cancel exactly, the fractions are O(1)
            const int bne = she[par][u + 1];
            const int eb = cae[i] + be - ze, el = cae[i] + bne - ze;
            p.grad_blank_w[o] = g * chain_exp((ca[i] + bb - z) + (float)eb);
            p.grad_lexical_w[o] = g * chain_exp((ca[i] + lb - z) + (float)el);
            ext_logaddexp(be, bb, bne, lb, be, beta);
          } else if constexpr (SR == LT_LOG) {
            p.grad_blank_w[o] = g * fast_exp(ca[i] + bb - z);
            p.grad_lexical_w[o] = g * fast_exp(ca[i] + lb - z);
            beta = S::plus(bb, lb);
          } else {
            p.grad_blank_w[o] = g * ca[i] * beta;
            p.grad_lexical_w[o] = g * ca[i] * bn;
            beta = S::plus(bb, lb);
          }
        }
        par ^= 1;
      }
    }
  }
}

template <int SR>
static int string_fwd_sr(const StrParams& p, cudaStream_t stream) {
  if (p.k < 1 && p.U1 <= 1024) {
    const bool small = p.U1 <= 256;
    if constexpr (SR == LT_LOG) {
      if (p.alpha_exp) {
        if (small) string_forward_fd_fast<SR, true, 256><<<p.B, block_for(p.U1), 0, stream>>>(p);
        else string_forward_fd_fast<SR, true, 1024><<<p.B, block_for(p.U1), 0, stream>>>(p);
        LT_LAUNCHED();
        return LT_OK;
      }
    }
    if (small) string_forward_fd_fast<SR, false, 256><<<p.B, block_for(p.U1), 0, stream>>>(p);
    else string_forward_fd_fast<SR, false, 1024><<<p.B, block_for(p.U1), 0, stream>>>(p);
    LT_LAUNCHED();
    return LT_OK;
  }
  const int block = block_for(p.U1);
  if constexpr (SR == LT_LOG) {
    if (p.alpha_exp) {                      // double chain, (e, f) storage
      const size_t dsmem = (sizeof(double) * 2 + sizeof(float) * 2) * p.U1;
      if (p.k >= 1) {
        LT_CUDA(cudaFuncSetAttribute(string_forward_ext_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
        string_forward_ext_kernel<true><<<p.B, block, dsmem, stream>>>(p);
      } else {
        LT_CUDA(cudaFuncSetAttribute(string_forward_ext_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
        string_forward_ext_kernel<false><<<p.B, block, dsmem, stream>>>(p);
      }
      LT_LAUNCHED();
      return LT_OK;
    }
  }
  const size_t smem = sizeof(float) * 2 * p.U1;
  if (p.k >= 1) {
    LT_CUDA(cudaFuncSetAttribute(string_forward_kernel<SR, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    string_forward_kernel<SR, true><<<p.B, block, smem, stream>>>(p);
  } else {
    LT_CUDA(cudaFuncSetAttribute(string_forward_kernel<SR, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    string_forward_kernel<SR, false><<<p.B, block, smem, stream>>>(p);
  }
  LT_LAUNCHED();
  return LT_OK;
}

template <int SR>
static int string_bwd_sr(const StrParams& p, cudaStream_t stream) {
  if constexpr (SR != LT_MAXTROPICAL) {
    if (p.k < 1 && p.U1 <= 1024) {
      const bool small = p.U1 <= 256;
      if constexpr (SR == LT_LOG) {
        if (p.alpha_exp_in) {
          if (small) string_backward_fd_fast<SR, true, 256><<<p.B, block_for(p.U1), 0, stream>>>(p);
          else string_backward_fd_fast<SR, true, 1024><<<p.B, block_for(p.U1), 0, stream>>>(p);
          LT_LAUNCHED();
          return LT_OK;
        }
      }
      if (small) string_backward_fd_fast<SR, false, 256><<<p.B, block_for(p.U1), 0, stream>>>(p);
      else string_backward_fd_fast<SR, false, 1024><<<p.B, block_for(p.U1), 0, stream>>>(p);
      LT_LAUNCHED();
      return LT_OK;
    }
  }
  const int block = block_for(p.U1);
  if constexpr (SR == LT_LOG) {
    if (p.alpha_exp_in) {
      const size_t dsmem = (sizeof(double) * 5 + sizeof(float) * 2) * p.U1;
      if (p.k >= 1) {
        LT_CUDA(cudaFuncSetAttribute(string_backward_ext_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
        string_backward_ext_kernel<true><<<p.B, block, dsmem, stream>>>(p);
      } else {
        LT_CUDA(cudaFuncSetAttribute(string_backward_ext_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
        string_backward_ext_kernel<false><<<p.B, block, dsmem, stream>>>(p);
      }
      LT_LAUNCHED();
      return LT_OK;
    }
  }
  const size_t smem = sizeof(float) * 4 * p.U1;
  if (p.k >= 1) {
    LT_CUDA(cudaFuncSetAttribute(string_backward_kernel<SR, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    string_backward_kernel<SR, true><<<p.B, block, smem, stream>>>(p);
  } else {
    LT_CUDA(cudaFuncSetAttribute(string_backward_kernel<SR, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    string_backward_kernel<SR, false><<<p.B, block, smem, stream>>>(p);
  }
  LT_LAUNCHED();
  return LT_OK;
}

int string_forward_launch(int semiring, const StrParams& p, cudaStream_t stream) {
  if (p.B == 0) return LT_OK;
  if ((size_t)p.U1 * 4 * sizeof(float) > 200 * 1024) {
    set_error("lt_string_forward: U+1 = %d label states do not fit in shared memory", p.U1);
    return LT_ERR_UNSUPPORTED;
  }
  switch (semiring) {
    case LT_REAL: return string_fwd_sr<LT_REAL>(p, stream);
    case LT_LOG: return string_fwd_sr<LT_LOG>(p, stream);
    case LT_MAXTROPICAL: return string_fwd_sr<LT_MAXTROPICAL>(p, stream);
  }
  set_error("lt_string_forward: unknown semiring %d", semiring);
  return LT_ERR_INVALID_ARGUMENT;
}

int string_backward_launch(int semiring, const StrParams& p, cudaStream_t stream) {
  if (p.B == 0) return LT_OK;
  if ((size_t)p.U1 * 4 * sizeof(float) > 200 * 1024) {
    set_error("lt_string_backward: U+1 = %d label states do not fit in shared memory", p.U1);
    return LT_ERR_UNSUPPORTED;
  }
  switch (semiring) {
    case LT_REAL: return string_bwd_sr<LT_REAL>(p, stream);
    case LT_LOG: return string_bwd_sr<LT_LOG>(p, stream);
    case LT_MAXTROPICAL: return string_bwd_sr<LT_MAXTROPICAL>(p, stream);
  }
  set_error("lt_string_backward: unknown semiring %d", semiring);
  return LT_ERR_INVALID_ARGUMENT;
}

}  // namespace lt
