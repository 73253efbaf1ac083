/*
 * CPU oracle in plain C (TEST INFRASTRUCTURE ONLY -- never linked into or
 * called from the product package last_torch_b200/).
 *
 * A restatement of the reference's Log-semiring GNAT loss and its gradient
 * (theadamsabra/last_torch, /root/reference at survey time):
 *   forward  recursion : lattices.py:436-462 with alignments.py:294-297
 *                        (FrameDependent) / :370-376 (FrameLabelDependent) and
 *                        FullNGram.forward_reduce contexts.py:207-230
 *   backward recursion : alignments.py:300-318 / :378-418 with
 *                        backward_broadcast contexts.py:232-256 and the padding
 *                        masks of lattices.py:775-779
 *   numerator          : lattices.py:250-377 (walk_states contexts.py:109-146,
 *                        string_forward alignments.py:327-329 / :427-432)
 *   loss               : lattices.py:183  (denominator - numerator)
 * Log (+) is the max-shifted form of semirings.py:247-255 / :279-286.
 *
 * It is pinned by tests/test_oracle_golden.py against oracle/lattice_oracle.py,
 * which in turn is pinned against the reference's golden vectors.  It exists so
 * that bench.py can time a multi-threaded CPU implementation of the same path
 * (OpenMP over utterances; utterances are independent) on the GPU box, where
 * the Python reference is not available.
 *
 * All arithmetic in double or float according to LT_ORACLE_REAL.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef LT_ORACLE_REAL
#define LT_ORACLE_REAL float
#endif
typedef LT_ORACLE_REAL real;
#define EXP(x) (sizeof(real) == 4 ? (double)expf((float)(x)) : exp((double)(x)))
#define LOG(x) (sizeof(real) == 4 ? (double)logf((float)(x)) : log((double)(x)))

typedef struct {
  int V, n, C, A, Alow, N, K, off;
} ngram_t;

static void make_ngram(int V, int n, ngram_t* g) {
  long long C = 0, pw = 1, A = 0, Alow = 0;
  for (int i = 0; i <= n; ++i) {
    if (i < n) A += pw;
    if (i < n - 1) Alow += pw;
    C += pw;
    if (i < n) pw *= V;
  }
  g->V = V; g->n = n; g->C = (int)C; g->A = (int)A; g->Alow = (int)Alow; g->N = (int)pw;
  g->K = (int)((C - Alow) * V / pw);
  g->off = n > 0 ? 1 : 0;
}

/* contexts.py:190-205 with zero-based label */
static int next_state(const ngram_t* g, int p, int y0) {
  if (p < g->Alow) return g->off + p * g->V + y0;
  return g->A + (int)((((long long)(p - g->Alow)) * g->V + y0) % g->N);
}

static real msafe(real m) { return isfinite(m) ? m : (real)0; }

static real logaddexp_r(real a, real b) {   /* semirings.py:247-255 */
  real c = a > b ? a : b;
  real cs = msafe(c);
  return cs + (real)LOG(EXP(a - cs) + EXP(b - cs));
}

/* out[q] = logsumexp over arcs into q of src[p] + lex[p,y]  (contexts.py:207-230) */
static void forward_reduce_log(const ngram_t* g, const real* src, const real* lex, real* out,
                               real* colmax) {
  const int V = g->V, lowV = g->Alow * g->V, N = g->N;
  if (g->n > 0) out[0] = -INFINITY;
  for (int f = 0; f < lowV; ++f) out[g->off + f] = src[f / V] + lex[f];
  real* o = out + g->off + lowV;
  const real* tail = lex + lowV;
  for (int j = 0; j < N; ++j) colmax[j] = -INFINITY;
  for (int kk = 0; kk < g->K; ++kk) {
    const real* row = tail + (size_t)kk * N;
    for (int j = 0; j < N; ++j) {
      const real x = src[(lowV + (size_t)kk * N + j) / V] + row[j];
      if (x > colmax[j]) colmax[j] = x;
    }
  }
  for (int j = 0; j < N; ++j) { colmax[j] = msafe(colmax[j]); o[j] = 0; }
  for (int kk = 0; kk < g->K; ++kk) {
    const real* row = tail + (size_t)kk * N;
    for (int j = 0; j < N; ++j) {
      const real x = src[(lowV + (size_t)kk * N + j) / V] + row[j];
      o[j] += (real)EXP(x - colmax[j]);
    }
  }
  for (int j = 0; j < N; ++j) o[j] = colmax[j] + (real)LOG(o[j]);
}

/*
 * Denominator: logZ, alphas and arc marginals (scaled by gscale[b]) for every
 * utterance.  k < 0 selects FrameDependent.  grad_* may be NULL (forward only).
 */
void oracle_lattice_log(int V, int n, int k, const real* blank, const real* lexical,
                        const int32_t* num_frames, int B, int T, const real* gscale,
                        real* log_z, real* alphas, real* grad_blank, real* grad_lexical) {
  ngram_t g;
  make_ngram(V, n, &g);
  const int C = g.C;
  const int nlev = k < 0 ? 0 : k;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    int nf = num_frames[b];
    if (nf < 0) nf = 0;
    if (nf > T) nf = T;
    real* alpha = (real*)malloc(sizeof(real) * C);
    real* nxt = (real*)malloc(sizeof(real) * C);
    real* red = (real*)malloc(sizeof(real) * C);
    real* colmax = (real*)malloc(sizeof(real) * (g.N > C ? g.N : C));
    real* levels = (real*)malloc(sizeof(real) * (size_t)(nlev + 1) * C * (T > 0 ? T : 1));
    real* al = alphas + (size_t)b * T * C;
    for (int c = 0; c < C; ++c) alpha[c] = c == 0 ? 0 : -INFINITY;   /* lattices.py:801-807 */
    for (int t = 0; t < T; ++t) {
      memcpy(al + (size_t)t * C, alpha, sizeof(real) * C);
      if (t >= nf) continue;                                        /* lattices.py:460-461 */
      const real* bl = blank + ((size_t)b * T + t) * C;
      const real* lx = lexical + ((size_t)b * T + t) * (size_t)C * V;
      if (k < 0) {                                                  /* alignments.py:294-297 */
        forward_reduce_log(&g, alpha, lx, red, colmax);
        for (int c = 0; c < C; ++c) nxt[c] = logaddexp_r(alpha[c] + bl[c], red[c]);
      } else {                                                      /* alignments.py:370-376 */
        real* lev = levels + (size_t)t * (nlev + 1) * C;
        memcpy(lev, alpha, sizeof(real) * C);
        for (int i = 0; i < k; ++i)
          forward_reduce_log(&g, lev + (size_t)i * C, lx, lev + (size_t)(i + 1) * C, colmax);
        for (int c = 0; c < C; ++c) {
          real m = -INFINITY;
          for (int i = 0; i <= k; ++i) { real x = lev[(size_t)i * C + c] + bl[c]; if (x > m) m = x; }
          const real ms = msafe(m);
          double s = 0;
          for (int i = 0; i <= k; ++i) s += EXP(lev[(size_t)i * C + c] + bl[c] - ms);
          nxt[c] = ms + (real)LOG(s);
        }
      }
      real* tmp = alpha; alpha = nxt; nxt = tmp;
    }
    {                                                               /* lattices.py:496 */
      real m = -INFINITY;
      for (int c = 0; c < C; ++c) if (alpha[c] > m) m = alpha[c];
      const real ms = msafe(m);
      double s = 0;
      for (int c = 0; c < C; ++c) s += EXP(alpha[c] - ms);
      log_z[b] = ms + (real)LOG(s);
    }
    if (grad_blank && grad_lexical) {
      const real z = log_z[b];
      const real gs = gscale ? gscale[b] : (real)1;
      real* beta = alpha;                 /* reuse buffers */
      real* nb = nxt;
      real* nb2 = red;
      for (int c = 0; c < C; ++c) beta[c] = 0;                       /* lattices.py:789-790 */
      for (int t = T - 1; t >= 0; --t) {
        real* gb = grad_blank + ((size_t)b * T + t) * C;
        real* gl = grad_lexical + ((size_t)b * T + t) * (size_t)C * V;
        if (t >= nf) {                                              /* lattices.py:775-779 */
          memset(gb, 0, sizeof(real) * C);
          memset(gl, 0, sizeof(real) * (size_t)C * V);
          continue;
        }
        const real* bl = blank + ((size_t)b * T + t) * C;
        const real* lx = lexical + ((size_t)b * T + t) * (size_t)C * V;
        const real* a = al + (size_t)t * C;
        if (k < 0) {                                                /* alignments.py:311-318 */
          for (int p = 0; p < C; ++p) {
            const real* row = lx + (size_t)p * V;
            real m = -INFINITY;
            for (int y = 0; y < V; ++y) {
              const real x = row[y] + beta[next_state(&g, p, y)];
              if (x > m) m = x;
            }
            const real ms = msafe(m);
            const real scale = a[p] - z;
            double s = 0;
            for (int y = 0; y < V; ++y) {
              const real x = row[y] + beta[next_state(&g, p, y)];
              s += EXP(x - ms);
              gl[(size_t)p * V + y] = gs * (real)EXP(x + scale);
            }
            const real bb = bl[p] + beta[p];
            gb[p] = gs * (real)EXP(bb + scale);
            nb[p] = logaddexp_r(bb, ms + (real)LOG(s));
          }
          real* tmp = beta; beta = nb; nb = tmp;
        } else {                                                    /* alignments.py:390-418 */
          const real* lev = levels + (size_t)t * (nlev + 1) * C;
          for (int p = 0; p < C; ++p) {
            double acc = 0;
            for (int i = 0; i <= k; ++i) acc += EXP(lev[(size_t)i * C + p] + bl[p] + beta[p] - z);
            gb[p] = gs * (real)acc;
            nb[p] = bl[p] + beta[p];
          }
          memset(gl, 0, sizeof(real) * (size_t)C * V);
          for (int j = k - 1; j >= 0; --j) {
            const real* la = lev + (size_t)j * C;
            for (int p = 0; p < C; ++p) {
              const real* row = lx + (size_t)p * V;
              real m = -INFINITY;
              for (int y = 0; y < V; ++y) {
                const real x = row[y] + nb[next_state(&g, p, y)];
                if (x > m) m = x;
              }
              const real ms = msafe(m);
              double s = 0;
              for (int y = 0; y < V; ++y) {
                const real x = row[y] + nb[next_state(&g, p, y)];
                s += EXP(x - ms);
                gl[(size_t)p * V + y] += gs * (real)EXP(x + la[p] - z);
              }
              nb2[p] = logaddexp_r(bl[p] + beta[p], ms + (real)LOG(s));
            }
            real* tmp = nb; nb = nb2; nb2 = tmp;
          }
          real* tmp = beta; beta = nb; nb = tmp;
        }
      }
      /* restore ownership for free() below */
      alpha = beta; nxt = nb; red = nb2;
    }
    free(alpha); free(nxt); free(red); free(colmax); free(levels);
  }
}

/*
 * Numerator (FrameDependent or FrameLabelDependent) on the label chain and its
 * gradient scattered into the dense gradient buffers with factor `sign`:
 * grad[b,t,state[u],label[u]-1] += sign * gscale[b] * posterior.
 */
void oracle_string_log(int V, int n, int k, const real* blank, const real* lexical,
                       const int32_t* num_frames, const int32_t* labels,
                       const int32_t* num_labels, int B, int T, int U, const real* gscale,
                       real sign, real* numerator, real* grad_blank, real* grad_lexical) {
  ngram_t g;
  make_ngram(V, n, &g);
  const int C = g.C, U1 = U + 1;
  const int kk = k < 0 ? 0 : k;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    int nf = num_frames[b];
    if (nf < 0) nf = 0;
    if (nf > T) nf = T;
    int* st = (int*)malloc(sizeof(int) * U1);
    int* lab = (int*)malloc(sizeof(int) * U1);
    st[0] = 0;                                                       /* contexts.py:109-146 */
    for (int u = 0; u < U; ++u) {
      const int y = labels[(size_t)b * U + u];
      st[u + 1] = y == 0 ? st[u] : next_state(&g, st[u], y - 1);
      lab[u] = y < 1 ? 1 : y;                                        /* lattices.py:314-315 */
    }
    lab[U] = 1;                                                      /* lattices.py:337-338 */
    real* bw = (real*)malloc(sizeof(real) * (size_t)(T > 0 ? T : 1) * U1);
    real* lw = (real*)malloc(sizeof(real) * (size_t)(T > 0 ? T : 1) * U1);
    real* al = (real*)malloc(sizeof(real) * (size_t)(T + 1) * U1);
    real* lasts = (real*)malloc(sizeof(real) * (size_t)(kk + 1) * U1);
    real* beta = (real*)malloc(sizeof(real) * U1);
    real* nb = (real*)malloc(sizeof(real) * U1);
    real* nb2 = (real*)malloc(sizeof(real) * U1);
    for (int t = 0; t < T; ++t) {                                    /* lattices.py:317-333 */
      const real* bl = blank + ((size_t)b * T + t) * C;
      const real* lx = lexical + ((size_t)b * T + t) * (size_t)C * V;
      for (int u = 0; u < U1; ++u) {
        bw[(size_t)t * U1 + u] = bl[st[u]];
        lw[(size_t)t * U1 + u] = lx[(size_t)st[u] * V + lab[u] - 1];
      }
    }
    for (int u = 0; u < U1; ++u) al[u] = u == 0 ? 0 : -INFINITY;
    for (int t = 0; t < nf; ++t) {
      const real* a = al + (size_t)t * U1;
      real* o = al + (size_t)(t + 1) * U1;
      const real* bl = bw + (size_t)t * U1;
      const real* lx = lw + (size_t)t * U1;
      if (k < 0) {                                                   /* alignments.py:327-329 */
        for (int u = 0; u < U1; ++u)
          o[u] = logaddexp_r(a[u] + bl[u], u > 0 ? a[u - 1] + lx[u - 1] : -INFINITY);
      } else {                                                       /* alignments.py:427-432 */
        memcpy(lasts, a, sizeof(real) * U1);
        for (int i = 0; i < k; ++i) {
          real* li = lasts + (size_t)i * U1; real* lo = lasts + (size_t)(i + 1) * U1;
          lo[0] = -INFINITY;
          for (int u = 1; u < U1; ++u) lo[u] = li[u - 1] + lx[u - 1];
        }
        for (int u = 0; u < U1; ++u) {
          real m = -INFINITY;
          for (int i = 0; i <= k; ++i) { real x = lasts[(size_t)i * U1 + u] + bl[u]; if (x > m) m = x; }
          const real ms = msafe(m);
          double s = 0;
          for (int i = 0; i <= k; ++i) s += EXP(lasts[(size_t)i * U1 + u] + bl[u] - ms);
          o[u] = ms + (real)LOG(s);
        }
      }
    }
    const int nl = num_labels[b];
    const real z = (nl >= 0 && nl < U1) ? al[(size_t)nf * U1 + nl] : -INFINITY;  /* lattices.py:375-377 */
    numerator[b] = z;
    if (grad_blank && grad_lexical && isfinite(z)) {
      const real gs = sign * (gscale ? gscale[b] : (real)1);
      for (int u = 0; u < U1; ++u) beta[u] = u == nl ? 0 : -INFINITY;
      for (int t = nf - 1; t >= 0; --t) {
        const real* a = al + (size_t)t * U1;
        const real* bl = bw + (size_t)t * U1;
        const real* lx = lw + (size_t)t * U1;
        real* gb = grad_blank + ((size_t)b * T + t) * C;
        real* gl = grad_lexical + ((size_t)b * T + t) * (size_t)C * V;
        if (k < 0) {
          for (int u = 0; u < U1; ++u) {
            const real lb = lx[u] + (u + 1 < U1 ? beta[u + 1] : -INFINITY);
            gb[st[u]] += gs * (real)EXP(a[u] + bl[u] + beta[u] - z);
            gl[(size_t)st[u] * V + lab[u] - 1] += gs * (real)EXP(a[u] + lb - z);
            nb[u] = logaddexp_r(bl[u] + beta[u], lb);
          }
          real* tmp = beta; beta = nb; nb = tmp;
        } else {
          memcpy(lasts, a, sizeof(real) * U1);
          for (int i = 0; i < k; ++i) {
            real* li = lasts + (size_t)i * U1; real* lo = lasts + (size_t)(i + 1) * U1;
            lo[0] = -INFINITY;
            for (int u = 1; u < U1; ++u) lo[u] = li[u - 1] + lx[u - 1];
          }
          for (int u = 0; u < U1; ++u) {
            double acc = 0;
            for (int i = 0; i <= k; ++i) acc += EXP(lasts[(size_t)i * U1 + u] + bl[u] + beta[u] - z);
            gb[st[u]] += gs * (real)acc;
            nb[u] = bl[u] + beta[u];
          }
          for (int j = k - 1; j >= 0; --j) {
            for (int u = 0; u < U1; ++u) {
              const real lb = lx[u] + (u + 1 < U1 ? nb[u + 1] : -INFINITY);
              gl[(size_t)st[u] * V + lab[u] - 1] += gs * (real)EXP(lasts[(size_t)j * U1 + u] + lb - z);
              nb2[u] = logaddexp_r(bl[u] + beta[u], lb);
            }
            real* tmp = nb; nb = nb2; nb2 = tmp;
          }
          real* tmp = beta; beta = nb; nb = tmp;
        }
      }
    }
    free(st); free(lab); free(bw); free(lw); free(al); free(lasts); free(beta); free(nb); free(nb2);
  }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int oracle_real_size(void) { return (int)sizeof(real); }

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline asks for all cores explicitly. */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
