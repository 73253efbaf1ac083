// Shared device/host helpers for the lattice kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/last_lattice.h"

// mbarrier.try_wait suspend-time hint (ns): the waiting warp is parked by the hardware and woken
// when the phase completes instead of re-polling the barrier.  Every poll is a shared-memory
// access on the L1 data pipe -- the pipe the tensor core's operand reads and the producers'
// stores already saturate in the joint kernels (ncu, profiles/r02: 236 M polls = a quarter of all
// issued instructions of the forward kernel before the hint).  Used by the joint (tensor-core)
// kernels only: measured -1.7 % on the forward kernel; the lattice kernels wait on their state
// exchange once per frame, where the wake-up latency of a parked warp costs more (+2.4 %).
#ifndef LT_MBAR_HINT
#define LT_MBAR_HINT ", 0x989680"
#endif

namespace lt {

// ---------------------------------------------------------------------------
// Error reporting (thread-local: autograd calls backward from its own thread).
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void note_launch();   // bumps the process-wide kernel-launch counter (lt_launch_count)

// Process-wide debug / test switches (lt_set_option / lt_get_option).  Each one is initialised
// ONCE, on first use, from the environment variable of the same name; the kernels' dispatch
// code only ever does an atomic load.
enum Option {
  OPT_JOINT_SIMT,             // LT_JOINT_SIMT: CUDA-core joint kernels instead of tcgen05
  OPT_JOINT_DGRAD_V1,         // LT_JOINT_DGRAD_V1: first-generation dgrad + streaming reduction
  OPT_JOINT_WGRAD_SIMT,       // LT_JOINT_WGRAD_SIMT: CUDA-core weight gradient
  OPT_JOINT_DGRAD_PAIR,       // LT_JOINT_DGRAD_PAIR: cta_group::2 split-row dgrad
  OPT_JOINT_DGRAD_MULTICAST,  // LT_JOINT_DGRAD_MULTICAST: TMA-multicast split-row dgrad
  OPT_TABLE_V1,               // LT_TABLE_V1: one-CTA NextStateTable kernels
  OPT_TABLE_CLUSTER,          // LT_TABLE_CLUSTER: force the NextStateTable cluster size
  OPT_JOINT_FWD_SS,           // LT_JOINT_FWD_SS: forward with the tanh operand in shared memory
  OPT_JOINT_FWD_CLUSTER,      // LT_JOINT_FWD_CLUSTER=1: no W_vocab multicast (one CTA per cluster)
  OPT_FLD_GENERIC,            // LT_FLD_GENERIC: FrameLabelDependent on bigram contexts uses the generic kernels
  OPT_LINEAR_SIMT,            // LT_LINEAR_SIMT: lt_linear_* on CUDA cores even where the tcgen05 kernels apply
  OPT_COUNT
};
int option(Option o);

#define LT_CHECK_ARG(cond, ...)                    \
  do {                                             \
    if (!(cond)) {                                 \
      ::lt::set_error(__VA_ARGS__);                \
      return LT_ERR_INVALID_ARGUMENT;              \
    }                                              \
  } while (0)

#define LT_CUDA(call)                                        \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return ::lt::cuda_fail(e__, #call); \
  } while (0)

// ---------------------------------------------------------------------------
// FullNGram geometry (contexts.py:181-230), computed once on the host.
//   states 0 .. A-1      : "ascending" n-grams of order < n
//   states A .. C-1      : the N = V^n full-order n-grams
//   rows p < Alow        : their V arcs lead to ascending states; arc (p,y)
//                          (flat index p*V+y) is the ONLY arc into state off+flat
//   rows p >= Alow       : flat tail index j = (p-Alow)*V + y leads to state
//                          A + (j mod N); every full-order state has K arcs,
//                          at tail offsets (q-A) + kk*N, kk = 0..K-1
// ---------------------------------------------------------------------------
struct NGram {
  int V, n, C;
  int A;        // sum_{i<n} V^i
  int Alow;     // sum_{i<n-1} V^i
  int N;        // V^n
  int K;        // (C - Alow) * V / N  (= V+1 for n>=1, V for n==0)
  int off;      // 1 if n > 0 else 0
  int pstride;  // source-state stride between consecutive kk: N / V (0 if n==0)
};

inline bool make_ngram(int V, int n, NGram* g) {
  if (V <= 0 || n < 0) return false;
  long long C = 0, pw = 1, A = 0, Alow = 0;
  for (int i = 0; i <= n; ++i) {
    if (i < n) A += pw;
    if (i < n - 1) Alow += pw;
    C += pw;
    if (i < n) pw *= V;
    if (C > (1ll << 28)) return false;
  }
  g->V = V; g->n = n; g->C = (int)C; g->A = (int)A; g->Alow = (int)Alow;
  g->N = (int)pw;
  g->K = (int)((C - Alow) * V / pw);
  g->off = n > 0 ? 1 : 0;
  g->pstride = n > 0 ? (int)(pw / V) : 0;
  return true;
}

// dest state of arc (p, y0) with y0 zero-based: contexts.py:190-205.
__host__ __device__ inline int ngram_next(const NGram& g, int p, int y0) {
  if (p < g.Alow) return g.off + p * g.V + y0;
  return g.A + ((p - g.Alow) * g.V + y0) % g.N;
}

// ---------------------------------------------------------------------------
// Semiring scalar algebra.
// ---------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }
__device__ __forceinline__ float pos_inf() { return __int_as_float(0x7f800000); }
__device__ __forceinline__ bool is_finite(float x) { return fabsf(x) <= 3.402823466e38f; }

// exp / log of the latency-bound kernels (generic lattice kernels, numerator chain, semiring
// ops, one-CTA table kernels): the accurate library versions (<= 2 ulp).  exp2f(x * log2e)
// rounds the product at ulp(x * log2e) -- 1e-6 relative at |x| ~ 20 -- and lg2.approx * ln2
// is as loose; these kernels are not instruction-bound, so they can afford the reference's
// accuracy.  The HBM-bound fast paths work in log2 units with bare MUFU ops instead
// (fast_ptx.cuh) and bound their error by renormalising the recursion state.
__device__ __forceinline__ float fast_exp(float x) { return expf(x); }
__device__ __forceinline__ float fast_log(float x) { return logf(x); }

template <int SR> struct Sr;

template <> struct Sr<LT_REAL> {
  __device__ static float zero() { return 0.f; }
  __device__ static float one() { return 1.f; }
  __device__ static float times(float a, float b) { return a * b; }
  __device__ static float plus(float a, float b) { return a + b; }
};

// semirings.py:247-255: c = max(a,b); non-finite c is replaced by 0.
__device__ __forceinline__ float log_add_exp(float a, float b) {
  float c = fmaxf(a, b);
  float cs = is_finite(c) ? c : 0.f;
  float z = fast_exp(a - cs) + fast_exp(b - cs);
  return cs + fast_log(z);
}

template <> struct Sr<LT_LOG> {
  __device__ static float zero() { return neg_inf(); }
  __device__ static float one() { return 0.f; }
  __device__ static float times(float a, float b) { return a + b; }
  __device__ static float plus(float a, float b) { return log_add_exp(a, b); }
};

template <> struct Sr<LT_MAXTROPICAL> {
  __device__ static float zero() { return neg_inf(); }
  __device__ static float one() { return 0.f; }
  __device__ static float times(float a, float b) { return a + b; }
  __device__ static float plus(float a, float b) { return fmaxf(a, b); }
};

// Running (+)-accumulator.  Log keeps (m, s) with value = msafe(m) + log(s),
// msafe(m) = m if finite else 0 (semirings.py:281-285); MaxTropical keeps the
// max and the FIRST arg-max (semirings.py:382); Real keeps the sum.
template <int SR> struct Acc;

template <> struct Acc<LT_REAL> {
  float s;
  __device__ void init() { s = 0.f; }
  __device__ void add(float x, int) { s += x; }
  __device__ void merge(const Acc& o) { s += o.s; }
  __device__ float value() const { return s; }
  __device__ int arg() const { return 0; }
};

__device__ __forceinline__ float msafe(float m) { return is_finite(m) ? m : 0.f; }
// Value of a Log accumulator (m, s) minus `shift`, rounded ONCE: (msafe(m) - shift) + log(s) in
// double.  -inf when the accumulator is empty (s == 0).
__device__ __forceinline__ float log_value_shifted(float m, float s, float shift) {
  return (float)(((double)msafe(m) - (double)shift) + (double)logf(s));
}
// Per-frame shift of the renormalised recursions: floor(max), an integer so that the running
// offset is exact; 0 when the maximum is not finite, and clamped so that offsets of absurd
// magnitude (weights like -1e30 standing in for -inf) stay inside int32.
__device__ __forceinline__ float norm_shift(float m) {
  if (!is_finite(m)) return 0.f;
  return fminf(fmaxf(floorf(m), -16777216.f), 16777216.f);
}

template <> struct Acc<LT_LOG> {
  float m, s;
  __device__ void init() { m = neg_inf(); s = 0.f; }
  // add one term (2 ex2 worst case; chunked callers use add_chunk instead)
  __device__ void add(float x, int) {
    float mn = fmaxf(m, x);
    float ms = msafe(mn);
    float sc = (m == neg_inf()) ? 0.f : fast_exp(msafe(m) - ms);
    s = s * sc + fast_exp(x - ms);
    m = mn;
  }
  // add a chunk with precomputed max `cm`: one rescale for the whole chunk.
  template <int N> __device__ void add_chunk(const float (&x)[N], float cm) {
    float mn = fmaxf(m, cm);
    float ms = msafe(mn);
    float sc = (m == neg_inf()) ? 0.f : fast_exp(msafe(m) - ms);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) acc += fast_exp(x[i] - ms);
    s = s * sc + acc;
    m = mn;
  }
  // The same with the terms given in DOUBLE (exact sums alpha + w of two floats): the running
  // maximum stays a float (it is only a shift), every exponent x - ms is formed in double and
  // rounded once, so a term near the maximum carries no rounding error of the sum alpha + w
  // (which is ulp(|alpha + w|) ~ 2e-6 at |w| ~ 30 when formed in float).
  __device__ void add_d(double x) {
    const float xf = (float)x;
    float mn = fmaxf(m, xf);
    float ms = msafe(mn);
    float sc = (m == neg_inf()) ? 0.f : fast_exp(msafe(m) - ms);
    s = s * sc + fast_exp((float)(x - (double)ms));
    m = mn;
  }
  template <int N> __device__ void add_chunk_d(const double (&x)[N], float cm) {
    float mn = fmaxf(m, cm);
    float ms = msafe(mn);
    float sc = (m == neg_inf()) ? 0.f : fast_exp(msafe(m) - ms);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) acc += fast_exp((float)(x[i] - (double)ms));
    s = s * sc + acc;
    m = mn;
  }
  __device__ void merge(const Acc& o) {
    float mn = fmaxf(m, o.m);
    float ms = msafe(mn);
    float sa = (m == neg_inf()) ? 0.f : fast_exp(msafe(m) - ms);
    float sb = (o.m == neg_inf()) ? 0.f : fast_exp(msafe(o.m) - ms);
    s = s * sa + o.s * sb;
    m = mn;
  }
  __device__ float value() const { return msafe(m) + fast_log(s); }
  __device__ int arg() const { return 0; }
};

template <> struct Acc<LT_MAXTROPICAL> {
  float m; int a;
  __device__ void init() { m = neg_inf(); a = 0; }
  // callers feed candidates in ASCENDING index order, so strict '>' keeps the
  // first arg-max (torch.argmax semantics, semirings.py:382).
  __device__ void add(float x, int idx) { if (x > m) { m = x; a = idx; } }
  __device__ void merge(const Acc& o) {
    if (o.m > m || (o.m == m && o.a < a)) { m = o.m; a = o.a; }
  }
  __device__ float value() const { return m; }
  __device__ int arg() const { return a; }
};

// ---------------------------------------------------------------------------
// Thread-block cluster helpers (raw PTX; cluster size 1 is legal too).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive_release();
  cluster_wait_acquire();
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// Address of the same shared-memory variable in CTA `rank` of the cluster.
__device__ __forceinline__ uint32_t map_shared_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_shared_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}

// Streaming (read-once) global loads: keep them out of L1.
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// tanh(a + b) from the factors ea = e^(2a), eb = e^(2b) (joint exponential tables):
//   tanh = 1 - 2 / (1 + ea * eb): ONE MUFU op (the reciprocal) and two FMA-pipe ops.
// The product may overflow to +inf (1 / inf = 0: tanh -> 1) or flush to 0 (tanh -> -1) -- that IS
// the saturated tanh -- so the tables only keep each FACTOR finite and non-zero: exponents clamped
// to +-126 in log2 units, exact for projections up to |x| = 43.6 each (fp32 tanh has saturated at
// |a + b| > 9.1; round 1 clamped to +-63 to keep the product normal, which was wrong by up to
// 0.05 when one projection exceeded 21.8 and the other cancelled it to within 9).
__device__ __forceinline__ float rcp_1p(float e) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
__device__ __forceinline__ float tanh_from_exp(float ea, float eb) {
  return fmaf(-2.f, rcp_1p(ea * eb), 1.f);
}

// 256-bit global loads (sm_100: LDG.E.256; p must be 32-byte aligned).  A warp that reads 32
// bytes per lane with two 128-bit loads touches every 128-byte line twice (16-byte accesses at a
// 32-byte lane stride): twice the L1 data-pipe wavefronts of one 256-bit load.
__device__ __forceinline__ void ldg_stream8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]),
                 "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void ldg_cached8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]),
                 "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// after a <<<>>> launch: count it and surface launch errors
#define LT_LAUNCHED()                   \
  do {                                  \
    ::lt::note_launch();                \
    LT_CUDA(cudaGetLastError());        \
  } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace lt
