// K5: Viterbi back-trace over the back-pointers written by the MaxTropical
// forward kernel.  Replaces "differentiate the tropical shortest distance
// w.r.t. a zero mask" (/root/reference/last_torch/lattices.py:221-247), which
// unrolls autograd through all T frames, by a pointer chase over
// [T, levels, C] int16 back-pointers staged through shared memory.
//
// Tie-breaking is fixed by the forward kernel (see lattice_forward.cu) and
// matches reference autograd: blank beats lexical (semirings.py:363), fewer
// expansions first (alignments.py:376), lowest source row-block first
// (semirings.py:382), first arg-max final state (lattices.py:496).
#include "common.cuh"
#include "params.cuh"

namespace lt {


// decode arc (source state, zero-based label) entering state q through
// row-block kk of the forward reduce
__device__ __forceinline__ void decode_arc(const NGram& g, int q, int kk, int* p, int* y) {
  const int lowV = g.Alow * g.V;
  int flat;
  if (q - g.off < lowV) flat = q - g.off;
  else flat = lowV + kk * g.N + (q - g.A);
  *p = flat / g.V;
  *y = flat % g.V;
}

__global__ void viterbi_backtrace_kernel(const VitParams p) {
  extern __shared__ __align__(16) unsigned char vsmem[];
  __shared__ float red_m[256];
  __shared__ int red_a[256];
  __shared__ int q_shared;
  const NGram& g = p.g;
  const int C = g.C, V = g.V;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nlev = p.k >= 1 ? p.k : 1;      // back-pointer levels per frame
  const int nlab = p.k >= 1 ? p.k + 1 : 1;  // alignment labels per frame
  const int nf = max(0, min(p.num_frames[b], p.T));

  // first arg-max of alpha_T
  float m = neg_inf(); int a = 0x7fffffff;
  for (int c = tid; c < C; c += blockDim.x) {
    const float v = p.alpha_final[(size_t)b * C + c];
    if (v > m || (v == m && c < a)) { m = v; a = c; }
  }
  if (a == 0x7fffffff) a = 0;
  red_m[tid] = m; red_a[tid] = a;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (tid < s) {
      const float om = red_m[tid + s]; const int oa = red_a[tid + s];
      if (om > red_m[tid] || (om == red_m[tid] && oa < red_a[tid])) { red_m[tid] = om; red_a[tid] = oa; }
    }
    __syncthreads();
  }
  if (tid == 0) q_shared = red_a[0];
  // labels default to 0 (blank / padding)
  for (size_t i = tid; i < (size_t)p.T * nlab; i += blockDim.x)
    p.labels[(size_t)b * p.T * nlab + i] = 0;
  __syncthreads();

  const float gscale = p.grad_dist ? p.grad_dist[b] : 1.f;
  const size_t bp_frame = (size_t)nlev * C;             // int16 entries per frame
  int16_t* s_bp = reinterpret_cast<int16_t*>(vsmem);
  uint8_t* s_term = vsmem + (size_t)p.frames_per_chunk * bp_frame * sizeof(int16_t);
  const int fpc = p.frames_per_chunk;

  int t_hi = nf;
  if (tid == 0 && p.path_states) {
    for (int t = nf; t <= p.T; ++t) p.path_states[(size_t)b * (p.T + 1) + t] = q_shared;
  }
  while (t_hi > 0) {
    const int t_lo = fpc > 0 ? max(0, t_hi - fpc) : 0;
    if (fpc > 0) {
      const int16_t* src = p.backptr + ((size_t)b * p.T + t_lo) * bp_frame;
      const size_t n = (size_t)(t_hi - t_lo) * bp_frame;
      for (size_t i = tid; i < n; i += blockDim.x) s_bp[i] = src[i];
      if (p.k >= 1) {
        const uint8_t* ts = p.termptr + ((size_t)b * p.T + t_lo) * C;
        const size_t nt = (size_t)(t_hi - t_lo) * C;
        for (size_t i = tid; i < nt; i += blockDim.x) s_term[i] = ts[i];
      }
      __syncthreads();
    }
    if (tid == 0) {
      int q = q_shared;
      for (int t = t_hi - 1; t >= t_lo; --t) {
        const size_t bt = (size_t)b * p.T + t;
        const int16_t* bp = fpc > 0 ? s_bp + (size_t)(t - t_lo) * bp_frame : p.backptr + bt * bp_frame;
        if (p.k < 1) {
          const int kk = bp[q];
          if (kk < 0) {
            if (p.grad_blank) p.grad_blank[bt * C + q] += gscale;
          } else {
            int src, y;
            decode_arc(g, q, kk, &src, &y);
            if (p.grad_lexical) p.grad_lexical[(bt * C + src) * V + y] += gscale;
            p.labels[bt] = y + 1;
            q = src;
          }
        } else {
          const uint8_t* tp = fpc > 0 ? s_term + (size_t)(t - t_lo) * C : p.termptr + bt * C;
          const int nexp = tp[q];
          if (p.grad_blank) p.grad_blank[bt * C + q] += gscale;
          for (int i = nexp - 1; i >= 0; --i) {
            const int kk = bp[(size_t)i * C + q];
            int src, y;
            decode_arc(g, q, kk, &src, &y);
            if (p.grad_lexical) p.grad_lexical[(bt * C + src) * V + y] += gscale;
            p.labels[bt * nlab + i] = y + 1;
            q = src;
          }
        }
        if (p.path_states) p.path_states[(size_t)b * (p.T + 1) + t] = q;
      }
      q_shared = q;
    }
    __syncthreads();
    t_hi = t_lo;
  }
}

int viterbi_launch(const VitParams& base, cudaStream_t stream) {
  VitParams p = base;
  if (p.B == 0) return LT_OK;
  const int nlev = p.k >= 1 ? p.k : 1;
  const size_t per_frame = (size_t)nlev * p.g.C * sizeof(int16_t) + (p.k >= 1 ? (size_t)p.g.C : 0);
  const size_t budget = 96 * 1024;
  int fpc = (int)(budget / per_frame);
  if (fpc > p.T) fpc = p.T;
  p.frames_per_chunk = fpc;
  // keep the uint8 region 2-byte aligned
  size_t smem = fpc > 0 ? (size_t)fpc * per_frame + 16 : 0;
  LT_CUDA(cudaFuncSetAttribute(viterbi_backtrace_kernel,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  viterbi_backtrace_kernel<<<p.B, 256, smem, stream>>>(p);
  LT_LAUNCHED();
  return LT_OK;
}

}  // namespace lt
