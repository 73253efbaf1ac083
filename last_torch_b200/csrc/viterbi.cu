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
#include "fast_ptx.cuh"
#include "params.cuh"

namespace lt {

using namespace fastptx;


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

// The back-pointers of one utterance are contiguous ([T][levels][C] int16 + [T][C] uint8).
// They are pulled through shared memory in chunks of `frames_per_chunk` frames with ONE bulk
// (TMA) copy per chunk and array, double-buffered: while thread 0 chases the pointers of a
// chunk, the previous frames are already in flight.  Bulk copies need 16-byte aligned
// addresses and sizes, frames do not start on such boundaries (C is odd), so a chunk is
// fetched from the aligned address below it and the data starts `head` bytes into the buffer.
struct BulkSpan {
  const unsigned char* src;   // 16-byte aligned
  uint32_t bytes;             // multiple of 16 (may be 0)
  uint32_t head;              // offset of the first wanted byte
  uint32_t tail_from, tail_bytes;   // the last < 16 bytes, copied by hand when the rounded-up
                                    // span would leave the tensor
};

__device__ __forceinline__ BulkSpan make_span(const unsigned char* base, size_t first, size_t nbytes,
                                              size_t total_bytes) {
  BulkSpan sp;
  const uintptr_t a0 = reinterpret_cast<uintptr_t>(base + first);
  const uintptr_t al = a0 & ~(uintptr_t)15;
  sp.src = reinterpret_cast<const unsigned char*>(al);
  sp.head = (uint32_t)(a0 - al);
  const uintptr_t end = a0 + nbytes;
  const uintptr_t tensor_end = reinterpret_cast<uintptr_t>(base) + total_bytes;
  uintptr_t end_al = (end + 15) & ~(uintptr_t)15;
  sp.tail_from = 0; sp.tail_bytes = 0;
  if (end_al > tensor_end) {            // never read past the tensor
    end_al = end & ~(uintptr_t)15;
    if (end_al < al) end_al = al;
    sp.tail_from = (uint32_t)(end_al - al);
    sp.tail_bytes = (uint32_t)(end - end_al);
  }
  sp.bytes = (uint32_t)(end_al - al);
  return sp;
}

__global__ void __launch_bounds__(256)
viterbi_backtrace_kernel(const VitParams p) {
  extern __shared__ __align__(128) unsigned char vsmem[];
  __shared__ float red_m[256];
  __shared__ int red_a[256];
  __shared__ int q_shared;
  __shared__ __align__(8) uint64_t bars[2];
  const NGram& g = p.g;
  const int C = g.C, V = g.V;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nlev = p.k >= 1 ? p.k : 1;      // back-pointer levels per frame
  const int nlab = p.k >= 1 ? p.k + 1 : 1;  // alignment labels per frame
  const int nf = max(0, min(p.num_frames[b], p.T));
  const bool fld = p.k >= 1;

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
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
  const int fpc = p.frames_per_chunk;
  // buffer layout: [2][bp_region | term_region], every region 16-byte aligned with 32 B slack
  const size_t bp_region = ((size_t)fpc * bp_frame * 2 + 32 + 15) & ~(size_t)15;
  const size_t tm_region = fld ? (((size_t)fpc * C + 32 + 15) & ~(size_t)15) : 0;
  const size_t buf_bytes = bp_region + tm_region;

  if (tid == 0 && p.path_states) {
    for (int t = nf; t <= p.T; ++t) p.path_states[(size_t)b * (p.T + 1) + t] = q_shared;
  }
  if (tid != 0 || nf == 0) return;

  // ------------------------------------------------------------ single chaser
  const unsigned char* bp_base = reinterpret_cast<const unsigned char*>(p.backptr);
  const unsigned char* tm_base = reinterpret_cast<const unsigned char*>(p.termptr);
  const size_t bp_total = (size_t)p.B * p.T * bp_frame * 2;
  const size_t tm_total = (size_t)p.B * p.T * C;
  const int nchunk = fpc > 0 ? (nf + fpc - 1) / fpc : 0;
  BulkSpan sb[2], st[2];
  auto issue = [&](int ci) {          // chunk ci covers frames [t_hi - fpc, t_hi), going backwards
    const int t_hi = nf - ci * fpc, t_lo = max(0, t_hi - fpc);
    const int s = ci & 1;
    unsigned char* dst = vsmem + (size_t)s * buf_bytes;
    sb[s] = make_span(bp_base, ((size_t)b * p.T + t_lo) * bp_frame * 2,
                      (size_t)(t_hi - t_lo) * bp_frame * 2, bp_total);
    uint32_t tx = sb[s].bytes;
    if (fld) {
      st[s] = make_span(tm_base, ((size_t)b * p.T + t_lo) * C, (size_t)(t_hi - t_lo) * C, tm_total);
      tx += st[s].bytes;
    }
    const uint32_t bar = smem_u32(&bars[s]);
    fence_proxy_async();     // the buffer was read with ordinary loads two chunks ago
    mbar_arrive_expect_tx(bar, tx);
    if (sb[s].bytes) bulk_load_1d(smem_u32(dst), sb[s].src, sb[s].bytes, bar);
    if (fld && st[s].bytes) bulk_load_1d(smem_u32(dst + bp_region), st[s].src, st[s].bytes, bar);
    // the few trailing bytes a rounded-up copy would read past the tensor
    for (uint32_t i = 0; i < sb[s].tail_bytes; ++i) dst[sb[s].tail_from + i] = sb[s].src[sb[s].tail_from + i];
    if (fld)
      for (uint32_t i = 0; i < st[s].tail_bytes; ++i)
        dst[bp_region + st[s].tail_from + i] = st[s].src[st[s].tail_from + i];
  };

  int q = q_shared;
  if (nchunk > 0) issue(0);
  int t_hi = nf;
  for (int ci = 0; t_hi > 0; ++ci) {
    const int t_lo = fpc > 0 ? max(0, t_hi - fpc) : 0;
    const int16_t* s_bp = nullptr;
    const uint8_t* s_term = nullptr;
    if (fpc > 0) {
      if (ci + 1 < nchunk) issue(ci + 1);
      const int s = ci & 1;
      mbar_wait(smem_u32(&bars[s]), (ci >> 1) & 1);
      const unsigned char* buf = vsmem + (size_t)s * buf_bytes;
      s_bp = reinterpret_cast<const int16_t*>(buf + sb[s].head);
      if (fld) s_term = buf + bp_region + st[s].head;
    }
    for (int t = t_hi - 1; t >= t_lo; --t) {
      const size_t bt = (size_t)b * p.T + t;
      const int16_t* bp = fpc > 0 ? s_bp + (size_t)(t - t_lo) * bp_frame : p.backptr + bt * bp_frame;
      if (!fld) {
        const int kk = bp[q];
        if (kk < 0) {
          if (p.grad_blank) atomicAdd(p.grad_blank + bt * C + q, gscale);
        } else {
          int src, y;
          decode_arc(g, q, kk, &src, &y);
          if (p.grad_lexical) atomicAdd(p.grad_lexical + (bt * C + src) * V + y, gscale);
          p.labels[bt] = y + 1;
          q = src;
        }
      } else {
        const uint8_t* tp = fpc > 0 ? s_term + (size_t)(t - t_lo) * C : p.termptr + bt * C;
        const int nexp = tp[q];
        if (p.grad_blank) atomicAdd(p.grad_blank + bt * C + q, gscale);
        for (int i = nexp - 1; i >= 0; --i) {
          const int kk = bp[(size_t)i * C + q];
          int src, y;
          decode_arc(g, q, kk, &src, &y);
          if (p.grad_lexical) atomicAdd(p.grad_lexical + (bt * C + src) * V + y, gscale);
          p.labels[bt * nlab + i] = y + 1;
          q = src;
        }
      }
      if (p.path_states) p.path_states[(size_t)b * (p.T + 1) + t] = q;
    }
    t_hi = t_lo;
  }
}

int viterbi_launch(const VitParams& base, cudaStream_t stream) {
  VitParams p = base;
  if (p.B == 0) return LT_OK;
  const int nlev = p.k >= 1 ? p.k : 1;
  const size_t bp_frame_bytes = (size_t)nlev * p.g.C * sizeof(int16_t);
  const size_t tm_frame_bytes = p.k >= 1 ? (size_t)p.g.C : 0;
  // two buffers of at most ~100 KB each; 0 frames per chunk = chase straight from global memory
  const size_t budget = 100 * 1024;
  int fpc = (int)((budget - 96) / (bp_frame_bytes + tm_frame_bytes));
  if (fpc > p.T) fpc = p.T;
  if (reinterpret_cast<uintptr_t>(p.backptr) % 2 != 0) fpc = 0;
  p.frames_per_chunk = fpc;
  size_t smem = 0;
  if (fpc > 0) {
    const size_t bp_region = ((size_t)fpc * bp_frame_bytes + 32 + 15) & ~(size_t)15;
    const size_t tm_region = p.k >= 1 ? (((size_t)fpc * p.g.C + 32 + 15) & ~(size_t)15) : 0;
    smem = 2 * (bp_region + tm_region);
  }
  LT_CUDA(cudaFuncSetAttribute(viterbi_backtrace_kernel,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  viterbi_backtrace_kernel<<<p.B, 256, smem, stream>>>(p);
  LT_LAUNCHED();
  return LT_OK;
}

}  // namespace lt
