// north_star item (4) / SURVEY X1: JointWeightFn FUSED into the recursion -- the logits of a frame
// are produced on chip, consumed by the semiring update and never written to HBM.
//
//   alpha_{t+1}[1+y] = alpha_t[1+y] (x) blank_t[1+y]  (+)  (+)_p alpha_t[p] (x) lexical_t[p, y]
//   lexical_t[p, :] = tanh(proj_ctx[p] + proj_frame[t]) . W_vocab^T + b_vocab      (weight_fns.py:194-227,
//   blank_t[p]      = tanh(proj_ctx[p] + proj_frame[t]) . w_blank   + b_blank       lattices.py:436-462)
//
// Scope ("where shapes allow"): the INFERENCE direction -- Log shortest distance and MaxTropical
// distance + back-pointers for the Viterbi back-trace (RecognitionLattice.shortest_path needs no
// gradient) -- of a bigram FrameDependent lattice with vocab_size <= 64 and hidden_size <= 128,
// where one CTA can hold everything an utterance needs: e^(2 proj_ctx) [C, H] and the joint tile
// tanh(.) [C, H] in shared memory (67 KB), a slice of W_vocab in registers.  HBM traffic per
// utterance-frame drops from C*(V+1)*4 bytes of logits (written by K4, read by K1) to the H*4
// bytes of proj_frame; device memory from O(B*T*C*V) to O(B*T*C).
//
// One CTA per utterance, 4 threads per output column (the V lexical labels and the blank column):
// thread (j, hq) keeps W[j, h] for its quarter of the hidden units in registers and, for every
// source state p, forms its part of the dot product with the joint row (broadcast reads of shared
// memory), two shuffles complete the logit, and the semiring accumulator of destination 1 + j is
// updated in registers.  fp32 FMAs throughout: the logits are the reference's fp32 logits to
// accumulation order (no bf16 operand split as in the tensor-core kernels).
//
// This is the CUDA-core form: the measured answer to "does fusion pay" at the shape where
// recomputing the logits is cheapest: it does NOT -- see the numbers there.
#include "common.cuh"
#include "fast_ptx.cuh"
#include "params.cuh"

namespace lt {

namespace {

using namespace fastptx;

constexpr int kFusedMaxV = 64;
constexpr int kFusedMaxH = 128;

struct FusedParams {
  int V, H, B, T;
  const float* pc;        // [C, H]
  const float* pf;        // [B, T, H]
  const float* w_blank;   // [H]
  const float* b_blank;   // [1]
  const float* w_vocab;   // [V, H]
  const float* b_vocab;   // [V]
  const int32_t* num_frames;
  float* dist;            // [B]
  float* alphas;          // [B, T, C] or null
  float* alpha_final;     // [B, C] or null
  int16_t* backptr;       // [B, T, C] or null (MaxTropical)
};

// Running log-sum-exp of one destination: ONE MUFU.EX2 per term (the running maximum only moves
// a few times per frame), natural-log units.
struct FusedLogAcc {
  float m, s;
  __device__ void init() { m = neg_inf(); s = 0.f; }
  __device__ void add(float x, int) {
    if (x > m) {
      s = (m == neg_inf() ? 0.f : s * ex2((m - x) * kLog2e)) + 1.f;
      m = x;
    } else if (x > neg_inf()) {
      s += ex2((x - m) * kLog2e);
    }
  }
  __device__ float value() const { return m == neg_inf() ? m : m + __log2f(s) * kLn2; }
  __device__ int arg() const { return 0; }
};
template <int SR> struct FusedAcc { using type = Acc<SR>; };
template <> struct FusedAcc<LT_LOG> { using type = FusedLogAcc; };

// HQ = H / 4: hidden units per thread.  Thread hq of a column owns the float4 groups
// i * 4 + hq (i = 0 .. HQ/4 - 1), so the 4 threads of a column read 64 contiguous bytes of a
// joint row per step: conflict-free LDS.128.
template <int SR, int HQ>
__global__ void __launch_bounds__(288, 1)
joint_lattice_forward_fused_kernel(const FusedParams p) {
  using S = Sr<SR>;
  constexpr int H = HQ * 4;
  constexpr int JS = H + 4;                       // joint row stride (floats)
  extern __shared__ __align__(16) float fsm[];
  const int V = p.V, C = V + 1;
  float* ec = fsm;                                // [C][H]   e^(2 proj_ctx)
  float* jt = ec + (size_t)C * H;                 // [C][JS]  tanh(proj_ctx + proj_frame_t)
  float* ef = jt + (size_t)C * JS;                // [H]      e^(2 proj_frame_t)
  float* al = ef + H;                             // [2][C + 3] alpha ping-pong
  float* bl = al + 2 * (C + 3);                   // [C + 3]  blank_t[p]
  const int tid = threadIdx.x, nth = blockDim.x;
  const int b = blockIdx.x;
  const int j = tid >> 2, hq = tid & 3;           // output column, hidden quarter
  const bool col = j <= V;                        // j == V: the blank column
  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;

  for (int i = tid; i < C * H; i += nth) {
    float a = p.pc[i] * 2.8853900817779268f;      // e^(2x) = 2^(2 log2(e) x), clamped like the
    a = fminf(fmaxf(a, -126.f), 126.f);             // tensor-core kernels' tables (joint_tc.cu)
    ec[i] = exp2f(a);
  }
  for (int c = tid; c < C; c += nth) {
    al[c] = c == 0 ? S::one() : S::zero();
    al[C + 3 + c] = S::zero();
  }
  float w[HQ];
  float bias = 0.f;
  if (col) {
    const float* wrow = j < V ? p.w_vocab + (size_t)j * H : p.w_blank;
#pragma unroll
    for (int i = 0; i < HQ / 4; ++i) {
      const float4 v4 = *reinterpret_cast<const float4*>(wrow + (i * 4 + hq) * 4);
      w[4 * i] = v4.x; w[4 * i + 1] = v4.y; w[4 * i + 2] = v4.z; w[4 * i + 3] = v4.w;
    }
    bias = j < V ? p.b_vocab[j] : p.b_blank[0];
  } else {
#pragma unroll
    for (int i = 0; i < HQ; ++i) w[i] = 0.f;
  }
  __syncthreads();

  float* cur = al;
  float* nxt = al + C + 3;
  for (int t = 0; t < nf; ++t) {
    // ---- joint tile of this frame
    const float* pft = p.pf + (bt0 + t) * H;
    for (int h = tid; h < H; h += nth) {
      float a = pft[h] * 2.8853900817779268f;
      a = fminf(fmaxf(a, -126.f), 126.f);
      ef[h] = exp2f(a);
    }
    if (p.alphas)
      for (int c = tid; c < C; c += nth) p.alphas[(bt0 + t) * C + c] = cur[c];
    __syncthreads();
    for (int i = tid; i < C * (H / 4); i += nth) {            // H is a power of two
      const int c = i / (H / 4), h4 = (i - c * (H / 4)) * 4;
      const float4 a = *reinterpret_cast<const float4*>(ec + c * H + h4);
      const float4 f = *reinterpret_cast<const float4*>(ef + h4);
      *reinterpret_cast<float4*>(jt + c * JS + h4) =
          make_float4(tanh_from_exp(a.x, f.x), tanh_from_exp(a.y, f.y), tanh_from_exp(a.z, f.z),
                      tanh_from_exp(a.w, f.w));
    }
    __syncthreads();

    // ---- logits of column j for every source state p, consumed on the spot
    // (every thread runs the loop -- the shuffles need whole warps; threads past the last column
    // multiply by zero weights and store nothing)
    typename FusedAcc<SR>::type acc; acc.init();
    {
      auto consume = [&](int s, float d) {
        const float logit = d + bias;
        if (j < V) acc.add(S::times(cur[s], logit), s);       // arc s --(j+1)--> state 1 + j
        else if (j == V && hq == 0) bl[s] = logit;            // blank weight of state s
      };
      // four source rows per step: 8 independent FMA chains and 4 shuffle pairs in flight
      int s = 0;
      for (; s + 4 <= C; s += 4) {
        float d[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) { d[r][0] = 0.f; d[r][1] = 0.f; }
#pragma unroll
        for (int i = 0; i < HQ / 4; ++i) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float4 x = *reinterpret_cast<const float4*>(jt + (s + r) * JS + hq * 4 + i * 16);
            d[r][0] = fmaf(x.x, w[4 * i], d[r][0]); d[r][1] = fmaf(x.y, w[4 * i + 1], d[r][1]);
            d[r][0] = fmaf(x.z, w[4 * i + 2], d[r][0]); d[r][1] = fmaf(x.w, w[4 * i + 3], d[r][1]);
          }
        }
        float e[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) e[r] = d[r][0] + d[r][1];
#pragma unroll
        for (int r = 0; r < 4; ++r) e[r] += __shfl_xor_sync(0xffffffffu, e[r], 1);
#pragma unroll
        for (int r = 0; r < 4; ++r) e[r] += __shfl_xor_sync(0xffffffffu, e[r], 2);
#pragma unroll
        for (int r = 0; r < 4; ++r) consume(s + r, e[r]);
      }
      for (; s < C; ++s) {
        const float* row = jt + s * JS + hq * 4;
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int i = 0; i < HQ / 4; ++i) {
          const float4 x = *reinterpret_cast<const float4*>(row + i * 16);
          d0 = fmaf(x.x, w[4 * i], d0); d1 = fmaf(x.y, w[4 * i + 1], d1);
          d0 = fmaf(x.z, w[4 * i + 2], d0); d1 = fmaf(x.w, w[4 * i + 3], d1);
        }
        float dd = d0 + d1;
        dd += __shfl_xor_sync(0xffffffffu, dd, 1);
        dd += __shfl_xor_sync(0xffffffffu, dd, 2);
        consume(s, dd);
      }
    }
    __syncthreads();
    if (col && hq == 0) {
      if (j < V) {
        const int q = 1 + j;
        const float a = S::times(cur[q], bl[q]);
        float v;
        if constexpr (SR == LT_MAXTROPICAL) {
          const bool take_blank = a >= acc.value();            // semirings.py:363
          v = take_blank ? a : acc.value();
          if (p.backptr) p.backptr[(bt0 + t) * C + q] = take_blank ? (int16_t)-1 : (int16_t)acc.arg();
        } else {
          acc.add(a, 0);
          v = acc.value();
        }
        nxt[q] = v;
      } else {                                                 // state 0: blank self-loop only
        nxt[0] = S::times(cur[0], bl[0]);
        if constexpr (SR == LT_MAXTROPICAL) { if (p.backptr) p.backptr[(bt0 + t) * C] = (int16_t)-1; }
      }
    }
    __syncthreads();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  // padding frames keep alpha (lattices.py:460-461) and are still recorded (:462)
  for (int c = tid; c < C; c += nth) {
    if (p.alphas)
      for (int t = nf; t < p.T; ++t) p.alphas[(bt0 + t) * C + c] = cur[c];
    if (p.alpha_final) p.alpha_final[(size_t)b * C + c] = cur[c];
  }
  if (tid < 32) {                                   // dist = (+)_c alpha_T[c]  (lattices.py:496)
    Acc<SR> acc; acc.init();
    for (int c = tid; c < C; c += 32) acc.add(cur[c], c);
    for (int o = 16; o > 0; o >>= 1) {
      Acc<SR> other = acc;
      if constexpr (SR == LT_LOG) {
        other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
        other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
      } else {
        other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
        other.a = __shfl_xor_sync(0xffffffffu, acc.a, o);
      }
      acc.merge(other);
    }
    if (tid == 0) p.dist[b] = acc.value();
  }
}

template <int SR, int HQ>
static int launch_fused(const FusedParams& p, cudaStream_t stream) {
  const int C = p.V + 1, H = HQ * 4;
  const size_t smem = sizeof(float) * ((size_t)C * H + (size_t)C * (H + 4) + H + 3 * (C + 3)) + 64;
  LT_CUDA(cudaFuncSetAttribute(joint_lattice_forward_fused_kernel<SR, HQ>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = round_up((p.V + 1) * 4, 32);
  joint_lattice_forward_fused_kernel<SR, HQ><<<p.B, threads, smem, stream>>>(p);
  LT_LAUNCHED();
  return LT_OK;
}

}  // namespace

bool joint_lattice_fused_supported(int semiring, int V, int n, int k, int H) {
  if (semiring != LT_LOG && semiring != LT_MAXTROPICAL) return false;
  if (n != 1 || k >= 1) return false;                     // bigram, FrameDependent
  if (V < 1 || V > kFusedMaxV) return false;
  return H == 32 || H == 64 || H == 128;
}

int joint_lattice_forward_fused_launch(int semiring, int V, int H, const float* pc, const float* pf,
                                       const float* w_blank, const float* b_blank,
                                       const float* w_vocab, const float* b_vocab,
                                       const int32_t* num_frames, int B, int T, float* dist,
                                       float* alphas, float* alpha_final, int16_t* backptr,
                                       cudaStream_t stream) {
  FusedParams p = {};
  p.V = V; p.H = H; p.B = B; p.T = T;
  p.pc = pc; p.pf = pf; p.w_blank = w_blank; p.b_blank = b_blank; p.w_vocab = w_vocab;
  p.b_vocab = b_vocab; p.num_frames = num_frames; p.dist = dist; p.alphas = alphas;
  p.alpha_final = alpha_final; p.backptr = semiring == LT_MAXTROPICAL ? backptr : nullptr;
#define LT_FUSED(SR)                                         \
  switch (H) {                                               \
    case 32: return launch_fused<SR, 8>(p, stream);          \
    case 64: return launch_fused<SR, 16>(p, stream);         \
    default: return launch_fused<SR, 32>(p, stream);         \
  }
  if (semiring == LT_LOG) { LT_FUSED(LT_LOG) }
  LT_FUSED(LT_MAXTROPICAL)
#undef LT_FUSED
}

}  // namespace lt
