// K4: JointWeightFn vocabulary projection
// (/root/reference/last_torch/weight_fns.py:194-227), whole-utterance form:
//   joint[m, :]   = tanh(proj_ctx[c, :] + proj_frame[n, :])      m = n*C + c
//   lexical[m, v] = joint[m, :] . w_vocab[v, :] + b_vocab[v]
//   blank[m]      = joint[m, :] . w_blank       + b_blank
// The [M, H] joint is never written to memory: it is generated on the fly as
// the A operand of the GEMM, and recomputed in the backward kernels.
//
// This translation unit holds the CUDA-core (fp32 FMA) implementation that
// serves every shape; joint_tc.cu adds the tcgen05 tensor-core path for the
// aligned large shapes and falls back to these kernels otherwise.
#include "common.cuh"
#include "params.cuh"
#include "joint_simt.cuh"

namespace lt {

__device__ __forceinline__ float tanh_acc(float x) { return tanhf(x); }

struct JointFwdA {
  const float* pc; const float* pf; int C, H;
  __device__ float operator()(int64_t m, int64_t k) const {
    const int64_t n = m / C; const int c = (int)(m % C);
    return tanh_acc(pc[(size_t)c * H + k] + pf[(size_t)n * H + k]);
  }
};
struct JointFwdB {
  const float* wv; const float* wb; int V, H;
  __device__ float operator()(int64_t k, int n) const {
    return n < V ? wv[(size_t)n * H + k] : wb[k];
  }
};
struct JointFwdEpi {
  float* blank; float* lexical; const float* bv; const float* bb; int V;
  __device__ void operator()(int64_t m, int n, float acc) const {
    if (n < V) lexical[(size_t)m * V + n] = acc + bv[n];
    else blank[m] = acc + __ldg(bb);
  }
};

// d pre-activation: (G . W) * (1 - h^2), reduced into the two projections.
struct JointBwd1A {
  const float* gl; const float* gb; int V;
  __device__ float operator()(int64_t m, int64_t k) const {
    return k < V ? gl[(size_t)m * V + k] : gb[m];
  }
};
struct JointBwd1B {
  const float* wv; const float* wb; int V, H;
  __device__ float operator()(int64_t k, int n) const {
    return k < V ? wv[(size_t)k * H + n] : wb[n];
  }
};
struct JointBwd1Epi {
  const float* pc; const float* pf; float* gpc; float* gpf; int C, H;
  __device__ void operator()(int64_t m, int j, float acc) const {
    const int64_t n = m / C; const int c = (int)(m % C);
    const float h = tanh_acc(pc[(size_t)c * H + j] + pf[(size_t)n * H + j]);
    const float gp = acc * (1.f - h * h);
    atomicAdd(gpc + (size_t)c * H + j, gp);
    atomicAdd(gpf + (size_t)n * H + j, gp);
  }
};

// d weights: out[v, j] = sum_m G[m, v] * h[m, j]; column j == H accumulates the bias.
struct JointBwd2A {
  const float* gl; const float* gb; int V;
  __device__ float operator()(int64_t v, int64_t m) const {
    return v < V ? gl[(size_t)m * V + v] : gb[m];
  }
};
struct JointBwd2B {
  const float* pc; const float* pf; int C, H;
  __device__ float operator()(int64_t m, int j) const {
    if (j == H) return 1.f;
    const int64_t n = m / C; const int c = (int)(m % C);
    return tanh_acc(pc[(size_t)c * H + j] + pf[(size_t)n * H + j]);
  }
};
struct JointBwd2Epi {
  float* gwv; float* gwb; float* gbv; float* gbb; int V, H;
  __device__ void operator()(int64_t v, int j, float acc) const {
    if (v < V) { if (j < H) atomicAdd(gwv + (size_t)v * H + j, acc); else atomicAdd(gbv + v, acc); }
    else { if (j < H) atomicAdd(gwb + j, acc); else atomicAdd(gbb, acc); }
  }
};

int joint_forward_simt(const float* pc, const float* pf, const float* wb, const float* bb,
                       const float* wv, const float* bv, int64_t N, int C, int H, int V,
                       float* blank, float* lexical, cudaStream_t stream) {
  const int64_t M = N * C;
  if (M == 0) return LT_OK;
  // grid.y is limited to 65535: fold extra rows into several launches
  const int64_t rows_per_launch = 65535ll * 64;
  for (int64_t r0 = 0; r0 < M; r0 += rows_per_launch) {
    const int64_t rows = min(rows_per_launch, M - r0);
    JointFwdA A{pc, pf, C, H};
    JointFwdB B{wv, wb, V, H};
    JointFwdEpi E{blank, lexical, bv, bb, V};
    // offset by r0 through shifted functors (r0 is a multiple of 64 but not of C)
    auto a = [=] __device__(int64_t m, int64_t k) { return A(m + r0, k); };
    auto e = [=] __device__(int64_t m, int n, float acc) { E(m + r0, n, acc); };
    dim3 g((V + 1 + 63) / 64, (unsigned)((rows + 63) / 64), 1);
    tile_gemm_kernel<<<g, 256, 0, stream>>>(rows, V + 1, (int64_t)H, (int64_t)H, a, B, e);
    LT_LAUNCHED();
  }
  return LT_OK;
}

int joint_backward_simt(const float* pc, const float* pf, const float* wb, const float* wv,
                        const float* gb, const float* gl, int64_t N, int C, int H, int V,
                        float* gpc, float* gpf, float* gwb, float* gbb, float* gwv, float* gbv,
                        int parts, cudaStream_t stream) {
  // parts: bit 0 = dgrad (grad_proj_ctx / grad_proj_frame), bit 1 = wgrad (weights, biases)
  const int64_t M = N * C;
  if (M == 0) return LT_OK;
  const int64_t rows_per_launch = 65535ll * 64;
  for (int64_t r0 = 0; (parts & 1) && r0 < M; r0 += rows_per_launch) {
    const int64_t rows = min(rows_per_launch, M - r0);
    JointBwd1A A{gl, gb, V};
    JointBwd1B B{wv, wb, V, H};
    JointBwd1Epi E{pc, pf, gpc, gpf, C, H};
    auto a = [=] __device__(int64_t m, int64_t k) { return A(m + r0, k); };
    auto e = [=] __device__(int64_t m, int j, float acc) { E(m + r0, j, acc); };
    dim3 g((H + 63) / 64, (unsigned)((rows + 63) / 64), 1);
    tile_gemm_kernel<<<g, 256, 0, stream>>>(rows, H, (int64_t)(V + 1), (int64_t)(V + 1), a, B, e);
    LT_LAUNCHED();
  }
  if (parts & 2) {
    // split the M-long reduction over blockIdx.z
    int64_t kchunk = 4096;
    int64_t nsplit = (M + kchunk - 1) / kchunk;
    if (nsplit > 65535) { kchunk = (M + 65534) / 65535; kchunk = (kchunk + 15) / 16 * 16; nsplit = (M + kchunk - 1) / kchunk; }
    JointBwd2A A{gl, gb, V};
    JointBwd2B B{pc, pf, C, H};
    JointBwd2Epi E{gwv, gwb, gbv, gbb, V, H};
    dim3 g((H + 1 + 63) / 64, (V + 1 + 63) / 64, (unsigned)nsplit);
    tile_gemm_kernel<<<g, 256, 0, stream>>>((int64_t)(V + 1), H + 1, M, kchunk, A, B, E);
    LT_LAUNCHED();
  }
  return LT_OK;
}

}  // namespace lt

using namespace lt;

extern "C" int64_t lt_joint_workspace_bytes(int64_t N, int C, int H, int V) {
  // W_vocab as bf16 hi + lo, then the exponential tables of proj_ctx and proj_frame
  return joint_split_bytes(H, V) + joint_table_bytes(N, C, H);
}

extern "C" int lt_joint_forward(const float* proj_ctx, const float* proj_frame,
                                const float* w_blank, const float* b_blank,
                                const float* w_vocab, const float* b_vocab, int64_t N, int C,
                                int H, int V, float* blank, float* lexical, void* workspace,
                                void* stream) {
  LT_CHECK_ARG(N >= 0 && C > 0 && H > 0 && V > 0, "lt_joint_forward: bad sizes N=%lld C=%d H=%d V=%d",
               (long long)N, C, H, V);
  if (N == 0) return LT_OK;
  LT_CHECK_ARG(proj_ctx && proj_frame && w_blank && b_blank && w_vocab && b_vocab && blank && lexical,
               "lt_joint_forward: NULL pointer");
  if (workspace && reinterpret_cast<uintptr_t>(workspace) % 128 == 0 &&
      joint_forward_ts_supported(N, C, H, V, lexical))
    return joint_forward_ts_launch(proj_ctx, proj_frame, w_blank, b_blank, w_vocab, b_vocab, N, C,
                                   H, V, blank, lexical, workspace, (cudaStream_t)stream);
  if (workspace && reinterpret_cast<uintptr_t>(workspace) % 128 == 0 &&
      joint_tc_supported(N, C, H, V, proj_ctx, proj_frame, lexical))
    return joint_forward_tc_launch(proj_ctx, proj_frame, w_blank, b_blank, w_vocab, b_vocab, N, C,
                                   H, V, blank, lexical, workspace, (cudaStream_t)stream);
  return joint_forward_simt(proj_ctx, proj_frame, w_blank, b_blank, w_vocab, b_vocab, N, C, H, V,
                            blank, lexical, (cudaStream_t)stream);
}

extern "C" int lt_joint_backward_split_supported(int64_t N, int C, int H, int V) {
  return joint_backward_split_supported(N, C, H, V) ? 1 : 0;
}

extern "C" int64_t lt_joint_backward_workspace_bytes(int64_t N, int C, int H, int V) {
  return joint_backward_workspace_bytes(N, C, H, V);
}

extern "C" int lt_joint_backward(const float* proj_ctx, const float* proj_frame,
                                 const float* w_blank, const float* w_vocab,
                                 const float* grad_blank, const float* grad_lexical, int64_t N,
                                 int C, int H, int V, float* grad_proj_ctx,
                                 float* grad_proj_frame, float* grad_w_blank, float* grad_b_blank,
                                 float* grad_w_vocab, float* grad_b_vocab, void* workspace,
                                 int grad_lexical_format, void* stream) {
  LT_CHECK_ARG(N >= 0 && C > 0 && H > 0 && V > 0, "lt_joint_backward: bad sizes N=%lld C=%d H=%d V=%d",
               (long long)N, C, H, V);
  if (N == 0) return LT_OK;
  LT_CHECK_ARG(proj_ctx && proj_frame && w_blank && w_vocab && grad_blank && grad_lexical &&
               grad_proj_ctx && grad_proj_frame && grad_w_blank && grad_b_blank && grad_w_vocab &&
               grad_b_vocab, "lt_joint_backward: NULL pointer");
  LT_CHECK_ARG(grad_lexical_format == 0 || grad_lexical_format == 1,
               "lt_joint_backward: grad_lexical_format must be 0 (fp32) or 1 (split rows)");
  int simt_parts = 3;
  const bool ws_ok = workspace && reinterpret_cast<uintptr_t>(workspace) % 256 == 0;
  int split = grad_lexical_format;
  if (split) {
    LT_CHECK_ARG(ws_ok && joint_backward_split_supported(N, C, H, V) &&
                     reinterpret_cast<uintptr_t>(grad_lexical) % 32 == 0,
                 "lt_joint_backward: split-row grad_lexical is not supported for this shape "
                 "(ask lt_joint_backward_split_supported first)");
  }
  const float* dgrad_gl = grad_lexical;      // what the tensor-core kernels read
  // e^(2 proj) tables of the tensor-core kernels (second region of the workspace)
  float* ec = ws_ok ? reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) +
                                               joint_split_bytes(H, V)) : nullptr;
  float* ef = ws_ok ? ec + (size_t)C * H : nullptr;
  bool tables = false;
  if (ws_ok && joint_dgrad_tc_supported(N, C, H, V, grad_lexical, proj_ctx, proj_frame)) {
    int rc = joint_dgrad_tc_launch(proj_ctx, proj_frame, w_blank, w_vocab, grad_blank,
                                   dgrad_gl, split != 0, N, C, H, V, grad_proj_ctx,
                                   grad_proj_frame, workspace, (cudaStream_t)stream);  // + tables
    if (rc) return rc;
    simt_parts = 2;
    tables = true;
  }
  if (ws_ok && joint_wgrad_tc_supported(N, C, H, V, grad_lexical, proj_ctx, proj_frame)) {
    if (!tables) {
      int rc = joint_exp_tables_launch(proj_ctx, proj_frame, N, C, H, ec, ef, (cudaStream_t)stream);
      if (rc) return rc;
    }
    int rc = joint_wgrad_tc_launch(ec, ef, grad_blank, dgrad_gl, split, N, C, H, V,
                                   grad_w_blank, grad_b_blank, grad_w_vocab, grad_b_vocab,
                                   (cudaStream_t)stream);
    if (rc) return rc;
    simt_parts &= ~2;
  }
  if (simt_parts == 0) return LT_OK;
  if (split) {
    set_error("lt_joint_backward: split-row grad_lexical reached a CUDA-core kernel");
    return LT_ERR_UNSUPPORTED;
  }
  return joint_backward_simt(proj_ctx, proj_frame, w_blank, w_vocab, grad_blank, grad_lexical, N,
                             C, H, V, grad_proj_ctx, grad_proj_frame, grad_w_blank, grad_b_blank,
                             grad_w_vocab, grad_b_vocab, simt_parts, (cudaStream_t)stream);
}

// fp32 rows [M, V] -> "split rows" (every row [V bf16 hi | V bf16 lo] in the same V*4 bytes): the
// operand form lt_joint_backward takes with grad_lexical_format = 1.
extern "C" int lt_joint_split_rows(const float* rows, void* out, int64_t M, int V, void* stream) {
  LT_CHECK_ARG(M >= 0 && V > 0 && V % 8 == 0, "lt_joint_split_rows: need V %% 8 == 0 (M=%lld V=%d)",
               (long long)M, V);
  if (M == 0) return LT_OK;
  LT_CHECK_ARG(rows && out && reinterpret_cast<uintptr_t>(rows) % 32 == 0 &&
                   reinterpret_cast<uintptr_t>(out) % 16 == 0,
               "lt_joint_split_rows: NULL or misaligned pointer");
  return joint_split_rows_launch(rows, out, M, V, (cudaStream_t)stream);
}

// ---- north_star (4): JointWeightFn fused into the recursion (joint_lattice_fused.cu) ----------
extern "C" int lt_joint_lattice_fused_supported(int semiring, int vocab_size, int context_size,
                                                int max_expansions, int H) {
  return joint_lattice_fused_supported(semiring, vocab_size, context_size, max_expansions, H) ? 1 : 0;
}

extern "C" int lt_joint_lattice_forward_fused(int semiring, int vocab_size, const float* proj_ctx,
                                              const float* proj_frame, const float* w_blank,
                                              const float* b_blank, const float* w_vocab,
                                              const float* b_vocab, const int32_t* num_frames,
                                              int B, int T, int H, float* dist, float* alphas,
                                              float* alpha_final, int16_t* backptr, void* stream) {
  LT_CHECK_ARG(joint_lattice_fused_supported(semiring, vocab_size, 1, LT_FRAME_DEPENDENT, H),
               "lt_joint_lattice_forward_fused: needs Log / MaxTropical, vocab_size <= 64 and "
               "H in {32, 64, 128} (got semiring %d, V=%d, H=%d)", semiring, vocab_size, H);
  LT_CHECK_ARG(B >= 0 && T >= 0, "lt_joint_lattice_forward_fused: bad sizes B=%d T=%d", B, T);
  if (B == 0) return LT_OK;
  LT_CHECK_ARG(proj_ctx && w_blank && b_blank && w_vocab && b_vocab && num_frames && dist &&
                   (T == 0 || proj_frame),
               "lt_joint_lattice_forward_fused: NULL pointer");
  auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  LT_CHECK_ARG(al16(w_vocab) && al16(w_blank),
               "lt_joint_lattice_forward_fused: w_vocab / w_blank must be 16-byte aligned");
  return joint_lattice_forward_fused_launch(semiring, vocab_size, H, proj_ctx, proj_frame, w_blank,
                                            b_blank, w_vocab, b_vocab, num_frames, B, T, dist,
                                            alphas, alpha_final, backptr, (cudaStream_t)stream);
}
