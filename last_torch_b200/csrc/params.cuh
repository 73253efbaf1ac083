// Kernel parameter blocks and launcher prototypes shared by capi.cu and the
// kernel translation units.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

struct CUtensorMap_st;

namespace lt {

struct FwdParams {
  NGram g;
  int k;            // max_expansions, or -1 for FrameDependent
  int B, T;
  int dslice;       // destination states per CTA
  int ppad;         // capacity of the partial buffers (entries)
  const float* blank;
  const float* lexical;
  const int32_t* num_frames;
  const float* alpha_init;
  float* dist;
  float* alphas;
  float* alpha_final;
  float* levels;
  int16_t* backptr;
  uint8_t* termptr;
  int32_t* alpha_norm;   // renormalised recursion state (lt_lattice_forward_norm) or nullptr
  int wlevels;           // LT_FLAG_LEVEL_WEIGHTS: k + 1 sets of weights per frame, else 1
};

struct BwdParams {
  NGram g;
  int k;            // max_expansions, or -1 for FrameDependent
  int B, T;
  int dslice;
  int lpr;          // lanes per row (power of two <= 32)
  const float* blank;
  const float* lexical;
  const int32_t* num_frames;
  const float* alphas;
  const float* levels;
  const float* dist;
  const float* grad_dist;
  float* grad_blank;
  float* grad_lexical;
  float* beta_final;
  const int32_t* alpha_norm;
  int wlevels;           // LT_FLAG_LEVEL_WEIGHTS: weights AND gradients are [.., k + 1, C(, V)]
  // lt_lattice_expectation (fast path only): sum of posterior * value instead of gradients
  const float* value_blank;
  const float* value_lexical;
  double* expect_part;   // [B, cluster size]
};

struct StrParams {
  int k;       // max_expansions or -1
  int B, T, U1;
  const float* blank_w;
  const float* lexical_w;
  const int32_t* num_frames;
  const int32_t* num_labels;
  float* dist;
  float* alphas;
  uint8_t* backptr;
  // backward only
  const float* alphas_in;
  const uint8_t* backptr_in;
  const float* dist_in;
  const float* grad_dist;
  float* grad_blank_w;
  float* grad_lexical_w;
  // (integer part, fraction) representation of the Log chain (lt_string_forward_norm)
  int32_t* alpha_exp;            // [B,T,U1] forward out
  int32_t* dist_norm;            // [B,2]    forward out: integer part, bits of the fraction
  const int32_t* alpha_exp_in;   // backward in
  const int32_t* dist_norm_in;
};
// (e, f) kernels exist for the Log semiring: register kernels for FrameDependent with
// U1 <= 1024, a double-precision chain otherwise (48 bytes of shared memory per label state)
inline bool string_norm_supported(int semiring, int k, int U1) {
  return semiring == LT_LOG && U1 >= 1 && (size_t)U1 * 48 <= 200 * 1024;
}

struct VitParams {
  NGram g;
  int k;    // -1 FrameDependent
  int B, T;
  int frames_per_chunk;   // 0: read back-pointers straight from global memory
  const int16_t* backptr;
  const uint8_t* termptr;
  const float* alpha_final;
  const int32_t* num_frames;
  int32_t* labels;
  int32_t* path_states;
  const float* grad_dist;
  float* grad_blank;
  float* grad_lexical;
};

int lattice_forward_generic_launch(int semiring, const NGram& g, int k, const FwdParams& base,
                                   unsigned flags, int sm_count, cudaStream_t stream);
int lattice_backward_generic_launch(int semiring, const NGram& g, int k, const BwdParams& base,
                                    unsigned flags, int sm_count, cudaStream_t stream);
// TMA / cluster fast path (lattice_fast2.cu): bigram FrameDependent, V in {64..256}; two CTAs
// of different utterances per SM.  lexical == nullptr skips the alignment test.
bool lattice_fast2_supported(const NGram& g, int k, unsigned flags, const void* lexical);
// FrameLabelDependent(k <= 3) on the same bigram shapes (lattice_fast2_fld.cu)
bool lattice_fast2_fld_supported(const NGram& g, int k, unsigned flags, const void* lexical);
int lattice_forward_fld2_launch(int semiring, const NGram& g, const FwdParams& base,
                                unsigned flags, cudaStream_t stream);
int lattice_backward_fld2_launch(int semiring, const NGram& g, const BwdParams& base,
                                 unsigned flags, cudaStream_t stream);
// which kernel family keeps alpha renormalised for this lattice (FwdParams::alpha_norm):
// 0 none, 1 the TMA fast path (log2 units), 2 the generic kernels (natural-log units).
// lexical == nullptr assumes 16-byte aligned weights.
int lattice_norm_family(int semiring, const NGram& g, int k, unsigned flags, const void* lexical);
int lattice_forward_fast2_launch(int semiring, const NGram& g, const FwdParams& base,
                                 unsigned flags, cudaStream_t stream);
int lattice_backward_fast2_launch(int semiring, const NGram& g, const BwdParams& base,
                                  unsigned flags, cudaStream_t stream);
// thread-per-column TMA path for context_size >= 2 (lattice_cols.cu), forward only
bool lattice_cols_supported(const NGram& g, int k, unsigned flags, const void* lexical);
int lattice_forward_cols_launch(int semiring, const NGram& g, int k, const FwdParams& base,
                                cudaStream_t stream);
// 8-lanes-per-row TMA backward for context_size >= 2, FrameDependent (lattice_rows.cu)
bool lattice_rows_supported(const NGram& g, int k, unsigned flags, const void* lexical,
                            const void* grad_lexical);
int lattice_backward_rows_launch(int semiring, const NGram& g, const BwdParams& base, int sm_count,
                                 cudaStream_t stream);
int viterbi_launch(const VitParams& base, cudaStream_t stream);
int string_gather_launch(int V, int C, const float* blank, const float* lexical,
                         const int32_t* states, const int32_t* labels, int B, int T, int U1,
                         float* blank_w, float* lexical_w, cudaStream_t stream);
int string_scatter_launch(int V, int C, const float* gbw, const float* glw,
                          const int32_t* states, const int32_t* labels, int B, int T, int U1,
                          float scale, const float* utt_scale, float* gblank, float* glex,
                          int split, cudaStream_t stream);
int walk_states_launch(const NGram& g, const int32_t* labels, const int32_t* num_labels, int B,
                       int U, int32_t* states, int32_t* next_labels, int32_t* bad,
                       cudaStream_t stream);
int stream_delay_launch(unsigned ns, cudaStream_t stream);
int string_forward_launch(int semiring, const StrParams& p, cudaStream_t stream);
int string_backward_launch(int semiring, const StrParams& p, cudaStream_t stream);
int alphas_denormalize_launch(float* alphas, const int32_t* alpha_norm, int B, int T, int C,
                              cudaStream_t stream);
int semiring_plus_forward_launch(int sr, const float* a, const float* b, float* out, int64_t n,
                                 cudaStream_t stream);
int semiring_plus_backward_launch(int sr, const float* a, const float* b, const float* g,
                                  float* ga, float* gb, int64_t n, cudaStream_t stream);
int semiring_sum_forward_launch(int sr, const float* a, int64_t outer, int64_t R, int64_t inner,
                                float* out, int32_t* argmax, cudaStream_t stream);
int semiring_sum_backward_launch(int sr, const float* a, const float* out, const int32_t* argmax,
                                 const float* g, int64_t outer, int64_t R, int64_t inner,
                                 float* ga, cudaStream_t stream);
// tcgen05 joint projection (joint_tc.cu)
bool joint_tc_supported(int64_t N, int C, int H, int V, const void* pc, const void* pf,
                        const void* lexical);
int joint_forward_tc_launch(const float* pc, const float* pf, const float* wb, const float* bb,
                            const float* wv, const float* bv, int64_t N, int C, int H, int V,
                            float* blank, float* lexical, void* workspace, cudaStream_t stream);
// the same with the tanh operand in tensor memory (joint_fwd_ts.cu)
bool joint_forward_ts_supported(int64_t N, int C, int H, int V, const void* lexical);
int joint_forward_ts_launch(const float* pc, const float* pf, const float* wb, const float* bb,
                            const float* wv, const float* bv, int64_t N, int C, int H, int V,
                            float* blank, float* lexical, void* workspace, cudaStream_t stream);
int joint_split_weights_launch(const float* w, __nv_bfloat16* hi, __nv_bfloat16* lo, int n,
                               cudaStream_t stream);
int joint_exp_table_launch(const float* x, float* out, long long n, cudaStream_t stream);
bool joint_dgrad_tc_supported(int64_t N, int C, int H, int V, const void* gl, const void* pc,
                              const void* pf);
int64_t joint_backward_workspace_bytes(int64_t N, int C, int H, int V);
// workspace layout of both directions: [W_vocab bf16 hi | lo] [e^(2 pc) [C,H] | e^(2 pf) [N,H]] ...
int64_t joint_split_bytes(int H, int V);
int64_t joint_table_bytes(int64_t N, int C, int H);
int joint_exp_tables_launch(const float* pc, const float* pf, int64_t N, int C, int H, float* ec,
                            float* ef, cudaStream_t stream);
// split: grad_lexical rows are [V bf16 hi | V bf16 lo] (only the fused dgrad takes that form)
int joint_dgrad_tc_launch(const float* pc, const float* pf, const float* wb, const float* wv,
                          const float* gb, const float* gl, int split, int64_t N, int C, int H,
                          int V, float* gpc, float* gpf, void* workspace, cudaStream_t stream);
bool joint_backward_split_supported(int64_t N, int C, int H, int V);
bool joint_dgrad2_pair(int H, int V);      // map_g box: 64 frames instead of 128
int joint_dgrad2_multicast(int H, int V);  // map_g box: 128 / this frames
int joint_split_rows_launch(const float* g, void* out, int64_t M, int V, cudaStream_t stream);
// fused dgrad + reductions (joint_dgrad2.cu)
bool joint_dgrad2_supported(int64_t N, int C, int H, int V, const void* gl, const void* pc,
                            const void* pf);
int joint_dgrad2_launch(const CUtensorMap_st& map_hi, const CUtensorMap_st& map_lo,
                        const CUtensorMap_st& map_g, int split,
                        const float* pc, const float* pf, const float* wb, const float* gb,
                        const float* gl, int64_t N, int C, int H, int V, float* gpc, float* gpf,
                        cudaStream_t stream);
bool joint_wgrad_tc_supported(int64_t N, int C, int H, int V, const void* gl, const void* pc,
                              const void* pf);
// ec / ef: the exponential tables (joint_exp_tables_launch), not the projections
// split: grad_lexical rows are [V bf16 hi | V bf16 lo]
int joint_wgrad_tc_launch(const float* ec, const float* ef, const float* gb, const float* gl,
                          int split, int64_t N, int C, int H, int V, float* gwb, float* gbb,
                          float* gwv, float* gbv, cudaStream_t stream);
// JointWeightFn fused into the forward recursion (joint_lattice_fused.cu): inference direction
bool joint_lattice_fused_supported(int semiring, int V, int n, int k, int H);
int joint_lattice_forward_fused_launch(int semiring, int V, int H, const float* pc, const float* pf,
                                       const float* w_blank, const float* b_blank,
                                       const float* w_vocab, const float* b_vocab,
                                       const int32_t* num_frames, int B, int T, float* dist,
                                       float* alphas, float* alpha_final, int16_t* backptr,
                                       cudaStream_t stream);
int pick_cluster_size(const NGram& g, int B, unsigned flags, int sm_count);

}  // namespace lt
