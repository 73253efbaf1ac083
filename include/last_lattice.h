/*
 * last_lattice.h -- C ABI of the B200 (sm_100a) lattice kernels.
 *
 * This is the drop-in boundary for the hot path of theadamsabra/last_torch:
 * the semiring forward/backward recursion over the frames x context-states
 * recognition lattice.  The reference has no FFI of its own (it is pure
 * Python); each entry point below replaces the reference Python function that
 * is cited beside it (file:line under /root/reference/last_torch/), and is what
 * a ctypes binding inside the reference would call.
 *
 * Conventions
 *   - all pointers are DEVICE pointers to dense row-major fp32 / int32 arrays
 *     owned by the caller; nothing is allocated or freed by the library;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *   - every function returns LT_OK (0) or an error code; the message is
 *     available from lt_last_error() (thread-local, so the autograd thread
 *     and the main thread do not race);
 *   - there is no global mutable state: calls are re-entrant.
 *
 * Shapes:  B utterances, T frames, V = vocab_size, n = context_size,
 *   C = sum_{i<=n} V^i context states (FullNGram, contexts.py:181-182),
 *   k = max_expansions (>= 1) for FrameLabelDependent, or LT_FRAME_DEPENDENT.
 *   blank   [B,T,C]     lexical [B,T,C,V]   (weight_fns.py:66-75)
 */
#ifndef LAST_LATTICE_H_
#define LAST_LATTICE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LT_OK 0
#define LT_ERR_INVALID_ARGUMENT 1
#define LT_ERR_CUDA 2
#define LT_ERR_UNSUPPORTED 3

/* semirings.py:143-173 (Real), :184-305 (Log), :308-401 (MaxTropical) */
#define LT_REAL 0
#define LT_LOG 1
#define LT_MAXTROPICAL 2

/* alignments.py:266-329 FrameDependent; any k >= 1 selects
 * FrameLabelDependent(max_expansions=k), alignments.py:331-432 */
#define LT_FRAME_DEPENDENT (-1)

/* flags for lt_lattice_forward / lt_lattice_backward */
#define LT_FLAG_FORCE_GENERIC 1u     /* never take the TMA/cluster fast path   */
#define LT_FLAG_CLUSTER_SHIFT 8      /* bits 8..11: force cluster size (1,2,4,8) */
#define LT_FLAG_LEVEL_WEIGHTS 32u    /* FrameLabelDependent(k) with one set of weights per alignment
                                        state (lattices.py:447-453, zero-valued masks added per
                                        state): blank is [B,T,k+1,C], lexical [B,T,k+1,C,V] (level i
                                        = blank[i] / lexical[i] of alignments.py:362-376; lexical[k]
                                        is unused) and the gradients have the same layout.  Generic
                                        kernels only (implies LT_FLAG_FORCE_GENERIC).               */
#define LT_FLAG_GRAD_SPLIT 16u       /* lt_lattice_backward: write grad_lexical as "split rows"
                                        (see lt_joint_backward); only when
                                        lt_lattice_backward_split_supported() returns 1 */

int lt_version(void);
const char* lt_last_error(void);
/* Number of CUDA kernels this library has launched in this process so far. */
unsigned long long lt_launch_count(void);
/* Debug / test switches, process-wide.  Each is initialised once from the environment variable
 * of the same name and can be changed at run time: LT_JOINT_SIMT, LT_JOINT_DGRAD_V1,
 * LT_JOINT_WGRAD_SIMT (CUDA-core / first-generation joint kernels), LT_JOINT_DGRAD_PAIR,
 * LT_JOINT_DGRAD_MULTICAST (measured-slower variants of the split-row dgrad), LT_JOINT_FWD_SS
 * (forward projection with the tanh operand staged in shared instead of tensor memory), LT_TABLE_V1,
 * LT_TABLE_CLUSTER (NextStateTable kernel selection).  lt_get_option returns -1 for an unknown
 * name. */
int lt_set_option(const char* name, int value);
int lt_get_option(const char* name);
/* Number of SMs, compute capability of the current device. */
int lt_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- K1: RecognitionLattice._forward (lattices.py:379-496, loop :856-892) ----
 * One persistent CTA cluster per utterance keeps alpha on chip for all T
 * frames.  alpha_{t+1} = alignment.forward(alpha_t, blank_t, lexical_t)
 * (alignments.py:294-297 / :370-376) with FullNGram.forward_reduce
 * (contexts.py:207-230); frames t >= num_frames[b] leave alpha unchanged
 * (lattices.py:460-461).
 *   alpha_init  [B,C] or NULL (NULL: one-hot semiring-one at the start state,
 *               lattices.py:801-807)
 *   dist        [B]      (+)_c alpha_T[c]                  (lattices.py:496)
 *   alphas      [B,T,C]  alpha_0 .. alpha_{T-1}, or NULL
 *   alpha_final [B,C]    alpha_T, or NULL
 *   levels      [B,T,k,C] FrameLabelDependent only: the k intermediate
 *               `last` vectors of alignments.py:372-375 (needed by the
 *               backward kernel); NULL otherwise
 *   backptr     [B,T,max(k,1),C] int16, MaxTropical only or NULL: arg-max
 *               source row-block of each forward_reduce (semirings.py:380-386);
 *               for FrameDependent -1 means "blank arc won" (semirings.py:363)
 *   termptr     [B,T,C] uint8, MaxTropical + FrameLabelDependent only: number
 *               of lexical expansions taken (first arg-max, alignments.py:376)
 */
int lt_lattice_forward(int semiring, int vocab_size, int context_size,
                       int max_expansions, const float* blank,
                       const float* lexical, const int32_t* num_frames, int B,
                       int T, const float* alpha_init, float* dist,
                       float* alphas, float* alpha_final, float* levels,
                       int16_t* backptr, uint8_t* termptr, unsigned flags,
                       void* stream);

/* ---- K2: backward (beta) recursion + arc posteriors as weight gradients ----
 * Intent of RecognitionLattice._backward (lattices.py:686-799) with
 * alignment.backward (alignments.py:300-318 / :378-418) and
 * FullNGram.backward_broadcast (contexts.py:232-256); replaces autograd through
 * the unrolled loop.  semiring is LT_LOG (marginals * grad_dist) or LT_REAL.
 *   grad_blank [B,T,C], grad_lexical [B,T,C,V]: fully overwritten (padding
 *   frames get zeros, lattices.py:775-779).
 *   beta_final [B,C] or NULL: beta_0 (for diagnostics / chunked use).
 */
int lt_lattice_backward(int semiring, int vocab_size, int context_size,
                        int max_expansions, const float* blank,
                        const float* lexical, const int32_t* num_frames, int B,
                        int T, const float* alphas, const float* levels,
                        const float* dist, const float* grad_dist,
                        float* grad_blank, float* grad_lexical,
                        float* beta_final, unsigned flags, void* stream);

/* ---- K1 / K2 with a RENORMALISED recursion state (Log semiring) -------------------
 * Same recursions, but alpha is kept as alpha_t = alpha~_t + off_t with an exact integer offset
 * (in log2 units) that follows floor(max_c alpha~_t[c]) from frame to frame, so every sum the
 * recursion rounds has magnitude O(10) instead of O(logZ).  At T = 1000 (logZ ~ 5.5e3, one fp32
 * ulp = 4.9e-4) the arc posteriors then agree with an fp64 evaluation to ~1e-6 relative; the
 * plain fp32 recursion -- and the fp32 reference, lattices.py:865-886 + autograd -- is at
 * 1e-4 .. 1e-3 there (profiles/r02_parity_errors.json).
 *   alpha_norm [B, T+3] int32: off_0 .. off_T, then the bits of r (fp32), then the unit u of the
 *              offsets (0: log2 units -- the TMA fast path works in log2; 1: natural log -- the
 *              generic kernels): logZ = off_T * (u ? 1 : ln 2) + r' with r' = r in the same unit.
 *              Written by the forward, read by the backward of the same lattice.
 *   alphas     [B,T,C] then holds alpha~_t (natural-log units); the TRUE alpha_t is
 *              alphas[b,t,c] + alpha_norm[b,t] * (u ? 1 : ln 2)  (lt_alphas_denormalize).
 * Only when lt_lattice_norm_supported() returns 1 (Log semiring; the bigram TMA fast path, the
 * context_size >= 2 thread-per-column forward where its 8-lanes-per-row backward also applies, and
 * the generic kernels: every FullNGram lattice except thread-per-column shapes whose vocabulary
 * the row backward does not cover);
 * alpha_norm == NULL selects the plain kernels (identical to lt_lattice_forward / _backward).
 */
int lt_lattice_norm_supported(int semiring, int vocab_size, int context_size,
                              int max_expansions, unsigned flags);
int lt_lattice_forward_norm(int semiring, int vocab_size, int context_size,
                            int max_expansions, const float* blank,
                            const float* lexical, const int32_t* num_frames, int B,
                            int T, const float* alpha_init, float* dist,
                            float* alphas, float* alpha_final, float* levels,
                            int16_t* backptr, uint8_t* termptr, int32_t* alpha_norm,
                            unsigned flags, void* stream);
int lt_lattice_backward_norm(int semiring, int vocab_size, int context_size,
                             int max_expansions, const float* blank,
                             const float* lexical, const int32_t* num_frames, int B,
                             int T, const float* alphas, const float* levels,
                             const float* dist, const float* grad_dist,
                             float* grad_blank, float* grad_lexical,
                             float* beta_final, const int32_t* alpha_norm,
                             unsigned flags, void* stream);
/* Expected value of an additive arc function under the lattice's path distribution -- what the
 * first-order expectation semiring (semirings.py:404-484, Expectation / LogLogExpectation) computes
 * when it is run through the forward recursion, evaluated here by forward-backward in one pass
 * over the weights and WITHOUT materialising the arc posteriors:
 *     expect[b] = sum over arcs a of  posterior_b(a) * value(a)
 * value_blank [B,T,C] / value_lexical [B,T,C,V]; both NULL: value(a) = the arc's own weight, so
 * that  entropy[b] = logZ[b] - expect[b]  (semirings_test.py:305-324) costs two reads of the
 * weights and no write.  `alphas`, `dist` (and `alpha_norm`, or NULL) come from
 * lt_lattice_forward(_norm) on the same weights.  expect_part [B, V/32] doubles: the partial sums
 * of the V/32 CTAs of an utterance's cluster (sum them; fixed order, bit-reproducible).
 * Log semiring, lattices with lt_lattice_expectation_supported() == 1 (the bigram TMA fast path);
 * LT_ERR_UNSUPPORTED otherwise (compose lt_lattice_backward's posteriors with the values). */
int lt_lattice_expectation_supported(int vocab_size, int context_size, int max_expansions,
                                     unsigned flags);
int lt_lattice_expectation(int vocab_size, int context_size, int max_expansions,
                           const float* blank, const float* lexical, const int32_t* num_frames,
                           int B, int T, const float* alphas, const float* dist,
                           const int32_t* alpha_norm, const float* value_blank,
                           const float* value_lexical, double* expect_part, unsigned flags,
                           void* stream);

/* alphas[b,t,c] += alpha_norm[b,t] * (unit) in place (the alphas RecognitionLattice._forward
 * returns, lattices.py:496). */
int lt_alphas_denormalize(float* alphas, const int32_t* alpha_norm, int B, int T, int C,
                          void* stream);

/* ---- K5: Viterbi back-trace (replaces the vjp trick of lattices.py:221-247) ----
 *   alpha_final [B,C] from the MaxTropical forward.
 *   labels      [B,T,max(k,0)+1] int32: TRUE 1-based lexical labels, 0 = blank
 *               / not taken (the reference reports y-1)
 *   path_states [B,T+1] int32 or NULL: context state before each frame
 *   grad_blank / grad_lexical: NULL, or PRE-ZEROED dense buffers that receive
 *               grad_dist[b] (or 1 if grad_dist is NULL) on the arcs of the path
 *               (semirings.py:366-369, :389-398).
 */
int lt_viterbi_backtrace(int vocab_size, int context_size, int max_expansions,
                         const int16_t* backptr, const uint8_t* termptr,
                         const float* alpha_final, const int32_t* num_frames,
                         int B, int T, int32_t* labels, int32_t* path_states,
                         const float* grad_dist, float* grad_blank,
                         float* grad_lexical, void* stream);

/* Scheduling aid for callers that run the numerator kernels on a side stream beside K1: one
 * thread that sleeps for `nanoseconds` (<= 1 ms) on `stream`.  K1's CTAs fill the register file
 * of an SM exactly (two per SM on 128 of the 148 SMs) and its clusters of 8 must be co-resident;
 * 33 clusters fit on a B200.  When K1 and the small numerator kernels become launchable at the
 * same moment (both wait for the kernel that produced the weights) and numerator CTAs are placed
 * first, each one takes an SM away from K1, a whole cluster -- an utterance -- cannot be placed
 * and starts only when the numerator retires: K1 then lasts up to twice as long (measured 2.25
 * and 2.73 ms against 1.5 ms).  Delaying the side stream by a few tens of microseconds lets the
 * block scheduler place K1 first; the numerator then runs on the SMs that are left. */
int lt_stream_delay(unsigned nanoseconds, void* stream);

/* ---- K3: numerator on the T x (U+1) label lattice (lattices.py:250-377) ----
 * gather: weight_step_scan + gather_weight (lattices.py:300-342, :830-845)
 *   states [B,U1] int32 = walk_states(labels) (contexts.py:109-146)
 *   next_labels [B,U1] int32 in [1,V] (labels ++ [1], label 0 read as 1)
 *   -> blank_w, lexical_w [B,T,U1]
 * scatter_add: its transpose,
 *   grad_dense[b,t,states[u],(label-1)] += scale * utt_scale[b] * g
 *   (utt_scale [B] or NULL = 1: lets the numerator posteriors be computed once,
 *   unscaled, and weighted by the upstream gradient of each utterance here)
 */
/* walk_states (contexts.py:109-146, FullNGram.next_state :190-205):
 *   labels [B,U] int32 in [0,V] -> states [B,U+1], next_labels [B,U+1] */
int lt_walk_states(int vocab_size, int context_size, const int32_t* labels,
                   int B, int U, int32_t* states, int32_t* next_labels,
                   void* stream);
/* The same with the padding / range rules a caller needs for untrusted label tensors:
 *   num_labels [B] or NULL: positions u >= num_labels[b] are padding and read as epsilon
 *              whatever they hold (-1, V+1, ...) -- they cannot influence the numerator;
 *   bad_labels [1] or NULL: += number of labels outside [0, V] BEFORE num_labels (the reference
 *              fails on those in one_hot, lattices.py:322); they are read as epsilon too, so the
 *              emitted states / next_labels are always in range (the gather / scatter kernels
 *              additionally clamp what they are given).  The caller pre-zeroes the counter. */
int lt_walk_states_checked(int vocab_size, int context_size, const int32_t* labels,
                           const int32_t* num_labels, int B, int U, int32_t* states,
                           int32_t* next_labels, int32_t* bad_labels, void* stream);
int lt_string_gather(int vocab_size, int num_states, const float* blank,
                     const float* lexical, const int32_t* states,
                     const int32_t* next_labels, int B, int T, int U1,
                     float* blank_w, float* lexical_w, void* stream);
int lt_string_scatter_add(int vocab_size, int num_states,
                          const float* grad_blank_w, const float* grad_lexical_w,
                          const int32_t* states, const int32_t* next_labels,
                          int B, int T, int U1, float scale,
                          const float* utt_scale, float* grad_blank,
                          float* grad_lexical, void* stream);
/* shortest_distance_step_scan (lattices.py:347-377) with
 * alignment.string_forward (alignments.py:327-329 / :427-432).
 *   dist [B] = alpha_T[num_labels[b]] (semiring zero if num_labels > U)
 *   alphas [B,T,U1] or NULL; backptr [B,T,U1] uint8 (MaxTropical) or NULL. */
int lt_string_forward(int semiring, int max_expansions, const float* blank_w,
                      const float* lexical_w, const int32_t* num_frames,
                      const int32_t* num_labels, int B, int T, int U1,
                      float* dist, float* alphas, uint8_t* backptr,
                      void* stream);
int lt_string_backward(int semiring, int max_expansions, const float* blank_w,
                       const float* lexical_w, const int32_t* num_frames,
                       const int32_t* num_labels, int B, int T, int U1,
                       const float* alphas, const uint8_t* backptr,
                       const float* dist, const float* grad_dist,
                       float* grad_blank_w, float* grad_lexical_w,
                       void* stream);

/* The same two with the Log chain carried as (integer part, fraction): every value is
 * e + f with e an exact int32 and f in [0, 1) (f = -inf for the semiring zero), so the sums the
 * chain rounds have magnitude O(1) whatever the numerator's magnitude (1e3 at T = 1000).
 *   alphas [B,T,U1] then holds the fractions, alpha_exp [B,T,U1] int32 the integer parts;
 *   dist_norm [B,2] int32 = (integer part, bits of the fp32 fraction) of dist[b].
 * Only when lt_string_norm_supported() returns 1 (Log: register kernels for FrameDependent with
 * U1 <= 1024, a double-precision chain for FrameLabelDependent and longer label strings); NULL
 * for both selects the plain fp32 kernels. */
int lt_string_norm_supported(int semiring, int max_expansions, int U1);
int lt_string_forward_norm(int semiring, int max_expansions, const float* blank_w,
                           const float* lexical_w, const int32_t* num_frames,
                           const int32_t* num_labels, int B, int T, int U1,
                           float* dist, float* alphas, uint8_t* backptr,
                           int32_t* alpha_exp, int32_t* dist_norm, void* stream);
int lt_string_backward_norm(int semiring, int max_expansions, const float* blank_w,
                            const float* lexical_w, const int32_t* num_frames,
                            const int32_t* num_labels, int B, int T, int U1,
                            const float* alphas, const uint8_t* backptr,
                            const float* dist, const float* grad_dist,
                            float* grad_blank_w, float* grad_lexical_w,
                            const int32_t* alpha_exp, const int32_t* dist_norm,
                            void* stream);

/* ---- the bias-free input projections of JointWeightFn (weight_fns.py:208-211) -----------
 * y[M,N] = x[M,K] . w[N,K]^T  (nn.Linear(K, N, bias=False)).  The input gradient is the same
 * call on (gy, w^T); the weight gradient gw[N,K] = gy[M,N]^T . x[M,K] is a two-pass reduction in
 * a fixed order (bit-reproducible) through a caller-provided workspace of
 * lt_linear_wgrad_workspace_bytes(M, K, N) bytes.  Large 64-aligned products run on tcgen05
 * tensor cores (bf16x3 operand split, ~2^-17 relative, fp32 accumulation in tensor memory; forward:
 * M >= 128, K % 64 == 0, N % 16 == 0; weight gradient: M >= 1024, N % 128 == 0, K % 64 == 0;
 * 16-byte aligned buffers), everything else on fp32 FMAs. */
int lt_linear_forward(const float* x, const float* w, float* y, int64_t M, int K, int N,
                      void* stream);
int64_t lt_linear_wgrad_workspace_bytes(int64_t M, int K, int N);
int lt_linear_wgrad(const float* gy, const float* x, float* gw, int64_t M, int K, int N,
                    void* workspace, void* stream);
/* 1 when both calls above run on tcgen05 tensor cores for this shape (bf16x3 operand split, fp32
 * accumulation; K % 64 == 0, N % 128 == 0) AND the product is large enough (M >= 4096) for that
 * to beat an fp32 FMA GEMM -- the host side then routes the projection here instead of to the
 * library sgemm behind nn.Linear (LT_LINEAR_SIMT=1 keeps the CUDA-core kernels). */
int lt_linear_tensor_core(int64_t M, int K, int N);

/* ---- semiring (+) on arbitrary tensors (semirings.py:202-220, :330-348) ----
 * plus: elementwise on n elements (inputs already broadcast & contiguous).
 * sum: a viewed as [outer, reduce, inner] -> out [outer, inner];
 *      argmax [outer, inner] int32 (MaxTropical, may be NULL otherwise).
 * backward kernels implement the "safe gradient" rules (semirings.py:222-241)
 * and the tie rules a>=b / first arg-max (semirings.py:363, :382).
 */
int lt_semiring_plus_forward(int semiring, const float* a, const float* b,
                             float* out, int64_t n, void* stream);
int lt_semiring_plus_backward(int semiring, const float* a, const float* b,
                              const float* grad_out, float* grad_a,
                              float* grad_b, int64_t n, void* stream);
int lt_semiring_sum_forward(int semiring, const float* a, int64_t outer,
                            int64_t reduce, int64_t inner, float* out,
                            int32_t* argmax, void* stream);
int lt_semiring_sum_backward(int semiring, const float* a, const float* out,
                             const int32_t* argmax, const float* grad_out,
                             int64_t outer, int64_t reduce, int64_t inner,
                             float* grad_a, void* stream);

/* ---- generic context DFA: contexts.NextStateTable (contexts.py:266-324) ------
 * The same recursions as lt_lattice_forward / _backward / lt_viterbi_backtrace for a
 * context dependency given as a table:
 *   table      [C,V] int32  next state of (p, y), y zero-based
 *   in_offsets [C+1], in_arcs [C*V] int32: CSR of INCOMING arcs -- the flat arc indices
 *              p*V+y grouped by destination state, ascending inside a group (built once
 *              by the host, e.g. a stable argsort of the table)
 *   backarc    [B,T,max(k,1),C] int32, MaxTropical only: winning incoming arc of every
 *              reduction (first arg-max in flat-arc order); for FrameDependent -1 means
 *              "blank arc won" (semirings.py:363)
 * forward_reduce / its gradient on arbitrary leading dims (w viewed as [outer, C, V]):
 *   out[o,q] = (+)_{p -y-> q} w[o,p,y] for EVERY semiring (the reference implements the
 *   Real semiring only); argarc [outer,C] int32 for MaxTropical.
 * FrameDependent lattices with V % 4 == 0 (and C <= 2048 forward) run on a CLUSTER of up to 8
 * CTAs per utterance (csrc/lattice_table2.cu: every CTA streams a slab of source rows with
 * bulk copies; partial values / new beta travel by st.async + mbarrier complete_tx, there is NO
 * cluster barrier in either loop); everything else runs one CTA per utterance.
 * lt_table_lattice_cluster() returns the cluster size the forward (backward != 0: the
 * backward) kernel of such a lattice would use with 16-byte aligned weights, 0 for the
 * one-CTA kernels.  Options (lt_set_option): LT_TABLE_CLUSTER=1|2|4|8 forces a size,
 * LT_TABLE_V1=1 the one-CTA kernels.
 */
int lt_table_lattice_cluster(int C, int V, int max_expansions, int backward);
int lt_table_lattice_forward(int semiring, int max_expansions, const int32_t* table,
                             const int32_t* in_offsets, const int32_t* in_arcs, int C, int V,
                             const float* blank, const float* lexical,
                             const int32_t* num_frames, int B, int T,
                             const float* alpha_init, float* dist, float* alphas,
                             float* alpha_final, float* levels, int32_t* backarc,
                             uint8_t* termptr, void* stream);
int lt_table_lattice_backward(int semiring, int max_expansions, const int32_t* table, int C,
                              int V, const float* blank, const float* lexical,
                              const int32_t* num_frames, int B, int T, const float* alphas,
                              const float* levels, const float* dist,
                              const float* grad_dist, float* grad_blank,
                              float* grad_lexical, void* stream);
int lt_table_viterbi_backtrace(int max_expansions, int C, int V, const int32_t* backarc,
                               const uint8_t* termptr, const float* alpha_final,
                               const int32_t* num_frames, int B, int T, int32_t* labels,
                               int32_t* path_states, const float* grad_dist,
                               float* grad_blank, float* grad_lexical, void* stream);
int lt_table_reduce_forward(int semiring, const float* w, const int32_t* in_offsets,
                            const int32_t* in_arcs, int64_t outer, int C, int V, float* out,
                            int32_t* argarc, void* stream);
int lt_table_reduce_backward(int semiring, const float* w, const float* out,
                             const int32_t* argarc, const float* grad_out,
                             const int32_t* table, int64_t outer, int C, int V,
                             float* grad_w, void* stream);

/* ---- local normalisation epilogues (weight_fns.py:99-136) ---------------------
 * hat_normalize / log_softmax_normalize on M rows of (blank [M], lexical [M,V]):
 *   LT_NORM_HAT          blank' = blank - softplus(blank), lex' = log_softmax(lex) - softplus(blank)
 *   LT_NORM_LOG_SOFTMAX  (blank', lex') = log_softmax(blank ++ lex)
 * backward: gradients w.r.t. the INPUTS from the cotangents of the outputs (the
 * normalisers are recomputed from the inputs).
 */
#define LT_NORM_HAT 0
#define LT_NORM_LOG_SOFTMAX 1
int lt_local_normalize_forward(int mode, const float* blank, const float* lexical,
                               int64_t M, int V, float* out_blank, float* out_lexical,
                               void* stream);
int lt_local_normalize_backward(int mode, const float* blank, const float* lexical,
                                const float* grad_out_blank,
                                const float* grad_out_lexical, int64_t M, int V,
                                float* grad_blank, float* grad_lexical, void* stream);

/* ---- K4: JointWeightFn (weight_fns.py:194-227) -------------------------------
 * joint = tanh(cache @ w_ctx^T + frames @ w_frame^T)   [N, C, H]
 * blank = joint @ w_blank + b_blank                     [N, C]
 * lexical = joint @ w_vocab^T + b_vocab                 [N, C, V]
 * N = number of frames (B*T for the whole-utterance call the lattice makes).
 * proj_ctx [C,H] and proj_frame [N,H] are the two small input projections
 * (computed by the caller; they are O(C*E*H + N*D*H) and not on the hot path).
 * The vocabulary projection runs on tcgen05 tensor cores with a 3-way bf16
 * split of both operands (fp32-accurate).
 */
/* Bytes of device scratch lt_joint_forward needs (bf16 hi/lo split of W_vocab and the
 * e^(2x) tables of the two projections, (C + N) * H floats). */
int64_t lt_joint_workspace_bytes(int64_t N, int C, int H, int V);
/* b_blank is a DEVICE scalar (the bias of Linear(H, 1), weight_fns.py:220), read by the kernel
 * epilogue like b_vocab: no host synchronisation on the way in. */
int lt_joint_forward(const float* proj_ctx, const float* proj_frame,
                     const float* w_blank, const float* b_blank,
                     const float* w_vocab, const float* b_vocab, int64_t N, int C,
                     int H, int V, float* blank, float* lexical, void* workspace,
                     void* stream);
/* Gradients w.r.t. the joint pre-activation, reduced to the two projections:
 *   grad_proj_ctx [C,H] (+= over N), grad_proj_frame [N,H] (+= over C),
 *   grad_w_blank [H], grad_b_blank [1], grad_w_vocab [V,H], grad_b_vocab [V].
 * All outputs must be pre-zeroed; they are accumulated with atomics. */
int lt_joint_backward(const float* proj_ctx, const float* proj_frame,
                      const float* w_blank, const float* w_vocab,
                      const float* grad_blank, const float* grad_lexical,
                      int64_t N, int C, int H, int V, float* grad_proj_ctx,
                      float* grad_proj_frame, float* grad_w_blank,
                      float* grad_b_blank, float* grad_w_vocab,
                      float* grad_b_vocab, void* workspace,
                      int grad_lexical_format, void* stream);
/* 1 if lt_lattice_backward can emit split rows for this lattice (TMA fast path). */
int lt_lattice_backward_split_supported(int semiring, int vocab_size, int context_size,
                                        int max_expansions, unsigned flags);
/* lt_string_scatter_add for a split-row grad_lexical (grad_blank stays fp32). */
int lt_string_scatter_add_split(int vocab_size, int num_states, const float* grad_blank_w,
                                const float* grad_lexical_w, const int32_t* states,
                                const int32_t* next_labels, int B, int T, int U1, float scale,
                                const float* utt_scale, float* grad_blank, float* grad_lexical,
                                void* stream);
/* grad_lexical_format: 0 = fp32 [N, C, V]; 1 = "split rows": every row of V floats is replaced,
 * in the same V*4 bytes, by [V bf16 hi | V bf16 lo] with hi + lo = the value to 2^-17 -- the
 * operand form of the tensor-core kernels, which lt_lattice_backward can emit directly
 * (LT_FLAG_GRAD_SPLIT), so the fused dgrad loads it with TMA and converts nothing.  Only when
 * lt_joint_backward_split_supported() returns 1. */
int lt_joint_backward_split_supported(int64_t N, int C, int H, int V);
/* fp32 rows [M, V] -> split rows (same bytes), for callers that hold fp32 gradients. */
int lt_joint_split_rows(const float* rows, void* out, int64_t M, int V, void* stream);
/* Bytes of device scratch for lt_joint_backward (W_vocab^T as bf16 hi/lo + the
 * [N*C, H] pre-activation gradient that is reduced into the two projections);
 * workspace may be NULL, which selects the CUDA-core kernels. */
int64_t lt_joint_backward_workspace_bytes(int64_t N, int C, int H, int V);

/* ---- north_star (4): JointWeightFn FUSED into the forward recursion ----------------------------
 * lt_joint_forward followed by lt_lattice_forward in ONE kernel: the logits of a frame are produced
 * on chip (fp32 FMAs on the CUDA cores), consumed by the semiring update and never written to HBM;
 * device memory is O(B*T*C) instead of O(B*T*C*V).  Inference direction only (no gradient): Log
 * shortest distance, and MaxTropical distance + back-pointers (lt_viterbi_backtrace with
 * context_size 1, LT_FRAME_DEPENDENT) -- what RecognitionLattice.shortest_path needs.
 * Bigram FullNGram, FrameDependent, vocab_size <= 64, H in {32, 64, 128}
 * (lt_joint_lattice_fused_supported); proj_frame is [B, T, H].  Measured against the unfused pair
 * in DESIGN.md section 6: the recompute on CUDA cores makes it SLOWER than materialising the
 * logits with the tensor-core kernel; it is the memory-saving path and the measured answer to
 * "does fusion pay at the shape where recompute is cheapest". */
int lt_joint_lattice_fused_supported(int semiring, int vocab_size, int context_size,
                                     int max_expansions, int H);
int lt_joint_lattice_forward_fused(int semiring, int vocab_size, const float* proj_ctx,
                                   const float* proj_frame, const float* w_blank,
                                   const float* b_blank, const float* w_vocab,
                                   const float* b_vocab, const int32_t* num_frames, int B,
                                   int T, int H, float* dist, float* alphas,
                                   float* alpha_final, int16_t* backptr, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* LAST_LATTICE_H_ */
