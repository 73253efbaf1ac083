// extern "C" entry points declared in include/last_lattice.h: argument
// validation, geometry, kernel-path dispatch and error reporting.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "params.cuh"

namespace lt {

static thread_local char g_error[512] = "";
static unsigned long long g_launches = 0;   // atomically bumped; read by lt_launch_count

void note_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return LT_ERR_CUDA;
}

static const char* const kOptionNames[OPT_COUNT] = {
    "LT_JOINT_SIMT", "LT_JOINT_DGRAD_V1", "LT_JOINT_WGRAD_SIMT", "LT_JOINT_DGRAD_PAIR",
    "LT_JOINT_DGRAD_MULTICAST", "LT_TABLE_V1", "LT_TABLE_CLUSTER", "LT_JOINT_FWD_SS",
    "LT_JOINT_FWD_CLUSTER", "LT_FLD_GENERIC", "LT_LINEAR_SIMT"};
static int g_options[OPT_COUNT];
static std::once_flag g_options_once;

static void init_options() {
  for (int i = 0; i < OPT_COUNT; ++i) {
    const char* e = getenv(kOptionNames[i]);
    int v = 0;
    if (e && *e) { v = atoi(e); if (v == 0 && e[0] != '0') v = 1; }
    __atomic_store_n(&g_options[i], v, __ATOMIC_RELAXED);
  }
}

int option(Option o) {
  std::call_once(g_options_once, init_options);
  return __atomic_load_n(&g_options[o], __ATOMIC_RELAXED);
}

static int find_option(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < OPT_COUNT; ++i)
    if (strcmp(name, kOptionNames[i]) == 0) return i;
  return -1;
}

static int device_sm_count(int* out) {
  int dev = 0;
  LT_CUDA(cudaGetDevice(&dev));
  LT_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return LT_OK;
}

static int check_arch() {
  int dev = 0, major = 0;
  LT_CUDA(cudaGetDevice(&dev));
  LT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("last_lattice kernels are built for sm_100a (B200) only; current device has "
              "compute capability major %d", major);
    return LT_ERR_UNSUPPORTED;
  }
  return LT_OK;
}

}  // namespace lt

using namespace lt;

extern "C" {

int lt_version(void) { return 100; }

const char* lt_last_error(void) { return g_error; }

unsigned long long lt_launch_count(void) { return launch_count(); }

int lt_set_option(const char* name, int value) {
  const int i = find_option(name);
  LT_CHECK_ARG(i >= 0, "lt_set_option: unknown option '%s'", name ? name : "(null)");
  std::call_once(g_options_once, init_options);
  __atomic_store_n(&g_options[i], value, __ATOMIC_RELAXED);
  return LT_OK;
}

int lt_get_option(const char* name) {
  const int i = find_option(name);
  return i < 0 ? -1 : option((Option)i);
}

int lt_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  LT_CUDA(cudaGetDevice(&dev));
  if (sm_count) LT_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) LT_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) LT_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return LT_OK;
}

static int check_common(const char* fn, int semiring, int V, int n, int k, int B, int T, NGram* g) {
  LT_CHECK_ARG(semiring == LT_REAL || semiring == LT_LOG || semiring == LT_MAXTROPICAL,
               "%s: unknown semiring %d", fn, semiring);
  LT_CHECK_ARG(V > 0, "%s: vocab_size should be > 0, but got vocab_size=%d", fn, V);
  LT_CHECK_ARG(n >= 0, "%s: context_size should be >= 0, but got context_size=%d", fn, n);
  LT_CHECK_ARG(k == LT_FRAME_DEPENDENT || (k >= 1 && k <= 254),
               "%s: max_expansions must be LT_FRAME_DEPENDENT or in [1, 254], got %d", fn, k);
  LT_CHECK_ARG(B >= 0 && T >= 0, "%s: negative batch (%d) or frame count (%d)", fn, B, T);
  LT_CHECK_ARG(make_ngram(V, n, g), "%s: FullNGram(vocab_size=%d, context_size=%d) has too many states",
               fn, V, n);
  LT_CHECK_ARG(V + 1 <= 32767, "%s: vocab_size %d exceeds the int16 back-pointer range", fn, V);
  return LT_OK;
}

int lt_lattice_forward(int semiring, int vocab_size, int context_size, int max_expansions,
                       const float* blank, const float* lexical, const int32_t* num_frames, int B,
                       int T, const float* alpha_init, float* dist, float* alphas,
                       float* alpha_final, float* levels, int16_t* backptr, uint8_t* termptr,
                       unsigned flags, void* stream) {
  return lt_lattice_forward_norm(semiring, vocab_size, context_size, max_expansions, blank, lexical,
                                 num_frames, B, T, alpha_init, dist, alphas, alpha_final, levels,
                                 backptr, termptr, nullptr, flags, stream);
}

}  // extern "C"

namespace lt {
int lattice_norm_family(int semiring, const NGram& g, int k, unsigned flags, const void* lexical) {
  if (semiring != LT_LOG) return 0;
  if (lattice_fast2_supported(g, k, flags, lexical)) return 1;
  if (lattice_fast2_fld_supported(g, k, flags, lexical)) return 1;
  // context_size >= 2: the thread-per-column forward and the 8-lanes-per-row backward share the
  // fast path's convention (log2 units); a forward without its backward stays plain
  const void* probe = lexical ? lexical : reinterpret_cast<const void*>(uintptr_t(256));
  if (lattice_cols_supported(g, k, flags, probe))
    return lattice_rows_supported(g, k, flags, probe, probe) ? 1 : 0;
  return 2;
}
}  // namespace lt

extern "C" {

int lt_lattice_norm_supported(int semiring, int vocab_size, int context_size, int max_expansions,
                              unsigned flags) {
  NGram g;
  if (!make_ngram(vocab_size, context_size, &g)) return 0;
  return lattice_norm_family(semiring, g, max_expansions, flags, nullptr) ? 1 : 0;
}

int lt_lattice_forward_norm(int semiring, int vocab_size, int context_size, int max_expansions,
                            const float* blank, const float* lexical, const int32_t* num_frames,
                            int B, int T, const float* alpha_init, float* dist, float* alphas,
                            float* alpha_final, float* levels, int16_t* backptr, uint8_t* termptr,
                            int32_t* alpha_norm, unsigned flags, void* stream) {
  NGram g;
  int rc = check_common("lt_lattice_forward", semiring, vocab_size, context_size, max_expansions, B, T, &g);
  if (rc) return rc;
  LT_CHECK_ARG(dist && num_frames, "lt_lattice_forward: dist and num_frames must not be NULL");
  LT_CHECK_ARG(T == 0 || (blank && lexical), "lt_lattice_forward: blank/lexical must not be NULL");
  if (B == 0) return LT_OK;
  if ((rc = check_arch())) return rc;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  FwdParams p = {};
  p.g = g; p.k = max_expansions; p.B = B; p.T = T;
  p.blank = blank; p.lexical = lexical; p.num_frames = num_frames; p.alpha_init = alpha_init;
  p.dist = dist; p.alphas = alphas; p.alpha_final = alpha_final;
  p.levels = max_expansions >= 1 ? levels : nullptr;
  p.backptr = semiring == LT_MAXTROPICAL ? backptr : nullptr;
  p.termptr = (semiring == LT_MAXTROPICAL && max_expansions >= 1) ? termptr : nullptr;
  p.alpha_norm = alpha_norm;
  p.wlevels = 1;
  if (flags & LT_FLAG_LEVEL_WEIGHTS) {
    LT_CHECK_ARG(max_expansions >= 1,
                 "lt_lattice_forward: LT_FLAG_LEVEL_WEIGHTS needs FrameLabelDependent");
    p.wlevels = max_expansions + 1;
    flags |= LT_FLAG_FORCE_GENERIC;
  }
  const bool fast2 = T > 0 && lattice_fast2_supported(g, max_expansions, flags, lexical);
  const bool fld2 = T > 0 && lattice_fast2_fld_supported(g, max_expansions, flags, lexical);
  if (alpha_norm && (T == 0 || !lattice_norm_family(semiring, g, max_expansions, flags, lexical))) {
    set_error("lt_lattice_forward_norm: this lattice has no renormalised kernel "
              "(lt_lattice_norm_supported, T > 0)");
    return LT_ERR_UNSUPPORTED;
  }
  if (fast2) return lattice_forward_fast2_launch(semiring, g, p, flags, (cudaStream_t)stream);
  if (fld2) return lattice_forward_fld2_launch(semiring, g, p, flags, (cudaStream_t)stream);
  if (T > 0 && lattice_cols_supported(g, max_expansions, flags, lexical))
    return lattice_forward_cols_launch(semiring, g, max_expansions, p, (cudaStream_t)stream);
  return lattice_forward_generic_launch(semiring, g, max_expansions, p, flags, sms,
                                        (cudaStream_t)stream);
}

int lt_lattice_backward(int semiring, int vocab_size, int context_size, int max_expansions,
                        const float* blank, const float* lexical, const int32_t* num_frames, int B,
                        int T, const float* alphas, const float* levels, const float* dist,
                        const float* grad_dist, float* grad_blank, float* grad_lexical,
                        float* beta_final, unsigned flags, void* stream) {
  return lt_lattice_backward_norm(semiring, vocab_size, context_size, max_expansions, blank,
                                  lexical, num_frames, B, T, alphas, levels, dist, grad_dist,
                                  grad_blank, grad_lexical, beta_final, nullptr, flags, stream);
}

int lt_lattice_backward_norm(int semiring, int vocab_size, int context_size, int max_expansions,
                             const float* blank, const float* lexical, const int32_t* num_frames,
                             int B, int T, const float* alphas, const float* levels,
                             const float* dist, const float* grad_dist, float* grad_blank,
                             float* grad_lexical, float* beta_final, const int32_t* alpha_norm,
                             unsigned flags, void* stream) {
  NGram g;
  int rc = check_common("lt_lattice_backward", semiring, vocab_size, context_size, max_expansions, B, T, &g);
  if (rc) return rc;
  LT_CHECK_ARG(semiring != LT_MAXTROPICAL,
               "lt_lattice_backward: MaxTropical gradients come from lt_viterbi_backtrace");
  LT_CHECK_ARG(num_frames && dist, "lt_lattice_backward: num_frames and dist must not be NULL");
  LT_CHECK_ARG(T == 0 || (blank && lexical && alphas && grad_blank && grad_lexical),
               "lt_lattice_backward: blank/lexical/alphas/grad buffers must not be NULL");
  LT_CHECK_ARG(max_expansions < 1 || levels || T == 0,
               "lt_lattice_backward: FrameLabelDependent needs the `levels` buffer of the forward");
  if (B == 0 || T == 0) return LT_OK;
  if ((rc = check_arch())) return rc;
  int sms = 0;
  if ((rc = device_sm_count(&sms))) return rc;
  BwdParams p = {};
  p.g = g; p.k = max_expansions; p.B = B; p.T = T;
  p.blank = blank; p.lexical = lexical; p.num_frames = num_frames;
  p.alphas = alphas; p.levels = levels; p.dist = dist; p.grad_dist = grad_dist;
  p.grad_blank = grad_blank; p.grad_lexical = grad_lexical; p.beta_final = beta_final;
  p.alpha_norm = alpha_norm;
  p.wlevels = 1;
  if (flags & LT_FLAG_LEVEL_WEIGHTS) {
    LT_CHECK_ARG(max_expansions >= 1,
                 "lt_lattice_backward: LT_FLAG_LEVEL_WEIGHTS needs FrameLabelDependent");
    p.wlevels = max_expansions + 1;
    flags |= LT_FLAG_FORCE_GENERIC;
  }
  const bool fld2 = lattice_fast2_fld_supported(g, max_expansions, flags, lexical) &&
                    reinterpret_cast<uintptr_t>(grad_lexical) % 16 == 0;
  const bool fast2 = fld2 || (lattice_fast2_supported(g, max_expansions, flags, lexical) &&
                              reinterpret_cast<uintptr_t>(grad_lexical) % 16 == 0);
  if (alpha_norm) {
    // the pair must stay inside one kernel family: the offsets' unit differs between them
    const int family = lattice_norm_family(semiring, g, max_expansions, flags, lexical);
    const bool rows = !fast2 && !(flags & LT_FLAG_GRAD_SPLIT) &&
                      lattice_rows_supported(g, max_expansions, flags, lexical, grad_lexical);
    if (family == 0 || (family == 1 && !fast2 && !rows)) {
      set_error("lt_lattice_backward_norm: `alphas` / `alpha_norm` are renormalised but this "
                "call cannot take the kernel family that wrote them (16-byte aligned gradients?)");
      return LT_ERR_UNSUPPORTED;
    }
    if (family == 2) {
      if (flags & LT_FLAG_GRAD_SPLIT) {
        set_error("lt_lattice_backward: LT_FLAG_GRAD_SPLIT needs the TMA fast path");
        return LT_ERR_UNSUPPORTED;
      }
      return lattice_backward_generic_launch(semiring, g, max_expansions, p, flags, sms,
                                             (cudaStream_t)stream);
    }
  }
  if (fld2) return lattice_backward_fld2_launch(semiring, g, p, flags, (cudaStream_t)stream);
  if (fast2) return lattice_backward_fast2_launch(semiring, g, p, flags, (cudaStream_t)stream);
  if (flags & LT_FLAG_GRAD_SPLIT) {
    set_error("lt_lattice_backward: LT_FLAG_GRAD_SPLIT needs the TMA fast path "
              "(lt_lattice_backward_split_supported)");
    return LT_ERR_UNSUPPORTED;
  }
  if (lattice_rows_supported(g, max_expansions, flags, lexical, grad_lexical))
    return lattice_backward_rows_launch(semiring, g, p, sms, (cudaStream_t)stream);
  return lattice_backward_generic_launch(semiring, g, max_expansions, p, flags, sms,
                                         (cudaStream_t)stream);
}

int lt_lattice_expectation_supported(int vocab_size, int context_size, int max_expansions,
                                     unsigned flags) {
  NGram g;
  if (vocab_size < 1 || context_size < 0 || !make_ngram(vocab_size, context_size, &g)) return 0;
  return lattice_fast2_supported(g, max_expansions, flags,
                                 reinterpret_cast<const void*>(uintptr_t(256))) ? 1 : 0;
}

int lt_lattice_expectation(int vocab_size, int context_size, int max_expansions,
                           const float* blank, const float* lexical, const int32_t* num_frames,
                           int B, int T, const float* alphas, const float* dist,
                           const int32_t* alpha_norm, const float* value_blank,
                           const float* value_lexical, double* expect_part, unsigned flags,
                           void* stream) {
  NGram g;
  int rc = check_common("lt_lattice_expectation", LT_LOG, vocab_size, context_size, max_expansions,
                        B, T, &g);
  if (rc) return rc;
  LT_CHECK_ARG(num_frames && dist && expect_part, "lt_lattice_expectation: NULL pointer");
  LT_CHECK_ARG((value_blank == nullptr) == (value_lexical == nullptr),
               "lt_lattice_expectation: value_blank and value_lexical go together");
  if (B == 0) return LT_OK;
  if ((rc = check_arch())) return rc;
  const int cl = vocab_size / 32;
  LT_CUDA(cudaMemsetAsync(expect_part, 0, sizeof(double) * (size_t)B * (cl > 0 ? cl : 1),
                          (cudaStream_t)stream));
  if (T == 0) return LT_OK;
  LT_CHECK_ARG(blank && lexical && alphas, "lt_lattice_expectation: NULL weights / alphas");
  if (!lattice_fast2_supported(g, max_expansions, flags, lexical) ||
      (value_lexical && reinterpret_cast<uintptr_t>(value_lexical) % 16 != 0)) {
    set_error("lt_lattice_expectation: no fused kernel for this lattice "
              "(lt_lattice_expectation_supported); use the arc posteriors of lt_lattice_backward");
    return LT_ERR_UNSUPPORTED;
  }
  BwdParams p = {};
  p.g = g; p.k = max_expansions; p.B = B; p.T = T;
  p.blank = blank; p.lexical = lexical; p.num_frames = num_frames;
  p.alphas = alphas; p.dist = dist; p.alpha_norm = alpha_norm; p.wlevels = 1;
  p.value_blank = value_blank; p.value_lexical = value_lexical; p.expect_part = expect_part;
  return lattice_backward_fast2_launch(LT_LOG, g, p, flags & ~LT_FLAG_GRAD_SPLIT,
                                       (cudaStream_t)stream);
}

int lt_alphas_denormalize(float* alphas, const int32_t* alpha_norm, int B, int T, int C,
                          void* stream) {
  LT_CHECK_ARG(B >= 0 && T >= 0 && C >= 0, "lt_alphas_denormalize: bad sizes B=%d T=%d C=%d", B, T, C);
  if (B == 0 || T == 0 || C == 0) return LT_OK;
  LT_CHECK_ARG(alphas && alpha_norm, "lt_alphas_denormalize: NULL pointer");
  return alphas_denormalize_launch(alphas, alpha_norm, B, T, C, (cudaStream_t)stream);
}

int lt_viterbi_backtrace(int vocab_size, int context_size, int max_expansions,
                         const int16_t* backptr, const uint8_t* termptr, const float* alpha_final,
                         const int32_t* num_frames, int B, int T, int32_t* labels,
                         int32_t* path_states, const float* grad_dist, float* grad_blank,
                         float* grad_lexical, void* stream) {
  NGram g;
  int rc = check_common("lt_viterbi_backtrace", LT_MAXTROPICAL, vocab_size, context_size,
                        max_expansions, B, T, &g);
  if (rc) return rc;
  LT_CHECK_ARG(alpha_final && num_frames && labels,
               "lt_viterbi_backtrace: alpha_final, num_frames and labels must not be NULL");
  LT_CHECK_ARG(T == 0 || backptr, "lt_viterbi_backtrace: backptr must not be NULL");
  LT_CHECK_ARG(max_expansions < 1 || termptr || T == 0,
               "lt_viterbi_backtrace: FrameLabelDependent needs termptr");
  if (B == 0) return LT_OK;
  if ((rc = check_arch())) return rc;
  VitParams p = {};
  p.g = g; p.k = max_expansions; p.B = B; p.T = T;
  p.backptr = backptr; p.termptr = termptr; p.alpha_final = alpha_final;
  p.num_frames = num_frames; p.labels = labels; p.path_states = path_states;
  p.grad_dist = grad_dist; p.grad_blank = grad_blank; p.grad_lexical = grad_lexical;
  return viterbi_launch(p, (cudaStream_t)stream);
}

int lt_walk_states(int vocab_size, int context_size, const int32_t* labels, int B, int U,
                   int32_t* states, int32_t* next_labels, void* stream) {
  return lt_walk_states_checked(vocab_size, context_size, labels, nullptr, B, U, states,
                                next_labels, nullptr, stream);
}

int lt_walk_states_checked(int vocab_size, int context_size, const int32_t* labels,
                           const int32_t* num_labels, int B, int U, int32_t* states,
                           int32_t* next_labels, int32_t* bad_labels, void* stream) {
  NGram g;
  LT_CHECK_ARG(make_ngram(vocab_size, context_size, &g),
               "lt_walk_states: bad FullNGram(vocab_size=%d, context_size=%d)", vocab_size,
               context_size);
  LT_CHECK_ARG(B >= 0 && U >= 0, "lt_walk_states: bad sizes B=%d U=%d", B, U);
  if (B == 0) return LT_OK;
  LT_CHECK_ARG(states && next_labels && (U == 0 || labels), "lt_walk_states: NULL pointer");
  return walk_states_launch(g, labels, num_labels, B, U, states, next_labels, bad_labels,
                            (cudaStream_t)stream);
}

int lt_stream_delay(unsigned nanoseconds, void* stream) {
  return stream_delay_launch(nanoseconds, (cudaStream_t)stream);
}

int lt_string_gather(int vocab_size, int num_states, const float* blank, const float* lexical,
                     const int32_t* states, const int32_t* next_labels, int B, int T, int U1,
                     float* blank_w, float* lexical_w, void* stream) {
  LT_CHECK_ARG(vocab_size > 0 && num_states > 0 && B >= 0 && T >= 0 && U1 >= 1,
               "lt_string_gather: bad sizes V=%d C=%d B=%d T=%d U1=%d", vocab_size, num_states, B, T, U1);
  if (B == 0 || T == 0) return LT_OK;
  LT_CHECK_ARG(blank && lexical && states && next_labels && blank_w && lexical_w,
               "lt_string_gather: NULL pointer");
  return string_gather_launch(vocab_size, num_states, blank, lexical, states, next_labels, B, T,
                              U1, blank_w, lexical_w, (cudaStream_t)stream);
}

int lt_string_scatter_add(int vocab_size, int num_states, const float* grad_blank_w,
                          const float* grad_lexical_w, const int32_t* states,
                          const int32_t* next_labels, int B, int T, int U1, float scale,
                          const float* utt_scale, float* grad_blank, float* grad_lexical,
                          void* stream) {
  LT_CHECK_ARG(vocab_size > 0 && num_states > 0 && B >= 0 && T >= 0 && U1 >= 1,
               "lt_string_scatter_add: bad sizes V=%d C=%d B=%d T=%d U1=%d", vocab_size, num_states, B, T, U1);
  if (B == 0 || T == 0) return LT_OK;
  LT_CHECK_ARG(grad_blank_w && grad_lexical_w && states && next_labels && grad_blank && grad_lexical,
               "lt_string_scatter_add: NULL pointer");
  return string_scatter_launch(vocab_size, num_states, grad_blank_w, grad_lexical_w, states,
                               next_labels, B, T, U1, scale, utt_scale, grad_blank, grad_lexical,
                               0, (cudaStream_t)stream);
}

int lt_string_scatter_add_split(int vocab_size, int num_states, const float* grad_blank_w,
                                const float* grad_lexical_w, const int32_t* states,
                                const int32_t* next_labels, int B, int T, int U1, float scale,
                                const float* utt_scale, float* grad_blank, float* grad_lexical,
                                void* stream) {
  LT_CHECK_ARG(vocab_size > 0 && num_states > 0 && B >= 0 && T >= 0 && U1 >= 1 && U1 <= 4096,
               "lt_string_scatter_add_split: bad sizes V=%d C=%d B=%d T=%d U1=%d", vocab_size,
               num_states, B, T, U1);
  if (B == 0 || T == 0) return LT_OK;
  LT_CHECK_ARG(grad_blank_w && grad_lexical_w && states && next_labels && grad_blank && grad_lexical,
               "lt_string_scatter_add_split: NULL pointer");
  return string_scatter_launch(vocab_size, num_states, grad_blank_w, grad_lexical_w, states,
                               next_labels, B, T, U1, scale, utt_scale, grad_blank, grad_lexical,
                               1, (cudaStream_t)stream);
}

int lt_lattice_backward_split_supported(int semiring, int vocab_size, int context_size,
                                        int max_expansions, unsigned flags) {
  if (semiring != LT_LOG && semiring != LT_REAL) return 0;
  if (vocab_size < 1 || context_size < 0) return 0;
  NGram g;
  if (check_common("lt_lattice_backward_split_supported", semiring, vocab_size, context_size,
                   max_expansions, 1, 1, &g))
    return 0;
  const void* probe = reinterpret_cast<const void*>(uintptr_t(256));
  return (lattice_fast2_supported(g, max_expansions, flags & ~LT_FLAG_GRAD_SPLIT, probe) ||
          lattice_fast2_fld_supported(g, max_expansions, flags & ~LT_FLAG_GRAD_SPLIT, probe)) ? 1 : 0;
}

static int check_string(const char* fn, int semiring, int k, int B, int T, int U1) {
  LT_CHECK_ARG(semiring == LT_REAL || semiring == LT_LOG || semiring == LT_MAXTROPICAL,
               "%s: unknown semiring %d", fn, semiring);
  LT_CHECK_ARG(k == LT_FRAME_DEPENDENT || (k >= 1 && k <= 254),
               "%s: max_expansions must be LT_FRAME_DEPENDENT or in [1, 254], got %d", fn, k);
  LT_CHECK_ARG(B >= 0 && T >= 0 && U1 >= 1, "%s: bad sizes B=%d T=%d U1=%d", fn, B, T, U1);
  return LT_OK;
}

int lt_string_forward(int semiring, int max_expansions, const float* blank_w,
                      const float* lexical_w, const int32_t* num_frames,
                      const int32_t* num_labels, int B, int T, int U1, float* dist, float* alphas,
                      uint8_t* backptr, void* stream) {
  return lt_string_forward_norm(semiring, max_expansions, blank_w, lexical_w, num_frames,
                                num_labels, B, T, U1, dist, alphas, backptr, nullptr, nullptr,
                                stream);
}

int lt_string_norm_supported(int semiring, int max_expansions, int U1) {
  return string_norm_supported(semiring, max_expansions, U1) ? 1 : 0;
}

int lt_string_forward_norm(int semiring, int max_expansions, const float* blank_w,
                           const float* lexical_w, const int32_t* num_frames,
                           const int32_t* num_labels, int B, int T, int U1, float* dist,
                           float* alphas, uint8_t* backptr, int32_t* alpha_exp,
                           int32_t* dist_norm, void* stream) {
  int rc = check_string("lt_string_forward", semiring, max_expansions, B, T, U1);
  if (rc) return rc;
  LT_CHECK_ARG((alpha_exp == nullptr) == (dist_norm == nullptr),
               "lt_string_forward_norm: alpha_exp and dist_norm go together");
  if (alpha_exp) {
    LT_CHECK_ARG(alphas, "lt_string_forward_norm: alpha_exp needs the alphas buffer");
    if (!string_norm_supported(semiring, max_expansions, U1)) {
      set_error("lt_string_forward_norm: no (integer, fraction) kernel for this chain "
                "(lt_string_norm_supported)");
      return LT_ERR_UNSUPPORTED;
    }
  }
  LT_CHECK_ARG(num_frames && num_labels && dist, "lt_string_forward: NULL pointer");
  LT_CHECK_ARG(T == 0 || (blank_w && lexical_w), "lt_string_forward: NULL weights");
  StrParams p = {};
  p.k = max_expansions; p.B = B; p.T = T; p.U1 = U1;
  p.blank_w = blank_w; p.lexical_w = lexical_w; p.num_frames = num_frames;
  p.num_labels = num_labels; p.dist = dist; p.alphas = alphas;
  p.backptr = semiring == LT_MAXTROPICAL ? backptr : nullptr;
  p.alpha_exp = alpha_exp; p.dist_norm = dist_norm;
  return string_forward_launch(semiring, p, (cudaStream_t)stream);
}

int lt_string_backward(int semiring, int max_expansions, const float* blank_w,
                       const float* lexical_w, const int32_t* num_frames,
                       const int32_t* num_labels, int B, int T, int U1, const float* alphas,
                       const uint8_t* backptr, const float* dist, const float* grad_dist,
                       float* grad_blank_w, float* grad_lexical_w, void* stream) {
  return lt_string_backward_norm(semiring, max_expansions, blank_w, lexical_w, num_frames,
                                 num_labels, B, T, U1, alphas, backptr, dist, grad_dist,
                                 grad_blank_w, grad_lexical_w, nullptr, nullptr, stream);
}

int lt_string_backward_norm(int semiring, int max_expansions, const float* blank_w,
                            const float* lexical_w, const int32_t* num_frames,
                            const int32_t* num_labels, int B, int T, int U1, const float* alphas,
                            const uint8_t* backptr, const float* dist, const float* grad_dist,
                            float* grad_blank_w, float* grad_lexical_w, const int32_t* alpha_exp,
                            const int32_t* dist_norm, void* stream) {
  int rc = check_string("lt_string_backward", semiring, max_expansions, B, T, U1);
  if (rc) return rc;
  LT_CHECK_ARG((alpha_exp == nullptr) == (dist_norm == nullptr),
               "lt_string_backward_norm: alpha_exp and dist_norm go together");
  if (alpha_exp && !string_norm_supported(semiring, max_expansions, U1)) {
    set_error("lt_string_backward_norm: no (integer, fraction) kernel for this chain "
              "(lt_string_norm_supported)");
    return LT_ERR_UNSUPPORTED;
  }
  if (B == 0 || T == 0) return LT_OK;
  LT_CHECK_ARG(num_frames && num_labels && dist && grad_blank_w && grad_lexical_w && blank_w && lexical_w,
               "lt_string_backward: NULL pointer");
  LT_CHECK_ARG(semiring == LT_MAXTROPICAL ? backptr != nullptr : alphas != nullptr,
               "lt_string_backward: needs alphas (Real/Log) or backptr (MaxTropical) from the forward");
  StrParams p = {};
  p.k = max_expansions; p.B = B; p.T = T; p.U1 = U1;
  p.blank_w = blank_w; p.lexical_w = lexical_w; p.num_frames = num_frames;
  p.num_labels = num_labels; p.alphas_in = alphas; p.backptr_in = backptr; p.dist_in = dist;
  p.grad_dist = grad_dist; p.grad_blank_w = grad_blank_w; p.grad_lexical_w = grad_lexical_w;
  p.alpha_exp_in = alpha_exp; p.dist_norm_in = dist_norm;
  return string_backward_launch(semiring, p, (cudaStream_t)stream);
}

int lt_semiring_plus_forward(int semiring, const float* a, const float* b, float* out, int64_t n,
                             void* stream) {
  LT_CHECK_ARG(n >= 0 && (n == 0 || (a && b && out)), "lt_semiring_plus_forward: bad arguments");
  return semiring_plus_forward_launch(semiring, a, b, out, n, (cudaStream_t)stream);
}

int lt_semiring_plus_backward(int semiring, const float* a, const float* b, const float* grad_out,
                              float* grad_a, float* grad_b, int64_t n, void* stream) {
  LT_CHECK_ARG(n >= 0 && (n == 0 || (a && b && grad_out && grad_a && grad_b)),
               "lt_semiring_plus_backward: bad arguments");
  return semiring_plus_backward_launch(semiring, a, b, grad_out, grad_a, grad_b, n,
                                       (cudaStream_t)stream);
}

int lt_semiring_sum_forward(int semiring, const float* a, int64_t outer, int64_t reduce,
                            int64_t inner, float* out, int32_t* argmax, void* stream) {
  LT_CHECK_ARG(outer >= 0 && reduce >= 0 && inner >= 0, "lt_semiring_sum_forward: negative extent");
  return semiring_sum_forward_launch(semiring, a, outer, reduce, inner, out, argmax,
                                     (cudaStream_t)stream);
}

int lt_semiring_sum_backward(int semiring, const float* a, const float* out, const int32_t* argmax,
                             const float* grad_out, int64_t outer, int64_t reduce, int64_t inner,
                             float* grad_a, void* stream) {
  LT_CHECK_ARG(outer >= 0 && reduce >= 0 && inner >= 0, "lt_semiring_sum_backward: negative extent");
  return semiring_sum_backward_launch(semiring, a, out, argmax, grad_out, outer, reduce, inner,
                                      grad_a, (cudaStream_t)stream);
}

}  // extern "C"
