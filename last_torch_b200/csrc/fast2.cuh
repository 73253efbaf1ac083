// Pieces shared by the bigram fast-path kernels (lattice_fast2.cu: FrameDependent,
// lattice_fast2_fld.cu: FrameLabelDependent): CTA geometry, utterance order, parameter blocks,
// gradient stores and the cluster launcher.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "fast_ptx.cuh"
#include "params.cuh"
#include "umma.cuh"

namespace lt {
namespace {

using namespace fastptx;

constexpr int kGroupThreads = 256;
constexpr int kGroupWarps = kGroupThreads / 32;
constexpr int kCols = 32;                       // destination columns (fwd) / source rows (bwd) per CTA

// Which utterance a cluster works on.  The block scheduler hands out clusters in blockIdx order
// as slots free up -- a greedy work queue -- so with more utterances than co-resident clusters
// (33 on a B200) the ORDER decides how well a ragged batch packs: longest first (LPT) keeps the
// tail short.  Every CTA ranks the utterances by length itself (rank(i) = number of utterances
// that are longer, or as long with a smaller index: O(B^2 / 256) compares per thread, a few
// microseconds once per kernel) and takes the one whose rank equals its cluster index; all CTAs
// of a cluster see the same num_frames and agree without communicating.
__device__ __forceinline__ int utterance_of_cluster(int cluster_id, const int32_t* num_frames,
                                                    int B, int T, int* slot) {
  if (B <= 32 || B > 4096) return cluster_id;          // one wave / ranking not worth its cost
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const int ni = max(0, min(num_frames[i], T));
    int rank = 0;
    for (int j = 0; j < B; ++j) {
      const int nj = max(0, min(num_frames[j], T));
      rank += (nj > ni || (nj == ni && j < i)) ? 1 : 0;
    }
    if (rank == cluster_id) *slot = i;
  }
  __syncthreads();
  return *slot;
}

__device__ __forceinline__ void group_sync(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(kGroupThreads) : "memory");
}

struct Fast2FwdParams {
  int B, T, stages;
  const float* blank;
  const float* lexical;
  const int32_t* num_frames;
  const float* alpha_init;
  float* dist;
  float* alphas;
  float* alpha_final;
  int16_t* backptr;
  int32_t* alpha_norm;   // NORM only: [B, T+3] = off_0 .. off_T (log2 units), bits of r, unit 0
  // FrameLabelDependent only (lattice_fast2_fld.cu)
  float* levels;         // [B, T, k, C] intermediate vectors last_1 .. last_k, or nullptr
  uint8_t* termptr;      // MaxTropical: [B, T, C] number of expansions of the best path into (t+1, q)
};

struct Fast2BwdParams {
  int B, T, stages;
  const float* blank;
  const float* lexical;
  const int32_t* num_frames;
  const float* alphas;
  const float* dist;
  const float* grad_dist;
  float* grad_blank;
  float* grad_lexical;
  float* beta_final;
  int split;             // LT_FLAG_GRAD_SPLIT: rows of [V bf16 hi | V bf16 lo] instead of fp32
  const int32_t* alpha_norm;   // NORM only: written by the NORM forward kernel
  const float* levels;   // FrameLabelDependent only: [B, T, k, C]
  // EXPECT (lt_lattice_expectation): no gradients are written; every CTA sums
  // posterior(arc) * value(arc) over its arcs into expect_part[b * CL + rank]
  const float* value_blank;     // [B, T, C] or nullptr: the arc's own weight
  const float* value_lexical;   // [B, T, C, V] or nullptr
  double* expect_part;
};

// Four consecutive gradients of one row.  fp32: one 16-byte store.  Split rows: the same 16
// bytes as two 8-byte stores -- 4 bf16 "hi" at element offset c4 of the row's first half, the 4
// bf16 residuals "lo" at the same offset of its second half (hi + lo = value to 2^-17).  It is
// the operand form of the tensor-core joint backward (joint_dgrad2.cu loads it by TMA).
__device__ __forceinline__ void store_grad4(float* row, int c4, int V, bool split, float4 v) {
  if (!split) {
    stg_stream4(row + c4, v);
  } else {
    uint32_t h0, l0, h1, l1;
    umma::split_pack2(v.x, v.y, h0, l0);
    umma::split_pack2(v.z, v.w, h1, l1);
    unsigned char* r = reinterpret_cast<unsigned char*>(row);
    asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" ::"l"(r + c4 * 2), "r"(h0),
                 "r"(h1) : "memory");
    asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" ::"l"(r + V * 2 + c4 * 2),
                 "r"(l0), "r"(l1) : "memory");
  }
}

// ------------------------------------------------------------------- host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn2() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
          cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

template <typename KernelT, typename... Args>
static int launch_fast2(KernelT kernel, int grid, int threads, size_t smem, int cluster,
                        cudaStream_t stream, Args... args) {
  LT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cluster > 8)
    LT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LT_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  note_launch();
  return LT_OK;
}

// One 256-thread CTA per (utterance, 32-column slice), at most 113 KB of shared memory so that
// TWO CTAs share an SM (33 clusters of 8 are co-resident on a B200).
constexpr int kSharedBudget = 113 * 1024;


}  // namespace
}  // namespace lt
