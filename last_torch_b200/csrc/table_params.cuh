// Parameters of the table-driven lattice kernels (contexts.NextStateTable,
// /root/reference/last_torch/contexts.py:266-324): lattice_table.cu (one CTA per utterance,
// every shape) and lattice_table2.cu (a cluster per utterance, FrameDependent).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lt {

struct TableParams {
  int C, V, k, B, T;
  const int32_t* table;        // [C, V] next state of (p, y)
  const int32_t* in_offsets;   // [C + 1]
  const int32_t* in_arcs;      // [C * V]
  const float* blank;
  const float* lexical;
  const int32_t* num_frames;
  const float* alpha_init;
  float* dist;
  float* alphas;
  float* alpha_final;
  float* levels;
  int32_t* backarc;            // [B, T, max(k,1), C]
  uint8_t* termptr;            // [B, T, C]
  // backward
  const float* alphas_in;
  const float* levels_in;
  const float* dist_in;
  const float* grad_dist;
  float* grad_blank;
  float* grad_lexical;
};

// lattice_table2.cu: cluster kernels; *_supported() decides, the launchers return LT_* codes.
bool table2_forward_supported(const TableParams& p);
bool table2_backward_supported(const TableParams& p);
int table2_cluster_size(int C, int V, int k, bool backward);
int table2_forward_launch(int semiring, const TableParams& p, cudaStream_t stream);
int table2_backward_launch(int semiring, const TableParams& p, cudaStream_t stream);

}  // namespace lt
