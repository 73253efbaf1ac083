// PTX wrappers shared by the TMA / cluster fast-path kernels (lattice_fast.cu,
// lattice_fast2.cu): mbarrier, bulk / tensor TMA loads, st.async DSMEM exchange and
// the log2-domain log-sum-exp helpers.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace lt {
namespace fastptx {

// ----------------------------------------------------------------- PTX helpers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LT_WAIT_LOOP%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LT_WAIT_DONE%=;\n"
      "bra LT_WAIT_LOOP%=;\n"
      "LT_WAIT_DONE%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// one MUFU.EX2 (max rel. error 2^-22; results below 2^-126 flush to +0)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// All-gather of the recursion state without a cluster barrier: a DSMEM store
// that also completes 4 bytes of the destination CTA's mbarrier transaction.
// The receiver arms the barrier with expect_tx(C * 4) once per frame and waits
// on it; no memory fence is involved (the mbarrier orders the data).
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::
                   "r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void xchg_store(float* base, int idx, float v, uint64_t* bar,
                                           uint32_t nrank) {
  const uint32_t a = smem_u32(base + idx), bb = smem_u32(bar);
  for (uint32_t r = 0; r < nrank; ++r)
    st_async_f32(map_shared_rank(a, r), v, map_shared_rank(bb, r));
}
// The same all-gather issued by the 8 lanes of a row group (lane `sl` of the group sends to rank
// `sl`): one shuffle and one store per lane instead of a loop of `nrank` stores on the owner lane.
// Every lane of the warp must call it (the shuffle is warp-wide); `v` is taken from the group's
// lane 0, `go` (uniform inside a group) masks groups that have nothing to send.
__device__ __forceinline__ void xchg_store_group8(float* base, int idx, float v, uint64_t* bar,
                                                  uint32_t nrank, int lane, bool go) {
  const float vv = __shfl_sync(0xffffffffu, v, lane & ~7);
  const uint32_t r = lane & 7;
  if (go && r < nrank)
    st_async_f32(map_shared_rank(smem_u32(base + idx), r), vv, map_shared_rank(smem_u32(bar), r));
}

// The Log-semiring fast kernels keep alpha / beta in LOG2 units on chip:
//   y = fma(w, log2(e), alpha2)  is one FFMA whose rounding error is the fp32
//   representation error of the sum itself (same as the reference's a + w), and
//   the max-shifted exponent y - m is then an exact difference fed to ex2.
// (m, s) pair of a running log2-sum-exp2: value = msafe(m) + log2(s).
__device__ __forceinline__ void lse2_merge(float& m, float& s, float om, float os) {
  const float mn = fmaxf(m, om);
  const float mns = msafe(mn);
  const float sa = (m == neg_inf()) ? 0.f : ex2(msafe(m) - mns);
  const float sb = (om == neg_inf()) ? 0.f : ex2(msafe(om) - mns);
  s = s * sa + os * sb;
  m = mn;
}
// log2(2^a + 2^b) with the non-finite-max rule of semirings.py:250-251
__device__ __forceinline__ float log2_add_exp2(float a, float b) {
  const float c = fmaxf(a, b);
  const float cs = msafe(c);
  return cs + __log2f(ex2(a - cs) + ex2(b - cs));
}
template <int SR> __device__ __forceinline__ float to_dom(float x) {
  return SR == LT_LOG ? x * kLog2e : x;
}
template <int SR> __device__ __forceinline__ float from_dom(float x) {
  return SR == LT_LOG ? x * kLn2 : x;
}
// weight (x) destination value: Log works in log2 units (w * log2e + v), Real is w * v
template <int SR> __device__ __forceinline__ float arc(float w, float v) {
  return SR == LT_LOG ? fmaf(w, kLog2e, v) : w * v;
}

__device__ __forceinline__ void bcast_f32(float* base, int idx, float v, uint32_t nrank) {
  const uint32_t a = smem_u32(base + idx);
  for (uint32_t r = 0; r < nrank; ++r) st_shared_cluster_f32(map_shared_rank(a, r), v);
}


}  // namespace fastptx
}  // namespace lt
