// Minimal hand-written tcgen05 / TMEM plumbing for sm_100a (no CUTLASS):
// shared-memory matrix descriptors (K-major, SWIZZLE_128B), instruction
// descriptors for kind::f16 (bf16 inputs, fp32 accumulate), TMEM alloc / ld /
// commit / fences.  Bit layouts follow the PTX ISA "tcgen05" matrix- and
// instruction-descriptor tables.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace lt {
namespace umma {

// ---- K-major operand tile in shared memory, 128-byte rows, SWIZZLE_128B --------
// Row r (an M or N index) occupies 128 bytes = 64 bf16 along K.  Rows are grouped
// in atoms of 8 rows x 128 B = 1024 B (tile base must be 1024-byte aligned); the
// 16-byte chunk c of row r is stored at chunk position c ^ (r & 7).
__device__ __forceinline__ uint32_t swizzled_offset(int row, int chunk16) {
  return (uint32_t)row * 128u + (uint32_t)((chunk16 ^ (row & 7)) << 4);
}

// 64-bit shared-memory matrix descriptor.
//   [0,14)  start address >> 4          [16,30) leading byte offset >> 4 (unused for
//   [32,46) stride byte offset >> 4 (= 1024 B between 8-row atoms)      swizzled K-major)
//   [46,48) version = 1 (Blackwell)      [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;            // SBO
  d |= (uint64_t)1 << 46;                       // version
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}

// 32-bit instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major.
//   [4,6) c_format 1=F32   [7,10) a_format 1=BF16   [10,13) b_format 1=BF16
//   [15] a_major 0=K  [16] b_major 0=K   [17,23) N>>3   [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- MN-major operand tile, SWIZZLE_128B ------------------------------------------
// The contiguous direction is M (or N): 64 bf16 along MN form one 128-byte row, 8
// consecutive K indices form a 1024-byte atom.  A tile [MN_total x K] is stored as
// atoms indexed (kb = k / 8, mnb = mn / 64) at  (kb * (MN_total / 64) + mnb) * 1024:
//   LBO = byte stride between consecutive 64-wide MN blocks  (1024)
//   SBO = byte stride between consecutive 8-deep K blocks    (MN_total / 64 * 1024)
// and inside an atom element (i = mn % 64, r = k % 8) sits at
//   r * 128 + (((i / 8) ^ r) << 4) + (i % 8) * 2.
__device__ __forceinline__ uint32_t mn_major_chunk_offset(int mn_total, int mn, int k) {
  // byte offset of the 16-byte chunk that holds elements mn .. mn+7 (mn % 8 == 0) at depth k
  const int kb = k >> 3, r = k & 7, mnb = mn >> 6, ch = (mn & 63) >> 3;
  return (uint32_t)(kb * (mn_total >> 6) + mnb) * 1024u + (uint32_t)r * 128u +
         (uint32_t)((ch ^ r) << 4);
}
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                            uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, D fp32, A/B bf16, both MN-major (bits 15 and 16 set)
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::
                   "r"(smem_result_addr), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void fence_before_thread_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory ("TS" form): A [128 x 16] bf16 sits in TMEM with row i in lane i and
// two consecutive K elements per 32-bit column (8 columns per instruction, even K in the low half);
// only B is read from shared memory.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Each thread of the warp writes N consecutive 32-bit columns of ITS TMEM lane.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// mbarrier arrive (count 1) when all previously issued MMAs of this thread completed.
__device__ __forceinline__ void commit(uint32_t mbar_smem_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                   "r"(mbar_smem_addr) : "memory");
}
// The same arrival delivered to the barrier at this offset in every CTA of the mask.
__device__ __forceinline__ void commit_mc(uint32_t mbar_smem_addr, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(mbar_smem_addr), "h"(cta_mask) : "memory");
}
// Each thread of the warp reads 32 consecutive fp32 columns of ITS TMEM lane.
// taddr = (lane_base << 16) | column; a warp may only touch lanes 32*(warp%4)..+31.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// fp32 -> (hi, lo) bf16 split: x ~= hi + lo with |x - hi - lo| <= 2^-17 |x|
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// Packed variant: two values per cvt.rn.bf16x2.f32 (one ALU-pipe F2FP instead of two
// XU-pipe F2F), bf16 -> fp32 by a 16-bit shift.  hi / lo hold (x0 low half, x1 high half).
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ void split_pack8(const float (&x)[8], uint4& hi, uint4& lo) {
  split_pack2(x[0], x[1], hi.x, lo.x);
  split_pack2(x[2], x[3], hi.y, lo.y);
  split_pack2(x[4], x[5], hi.z, lo.z);
  split_pack2(x[6], x[7], hi.w, lo.w);
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

}  // namespace umma
}  // namespace lt
