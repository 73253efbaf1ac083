// JointWeightFn backward, weight gradient, on CTA PAIRS (tcgen05 cta_group::2):
//   grad_W_vocab[v, j] = sum_m G[m, v] * h[m, j]        h = tanh(pc[c] + pf[n])
// Second generation of joint_wgrad_tc_kernel (joint_tc.cu).  That kernel is bound by its
// producer warps: every CTA converts (bf16x3 split) ALL V columns of its gradient rows and
// recomputes tanh for ALL 256 hidden columns of its block.  Here a pair of CTAs shares one
// (hidden block, row range) work item and one tcgen05.mma.cta_group::2 of M = 256 (= V), N = 256:
//   * CTA r owns the 128 vocabulary rows v in [128 r, 128 r + 128) of the accumulator, so its
//     producers convert only that half of every gradient row (A operand, MN-major);
//   * the B operand h^T is split along N: CTA r recomputes tanh for 128 of the 256 hidden
//     columns only; the hardware feeds both halves to both SMs;
// so the producer work per SM halves, the stage shrinks from 64 KB to 32 KB (6 stages instead
// of 3) and the [128 x 256] fp32 accumulator takes 256 TMEM columns per CTA.
// STATUS: correct (tests/test_gpu_umma.py) but slower than the single-CTA kernel on a B200
// (10.8 vs 8.4 ms at configs[1]); opt-in with LT_JOINT_WGRAD_PAIR=1.
// Barriers: full[s] lives in the leader (rank 0) and collects the producers of BOTH CTAs;
// empty[s] / done are per CTA, signalled by multicast tcgen05.commit.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "params.cuh"
#include "umma.cuh"

namespace lt {
namespace {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ void mbar_init_n(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LTW_WAIT%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LTW_DONE%=;\n"
      "bra LTW_WAIT%=;\n"
      "LTW_DONE%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LTWC_WAIT%=:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LTWC_DONE%=;\n"
      "bra LTWC_WAIT%=;\n"
      "LTWC_DONE%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask)
               : "memory");
}
__device__ __forceinline__ void mma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::
                   "r"(smem_result_addr), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (2.f * kLog2e)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return fmaf(-2.f, r, 1.f);
}

constexpr int kW2Threads = 544;         // warp 0: MMA issuer (leader); warps 1-16: producers
constexpr int kW2Producers = 512;
constexpr int kW2Stages = 6;
constexpr int kW2K = 32;                // joint rows per stage
constexpr int kW2Passes = kW2K / 32;    // rows per producer thread and stage
constexpr int kW2Half = 128;            // vocabulary rows / hidden columns per CTA

struct Wgrad2Params {
  const float* pc;       // [C, H]
  const float* pf;       // [N, H]
  const float* gl;       // [M, V]   V == 256
  const float* gb;       // [M]
  long long M, rows_per_pair;
  int C, H, V;
  float* gwv;            // [V, H]
  float* gwb;            // [H]
  float* gbv;            // [V]
  float* gbb;            // [1]
};

__global__ void __launch_bounds__(kW2Threads, 1)
joint_wgrad2_kernel(const Wgrad2Params p) {
  extern __shared__ __align__(1024) unsigned char w2smem_raw[];
  unsigned char* base = w2smem_raw + ((1024u - (smem_u32(w2smem_raw) & 1023u)) & 1023u);
  const int V = p.V, H = p.H;
  const uint32_t op_bytes = kW2Half * kW2K * 2;           // one [128 x 32] bf16 operand tile
  const uint32_t stage_bytes = 4 * op_bytes;              // A_hi | A_lo | B_hi | B_lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kW2Stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kW2Stages;
  uint64_t* done = bars + 2 * kW2Stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int nj = H / 256;
  const int jblk = pair % nj;
  const long long m_lo = (long long)(pair / nj) * p.rows_per_pair;
  const long long m_hi = min(p.M, m_lo + p.rows_per_pair);
  const bool has_work = m_lo < m_hi;                      // uniform for the pair
  const int nchunks = has_work ? (int)((m_hi - m_lo + kW2K - 1) / kW2K) : 0;

  if (tid == 0) {
    for (int s = 0; s < kW2Stages; ++s) {
      mbar_init_n(smem_u32(&full[s]), 2 * kW2Producers / 32);     // one arrival per producer warp
      mbar_init_n(smem_u32(&empty[s]), 1);
    }
    mbar_init_n(smem_u32(done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc_2cta(smem_u32(tmem_slot), 256);
  umma::fence_before_thread_sync();
  __syncthreads();
  cluster_sync_all();
  umma::fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && rank == 0 && nchunks > 0) {
      const uint32_t idesc = umma::make_idesc_bf16_mn(256, 256);
      const uint32_t sbo = (uint32_t)(kW2Half / 64) * 1024;       // between 8-deep K blocks
      for (int ch = 0; ch < nchunks; ++ch) {
        const int s = ch % kW2Stages;
        mbar_wait_parity_cluster(smem_u32(&full[s]), (ch / kW2Stages) & 1);
        umma::fence_after_thread_sync();
        const uint32_t sa = smem_u32(base) + s * stage_bytes;
        const uint32_t sb = sa + 2 * op_bytes;
#pragma unroll
        for (int ks = 0; ks < kW2K / 16; ++ks) {
          const uint32_t off = 2 * ks * sbo;
          const uint64_t dah = umma::make_smem_desc_mn_sw128(sa + off, 1024, sbo);
          const uint64_t dal = umma::make_smem_desc_mn_sw128(sa + op_bytes + off, 1024, sbo);
          const uint64_t dbh = umma::make_smem_desc_mn_sw128(sb + off, 1024, sbo);
          const uint64_t dbl = umma::make_smem_desc_mn_sw128(sb + op_bytes + off, 1024, sbo);
          mma_bf16_2cta(tmem, dah, dbh, idesc, (ch | ks) > 0);
          mma_bf16_2cta(tmem, dah, dbl, idesc, 1);
          mma_bf16_2cta(tmem, dal, dbh, idesc, 1);
        }
        commit_2cta(smem_u32(&empty[s]));
      }
      commit_2cta(smem_u32(done));
    }
  } else if (nchunks > 0) {
    // producers: thread = (rows k0, k0 + 32 of the stage, 16-byte chunk) for BOTH operands
    const int pidx = tid - 32;                           // 0 .. 511
    const int chk = pidx & 15, k0 = pidx >> 4;           // 16 chunks x 32 rows per pass
    const int v0 = (int)rank * kW2Half + chk * 8;                    // A: vocabulary columns
    const int j0 = jblk * 256 + (int)rank * kW2Half + chk * 8;       // B: hidden columns
    float bv_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float wb_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float bb_acc = 0.f;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t a_off[kW2Passes];
#pragma unroll
    for (int q = 0; q < kW2Passes; ++q) a_off[q] = umma::mn_major_chunk_offset(kW2Half, chk * 8, k0 + 32 * q);
    // loads of the next stage are issued before the current one is converted
    float4 nax[kW2Passes][2], nbp[kW2Passes][2], nbf[kW2Passes][2];
    float ngb[kW2Passes];
    auto issue = [&](int ch) {
#pragma unroll
      for (int q = 0; q < kW2Passes; ++q) {
        const long long m = m_lo + (long long)ch * kW2K + k0 + 32 * q;
        nax[q][0] = nax[q][1] = nbp[q][0] = nbp[q][1] = nbf[q][0] = nbf[q][1] = z4;
        ngb[q] = 0.f;
        if (ch < nchunks && m < m_hi) {
          const long long n = m / p.C;
          const int c = (int)(m - n * p.C);
          const float* g = p.gl + (size_t)m * V + v0;
          nax[q][0] = ldg_stream4(g);
          nax[q][1] = ldg_stream4(g + 4);
          const float* pcr = p.pc + (size_t)c * H + j0;
          const float* pfr = p.pf + (size_t)n * H + j0;
          nbp[q][0] = __ldg(reinterpret_cast<const float4*>(pcr));
          nbp[q][1] = __ldg(reinterpret_cast<const float4*>(pcr + 4));
          nbf[q][0] = __ldg(reinterpret_cast<const float4*>(pfr));
          nbf[q][1] = __ldg(reinterpret_cast<const float4*>(pfr + 4));
          ngb[q] = __ldg(p.gb + m);
        }
      }
    };
    issue(0);
    for (int ch = 0; ch < nchunks; ++ch) {
      const int s = ch % kW2Stages;
      uint4 ah[kW2Passes], al[kW2Passes], bh[kW2Passes], bl[kW2Passes];
      float4 cax[kW2Passes][2], cbp[kW2Passes][2], cbf[kW2Passes][2];
      float cgb[kW2Passes];
#pragma unroll
      for (int q = 0; q < kW2Passes; ++q) {
        cax[q][0] = nax[q][0]; cax[q][1] = nax[q][1];
        cbp[q][0] = nbp[q][0]; cbp[q][1] = nbp[q][1];
        cbf[q][0] = nbf[q][0]; cbf[q][1] = nbf[q][1];
        cgb[q] = ngb[q];
      }
      issue(ch + 1);
#pragma unroll
      for (int q = 0; q < kW2Passes; ++q) {
        const bool live = m_lo + (long long)ch * kW2K + k0 + 32 * q < m_hi;
        const float x[8] = {cax[q][0].x, cax[q][0].y, cax[q][0].z, cax[q][0].w,
                            cax[q][1].x, cax[q][1].y, cax[q][1].z, cax[q][1].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) bv_acc[e] += x[e];
        umma::split_pack8(x, ah[q], al[q]);
        float t[8] = {cbp[q][0].x + cbf[q][0].x, cbp[q][0].y + cbf[q][0].y,
                      cbp[q][0].z + cbf[q][0].z, cbp[q][0].w + cbf[q][0].w,
                      cbp[q][1].x + cbf[q][1].x, cbp[q][1].y + cbf[q][1].y,
                      cbp[q][1].z + cbf[q][1].z, cbp[q][1].w + cbf[q][1].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          t[e] = live ? tanh_fast(t[e]) : 0.f;
          wb_acc[e] = fmaf(cgb[q], t[e], wb_acc[e]);
        }
        if (chk == 0) bb_acc += cgb[q];
        umma::split_pack8(t, bh[q], bl[q]);
      }
      mbar_wait_parity(smem_u32(&empty[s]), ((ch / kW2Stages) & 1) ^ 1);
      unsigned char* a_hi = base + s * stage_bytes;
#pragma unroll
      for (int q = 0; q < kW2Passes; ++q) {
        *reinterpret_cast<uint4*>(a_hi + a_off[q]) = ah[q];
        *reinterpret_cast<uint4*>(a_hi + op_bytes + a_off[q]) = al[q];
        *reinterpret_cast<uint4*>(a_hi + 2 * op_bytes + a_off[q]) = bh[q];
        *reinterpret_cast<uint4*>(a_hi + 3 * op_bytes + a_off[q]) = bl[q];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(smem_u32(&full[s]));
    }
    // side sums: bias / blank-projection gradients
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(p.gwb + j0 + e, wb_acc[e]);
    if (jblk == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(p.gbv + v0 + e, bv_acc[e]);
      if (chk == 0 && rank == 0) atomicAdd(p.gbb, bb_acc);
    }
  }
  // epilogue: warps 0-3 of each CTA add its 128 vocabulary rows x 256 hidden columns
  __syncthreads();
  if (warp < 4 && nchunks > 0) {
    mbar_wait_parity(smem_u32(done), 0);
    umma::fence_after_thread_sync();
    float* out = p.gwv + (size_t)((int)rank * kW2Half + warp * 32 + lane) * H + jblk * 256;
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) atomicAdd(out + c0 + i, v[i]);
    }
  }
  umma::fence_before_thread_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_2cta(tmem, 256);
}

}  // namespace

bool joint_wgrad2_supported(int64_t N, int C, int H, int V, const void* gl, const void* pc,
                            const void* pf) {
  // Opt-in: measured 10.8 ms against 8.4 ms for the single-CTA kernel at configs[1], although
  // the producer work per SM halves -- the pair's handshake (cluster-scope arrivals on the
  // leader's barrier, multicast commits) costs more than the operand work it saves at 32 rows
  // per stage.  Kept, with the forward pair kernel, as the validated cta_group::2 plumbing.
  if (getenv("LT_JOINT_SIMT") || getenv("LT_JOINT_WGRAD_SIMT") || !getenv("LT_JOINT_WGRAD_PAIR"))
    return false;
  if (V != 256 || H % 256 != 0 || H > 4096) return false;
  if (N * (int64_t)C < 1) return false;
  auto al = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  return al(gl) && al(pc) && al(pf);
}

int joint_wgrad2_launch(const float* pc, const float* pf, const float* gb, const float* gl,
                        int64_t N, int C, int H, int V, float* gwb, float* gbb, float* gwv,
                        float* gbv, cudaStream_t stream) {
  Wgrad2Params p = {};
  p.pc = pc; p.pf = pf; p.gl = gl; p.gb = gb;
  p.M = (long long)N * C; p.C = C; p.H = H; p.V = V;
  p.gwv = gwv; p.gwb = gwb; p.gbv = gbv; p.gbb = gbb;
  int dev = 0, sms = 0;
  LT_CUDA(cudaGetDevice(&dev));
  LT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int nj = H / 256;
  long long ranges = (sms / 2) / nj;
  if (ranges < 1) ranges = 1;
  long long rows = (p.M + ranges - 1) / ranges;
  rows = (rows + kW2K - 1) / kW2K * kW2K;
  ranges = (p.M + rows - 1) / rows;
  p.rows_per_pair = rows;
  const size_t smem = (size_t)kW2Stages * 4 * kW2Half * kW2K * 2 + 16 * 8 + 16 + 1024;
  LT_CUDA(cudaFuncSetAttribute(joint_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * ranges * nj));
  cfg.blockDim = dim3(kW2Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LT_CUDA(cudaLaunchKernelEx(&cfg, joint_wgrad2_kernel, p));
  note_launch();
  return LT_OK;
}

}  // namespace lt
