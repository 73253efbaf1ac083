// Table-driven lattice kernels, second generation: a CLUSTER per utterance.
//
// contexts.NextStateTable (/root/reference/last_torch/contexts.py:266-324) with the
// FrameDependent alignment lattice (alignments.py:286-318); same results as lattice_table.cu
// (which stays the path for FrameLabelDependent and for shapes outside the limits below).
//
//   * a cluster of CL CTAs owns one utterance; CTA r owns the R = ceil(C / CL) SOURCE rows
//     [r R, r R + R) of every frame's [C, V] lexical weights: one contiguous slab, streamed
//     by one `cp.async.bulk` per frame into a ring of NS stages (full mbarriers), NS frames
//     ahead of the recursion -- HBM sees only long sequential reads, each byte once;
//   * forward: destination-major PULL restricted to the local slab.  The global CSR of
//     incoming arcs is ascending in p V + y inside a destination, so the arcs of destination
//     q that start in this CTA's rows are a contiguous piece of q's segment; the pieces are
//     found once by binary search and copied to shared memory as 16-bit slab offsets.  Every
//     thread reduces the local arcs of its destinations (no atomics in any semiring, first
//     arg-max in flat-arc order), writes a partial (m, s) / (max, arc) / sum, and after ONE
//     cluster barrier every CTA merges the CL partials of every destination in rank order
//     (= ascending arc order) -- all CTAs hold the full new alpha, nothing is broadcast;
//   * backward: source-major, one warp per local row, beta'[table[p, y]] a shared-memory
//     gather, ONE exponential per arc for the row log-sum-exp and the arc posterior; the new
//     beta of a row is stored into every CTA's buffer, ONE cluster barrier per frame;
//   * partial / beta buffers alternate per frame, so the single barrier also covers the
//     write-after-read hazards, and the slab stage of a frame is refilled right after it.
#include <stdlib.h>

#include "common.cuh"
#include "fast_ptx.cuh"
#include "params.cuh"
#include "table_params.cuh"

namespace lt {

namespace {

using namespace fastptx;

constexpr int kT2Threads = 256;
constexpr int kT2Warps = kT2Threads / 32;
constexpr int kT2MaxQ = 4;          // destinations per thread in the forward kernel
constexpr int kT2MaxStages = 4;

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int lo, int hi,
                                               int key) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ float2 ld_cluster_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr)
               : "memory");
  return v;
}

// ============================================================== forward ==
template <int SR>
__global__ void __launch_bounds__(kT2Threads)
table_forward2_kernel(const TableParams p, const int R, const int NS, const uint32_t magic) {
  using S = Sr<SR>;
  extern __shared__ __align__(128) unsigned char t2sm[];
  const int C = p.C, V = p.V, Cp = (C + 3) & ~3;
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank(), CL = cluster_nctarank();
  const int b = blockIdx.x / CL;
  const int row0 = rank * R, nrows = min(R, C - row0);
  const int base = row0 * V, lim = (row0 + nrows) * V;
  const uint32_t slab_bytes = (uint32_t)nrows * V * 4;
  const size_t stage_floats = (size_t)R * V;

  float* slabs = reinterpret_cast<float*>(t2sm);
  unsigned char* ptr = t2sm + (size_t)NS * stage_floats * 4;
  float* alpha = reinterpret_cast<float*>(ptr); ptr += (size_t)2 * Cp * 4;
  float2* part = reinterpret_cast<float2*>(ptr); ptr += (size_t)2 * Cp * 8;
  int* seg_lo = reinterpret_cast<int*>(ptr); ptr += (size_t)Cp * 4;
  int* seg_n = reinterpret_cast<int*>(ptr); ptr += (size_t)Cp * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ptr); ptr += kT2MaxStages * 8;
  uint16_t* arcs = reinterpret_cast<uint16_t*>(ptr);        // [nrows * V] slab offsets

  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  // local piece of every destination's arc segment (global positions in seg_lo for now)
  for (int q = tid; q < C; q += kT2Threads) {
    const int lo = p.in_offsets[q], hi = p.in_offsets[q + 1];
    const int a0 = lower_bound_i32(p.in_arcs, lo, hi, base);
    const int a1 = lower_bound_i32(p.in_arcs, a0, hi, lim);
    seg_lo[q] = a0;
    seg_n[q] = a1 - a0;
  }
  for (int c = tid; c < C; c += kT2Threads)
    alpha[c] = p.alpha_init ? p.alpha_init[(size_t)b * C + c] : (c == 0 ? S::one() : S::zero());
  __syncthreads();
  int* gpos = reinterpret_cast<int*>(part);                 // scratch: global segment starts
  for (int q = tid; q < C; q += kT2Threads) gpos[q] = seg_lo[q];
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int q = 0; q < C; ++q) { seg_lo[q] = run; run += seg_n[q]; }
  }
  __syncthreads();
  for (int q = tid; q < C; q += kT2Threads) {
    const int n = seg_n[q], g0 = gpos[q], l0 = seg_lo[q];
    for (int i = 0; i < n; ++i) arcs[l0 + i] = (uint16_t)(p.in_arcs[g0 + i] - base);
  }
  __syncthreads();
  cluster_sync_all();

  auto issue = [&](int t) {
    const int s = t % NS;
    const uint32_t bar = smem_u32(&bars[s]);
    mbar_arrive_expect_tx(bar, slab_bytes);
    bulk_load_1d(smem_u32(slabs + (size_t)s * stage_floats),
                 p.lexical + (bt0 + t) * (size_t)C * V + base, slab_bytes, bar);
  };
  if (tid == 0)
    for (int t = 0; t < NS && t < nf; ++t) issue(t);

  float* cur = alpha;
  float* nxt = alpha + Cp;
  for (int t = 0; t < nf; ++t) {
    const int stage = t % NS;
    float2* pbuf = part + (size_t)(t & 1) * Cp;
    float bl[kT2MaxQ];
#pragma unroll
    for (int j = 0; j < kT2MaxQ; ++j) {
      const int q = tid + j * kT2Threads;
      bl[j] = q < C ? ldg_stream(p.blank + (bt0 + t) * C + q) : 0.f;
    }
    mbar_wait(smem_u32(&bars[stage]), (t / NS) & 1);
    const float* slab = slabs + (size_t)stage * stage_floats;
    const float* src = cur + row0;
#pragma unroll
    for (int j = 0; j < kT2MaxQ; ++j) {
      const int q = tid + j * kT2Threads;
      if (q >= C) break;
      if (p.alphas && (uint32_t)q % CL == rank) p.alphas[(bt0 + t) * C + q] = cur[q];
      const uint16_t* al = arcs + seg_lo[q];
      const int n = seg_n[q];
      Acc<SR> acc;
      acc.init();
      int i = 0;
      if constexpr (SR == LT_LOG) {
        for (; i + 4 <= n; i += 4) {
          float x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t a = al[i + u];
            x[u] = src[__umulhi(a, magic)] + slab[a];
          }
          acc.add_chunk(x, fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])));
        }
      }
      for (; i < n; ++i) {
        const uint32_t a = al[i];
        acc.add(S::times(src[__umulhi(a, magic)], slab[a]), base + (int)a);
      }
      float2 out;
      if constexpr (SR == LT_LOG) out = make_float2(acc.m, acc.s);
      else if constexpr (SR == LT_MAXTROPICAL) out = make_float2(acc.m, __int_as_float(acc.a));
      else out = make_float2(acc.s, 0.f);
      pbuf[q] = out;
    }
    __syncwarp();
    cluster_sync_all();      // partials of every CTA visible; this CTA is done with the stage
    if (tid == 0 && t + NS < nf) issue(t + NS);

#pragma unroll
    for (int j = 0; j < kT2MaxQ; ++j) {
      const int q = tid + j * kT2Threads;
      if (q >= C) break;
      const uint32_t mine = smem_u32(&pbuf[q]);
      Acc<SR> tot;
      tot.init();
      for (uint32_t r = 0; r < CL; ++r) {
        const float2 pr = ld_cluster_f2(map_shared_rank(mine, r));
        Acc<SR> o;
        if constexpr (SR == LT_LOG) { o.m = pr.x; o.s = pr.y; }
        else if constexpr (SR == LT_MAXTROPICAL) { o.m = pr.x; o.a = __float_as_int(pr.y); }
        else { o.s = pr.x; }
        if (r == 0) tot = o; else tot.merge(o);
      }
      const float a0 = S::times(cur[q], bl[j]);
      float v;
      if constexpr (SR == LT_MAXTROPICAL) {
        const float rr = tot.value();
        const bool take_blank = a0 >= rr;            // semirings.py:363
        v = take_blank ? a0 : rr;
        if (p.backarc && (uint32_t)q % CL == rank)
          p.backarc[(bt0 + t) * C + q] = take_blank ? -1 : tot.arg();
      } else {
        v = S::plus(a0, tot.value());
      }
      nxt[q] = v;
    }
    __syncthreads();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }

  for (int c = tid; c < C; c += kT2Threads) {
    if ((uint32_t)c % CL != rank) continue;
    if (p.alphas)
      for (int t = nf; t < p.T; ++t) p.alphas[(bt0 + t) * C + c] = cur[c];
    if (p.alpha_final) p.alpha_final[(size_t)b * C + c] = cur[c];
  }
  if (rank == 0 && tid < 32) {         // dist = (+)_c alpha_T[c]  (lattices.py:496)
    Acc<SR> acc;
    acc.init();
    for (int c = tid; c < C; c += 32) acc.add(cur[c], c);
    for (int o = 16; o > 0; o >>= 1) {
      Acc<SR> other;
      if constexpr (SR == LT_LOG) {
        other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
        other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
      } else if constexpr (SR == LT_MAXTROPICAL) {
        other.m = __shfl_xor_sync(0xffffffffu, acc.m, o);
        other.a = __shfl_xor_sync(0xffffffffu, acc.a, o);
      } else {
        other.s = __shfl_xor_sync(0xffffffffu, acc.s, o);
      }
      acc.merge(other);
    }
    if (tid == 0) p.dist[b] = acc.value();
  }
  cluster_sync_all();      // nobody leaves while its partials may still be read
}

// ============================================================= backward ==
template <int SR>
__global__ void __launch_bounds__(kT2Threads)
table_backward2_kernel(const TableParams p, const int R, const int NS) {
  using S = Sr<SR>;
  extern __shared__ __align__(128) unsigned char t2sm[];
  const int C = p.C, V = p.V, Cp = (C + 3) & ~3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank(), CL = cluster_nctarank();
  const int b = blockIdx.x / CL;
  const int row0 = rank * R, nrows = min(R, C - row0);
  const int base = row0 * V;
  const uint32_t slab_bytes = (uint32_t)nrows * V * 4;
  const size_t stage_floats = (size_t)R * V;

  float* slabs = reinterpret_cast<float*>(t2sm);
  unsigned char* ptr = t2sm + (size_t)NS * stage_floats * 4;
  float* beta_buf = reinterpret_cast<float*>(ptr); ptr += (size_t)2 * Cp * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ptr); ptr += kT2MaxStages * 8;
  uint16_t* tbl = reinterpret_cast<uint16_t*>(ptr);         // [nrows * V] next states of the slab

  const int nf = max(0, min(p.num_frames[b], p.T));
  const size_t bt0 = (size_t)b * p.T;
  const float logz = p.dist_in[b];
  const float gscale = p.grad_dist ? p.grad_dist[b] : 1.f;
  const bool scale_ok = (SR != LT_LOG) || is_finite(logz);

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int c = tid; c < 2 * Cp; c += kT2Threads) beta_buf[c] = S::one();   // lattices.py:789-790
  for (int i = tid; i < nrows * V; i += kT2Threads) tbl[i] = (uint16_t)p.table[base + i];
  __syncthreads();
  cluster_sync_all();

  auto issue = [&](int it) {
    const int t = nf - 1 - it;
    const int s = it % NS;
    const uint32_t bar = smem_u32(&bars[s]);
    mbar_arrive_expect_tx(bar, slab_bytes);
    bulk_load_1d(smem_u32(slabs + (size_t)s * stage_floats),
                 p.lexical + (bt0 + t) * (size_t)C * V + base, slab_bytes, bar);
  };
  if (tid == 0)
    for (int it = 0; it < NS && it < nf; ++it) issue(it);

  // padding frames: zero gradients of this CTA's rows (lattices.py:775-779)
  for (int t = nf; t < p.T; ++t) {
    float* gl = p.grad_lexical + (bt0 + t) * (size_t)C * V + base;
    for (int i = tid * 4; i < nrows * V; i += kT2Threads * 4)
      stg_stream4(gl + i, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int r = tid; r < nrows; r += kT2Threads) p.grad_blank[(bt0 + t) * C + row0 + r] = 0.f;
  }

  // lane l of a warp fetches alpha_t[p], blank_t[p] of the warp's l-th row, one frame ahead
  const int my_lr = warp + lane * kT2Warps;
  float n_alpha = 0.f, n_blank = 0.f;
  if (nf > 0 && my_lr < nrows) {
    n_alpha = p.alphas_in[(bt0 + nf - 1) * C + row0 + my_lr];
    n_blank = ldg_stream(p.blank + (bt0 + nf - 1) * C + row0 + my_lr);
  }

  for (int it = 0; it < nf; ++it) {
    const int t = nf - 1 - it;
    const int stage = it % NS;
    const float c_alpha = n_alpha, c_blank = n_blank;
    if (t > 0 && my_lr < nrows) {
      n_alpha = p.alphas_in[(bt0 + t - 1) * C + row0 + my_lr];
      n_blank = ldg_stream(p.blank + (bt0 + t - 1) * C + row0 + my_lr);
    }
    const float* beta = beta_buf + (size_t)(it & 1) * Cp;         // beta_{t+1}
    float* nxt = beta_buf + (size_t)((it + 1) & 1) * Cp;
    const float* alpha = p.alphas_in + (bt0 + t) * C;
    const float* blank = p.blank + (bt0 + t) * C;
    float* gb = p.grad_blank + (bt0 + t) * C;
    float* gl = p.grad_lexical + (bt0 + t) * (size_t)C * V;
    mbar_wait(smem_u32(&bars[stage]), (it / NS) & 1);
    const float* slab = slabs + (size_t)stage * stage_floats;
    for (int lr = warp; lr < nrows; lr += kT2Warps) {
      const int prow = row0 + lr;
      const int li = (lr - warp) / kT2Warps;                 // warp-uniform
      const float a = li < 32 ? __shfl_sync(0xffffffffu, c_alpha, li) : alpha[prow];
      const float bk = li < 32 ? __shfl_sync(0xffffffffu, c_blank, li) : ldg_stream(blank + prow);
      const float* row = slab + (size_t)lr * V;
      const uint16_t* trow = tbl + (size_t)lr * V;
      float* grow = gl + (size_t)prow * V;
      float rowv;
      if (V <= 256) {
        // the whole row in registers: 8 values per lane
        float x[8], bv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int y = lane + 32 * i;
          if (y < V) {
            bv[i] = beta[trow[y]];
            x[i] = S::times(row[y], bv[i]);
          } else {
            bv[i] = 0.f;
            x[i] = S::zero();
          }
        }
        if constexpr (SR == LT_LOG) {
          float m = x[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) m = fmaxf(m, x[i]);
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          const float ms = msafe(m);
          const float rs = scale_ok ? gscale * fast_exp(a + ms - logz) : 0.f;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int y = lane + 32 * i;
            const float e = fast_exp(x[i] - ms);
            s += e;
            if (y < V) __stcs(grow + y, e * rs);
          }
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          rowv = ms + fast_log(s);
        } else {
          const float ga = gscale * a;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int y = lane + 32 * i;
            s += x[i];
            if (y < V) __stcs(grow + y, ga * bv[i]);
          }
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          rowv = s;
        }
      } else {
        if constexpr (SR == LT_LOG) {
          float m = neg_inf();
          for (int y = lane; y < V; y += 32) m = fmaxf(m, row[y] + beta[trow[y]]);
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          const float ms = msafe(m);
          const float rs = scale_ok ? gscale * fast_exp(a + ms - logz) : 0.f;
          float s = 0.f;
          for (int y = lane; y < V; y += 32) {
            const float e = fast_exp(row[y] + beta[trow[y]] - ms);
            s += e;
            grow[y] = e * rs;
          }
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          rowv = ms + fast_log(s);
        } else {
          const float ga = gscale * a;
          float s = 0.f;
          for (int y = lane; y < V; y += 32) {
            const float bvv = beta[trow[y]];
            s += row[y] * bvv;
            grow[y] = ga * bvv;
          }
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          rowv = s;
        }
      }
      if (lane == 0) {
        const float bp = beta[prow];
        const float bb = S::times(bk, bp);
        if constexpr (SR == LT_LOG) gb[prow] = scale_ok ? gscale * fast_exp(a + bb - logz) : 0.f;
        else gb[prow] = gscale * a * bp;
        bcast_f32(nxt, prow, S::plus(bb, rowv), CL);
      }
    }
    __syncwarp();
    cluster_sync_all();      // beta_t complete everywhere; this CTA is done with the stage
    if (tid == 0 && it + NS < nf) issue(it + NS);
  }
  cluster_sync_all();
}

// ------------------------------------------------------------------ host ----
struct T2Geom {
  int CL, R, NS;
  size_t smem;
  uint32_t magic;
};

static size_t t2_fixed_bytes(const TableParams& p, int R, bool backward) {
  const size_t Cp = (p.C + 3) & ~3;
  if (backward) return 2 * Cp * 4 + kT2MaxStages * 8 + (((size_t)R * p.V * 2 + 15) & ~(size_t)15);
  return 2 * Cp * 4 + 2 * Cp * 8 + 2 * Cp * 4 + kT2MaxStages * 8 + (((size_t)R * p.V * 2 + 15) & ~(size_t)15);
}

static bool t2_geometry(const TableParams& p, bool backward, T2Geom* g) {
  if (getenv("LT_TABLE_V1")) return false;
  if (p.k >= 1 || p.T <= 0) return false;
  if (p.V % 4 != 0) return false;                                   // 16-byte bulk copies
  if (reinterpret_cast<uintptr_t>(p.lexical) % 16 != 0) return false;
  if (backward && reinterpret_cast<uintptr_t>(p.grad_lexical) % 16 != 0) return false;
  if (!backward && p.C > kT2Threads * kT2MaxQ) return false;
  if (p.C > 65535) return false;                                    // 16-bit next states
  const char* env = getenv("LT_TABLE_CLUSTER");
  const int forced = env ? atoi(env) : 0;
  for (int cl = 1; cl <= 8; cl <<= 1) {
    const int R = (p.C + cl - 1) / cl;
    if ((cl - 1) * R >= p.C) continue;                              // a rank without rows
    const size_t slab = (size_t)R * p.V * 4;
    if (forced ? cl != forced : (slab > 40 * 1024 && cl < 8)) continue;
    if (!backward && (size_t)R * p.V > 65535) return false;         // 16-bit slab offsets
    const size_t fixed = t2_fixed_bytes(p, R, backward);
    size_t budget = 113 * 1024;                                     // two CTAs per SM
    if (fixed + 2 * slab > budget) budget = 227 * 1024;
    if (fixed + 2 * slab > budget) return false;
    int ns = (int)((budget - fixed) / slab);
    if (ns > kT2MaxStages) ns = kT2MaxStages;
    g->CL = cl; g->R = R; g->NS = ns;
    g->smem = fixed + (size_t)ns * slab;
    g->magic = (uint32_t)(((1ull << 32) + p.V - 1) / p.V);          // a / V for a < 2^16
    return true;
  }
  return false;
}

template <typename KernelT, typename... Args>
static int launch_t2(KernelT kernel, const T2Geom& g, int B, cudaStream_t stream, Args... args) {
  LT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)B * g.CL);
  cfg.blockDim = dim3(kT2Threads);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LT_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  note_launch();
  return LT_OK;
}

}  // namespace

int table2_cluster_size(int C, int V, int k, bool backward) {
  TableParams p = {};
  p.C = C; p.V = V; p.k = k; p.B = 1; p.T = 1;
  T2Geom g;
  return t2_geometry(p, backward, &g) ? g.CL : 0;     // null pointers count as aligned
}

bool table2_forward_supported(const TableParams& p) {
  T2Geom g;
  return t2_geometry(p, false, &g);
}

bool table2_backward_supported(const TableParams& p) {
  T2Geom g;
  return t2_geometry(p, true, &g);
}

int table2_forward_launch(int semiring, const TableParams& p, cudaStream_t stream) {
  T2Geom g;
  if (!t2_geometry(p, false, &g)) { set_error("table cluster path: unsupported shape"); return LT_ERR_UNSUPPORTED; }
  if (semiring == LT_LOG)
    return launch_t2(table_forward2_kernel<LT_LOG>, g, p.B, stream, p, g.R, g.NS, g.magic);
  if (semiring == LT_MAXTROPICAL)
    return launch_t2(table_forward2_kernel<LT_MAXTROPICAL>, g, p.B, stream, p, g.R, g.NS, g.magic);
  return launch_t2(table_forward2_kernel<LT_REAL>, g, p.B, stream, p, g.R, g.NS, g.magic);
}

int table2_backward_launch(int semiring, const TableParams& p, cudaStream_t stream) {
  T2Geom g;
  if (!t2_geometry(p, true, &g)) { set_error("table cluster path: unsupported shape"); return LT_ERR_UNSUPPORTED; }
  if (semiring == LT_LOG)
    return launch_t2(table_backward2_kernel<LT_LOG>, g, p.B, stream, p, g.R, g.NS);
  return launch_t2(table_backward2_kernel<LT_REAL>, g, p.B, stream, p, g.R, g.NS);
}

}  // namespace lt
