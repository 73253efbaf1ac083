// L2-hit and HBM read bandwidth of this GPU as seen by plain 16-byte loads: every CTA sweeps the
// whole buffer (so every SM pulls every line), buffer sizes from inside the 126 MB L2 to well
// beyond it.  The FrameLabelDependent kernels stream a frame k times and take the re-reads from
// L2: this is the number their L2 -> SM traffic is compared with.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2bw tools/l2_bw.cu && /tmp/l2bw
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) read_kernel(const float4* __restrict__ buf, size_t n4, int reps,
                                                   float* out) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    // rotate the starting point per repetition so that the access order differs from the last pass
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 4
    for (; i < n4; i += stride) {
      float4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(buf + i));
      acc += (v.x + v.y) + (v.z + v.w);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t max_bytes = (size_t)2048 << 20;
  float4* buf;
  float* out;
  cudaMalloc(&buf, max_bytes);
  cudaMalloc(&out, 4);
  cudaMemset(buf, 0, max_bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int mbs[] = {8, 16, 32, 48, 64, 80, 96, 112, 128, 192, 512, 2048};
  for (int mb : mbs) {
    const size_t bytes = (size_t)mb << 20, n4 = bytes / 16;
    const int reps = mb <= 128 ? 64 : (mb <= 512 ? 16 : 6);
    for (int blocks_per_sm = 2; blocks_per_sm <= 4; blocks_per_sm += 2) {
      read_kernel<<<sms * blocks_per_sm, 512>>>(buf, n4, 2, out);     // warm (fills L2)
      cudaEventRecord(a);
      read_kernel<<<sms * blocks_per_sm, 512>>>(buf, n4, reps, out);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, a, b);
      printf("buffer %5d MB  %d CTAs/SM x 512 thr: %8.1f GB/s\n", mb, blocks_per_sm,
             (double)bytes * reps / ms / 1e6);
    }
  }
  return cudaGetLastError() != cudaSuccess;
}
