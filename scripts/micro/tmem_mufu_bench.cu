// Microbenchmarks that size the attention kernels (round 2): per-SM throughput of tcgen05.ld (TMEM -> registers)
// and of MUFU.EX2, as a function of the number of warps.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -I mertools_b200/csrc scripts/micro/tmem_mufu_bench.cu -o gpurun_out/tmem_mufu_bench ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include "mer_common.cuh"
void mer_set_error(const char*, ...) {}
using namespace mer;

template <int DEPTH>  // loads in flight before a wait
__global__ void tmem_ld_kernel(int iters, long long* cycles, unsigned* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[DEPTH][32];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) tmem_ld_32x32(base + ((i * DEPTH + d) * 32 & 255) + (warp >> 2) * 0, r[d]);
    tmem_ld_wait();
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[d][j];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

__global__ void ex2_kernel(int iters, long long* cycles, float* sink) {
  float x[8];
  for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0;
  for (int j = 0; j < 8; ++j) s += x[j];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  long long* cyc; unsigned* sink;
  cudaMalloc(&cyc, 1024 * 8); cudaMalloc(&sink, 1024 * 1024 * 4);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    for (int depth : {1, 2, 4}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (depth == 1) tmem_ld_kernel<1><<<1, warps * 32>>>(iters, cyc, sink);
        if (depth == 2) tmem_ld_kernel<2><<<1, warps * 32>>>(iters, cyc, sink);
        if (depth == 4) tmem_ld_kernel<4><<<1, warps * 32>>>(iters, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      const double bytes = (double)iters * depth * 4096.0 * warps;
      printf("tcgen05.ld.32x32b.x32: %2d warps, %d loads per wait: %.1f B/clk/SM (%.0f clk per load per warp)\n", warps, depth,
             bytes / h, (double)h / (iters * depth));
    }
  }
  for (int warps : {4, 8, 16, 32}) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      ex2_kernel<<<1, warps * 32>>>(iters, cyc, (float*)sink);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("MUFU.EX2: %2d warps: %.2f results/clk/SM\n", warps, (double)iters * 8 * warps * 32 / h);
  }
  return 0;
}
