// How fast can one SM pull strided 128-byte rows through TMA?  (Sizing of the attention kernels' operand loads:
// the phase traces show ~96 KB per item arriving in ~8k cycles.)  Every CTA (one per SM) loads boxes of ROWS x 128 B
// from a 2-D fp16 tensor [n_rows][pitch] with DEPTH boxes in flight; the buffer is either small (L2-resident) or
// large (streams from HBM).  Prints bytes per clock and SM.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include "mer_common.cuh"
void mer_set_error(const char*, ...) {}
int mer_make_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&fn, cudaEnableDefault, &q);
  }
  uint32_t es[5] = {1, 1, 1, 1, 1};
  return fn(out, dtype, rank, const_cast<void*>(base), dims, strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}
using namespace mer;

template <int DEPTH>
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, int rows_per_box, int n_rows,
                                                     int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bars[DEPTH];
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = rows_per_box * 128;
    int row = (blockIdx.x * 7919) % (n_rows - rows_per_box);
    const long long t0 = clock64();
    for (int i = 0; i < iters + DEPTH; ++i) {
      const int s = i % DEPTH;
      if (i >= DEPTH) mbar_wait(&bars[s], ((i / DEPTH) - 1) & 1);
      if (i < iters) {
        mbar_expect_tx(&bars[s], bytes);
        tma_load_2d(smem + s * 32768, &tm, &bars[s], (blockIdx.x % 12) * 64, row);
        row = (row + 1237 * rows_per_box) % (n_rows - rows_per_box);
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  int sms = 148;
  long long* cyc;
  cudaMalloc(&cyc, 1024 * 8);
  for (int big = 0; big < 2; ++big) {
    const long long n_rows = big ? 400000 : 8000;  // x 4608 B: 1.8 GB (HBM) | 37 MB (L2-resident)
    uint16_t* buf;
    cudaMalloc(&buf, n_rows * 4608);
    cudaMemset(buf, 0, n_rows * 4608);
    for (int rows : {64, 128, 256}) {
      CUtensorMap tm;
      const uint64_t dims[2] = {2304, (uint64_t)n_rows};
      const uint64_t strides[1] = {4608};
      const uint32_t box[2] = {64, (uint32_t)rows};
      if (mer_make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) {
        printf("tmap failed\n");
        return 1;
      }
      for (int depth : {1, 2, 4, 6}) {
        const int iters = 400;
        const size_t smem = 6 * 32768 + 1024;
        for (int rep = 0; rep < 2; ++rep) {
#define RUN(D) { cudaFuncSetAttribute(tma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                 tma_kernel<D><<<sms, 128, smem>>>(tm, rows, (int)n_rows, iters, cyc); }
          if (depth == 1) RUN(1) else if (depth == 2) RUN(2) else if (depth == 4) RUN(4) else RUN(6)
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%s  box %3d rows x 128 B, %d in flight: %.1f B/clk/SM  (%.0f clk per box)\n", big ? "HBM" : "L2 ", rows, depth,
               (double)iters * rows * 128 / mx, (double)mx / iters);
      }
    }
    cudaFree(buf);
  }
  return 0;
}
