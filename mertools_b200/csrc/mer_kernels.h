// mer_kernels.h — internal launch prototypes shared between the .cu files of libmer_b200.so.
// The public C ABI lives in include/mer_b200.h; the structs used here are defined there.
#pragma once
#include <cuda_runtime.h>
#include "../../include/mer_b200.h"

// gemm_tf32.cu
int mer_gemm_tf32_launch(const MerGemmDesc* g, cudaStream_t stream);

// rowwise.cu
int mer_layernorm_launch(const float* x, const float* gamma, const float* beta, float* y,
                         float* acc, long long rows, int dim, float eps, int flags,
                         cudaStream_t stream);

// attention.cu
int mer_attention_launch(const float* qkv, float* ctx, const int* cu_seqlens, int n_seq,
                         int max_seqlen, int heads, int flags, cudaStream_t stream);
