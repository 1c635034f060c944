// hubert_frontend.cu — the HBM-bound head of the audio encoder (SURVEY.md kernel rows A0/A1):
//
//   wave_normalize   : HF Wav2Vec2FeatureExtractor zero-mean / unit-variance, eps 1e-7
//                      (feature_extraction_wav2vec2.py:78-97, called at
//                      extract_audio_huggingface.py:94)
//   conv0_stats      : per (clip, channel) sum / sum-of-squares of Conv1d(1->512, k=10, s=5)
//                      over time, for GroupNorm(512 groups) (modeling_hubert.py:154-175)
//   conv0_apply      : recompute conv0, normalise, affine, exact GELU, write the TIME-MAJOR
//                      activation [B, T0_pad, 512] that the conv1 implicit GEMM reads (split bf16
//                      hi|lo rows for the BF16X3 GEMM, or tf32-rounded fp32).
//
// The fp32 conv0 output (32.8 MB per 5 s clip) is never materialised un-normalised: the waveform
// (320 KB per clip) is read three times instead.  Algorithmic traffic per clip:
// 3 x 320 KB in + 15,999 x 512 x 4 B = 32.8 MB out.
#include <stdlib.h>

#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

constexpr int C0 = 512;
constexpr int K0 = 10;
constexpr int S0 = 5;

__device__ __forceinline__ double block_sum_double(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) r += sh[i];
  __syncthreads();
  return r;
}

// one block per clip; two passes over the waveform (mean, then variance about the mean)
__global__ void __launch_bounds__(1024)
wave_normalize_kernel(const float* __restrict__ in, float* __restrict__ out, int L, long long ld_in,
                      long long ld_out) {
  __shared__ double sh[32];
  const float* x = in + (long long)blockIdx.x * ld_in;
  float* y = out + (long long)blockIdx.x * ld_out;
  double s = 0.0;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s += (double)x[i];
  const double mean = block_sum_double(s, sh) / (double)L;
  double q = 0.0;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const double d = (double)x[i] - mean;
    q += d * d;
  }
  const double var = block_sum_double(q, sh) / (double)L;
  const float meanf = (float)mean;
  const float denom = sqrtf((float)var + 1e-7f);
  for (int i = threadIdx.x; i < L; i += blockDim.x) y[i] = (x[i] - meanf) / denom;
}

// ragged batch: row b holds lengths[b] samples (statistics over those), the rest of the row is written as zeros
__global__ void __launch_bounds__(1024)
wave_normalize_ragged_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ lengths,
                             int L, long long ld_in, long long ld_out) {
  __shared__ double sh[32];
  const float* x = in + (long long)blockIdx.x * ld_in;
  float* y = out + (long long)blockIdx.x * ld_out;
  const int n = lengths[blockIdx.x];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
  const double mean = block_sum_double(s, sh) / (double)n;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)x[i] - mean;
    q += d * d;
  }
  const double var = block_sum_double(q, sh) / (double)n;
  const float meanf = (float)mean;
  const float denom = sqrtf((float)var + 1e-7f);
  for (int i = threadIdx.x; i < L; i += blockDim.x) y[i] = i < n ? (x[i] - meanf) / denom : 0.f;
}

// grid (chunks, B); 512 threads = one channel each; each block covers TCHUNK output frames.
constexpr int TCHUNK = 128;

// RAGGED: clip b has t0s[b] <= T0 frames; GroupNorm statistics and (in apply) their divisor use that count.
template <bool RAGGED>
__global__ void __launch_bounds__(C0)
conv0_stats_kernel(const float* __restrict__ wave, long long ld_wave, const float* __restrict__ w0,
                   int T0, double* __restrict__ stats /*[B,512,2]*/, const int* __restrict__ t0s) {
  __shared__ float xs[TCHUNK * S0 + K0];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TCHUNK;
  if (RAGGED) T0 = t0s[b];
  if (RAGGED && t0 >= T0) return;  // block-uniform: this chunk lies past the clip
  const int nt = min(TCHUNK, T0 - t0);
  const float* x = wave + (long long)b * ld_wave + (long long)t0 * S0;
  const int nx = (nt - 1) * S0 + K0;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = x[i];
  __syncthreads();
  const int c = threadIdx.x;
  float w[K0];
#pragma unroll
  for (int k = 0; k < K0; ++k) w[k] = __ldg(w0 + c * K0 + k);
  float s = 0.f, q = 0.f;
  for (int t = 0; t < nt; ++t) {
    float y = 0.f;
#pragma unroll
    for (int k = 0; k < K0; ++k) y = fmaf(w[k], xs[t * S0 + k], y);
    s += y;
    q = fmaf(y, y, q);
  }
  atomicAdd(&stats[((long long)b * C0 + c) * 2 + 0], (double)s);
  atomicAdd(&stats[((long long)b * C0 + c) * 2 + 1], (double)q);
}

template <bool RAGGED>
__global__ void __launch_bounds__(C0)
conv0_apply_kernel(const float* __restrict__ wave, long long ld_wave, const float* __restrict__ w0,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   const double* __restrict__ stats, int T0, long long out_bstride /*floats*/,
                   int split_out, float* __restrict__ out, const int* __restrict__ t0s) {
  __shared__ float xs[TCHUNK * S0 + K0];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TCHUNK;
  const int nt = min(TCHUNK, T0 - t0);  // RAGGED: frames past the clip's own count are still written (finite, unused)
  if (RAGGED) T0 = t0s[b];              // ... but the statistics were taken over the clip's own frames
  const float* x = wave + (long long)b * ld_wave + (long long)t0 * S0;
  const int nx = (nt - 1) * S0 + K0;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = x[i];
  __syncthreads();
  const int c = threadIdx.x;
  float w[K0];
#pragma unroll
  for (int k = 0; k < K0; ++k) w[k] = __ldg(w0 + c * K0 + k);
  // GroupNorm with num_groups == channels: biased variance over time, eps 1e-5
  const double sum = stats[((long long)b * C0 + c) * 2 + 0];
  const double sq = stats[((long long)b * C0 + c) * 2 + 1];
  const double mean_d = sum / (double)T0;
  double var_d = sq / (double)T0 - mean_d * mean_d;
  if (var_d < 0.0) var_d = 0.0;
  const float mean = (float)mean_d;
  const float rstd = (float)(1.0 / sqrt(var_d + 1e-5));
  const float g = __ldg(gamma + c) * rstd;
  const float bt = __ldg(beta + c);
  float* orow = out + (long long)b * out_bstride + (long long)t0 * C0;
  for (int t = 0; t < nt; ++t) {
    float y = 0.f;
#pragma unroll
    for (int k = 0; k < K0; ++k) y = fmaf(w[k], xs[t * S0 + k], y);
    const float v = gelu_erf_fast((y - mean) * g + bt);
    if (split_out) store_split1(orow + (long long)t * C0, c, v);
    else orow[(long long)t * C0 + c] = round_tf32(v);
  }
}

// ---- opt-in variants (MER_CONV0_PACKED=1): two adjacent channels per thread on the packed fp32 pipe (FFMA2 with
// the sample as a broadcast operand), half the shared-memory loads per output, the packed GELU, 4 + 4 byte split
// stores.  Per channel the conv / statistics arithmetic is the same sequence of fma.rn as the scalar kernels. ----
__global__ void __launch_bounds__(C0 / 2)
conv0_stats2_kernel(const float* __restrict__ wave, long long ld_wave, const float* __restrict__ w0,
                    int T0, double* __restrict__ stats /*[B,512,2]*/) {
  __shared__ float xs[TCHUNK * S0 + K0];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TCHUNK;
  const int nt = min(TCHUNK, T0 - t0);
  const float* x = wave + (long long)b * ld_wave + (long long)t0 * S0;
  const int nx = (nt - 1) * S0 + K0;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = x[i];
  __syncthreads();
  const int c = 2 * threadIdx.x;
  uint64_t w[K0];
#pragma unroll
  for (int k = 0; k < K0; ++k) w[k] = pack2(__ldg(w0 + c * K0 + k), __ldg(w0 + (c + 1) * K0 + k));
  uint64_t s = pack2(0.f, 0.f), q = s;
  for (int t = 0; t < nt; ++t) {
    uint64_t y = pack2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K0; ++k) {
      const float xv = xs[t * S0 + k];
      y = fma2(w[k], pack2(xv, xv), y);
    }
    s = add2(s, y);
    q = fma2(y, y, q);
  }
  float s0, s1, q0, q1;
  unpack2(s, s0, s1);
  unpack2(q, q0, q1);
  double* st = stats + ((long long)b * C0 + c) * 2;
  atomicAdd(st + 0, (double)s0);
  atomicAdd(st + 1, (double)q0);
  atomicAdd(st + 2, (double)s1);
  atomicAdd(st + 3, (double)q1);
}

__global__ void __launch_bounds__(C0 / 2)
conv0_apply2_kernel(const float* __restrict__ wave, long long ld_wave, const float* __restrict__ w0,
                    const float* __restrict__ gamma, const float* __restrict__ beta,
                    const double* __restrict__ stats, int T0, long long out_bstride /*floats*/,
                    float* __restrict__ out /*split bf16 rows*/) {
  __shared__ float xs[TCHUNK * S0 + K0];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TCHUNK;
  const int nt = min(TCHUNK, T0 - t0);
  const float* x = wave + (long long)b * ld_wave + (long long)t0 * S0;
  const int nx = (nt - 1) * S0 + K0;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = x[i];
  __syncthreads();
  const int c = 2 * threadIdx.x;
  uint64_t w[K0];
#pragma unroll
  for (int k = 0; k < K0; ++k) w[k] = pack2(__ldg(w0 + c * K0 + k), __ldg(w0 + (c + 1) * K0 + k));
  float nmean[2], g[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {  // GroupNorm with num_groups == channels: biased variance over time, eps 1e-5
    const double sum = stats[((long long)b * C0 + c + j) * 2 + 0];
    const double sq = stats[((long long)b * C0 + c + j) * 2 + 1];
    const double mean_d = sum / (double)T0;
    double var_d = sq / (double)T0 - mean_d * mean_d;
    if (var_d < 0.0) var_d = 0.0;
    nmean[j] = -(float)mean_d;
    g[j] = __ldg(gamma + c + j) * (float)(1.0 / sqrt(var_d + 1e-5));
  }
  const uint64_t nmean2 = pack2(nmean[0], nmean[1]), g2 = pack2(g[0], g[1]);
  const uint64_t bt2 = pack2(__ldg(beta + c), __ldg(beta + c + 1));
  float* orow = out + (long long)b * out_bstride + (long long)t0 * C0;
  for (int t = 0; t < nt; ++t) {
    uint64_t y = pack2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K0; ++k) {
      const float xv = xs[t * S0 + k];
      y = fma2(w[k], pack2(xv, xv), y);
    }
    float a0, a1, v0, v1;
    unpack2(fma2(add2(y, nmean2), g2, bt2), a0, a1);
    gelu_erf_fast2(a0, a1, v0, v1);
    store_split2(orow + (long long)t * C0, c, v0, v1);
  }
}

// conv0 (k = 10, stride 5, + bias) -> LayerNorm over the 512 channels -> GELU, the first layer of the
// feat_extract_norm="layer" feature encoder (HF HubertLayerNormConvLayer; hubert-large family).
// One warp = LN_ROWS consecutive frames; lane owns channels lane + 32 j (j < 16), weights transposed in
// smem ([tap][channel]: conflict-free), two-pass LayerNorm in registers via warp shuffles.
constexpr int LN_ROWS = 4;
__global__ void __launch_bounds__(256)
conv0_ln_kernel(const float* __restrict__ wave, long long ld_wave, const float* __restrict__ w0,
                const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                int T0, long long out_bstride, float* __restrict__ out) {
  __shared__ float ws[K0][C0];  // 20 KB
  __shared__ float xs[8][LN_ROWS * S0 + K0];
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < C0 * K0; i += blockDim.x) ws[i % K0][i / K0] = __ldg(w0 + i);
  __syncthreads();
  const int t0 = (blockIdx.x * 8 + warp) * LN_ROWS;
  if (t0 >= T0) return;
  const int nt = min(LN_ROWS, T0 - t0);
  const float* x = wave + (long long)b * ld_wave + (long long)t0 * S0;
  for (int i = lane; i < (nt - 1) * S0 + K0; i += 32) xs[warp][i] = x[i];
  __syncwarp();
  float acc[LN_ROWS][16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float bj = bias ? __ldg(bias + lane + 32 * j) : 0.f;
#pragma unroll
    for (int r = 0; r < LN_ROWS; ++r) acc[r][j] = bj;
  }
#pragma unroll
  for (int k = 0; k < K0; ++k) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float w = ws[k][lane + 32 * j];
#pragma unroll
      for (int r = 0; r < LN_ROWS; ++r) acc[r][j] = fmaf(w, xs[warp][r * S0 + k], acc[r][j]);
    }
  }
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    if (r >= nt) break;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[r][j];
    const float mean = warp_sum(s) * (1.0f / C0);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      acc[r][j] -= mean;
      q += acc[r][j] * acc[r][j];
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C0) + 1e-5f);
    float* orow = out + (long long)b * out_bstride + (long long)(t0 + r) * C0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = lane + 32 * j;
      const float v = gelu_erf_fast(acc[r][j] * rstd * __ldg(gamma + c) + __ldg(beta + c));
      store_split1(orow, c, v);
    }
  }
}

}  // namespace

int mer_wave_normalize_launch(const float* in, float* out, int B, int L, long long ld_in,
                              long long ld_out, cudaStream_t stream, const int* lengths) {
  MER_REQUIRE(in && out && B > 0 && L > 0, "mer_wave_normalize: bad arguments");
  if (lengths)
    wave_normalize_ragged_kernel<<<B, 1024, 0, stream>>>(in, out, lengths, L, ld_in, ld_out);
  else
    wave_normalize_kernel<<<B, 1024, 0, stream>>>(in, out, L, ld_in, ld_out);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}

int mer_hubert_conv0_launch(const float* wave, long long ld_wave, int B, int L, const float* w0,
                            const float* gamma, const float* beta, double* stats, float* out,
                            long long out_bstride, int split_out, cudaStream_t stream, const int* t0s) {
  const int T0 = (L - K0) / S0 + 1;
  MER_REQUIRE(T0 > 0, "mer_hubert_conv0: waveform too short (%d samples)", L);
  MER_CUDA_CHECK(cudaMemsetAsync(stats, 0, (size_t)B * C0 * 2 * sizeof(double), stream));
  dim3 grid((T0 + TCHUNK - 1) / TCHUNK, B);
  // both passes: the waveform in twice, the [T0, 512] operand (4 B per element) out once
  const int prof = mer_prof_begin(MER_PROF_CONV0, (double)B * (2.0 * L * 4.0 + (double)T0 * C0 * 4.0), stream);
  const char* pk = getenv("MER_CONV0_PACKED");  // read at every launch: tests run both forms in one process
  if (t0s) {  // ragged batch: per-clip frame counts for the GroupNorm statistics
    conv0_stats_kernel<true><<<grid, C0, 0, stream>>>(wave, ld_wave, w0, T0, stats, t0s);
    MER_CUDA_CHECK(cudaGetLastError());
    conv0_apply_kernel<true><<<grid, C0, 0, stream>>>(wave, ld_wave, w0, gamma, beta, stats, T0, out_bstride, split_out,
                                                      out, t0s);
  } else if (!(pk && atoi(pk) == 0) && split_out) {  // default since round 2 (0.21 -> 0.29 of HBM peak measured); MER_CONV0_PACKED=0: one channel per thread
    conv0_stats2_kernel<<<grid, C0 / 2, 0, stream>>>(wave, ld_wave, w0, T0, stats);
    MER_CUDA_CHECK(cudaGetLastError());
    conv0_apply2_kernel<<<grid, C0 / 2, 0, stream>>>(wave, ld_wave, w0, gamma, beta, stats, T0, out_bstride, out);
  } else {
    conv0_stats_kernel<false><<<grid, C0, 0, stream>>>(wave, ld_wave, w0, T0, stats, nullptr);
    MER_CUDA_CHECK(cudaGetLastError());
    conv0_apply_kernel<false><<<grid, C0, 0, stream>>>(wave, ld_wave, w0, gamma, beta, stats, T0,
                                                       out_bstride, split_out, out, nullptr);
  }
  mer_prof_end(prof, stream);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(2);
  return 0;
}

// ragged batch helpers: clip b occupies rows [b * Tmax, b * Tmax + tb[b]) of a padded [B, Tmax, dim] activation
__global__ void __launch_bounds__(256)
zero_tail_rows_f16_kernel(uint4* __restrict__ x, const int* __restrict__ tb, int Tmax, int dim8) {
  const int b = blockIdx.y;
  const long long n = (long long)(Tmax - tb[b]) * dim8;  // 16-byte slots of the clip's tail rows
  uint4* tail = x + ((long long)b * Tmax + tb[b]) * dim8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    tail[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(256)
pack_rows_kernel(const float4* __restrict__ padded, const int* __restrict__ cu, int Tmax, int dim4,
                 float4* __restrict__ packed) {
  const int b = blockIdx.y;
  const long long n = (long long)(cu[b + 1] - cu[b]) * dim4;
  const float4* src = padded + (long long)b * Tmax * dim4;
  float4* dst = packed + (long long)cu[b] * dim4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

int mer_zero_tail_rows_f16_launch(void* x16, const int* tb, int B, int Tmax, int dim, cudaStream_t stream) {
  MER_REQUIRE(x16 && tb && B > 0 && Tmax > 0 && dim % 8 == 0, "mer_zero_tail_rows_f16: bad operands");
  zero_tail_rows_f16_kernel<<<dim3(32, B), 256, 0, stream>>>(static_cast<uint4*>(x16), tb, Tmax, dim / 8);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}

int mer_pack_rows_launch(const float* padded, const int* cu, int B, int Tmax, int dim, float* packed,
                         cudaStream_t stream) {
  MER_REQUIRE(padded && cu && packed && padded != packed && B > 0 && dim % 4 == 0, "mer_pack_rows: bad operands");
  pack_rows_kernel<<<dim3(32, B), 256, 0, stream>>>(reinterpret_cast<const float4*>(padded), cu, Tmax, dim / 4,
                                                    reinterpret_cast<float4*>(packed));
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}

int mer_hubert_conv0_ln_launch(const float* wave, long long ld_wave, int B, int L, const float* w0,
                               const float* bias, const float* gamma, const float* beta, float* out,
                               long long out_bstride, cudaStream_t stream) {
  const int T0 = (L - K0) / S0 + 1;
  MER_REQUIRE(T0 > 0, "mer_hubert_conv0_ln: waveform too short (%d samples)", L);
  dim3 grid((T0 + 8 * LN_ROWS - 1) / (8 * LN_ROWS), B);
  conv0_ln_kernel<<<grid, 256, 0, stream>>>(wave, ld_wave, w0, bias, gamma, beta, T0, out_bstride, out);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}

extern "C" int mer_wave_normalize(const float* in, float* out, int batch, int n_samples,
                                  long long ld_in, long long ld_out, void* stream) {
  return mer_wave_normalize_launch(in, out, batch, n_samples, ld_in, ld_out,
                                   static_cast<cudaStream_t>(stream));
}
