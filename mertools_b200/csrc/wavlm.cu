// wavlm.cu — what WavLM's attention (HF modeling_wavlm.py WavLMAttention; extract_audio_huggingface.py:36-37 lists
// wavlm-base / wavlm-large) needs beyond the shared kernels: the per-(token, head) gate of the relative position bias
// and softmax attention with an additive bias.  The bias table [heads, T, T] (bucketed relative positions through
// layer 0's embedding) is the same for every clip and layer and is built once by the host; the gate depends on the
// layer input.  fp32 CUDA-core kernels, not tuned: first correct path for this model family.
#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

// one warp per (token, head): 8 dot products of length 64
__global__ void __launch_bounds__(256)
wavlm_gate_kernel(const float* __restrict__ x, long long tokens, int heads, const float* __restrict__ w,
                  const float* __restrict__ b, const float* __restrict__ c, float* __restrict__ gate) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= tokens * heads) return;
  const long long tok = wid / heads;
  const int h = (int)(wid % heads);
  const float* xr = x + (tok * heads + h) * 64;
  const float x0 = xr[lane], x1 = xr[lane + 32];
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const float d = warp_sum(fmaf(__ldg(w + r * 64 + lane), x0, __ldg(w + r * 64 + lane + 32) * x1));
    s[r >> 2] += d + __ldg(b + r);
  }
  if (lane == 0) {
    const float ga = 1.0f / (1.0f + expf(-s[0])), gb = 1.0f / (1.0f + expf(-s[1]));
    gate[wid] = ga * (gb * __ldg(c + h) - 1.0f) + 2.0f;
  }
}

// grid (ceil(T / 8), heads, batch), 8 warps = 8 query rows; scores of a row live in the warp's slice of shared memory
__global__ void __launch_bounds__(256)
biased_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ bias, const float* __restrict__ rowscale,
                        int T, int heads, float* __restrict__ ctx, int round_out) {
  extern __shared__ float ba_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= T) return;  // warp-uniform; only warp-level synchronisation below
  float* qs = ba_sm + warp * 64;
  float* sc = ba_sm + 8 * 64 + (long long)warp * T;
  const int D = heads * 64, ld = 3 * D;
  const long long tok = (long long)b * T + i;
  const float* base = qkv + (long long)b * T * ld + h * 64;
  qs[lane] = qkv[tok * ld + h * 64 + lane] * 0.125f;          // q scaled before the product, as
  qs[lane + 32] = qkv[tok * ld + h * 64 + lane + 32] * 0.125f;  // F.multi_head_attention_forward does
  __syncwarp();
  const float rs = rowscale ? rowscale[tok * heads + h] : 1.0f;
  const float* brow = bias + ((long long)h * T + i) * T;
  float mx = -INFINITY;
  for (int j = lane; j < T; j += 32) {
    const float4* kr = reinterpret_cast<const float4*>(base + (long long)j * ld + D);
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const float4 kv = __ldg(kr + d);
      a = fmaf(qs[4 * d], kv.x, a);
      a = fmaf(qs[4 * d + 1], kv.y, a);
      a = fmaf(qs[4 * d + 2], kv.z, a);
      a = fmaf(qs[4 * d + 3], kv.w, a);
    }
    a = fmaf(rs, __ldg(brow + j), a);
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float p = expf(sc[j] - mx);
    sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float a0 = 0.f, a1 = 0.f;
  const float* vb = base + 2 * D;
  for (int j = 0; j < T; ++j) {
    const float p = sc[j];
    a0 = fmaf(p, __ldg(vb + (long long)j * ld + lane), a0);
    a1 = fmaf(p, __ldg(vb + (long long)j * ld + lane + 32), a1);
  }
  const float inv = 1.0f / sum;
  a0 *= inv;
  a1 *= inv;
  if (round_out) {
    a0 = round_tf32(a0);
    a1 = round_tf32(a1);
  }
  ctx[tok * D + h * 64 + lane] = a0;
  ctx[tok * D + h * 64 + lane + 32] = a1;
}

}  // namespace

extern "C" {

int mer_wavlm_gate(const float* x, long long tokens, int heads, const float* w, const float* b, const float* c,
                   float* gate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MER_REQUIRE(x && w && b && c && gate && tokens > 0 && heads > 0, "mer_wavlm_gate: bad arguments");
  const long long warps = tokens * heads;
  wavlm_gate_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(x, tokens, heads, w, b, c, gate);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}

int mer_biased_attention(const float* qkv, const float* bias, const float* rowscale, int batch, int T, int heads,
                         float* ctx, int round_tf32_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MER_REQUIRE(qkv && bias && ctx && batch > 0 && batch <= 65535 && heads > 0 && heads <= 65535 && T > 0 && T <= 1024,
              "mer_biased_attention: batch %d, heads %d, T %d (<= 1024)", batch, heads, T);
  MER_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "mer_biased_attention: qkv must be 16-byte aligned");
  const size_t smem = (size_t)(8 * 64 + 8 * T) * sizeof(float);  // <= 34 KB
  biased_attention_kernel<<<dim3((T + 7) / 8, heads, batch), 256, smem, stream>>>(qkv, bias, rowscale, T, heads, ctx,
                                                                                 round_tf32_out);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}

}  // extern "C"
