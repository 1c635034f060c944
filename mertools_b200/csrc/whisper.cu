// whisper.cu — the two pieces of the reference's Whisper branch (extract_audio_huggingface.py:83-91) that the shared
// GEMM / LayerNorm / attention kernels do not cover:
//
//   whisper_logmel   : WhisperFeatureExtractor (HF feature_extraction_whisper.py: 30 s of audio, reflect-padded STFT
//                      with a periodic Hann window of 400 samples and hop 160, power spectrum, 80 Slaney mel bands,
//                      log10 with floor 1e-10, last frame dropped, clamp to (clip max - 8), (x + 4) / 4) written
//                      TIME-MAJOR [B, 3000, ld] as the operand of the first convolution (a 3-tap GEMM).
//   small_attention  : softmax(q k^T / 8) v for a handful of query rows per (clip, head) — the decoder's causal
//                      self-attention over its two start tokens and its cross-attention over the 1500 encoder frames.
//
// Direct 400-point DFT through a cos / sin table (the size is not a power of two; 0.16 MFLOP per frame, 0.5 GFLOP per
// clip): ALU work, negligible next to the encoder.
#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

constexpr int WH_NFFT = 400, WH_HOP = 160, WH_BINS = 201, WH_MELS = 80, WH_FRAMES = 3000, WH_SAMPLES = 480000;

// one block per (frame, clip): windowed frame -> 201 power bins -> 80 log10 mel values; per-clip maximum by atomics
__global__ void __launch_bounds__(256)
whisper_logmel_kernel(const float* __restrict__ wave, long long ld_wave, const float* __restrict__ mel /*[201][80]*/,
                      float* __restrict__ out, int ld_out, int* __restrict__ clip_max_bits) {
  __shared__ float fr[WH_NFFT];
  __shared__ float cs[WH_NFFT], sn[WH_NFFT];
  __shared__ float spec[WH_BINS];
  const int t = blockIdx.x, b = blockIdx.y;
  const float* x = wave + (long long)b * ld_wave;
  for (int n = threadIdx.x; n < WH_NFFT; n += blockDim.x) {
    int i = t * WH_HOP + n - WH_NFFT / 2;                 // index into the un-padded 30 s signal
    if (i < 0) i = -i;                                    // numpy "reflect": the edge sample is not repeated
    if (i >= WH_SAMPLES) i = 2 * (WH_SAMPLES - 1) - i;
    float sv, cv;
    sincospif(2.0f * (float)n / (float)WH_NFFT, &sv, &cv);
    fr[n] = x[i] * (0.5f - 0.5f * cv);                    // periodic Hann
    cs[n] = cv;
    sn[n] = sv;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < WH_BINS; k += blockDim.x) {
    float re = 0.f, im = 0.f;
    int ph = 0;                                           // (k * n) mod 400
    for (int n = 0; n < WH_NFFT; ++n) {
      re = fmaf(fr[n], cs[ph], re);
      im = fmaf(fr[n], sn[ph], im);
      ph += k;
      if (ph >= WH_NFFT) ph -= WH_NFFT;
    }
    spec[k] = re * re + im * im;
  }
  __syncthreads();
  float local_max = -INFINITY;
  for (int m = threadIdx.x; m < WH_MELS; m += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < WH_BINS; ++k) a = fmaf(spec[k], __ldg(mel + k * WH_MELS + m), a);
    const float v = log10f(fmaxf(a, 1e-10f));
    out[((long long)b * WH_FRAMES + t) * ld_out + m] = v;
    local_max = fmaxf(local_max, v);
  }
  local_max = warp_max(local_max);
  if ((threadIdx.x & 31) == 0 && local_max > -INFINITY) {
    // order-preserving integer image of a float (values here are finite): atomicMax on it
    const int bits = __float_as_int(local_max);
    atomicMax(clip_max_bits + b, bits >= 0 ? bits : bits ^ 0x7fffffff);
  }
}

// (max(x, clip_max - 8) + 4) / 4, optionally TF32-rounded (GEMM operand); pad columns [80, ld) are zeroed
__global__ void __launch_bounds__(256)
whisper_logmel_finish_kernel(float* __restrict__ feat, int ld, const int* __restrict__ clip_max_bits, int round_tf32_out,
                             long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % ld);
  const long long b = idx / ((long long)ld * WH_FRAMES);
  float v = 0.f;
  if (c < WH_MELS) {
    const int bits = clip_max_bits[b];
    const float mx = __int_as_float(bits >= 0 ? bits : bits ^ 0x7fffffff);
    v = (fmaxf(feat[idx], mx - 8.0f) + 4.0f) * 0.25f;
    if (round_tf32_out) v = round_tf32(v);
  }
  feat[idx] = v;
}

// one block per (clip, head); nq <= 8 query rows, nk <= 1536 keys; q / k / v rows are [*, ld] with this head's 64
// columns at head * 64 (+ the caller's column offset, folded into the pointers)
constexpr int SA_MAXK = 1536;
__global__ void __launch_bounds__(128)
small_attention_kernel(const float* __restrict__ q, int ld_q, const float* __restrict__ k, int ld_k,
                       const float* __restrict__ v, int ld_v, int nq, int nk, int causal, float* __restrict__ out,
                       int ld_o) {
  __shared__ float sc[SA_MAXK];
  __shared__ float qs[64];
  __shared__ float red[4];
  const int b = blockIdx.y, h = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* kb = k + (long long)b * nk * ld_k + h * 64;
  const float* vb = v + (long long)b * nk * ld_v + h * 64;
  for (int i = 0; i < nq; ++i) {
    const int nvis = causal ? min(nk, i + 1) : nk;        // keys this query may see
    if (threadIdx.x < 64) qs[threadIdx.x] = q[((long long)b * nq + i) * ld_q + h * 64 + threadIdx.x];
    __syncthreads();
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < nvis; j += blockDim.x) {
      const float4* kr = reinterpret_cast<const float4*>(kb + (long long)j * ld_k);
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float4 kv = __ldg(kr + d);
        a = fmaf(qs[4 * d], kv.x, a);
        a = fmaf(qs[4 * d + 1], kv.y, a);
        a = fmaf(qs[4 * d + 2], kv.z, a);
        a = fmaf(qs[4 * d + 3], kv.w, a);
      }
      a *= 0.125f;
      sc[j] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int j = threadIdx.x; j < nvis; j += blockDim.x) {
      const float p = expf(sc[j] - mx);
      sc[j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = (red[0] + red[1]) + (red[2] + red[3]);
    // out[d] = sum_j p_j v_j[d]: two key halves per output column (threads 0..63 and 64..127), combined through smem
    const int d = threadIdx.x & 63, half = threadIdx.x >> 6;
    float acc = 0.f;
    for (int j = half; j < nvis; j += 2) acc = fmaf(sc[j], __ldg(vb + (long long)j * ld_v + d), acc);
    __syncthreads();                                      // everyone is done reading red[] and sc[] scores
    if (half == 1) qs[d] = acc;
    __syncthreads();
    if (half == 0) out[((long long)b * nq + i) * ld_o + h * 64 + d] = (acc + qs[d]) / sum;
    __syncthreads();
  }
}

}  // namespace

extern "C" {

int mer_whisper_logmel(const float* waves, int batch, long long ld_wave, const float* mel_filters, float* out,
                       int ld_out, int round_tf32_out, int* scratch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MER_REQUIRE(waves && mel_filters && out && scratch && batch > 0 && ld_wave >= WH_SAMPLES && ld_out >= WH_MELS,
              "mer_whisper_logmel: bad arguments (rows of 480000 samples, ld_out >= 80)");
  MER_CUDA_CHECK(cudaMemsetAsync(scratch, 0x80, (size_t)batch * sizeof(int), stream));  // 0x80808080: below the image of every value here (log10 >= -10)
  whisper_logmel_kernel<<<dim3(WH_FRAMES, batch), 256, 0, stream>>>(waves, ld_wave, mel_filters, out, ld_out, scratch);
  MER_CUDA_CHECK(cudaGetLastError());
  const long long total = (long long)batch * WH_FRAMES * ld_out;
  whisper_logmel_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(out, ld_out, scratch, round_tf32_out,
                                                                                    total);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(2);
  return 0;
}

int mer_small_attention(const float* q, int ld_q, const float* k, int ld_k, const float* v, int ld_v, int batch, int heads,
                        int nq, int nk, int causal, float* out, int ld_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MER_REQUIRE(q && k && v && out && batch > 0 && heads > 0 && nq > 0 && nq <= 8 && nk > 0 && nk <= SA_MAXK,
              "mer_small_attention: %d queries (<= 8), %d keys (<= %d)", nq, nk, SA_MAXK);
  MER_REQUIRE(ld_k % 4 == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0, "mer_small_attention: k rows must be 16-byte aligned");
  small_attention_kernel<<<dim3(heads, batch), 128, 0, stream>>>(q, ld_q, k, ld_k, v, ld_v, nq, nk, causal, out, ld_out);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}

}  // extern "C"
