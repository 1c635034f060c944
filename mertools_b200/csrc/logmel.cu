// logmel.cu — log mel spectrogram front-end of the reference's VGGish audio path
// (MERBench/feature_extraction/audio/vggish/mel_features.py:21-223 with the constants of
// vggish_params.py:22-34: 16 kHz, 25 ms periodic-Hann window = 400 samples, 10 ms hop = 160, FFT 512,
// 64 HTK mel bands over 125-7500 Hz, log(mel + 0.01)).
//
// One block = FRAMES_PER_BLOCK consecutive frames of one clip.  The 400 windowed samples of each frame sit
// in shared memory; thread k computes DFT bin k (k = 0..256) of every frame of the block as a direct sum
// over a 512-entry cos/sin table (fp32, |err| ~1e-6 of the magnitude: the reference is float64 numpy, the
// parity bar is 1e-3); the 257 magnitudes go back to shared memory and warp w reduces mel bands w, w+8, ...
// with shuffles over the band's non-zero bin range [lo, hi) (the HTK triangles overlap only pairwise).
// Algorithmic traffic: 4 B per input sample in (each sample is read by 2.5 frames, from L1/L2) and
// 64 x 4 B per frame out.
#include <math.h>

#include <vector>

#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

constexpr int WIN = 400, HOP = 160, NFFT = 512, NBIN = 257, NMEL = 64;
constexpr int FRAMES_PER_BLOCK = 8;
constexpr int LM_THREADS = 288;  // 9 warps: bins 0..256 on threads 0..256

__global__ void __launch_bounds__(LM_THREADS)
logmel_kernel(const float* __restrict__ wave, long long ld_wave, int n_frames, const float* __restrict__ window,
              const float* __restrict__ mel_w /*[NBIN][NMEL]*/, const int* __restrict__ mel_lo,
              const int* __restrict__ mel_hi, float log_offset, float* __restrict__ out /*[B, n_frames, NMEL]*/) {
  __shared__ float cs[NFFT], sn[NFFT];
  __shared__ float xs[FRAMES_PER_BLOCK][WIN];
  __shared__ float mag[FRAMES_PER_BLOCK][NBIN + 3];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FRAMES_PER_BLOCK;
  const int nf = min(FRAMES_PER_BLOCK, n_frames - f0);
  const int tid = threadIdx.x;
  for (int i = tid; i < NFFT; i += LM_THREADS) {
    float s, c;
    sincospif(2.0f * (float)i / (float)NFFT, &s, &c);
    cs[i] = c;
    sn[i] = s;
  }
  const float* x = wave + (long long)b * ld_wave + (long long)f0 * HOP;
  for (int i = tid; i < nf * WIN; i += LM_THREADS) {
    const int f = i / WIN, n = i - f * WIN;
    xs[f][n] = x[f * HOP + n] * window[n];
  }
  __syncthreads();
  if (tid < NBIN) {
    float re[FRAMES_PER_BLOCK], im[FRAMES_PER_BLOCK];
#pragma unroll
    for (int f = 0; f < FRAMES_PER_BLOCK; ++f) re[f] = im[f] = 0.f;
    int idx = 0;  // (tid * n) mod 512
    for (int n = 0; n < WIN; ++n) {
      const float c = cs[idx], s = sn[idx];
#pragma unroll
      for (int f = 0; f < FRAMES_PER_BLOCK; ++f) {
        const float v = xs[f][n];  // broadcast
        re[f] = fmaf(v, c, re[f]);
        im[f] = fmaf(v, s, im[f]);
      }
      idx = (idx + tid) & (NFFT - 1);
    }
#pragma unroll
    for (int f = 0; f < FRAMES_PER_BLOCK; ++f) mag[f][tid] = sqrtf(re[f] * re[f] + im[f] * im[f]);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int m = warp; m < NMEL; m += LM_THREADS / 32) {
    const int lo = mel_lo[m], hi = mel_hi[m];
    for (int f = 0; f < nf; ++f) {
      float acc = 0.f;
      for (int k = lo + lane; k < hi; k += 32) acc = fmaf(mag[f][k], __ldg(mel_w + k * NMEL + m), acc);
      acc = warp_sum(acc);
      if (lane == 0) out[((long long)b * n_frames + f0 + f) * NMEL + m] = logf(acc + log_offset);
    }
  }
}

struct MelTables { float* window; float* mel_w; int* lo; int* hi; };
MelTables g_mel = {nullptr, nullptr, nullptr, nullptr};

// periodic Hann (mel_features.py:48-69) and the HTK mel matrix (:96-164), in double as numpy does
int mel_tables(const MelTables** res, cudaStream_t stream) {
  if (g_mel.window) { *res = &g_mel; return 0; }
  std::vector<float> win(WIN), w((size_t)NBIN * NMEL);
  std::vector<int> lo(NMEL), hi(NMEL);
  const double pi = 3.14159265358979323846;
  for (int n = 0; n < WIN; ++n) win[n] = (float)(0.5 - 0.5 * cos(2.0 * pi / WIN * n));
  const double nyquist = 16000.0 / 2.0, lower = 125.0, upper = 7500.0;
  auto h2m = [](double hz) { return 1127.0 * log(1.0 + hz / 700.0); };
  std::vector<double> bins_mel(NBIN), edges(NMEL + 2);
  for (int k = 0; k < NBIN; ++k) bins_mel[k] = h2m(nyquist * k / (NBIN - 1));  // np.linspace(0, nyquist, NBIN)
  const double e0 = h2m(lower), e1 = h2m(upper);
  for (int i = 0; i < NMEL + 2; ++i) edges[i] = e0 + (e1 - e0) * i / (NMEL + 1);
  for (int m = 0; m < NMEL; ++m) {
    const double lo_e = edges[m], ce = edges[m + 1], up = edges[m + 2];
    lo[m] = NBIN;
    hi[m] = 0;
    for (int k = 0; k < NBIN; ++k) {
      const double ls = (bins_mel[k] - lo_e) / (ce - lo_e), us = (up - bins_mel[k]) / (up - ce);
      double v = ls < us ? ls : us;
      if (v < 0.0) v = 0.0;
      if (k == 0) v = 0.0;  // mel_weights_matrix[0, :] = 0
      w[(size_t)k * NMEL + m] = (float)v;
      if (v > 0.0) {
        if (k < lo[m]) lo[m] = k;
        if (k + 1 > hi[m]) hi[m] = k + 1;
      }
    }
    if (lo[m] > hi[m]) lo[m] = hi[m] = 0;
  }
  MelTables t;
  MER_CUDA_CHECK(cudaMalloc(&t.window, WIN * sizeof(float)));
  MER_CUDA_CHECK(cudaMalloc(&t.mel_w, w.size() * sizeof(float)));
  MER_CUDA_CHECK(cudaMalloc(&t.lo, NMEL * sizeof(int)));
  MER_CUDA_CHECK(cudaMalloc(&t.hi, NMEL * sizeof(int)));
  MER_CUDA_CHECK(cudaMemcpyAsync(t.window, win.data(), WIN * sizeof(float), cudaMemcpyHostToDevice, stream));
  MER_CUDA_CHECK(cudaMemcpyAsync(t.mel_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
  MER_CUDA_CHECK(cudaMemcpyAsync(t.lo, lo.data(), NMEL * sizeof(int), cudaMemcpyHostToDevice, stream));
  MER_CUDA_CHECK(cudaMemcpyAsync(t.hi, hi.data(), NMEL * sizeof(int), cudaMemcpyHostToDevice, stream));
  MER_CUDA_CHECK(cudaStreamSynchronize(stream));
  g_mel = t;
  *res = &g_mel;
  return 0;
}

}  // namespace

extern "C" int mer_logmel_num_frames(int n_samples) {
  return n_samples < WIN ? 0 : 1 + (n_samples - WIN) / HOP;
}

extern "C" int mer_logmel(const float* wave, int batch, int n_samples, long long ld_wave, float* out,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MER_REQUIRE(wave && out && batch > 0 && n_samples >= WIN && ld_wave >= n_samples, "mer_logmel: bad arguments");
  const int n_frames = mer_logmel_num_frames(n_samples);
  const MelTables* t;
  if (int rc = mel_tables(&t, stream)) return rc;
  dim3 grid((n_frames + FRAMES_PER_BLOCK - 1) / FRAMES_PER_BLOCK, batch);
  logmel_kernel<<<grid, LM_THREADS, 0, stream>>>(wave, ld_wave, n_frames, t->window, t->mel_w, t->lo, t->hi, 0.01f,
                                                 out);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}
