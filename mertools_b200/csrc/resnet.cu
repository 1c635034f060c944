// resnet.cu — ResNet-18 frame encoder of the reference's ImageNet CNN extractor
// (MERBench/feature_extraction/visual/extract_imagenet_embedding.py:47-55: torchvision resnet18 without its
// fc layer on Resize(224) / ToTensor / Normalize(ImageNet) frames -> one 512-vector per frame).
//
// Every convolution (BatchNorm folded into weight and bias at load time) is an im2col gather into an fp16
// operand followed by the shared tcgen05 GEMM (MER_GEMM_F16) with its epilogue doing bias (+ identity)
// + ReLU; activations stay NHWC fp32 between layers (the GEMM's residual input is fp32).  64-channel layers
// are stored with 128 channels (upper half zero) because the GEMM's narrowest column block is 128; the gather
// reads only the real channels, so K is not inflated.  Max-pool and the im2col gathers are plain coalesced
// kernels (HBM-bound); the global average pool is the shared segment reduce.
// Per frame: 1.82 GFLOP algorithmic (2.4 executed with the 64 -> 128 column padding).
#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

// conv1 gather: uint8 BGR frames [n, 224, 224, 3] -> rows (n, oy, ox) of K = 7*7*3 (ky, kx, c) fp16 values
// ((pix/255 - mean[c]) / std[c], RGB order: ToTensor + Normalize), zero outside the image and for k >= 147.
__global__ void __launch_bounds__(256)
im2col_stem_kernel(const uint8_t* __restrict__ frames, int H, int W, int OH, int OW, int kpad, float m0, float m1,
                   float m2, float s0, float s1, float s2, uint16_t* __restrict__ out, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % kpad);
  const long long row = idx / kpad;
  const int ox = (int)(row % OW), oy = (int)((row / OW) % OH);
  const long long n = row / ((long long)OW * OH);
  float v = 0.f;
  if (k < 147) {
    const int c = k % 3, kx = (k / 3) % 7, ky = k / 21;
    const int iy = oy * 2 - 3 + ky, ix = ox * 2 - 3 + kx;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const float pix = (float)frames[((n * H + iy) * W + ix) * 3 + (2 - c)];
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
      v = (pix * 0.00392156862745098f - mean) / sd;
    }
  }
  out[idx] = (uint16_t)(pack_f16x2(v, 0.f) & 0xffffu);
}

// generic gather: NHWC fp32 activations [n, H, W, cs] (first C channels real) -> fp16 rows (n, oy, ox) of
// K = k*k*C in (ky, kx, c) order; one thread = 8 consecutive channels (16 bytes out).
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ x, int H, int W, int cs, int C, int ksz, int stride, int pad, int OH, int OW,
              uint4* __restrict__ out, long long total8) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total8) return;
  const int c8 = C / 8;
  const int kc = ksz * ksz * c8;
  const int q = (int)(idx % kc);
  const long long row = idx / kc;
  const int cg = q % c8, kx = (q / c8) % ksz, ky = q / (c8 * ksz);
  const int ox = (int)(row % OW), oy = (int)((row / OW) % OH);
  const long long n = row / ((long long)OW * OH);
  const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
    const float4* src = reinterpret_cast<const float4*>(x + ((n * H + iy) * W + ix) * (long long)cs + cg * 8);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    o = make_uint4(pack_f16x2(a.x, a.y), pack_f16x2(a.z, a.w), pack_f16x2(b.x, b.y), pack_f16x2(b.z, b.w));
  }
  out[idx] = o;
}

// MaxPool2d(3, stride 2, padding 1) on NHWC fp32, 4 channels per thread
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const float4* __restrict__ x, int H, int W, int c4, int OH, int OW, float4* __restrict__ y,
                    long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % c4);
  const long long pos = idx / c4;
  const int ox = (int)(pos % OW), oy = (int)((pos / OW) % OH);
  const long long n = pos / ((long long)OW * OH);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= W) continue;
      const float4 v = __ldg(x + ((n * H + iy) * W + ix) * (long long)c4 + c);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  y[idx] = m;
}

struct Shape { int H, W, C, Cs; };  // C real channels, Cs stored channels (>= 128)

int conv(const MerResnetConv& cv, const float* x, Shape in, int n, uint16_t* col, const float* res, bool relu,
         float* y, Shape* out, cudaStream_t st) {
  const int OH = (in.H + 2 * cv.pad - cv.k) / cv.stride + 1, OW = (in.W + 2 * cv.pad - cv.k) / cv.stride + 1;
  const long long rows = (long long)n * OH * OW;
  const int K = cv.k * cv.k * in.C;
  MER_REQUIRE(cv.cin == in.C && cv.kpad == K && K % 64 == 0 && in.C % 8 == 0, "resnet conv: geometry (cin %d K %d)",
              cv.cin, K);
  const long long total8 = rows * (K / 8);
  im2col_kernel<<<(unsigned)((total8 + 255) / 256), 256, 0, st>>>(x, in.H, in.W, in.Cs, in.C, cv.k, cv.stride, cv.pad,
                                                                  OH, OW, reinterpret_cast<uint4*>(col), total8);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  MerGemmDesc g;
  memset(&g, 0, sizeof(g));
  g.A = reinterpret_cast<const float*>(col);
  g.W = static_cast<const float*>(cv.w);
  g.rows_per_batch = (int)rows;
  g.a_rows_dim = (int)rows;
  g.batches = 1;
  g.N = cv.cout_pad;
  g.K_inner = K;
  g.taps = 1;
  g.P = 1;
  g.a_phase_stride = K;
  g.a_row_stride = K;
  g.a_batch_stride = rows * K;
  g.mode = MER_GEMM_F16;
  g.ep.bias = cv.b;
  g.ep.res = res;
  g.ep.out = y;
  g.ep.ld_out = cv.cout_pad;
  g.ep.ld_res = cv.cout_pad;
  g.ep.flags = relu ? MER_EPI_RELU : 0;
  if (int rc = mer_gemm_launch(&g, st)) return rc;
  *out = Shape{OH, OW, cv.cout, cv.cout_pad};
  return 0;
}

}  // namespace

namespace {
struct ResnetPlan { long long off_a0, off_p, off_q, off_r, off_col, off_cu, total; };
ResnetPlan resnet_plan(int n_frames) {
  const long long n = n_frames;
  auto al = [](long long x) { return (x + 255) & ~255ll; };
  ResnetPlan p;
  long long o = 0;
  p.off_a0 = o;  o += al(n * 112 * 112 * 128 * 4);   // conv1 output (64 real channels stored as 128)
  p.off_p = o;   o += al(n * 56 * 56 * 128 * 4);     // residual stream
  p.off_q = o;   o += al(n * 56 * 56 * 128 * 4);     // block-internal activation
  p.off_r = o;   o += al(n * 28 * 28 * 128 * 4);     // downsample branch (first needed at 28 x 28)
  const long long c1 = n * 112 * 112 * 192 * 2, c2 = n * 56 * 56 * 576 * 2;
  p.off_col = o; o += al(c1 > c2 ? c1 : c2);         // fp16 im2col operand
  p.off_cu = o;  o += al((n + 1) * 4);
  p.total = o;
  return p;
}
}  // namespace

extern "C" {

long long mer_resnet18_workspace_bytes(int n_frames) { return resnet_plan(n_frames).total; }

int mer_resnet18_forward(const MerResnet18Model* m, const uint8_t* frames_bgr, int n_frames, void* workspace,
                         long long workspace_bytes, float* out_feats, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  MER_REQUIRE(m && frames_bgr && workspace && out_feats && n_frames > 0, "mer_resnet18_forward: bad operands");
  const ResnetPlan p = resnet_plan(n_frames);
  MER_REQUIRE(workspace_bytes >= p.total, "mer_resnet18_forward: workspace %lld B < required %lld B", workspace_bytes,
              p.total);
  MER_REQUIRE((long long)n_frames * 112 * 112 < (1ll << 31), "mer_resnet18_forward: too many frames per call");
  const long long n = n_frames;
  char* ws = static_cast<char*>(workspace);
  float* a0 = reinterpret_cast<float*>(ws + p.off_a0);
  float* x = reinterpret_cast<float*>(ws + p.off_p);
  float* t1 = reinterpret_cast<float*>(ws + p.off_q);
  float* t2 = reinterpret_cast<float*>(ws + p.off_r);
  uint16_t* col = reinterpret_cast<uint16_t*>(ws + p.off_col);
  int* offsets = reinterpret_cast<int*>(ws + p.off_cu);

  // stem: conv1 7x7/2 (+BN folded) + ReLU
  {
    const MerResnetConv& cv = m->convs[0];
    MER_REQUIRE(cv.k == 7 && cv.kpad == 192 && cv.cout_pad == 128, "mer_resnet18_forward: conv1 packing");
    const long long rows = n * 112 * 112, total = rows * 192;
    im2col_stem_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(frames_bgr, 224, 224, 112, 112, 192,
                                                                        m->mean[0], m->mean[1], m->mean[2], m->std[0],
                                                                        m->std[1], m->std[2], col, total);
    MER_CUDA_CHECK(cudaGetLastError());
    mer_count_launches(1);
    MerGemmDesc g;
    memset(&g, 0, sizeof(g));
    g.A = reinterpret_cast<const float*>(col);
    g.W = static_cast<const float*>(cv.w);
    g.rows_per_batch = (int)rows;
    g.a_rows_dim = (int)rows;
    g.batches = 1;
    g.N = 128;
    g.K_inner = 192;
    g.taps = 1;
    g.P = 1;
    g.a_phase_stride = 192;
    g.a_row_stride = 192;
    g.a_batch_stride = rows * 192;
    g.mode = MER_GEMM_F16;
    g.ep.bias = cv.b;
    g.ep.out = a0;
    g.ep.ld_out = 128;
    g.ep.flags = MER_EPI_RELU;
    if (int rc = mer_gemm_launch(&g, st)) return rc;
  }
  // maxpool 3x3/2 -> x [n,56,56,128]
  {
    const long long total = n * 56 * 56 * 32;
    maxpool3x3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(a0), 112, 112, 32,
                                                                         56, 56, reinterpret_cast<float4*>(x), total);
    MER_CUDA_CHECK(cudaGetLastError());
    mer_count_launches(1);
  }
  // four stages of two BasicBlocks (torchvision resnet.py: conv-bn-relu, conv-bn, (+downsample), add, relu).
  // The block output overwrites the stream buffer: its residual is either the stream itself (read and written
  // by the same epilogue thread) or the downsample branch in t2.
  Shape s{56, 56, 64, 128};
  int ci = 1;
  for (int stage = 0; stage < 4; ++stage) {
    for (int blk = 0; blk < 2; ++blk) {
      const bool down = stage > 0 && blk == 0;
      Shape s1, s2, sd;
      if (int rc = conv(m->convs[ci], x, s, n_frames, col, nullptr, true, t1, &s1, st)) return rc;
      const float* identity = x;
      if (down) {
        if (int rc = conv(m->convs[ci + 2], x, s, n_frames, col, nullptr, false, t2, &sd, st)) return rc;
        identity = t2;
      }
      if (int rc = conv(m->convs[ci + 1], t1, s1, n_frames, col, identity, true, x, &s2, st)) return rc;
      s = s2;
      ci += down ? 3 : 2;
    }
  }
  // global average pool over the 7 x 7 positions
  MER_REQUIRE(s.H == 7 && s.W == 7 && s.C == 512 && s.Cs == 512, "mer_resnet18_forward: unexpected final shape");
  if (int rc = mer_iota_offsets_launch(offsets, n_frames, 49, st)) return rc;
  return mer_segment_reduce_launch(x, offsets, offsets + 1, n_frames, 512, MER_SEG_MEAN, out_feats, st);
}

}  // extern "C"
