// posconv.cu — HubertPositionalConvEmbedding (HF modeling_hubert.py:45-103) fused with the
// residual add:   x1[b,t,:] = x0[b,t,:] + GELU( bias + GroupedConv1d_{k=128,pad=64,g=16}(x0)[b,t,:] )
// (the conv yields T+1 frames; HubertSamePadLayer drops the last, so output t reads input frames
// t-64 .. t+63 with zero padding at both ends of EACH sequence).  The weight-norm is folded on the
// host at load time (mertools_b200/weights.py); weights arrive as Wp[g][tap][o][c] (tf32-rounded).
//
// One CTA = (sequence b, group g, 64 output frames): the 191 x 48 input window sits in shared
// memory once and serves all 128 taps (Toeplitz reuse); weights stream through smem, 2 taps per chunk
// (double-buffered cp.async); math is mma.sync m16n8k8 TF32, fp32 accumulate.
// FLOPs per 5 s clip: 2 * 249 * 768 * 48 * 128 = 2.35 GF.  [round 1: legacy tensor path]
#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

constexpr int GC = 48;       // channels per group (in and out)
constexpr int NG = 16;       // groups
constexpr int KT = 128;      // taps
constexpr int PADL = 64;     // left zero padding
constexpr int BT = 64;       // output frames per CTA
constexpr int WIN = BT + KT - 1;  // 191 input frames
constexpr int LDX = 52;      // padded pitch of the window rows (floats)
constexpr int LDW = 52;      // padded pitch of weight rows
constexpr int TAPC = 2;      // taps per weight chunk (80 KB smem total -> 2 CTAs per SM)
constexpr int WCHUNK = TAPC * GC * LDW;  // floats per chunk buffer
constexpr int PC_THREADS = 128;
constexpr int PC_SMEM = (WIN * LDX + 2 * WCHUNK) * 4;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(PC_THREADS)
posconv_kernel(const float* __restrict__ x0, const float* __restrict__ wp,
               const float* __restrict__ bias, const int* __restrict__ cu_seqlens,
               float* __restrict__ x1) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                 // [WIN][LDX]
  float* ws = sm + WIN * LDX;     // [2][TAPC][GC][LDW]

  const int seq = blockIdx.z;
  const int grp = blockIdx.y;
  const int start = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - start;
  const int t0 = blockIdx.x * BT;
  if (t0 >= len) return;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  const float* wg = wp + (long long)grp * KT * GC * GC;
  auto load_w = [&](int chunk, int buf) {
    // TAPC*48 rows of 48 floats = 12 x 16B per row
    const float* src = wg + (long long)chunk * TAPC * GC * GC;
    float* dst = ws + buf * WCHUNK;
    for (int i = tid; i < TAPC * GC * 12; i += PC_THREADS) {
      const int row = i / 12, c4 = i % 12;
      cp_async16(dst + row * LDW + c4 * 4, src + row * GC + c4 * 4);
    }
    cp_async_commit();
  };
  load_w(0, 0);

  // input window: frames t0-64 .. t0+126 of this sequence, this group's 48 channels (tf32-rounded)
  for (int i = tid; i < WIN * 12; i += PC_THREADS) {
    const int r = i / 12, c4 = i % 12;
    const int tt = t0 - PADL + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt >= 0 && tt < len)
      v = *reinterpret_cast<const float4*>(x0 + (long long)(start + tt) * 768 + grp * GC + c4 * 4);
    v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
    *reinterpret_cast<float4*>(xs + r * LDX + c4 * 4) = v;
  }

  float acc[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  const int row_lo = warp * 16 + g;  // output frame (within the tile) of fragment rows g / g+8
  constexpr int NCHUNK = KT / TAPC;
  for (int ch = 0; ch < NCHUNK; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < NCHUNK) {
      load_w(ch + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wt = ws + buf * WCHUNK;
#pragma unroll 2
    for (int j = 0; j < TAPC; ++j) {
      const int tap = ch * TAPC + j;
      const float* xr_lo = xs + (row_lo + tap) * LDX + t;
      const float* xr_hi = xr_lo + 8 * LDX;
      const float* wj = wt + j * GC * LDW;
#pragma unroll
      for (int ks = 0; ks < GC / 8; ++ks) {
        uint32_t a[4];
        a[0] = __float_as_uint(xr_lo[ks * 8]);
        a[1] = __float_as_uint(xr_hi[ks * 8]);
        a[2] = __float_as_uint(xr_lo[ks * 8 + 4]);
        a[3] = __float_as_uint(xr_hi[ks * 8 + 4]);
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) {
          const float* wr = wj + (nt * 8 + g) * LDW + ks * 8 + t;
          mma_tf32(acc[nt], a, __float_as_uint(wr[0]), __float_as_uint(wr[4]));
        }
      }
    }
    __syncthreads();
  }

  // epilogue: + bias, GELU, + residual x0 (exact fp32 from global), write x1
  const int f_lo = t0 + row_lo, f_hi = f_lo + 8;
#pragma unroll
  for (int nt = 0; nt < 6; ++nt) {
    const int col = grp * GC + nt * 8 + 2 * t;
    const float b0 = __ldg(bias + col), b1 = __ldg(bias + col + 1);
    if (f_lo < len) {
      const long long off = (long long)(start + f_lo) * 768 + col;
      const float2 r = *reinterpret_cast<const float2*>(x0 + off);
      float2 o = make_float2(r.x + gelu_erf_fast(acc[nt][0] + b0), r.y + gelu_erf_fast(acc[nt][1] + b1));
      *reinterpret_cast<float2*>(x1 + off) = o;
    }
    if (f_hi < len) {
      const long long off = (long long)(start + f_hi) * 768 + col;
      const float2 r = *reinterpret_cast<const float2*>(x0 + off);
      float2 o = make_float2(r.x + gelu_erf_fast(acc[nt][2] + b0), r.y + gelu_erf_fast(acc[nt][3] + b1));
      *reinterpret_cast<float2*>(x1 + off) = o;
    }
  }
}

}  // namespace

int mer_posconv_launch(const float* x0, const float* wp, const float* bias, const int* cu_seqlens,
                       int n_seq, int max_seqlen, float* x1, cudaStream_t stream) {
  MER_REQUIRE(x0 && wp && bias && cu_seqlens && x1 && x0 != x1, "mer_posconv: bad operands");
  if (n_seq <= 0 || max_seqlen <= 0) return 0;
  static MerPerDevice attr_set;
  if (attr_set.needs_setup()) {
    MER_CUDA_CHECK(cudaFuncSetAttribute(posconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        PC_SMEM));
    attr_set.mark();
  }
  dim3 grid((max_seqlen + BT - 1) / BT, NG, n_seq);
  // 2 * frames * 768 outputs * (48 inputs * 128 taps); frames bounded by n_seq * max_seqlen (exact for
  // equal-length batches)
  const int prof = mer_prof_begin(MER_PROF_POSCONV, 2.0 * (double)n_seq * max_seqlen * 768.0 * 48.0 * 128.0, stream);
  posconv_kernel<<<grid, PC_THREADS, PC_SMEM, stream>>>(x0, wp, bias, cu_seqlens, x1);
  mer_prof_end(prof, stream);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}
