// attention_f16_long.cu — tcgen05 attention on fp16 operands for sequences of 250 .. 505 tokens (head_dim 64).
//
// The rows attention_f16.cu cannot take: audio rows of 5 .. 10 s (250 .. 499 HuBERT frames, the reference's
// split_into_batch rows, extract_audio_huggingface.py:40-50) and CLIP L/14's 257 tokens.  Until round 2 these fell to the
// mma.sync flash kernel of attention.cu.  Same operands as attention_f16.cu (q | k fp16 rows and V^T fp16 from the QKV
// GEMM epilogue; ctx in the operand format of the out-proj GEMM: fp16, tf32-rounded fp32 or bf16 hi | lo split rows) and the same reference op (HF eager / sdpa attention,
// modeling_hubert.py:372-405, modeling_clip.py attention).
//
// Persistent, one CTA per SM, work item = (sequence, head).  K (<= 512 keys x 128 B) and V^T (8 chunks of 64 keys) of the
// item stay resident in shared memory while its query tiles (128 rows each, up to 4) stream through two Q buffers:
//   warp 0       TMA producer: K and V^T once per item (refilled as soon as the last S = Q K^T / P V product of the
//                previous item has completed), one 128-row Q tile per tile
//   warp 1       tcgen05 issuer: S = Q_t K^T as one or two UMMAs of N <= 256 per 16 head dims into TMEM columns [0, NK)
//                (all 512 columns at 499 frames); O_t = P V with P read from tensor memory (tcgen05.mma, A in TMEM)
//   warps 2..17  softmax + epilogue: FOUR warps share each 32-row quarter of the tile, each owning a contiguous range
//                of 16-key steps: range 0 = steps [0, 8), the rest in three equal parts.  Row maxima / sums meet
//                through shared memory under a 128-thread named barrier.  Each warp turns its scores into fp16
//                probabilities 16 keys at a time and writes them with tcgen05.st over the START of its own range (step j
//                of a range at column c lands in [c + 8 j, c + 8 j + 8), columns that thread has already consumed).
//                O_t accumulates in columns [64, 128): the second half of range 0's scores, dead once every warp has
//                arrived, and the reason range 0 is pinned to 8 steps.  Epilogue: each warp takes 16 head dims of its
//                32 rows (O / sum -> fp16 -> one 32-byte sector per row of ctx).
// One tile is in flight per SM (its S needs the whole tensor memory), so the chain S -> max -> exp -> P V -> O is
// serial; what the long key axis buys back is amortisation: per tile 2 x 1k cycles of UMMA against ~4k of exponentials.
// TMA boxes start on 16-byte boundaries: the key axis begins at the sequence start rounded down to a multiple of 8
// tokens; the (up to 7) leading foreign keys and the tail beyond the sequence are masked.
// Algorithmic HBM traffic per token and layer: 4.5 KB of q | k | v^T in, 1.5 KB of ctx out (K / V^T re-reads of the other
// heads' CTAs hit L2).
#include <stdlib.h>

#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

constexpr int HD = 64;
constexpr int LONG_THREADS = 576;          // producer, MMA issuer, 16 softmax / epilogue warps
constexpr int MAX_KEYS = 512;              // shifted key axis, padded to the UMMA step
constexpr int K_BYTES = MAX_KEYS * 128;    // 64 KB: one 128-byte swizzle row per key
constexpr int VT_CHUNK = HD * 128;         // V^T chunk: 64 d-rows x 64 keys
constexpr int V_BYTES = 8 * VT_CHUNK;      // 64 KB
constexpr int QTILE_BYTES = 128 * 128;
constexpr int SMEM_K = 0;
constexpr int SMEM_V = K_BYTES;
constexpr int SMEM_Q = SMEM_V + V_BYTES;               // two Q tile buffers
constexpr int SMEM_BAR = SMEM_Q + 2 * QTILE_BYTES;     // 160 KB
constexpr int SMEM_XCHG = SMEM_BAR + 256;              // [128 rows][4 ranges] row max, then the same for the row sums
constexpr int LONG_SMEM = SMEM_XCHG + 2 * 128 * 4 * 4 + 1024;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t O_COL = 64;
constexpr int R0_STEPS = 8;  // 16-key steps of range 0: its scores cover columns [0, 128), its P [0, 64), O [64, 128)

__device__ __forceinline__ float fast_ex2(float x) {  // MUFU.EX2, flush-to-zero
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {  // SW128, SBO 1024
  return static_cast<uint64_t>((addr & 0x3FFFF) >> 4) | (1ull << 16) | (uint64_t(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

// the 16-key steps [k_first, k_first + nq) of key range r (0..3) when the item has nks steps; MMA and softmax warps
// must agree on this
__device__ __forceinline__ void key_range(int r, int nks, int& k_first, int& nq) {
  const int n0 = nks < R0_STEPS ? nks : R0_STEPS;
  if (r == 0) {
    k_first = 0;
    nq = n0;
    return;
  }
  const int rest = nks - n0, per = (rest + 2) / 3;
  k_first = n0 + (r - 1) * per;
  int n = rest - (r - 1) * per;
  nq = n < 0 ? 0 : (n > per ? per : n);
  if (nq == 0) k_first = n0;  // empty range: any in-bounds column
}

// POLY: of the 8 exponential pairs per 16-key step, how many run as a polynomial on the FMA pipe (attention_f16.cu)
template <int POLY>
__global__ void __launch_bounds__(LONG_THREADS, 1)
attention_f16_long_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_vt,
                          void* __restrict__ ctx_, const int* __restrict__ cu_seqlens, int n_seq, int heads,
                          int out_mode) {
  // out_mode: the operand format of the out-proj GEMM that reads ctx: 3 = fp16, 2 = bf16 hi | lo split rows (BF16X3),
  // 1 = tf32-rounded fp32, 0 = fp32
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* bar_k = bars + 0;       // producer -> MMA: K of the item
  uint64_t* bar_v = bars + 1;       // producer -> MMA: V^T of the item
  uint64_t* bar_q = bars + 2;       // [2] producer -> MMA: Q tile buffer
  uint64_t* bar_kfree = bars + 4;   // MMA -> producer: the item's last S product has read K
  uint64_t* bar_vfree = bars + 5;   // MMA -> producer: the item's last P V product has read V^T
  uint64_t* bar_qfree = bars + 6;   // [2] MMA -> producer: the S product of the tile has read the Q buffer
  uint64_t* bar_sfull = bars + 8;   // MMA -> softmax: S complete
  uint64_t* bar_pready = bars + 9;  // softmax (16 warps) -> MMA: P sits in tensor memory
  uint64_t* bar_ofull = bars + 10;  // MMA -> softmax: O complete
  uint64_t* bar_otfree = bars + 11; // softmax (16 warps) -> MMA: O has been read out of tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_seq * heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_vt);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_kfree, 1);
    mbar_init(bar_vfree, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_q[b], 1);
      mbar_init(&bar_qfree[b], 1);
    }
    mbar_init(bar_sfull, 1);
    mbar_init(bar_pready, 16);
    mbar_init(bar_ofull, 1);
    mbar_init(bar_otfree, 16);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform; one elected lane issues) =====================
    uint32_t item_n = 0, g = 0;  // items and tiles this CTA has started
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++item_n) {
      const int seq = it / heads, h = it % heads;
      const int start = cu_seqlens[seq];
      const int len = cu_seqlens[seq + 1] - start;
      const int n_mt = (len + 127) >> 7;
      const int a_start = start & ~7;
      const int Lk = (start - a_start) + len;
      const int nb = (Lk + 127) >> 7, n_vc = (Lk + 63) >> 6;
      auto load_q = [&](int t) {
        const uint32_t gt = g + t, b = gt & 1, use = gt >> 1;
        if (use > 0) mbar_wait(&bar_qfree[b], (use - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(&bar_q[b], QTILE_BYTES);
          tma_load_2d(smem + SMEM_Q + b * QTILE_BYTES, &tmap_qkv, &bar_q[b], h * HD, start + t * 128);
        }
        __syncwarp();
      };
      if (item_n > 0) mbar_wait(bar_kfree, (item_n - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_k, (uint32_t)(nb * 16384));
        for (int b = 0; b < nb; ++b)
          tma_load_2d(smem + SMEM_K + b * 16384, &tmap_qkv, bar_k, heads * HD + h * HD, a_start + b * 128);
      }
      __syncwarp();
      for (int t = 0; t < n_mt && t < 2; ++t) load_q(t);
      if (item_n > 0) mbar_wait(bar_vfree, (item_n - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_v, (uint32_t)(n_vc * VT_CHUNK));
        for (int c = 0; c < n_vc; ++c)
          tma_load_2d(smem + SMEM_V + c * VT_CHUNK, &tmap_vt, bar_v, a_start + c * 64, h * HD);
      }
      __syncwarp();
      for (int t = 2; t < n_mt; ++t) load_q(t);
      g += n_mt;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform; one elected lane issues and commits) =====================
    const uint64_t desc_k = desc_kmajor(smem_u32(smem + SMEM_K));
    const uint64_t desc_v = desc_kmajor(smem_u32(smem + SMEM_V));
    const uint64_t desc_q0 = desc_kmajor(smem_u32(smem + SMEM_Q));
    const uint32_t idesc_o = umma_idesc(0, 128, HD);
    uint32_t item_n = 0, g = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++item_n) {
      const int seq = it / heads;
      const int start = cu_seqlens[seq];
      const int len = cu_seqlens[seq + 1] - start;
      const int n_mt = (len + 127) >> 7;
      const int NK = ((start & 7) + len + 15) & ~15;
      const int nks = NK >> 4;
      const int N1 = NK < 256 ? NK : 256, N2 = NK - N1;
      const uint32_t idesc_s1 = umma_idesc(0, 128, N1), idesc_s2 = umma_idesc(0, 128, N2 > 0 ? N2 : 16);
      mbar_wait(bar_k, item_n & 1);
      for (int t = 0; t < n_mt; ++t, ++g) {
        const uint32_t b = g & 1;
        mbar_wait(&bar_q[b], (g >> 1) & 1);
        if (g > 0) mbar_wait(bar_otfree, (g - 1) & 1);  // the previous tile's O (and P) have left tensor memory
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = desc_q0 + (uint64_t)((b * QTILE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base, da + 2 * k, desc_k + 2 * k, idesc_s1, k != 0);
          if (N2 > 0) {
            const uint64_t dk2 = desc_k + (uint64_t)((256 * 128) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base + 256, da + 2 * k, dk2 + 2 * k, idesc_s2, k != 0);
          }
          tc_commit(bar_sfull);
          tc_commit(&bar_qfree[b]);
          if (t == n_mt - 1) tc_commit(bar_kfree);
        }
        __syncwarp();
        if (t == 0) mbar_wait(bar_v, item_n & 1);
        mbar_wait(bar_pready, g & 1);
        tc_fence_after();
        if (elect_one()) {
          uint32_t acc = 0;
          for (int r = 0; r < 4; ++r) {
            int kf, nq;
            key_range(r, nks, kf, nq);
            uint32_t a_col = tmem_base + 16 * kf;
            for (int j = 0; j < nq; ++j, a_col += 8) {
              const int ks = kf + j;
              const uint64_t bd = desc_v + (uint64_t)((ks >> 2) * (VT_CHUNK >> 4) + (ks & 3) * 2);
              tc_mma_f16_ts(tmem_base + O_COL, a_col, bd, idesc_o, acc);
              acc = 1;
            }
          }
          tc_commit(bar_ofull);
          if (t == n_mt - 1) tc_commit(bar_vfree);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax + epilogue: 16 warps, lane quarter q = warp & 3, key range kr = (warp - 2) / 4 =====
    const int kr = (warp - 2) >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
    const int ldc = heads * HD;
    const int r_tile = q * 32 + lane;  // row inside the 128-row tile
    float* xmax = reinterpret_cast<float*>(smem + SMEM_XCHG) + r_tile * 4;
    float* xsum = xmax + 512;
    const uint32_t quarter_bar = 1 + q;  // named barrier of the four warps that share these 32 rows
    uint32_t g = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int seq = it / heads, h = it % heads;
      const int start = cu_seqlens[seq];
      const int len = cu_seqlens[seq + 1] - start;
      const int n_mt = (len + 127) >> 7;
      const int shift = start & 7;
      const int Lk = shift + len;
      const int NK = (Lk + 15) & ~15;
      int k_first, nq;
      key_range(kr, NK >> 4, k_first, nq);
      const int cb0 = 16 * k_first;
      for (int t = 0; t < n_mt; ++t, ++g) {
        mbar_wait(bar_sfull, g & 1);
        tc_fence_after();
        const int row = t * 128 + r_tile;  // row inside the sequence
        if (t * 128 + q * 32 >= len) {
          // all 32 rows of this quarter lie beyond the sequence (all four warps of the quarter take this branch): only
          // the hand-shakes; their P columns keep whatever S left there, rows are independent and these are never stored
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_pready);
            mbar_arrive(bar_otfree);
          }
          continue;
        }
        // pass 1: maximum over this warp's key range, 32 columns per load
        float mx0 = -INFINITY, mx1 = -INFINITY;
        {
          uint32_t w[32];
          const int n32 = (nq + 1) >> 1;  // the last load may reach 16 columns past the range (ignored)
          for (int i = 0; i < n32; ++i) {
            tmem_ld_32x32(t_lane + cb0 + 32 * i, w);
            tmem_ld_wait();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int c0 = cb0 + 32 * i + 16 * hh;
              if (2 * i + hh < nq) {
                if (c0 >= shift && c0 + 16 <= Lk) {
#pragma unroll
                  for (int j = 0; j < 16; j += 4) {
                    mx0 = max3(mx0, __uint_as_float(w[16 * hh + j]), __uint_as_float(w[16 * hh + j + 1]));
                    mx1 = max3(mx1, __uint_as_float(w[16 * hh + j + 2]), __uint_as_float(w[16 * hh + j + 3]));
                  }
                } else if (c0 < Lk) {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (c0 + j >= shift && c0 + j < Lk) mx0 = fmaxf(mx0, __uint_as_float(w[16 * hh + j]));
                }
              }
            }
          }
        }
        uint32_t r[2][16];
        if (nq > 0) tmem_ld_32x16(t_lane + cb0, r[0]);  // pass 2's first step: under way during the exchange
        xmax[kr] = fmaxf(mx0, mx1);
        asm volatile("bar.sync %0, 128;" ::"r"(quarter_bar) : "memory");
        const float4 m4 = *reinterpret_cast<const float4*>(xmax);
        const float mb = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w)) * SCALE_LOG2;
        // pass 2: each 16-key step becomes 8 columns of fp16 P written over this warp's own consumed range
        const uint64_t scale2 = pack2(SCALE_LOG2, SCALE_LOG2), nmb2 = pack2(-mb, -mb);
        uint64_t acc2 = pack2(0.f, 0.f);
        float sum = 0.f;
        for (int s0 = 0; s0 < nq; s0 += 2) {
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int s = s0 + b, c0 = cb0 + 16 * s;
            if (s < nq) {
              tmem_ld_wait();
              if (s + 1 < nq) tmem_ld_32x16(t_lane + c0 + 16, r[b ^ 1]);
              uint32_t pk[8];
              if (c0 >= shift && c0 + 16 <= Lk) {
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  float x0, x1;
                  const uint64_t x2 = fma2(pack2(__uint_as_float(r[b][j]), __uint_as_float(r[b][j + 1])), scale2, nmb2);
                  if ((POLY >= 1 && j == 8) || (POLY >= 2 && j == 2) || (POLY >= 3 && j == 12)) {
                    ex2_poly2(x2, x0, x1);
                  } else {
                    unpack2(x2, x0, x1);
                    x0 = fast_ex2(x0);
                    x1 = fast_ex2(x1);
                  }
                  acc2 = add2(acc2, pack2(x0, x1));
                  pk[j >> 1] = pack_f16x2(x0, x1);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  float x0 = 0.f, x1 = 0.f;
                  if (c0 + j >= shift && c0 + j < Lk) x0 = fast_ex2(fmaf(__uint_as_float(r[b][j]), SCALE_LOG2, -mb));
                  if (c0 + j + 1 >= shift && c0 + j + 1 < Lk)
                    x1 = fast_ex2(fmaf(__uint_as_float(r[b][j + 1]), SCALE_LOG2, -mb));
                  sum += x0 + x1;
                  pk[j >> 1] = pack_f16x2(x0, x1);
                }
              }
              tmem_st_32x8(t_lane + cb0 + 8 * s, pk);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pready);
        float s_lo, s_hi;
        unpack2(acc2, s_lo, s_hi);
        xsum[kr] = sum + (s_lo + s_hi);
        asm volatile("bar.sync %0, 128;" ::"r"(quarter_bar) : "memory");
        const float4 s4 = *reinterpret_cast<const float4*>(xsum);
        const float inv = 1.0f / ((s4.x + s4.y) + (s4.z + s4.w));
        // epilogue: head dims [16 kr, 16 kr + 16) of this row: O / sum -> fp16 -> one 32-byte sector of ctx
        mbar_wait(bar_ofull, g & 1);
        tc_fence_after();
        uint32_t o[16];
        tmem_ld_32x16(t_lane + O_COL + 16 * kr, o);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_otfree);
        if (row < len) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(o[j]) * inv;
          const int col = h * HD + 16 * kr;  // first of this thread's 16 ctx columns
          if (out_mode == 3) {
            uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(ctx_) + (long long)(start + row) * ldc + col);
            dst[0] = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
            dst[1] = make_uint4(pack_f16x2(v[8], v[9]), pack_f16x2(v[10], v[11]), pack_f16x2(v[12], v[13]),
                                pack_f16x2(v[14], v[15]));
          } else if (out_mode == 2) {
            float* row_base = static_cast<float*>(ctx_) + (long long)(start + row) * ldc;  // a split row is as wide as an fp32 row
#pragma unroll
            for (int j = 0; j < 16; j += 4) store_split4(row_base, col + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          } else {
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(ctx_) + (long long)(start + row) * ldc + col);
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 f = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (out_mode == 1) {
                f.x = round_tf32(f.x); f.y = round_tf32(f.y); f.z = round_tf32(f.z); f.w = round_tf32(f.w);
              }
              dst[j >> 2] = f;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

bool mer_attention_f16_long_supported(int max_seqlen) { return max_seqlen > 0 && max_seqlen <= MAX_KEYS - 7; }

// Operands as mer_attention_f16_launch: qkv16 fp16 [tokens, 3*heads*64] (V columns unused), vt16 fp16 [heads*64, vt_ld]
// with vt[d, token]; ctx [tokens, heads*64] in the format `out_mode` names (3 fp16, 2 bf16 hi | lo split rows, 1 tf32-
// rounded fp32, 0 fp32).  Sequences of up to 505 tokens.
int mer_attention_f16_long_launch(const void* qkv16, const void* vt16, long long vt_ld, void* ctx16,
                                  const int* cu_seqlens, int n_seq, long long tokens, int heads, cudaStream_t stream,
                                  int max_seqlen, int out_mode) {
  MER_REQUIRE(qkv16 && vt16 && ctx16 && cu_seqlens, "mer_attention_f16_long: null operand");
  MER_REQUIRE(out_mode >= 0 && out_mode <= 3, "mer_attention_f16_long: out_mode %d", out_mode);
  MER_REQUIRE(vt_ld >= tokens && vt_ld % 8 == 0, "mer_attention_f16_long: V^T pitch %lld must be a multiple of 8 >= tokens",
              vt_ld);
  MER_REQUIRE(mer_attention_f16_long_supported(max_seqlen), "mer_attention_f16_long: max_seqlen %d (1 .. %d)", max_seqlen,
              MAX_KEYS - 7);
  CUtensorMap tm, tv;
  {
    const uint64_t dims[2] = {(uint64_t)(3 * heads * HD), (uint64_t)tokens};
    const uint64_t strides[1] = {(uint64_t)(3 * heads * HD) * 2ull};
    const uint32_t box[2] = {64, 128};
    if (int rc = mer_make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, qkv16, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)tokens, (uint64_t)(heads * HD)};
    const uint64_t strides[1] = {(uint64_t)vt_ld * 2ull};
    const uint32_t box[2] = {64, HD};
    if (int rc = mer_make_tmap(&tv, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, vt16, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  const char* poly_env = getenv("MER_ATT_F16_POLY");
  const int poly = poly_env ? atoi(poly_env) : 1;
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, void*, const int*, int, int, int);
  Kern kern = poly == 0 ? attention_f16_long_kernel<0> : poly == 1 ? attention_f16_long_kernel<1>
            : poly == 2 ? attention_f16_long_kernel<2> : attention_f16_long_kernel<3>;
  static MerPerDevice attr_set;
  if (attr_set.needs_setup()) {
    const Kern all[] = {attention_f16_long_kernel<0>, attention_f16_long_kernel<1>, attention_f16_long_kernel<2>,
                        attention_f16_long_kernel<3>};
    for (Kern k : all) MER_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, LONG_SMEM));
    attr_set.mark();
  }
  const long long items = (long long)n_seq * heads;
  if (items <= 0 || tokens <= 0) return 0;
  int grid = mer_num_sms();
  if (items < grid) grid = (int)items;
  const double s_avg = (double)tokens / n_seq;
  const int prof = mer_prof_begin(MER_PROF_ATT_LONG, 4.0 * s_avg * s_avg * HD * (double)items, stream);
  kern<<<grid, LONG_THREADS, LONG_SMEM, stream>>>(tm, tv, ctx16, cu_seqlens, n_seq, heads, out_mode);
  mer_prof_end(prof, stream);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}
