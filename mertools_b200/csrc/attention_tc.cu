// attention_tc.cu — tcgen05 attention for sequences of up to 256 tokens, head_dim 64, TF32.
//
// softmax(Q K^T / 8) V per (sequence, head) for the ViT (S = 197) and HuBERT (S = 249) stacks;
// longer sequences take the flash-style kernel in attention.cu.  Replaces the same reference ops
// (HF eager/sdpa attention, modeling_vit.py:171-196, modeling_hubert.py:262-345).
//
// Persistent, one CTA per SM, work item = (sequence, head):
//   warp 0      TMA producer: K (all keys), V^T (keys contiguous; written transposed by the QKV GEMM
//               epilogue) and the 128-row Q tiles of the item -> 128B-swizzled smem
//   warp 1      tcgen05 issuer:  S = Q_tile K^T  (UMMA 128 x NK x 8, kind::tf32) into TMEM columns
//               [0,256);  then, per 64-key chunk of P staged in smem by the softmax warps,
//               O += P_chunk V_chunk  (UMMA 128 x 64 x 8) into TMEM columns [256,320)
//   warps 2..9  two softmax + epilogue groups of 4 warps, one per 128-row query tile of the item, running
//               concurrently: S rows TMEM -> registers (thread = query row), max, exp2, sum, tf32-rounded
//               P chunks -> swizzled smem (A operand of the second MMA; the buffer is the tile's dead Q
//               tile); O / sum -> swizzled smem transpose -> 512-byte coalesced stores into ctx
// TMEM: tile t owns columns [256t, 256t+256): S_t, with O_t accumulating over S_t's first 64 columns
// (dead once the first P chunk has been taken).
// The score matrix never leaves the SM.  Algorithmic HBM traffic per token and layer: 9 KB of
// q|k|v^T in, 3 KB of ctx out.
// (A first version fed P to the second MMA straight from TMEM and V as an MN-major operand; both
// gave wrong results on the device and were replaced by these two known-good K-major smem operands.)
#include <stdlib.h>

#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

constexpr int HD = 64;
constexpr int MAXS = 256;
constexpr int TC_THREADS = 320;                   // producer, MMA issuer, 2 x 4 softmax/epilogue warps
constexpr int CHUNK_BYTES = MAXS * 128;           // one 32-float column chunk of K: [256][128 B]
constexpr int QTILE_BYTES = 2 * 128 * 128;        // one 128-row Q tile: 2 chunks x [128][128 B]
constexpr int VT_CHUNK = HD * 128;                // V^T chunk: 64 d-rows x 32 keys (128 B)
constexpr int SMEM_K = 0;
constexpr int SMEM_V = 2 * CHUNK_BYTES;           // 8 V^T chunks = 64 KB
constexpr int SMEM_Q = 4 * CHUNK_BYTES;           // Q tile t; after S_t it becomes the P-chunk buffer of
                                                  // tile t (2 x [128 rows][128 B])
constexpr int SMEM_BAR = SMEM_Q + 2 * QTILE_BYTES;
constexpr int TC_SMEM = SMEM_BAR + 256 + 1024;
constexpr uint32_t TILE_COLS = 256, TMEM_COLS = 512;  // tile t: S_t at [256t, 256t+NK), O_t aliases its first 64

__device__ __forceinline__ float fast_ex2(float x) {  // MUFU.EX2, flush-to-zero: 2 ulp, no range fix-up
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {  // SW128, SBO 1024
  return static_cast<uint64_t>((addr & 0x3FFFF) >> 4) | (1ull << 16) | (uint64_t(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

// VER 1 masks every score against the valid key range; VER 2 works in 16-key granules, unmasked (FMNMX3 /
// FFMA2 / FADD2) wherever a granule lies inside [shift, Lk), and skips granules beyond the UMMA key count.
template <int VER>
__global__ void __launch_bounds__(TC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                    const __grid_constant__ CUtensorMap tmap_vt, float* __restrict__ ctx,
                    const int* __restrict__ cu_seqlens, int n_seq, int heads, int out_mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* bar_k = bars + 0;       // producer -> MMA: K tile of the item
  uint64_t* bar_q = bars + 1;       // producer -> MMA: Q tiles
  uint64_t* bar_v = bars + 2;       // producer -> MMA: V^T chunks
  uint64_t* bar_sfull = bars + 3;   // [2] MMA -> softmax group t / producer: S_t complete
  uint64_t* bar_pready = bars + 5;  // [2] softmax group t -> MMA: a P chunk of tile t sits in smem
  uint64_t* bar_pfree = bars + 7;   // [2] MMA -> softmax group t: that chunk has been consumed
  uint64_t* bar_ofull = bars + 9;   // [2] MMA -> softmax group t: O_t complete
  uint64_t* bar_ofree = bars + 11;  // [2] softmax group t -> producer: output staging (V^T region) consumed
  uint64_t* bar_otfree = bars + 13; // [2] softmax group t -> MMA: O_t has been read out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_seq * heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_vt);
    mbar_init(bar_k, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_v, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bar_sfull[t], 1);
      mbar_init(&bar_pready[t], 4);
      mbar_init(&bar_pfree[t], 1);
      mbar_init(&bar_ofull[t], 1);
      mbar_init(&bar_ofree[t], 4);
      mbar_init(&bar_otfree[t], 4);
    }
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
    // ===================== TMA producer =====================
    // warp-uniform loop (coordinates stay in uniform registers); one elected lane issues the copies
    uint32_t uses[2] = {0, 0};  // how often tile slot t has been used so far
    int prev_nmt = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int seq = it / heads, h = it % heads;
      const int start = cu_seqlens[seq];
      const int len = cu_seqlens[seq + 1] - start;
      const int n_mt = (len + 127) >> 7;
      // TMA needs 16-byte aligned inner-dimension starts: V^T (keys contiguous) is read from the
      // token rounded down to a multiple of 4, and K rows likewise, so both MMAs see the same key
      // axis k' = key + shift; the (up to 3) leading keys of the previous sequence are masked
      const int a_start = start & ~3, shift = start - a_start;
      const int Lk = shift + len;
      const int nb = (Lk + 127) >> 7;
      // K: the previous item's S MMAs are done
      for (int t = 0; t < prev_nmt; ++t) mbar_wait(&bar_sfull[t], (uses[t] - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_k, (uint32_t)(2 * nb * 16384));
        for (int c = 0; c < 2; ++c)
          for (int b = 0; b < nb; ++b)
            tma_load_2d(smem + SMEM_K + c * CHUNK_BYTES + b * 16384, &tmap_qkv, bar_k,
                        heads * HD + h * HD + c * 32, a_start + b * 128);
      }
      __syncwarp();
      // Q tiles: the previous item's P V MMAs are done (the Q_t regions double as P-chunk buffers)
      for (int t = 0; t < prev_nmt; ++t) mbar_wait(&bar_ofull[t], (uses[t] - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(bar_q, (uint32_t)(n_mt * QTILE_BYTES));
        for (int t = 0; t < n_mt; ++t)
          for (int c = 0; c < 2; ++c)
            tma_load_2d(smem + SMEM_Q + t * QTILE_BYTES + c * 16384, &tmap_qkv, bar_q, h * HD + c * 32,
                        start + t * 128);
      }
      __syncwarp();
      // V^T: the previous item's epilogues are done (they stage their output in the V^T region)
      for (int t = 0; t < prev_nmt; ++t) mbar_wait(&bar_ofree[t], (uses[t] - 1) & 1);
      const int n_vc = (Lk + 31) >> 5;
      if (elect_one()) {
        mbar_expect_tx(bar_v, (uint32_t)(n_vc * VT_CHUNK));
        for (int c = 0; c < n_vc; ++c)
          tma_load_2d(smem + SMEM_V + c * VT_CHUNK, &tmap_vt, bar_v, a_start + c * 32, h * HD);
      }
      __syncwarp();
      for (int t = 0; t < n_mt; ++t) ++uses[t];
      prev_nmt = n_mt;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop; one elected lane issues the MMAs and their commits
    uint32_t uses[2] = {0, 0};
    uint32_t g[2] = {0, 0};  // P chunks consumed per tile slot
    uint32_t item_n = 0;
    const uint64_t desc_q = desc_kmajor(smem_u32(smem + SMEM_Q));
    const uint64_t desc_k = desc_kmajor(smem_u32(smem + SMEM_K));
    const uint64_t desc_v = desc_kmajor(smem_u32(smem + SMEM_V));
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++item_n) {
      const int seq = it / heads;
      const int start = cu_seqlens[seq];
      const int len = cu_seqlens[seq + 1] - start;
      const int n_mt = (len + 127) >> 7;
      const int NK = ((start & 3) + len + 15) & ~15;  // shifted key axis, padded to the UMMA N step
      const int n_pc = (NK + 63) >> 6;
      const uint32_t idesc_s = umma_idesc(2, 128, NK);
      const uint32_t idesc_o = umma_idesc(2, 128, HD);
      mbar_wait(bar_k, item_n & 1);
      mbar_wait(bar_q, item_n & 1);
      // ---- S_t = Q_t K^T for both tiles ----
      for (int t = 0; t < n_mt; ++t) {
        if (uses[t] > 0) mbar_wait(&bar_otfree[t], (uses[t] - 1) & 1);  // S_t / O_t columns free again
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint64_t da = desc_q + (uint64_t)((t * QTILE_BYTES + c * 16384) >> 4);
            const uint64_t db = desc_k + (uint64_t)((c * CHUNK_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_tf32(tmem_base + t * TILE_COLS, da + 2 * k, db + 2 * k, idesc_s, (c | k) != 0);
          }
          tc_commit(&bar_sfull[t]);
        }
        __syncwarp();
      }
      // ---- O_t += P_t chunk * V chunk, the two tiles interleaved ----
      mbar_wait(bar_v, item_n & 1);
      for (int pc = 0; pc < n_pc; ++pc) {
        const int keys = min(64, NK - pc * 64);
        for (int t = 0; t < n_mt; ++t) {
          mbar_wait(&bar_pready[t], g[t] & 1);
          tc_fence_after();
          if (elect_one()) {
            for (int k8 = 0; k8 < keys / 8; ++k8) {
              const int sub = k8 >> 2, k = k8 & 3;  // 32-key sub-chunk, 8-key step inside it
              const uint64_t da = desc_q + (uint64_t)((t * QTILE_BYTES + sub * 16384) >> 4) + 2 * k;
              const uint64_t db = desc_v + (uint64_t)(((2 * pc + sub) * VT_CHUNK) >> 4) + 2 * k;
              tc_mma_tf32(tmem_base + t * TILE_COLS, da, db, idesc_o, (pc | k8) != 0);
            }
            tc_commit(&bar_pfree[t]);
            if (pc == n_pc - 1) tc_commit(&bar_ofull[t]);
          }
          __syncwarp();
          ++g[t];
        }
      }
      for (int t = 0; t < n_mt; ++t) ++uses[t];
    }
  } else {
    // ===================== softmax + epilogue: group 0 = warps 2..5 (tile 0), group 1 = warps 6..9 =====
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16) + grp * TILE_COLS;
    constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
    const int ldc = heads * HD;
    const int r_tile = q * 32 + lane;  // row inside the 128-row tile
    float* p_lo = reinterpret_cast<float*>(smem + SMEM_Q + grp * QTILE_BYTES) + r_tile * 32;
    float* p_hi = p_lo + 4096;         // second 32-key sub-chunk ([128][32] floats further)
    // output staging: 32 KB of the V^T region per tile (V^T is dead once O_t is complete)
    float* o_lo = reinterpret_cast<float*>(smem + SMEM_V + grp * QTILE_BYTES) + r_tile * 32;
    float* o_hi = o_lo + 4096;
    float* stg = reinterpret_cast<float*>(smem + SMEM_V + grp * QTILE_BYTES) + q * 32 * 32;  // this warp's rows
    const int sub_r = lane >> 3, sub_c = lane & 7;
    uint32_t uses = 0, G = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int seq = it / heads, h = it % heads;
      const int start = cu_seqlens[seq];
      const int len = cu_seqlens[seq + 1] - start;
      const int n_mt = (len + 127) >> 7;
      if (grp >= n_mt) continue;  // single-tile item: group 1 has nothing to do
      const int shift = start & 3;  // keys live at columns [shift, shift + len) of S
      const int Lk = shift + len;
      const int n_chunks = (Lk + 31) >> 5;
      const int n_pc = (((Lk + 15) & ~15) + 63) >> 6;
      const int row = grp * 128 + r_tile;  // query row inside the sequence
      mbar_wait(&bar_sfull[grp], uses & 1);
      tc_fence_after();
      float mb, sum = 0.f;
      if constexpr (VER == 2) {
        if (grp * 128 + q * 32 >= len) {
          // all 32 query rows of this warp lie beyond the sequence (rows 224..255 at 197 tokens): nothing to
          // compute -- the P rows it would write feed output rows that are never stored -- but the chunk
          // hand-shake with the MMA warp still counts four arrivals per tile
          mb = 0.f;
          sum = 1.f;
          for (int pc = 0; pc < n_pc; ++pc, ++G) {
            mbar_wait(&bar_pfree[grp], (G & 1) ^ 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_pready[grp]);
          }
        } else {
          const int NK = (Lk + 15) & ~15;  // the UMMA key count of this item (MMA warp: same expression)
          float mx0 = -INFINITY, mx1 = -INFINITY;
          for (int c = 0; c < n_chunks; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_lane + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
              const int c0 = c * 32 + gi * 16;
              if (c0 >= shift && c0 + 16 <= Lk) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  mx0 = max3(mx0, __uint_as_float(r[gi * 16 + j]), __uint_as_float(r[gi * 16 + j + 1]));
                  mx1 = max3(mx1, __uint_as_float(r[gi * 16 + j + 2]), __uint_as_float(r[gi * 16 + j + 3]));
                }
              } else if (c0 < Lk) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + j >= shift && c0 + j < Lk) mx0 = fmaxf(mx0, __uint_as_float(r[gi * 16 + j]));
              }
            }
          }
          mb = fmaxf(mx0, mx1) * SCALE_LOG2;
          const uint64_t scale2 = pack2(SCALE_LOG2, SCALE_LOG2), nmb2 = pack2(-mb, -mb);
          uint64_t acc2 = pack2(0.f, 0.f);
          for (int pc = 0; pc < n_pc; ++pc, ++G) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(t_lane + pc * 64, r0);
            tmem_ld_32x32(t_lane + pc * 64 + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {  // in place: scores -> tf32-rounded probabilities
              uint32_t* rr = gi < 2 ? r0 + gi * 16 : r1 + (gi - 2) * 16;
              const int c0 = pc * 64 + gi * 16;
              if (c0 >= shift && c0 + 16 <= Lk) {
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  float a, b;
                  unpack2(fma2(pack2(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1])), scale2, nmb2), a, b);
                  a = fast_ex2(a);
                  b = fast_ex2(b);
                  acc2 = add2(acc2, pack2(a, b));
                  rr[j] = __float_as_uint(round_tf32(a));
                  rr[j + 1] = __float_as_uint(round_tf32(b));
                }
              } else if (c0 < NK) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float a = 0.f;
                  if (c0 + j >= shift && c0 + j < Lk) a = fast_ex2(fmaf(__uint_as_float(rr[j]), SCALE_LOG2, -mb));
                  sum += a;
                  rr[j] = __float_as_uint(round_tf32(a));
                }
              }
            }
            const int keys = min(64, NK - pc * 64);  // keys of this chunk the P V MMAs read (multiple of 16)
            mbar_wait(&bar_pfree[grp], (G & 1) ^ 1);  // the previous chunk's MMAs have read the buffer
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // 16-byte slot j = keys 4j .. 4j+3 of either 32-key sub-chunk
              const int sj = (j ^ (r_tile & 7)) << 2;
              if (4 * j < keys)
                *reinterpret_cast<uint4*>(p_lo + sj) = make_uint4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
              if (32 + 4 * j < keys)
                *reinterpret_cast<uint4*>(p_hi + sj) = make_uint4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_pready[grp]);
          }
          float s_lo, s_hi;
          unpack2(acc2, s_lo, s_hi);
          sum += s_lo + s_hi;
        }
      } else {
      // pass 1: row maximum over the valid keys
      float mx = -INFINITY;
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_lane + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j >= shift && c * 32 + j < Lk) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
      mb = mx * SCALE_LOG2;
      // pass 2, per 64-key chunk: p = exp2((s - max) / 8 * log2 e) -> row sum, tf32-rounded P into the
      // swizzled smem chunk (A operand of the P V MMA)
      for (int pc = 0; pc < n_pc; ++pc, ++G) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_lane + pc * 64, r0);
        tmem_ld_32x32(t_lane + pc * 64 + 32, r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float p0 = 0.f, p1 = 0.f;
          const int k0 = pc * 64 + j, k1 = k0 + 32;
          if (k0 >= shift && k0 < Lk) p0 = fast_ex2(fmaf(__uint_as_float(r0[j]), SCALE_LOG2, -mb));
          if (k1 >= shift && k1 < Lk) p1 = fast_ex2(fmaf(__uint_as_float(r1[j]), SCALE_LOG2, -mb));
          sum += p0 + p1;
          r0[j] = __float_as_uint(round_tf32(p0));
          r1[j] = __float_as_uint(round_tf32(p1));
        }
        mbar_wait(&bar_pfree[grp], (G & 1) ^ 1);  // the previous chunk's MMAs have read the buffer
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int sj = (j ^ (r_tile & 7)) << 2;
          *reinterpret_cast<uint4*>(p_lo + sj) = make_uint4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
          *reinterpret_cast<uint4*>(p_hi + sj) = make_uint4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]);
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_pready[grp]);
      }
      }
      const float inv = 1.0f / sum;
      // epilogue: O / sum -> (swizzled smem transpose) -> 512-byte coalesced stores into ctx
      mbar_wait(&bar_ofull[grp], uses & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(t_lane, o0);
      tmem_ld_32x32(t_lane + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_otfree[grp]);  // the MMA warp may start the next item's S_t
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t* o = half ? o1 : o0;
        float* dst_row = (half ? o_hi : o_lo);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v = make_float4(__uint_as_float(o[4 * j]) * inv, __uint_as_float(o[4 * j + 1]) * inv,
                                 __uint_as_float(o[4 * j + 2]) * inv, __uint_as_float(o[4 * j + 3]) * inv);
          if (out_mode == 2) {
            // split bf16 row group [32 hi | 32 lo]: 8 bytes of each half per j
            const float hx = bf16_round(v.x), hy = bf16_round(v.y), hz = bf16_round(v.z), hw = bf16_round(v.w);
            uint2* hi = reinterpret_cast<uint2*>(dst_row + (((j >> 1) ^ (r_tile & 7)) << 2)) + (j & 1);
            uint2* lo = reinterpret_cast<uint2*>(dst_row + (((4 + (j >> 1)) ^ (r_tile & 7)) << 2)) + (j & 1);
            *hi = make_uint2(pack_bf16x2(hx, hy), pack_bf16x2(hz, hw));
            *lo = make_uint2(pack_bf16x2(v.x - hx, v.y - hy), pack_bf16x2(v.z - hz, v.w - hw));
          } else {
            if (out_mode == 1) {
              v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
            }
            *reinterpret_cast<float4*>(dst_row + ((j ^ (r_tile & 7)) << 2)) = v;
          }
        }
      }
      __syncwarp();
      const int row0 = grp * 128 + q * 32;  // first sequence row of this warp's 32
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float* src = stg + half * 4096;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + sub_r;
          const int rt = q * 32 + rr;  // row inside the tile (swizzle key)
          const float4 d = *reinterpret_cast<const float4*>(src + rr * 32 + ((sub_c ^ (rt & 7)) << 2));
          if (row0 + rr < len) {
            const long long e = (long long)(start + row0 + rr) * ldc + h * HD + half * 32 + sub_c * 4;
            if (out_mode == 3)  // fp16 ctx (operand of an F16 out-proj GEMM)
              *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(ctx) + e) =
                  make_uint2(pack_f16x2(d.x, d.y), pack_f16x2(d.z, d.w));
            else
              *reinterpret_cast<float4*>(ctx + e) = d;
          }
        }
      }
      (void)row;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ofree[grp]);
      ++uses;
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

int mer_attention_tc_launch(const float* qkv, const float* vt, long long vt_ld, float* ctx,
                            const int* cu_seqlens, int n_seq, long long tokens, int heads, int flags,
                            cudaStream_t stream) {
  MER_REQUIRE(vt && vt_ld >= tokens && vt_ld % 4 == 0, "mer_attention_tc: V^T buffer missing or mis-pitched");
  CUtensorMap tm, tv;
  {
    const uint64_t dims[2] = {(uint64_t)(3 * heads * HD), (uint64_t)tokens};
    const uint64_t strides[1] = {(uint64_t)(3 * heads * HD) * 4ull};
    const uint32_t box[2] = {32, 128};
    if (int rc = mer_make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, qkv, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)tokens, (uint64_t)(heads * HD)};
    const uint64_t strides[1] = {(uint64_t)vt_ld * 4ull};
    const uint32_t box[2] = {32, HD};
    if (int rc = mer_make_tmap(&tv, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, vt, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  // MER_ATT_TC_VER=2 selects the granule softmax (see the kernel's header comment); read at every launch so that
  // a test can run both versions in one process.
  const char* ver_env = getenv("MER_ATT_TC_VER");
  const bool ver2 = !(ver_env && atoi(ver_env) == 1);  // granule softmax is the default since round 2 (measured)
  auto kern = ver2 ? attention_tc_kernel<2> : attention_tc_kernel<1>;
  static MerPerDevice attr_set;
  if (attr_set.needs_setup()) {
    MER_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    MER_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    attr_set.mark();
  }
  const long long items = (long long)n_seq * heads;
  int grid = mer_num_sms();
  if (items < grid) grid = (int)items;
  const int out_mode = (flags & MER_EPI_OUT_F16) ? 3 : (flags & MER_EPI_SPLIT_BF16) ? 2 : ((flags & MER_EPI_ROUND_TF32) ? 1 : 0);
  const double s_avg = (double)tokens / n_seq;  // exact for equal-length batches
  const int prof = mer_prof_begin(MER_PROF_ATT_TC, 4.0 * s_avg * s_avg * HD * (double)items, stream);
  kern<<<grid, TC_THREADS, TC_SMEM, stream>>>(tm, tv, ctx, cu_seqlens, n_seq, heads, out_mode);
  mer_prof_end(prof, stream);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}
