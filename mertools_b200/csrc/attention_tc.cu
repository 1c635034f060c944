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
//   warps 2..5  softmax + epilogue: S rows TMEM -> registers (thread = query row), max, exp2, sum,
//               tf32-rounded P chunks -> swizzled smem (A operand of the second MMA); O / sum -> ctx
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
constexpr int TC_THREADS = 192;
constexpr int CHUNK_BYTES = MAXS * 128;           // one 32-float column chunk of K or V: [256][128 B]
constexpr int QTILE_BYTES = 2 * 128 * 128;        // one 128-row Q tile: 2 chunks x [128][128 B]
constexpr int VT_CHUNK = HD * 128;                // V^T chunk: 64 d-rows x 32 keys (128 B)
constexpr int SMEM_K = 0;
constexpr int SMEM_V = 2 * CHUNK_BYTES;           // 8 V^T chunks = 64 KB
constexpr int SMEM_Q = 4 * CHUNK_BYTES;
constexpr int SMEM_P = SMEM_Q + 2 * QTILE_BYTES;  // P chunk: 2 x [128 rows][32 keys] = 32 KB
constexpr int SMEM_BAR = SMEM_P + 2 * 16384;
constexpr int TC_SMEM = SMEM_BAR + 256 + 1024;
constexpr uint32_t S_COL = 0, O_COL = 256, TMEM_COLS = 512;

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {  // SW128, SBO 1024
  return static_cast<uint64_t>((addr & 0x3FFFF) >> 4) | (1ull << 16) | (uint64_t(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                    const __grid_constant__ CUtensorMap tmap_vt, float* __restrict__ ctx,
                    const int* __restrict__ cu_seqlens, int n_seq, int heads, int out_mode) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (not through an integer round trip) so the compiler keeps the
  // shared address space of everything derived from it (st.shared / ld.shared, not generic ST / LD)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* bar_kq = bars + 0;
  uint64_t* bar_v = bars + 1;
  // sfull / ofull exist twice, indexed by (tile counter & 1) with phase (tile counter >> 1) & 1: the
  // producer looks back at BOTH tiles of the previous item, and a parity wait is only unambiguous
  // when the waiter can never be two completions behind on the same barrier
  uint64_t* bar_sfull = bars + 2;   // [2] MMA -> softmax/producer: S tile complete
  uint64_t* bar_pready = bars + 4;  // softmax -> MMA: a P chunk sits in smem
  uint64_t* bar_ofull = bars + 5;   // [2] MMA -> softmax/producer: O complete (all MMAs of the tile done)
  uint64_t* bar_ofree = bars + 7;
  uint64_t* bar_pfree = bars + 8;   // MMA -> softmax: the P chunk buffer has been consumed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_seq * heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_vt);
    mbar_init(bar_kq, 1);
    mbar_init(bar_v, 1);
    mbar_init(&bar_sfull[0], 1);
    mbar_init(&bar_sfull[1], 1);
    mbar_init(bar_pready, 4);
    mbar_init(bar_pfree, 1);
    mbar_init(&bar_ofull[0], 1);
    mbar_init(&bar_ofull[1], 1);
    mbar_init(bar_ofree, 4);
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
    if (lane == 0) {
      uint32_t tiles_done = 0;  // tiles of all previous items of this CTA
      uint32_t prev_tiles = 0;  // tiles of the previous item (0 for the first)
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
        const int nb = (Lk + 127) >> 7;  // 128-row boxes of K actually needed
        // K + Q tile 0: the previous item's S MMAs (last tile) must have consumed K and Q
        for (uint32_t tt = tiles_done - prev_tiles; tt < tiles_done; ++tt)
          mbar_wait(&bar_sfull[tt & 1], (tt >> 1) & 1);
        mbar_expect_tx(bar_kq, (uint32_t)(2 * nb * 16384 + 2 * 16384));
        for (int c = 0; c < 2; ++c) {
          for (int b = 0; b < nb; ++b)
            tma_load_2d(smem + SMEM_K + c * CHUNK_BYTES + b * 16384, &tmap_qkv, bar_kq,
                        heads * HD + h * HD + c * 32, a_start + b * 128);
          tma_load_2d(smem + SMEM_Q + c * 16384, &tmap_qkv, bar_kq, h * HD + c * 32, start);
        }
        // V (+ Q tile 1): the previous item's P V MMAs must have consumed V
        for (uint32_t tt = tiles_done - prev_tiles; tt < tiles_done; ++tt)
          mbar_wait(&bar_ofull[tt & 1], (tt >> 1) & 1);
        const int n_vc = (Lk + 31) >> 5;  // 32-key chunks of V^T
        mbar_expect_tx(bar_v, (uint32_t)(n_vc * VT_CHUNK + (n_mt > 1 ? 2 * 16384 : 0)));
        for (int c = 0; c < n_vc; ++c)
          tma_load_2d(smem + SMEM_V + c * VT_CHUNK, &tmap_vt, bar_v, a_start + c * 32, h * HD);
        if (n_mt > 1)
          for (int c = 0; c < 2; ++c)
            tma_load_2d(smem + SMEM_Q + QTILE_BYTES + c * 16384, &tmap_qkv, bar_v, h * HD + c * 32,
                        start + 128);
        tiles_done += n_mt;
        prev_tiles = n_mt;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t T = 0;       // global tile counter of this CTA
      uint32_t G = 0;       // global P-chunk counter of this CTA
      uint32_t item_n = 0;  // items processed
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++item_n) {
        const int seq = it / heads;
        const int start = cu_seqlens[seq];
        const int len = cu_seqlens[seq + 1] - start;
        const int n_mt = (len + 127) >> 7;
        const int NK = ((start & 3) + len + 15) & ~15;  // shifted key axis, padded to the UMMA N step
        const int n_pc = (NK + 63) >> 6;  // 64-key chunks of P
        const uint32_t idesc_s = umma_idesc(2, 128, NK);
        const uint32_t idesc_o = umma_idesc(2, 128, HD);
        for (int t = 0; t < n_mt; ++t, ++T) {
          if (t == 0) mbar_wait(bar_kq, item_n & 1);
          else mbar_wait(bar_v, item_n & 1);  // Q tile 1 travels with V^T
          tc_fence_after();
          // ---- S = Q_t K^T ----
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint64_t da = desc_kmajor(smem_u32(smem + SMEM_Q + t * QTILE_BYTES + c * 16384));
            const uint64_t db = desc_kmajor(smem_u32(smem + SMEM_K + c * CHUNK_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_tf32(tmem_base + S_COL, da + 2 * k, db + 2 * k, idesc_s, (c | k) != 0);
          }
          tc_commit(&bar_sfull[T & 1]);
          // ---- O = sum over 64-key chunks of P_chunk V_chunk ----
          if (t == 0) mbar_wait(bar_v, item_n & 1);
          mbar_wait(bar_ofree, (T & 1) ^ 1);
          for (int pc = 0; pc < n_pc; ++pc, ++G) {
            mbar_wait(bar_pready, G & 1);
            tc_fence_after();
            const int keys = min(64, NK - pc * 64);
            for (int k8 = 0; k8 < keys / 8; ++k8) {
              const int sub = k8 >> 2, k = k8 & 3;  // 32-key sub-chunk, 8-key step inside it
              const uint64_t da = desc_kmajor(smem_u32(smem + SMEM_P + sub * 16384)) + 2 * k;
              const uint64_t db = desc_kmajor(smem_u32(smem + SMEM_V + (2 * pc + sub) * VT_CHUNK)) + 2 * k;
              tc_mma_tf32(tmem_base + O_COL, da, db, idesc_o, (pc | k8) != 0);
            }
            tc_commit(bar_pfree);
          }
          tc_commit(&bar_ofull[T & 1]);
          mbar_wait(&bar_ofull[T & 1], (T >> 1) & 1);  // S is rewritten by the next tile
        }
      }
    }
  } else {
    // ===================== softmax + epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
    const int ldc = heads * HD;
    uint32_t T = 0, G = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int seq = it / heads, h = it % heads;
      const int start = cu_seqlens[seq];
      const int len = cu_seqlens[seq + 1] - start;
      const int n_mt = (len + 127) >> 7;
      const int shift = start & 3;   // keys live at columns [shift, shift + len) of S
      const int Lk = shift + len;
      const int n_chunks = (Lk + 31) >> 5;
      for (int t = 0; t < n_mt; ++t, ++T) {
        const int row = t * 128 + q * 32 + lane;  // query row inside the sequence
        mbar_wait(&bar_sfull[T & 1], (T >> 1) & 1);
        tc_fence_after();
        // pass 1: row maximum over the valid keys
        float mx = -INFINITY;
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_lane + S_COL + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j >= shift && c * 32 + j < Lk) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        const float mb = mx * SCALE_LOG2;
        // pass 2, per 64-key chunk: p = exp2((s - max) / 8 * log2 e) -> row sum, tf32-rounded P into
        // the swizzled smem chunk (A operand of the P V MMA)
        float sum = 0.f;
        const int r_tile = q * 32 + lane;  // row inside the 128-row tile
        const int NKs = (Lk + 15) & ~15;
        const int n_pc = (NKs + 63) >> 6;
        for (int pc = 0; pc < n_pc; ++pc, ++G) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(t_lane + S_COL + pc * 64, r0);
          tmem_ld_32x32(t_lane + S_COL + pc * 64 + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float p0 = 0.f, p1 = 0.f;
            const int k0 = pc * 64 + j, k1 = k0 + 32;
            if (k0 >= shift && k0 < Lk) p0 = exp2f(fmaf(__uint_as_float(r0[j]), SCALE_LOG2, -mb));
            if (k1 >= shift && k1 < Lk) p1 = exp2f(fmaf(__uint_as_float(r1[j]), SCALE_LOG2, -mb));
            sum += p0 + p1;
            r0[j] = __float_as_uint(round_tf32(p0));
            r1[j] = __float_as_uint(round_tf32(p1));
          }
          mbar_wait(bar_pfree, (G & 1) ^ 1);  // previous chunk's MMAs have read the buffer
          float* p_lo = reinterpret_cast<float*>(smem + SMEM_P) + r_tile * 32;
          float* p_hi = reinterpret_cast<float*>(smem + SMEM_P + 16384) + r_tile * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int sj = (j ^ (r_tile & 7)) << 2;
            *reinterpret_cast<uint4*>(p_lo + sj) = make_uint4(r0[4 * j], r0[4 * j + 1], r0[4 * j + 2], r0[4 * j + 3]);
            *reinterpret_cast<uint4*>(p_hi + sj) = make_uint4(r1[4 * j], r1[4 * j + 1], r1[4 * j + 2], r1[4 * j + 3]);
          }
          fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_pready);
        }
        const float inv = 1.0f / sum;
        // epilogue: O / sum -> ctx
        mbar_wait(&bar_ofull[T & 1], (T >> 1) & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_lane + O_COL, o0);
        tmem_ld_32x32(t_lane + O_COL + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ofree);
        if (row < len) {
          float* dst = ctx + (long long)(start + row) * ldc;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t* o = half ? o1 : o0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv,
                                     __uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
              const int col = h * HD + half * 32 + j;
              if (out_mode == 2) {
                store_split4(dst, col, v);
              } else {
                if (out_mode == 1) {
                  v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
                }
                *reinterpret_cast<float4*>(dst + col) = v;
              }
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
  static bool attr_set = false;
  if (!attr_set) {
    MER_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TC_SMEM));
    attr_set = true;
  }
  const long long items = (long long)n_seq * heads;
  int grid = mer_num_sms();
  if (items < grid) grid = (int)items;
  const int out_mode = (flags & MER_EPI_SPLIT_BF16) ? 2 : ((flags & MER_EPI_ROUND_TF32) ? 1 : 0);
  attention_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(tm, tv, ctx, cu_seqlens, n_seq, heads, out_mode);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}
