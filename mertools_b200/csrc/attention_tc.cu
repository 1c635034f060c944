// attention_tc.cu — tcgen05 attention for sequences of up to 256 tokens, head_dim 64, TF32.
//
// softmax(Q K^T / 8) V per (sequence, head) for the ViT (S = 197) and HuBERT (S = 249) stacks;
// longer sequences take the flash-style kernel in attention.cu.  Replaces the same reference ops
// (HF eager/sdpa attention, modeling_vit.py:171-196, modeling_hubert.py:262-345).
//
// Persistent, one CTA per SM, work item = (sequence, head):
//   warp 0      TMA producer: K, V (all keys) and the 128-row Q tiles of the item -> 128B-swizzled smem
//   warp 1      tcgen05 issuer:  S = Q_tile K^T  (UMMA 128 x NK x 8, kind::tf32, both operands K-major
//               in smem) into TMEM columns [0,256);  then  O = P V  with P read straight from TMEM
//               (A operand in tensor memory) and V as an MN-major smem operand, into columns [256,320)
//   warps 2..5  softmax + epilogue: S rows TMEM -> registers (thread = query row), max, exp2, sum,
//               tf32-rounded P written back IN PLACE with tcgen05.st; after P V: O / sum -> ctx
// The score matrix never leaves the SM.  Algorithmic HBM traffic per token and layer: 9 KB of qkv in,
// 3 KB of ctx out.
#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

constexpr int HD = 64;
constexpr int MAXS = 256;
constexpr int TC_THREADS = 192;
constexpr int CHUNK_BYTES = MAXS * 128;           // one 32-float column chunk of K or V: [256][128 B]
constexpr int QTILE_BYTES = 2 * 128 * 128;        // one 128-row Q tile: 2 chunks x [128][128 B]
constexpr int SMEM_K = 0;
constexpr int SMEM_V = 2 * CHUNK_BYTES;
constexpr int SMEM_Q = 4 * CHUNK_BYTES;
constexpr int SMEM_BAR = SMEM_Q + 2 * QTILE_BYTES;
constexpr int TC_SMEM = SMEM_BAR + 256 + 1024;
constexpr uint32_t S_COL = 0, O_COL = 256, TMEM_COLS = 512;

__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {  // SW128, SBO 1024
  return static_cast<uint64_t>((addr & 0x3FFFF) >> 4) | (1ull << 16) | (uint64_t(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// MN-major SW128 operand: 32-float (128 B) atoms along MN `lbo` bytes apart, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo) {
  return static_cast<uint64_t>((addr & 0x3FFFF) >> 4) | (uint64_t(lbo >> 4) << 16) |
         (uint64_t(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
      "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, float* __restrict__ ctx,
                    const int* __restrict__ cu_seqlens, int n_seq, int heads, int out_mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
  uint64_t* bar_kq = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_sfull = bars + 2;
  uint64_t* bar_pfull = bars + 3;
  uint64_t* bar_ofull = bars + 4;
  uint64_t* bar_ofree = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_seq * heads;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(bar_kq, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_sfull, 1);
    mbar_init(bar_pfull, 4);
    mbar_init(bar_ofull, 1);
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
      bool first = true;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int seq = it / heads, h = it % heads;
        const int start = cu_seqlens[seq];
        const int len = cu_seqlens[seq + 1] - start;
        const int n_mt = (len + 127) >> 7;
        const int nb = n_mt;  // 128-row boxes of K / V actually needed
        // K + Q tile 0: the previous item's S MMAs (last tile) must have consumed K and Q
        if (!first) mbar_wait(bar_sfull, (tiles_done - 1) & 1);
        mbar_expect_tx(bar_kq, (uint32_t)(2 * nb * 16384 + 2 * 16384));
        for (int c = 0; c < 2; ++c) {
          for (int b = 0; b < nb; ++b)
            tma_load_2d(smem + SMEM_K + c * CHUNK_BYTES + b * 16384, &tmap_qkv, bar_kq,
                        heads * HD + h * HD + c * 32, start + b * 128);
          tma_load_2d(smem + SMEM_Q + c * 16384, &tmap_qkv, bar_kq, h * HD + c * 32, start);
        }
        // V (+ Q tile 1): the previous item's P V MMAs must have consumed V
        if (!first) mbar_wait(bar_ofull, (tiles_done - 1) & 1);
        mbar_expect_tx(bar_v, (uint32_t)(2 * nb * 16384 + (n_mt > 1 ? 2 * 16384 : 0)));
        for (int c = 0; c < 2; ++c) {
          for (int b = 0; b < nb; ++b)
            tma_load_2d(smem + SMEM_V + c * CHUNK_BYTES + b * 16384, &tmap_qkv, bar_v,
                        2 * heads * HD + h * HD + c * 32, start + b * 128);
          if (n_mt > 1)
            tma_load_2d(smem + SMEM_Q + QTILE_BYTES + c * 16384, &tmap_qkv, bar_v, h * HD + c * 32,
                        start + 128);
        }
        tiles_done += n_mt;
        first = false;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t T = 0;       // global tile counter of this CTA
      uint32_t item_n = 0;  // items processed
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++item_n) {
        const int seq = it / heads;
        const int len = cu_seqlens[seq + 1] - cu_seqlens[seq];
        const int n_mt = (len + 127) >> 7;
        const int NK = (len + 15) & ~15;  // keys, padded to the UMMA N granularity
        const uint32_t idesc_s = umma_idesc(2, 128, NK);
        const uint32_t idesc_o = umma_idesc(2, 128, HD) | (1u << 16);  // B (= V) is MN-major
        for (int t = 0; t < n_mt; ++t, ++T) {
          if (t == 0) mbar_wait(bar_kq, item_n & 1);
          else mbar_wait(bar_v, item_n & 1);  // Q tile 1 travels with V
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
          tc_commit(bar_sfull);
          // ---- O = P V ----
          mbar_wait(bar_pfull, T & 1);
          if (t == 0) mbar_wait(bar_v, item_n & 1);
          mbar_wait(bar_ofree, (T & 1) ^ 1);
          tc_fence_after();
          const uint32_t v0 = smem_u32(smem + SMEM_V);
          for (int j = 0; j < NK / 8; ++j)
            mma_tf32_ts(tmem_base + O_COL, tmem_base + S_COL + 8 * j,
                        desc_mnmajor(v0 + j * 1024, CHUNK_BYTES), idesc_o, j != 0);
          tc_commit(bar_ofull);
          mbar_wait(bar_ofull, T & 1);  // S region is rewritten by the next tile's first MMA
        }
      }
    }
  } else {
    // ===================== softmax + epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;
    const int ldc = heads * HD;
    uint32_t T = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int seq = it / heads, h = it % heads;
      const int start = cu_seqlens[seq];
      const int len = cu_seqlens[seq + 1] - start;
      const int n_mt = (len + 127) >> 7;
      const int n_chunks = (len + 31) >> 5;
      for (int t = 0; t < n_mt; ++t, ++T) {
        const int row = t * 128 + q * 32 + lane;  // query row inside the sequence
        mbar_wait(bar_sfull, T & 1);
        tc_fence_after();
        // pass 1: row maximum over the valid keys
        float mx = -INFINITY;
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_lane + S_COL + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < len) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        const float mb = mx * SCALE_LOG2;
        // pass 2: p = exp2((s - max) / 8 * log2 e), row sum, tf32-rounded P back in place
        float sum = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_lane + S_COL + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float p = 0.f;
            if (c * 32 + j < len) p = exp2f(fmaf(__uint_as_float(r[j]), SCALE_LOG2, -mb));
            sum += p;
            r[j] = __float_as_uint(round_tf32(p));
          }
          tmem_st_32x32(t_lane + S_COL + c * 32, r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pfull);
        const float inv = 1.0f / sum;
        // epilogue: O / sum -> ctx
        mbar_wait(bar_ofull, T & 1);
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

int mer_attention_tc_launch(const float* qkv, float* ctx, const int* cu_seqlens, int n_seq,
                            long long tokens, int heads, int flags, cudaStream_t stream) {
  CUtensorMap tm;
  const uint64_t dims[2] = {(uint64_t)(3 * heads * HD), (uint64_t)tokens};
  const uint64_t strides[1] = {(uint64_t)(3 * heads * HD) * 4ull};
  const uint32_t box[2] = {32, 128};
  if (int rc = mer_make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, qkv, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
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
  attention_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(tm, ctx, cu_seqlens, n_seq, heads, out_mode);
  MER_CUDA_CHECK(cudaGetLastError());
  mer_count_launches(1);
  return 0;
}
