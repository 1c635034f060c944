// gemm_tf32.cu — the one GEMM every encoder layer goes through.
//
//   out[b*out_bstride + out_row0 + m, n] =
//       round?( act( sum_k A[b, m, k] * W[n, k] + bias[n] ) + res[b*res_bstride + res_row0 + m, n] )
//
// A is a (K, rows, batches) tensor described by a TMA map with ARBITRARY row / batch strides,
// which is how the strided HuBERT convolutions (time-major activations, overlapping windows)
// and the ViT patch embedding run through the same kernel as the Linear layers
// (reference ops: HF ViT/HuBERT/BERT nn.Linear + nn.Conv1d, see DESIGN.md kernel table).
// W is the nn.Linear weight as stored: [N, K] row-major == K-major B operand.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer: A/B tiles -> 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1      MMA issuer  : tcgen05.mma.kind::tf32, UMMA 128 x BLOCK_N x 8, fp32 accum in TMEM
//   warp 2      TMEM allocator
//   warps 4..7  epilogue    : tcgen05.ld TMEM -> registers -> bias/GELU/residual -> st.global
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

using namespace mer;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;  // 32 tf32 = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;    // 32 bytes of K per tcgen05.mma
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 4;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +align slack
};

template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const MerGemmEpilogue ep,
                 int rows_per_batch, int batches, int N, int K, int K_inner, int P) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint64_t* tempty_bar = bars + 2 * Cfg::kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (rows_per_batch + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = N / BLOCK_N;
  const int num_tiles = batches * m_tiles * n_tiles;
  const int num_kb = K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_blk = t % n_tiles;
        const int mb = t / n_tiles;
        const int b = mb / m_tiles;
        const int mt = mb % m_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int kk = kb * BLOCK_K;
          const int tap = kk / K_inner;
          tma_load_4d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kk - tap * K_inner,
                      tap % P, mt * BLOCK_M + tap / P, b);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * BLOCK_K,
                      n_blk * BLOCK_N);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(2, BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance the start address by k*32 bytes inside the 128B swizzle row (>>4 => +2)
            tc_mma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int ew = warp - EPI_WARP0;  // == warp % 4: the TMEM lane quarter this warp may read
    int as = 0;
    uint32_t aphase = 0;
    const bool do_gelu = (ep.flags & MER_EPI_GELU) != 0;
    const bool do_round = (ep.flags & MER_EPI_ROUND_TF32) != 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n_blk = t % n_tiles;
      const int mb = t / n_tiles;
      const int b = mb / m_tiles;
      const int mt = mb % m_tiles;
      const int m = mt * BLOCK_M + ew * 32 + lane;  // row inside the batch entry
      const bool valid = m < rows_per_batch;
      float* out_row =
          ep.out + ((long long)b * ep.out_bstride + ep.out_row0 + m) * (long long)ep.ld_out;
      const float* res_row =
          ep.res ? ep.res + ((long long)b * ep.res_bstride + ep.res_row0 + m) * (long long)ep.ld_res
                 : nullptr;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        const int n0 = n_blk * BLOCK_N + c * 32;
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v;
            v.x = __uint_as_float(r[j + 0]);
            v.y = __uint_as_float(r[j + 1]);
            v.z = __uint_as_float(r[j + 2]);
            v.w = __uint_as_float(r[j + 3]);
            if (ep.bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
              v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            if (do_gelu) {
              v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
            }
            if (res_row) {
              const float4 rr = *reinterpret_cast<const float4*>(res_row + n0 + j);
              v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            }
            if (do_round) {
              v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w);
            }
            *reinterpret_cast<float4*>(out_row + n0 + j) = v;
          }
        }
      }
      // all TMEM reads of this stage are complete (wait::ld above) -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_N>
int launch_gemm(const MerGemmDesc* g, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {(uint64_t)g->K_inner, (uint64_t)g->P, (uint64_t)g->a_rows_dim,
                              (uint64_t)g->batches};
    const uint64_t strides[3] = {(uint64_t)g->a_phase_stride * 4ull,
                                 (uint64_t)g->a_row_stride * 4ull,
                                 (uint64_t)g->a_batch_stride * 4ull};
    const uint32_t box[4] = {BLOCK_K, 1, BLOCK_M, 1};
    if (int rc = mer_make_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g->A, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    const int K = g->K_inner * g->taps;
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)g->N};
    const uint64_t strides[1] = {(uint64_t)K * 4ull};
    const uint32_t box[2] = {BLOCK_K, BLOCK_N};
    if (int rc = mer_make_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g->W, dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    MER_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32_kernel<BLOCK_N>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_set = true;
  }
  const int m_tiles = (g->rows_per_batch + BLOCK_M - 1) / BLOCK_M;
  const long long tiles = (long long)g->batches * m_tiles * (g->N / BLOCK_N);
  int grid = mer_num_sms();
  if (tiles < grid) grid = (int)tiles;
  gemm_tf32_kernel<BLOCK_N><<<grid, NUM_THREADS, Cfg::kSmemBytes, stream>>>(
      ta, tb, g->ep, g->rows_per_batch, g->batches, g->N, g->K_inner * g->taps, g->K_inner, g->P);
  MER_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

int mer_gemm_tf32_launch(const MerGemmDesc* g, cudaStream_t stream) {
  MER_REQUIRE(g && g->A && g->W && g->ep.out, "mer_gemm_tf32: null operand");
  MER_REQUIRE(g->K_inner > 0 && g->K_inner % BLOCK_K == 0 && g->taps > 0 && g->P > 0,
              "mer_gemm_tf32: K_inner=%d must be a positive multiple of %d (taps=%d P=%d)",
              g->K_inner, BLOCK_K, g->taps, g->P);
  MER_REQUIRE(g->a_rows_dim >= g->rows_per_batch, "mer_gemm_tf32: a_rows_dim < rows_per_batch");
  MER_REQUIRE(g->N > 0 && g->N % 128 == 0, "mer_gemm_tf32: N=%d must be a multiple of 128", g->N);
  MER_REQUIRE(g->rows_per_batch > 0 && g->batches > 0, "mer_gemm_tf32: empty problem");
  MER_REQUIRE(g->a_row_stride % 4 == 0 && g->a_batch_stride % 4 == 0 && g->a_phase_stride % 4 == 0,
              "mer_gemm_tf32: A strides must be multiples of 16 bytes");
  MER_REQUIRE(g->ep.ld_out % 4 == 0 && (g->ep.res == nullptr || g->ep.ld_res % 4 == 0),
              "mer_gemm_tf32: out/res leading dims must be multiples of 4 floats");
  const int m_tiles = (g->rows_per_batch + BLOCK_M - 1) / BLOCK_M;
  const long long tiles256 = (g->N % 256 == 0) ? (long long)g->batches * m_tiles * (g->N / 256) : 0;
  // 128 x 256 tiles whenever they fill the machine; 128 x 128 for small problems / N % 256 != 0
  if (tiles256 >= mer_num_sms() && g->force_block_n != 128) return launch_gemm<256>(g, stream);
  if (g->force_block_n == 256 && tiles256 > 0) return launch_gemm<256>(g, stream);
  return launch_gemm<128>(g, stream);
}
