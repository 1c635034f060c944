// encoder.cu — host-side launch sequences of the encoder forwards (no kernels here).
//
// mer_run_stack   : the 12-layer transformer stack shared by the three modalities
//                   pre-LN  (HF ViTLayer,            modeling_vit.py:328-346)
//                   post-LN (HF HubertEncoderLayer,  modeling_hubert.py:372-405; BertLayer)
// mer_vit_forward : frames (uint8 BGR) -> patchify -> patch-embed GEMM (+bias +pos) -> stack ->
//                   token-sum readout  (reference: extract_vision_huggingface.py:137-144)
//
// Per layer and token the chain moves (fp32 activations): LN 6 KB x2, QKV 3+9 KB, attention
// 9+3 KB, out-proj 3+3+3 KB, FC1 3+12 KB, FC2 12+3+3 KB  = 72 KB  (see DESIGN.md).
#include "mer_common.cuh"
#include "mer_kernels.h"

namespace {

constexpr int D = 768;
constexpr int DQKV = 2304;
constexpr int DFF = 3072;
constexpr int HEADS = 12;

int linear(const float* A, const float* W, const float* bias, const float* res, float* out,
           long long M, int N, int K, int flags, cudaStream_t stream) {
  MerGemmDesc g;
  memset(&g, 0, sizeof(g));
  g.A = A;
  g.W = W;
  g.rows_per_batch = (int)M;
  g.a_rows_dim = (int)M;
  g.batches = 1;
  g.N = N;
  g.K_inner = K;
  g.taps = 1;
  g.P = 1;
  g.a_phase_stride = K;
  g.a_row_stride = K;
  g.a_batch_stride = (long long)K * M;
  g.ep.bias = bias;
  g.ep.res = res;
  g.ep.out = out;
  g.ep.ld_out = N;
  g.ep.ld_res = N;
  g.ep.flags = flags;
  return mer_gemm_tf32_launch(&g, stream);
}

#define MER_TRY(expr)          \
  do {                         \
    if (int _rc = (expr)) return _rc; \
  } while (0)

}  // namespace

int mer_run_stack(const MerStackArgs& a, cudaStream_t stream) {
  MER_REQUIRE(a.tokens > 0 && a.tokens < (1ll << 31), "mer_run_stack: bad token count %lld", a.tokens);
  const long long M = a.tokens;
  const size_t hs_bytes = (size_t)M * D * sizeof(float);
  if (a.opt_hidden) MER_CUDA_CHECK(cudaMemcpyAsync(a.opt_hidden, a.x, hs_bytes, cudaMemcpyDeviceToDevice, stream));
  for (int l = 0; l < a.n_layers; ++l) {
    const MerLayerWeights& w = a.layers[l];
    const int first_acc = a.n_layers - a.acc_last;  // hidden state index l+1 > first_acc is summed
    if (a.pre_ln) {
      // x = x + Wo * Attn(LN1(x));  x = x + W2 * GELU(W1 * LN2(x))
      MER_TRY(mer_layernorm_launch(a.x, w.ln1_g, w.ln1_b, a.xn, nullptr, M, D, a.eps,
                                   MER_LN_ROUND_TF32, stream));
      MER_TRY(linear(a.xn, w.w_qkv, w.b_qkv, nullptr, a.qkv, M, DQKV, D, MER_EPI_ROUND_TF32, stream));
      MER_TRY(mer_attention_launch(a.qkv, a.xn, a.cu_seqlens, a.n_seq, a.max_seqlen, HEADS,
                                   MER_EPI_ROUND_TF32, stream));
      MER_TRY(linear(a.xn, w.w_o, w.b_o, a.x, a.x, M, D, D, 0, stream));
      MER_TRY(mer_layernorm_launch(a.x, w.ln2_g, w.ln2_b, a.xn, nullptr, M, D, a.eps,
                                   MER_LN_ROUND_TF32, stream));
      MER_TRY(linear(a.xn, w.w_fc1, w.b_fc1, nullptr, a.h, M, DFF, D,
                     MER_EPI_GELU | MER_EPI_ROUND_TF32, stream));
      MER_TRY(linear(a.h, w.w_fc2, w.b_fc2, a.x, a.x, M, D, DFF, 0, stream));
    } else {
      // x = LN1(x + Wo * Attn(x));  x = LN2(x + W2 * GELU(W1 * x))   (x enters tf32-rounded)
      MER_TRY(linear(a.x, w.w_qkv, w.b_qkv, nullptr, a.qkv, M, DQKV, D, MER_EPI_ROUND_TF32, stream));
      MER_TRY(mer_attention_launch(a.qkv, a.xn, a.cu_seqlens, a.n_seq, a.max_seqlen, HEADS,
                                   MER_EPI_ROUND_TF32, stream));
      // the pre-LN sum goes to the (now dead) qkv buffer: ctx in xn is still being read
      MER_TRY(linear(a.xn, w.w_o, w.b_o, a.x, a.qkv, M, D, D, 0, stream));
      MER_TRY(mer_layernorm_launch(a.qkv, w.ln1_g, w.ln1_b, a.x, nullptr, M, D, a.eps,
                                   MER_LN_ROUND_TF32, stream));
      MER_TRY(linear(a.x, w.w_fc1, w.b_fc1, nullptr, a.h, M, DFF, D,
                     MER_EPI_GELU | MER_EPI_ROUND_TF32, stream));
      MER_TRY(linear(a.h, w.w_fc2, w.b_fc2, a.x, a.xn, M, D, DFF, 0, stream));
      int fl = MER_LN_ROUND_TF32;
      float* acc = nullptr;
      if (a.acc && a.acc_last > 0 && l + 1 > first_acc) {
        acc = a.acc;
        fl |= (l + 1 == first_acc + 1) ? MER_LN_ACC_INIT : MER_LN_ACC_ADD;
      }
      MER_TRY(mer_layernorm_launch(a.xn, w.ln2_g, w.ln2_b, a.x, acc, M, D, a.eps, fl, stream));
    }
    if (a.opt_hidden) {
      // post-LN: the hidden state is the un-rounded LayerNorm output; re-derive it exactly from
      // the pre-LN sum still sitting in xn (debug/parity path only)
      float* dst = a.opt_hidden + (size_t)(l + 1) * M * D;
      if (a.pre_ln) {
        MER_CUDA_CHECK(cudaMemcpyAsync(dst, a.x, hs_bytes, cudaMemcpyDeviceToDevice, stream));
      } else {
        MER_TRY(mer_layernorm_launch(a.xn, w.ln2_g, w.ln2_b, dst, nullptr, M, D, a.eps, 0, stream));
      }
    }
  }
  return 0;
}

extern "C" {

long long mer_vit_workspace_bytes(int n_frames) {
  const long long M = (long long)n_frames * 197;
  // x, xn, qkv, h (+ patch operand aliasing h) + offsets
  return (M * (D + D + DQKV + DFF)) * 4 + ((long long)n_frames + 1) * 4 + 1024;
}

int mer_vit_forward(const MerVitModel* m, const uint8_t* frames_bgr, int n_frames, void* workspace,
                    long long workspace_bytes, float* out_frame_feats, float* opt_hidden,
                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MER_REQUIRE(m && frames_bgr && workspace && out_frame_feats, "mer_vit_forward: null operand");
  MER_REQUIRE(n_frames > 0, "mer_vit_forward: n_frames=%d", n_frames);
  MER_REQUIRE(workspace_bytes >= mer_vit_workspace_bytes(n_frames),
              "mer_vit_forward: workspace %lld B < required %lld B", workspace_bytes,
              mer_vit_workspace_bytes(n_frames));
  MER_REQUIRE((long long)n_frames * 197 < (1ll << 31) / 4, "mer_vit_forward: too many frames");
  const long long M = (long long)n_frames * 197;
  float* x = static_cast<float*>(workspace);
  float* xn = x + M * D;
  float* qkv = xn + M * D;
  float* h = qkv + M * DQKV;
  int* offsets = reinterpret_cast<int*>(h + M * DFF);
  float* a_patch = h;  // [n_frames*196, 768] patch operand lives in the (not yet used) FFN buffer

  MER_TRY(mer_iota_offsets_launch(offsets, n_frames, 197, stream));
  MER_TRY(mer_vit_patchify_launch(frames_bgr, n_frames, a_patch, stream));
  MER_TRY(mer_vit_cls_rows_launch(m->cls_pos0, x, n_frames, stream));
  {
    MerGemmDesc g;
    memset(&g, 0, sizeof(g));
    g.A = a_patch;
    g.W = m->patch_w;
    g.rows_per_batch = 196;
    g.a_rows_dim = 196;
    g.batches = n_frames;
    g.N = D;
    g.K_inner = D;
    g.taps = 1;
    g.P = 1;
    g.a_phase_stride = D;
    g.a_row_stride = D;
    g.a_batch_stride = 196ll * D;
    g.ep.bias = m->patch_b;
    g.ep.res = m->pos_rest;  // + position_embeddings[1:], same for every frame
    g.ep.res_bstride = 0;
    g.ep.out = x;
    g.ep.out_bstride = 197;
    g.ep.out_row0 = 1;
    g.ep.ld_out = D;
    g.ep.ld_res = D;
    MER_TRY(mer_gemm_tf32_launch(&g, stream));
  }
  MerStackArgs a;
  memset(&a, 0, sizeof(a));
  a.layers = m->layers;
  a.n_layers = m->n_layers;
  a.pre_ln = 1;
  a.eps = m->ln_eps;
  a.tokens = M;
  a.cu_seqlens = offsets;
  a.n_seq = n_frames;
  a.max_seqlen = 197;
  a.x = x;
  a.xn = xn;
  a.qkv = qkv;
  a.h = h;
  a.opt_hidden = opt_hidden;
  MER_TRY(mer_run_stack(a, stream));
  MER_TRY(mer_segment_reduce_launch(x, offsets, n_frames, D, MER_SEG_SUM, out_frame_feats, stream));
  return 0;
}

}  // extern "C"
