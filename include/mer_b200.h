/* mer_b200.h — C ABI of libmer_b200.so (B200 / sm_100a only).
 *
 * Drop-in boundary for the MERTools hot path: the reference has no FFI; the
 * seam a maintainer would bind is the model-object call
 *     model(x, output_hidden_states=True).hidden_states  ->  readout
 * in MERBench/feature_extraction/{visual/extract_vision_huggingface.py:140-144,
 * audio/extract_audio_huggingface.py:93-100, text/extract_text_huggingface.py:222-231}
 * and the fusion step of MERBench/main-release.py:31-66.  Every function below cites the
 * reference lines it replaces.  INTEGRATION.md shows the ctypes stub on the reference side.
 *
 * Conventions: all pointers are DEVICE pointers owned by the caller unless a parameter is
 * documented as host; `stream` is a cudaStream_t passed as void*; return 0 on success, non-zero
 * on failure with the message available from mer_last_error().  No global state besides the
 * per-thread error string and cached device attributes.  No CPU fallback exists.
 */
#ifndef MER_B200_H_
#define MER_B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define MER_API __attribute__((visibility("default")))
#else
#define MER_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------- */
MER_API const char* mer_last_error(void);
MER_API int mer_abi_version(void);
/* 0 when the current device is compute capability 10.x, non-zero (and an error string) otherwise */
MER_API int mer_check_device(void);

/* ---- GEMM (nn.Linear / Conv1d-as-GEMM / patch-embed) ---------------------------------- */
enum { MER_EPI_GELU = 1, MER_EPI_ROUND_TF32 = 2 };

typedef struct MerGemmEpilogue {
  const float* bias; /* [N] or NULL */
  const float* res;  /* residual rows or NULL */
  float* out;
  long long out_bstride; /* out row = b*out_bstride + out_row0 + m */
  long long out_row0;
  long long res_bstride; /* res row = b*res_bstride + res_row0 + m */
  long long res_row0;
  int ld_out; /* floats */
  int ld_res;
  int flags; /* MER_EPI_* */
} MerGemmEpilogue;

/* A[b, m, tap*K_inner + c] = base[b*a_batch_stride + (m + tap / P)*a_row_stride +
 *                                 (tap % P)*a_phase_stride + c]        (element strides)
 * Plain row-major A[M,K]: taps = 1, P = 1, K_inner = K, batches = 1, a_row_stride = K.
 * Conv1d(kernel k, stride s) over time-major activations x[b, tau, c]:
 *   taps = k, P = s, K_inner = C_in, a_phase_stride = C_in, a_row_stride = s*C_in.           */
typedef struct MerGemmDesc {
  const float* A;
  const float* W; /* [N, taps*K_inner] row-major (nn.Linear layout) */
  int rows_per_batch;
  int a_rows_dim; /* addressable row groups per batch entry (>= rows_per_batch + (taps-1)/P) */
  int batches;
  int N;
  int K_inner;
  int taps;
  int P;
  long long a_phase_stride;
  long long a_row_stride;
  long long a_batch_stride;
  int force_block_n; /* 0 = auto, 128 or 256 */
  MerGemmEpilogue ep;
} MerGemmDesc;

/* tcgen05 TF32 GEMM with fused epilogue.  Replaces torch nn.Linear / nn.Conv1d calls inside
 * HF ViTLayer / HubertEncoderLayer / BertLayer reached from the reference extractors. */
MER_API int mer_gemm_tf32(const MerGemmDesc* desc, void* stream);

/* ---- row-wise kernels ------------------------------------------------------------------ */
enum { MER_LN_ROUND_TF32 = 1, MER_LN_ACC_INIT = 2, MER_LN_ACC_ADD = 4 };
/* y = LayerNorm(x) * gamma + beta over the last dim (768 or 512).  Optional side buffer acc
 * (same shape): acc = y (ACC_INIT) or acc += y (ACC_ADD) — the "sum of the last four hidden
 * states" readout of extract_audio_huggingface.py:98 / extract_text_huggingface.py:226. */
MER_API int mer_layernorm(const float* x, const float* gamma, const float* beta, float* y, float* acc,
                  long long rows, int dim, float eps, int flags, void* stream);

/* in-place round-to-nearest fp32 -> tf32 (weights at load time) */
MER_API int mer_round_tf32(float* x, long long n, void* stream);

/* ---- attention ------------------------------------------------------------------------- */
/* softmax(Q K^T / 8) V per (sequence, head); head_dim 64.  qkv is [tokens, 3*heads*64] with
 * Q | K | V column blocks, sequences packed back to back, cu_seqlens[n_seq+1] (device, int32).
 * ctx is [tokens, heads*64].  flags: MER_EPI_ROUND_TF32 rounds ctx for the out-proj GEMM.
 * Replaces HF eager/sdpa attention (modeling_vit.py:171-196, modeling_hubert.py:262-345). */
MER_API int mer_attention(const float* qkv, float* ctx, const int32_t* cu_seqlens, int n_seq,
                  int max_seqlen, int heads, int flags, void* stream);

/* ---- segment reduce (readouts) ------------------------------------------------------------ */
enum { MER_SEG_SUM = 0, MER_SEG_MEAN = 1 };
/* out[s, :] = sum or mean of in[begins[s] : ends[s], :] (dim % 4 == 0; begins/ends device int32,
 * n_seg entries each; for back-to-back segments pass offsets and offsets+1).  Empty segments give zeros (extract_text_huggingface.py:236-249 writes zeros
 * for an empty sentence).  Replaces hidden_states[-1].sum(dim=1) / np.mean(axis=0)
 * (extract_vision_huggingface.py:144,187-188; extract_audio_huggingface.py:105-108). */
MER_API int mer_segment_reduce(const float* in, const int32_t* begins, const int32_t* ends, int n_seg,
                               int dim, int mode, float* out, void* stream);

/* ---- transformer encoder stack shared by the three modalities ----------------------------------- */
typedef struct MerLayerWeights {
  const float* ln1_g; /* ViT: layernorm_before | HuBERT: layer_norm | BERT: attention.output.LayerNorm */
  const float* ln1_b;
  const float* w_qkv; /* [2304, 768] = rows Q | K | V, tf32-rounded */
  const float* b_qkv; /* [2304] */
  const float* w_o;   /* [768, 768] */
  const float* b_o;
  const float* ln2_g; /* ViT: layernorm_after | HuBERT: final_layer_norm | BERT: output.LayerNorm */
  const float* ln2_b;
  const float* w_fc1; /* [3072, 768] */
  const float* b_fc1;
  const float* w_fc2; /* [768, 3072] */
  const float* b_fc2;
} MerLayerWeights;

/* ---- ViT-B/16 frame encoder (visual) ------------------------------------------------------------ */
typedef struct MerVitModel {
  int n_layers;        /* 12 */
  float ln_eps;        /* 1e-12 */
  const float* patch_w;   /* [768, 768]  conv weight flattened (c, ph, pw), tf32-rounded */
  const float* patch_b;   /* [768] */
  const float* cls_pos0;  /* [768]  cls_token + position_embeddings[0] */
  const float* pos_rest;  /* [196, 768] position_embeddings[1:] */
  const MerLayerWeights* layers; /* host array of n_layers entries (device pointers inside) */
} MerVitModel;

/* bytes of caller-provided device workspace for n_frames frames */
MER_API long long mer_vit_workspace_bytes(int n_frames);

/* frames: uint8 [n_frames, 224, 224, 3] BGR (the reference's openface_face/<vid>/<vid>.npy layout).
 * Does, on the device: BGR->RGB, x/255, (x-.5)/.5 (HF ViTImageProcessor as invoked at
 * extract_vision_huggingface.py:137-138), the 12-layer pre-LN ViT forward (HF modeling_vit.py),
 * and the readout hidden_states[-1].sum(dim=1) (extract_vision_huggingface.py:143-144).
 * out_frame_feats: [n_frames, 768].  opt_hidden: NULL or [(n_layers+1), n_frames*197, 768] to
 * receive every hidden state (parity tests). */
MER_API int mer_vit_forward(const MerVitModel* model, const uint8_t* frames_bgr, int n_frames,
                            void* workspace, long long workspace_bytes, float* out_frame_feats,
                            float* opt_hidden, void* stream);

/* ---- HuBERT-base audio encoder ------------------------------------------------------------------ */
typedef struct MerHubertModel {
  int n_layers;  /* 12 (>= 4: the readout sums the last four hidden states) */
  float ln_eps;  /* 1e-5 */
  const float* conv0_w;    /* [512, 10] */
  const float* gn_g;       /* GroupNorm(512 groups) affine, [512] */
  const float* gn_b;
  const float* conv_w[6];  /* conv1..6, [512, k*512] laid out [out][tap][in], tf32-rounded */
  const float* fp_ln_g;    /* feature_projection.layer_norm [512] */
  const float* fp_ln_b;
  const float* fp_w;       /* [768, 512] tf32-rounded */
  const float* fp_b;
  const float* pos_w;      /* [16][128][48][48] = [group][tap][out][in], weight-norm folded, tf32 */
  const float* pos_b;      /* [768] */
  const float* enc_ln_g;   /* encoder.layer_norm */
  const float* enc_ln_b;
  const MerLayerWeights* layers;
} MerHubertModel;

/* frames produced for n_samples input samples (conv kernels 10,3,3,3,3,2,2 / strides 5,2,2,2,2,2,2) */
MER_API int mer_hubert_num_frames(int n_samples);
MER_API long long mer_hubert_workspace_bytes(int batch, int n_samples);

/* wave: fp32 [batch, n_samples] raw samples (every row the same length; the reference feeds one
 * clip at a time, or 10 s rows from split_into_batch, extract_audio_huggingface.py:40-50,95).
 * normalize != 0 applies the Wav2Vec2FeatureExtractor zero-mean/unit-variance step per row (:94).
 * Then HubertModel forward (HF modeling_hubert.py) and the readout
 * torch.stack(hidden_states)[[-4,-3,-2,-1]].sum(0) (:98).
 * out_frames: NULL or [batch*T, 768] (FRAME level, :100); out_utt: NULL or [batch, 768] = mean over
 * each row's T frames (UTTERANCE level for clips <= 10 s, :105-108). */
MER_API int mer_hubert_forward(const MerHubertModel* model, const float* wave, int batch, int n_samples,
                               int normalize, void* workspace, long long workspace_bytes,
                               float* out_frames, float* out_utt, float* opt_hidden, void* stream);

/* ---- BERT / RoBERTa-base text encoder ------------------------------------------------------------ */
typedef struct MerBertModel {
  int n_layers;
  float ln_eps;             /* 1e-12 (BERT) / 1e-5 (roberta-base checkpoint) */
  const float* word_emb;    /* [V, 768] */
  const float* pos_emb;     /* [P, 768] */
  const float* type_emb0;   /* token_type_embeddings[0], [768] */
  const float* emb_ln_g;
  const float* emb_ln_b;
  const MerLayerWeights* layers;
} MerBertModel;

MER_API long long mer_bert_workspace_bytes(int tokens, int n_seq);

/* Packed variable-length batch of tokenised sentences (ids from the HF tokenizer on the host, as in
 * extract_text_huggingface.py:222).  ids/pos_ids: device int32 [tokens]; cu_seqlens: device int32
 * [n_seq+1]; seg_begins/seg_ends: device int32 [n_seq], the token range kept by the reference's
 * outputs[0, start:end] slice (:228-231).  out_tokens: NULL or [tokens, 768] = sum of the last four
 * hidden states (:226); out_utt: NULL or [n_seq, 768] = mean over the kept range (:243-249). */
MER_API int mer_bert_forward(const MerBertModel* model, const int32_t* ids, const int32_t* pos_ids,
                             const int32_t* cu_seqlens, int n_seq, int tokens, int max_seqlen,
                             const int32_t* seg_begins, const int32_t* seg_ends, void* workspace,
                             long long workspace_bytes, float* out_tokens, float* out_utt,
                             float* opt_hidden, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MER_B200_H_ */
