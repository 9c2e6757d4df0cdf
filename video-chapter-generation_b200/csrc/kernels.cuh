// Launchers of the memory-bound / small kernels around the tcgen05 GEMMs (definitions in kernels.cu, attention.cu,
// tail.cu, pack.cu).  `fp32` selects the element type of activations: false = bf16, true = fp32.
#pragma once
#include "../../include/vcg.h"
#include <cuda_runtime.h>
#include <cstdint>

namespace vcg {

// geometry of the zero-padded NHWC4 stem input (pixel (0,0) of the buffer is input pixel (-3,-3))
constexpr int kImg = 224;
constexpr int kStemPad = 3;
constexpr int kStemHp = 230;
constexpr int kStemWp = 240;
constexpr int kStemOut = 112;
constexpr int kBertHidden = 768;
constexpr int kBertHeads = 12;
constexpr int kBertFfn = 3072;
constexpr int kVisionDim = 2048;

// n_frames > 0: source frame indices are clamped to [0, n_frames) (the kernel never reads outside the frame buffer)
void launch_preprocess_u8(const uint8_t* frames, const int32_t* frame_index, int n, void* out, bool fp32,
                          cudaStream_t s, int n_frames = 0);
// clip-structured gather: image i = (b, t) reads frame clip_start[b] + t
// resize.cu: bilinear resize (PIL / torchvision BILINEAR, bit-identical) fused with the pre-processing
int resize_ksize(int in_size, int out_size);
void launch_resize_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* kk, cudaStream_t s);
void launch_resize_preprocess_u8(const uint8_t* frames, const int32_t* frame_index, const int32_t* clip_start, int T, int n,
                                 int n_frames, int Hs, int Ws, const int32_t* xb, const int32_t* xk, const int32_t* yb,
                                 const int32_t* yk, void* out_stem, uint8_t* out_u8, bool fp32, cudaStream_t s);
void launch_preprocess_u8_clips(const uint8_t* frames, const int32_t* clip_start, int B, int T, void* out, bool fp32,
                                cudaStream_t s, int n_frames = 0);
void launch_nchw_to_stem(const float* img, int n, void* out, bool fp32, cudaStream_t s);
void launch_maxpool_tsm(const void* in, int n, void* out, void* out_shifted, int T, int fold, bool fp32,
                        cudaStream_t s);
// out: fp32 [n, C]; out_act (optional): the same values in the activation type (GEMM operand of the vision projection)
void launch_avgpool(const void* in, int n, int hw, int C, float* out, void* out_act, cudaStream_t s, bool fp32);
// token packing for variable-length BERT (kernels.cu): cu [B+1], tok_src / key_ok [B*L], total [1]
void launch_bert_pack(const int64_t* mask, int B, int L, int32_t* cu, int32_t* tok_src, uint8_t* key_ok, int32_t* total,
                      cudaStream_t s);
// out[b] = x[row_of ? row_of[b] : b*stride]  (rows of 768)
void launch_gather_rows768(const void* x, const int32_t* row_of, int stride, int B, void* out, bool fp32, cudaStream_t s);
// tok_src / rows_dev are nullptr in the un-packed layout (row m = token m of the [B, L] matrix)
// token ids are clamped to [0, vocab)
void launch_bert_embed_ln(const int64_t* ids, int rows, int L, const int32_t* tok_src, const int32_t* rows_dev, int vocab,
                          const void* word, const void* pos, const void* type, const float* gamma, const float* beta,
                          void* out, bool fp32, cudaStream_t s);
// bf16 path, LayerNorm folded into the GEMMs: out[b] = LayerNorm(x[row_of[b]]) for the B pooled rows
void launch_gather_ln_rows768(const void* x, const int32_t* row_of, int stride, int B, const float* gamma, const float* beta,
                              float eps, void* out, cudaStream_t s);
// W' = bf16(W * gamma) [N][K], c1[n] = sum_k W'[n,k], c2b[n] = bias[n] + sum_k beta[k] W[n,k]
void launch_fold_ln_linear(const float* w, const float* gamma, const float* beta, const float* bias, int N, int K, void* w_out,
                           float* c1, float* c2b, cudaStream_t s);
void launch_layernorm(const void* x, const float* gamma, const float* beta, void* y, int rows, int cols, float eps,
                      bool fp32, cudaStream_t s, const int32_t* rows_dev = nullptr);
// cu / key_ok: packed layout (nullptr: rows b*L.., int64 mask)
// qkv_rows: rows of the qkv allocation that may be read (> 0 enables the tcgen05 kernel for packed bf16, L <= 128)
void launch_bert_attention(const void* qkv, const int64_t* mask, const int32_t* cu, const uint8_t* key_ok, void* ctx, int B,
                           int L, bool fp32, cudaStream_t s, long qkv_rows = 0, const void* items = nullptr,
                           const int32_t* n_items = nullptr);
void launch_bert_attention_tc(const void* qkv, const int32_t* cu, const uint8_t* key_ok, void* ctx, int B, long rows,
                              cudaStream_t s, const void* items = nullptr, const int32_t* n_items = nullptr);   // attention_tc.cu
void launch_attention_items(const int32_t* cu, int B, void* items, int32_t* n_items, cudaStream_t s);

struct TailParams {
  // inputs
  int T, H;                // frames per clip, head hidden size (128)
  const void* lang_out;    // [B, H]    relu(W_l lang_emb)       activation type (bf16 / fp32), written by a GEMM
  const void* vis_out;     // [B*T, H]  relu(W_v vision_emb[t])  fp32, written by a 3xTF32 GEMM
  int head_type;           // 0 mlp, 1 attn
  // head weights (fp32; *_t are transposed to [in][out])
  const float* head_w;     // mlp: [2][(T+1)*H]
  const float* head_b;     // mlp: [2]
  const float* q_w_t; const float* q_b;   // attn: [H][H], [H]
  const float* k_w_t; const float* k_b;
  const float* v_w_t; const float* v_b;
  const float* proj_w; const float* proj_b;  // [2][H], [2]
  // outputs
  float* logits;           // [B,2]
  float* probs;            // [B,2]
};
void launch_head_final(const TailParams& p, int B, bool fp32, cudaStream_t s);     // head + softmax
// single-modality heads: logits = x[B,D] w[2,D]^T + b, probs = softmax (all fp32)
void launch_linear_head2(const float* x, const float* w, const float* b, int B, int D, float* logits, float* probs,
                         cudaStream_t s);

// post-processing (postprocess.cu): labels / run-length cut points per video, precision-recall hit counts
void launch_cut_points(const float* logits, const int32_t* offsets, int n_videos, int clip_frames, int max_offset, int cap,
                       int32_t* labels_out, int32_t* cuts, int32_t* counts, cudaStream_t s);
void launch_pr_hits(const int32_t* gt, const int32_t* gt_off, const int32_t* pred, const int32_t* pred_off, int n_videos,
                    int32_t* hits, cudaStream_t s);

void launch_auc_ap(const float* scores, const int32_t* labels, const int32_t* offsets, int n_videos, double* auc, double* ap,
                   cudaStream_t s);

// window model, post-backbone part (window.cu); parameter structs are those of include/vcg.h
void launch_mlp_chain(const float* x0, int dim0, long stride0, const float* x1, int dim1, long stride1, int rows,
                      const vcg_mlp_op* ops, int n_ops, float* out, long out_stride, cudaStream_t s);
void launch_cross_attention(const vcg_cross_attn_params& p, const float* lang, const float* vision, int B, int T, float* out,
                            cudaStream_t s);
void launch_self_attention_first(const vcg_self_attn_params& p, const float* vision, const float* lang, int B, int T,
                                 float* out, cudaStream_t s);
void launch_bilinear_contract(const float* y, const float* x1, const float* bias, int rows, int in1, int out_features,
                              float* out, cudaStream_t s);
void launch_center_attention(const vcg_center_attn_params& p, const float* x, int B, int W, float* out, cudaStream_t s);
void launch_window_stack(const vcg_window_stack_params& p, const float* x, int B, int W, float* logits, float* probs,
                         cudaStream_t s);

// weight packing (fp32 state-dict tensors -> kernel layouts)
void launch_pack_conv(const float* w /*[Cout,Cin,k,k]*/, const float* bn_w, const float* bn_b, const float* bn_mean,
                      const float* bn_var, float eps, int Cout, int Cin, int k, void* w_out /*[Cout][k][k][Cin]*/,
                      float* bias_out, bool fp32, cudaStream_t s);
void launch_pack_conv1_shared(const float* w /*[Cout,Cin,1,1]*/, const float* bn_w, const float* bn_b, const float* bn_mean,
                              const float* bn_var, float eps, int Cout, int Cin, int fold, void* w_out /*[Cout][3][Cin] bf16*/,
                              float* bias_out, cudaStream_t s);
void launch_pack_stem(const float* w /*[64,3,7,7]*/, const float* bn_w, const float* bn_b, const float* bn_mean,
                      const float* bn_var, float eps, void* w_out /*[64][R][8][4]*/, float* bias_out, bool fp32,
                      cudaStream_t s);
void launch_convert(const float* in, void* out, long n, bool fp32, cudaStream_t s);       // fp32 -> T copy
void launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t s);    // [r][c] -> [c][r]
void launch_checksum(const void* p, size_t bytes, unsigned long long* out, cudaStream_t s);   // debug
void launch_cast_to_f32(const void* in, float* out, long n, bool fp32, cudaStream_t s);
// batch-statistics BatchNorm (opt-in mode of reference caller #1) and the stand-alone temporal shift
int bn_stats_partials(long rows, int C);       // CTAs of the statistics pass = rows of `partial` ([n][2][C] doubles)
void launch_bn_batch_stats(const void* x, long rows, int C, float eps, double* partial, float* mean, float* rstd, bool fp32,
                           cudaStream_t s);
void launch_bn_apply(const void* x, long rows, int C, const float* mean, const float* rstd, const float* gamma,
                     const float* beta, const void* residual, bool relu, void* out, bool fp32, cudaStream_t s);
void launch_tsm_shift(const void* x, long n, int hw, int C, int T, int fold, void* out, bool fp32, cudaStream_t s);

}  // namespace vcg
