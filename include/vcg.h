/*
 * vcg.h — C ABI of libvcg_b200.so, the sm_100a implementation of the chapter-boundary scoring path of
 * SeoYeonnLee/Video-Chapter-Generation (TwoStream point model: BERT-base text stream + ResNet-50-TSM vision
 * stream + ChapterHead).
 *
 * The reference has no FFI of its own: its boundary is the Python module API
 * (video_chapter_generation/model/fusion/two_stream.py:99-194).  This header is what the Python mirror of that
 * API (video-chapter-generation_b200/model/...) binds with ctypes; every entry point names the reference code it
 * replaces.  All pointers are plain device pointers unless a name ends in _host; no torch types cross the boundary.
 * Every function returning int returns 0 on success and non-zero on failure, in which case vcg_last_error()
 * holds the message (CUDA allocation failures contain the words "out of memory", which the reference's handler in
 * convert2vision_emb.py:208-215 looks for).  An engine is not re-entrant; use one engine per host thread / device.
 */
#ifndef VCG_B200_H
#define VCG_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define VCG_API __attribute__((visibility("default")))
#else
#define VCG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vcg_engine vcg_engine;

enum { VCG_HEAD_MLP = 0, VCG_HEAD_ATTN = 1 };       /* two_stream.py:63-68 head_type "mlp" / "attn"            */
enum { VCG_PREC_BF16 = 0, VCG_PREC_FP32 = 1 };      /* bf16 tensor-core mode / 3xTF32 fp32-accurate mode       */
enum { VCG_VISION_R50TSM = 0, VCG_VISION_NONE = 1 };/* resnet50_tsm.py:15-19 / precomputed vision embeddings   */
enum { VCG_DTYPE_F32 = 0, VCG_DTYPE_I64 = 1 };
/* which reference model the engine is: TwoStream (two_stream.py), or one of the single-modality scorers of
 * --data_mode image / text: Resnet50TSM / Resnet50 (resnet50_tsm.py:68-77, resnet50.py:64-73), BertHugface
 * (bert_hugface.py:98-132, pretrain_stage=False).  The single-modality state dicts use base_model.* / head.* keys. */
enum { VCG_MODALITY_TWO_STREAM = 0, VCG_MODALITY_VISION = 1, VCG_MODALITY_TEXT = 2,
       VCG_MODALITY_EMBED = 3 /* both backbones, no head: vcg_embed (window model, two_stream_window.py:404-428) */ };
enum { VCG_ACT_NONE = 0, VCG_ACT_RELU = 1, VCG_ACT_GELU = 2, VCG_ACT_TANH = 3 };

typedef struct vcg_config {
  int32_t clip_frames;   /* T: segment_size of TwoStream / n_segment of TemporalShift (two_stream.py:100-107)   */
  int32_t max_tokens;    /* L: longest text_ids row the engine will see (reference: max_text_len = 100)         */
  int32_t hidden_size;   /* ChapterHead hidden size, 128 in every reference caller                              */
  int32_t head_type;     /* VCG_HEAD_*                                                                          */
  int32_t precision;     /* VCG_PREC_*                                                                          */
  int32_t vision;        /* VCG_VISION_*                                                                        */
  int32_t max_batch;     /* clips scored per internal pass (workspace is sized for this)                        */
  int32_t shift_div;     /* TSM fold divisor, 8 in the reference (resnet50_tsm.py:16); 0 = plain ResNet-50      */
  int32_t modality;      /* VCG_MODALITY_*                                                                      */
} vcg_config;

/* Lifetime ------------------------------------------------------------------------------------------------- */
VCG_API int vcg_create(const vcg_config* cfg, vcg_engine** out);
VCG_API void vcg_destroy(vcg_engine* e);
VCG_API const char* vcg_last_error(void);
VCG_API const char* vcg_version(void);

/* Weights: one call per entry of TwoStream.state_dict() (key schema: SURVEY.md 8b; replaces
 * nn.Module.load_state_dict at test_video_segment_point.py:102-105).  dev_ptr is fp32 (or int64 for
 * num_batches_tracked, ignored).  The data is copied/repacked; the caller keeps ownership.
 * vcg_finalize folds BatchNorm (eval mode, running statistics) into the conv weights, packs everything into the
 * kernels' layouts and checks that no tensor is missing. */
VCG_API int vcg_load_tensor(vcg_engine* e, const char* key, const void* dev_ptr, const int64_t* shape, int32_t ndim,
                    int32_t dtype, void* stream);
VCG_API int vcg_finalize(vcg_engine* e, void* stream);

/* TwoStream.forward (two_stream.py:172-194).
 *   img_clip     [B,T,3,224,224] fp32, already normalised (test_video_segment_point.py:142-145), or NULL
 *   vision_emb   [B,T,2048] fp32 precomputed embeddings (convert2vision_emb.py output), used when img_clip is NULL
 *   text_ids, attention_mask  [B,L] int64
 *   logits, probs             [B,2] fp32 out
 *   vision_emb_out [B,T,2048] fp32 out or NULL;  lang_emb_out [B,768] fp32 out or NULL   (return_emb=True)
 * Work is enqueued on `stream` (a cudaStream_t); B may exceed max_batch (processed in passes). */
VCG_API int vcg_forward(vcg_engine* e, const float* img_clip, const float* vision_emb, const int64_t* text_ids,
                const int64_t* attention_mask, int32_t B, int32_t L, float* logits, float* probs,
                float* vision_emb_out, float* lang_emb_out, void* stream);

/* Single-modality scorers (engine created with VCG_MODALITY_VISION / VCG_MODALITY_TEXT).
 *   vcg_forward_vision = Resnet50TSM.forward / Resnet50.forward (resnet50_tsm.py:68-77, resnet50.py:64-73):
 *       img_clip [B,T,3,224,224] fp32 -> backbone -> [B, T*2048] -> head Linear(T*2048, 2) -> softmax
 *   vcg_forward_text   = BertHugface.forward with pretrain_stage=False (bert_hugface.py:98-132):
 *       pooler_output [B,768] -> head Linear(768, 2) -> softmax
 * logits, probs [B,2] fp32 out; vision_emb_out [B,T,2048] / lang_emb_out [B,768] fp32 out or NULL. */
VCG_API int vcg_forward_vision(vcg_engine* e, const float* img_clip, int32_t B, float* logits, float* probs,
                       float* vision_emb_out, void* stream);
VCG_API int vcg_forward_text(vcg_engine* e, const int64_t* text_ids, const int64_t* attention_mask, int32_t B, int32_t L,
                     float* logits, float* probs, float* lang_emb_out, void* stream);

/* Backbone embeddings only (engine created with VCG_MODALITY_EMBED; state-dict keys lang_model.* / vision_model.*):
 * what the window model computes per clip before its own heads (two_stream_window.py:404-428).
 *   vision_emb_out [B,T,2048], lang_emb_out [B,768] (BertPooler output), both fp32. */
VCG_API int vcg_embed(vcg_engine* e, const float* img_clip, const int64_t* text_ids, const int64_t* attention_mask, int32_t B,
              int32_t L, float* vision_emb_out, float* lang_emb_out, void* stream);
/* The same from decoded uint8 HWC frames [n_frames,224,224,3] (device): clip b = frames clip_start[b] .. +T-1, or, with
 * clip_start == NULL, the regular grid first_start + b*clip_stride (the ResNet stem then runs once per distinct frame). */
VCG_API int vcg_embed_u8(vcg_engine* e, const uint8_t* frames_u8, int32_t n_frames, const int32_t* clip_start,
                 int32_t first_start, int32_t clip_stride, const int64_t* text_ids, const int64_t* attention_mask,
                 int32_t B, int32_t L, float* vision_emb_out, float* lang_emb_out, void* stream);

/* Sliding-window scoring of one video straight from decoded frames (replaces ToTensor+Normalize at
 * test_video_segment_point.py:142-145, the clip gather of infer_youtube_video_dataset.py:117 and the forward).
 *   frames_u8   [n_frames,224,224,3] uint8 HWC (device)
 *   clip_start  [B] int32 (device): first frame of every candidate clip; clip b = frames start..start+T-1
 *   the rest as vcg_forward. */
VCG_API int vcg_score_clips_u8(vcg_engine* e, const uint8_t* frames_u8, int32_t n_frames, const int32_t* clip_start,
                       const int64_t* text_ids, const int64_t* attention_mask, int32_t B, int32_t L, float* logits,
                       float* probs, void* stream);

/* Source frame size of every uint8 entry point of this engine (vcg_score_*_u8*, vcg_embed_u8): frames_u8 is then
 * [n_frames, src_h, src_w, 3] and every frame is resized to 224 x 224 inside the pre-processing kernel with the bilinear
 * filter the reference applies to PIL frames (GroupScale, data/transforms.py:79-92 = torchvision Resize(BILINEAR) =
 * PIL.Image.resize, bit-identical on the uint8 result).  Default 224 x 224 = no resize (the frames ffmpeg -s 224x224
 * produced, video_chapter_youtube_dataset/extract_video_to_frames.py:28). */
VCG_API int vcg_set_frame_size(vcg_engine* e, int32_t src_h, int32_t src_w, void* stream);

/* Same with the clip starts ALSO given on the host (clip_start_host [B] int32, same values as the device array): the
 * engine plans its vision passes from them, so every run of overlapping clips on a regular grid (each video of a
 * video-major clip list: infer_youtube_video_dataset.py:117) shares pre-processing / stem / max-pool between its clips,
 * and the starts are bounds-checked before anything is enqueued.  Everything else stays on the device. */
VCG_API int vcg_score_clips_u8_planned(vcg_engine* e, const uint8_t* frames_u8, int32_t n_frames, const int32_t* clip_start,
                               const int32_t* clip_start_host, const int64_t* text_ids, const int64_t* attention_mask,
                               int32_t B, int32_t L, float* logits, float* probs, void* stream);

/* Same for clips on a regular grid — clip b = frames first_start + b*clip_stride .. +T-1, the reference's
 * range(0, n_frames - T, 4) (infer_youtube_video_dataset.py:117).  Overlapping clips share frames, so pre-processing,
 * the ResNet stem, the max-pool and layer1.0's downsample run once per UNIQUE frame (TSM makes everything after
 * layer1.0.conv1 clip-specific).  Results equal vcg_score_clips_u8 up to fp32 summation order. */
VCG_API int vcg_score_video_u8(vcg_engine* e, const uint8_t* frames_u8, int32_t n_frames, int32_t first_start,
                       int32_t clip_stride, const int64_t* text_ids, const int64_t* attention_mask, int32_t B, int32_t L,
                       float* logits, float* probs, void* stream);

/* Same, with HOST buffers (pinned for full speed): the host->device copies of frames / ids / mask and the
 * device->host copy of logits+probs are issued inside the call; returns after the results are on the host. */
VCG_API int vcg_score_clips_u8_host(vcg_engine* e, const uint8_t* frames_u8_host, int32_t n_frames,
                            const int32_t* clip_start_host, const int64_t* text_ids_host,
                            const int64_t* attention_mask_host, int32_t B, int32_t L, float* logits_host,
                            float* probs_host, void* stream);

/* TwoStream.forward with HOST buffers (pinned for full speed): img_clip_host [B,T,3,224,224] fp32 or, when it is
 * NULL, vision_emb_host [B,T,2048] fp32; ids / mask [B,L] int64; results land in logits_host / probs_host [B,2].
 * All host<->device copies are issued inside the call, which returns once the results are on the host. */
VCG_API int vcg_forward_host(vcg_engine* e, const float* img_clip_host, const float* vision_emb_host,
                     const int64_t* text_ids_host, const int64_t* attention_mask_host, int32_t B, int32_t L,
                     float* logits_host, float* probs_host, void* stream);

/* Per-kernel CUDA-event profile (bench.py's roofline): between begin and end every launch is bracketed by events
 * on the launching stream; end synchronises and returns one entry per "<kernel>|<layer>" name with the summed
 * device time and the algorithmic FLOPs / bytes of those launches. */
typedef struct vcg_profile_entry {
  char name[64];
  int64_t launches;
  double ms;
  double flops;
  double bytes;
} vcg_profile_entry;
VCG_API int vcg_profile_begin(vcg_engine* e);
VCG_API int vcg_profile_end(vcg_engine* e, void* stream, vcg_profile_entry* out, int32_t max_entries, int32_t* n_out);

/* Number of kernel launches issued by this engine since creation (bench.py's gpu_launches). */
VCG_API int64_t vcg_launch_count(const vcg_engine* e);

/* Debugging aid (engines created while VCG_DEBUG_CHECKSUM=1 is set): 64-bit checksums of the stem input and of every
 * vision-stream kernel's output (and shifted copy) of the LAST vision pass, in launch order - tools/stress_checksums.py
 * uses them to find the first kernel whose output differs between two runs on identical inputs. */
/* Debug: the launches number [from, to) counted from this call go out without the programmatic-dependent-launch attribute
 * (ordinary stream order), to bisect an ordering problem to one kernel boundary; (0, 0) switches the window off. */
VCG_API int vcg_debug_pdl_window(int32_t from, int32_t to);
VCG_API int vcg_debug_checksums(vcg_engine* e, uint64_t* out_host, int32_t max_n, int32_t* n_out, void* stream);

/* Stand-alone operators (what the engine is built from; used by the parity tests) ------------------------- */

/* uint8 HWC frames -> normalised, zero-padded stem input (4 channels per pixel, channel 3 = 0; pixel (0,0) of the
 * buffer is input pixel (-3,-3)): fp32 plain NHWC4 [n,230,240,4]; bf16 with row pairs interleaved per pixel
 * [n,115,240,2,4] (one stem K block = 8 pixels x 2 rows = 128 contiguous bytes):
 * (x/255 - mean_c)/std_c, the ToTensor+Normalize of test_video_segment_point.py:142-145.  frame_index may be
 * NULL (identity) or [n] int32 frame numbers to gather. */
VCG_API int vcg_op_preprocess_u8(const uint8_t* frames_u8, const int32_t* frame_index, int32_t n, void* out_padded,
                         int32_t precision, void* stream);
/* fp32 NCHW [n,3,224,224] -> the same padded NHWC4 layout. */
/* Stand-alone resize: uint8 HWC [n,src_h,src_w,3] -> 224 x 224, written as uint8 HWC [n,224,224,3] (out_u8) and / or as
 * the normalised zero-padded stem input (out_padded, as vcg_op_preprocess_u8); either may be NULL. */
VCG_API int vcg_op_resize_u8(const uint8_t* frames_u8, int32_t n, int32_t src_h, int32_t src_w, uint8_t* out_u8, void* out_padded,
                     int32_t precision, void* stream);
VCG_API int vcg_op_nchw_to_stem(const float* img, int32_t n, void* out_padded, int32_t precision, void* stream);

/* out[M,N] = act(A[M,K] W[N,K]^T + bias (+ residual)) on tcgen05 tensor cores; element type bf16 (precision 0) or
 * fp32 via 3xTF32 (precision 1); bias fp32.  lda/ld_res/ld_out in elements. */
VCG_API int vcg_op_gemm(const void* A, int64_t lda, const void* W, const float* bias, const void* residual, int32_t ld_res,
                void* out, int32_t ld_out, int32_t M, int32_t N, int32_t K, int32_t act, int32_t precision,
                void* stream);

/* NHWC convolution (k = 1 or 3, stride 1 or 2, pad k/2) as an implicit GEMM; weights [Cout][k][k][Cin].
 * tsm_in (optional, 1x1 only): buffer [n*H*W, tsm_in_ch] that replaces input channels [0,tsm_in_ch) — the
 * temporally shifted channels of ops/temporal_shift.py:34-51.  tsm_out (optional): the kernel also scatters output
 * channels [0,fold) to frame t-1 and [fold,2*fold) to frame t+1 of this buffer [n*Ho*Wo, 2*fold]. */
VCG_API int vcg_op_conv2d_nhwc(const void* in, int32_t n, int32_t H, int32_t W, int32_t Cin, const void* weight,
                       const float* bias, const void* residual, void* out, int32_t Cout, int32_t k, int32_t stride,
                       int32_t act, int32_t precision, const void* tsm_in, int32_t tsm_in_ch, void* tsm_out,
                       int32_t tsm_fold, int32_t clip_frames, void* stream);

/* ResNet stem conv (7x7/2, folded BN, ReLU) over the padded stem input; weight bf16 [64][4 row pairs][8 px][2 rows][4 ch],
 * fp32 [64][7 rows][8 px][4 ch] (zero for row 7, pixel 7, channel 3); out NHWC [n,112,112,64]. */
/* Fused tail of a ResNet bottleneck (layer1 / layer2, bf16): relu(conv3x3(in, w2) + bias2) -> relu(conv1x1(., w3) + bias3
 * + residual), optionally scattering the next bottleneck's temporally shifted channels (as vcg_op_conv2d_nhwc).
 *   in [n,H,W,P], w2 [P,3,3,P], w3 [4P,P] bf16; bias2 [P], bias3 [4P] fp32; residual / out [n,H/stride,W/stride,4P] bf16
 *   variant 0: per-tap TMA boxes (conv23.cuh, P = 64 / 128); 1: halo patch resident in shared memory (conv23h.cuh,
 *   P = 64, stride 1). */
VCG_API int vcg_op_bottleneck_tail(const void* in, int32_t n, int32_t H, int32_t W, int32_t P, int32_t stride, const void* w2,
                           const float* bias2, const void* w3, const float* bias3, const void* residual, void* out,
                           void* tsm_out, int32_t tsm_fold, int32_t clip_frames, int32_t variant, void* stream);

VCG_API int vcg_op_stem_conv(const void* in_padded, int32_t n, const void* weight, const float* bias, void* out,
                     int32_t precision, void* stream);

/* 3x3/2 max-pool NHWC [n,112,112,64] -> x [n,56,56,64] and its temporally shifted copy (fold = 64/shift_div). */
VCG_API int vcg_op_maxpool_tsm(const void* in, int32_t n, void* out, void* out_shifted, int32_t clip_frames,
                       int32_t shift_div, int32_t precision, void* stream);

/* ---- Batch-statistics BatchNorm mode (opt-in; reference caller #1 nulls the running statistics of every BatchNorm2d after
 * .eval(), test_video_segment_point.py:116-122, so F.batch_norm normalises with the statistics of the batch it is given).
 * The mirror composes the vision stream of ONE forward call from these operators (vcg_b200/bn_batch.py); all pointers are
 * device pointers, activations NHWC in the given precision. */
/* conv 7x7/2 of the padded stem input with an explicit activation (VCG_ACT_*) and an optional bias (may be NULL). */
VCG_API int vcg_op_stem_conv_act(const void* in_padded, int32_t n, const void* weight, const float* bias, void* out,
                                 int32_t act, int32_t precision, void* stream);
/* Rows of the workspace vcg_op_bn_batch_stats needs: partial is [rows_of_partial][2][C] doubles. */
VCG_API int32_t vcg_op_bn_partials(int64_t rows, int32_t C);
/* Per-channel mean and 1 / sqrt(biased variance + eps) of x [rows, C] (C = 64 * 2^k <= 2048), deterministic. */
VCG_API int vcg_op_bn_batch_stats(const void* x, int64_t rows, int32_t C, float eps, double* partial, float* mean,
                                  float* rstd, int32_t precision, void* stream);
/* out = relu?((x - mean) * rstd * gamma + beta (+ residual)) over x [rows, C]. */
VCG_API int vcg_op_bn_apply(const void* x, int64_t rows, int32_t C, const float* mean, const float* rstd,
                            const float* gamma, const float* beta, const void* residual, int32_t relu, void* out,
                            int32_t precision, void* stream);
/* TemporalShift.shift (ops/temporal_shift.py:34-51) of x [n, hw, C], n = clips * clip_frames, fold = C / shift_div. */
VCG_API int vcg_op_tsm_shift(const void* x, int64_t n, int32_t hw, int32_t C, int32_t clip_frames, int32_t fold, void* out,
                             int32_t precision, void* stream);
/* AdaptiveAvgPool2d(1) of x [n, hw, C] -> fp32 [n, C]. */
VCG_API int vcg_op_avgpool(const void* x, int32_t n, int32_t hw, int32_t C, float* out, int32_t precision, void* stream);

/* softmax(Q K^T / 8 + key mask) V for BERT: qkv [B*L, 2304] (Q|K|V, 12 heads x 64) -> ctx [B*L, 768]. */
VCG_API int vcg_op_bert_attention(const void* qkv, const int64_t* attention_mask, void* ctx, int32_t B, int32_t L,
                          int32_t precision, void* stream);

/* The same attention on the token-packed layout the engine uses (bf16 only): clip b owns rows cu[b] .. cu[b+1]-1 of
 * qkv [rows, 2304] / ctx [rows, 768]; key_ok[row] != 0 marks keys that may be attended; max_len >= every clip length.
 * `rows` = rows of the qkv allocation (>= cu[B] + 31, finite values everywhere).  max_len <= 128 runs the tcgen05
 * kernel (one 128 x 128 score tile per clip and head), longer clips the mma.sync kernel with an online softmax. */
VCG_API int vcg_op_bert_attention_packed(const void* qkv, const int32_t* cu, const uint8_t* key_ok, void* ctx, int32_t B,
                                 int32_t max_len, int64_t rows, void* stream);

/* Post-processing on the device, bit-identical to the reference's Python (all buffers device memory).
 * vcg_op_cut_points: video v owns clips [video_offsets[v], video_offsets[v+1]) of logits [N,2];
 *   labels_out [N] (or NULL) = argmax as logits.topk(1) (test_video_segment_point.py:201-203);
 *   cut_points [n_videos, cap], counts [n_videos] = convert_clip_label2cut_point(labels, clip_frames, max_offset)
 *   (eval_utils/eval_utils.py:3-18: half-to-even rounding, trailing run dropped).  counts[v] > cap: buffer too small.
 * vcg_op_pr_hits: per video the six hit counts of calculate_pr (eval_utils.py:21-92):
 *   hits [n_videos, 6] = {gt->pred exact, <=3 s, <=5 s, pred->gt exact, <=3 s, <=5 s}. */
/* vcg_op_auc_ap: per video, over its clips' scores (softmax prob of class 1) and 0/1 ground-truth labels:
 *   auc[v] = sklearn.metrics.auc(*roc_curve(labels, scores, pos_label=1)[:2]) (nan when a video has one class only),
 *   ap[v]  = sklearn.metrics.average_precision_score(labels, scores) (0 without positives)
 *   (test_video_segment_point.py:253-257, 299-303; scikit-learn is an unpinned dependency, requirements.txt:16). */
VCG_API int vcg_op_auc_ap(const float* scores, const int32_t* labels, const int32_t* video_offsets, int32_t n_videos,
                  double* auc, double* ap, void* stream);
VCG_API int vcg_op_cut_points(const float* logits, const int32_t* video_offsets, int32_t n_videos, int32_t clip_frames,
                      int32_t max_offset, int32_t cap, int32_t* labels_out, int32_t* cut_points, int32_t* counts,
                      void* stream);
VCG_API int vcg_op_pr_hits(const int32_t* gt, const int32_t* gt_offsets, const int32_t* pred, const int32_t* pred_offsets,
                   int32_t n_videos, int32_t* hits, void* stream);

/* ---- window ("update") model, post-backbone part (two_stream_window.py, stacked_window_self_attention.py) --------
 * All tensors fp32 device memory; weights are the reference's nn.Linear / nn.LayerNorm parameters as they are
 * (Linear weight [out, in] row-major).  The host mirror (model/fusion/two_stream_window.py) strings these together. */
enum { VCG_MLP_LINEAR = 0, VCG_MLP_LAYERNORM = 1, VCG_MLP_RELU = 2, VCG_MLP_GELU = 3,
       VCG_MLP_MULHALVES = 4,  /* row of 2n -> n: row[i] * row[n+i] ("multiplication" head, two_stream_window.py:277) */
       VCG_MLP_MEANGROUPS = 5, /* row of g*out_dim -> out_dim: mean over the g groups (two_stream_domain_specific.py:344) */
       VCG_MLP_SOFTMAX = 6,    /* softmax over the row */
       VCG_MLP_SAVE = 7,       /* keep a copy of the row (<= 1024 wide) ...                                        */
       VCG_MLP_ADDSAVED = 8 }; /* ... and add it back later: residual connections (window_self_attention.py:165-169) */
typedef struct vcg_mlp_op {
  int32_t type;          /* VCG_MLP_*                                                    */
  int32_t in_dim;        /* LINEAR: input features                                       */
  int32_t out_dim;       /* LINEAR: output features                                      */
  float eps;             /* LAYERNORM: epsilon (nn.LayerNorm default 1e-5)               */
  const void* w;         /* LINEAR: weight [out, in]; LAYERNORM: weight [dim]            */
  const void* b;         /* LINEAR: bias [out] or NULL; LAYERNORM: bias [dim]            */
} vcg_mlp_op;
/* One row per CTA through an nn.Sequential of Linear / LayerNorm / ReLU / GELU (Dropout is the identity in eval):
 * the per-position lang / vision projection heads and the "mlp" fusion head of ChapterHead
 * (two_stream_window.py:146-185, 252-268).  The input row is the concatenation of x0[r] (dim0) and x1[r] (dim1). */
VCG_API int vcg_op_mlp_chain(const float* x0, int32_t dim0, int64_t stride0, const float* x1, int32_t dim1, int64_t stride1,
                     int32_t rows, const vcg_mlp_op* ops, int32_t n_ops, float* out, int64_t out_stride, void* stream);

typedef struct vcg_cross_attn_params {   /* CrossAttention (two_stream_window.py:11-88), hidden size 128 */
  int32_t num_heads;
  const float *lang_norm_w, *lang_norm_b, *vision_norm_w, *vision_norm_b;
  const float *pos_w, *pos_b;            /* frame_pos_encoding = Linear(1, 128): weight [128,1], bias [128] */
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;
} vcg_cross_attn_params;
/* lang [B,128], vision [B,T,128] -> out [B,128] */
VCG_API int vcg_op_cross_attention(const vcg_cross_attn_params* p, const float* lang, const float* vision, int32_t B,
                           int32_t T, float* out, void* stream);

typedef struct vcg_self_attn_params {    /* SelfAttention (two_stream_window.py:91-131), n_embd 128 */
  int32_t num_heads;                     /* 4 in the reference (:237) */
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;   /* query / key / value / proj */
} vcg_self_attn_params;
/* head_type "self_attn" (:281-283): tokens = the T frame vectors of vision [B,T,128] followed by lang [B,128]; unmasked
 * attention, and only the FIRST token's context goes through proj (:129) -> out [B,128] */
VCG_API int vcg_op_self_attention_first(const vcg_self_attn_params* p, const float* vision, const float* lang, int32_t B,
                                int32_t T, float* out, void* stream);

/* Contraction step of nn.Bilinear (head_type "bilinear", :189-191, 269-271): y [rows, out*in1] holds
 * y[r, o*in1 + i] = sum_j W[o,i,j] x2[r,j] (one vcg_op_gemm over the weight viewed as [out*in1, in2]);
 * out[r,o] = bias[o] + sum_i x1[r,i] * y[r, o*in1 + i]. */
VCG_API int vcg_op_bilinear_contract(const float* y, const float* x1, const float* bias, int32_t rows, int32_t in1,
                             int32_t out_features, float* out, void* stream);

typedef struct vcg_center_attn_params {  /* window attention with the CENTRE clip as the only query, hidden size 128:
                                          * WindowSelfAttention (two_stream_domain_specific.py:92-135, of whose output
                                          * only the centre row is used, :354-356) and VideoChapterWindowAttention
                                          * (window_self_attention.py:80-121) */
  int32_t num_heads;
  int32_t bias_head_stride;              /* elements between heads in pos_bias                                      */
  int32_t bias_offset;                   /* first element of the centre query's row: centre*(2w+1) or 0             */
  int32_t add_residual;                  /* 1: out += x[:, centre] (the un-normalised input, window_self_attention.py:160-163) */
  const float *pre_norm_w, *pre_norm_b;  /* LayerNorm BEFORE the positions are added, or NULL                        */
  const float *post_norm_w, *post_norm_b;/* LayerNorm AFTER the positions are added, or NULL                         */
  const float *pos_w, *pos_b;            /* position_encoding.0 = Linear(1, 128)                                     */
  const float *pos_norm_w, *pos_norm_b;  /* position_encoding.1 = LayerNorm(128)                                     */
  const float *pos_bias;                 /* window_pos_bias                                                          */
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b;
  const float *o_w, *o_b;                /* single Linear output projection, or NULL (context returned as is)        */
} vcg_center_attn_params;
/* x [B,W,128] -> out [B,128]; positions (t - W/2) / (W/2 + 1e-6) */
VCG_API int vcg_op_center_attention(const vcg_center_attn_params* p, const float* x, int32_t B, int32_t W, float* out,
                            void* stream);

typedef struct vcg_window_layer {        /* VideoChapterBlock (stacked_window_self_attention.py:99-148) */
  const float *attn_norm_w, *attn_norm_b, *ffn_norm_w, *ffn_norm_b;
  const float *pos_w, *pos_b;            /* position_encoding = Linear(1, 128) */
  const float *pos_bias;                 /* window_pos_bias [1, 16, 1, 2w+1] */
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;
  const float *f0_w, *f0_b, *f1_w, *f1_b, *f2_w, *f2_b, *f3_w, *f3_b;   /* ffn: 128->256->512->256->128 */
} vcg_window_layer;
typedef struct vcg_window_stack_params { /* StackedVideoChapterAttention (:151-223) */
  int32_t num_layers;                    /* 6 in the reference */
  int32_t pos_bias_stride;               /* 2w+1 */
  vcg_window_layer layers[8];
  const float *final_norm_w, *final_norm_b;
  const float *cls_w[5], *cls_b[5];      /* classifier Linears: 128->128->128->64->32->2 */
  const float *cls_norm_w[4], *cls_norm_b[4];
} vcg_window_stack_params;
/* x [B, W, 128] (W = 2w+1 fused clip embeddings) -> logits, probs [B,2] of the middle clip */
VCG_API int vcg_op_window_stack(const vcg_window_stack_params* p, const float* x, int32_t B, int32_t W, float* logits,
                        float* probs, void* stream);

/* y = LayerNorm(x) * gamma + beta over rows of 768, eps 1e-12 (modeling_bert.py BertSelfOutput/BertOutput). */
VCG_API int vcg_op_layernorm(const void* x, const float* gamma, const float* beta, void* y, int32_t rows, int32_t cols,
                     float eps, int32_t precision, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VCG_B200_H */
