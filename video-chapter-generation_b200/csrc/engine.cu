// libvcg_b200.so — engine and C ABI (include/vcg.h).
//
// The engine owns packed weights and a workspace sized for `max_batch` clips, and builds (once per batch size) a
// "plan": the ordered list of kernel launches — tcgen05 implicit-GEMM convs / GEMMs with their TMA tensor maps and
// the memory-bound kernels between them — that scores a chunk of clips.  See DESIGN.md for the data layout.
#include "../../include/vcg.h"
#include "conv_gemm_host.h"
#include "kernels.cuh"

#include <cuda_bf16.h>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

using namespace vcg;

namespace vcg {
PdlDebugWindow& pdl_debug_window() {
  static PdlDebugWindow w;
  return w;
}
}  // namespace vcg

namespace {

thread_local std::string g_last_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  void alloc(size_t n, bool zero = false) {
    release();
    if (n == 0) n = 16;
    VCG_CUDA(cudaMalloc(&p, n));
    bytes = n;
    if (zero) VCG_CUDA(cudaMemset(p, 0, n));
  }
  void ensure(size_t n) {
    if (n > bytes) alloc(n);
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

struct RawTensor {
  std::unique_ptr<DevBuf> buf;
  std::vector<int64_t> shape;
  int dtype = VCG_DTYPE_F32;
  long numel() const {
    long n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

struct ConvW {
  DevBuf w, bias;
  int Cin = 0, Cout = 0, k = 1, stride = 1;
};
struct Bottleneck {
  ConvW c1, c2, c3, ds;
  ConvW c1_shared;   // block 0 only: conv1 as three temporal taps over per-unique-frame input (bf16, TSM)
  bool has_ds = false;
  int stride = 1, Cin = 0, planes = 0, H = 0;   // H = input spatial size
};
struct LinearW {
  DevBuf w, bias;   // w: [N][K] in the activation element type, bias fp32
  int N = 0, K = 0;
};
struct BertLayerW {
  LinearW qkv, out, ffn1, ffn2;
  DevBuf ln1_g, ln1_b, ln2_g, ln2_b;
  // bf16 path, LayerNorm folded into the consuming GEMMs (kernels.cu fold_ln_linear_kernel): the bias slot of the
  // LinearW holds c2b, `c1` the column sums.  qkv_f folds the PREVIOUS layer's output LayerNorm (absent in layer 0).
  LinearW qkv_f, ffn1_f;
  DevBuf qkv_c1, ffn1_c1;
};

// One step of a plan.
struct Step {
  enum Kind { CONV_GEMM, CONV23, MAXPOOL, AVGPOOL, EMBED, LAYERNORM, ATTENTION, ZERO_STATS } kind;
  ConvGemmLaunch gemm;
  Conv23Launch c23;   // CONV23: fused conv2 + conv3 of a bottleneck
  // generic arguments for the small kernels
  const void* in = nullptr;
  void* out = nullptr;
  void* out2 = nullptr;
  const float* g = nullptr;
  const float* b = nullptr;
  int n = 0, a = 0, c = 0;
};

// CUDA-event profile of one launch (collected between vcg_profile_begin / vcg_profile_end)
struct ProfRec {
  std::string name;     // "<kernel>|<layer>"
  double flops = 0, bytes = 0;
  bool packed = false;   // work scales with the number of packed BERT tokens (resolved at profile end)
  cudaEvent_t start = nullptr, stop = nullptr;
};

struct VisionPlan {
  int B = 0;
  std::vector<Step> steps;   // stem .. last bottleneck (avgpool is issued by the caller: its destination varies)
  const void* final_act = nullptr;
};
struct BertPlan {
  int B = 0, L = 0;
  std::vector<Step> steps;   // everything after the embedding (whose ids pointer varies)
  const void* final_hidden = nullptr;   // last layer's output
  bool final_is_raw = false;            // ... is a raw pre-LayerNorm matrix (LayerNorm folded into the GEMMs)
};

}  // namespace

struct vcg_engine {
  vcg_config cfg{};
  bool fp32 = false;
  bool finalized = false;
  int T = 16, Lmax = 100, H = 128, Bv = 32, Bt = 256;
  int64_t launches = 0;
  std::unordered_map<std::string, RawTensor> raw;

  // vision weights
  ConvW stem;
  std::vector<Bottleneck> blocks;
  bool tsm = true;
  // text weights
  DevBuf word, pos, type, emb_g, emb_b;
  int vocab = 0;
  std::vector<BertLayerW> layers;
  // tail weights (fp32)
  LinearW pooler;                       // [768][768] + bias, activation type (tcgen05 GEMM, tanh epilogue)
  DevBuf lang_w, vis_w;                 // bias-free projections [128][768], [128][2048], activation type
  DevBuf head_w, head_b, q_w_t, q_b, k_w_t, k_b, v_w_t, v_b, proj_w, proj_b;   // head, fp32

  // vision workspace (sized for Bv clips)
  DevBuf stem_in, stem_out, x0, xa, xb, dsbuf, mid1, mid2, vis_emb, vis_emb_act, vis_out, lang_out, pooled, pooled32;
  std::vector<std::unique_ptr<DevBuf>> shifted;   // one per bottleneck: its temporally shifted input channels
  // text workspace (sized for Bt clips x Lmax tokens)
  DevBuf hid, hid2, qkv, ctx, tmp, ffn, cls;
  // token packing (variable-length BERT): cu [Bt+1], tok_src / key_ok [Bt*Lmax], m_total [1]
  DevBuf pk_cu, pk_src, pk_ok, pk_total, pk_items, pk_nitems;   // + work list of the tcgen05 attention kernel
  // LayerNorm folded into the GEMMs (bf16): per layer two (sum, sum of squares) row-statistics arrays
  bool ln_fused = false;
  DevBuf ln_stats;
  size_t ln_stats_rows = 0;
  int last_bert_rows = 0;   // bt*L of the most recent BERT pass (profile scaling)
  // source frame size of the uint8 entry points (224 x 224 = no resize) and the resize coefficient tables
  int src_h = kImg, src_w = kImg;
  DevBuf rs_xb, rs_xk, rs_yb, rs_yk;
  bool resizing() const { return src_h != kImg || src_w != kImg; }
  size_t frame_bytes() const { return static_cast<size_t>(src_h) * src_w * 3; }
  // host-call staging
  DevBuf st_frames, st_ids, st_mask, st_start, st_logits, st_probs;

  std::map<int, VisionPlan> vplans;
  std::map<std::pair<int, int>, VisionPlan> vplans_shared;   // (clips, clip stride): stem computed once per unique frame
  std::map<std::pair<int, int>, BertPlan> bplans;
  std::map<std::pair<int, int>, std::vector<ConvGemmLaunch>> lang_tail_plans;   // (clips, L) -> pooler, lang projection
  std::map<int, ConvGemmLaunch> vis_proj_plans;                                 // frames -> vision projection

  // host-buffer entry points: copies run on a side stream and overlap the compute stream
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  cudaEvent_t copy_event(size_t i) {
    while (copy_events.size() <= i) {
      cudaEvent_t ev;
      VCG_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      copy_events.push_back(ev);
    }
    return copy_events[i];
  }
  cudaStream_t get_copy_stream() {
    if (!copy_stream) VCG_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    return copy_stream;
  }

  // debug (VCG_DEBUG_CHECKSUM=1): checksum of every vision step's output, read back by vcg_debug_checksums
  bool debug_checksum = false;
  DevBuf dbg;
  int dbg_n = 0;
  // profiling
  bool profiling = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  size_t events_used = 0;
  cudaEvent_t next_event() {
    if (events_used == event_pool.size()) {
      cudaEvent_t ev;
      VCG_CUDA(cudaEventCreate(&ev));
      event_pool.push_back(ev);
    }
    return event_pool[events_used++];
  }
  ~vcg_engine() {
    for (auto ev : event_pool) cudaEventDestroy(ev);
    for (auto ev : copy_events) cudaEventDestroy(ev);
    if (copy_stream) cudaStreamDestroy(copy_stream);
  }

  int es() const { return fp32 ? 4 : 2; }
  // state-dict prefixes: TwoStream keeps the backbones as vision_model.* / lang_model.*, the single-modality
  // models (Resnet50TSM, Resnet50, BertHugface) as base_model.*
  std::string vis_prefix() const { return cfg.modality == VCG_MODALITY_VISION ? "base_model." : "vision_model."; }
  std::string lang_prefix() const { return cfg.modality == VCG_MODALITY_TEXT ? "base_model." : "lang_model."; }
};

namespace {

// ------------------------------------------------------------------------------------------------ weights
const RawTensor& need(vcg_engine* e, const std::string& key, std::initializer_list<int64_t> shape = {}) {
  auto it = e->raw.find(key);
  if (it == e->raw.end()) throw Error("vcg: missing state-dict tensor '" + key + "'");
  if (shape.size()) {
    std::vector<int64_t> s(shape);
    if (s != it->second.shape) {
      std::string got;
      for (auto d : it->second.shape) got += std::to_string(d) + ",";
      throw Error("vcg: tensor '" + key + "' has unexpected shape [" + got + "]");
    }
  }
  return it->second;
}
const float* fptr(const RawTensor& t) { return t.buf->as<float>(); }

void copy_f32(DevBuf& dst, const RawTensor& t, cudaStream_t s) {
  dst.alloc(t.numel() * sizeof(float));
  VCG_CUDA(cudaMemcpyAsync(dst.p, t.buf->p, t.numel() * sizeof(float), cudaMemcpyDeviceToDevice, s));
}
void convert_to(DevBuf& dst, const RawTensor& t, bool fp32, cudaStream_t s) {
  dst.alloc(t.numel() * (fp32 ? 4 : 2));
  launch_convert(fptr(t), dst.p, t.numel(), fp32, s);
}
void transpose_to(DevBuf& dst, const RawTensor& t, cudaStream_t s) {   // [r][c] -> [c][r], fp32
  VCG_REQUIRE(t.shape.size() == 2, "transpose expects a matrix");
  dst.alloc(t.numel() * sizeof(float));
  launch_transpose(fptr(t), dst.as<float>(), static_cast<int>(t.shape[0]), static_cast<int>(t.shape[1]), s);
}

void pack_conv_bn(vcg_engine* e, ConvW& cw, const std::string& wkey, const std::string& bnkey, int Cin, int Cout, int k,
                  int stride, cudaStream_t s) {
  const RawTensor& w = need(e, wkey, {Cout, Cin, k, k});
  const RawTensor& g = need(e, bnkey + ".weight", {Cout});
  const RawTensor& b = need(e, bnkey + ".bias", {Cout});
  const RawTensor& m = need(e, bnkey + ".running_mean", {Cout});
  const RawTensor& v = need(e, bnkey + ".running_var", {Cout});
  cw.Cin = Cin; cw.Cout = Cout; cw.k = k; cw.stride = stride;
  cw.w.alloc(static_cast<size_t>(Cout) * Cin * k * k * e->es());
  cw.bias.alloc(Cout * sizeof(float));
  launch_pack_conv(fptr(w), fptr(g), fptr(b), fptr(m), fptr(v), 1e-5f, Cout, Cin, k, cw.w.p, cw.bias.as<float>(),
                   e->fp32, s);
}

std::string conv1_key(vcg_engine* e, const std::string& prefix) {
  // TemporalShift wraps conv1 (ops/temporal_shift.py:138), which inserts ".net" into the key
  if (e->raw.count(prefix + ".conv1.net.weight")) return prefix + ".conv1.net.weight";
  return prefix + ".conv1.weight";
}

void pack_linear(vcg_engine* e, LinearW& lw, const std::string& prefix, int N, int K, cudaStream_t s) {
  const RawTensor& w = need(e, prefix + ".weight", {N, K});
  const RawTensor& b = need(e, prefix + ".bias", {N});
  lw.N = N; lw.K = K;
  convert_to(lw.w, w, e->fp32, s);
  copy_f32(lw.bias, b, s);
}

void finalize_vision(vcg_engine* e, cudaStream_t s) {
  const std::string vm = e->vis_prefix();
  {
    const RawTensor& w = need(e, vm + "conv1.weight", {64, 3, 7, 7});
    const RawTensor& g = need(e, vm + "bn1.weight", {64});
    const RawTensor& b = need(e, vm + "bn1.bias", {64});
    const RawTensor& m = need(e, vm + "bn1.running_mean", {64});
    const RawTensor& v = need(e, vm + "bn1.running_var", {64});
    e->stem.Cin = 3; e->stem.Cout = 64; e->stem.k = 7; e->stem.stride = 2;
    e->stem.w.alloc(static_cast<size_t>(64) * (e->fp32 ? 7 * 32 : 4 * 64) * e->es());
    e->stem.bias.alloc(64 * sizeof(float));
    launch_pack_stem(fptr(w), fptr(g), fptr(b), fptr(m), fptr(v), 1e-5f, e->stem.w.p, e->stem.bias.as<float>(), e->fp32, s);
  }
  const int nblocks[4] = {3, 4, 6, 3};
  const int planes[4] = {64, 128, 256, 512};
  int inplanes = 64, Hs = 56;
  e->blocks.clear();
  e->blocks.resize(16);
  int bi = 0;
  for (int st = 0; st < 4; ++st) {
    for (int i = 0; i < nblocks[st]; ++i, ++bi) {
      Bottleneck& bk = e->blocks[bi];
      const std::string pre = vm + "layer" + std::to_string(st + 1) + "." + std::to_string(i);
      bk.stride = (i == 0 && st > 0) ? 2 : 1;
      bk.Cin = inplanes; bk.planes = planes[st]; bk.H = Hs;
      bk.has_ds = (i == 0);
      pack_conv_bn(e, bk.c1, conv1_key(e, pre), pre + ".bn1", inplanes, planes[st], 1, 1, s);
      pack_conv_bn(e, bk.c2, pre + ".conv2.weight", pre + ".bn2", planes[st], planes[st], 3, bk.stride, s);
      pack_conv_bn(e, bk.c3, pre + ".conv3.weight", pre + ".bn3", planes[st], planes[st] * 4, 1, 1, s);
      if (bk.has_ds)
        pack_conv_bn(e, bk.ds, pre + ".downsample.0.weight", pre + ".downsample.1", inplanes, planes[st] * 4, 1,
                     bk.stride, s);
      inplanes = planes[st] * 4;
      Hs /= bk.stride;
    }
  }
  if (e->tsm && !e->fp32) {   // layer1.0.conv1 as a 3-tap temporal conv for the shared-stem path
    Bottleneck& b0 = e->blocks[0];
    const std::string pre = vm + "layer1.0";
    const RawTensor& w = need(e, conv1_key(e, pre), {64, 64, 1, 1});
    b0.c1_shared.Cin = 64; b0.c1_shared.Cout = 64; b0.c1_shared.k = 1; b0.c1_shared.stride = 1;
    b0.c1_shared.w.alloc(static_cast<size_t>(64) * 3 * 64 * 2);
    b0.c1_shared.bias.alloc(64 * sizeof(float));
    launch_pack_conv1_shared(fptr(w), fptr(need(e, pre + ".bn1.weight")), fptr(need(e, pre + ".bn1.bias")),
                             fptr(need(e, pre + ".bn1.running_mean")), fptr(need(e, pre + ".bn1.running_var")), 1e-5f, 64, 64,
                             64 / e->cfg.shift_div, b0.c1_shared.w.p, b0.c1_shared.bias.as<float>(), s);
  }
  // workspace
  const size_t nF = static_cast<size_t>(e->Bv) * e->T, es = e->es();
  e->stem_in.alloc(nF * kStemHp * kStemWp * 4 * es, /*zero=*/true);   // borders stay zero forever
  e->stem_out.alloc(nF * 112 * 112 * 64 * es);
  e->x0.alloc(nF * 56 * 56 * 64 * es);
  e->xa.alloc(nF * 56 * 56 * 256 * es);
  e->xb.alloc(nF * 56 * 56 * 256 * es);
  e->dsbuf.alloc(nF * 56 * 56 * 256 * es);
  e->mid1.alloc(nF * 56 * 56 * 128 * es);
  e->mid2.alloc(nF * 56 * 56 * 64 * es);
  e->shifted.clear();
  for (int i = 0; i < 16; ++i) {
    const Bottleneck& bk = e->blocks[i];
    // two shifted channel groups of `fold = Cin / shift_div` channels each (block 0: the max-pool writes a full copy)
    const int ch = (i == 0) ? 64 : (e->tsm ? 2 * bk.Cin / e->cfg.shift_div : 0);
    auto buf = std::make_unique<DevBuf>();
    if (e->tsm) buf->alloc(nF * bk.H * bk.H * ch * es, /*zero=*/true);   // never-written edge frames stay zero
    e->shifted.push_back(std::move(buf));
  }
}

void finalize_text(vcg_engine* e, cudaStream_t s) {
  const std::string lm = e->lang_prefix();
  const RawTensor& word = need(e, lm + "embeddings.word_embeddings.weight");
  VCG_REQUIRE(word.shape.size() == 2 && word.shape[1] == kBertHidden, "word embedding must be [vocab,768]");
  const RawTensor& pos = need(e, lm + "embeddings.position_embeddings.weight");
  VCG_REQUIRE(pos.shape.size() == 2 && pos.shape[1] == kBertHidden && pos.shape[0] >= e->Lmax,
              "position embedding table shorter than max_tokens");
  const RawTensor& type = need(e, lm + "embeddings.token_type_embeddings.weight");
  e->vocab = static_cast<int>(word.shape[0]);
  convert_to(e->word, word, e->fp32, s);
  convert_to(e->pos, pos, e->fp32, s);
  convert_to(e->type, type, e->fp32, s);
  copy_f32(e->emb_g, need(e, lm + "embeddings.LayerNorm.weight", {kBertHidden}), s);
  copy_f32(e->emb_b, need(e, lm + "embeddings.LayerNorm.bias", {kBertHidden}), s);
  int n_layers = 0;
  while (e->raw.count(lm + "encoder.layer." + std::to_string(n_layers) + ".attention.self.query.weight")) ++n_layers;
  VCG_REQUIRE(n_layers > 0, "no BERT encoder layers found in the state dict");
  e->layers.clear();
  e->layers.resize(n_layers);
  for (int i = 0; i < n_layers; ++i) {
    BertLayerW& lw = e->layers[i];
    const std::string pre = lm + "encoder.layer." + std::to_string(i) + ".";
    // fused QKV: rows [0,768) query, [768,1536) key, [1536,2304) value
    lw.qkv.N = 3 * kBertHidden; lw.qkv.K = kBertHidden;
    lw.qkv.w.alloc(static_cast<size_t>(3) * kBertHidden * kBertHidden * e->es());
    lw.qkv.bias.alloc(3 * kBertHidden * sizeof(float));
    const char* names[3] = {"query", "key", "value"};
    for (int j = 0; j < 3; ++j) {
      const RawTensor& w = need(e, pre + "attention.self." + names[j] + ".weight", {kBertHidden, kBertHidden});
      const RawTensor& b = need(e, pre + "attention.self." + names[j] + ".bias", {kBertHidden});
      launch_convert(fptr(w), static_cast<uint8_t*>(lw.qkv.w.p) + static_cast<size_t>(j) * kBertHidden * kBertHidden * e->es(),
                     static_cast<long>(kBertHidden) * kBertHidden, e->fp32, s);
      VCG_CUDA(cudaMemcpyAsync(lw.qkv.bias.as<float>() + j * kBertHidden, b.buf->p, kBertHidden * sizeof(float),
                               cudaMemcpyDeviceToDevice, s));
    }
    pack_linear(e, lw.out, pre + "attention.output.dense", kBertHidden, kBertHidden, s);
    copy_f32(lw.ln1_g, need(e, pre + "attention.output.LayerNorm.weight", {kBertHidden}), s);
    copy_f32(lw.ln1_b, need(e, pre + "attention.output.LayerNorm.bias", {kBertHidden}), s);
    pack_linear(e, lw.ffn1, pre + "intermediate.dense", kBertFfn, kBertHidden, s);
    pack_linear(e, lw.ffn2, pre + "output.dense", kBertHidden, kBertFfn, s);
    copy_f32(lw.ln2_g, need(e, pre + "output.LayerNorm.weight", {kBertHidden}), s);
    copy_f32(lw.ln2_b, need(e, pre + "output.LayerNorm.bias", {kBertHidden}), s);
  }
  {
    // opt-in (VCG_LN_FUSED=1): measured on B200 the LayerNorm terms lengthen the latency-bound GEMM epilogues by more
    // (0.48 ms per 256-clip pass) than the 24 stand-alone LayerNorm launches cost (0.34 ms); see DESIGN.md
    const char* v = getenv("VCG_LN_FUSED");
    e->ln_fused = !e->fp32 && v && atoi(v) == 1;
  }
  if (e->ln_fused) {
    for (int i = 0; i < n_layers; ++i) {
      BertLayerW& lw = e->layers[i];
      const std::string pre = lm + "encoder.layer." + std::to_string(i) + ".";
      // FFN-in consumes LayerNorm1 of this layer
      lw.ffn1_f.N = kBertFfn; lw.ffn1_f.K = kBertHidden;
      lw.ffn1_f.w.alloc(static_cast<size_t>(kBertFfn) * kBertHidden * 2);
      lw.ffn1_f.bias.alloc(kBertFfn * sizeof(float));
      lw.ffn1_c1.alloc(kBertFfn * sizeof(float));
      launch_fold_ln_linear(fptr(need(e, pre + "intermediate.dense.weight")), lw.ln1_g.as<float>(), lw.ln1_b.as<float>(),
                            fptr(need(e, pre + "intermediate.dense.bias")), kBertFfn, kBertHidden, lw.ffn1_f.w.p,
                            lw.ffn1_c1.as<float>(), lw.ffn1_f.bias.as<float>(), s);
      if (i == 0) continue;
      // QKV consumes LayerNorm2 of the previous layer
      const BertLayerW& prev = e->layers[i - 1];
      lw.qkv_f.N = 3 * kBertHidden; lw.qkv_f.K = kBertHidden;
      lw.qkv_f.w.alloc(static_cast<size_t>(3) * kBertHidden * kBertHidden * 2);
      lw.qkv_f.bias.alloc(3 * kBertHidden * sizeof(float));
      lw.qkv_c1.alloc(3 * kBertHidden * sizeof(float));
      const char* names[3] = {"query", "key", "value"};
      for (int j = 0; j < 3; ++j)
        launch_fold_ln_linear(fptr(need(e, pre + "attention.self." + names[j] + ".weight")), prev.ln2_g.as<float>(),
                              prev.ln2_b.as<float>(), fptr(need(e, pre + "attention.self." + names[j] + ".bias")),
                              kBertHidden, kBertHidden,
                              static_cast<uint8_t*>(lw.qkv_f.w.p) + static_cast<size_t>(j) * kBertHidden * kBertHidden * 2,
                              lw.qkv_c1.as<float>() + j * kBertHidden, lw.qkv_f.bias.as<float>() + j * kBertHidden, s);
    }
  }
  pack_linear(e, e->pooler, lm + "pooler.dense", kBertHidden, kBertHidden, s);
  // workspace: +128 rows of slack so that a TMA box starting at the last valid row never leaves the allocation
  const size_t rows = static_cast<size_t>(e->Bt) * e->Lmax + 128, es = e->es();
  e->hid.alloc(rows * kBertHidden * es);
  e->hid2.alloc(rows * kBertHidden * es);
  e->qkv.alloc(rows * 3 * kBertHidden * es, /*zero=*/true);   // rows past the packed tokens are read (and multiplied by 0)
  e->ctx.alloc(rows * kBertHidden * es);
  e->tmp.alloc(rows * kBertHidden * es);
  e->ffn.alloc(rows * kBertFfn * es);
  e->cls.alloc((static_cast<size_t>(e->Bt) + 128) * kBertHidden * es);
  e->pk_cu.alloc((e->Bt + 1) * sizeof(int32_t));
  e->pk_src.alloc(rows * sizeof(int32_t));
  e->pk_ok.alloc(rows);
  e->pk_total.alloc(sizeof(int32_t), /*zero=*/true);
  e->pk_items.alloc(static_cast<size_t>(e->Bt) * kBertHeads * 8);
  e->pk_nitems.alloc(sizeof(int32_t), /*zero=*/true);
  if (e->ln_fused) {
    e->ln_stats_rows = rows;
    e->ln_stats.alloc(static_cast<size_t>(2) * n_layers * rows * sizeof(float2), /*zero=*/true);
  }
}

void finalize_head(vcg_engine* e, cudaStream_t s) {
  const std::string fh = "fusion_head.";
  const int H = e->H, T = e->T;
  if (e->cfg.modality == VCG_MODALITY_EMBED) {        // backbones only
    const size_t max_frames = static_cast<size_t>(e->Bv) * T;
    if (e->cfg.vision == VCG_VISION_R50TSM) {
      e->vis_emb.alloc(max_frames * kVisionDim * sizeof(float));
      if (!e->fp32) e->vis_emb_act.alloc(max_frames * kVisionDim * e->es());
    }
    e->pooled.alloc((static_cast<size_t>(e->Bt) + 128) * kBertHidden * e->es());
    return;
  }
  if (e->cfg.modality != VCG_MODALITY_TWO_STREAM) {   // nn.Linear(D, 2) on the backbone embedding
    const int64_t D = e->cfg.modality == VCG_MODALITY_VISION ? static_cast<int64_t>(T) * kVisionDim : kBertHidden;
    copy_f32(e->head_w, need(e, "head.weight", {2, D}), s);
    copy_f32(e->head_b, need(e, "head.bias", {2}), s);
    if (e->cfg.modality == VCG_MODALITY_VISION) {
      const size_t max_frames = static_cast<size_t>(e->Bv) * T;
      e->vis_emb.alloc(max_frames * kVisionDim * sizeof(float));
      if (!e->fp32) e->vis_emb_act.alloc(max_frames * kVisionDim * e->es());
    } else {
      e->pooled.alloc((static_cast<size_t>(e->Bt) + 128) * kBertHidden * e->es());
      e->pooled32.alloc(static_cast<size_t>(e->Bt) * kBertHidden * sizeof(float));
    }
    return;
  }
  convert_to(e->lang_w, need(e, fh + "lang_proj_head.weight", {H, kBertHidden}), e->fp32, s);
  // the vision projection runs as a 3xTF32 GEMM on the fp32 average-pooled embeddings in BOTH precisions and hands fp32
  // tokens to the fp32 head: rounding the 2048-d embedding, W_v and the 128-d tokens to bf16 cost a third of the bf16
  // mode's whole margin error (2.8e-3 -> 1.8e-3 on the 146-clip golden video) for 0.3 % of the step time
  convert_to(e->vis_w, need(e, fh + "vision_proj_head.weight", {H, kVisionDim}), true, s);
  if (e->cfg.head_type == VCG_HEAD_MLP) {
    copy_f32(e->head_w, need(e, fh + "head.weight", {2, static_cast<int64_t>(T + 1) * H}), s);
    copy_f32(e->head_b, need(e, fh + "head.bias", {2}), s);
  } else {
    transpose_to(e->q_w_t, need(e, fh + "head.query.weight", {H, H}), s);
    copy_f32(e->q_b, need(e, fh + "head.query.bias", {H}), s);
    transpose_to(e->k_w_t, need(e, fh + "head.key.weight", {H, H}), s);
    copy_f32(e->k_b, need(e, fh + "head.key.bias", {H}), s);
    transpose_to(e->v_w_t, need(e, fh + "head.value.weight", {H, H}), s);
    copy_f32(e->v_b, need(e, fh + "head.value.bias", {H}), s);
    copy_f32(e->proj_w, need(e, fh + "head.proj.weight", {2, H}), s);
    copy_f32(e->proj_b, need(e, fh + "head.proj.bias", {2}), s);
  }
  const size_t max_frames = static_cast<size_t>(std::max(e->Bv, e->Bt)) * T;
  e->vis_emb.alloc(max_frames * kVisionDim * sizeof(float));
  e->vis_out.alloc((max_frames + 128) * H * sizeof(float));
  e->lang_out.alloc((static_cast<size_t>(e->Bt) + 128) * H * e->es());
  e->pooled.alloc((static_cast<size_t>(e->Bt) + 128) * kBertHidden * e->es());
}

// ------------------------------------------------------------------------------------------------ plans
// Can the stem / max-pool / layer1.0 downsample be computed once per unique frame for clips start0 + b*stride ?
bool shared_stem_ok(const vcg_engine* e, int stride) {
  return e->tsm && !e->fp32 && stride >= 1 && stride < e->T && e->T % 2 == 0;
}

// clip_stride == 0: every clip brings its own T frames (N = B*T images everywhere).
// clip_stride  > 0: clips overlap (clip b = unique frames b*stride .. +T-1): stem, max-pool and layer1.0's downsample run
//                   once per unique frame; layer1.0.conv1 and the residual of layer1.0.conv3 read them through clip views.
VisionPlan& vision_plan(vcg_engine* e, int B, int clip_stride = 0) {
  const bool shared = clip_stride > 0;
  if (shared) {
    auto it = e->vplans_shared.find({B, clip_stride});
    if (it != e->vplans_shared.end()) return it->second;
  } else {
    auto it = e->vplans.find(B);
    if (it != e->vplans.end()) return it->second;
  }
  VisionPlan plan;
  plan.B = B;
  const int N = B * e->T;
  const int U = shared ? clip_stride * (B - 1) + e->T : N;   // images through the stem
  const bool fp = e->fp32;
  {
    Step st{};
    st.kind = Step::CONV_GEMM;
    Epilogue ep;
    ep.bias = e->stem.bias.as<float>();
    ep.act = ACT_RELU;
    st.gemm = build_stem(e->stem_in.p, U, kStemHp, kStemWp, kStemOut, kStemOut, e->stem.w.p, 64, e->stem_out.p, fp, ep);
    plan.steps.push_back(st);
  }
  {
    Step st{};
    st.kind = Step::MAXPOOL;
    st.in = e->stem_out.p; st.out = e->x0.p; st.out2 = (e->tsm && !shared) ? e->shifted[0]->p : nullptr;
    st.n = U; st.a = e->T; st.c = e->tsm ? 64 / e->cfg.shift_div : 0;
    plan.steps.push_back(st);
  }
  const void* x = e->x0.p;
  void* pingpong[2] = {e->xa.p, e->xb.p};
  int pp = 0;
  // blocks whose conv1 reads the temporally shifted channels straight from frames t+1 / t-1 of its input (Cin >= 512);
  // the others (layer1, layer2.0) get a shifted copy scattered by the producing epilogue.  VCG_TSM_DIRECT=0: never.
  static const bool direct_on = !(getenv("VCG_TSM_DIRECT") && atoi(getenv("VCG_TSM_DIRECT")) == 0);
  auto tsm_direct = [&](int blk) {
    const Bottleneck& b = e->blocks[blk];
    return direct_on && e->tsm && blk > 0 &&
           conv1_tsm_direct_ok(b.H, b.H, N, b.Cin, b.Cin / e->cfg.shift_div, e->T, fp);
  };
  static const char* const kNames[4][4] = {{"l1.conv1", "l1.conv2", "l1.conv3", "l1.downsample"},
                                           {"l2.conv1", "l2.conv2", "l2.conv3", "l2.downsample"},
                                           {"l3.conv1", "l3.conv2", "l3.conv3", "l3.downsample"},
                                           {"l4.conv1", "l4.conv2", "l4.conv3", "l4.downsample"}};
  for (int i = 0; i < 16; ++i) {
    const Bottleneck& bk = e->blocks[i];
    const int stage = i < 3 ? 0 : i < 7 ? 1 : i < 13 ? 2 : 3;
    const int H = bk.H, Ho = H / bk.stride, Cout = bk.planes * 4;
    void* xnext = pingpong[pp];
    pp ^= 1;
    {   // conv1 1x1 (+ temporal shift folded into the operand load) + BN + ReLU
      Step st{}; st.kind = Step::CONV_GEMM;
      Epilogue ep; ep.bias = bk.c1.bias.as<float>(); ep.act = ACT_RELU;
      const void* sh = e->tsm ? e->shifted[i]->p : nullptr;
      const int sh_ch = e->tsm ? ((i == 0) ? 64 : 2 * (bk.Cin / e->cfg.shift_div)) : 0;
      if (shared && i == 0) {
        ep.bias = bk.c1_shared.bias.as<float>();
        st.gemm = build_conv1_shared(x, B, e->T, clip_stride, H, H, bk.Cin, bk.c1_shared.w.p, bk.planes, e->mid1.p, ep,
                                     kNames[stage][0]);
      } else if (tsm_direct(i)) {   // shifted channel groups straight from frames t+1 / t-1 of the block input
        st.gemm = build_conv1_tsm_direct(x, N, e->T, H, H, bk.Cin, bk.Cin / e->cfg.shift_div, bk.c1.w.p, bk.planes, e->mid1.p, ep,
                                         kNames[stage][0]);
      } else {
        st.gemm = build_conv(x, N, H, H, bk.Cin, bk.c1.w.p, bk.planes, 1, 1, e->mid1.p, fp, ep, sh, sh_ch, kNames[stage][0]);
      }
      plan.steps.push_back(st);
    }
    const bool fuse23 = conv23_ok(bk.planes, fp);   // layer1 / layer2: conv2 + conv3 in one kernel (conv23.cuh)
    if (!fuse23) {   // conv2 3x3 (stride) + BN + ReLU
      Step st{}; st.kind = Step::CONV_GEMM;
      Epilogue ep; ep.bias = bk.c2.bias.as<float>(); ep.act = ACT_RELU;
      st.gemm = build_conv(e->mid1.p, N, H, H, bk.planes, bk.c2.w.p, bk.planes, 3, bk.stride, e->mid2.p, fp, ep, nullptr, 0, kNames[stage][1]);
      plan.steps.push_back(st);
    }
    const void* identity = x;
    if (bk.has_ds) {   // downsample 1x1 (stride) + BN on the un-shifted block input
      Step st{}; st.kind = Step::CONV_GEMM;
      Epilogue ep; ep.bias = bk.ds.bias.as<float>(); ep.act = ACT_NONE;
      st.gemm = build_conv(x, (shared && i == 0) ? U : N, H, H, bk.Cin, bk.ds.w.p, Cout, 1, bk.stride, e->dsbuf.p, fp, ep, nullptr,
                           0, kNames[stage][3]);
      plan.steps.push_back(st);
      identity = e->dsbuf.p;
    }
    {   // conv3 1x1 + BN + residual + ReLU, scattering the next block's shifted channels
      Step st{}; st.kind = Step::CONV_GEMM;
      Epilogue ep; ep.bias = bk.c3.bias.as<float>(); ep.act = ACT_RELU;
      ep.residual = identity; ep.ld_res = Cout;
      if (shared && i == 0) { ep.res_clip_T = e->T; ep.res_clip_stride = clip_stride; }
      if (e->tsm && i + 1 < 16 && !tsm_direct(i + 1)) {
        ep.tsm_out = e->shifted[i + 1]->p;
        ep.tsm_fold = Cout / e->cfg.shift_div;
        ep.tsm_ld = 2 * ep.tsm_fold;
        ep.T = e->T;
      }
      if (fuse23 && conv23h_ok(bk.planes, bk.stride, H, H, fp)) {
        st.kind = Step::CONV23;
        st.c23 = build_conv23h(e->mid1.p, N, H, H, bk.c2.w.p, bk.c2.bias.as<float>(), bk.c3.w.p, xnext, ep, kNames[stage][2]);
      } else if (fuse23 && conv23h2_ok(bk.planes, bk.stride, H, H, fp, ep.tsm_out)) {
        st.kind = Step::CONV23;
        st.c23 = build_conv23h(e->mid1.p, N, H, H, bk.c2.w.p, bk.c2.bias.as<float>(), bk.c3.w.p, xnext, ep, kNames[stage][2], 128);
      } else if (fuse23) {
        st.kind = Step::CONV23;
        st.c23 = build_conv23(e->mid1.p, N, H, H, bk.planes, bk.stride, bk.c2.w.p, bk.c2.bias.as<float>(), bk.c3.w.p, xnext, ep,
                              kNames[stage][2]);
      } else {
        st.gemm = build_conv(e->mid2.p, N, Ho, Ho, bk.planes, bk.c3.w.p, Cout, 1, 1, xnext, fp, ep, nullptr, 0, kNames[stage][2]);
      }
      plan.steps.push_back(st);
    }
    x = xnext;
  }
  plan.final_act = x;
  if (shared) return e->vplans_shared.emplace(std::make_pair(B, clip_stride), std::move(plan)).first->second;
  return e->vplans.emplace(B, std::move(plan)).first->second;
}

BertPlan& bert_plan(vcg_engine* e, int B, int L) {
  auto key = std::make_pair(B, L);
  auto it = e->bplans.find(key);
  if (it != e->bplans.end()) return it->second;
  BertPlan plan;
  plan.B = B; plan.L = L;
  const int M = B * L;
  const bool fp = e->fp32;
  auto gemm_step = [&](const void* A, const LinearW& w, void* out, int act, const Epilogue& ep0, const char* name) {
    Step st{}; st.kind = Step::CONV_GEMM;
    Epilogue ep = ep0; ep.bias = w.bias.as<float>(); ep.act = act; ep.ld_res = w.N;
    st.gemm = build_gemm(A, w.K, w.w.p, out, w.N, M, w.N, w.K, fp, ep, name);
    st.gemm.p.m_dev = e->pk_total.as<int32_t>();   // only the packed rows are computed
    plan.steps.push_back(st);
  };
  auto res_ep = [](const void* res) { Epilogue ep; ep.residual = res; return ep; };
  if (!e->ln_fused) {
    auto ln_step = [&](const void* x, const DevBuf& g, const DevBuf& b, void* y) {
      Step st{}; st.kind = Step::LAYERNORM;
      st.in = x; st.out = y; st.g = g.as<float>(); st.b = b.as<float>(); st.n = M;
      plan.steps.push_back(st);
    };
    for (size_t i = 0; i < e->layers.size(); ++i) {
      const BertLayerW& lw = e->layers[i];
      gemm_step(e->hid.p, lw.qkv, e->qkv.p, ACT_NONE, Epilogue{}, "bert.qkv");
      { Step st{}; st.kind = Step::ATTENTION; st.in = e->qkv.p; st.out = e->ctx.p; st.n = B; st.a = L; plan.steps.push_back(st); }
      gemm_step(e->ctx.p, lw.out, e->tmp.p, ACT_NONE, res_ep(e->hid.p), "bert.attn_out");
      ln_step(e->tmp.p, lw.ln1_g, lw.ln1_b, e->hid2.p);
      gemm_step(e->hid2.p, lw.ffn1, e->ffn.p, ACT_GELU, Epilogue{}, "bert.ffn_in");
      gemm_step(e->ffn.p, lw.ffn2, e->tmp.p, ACT_NONE, res_ep(e->hid2.p), "bert.ffn_out");
      ln_step(e->tmp.p, lw.ln2_g, lw.ln2_b, e->hid.p);
    }
    plan.final_hidden = e->hid.p;
  } else {
    // LayerNorm folded into the GEMMs: t1 = attn_out + x (raw, in tmp) and t2 = ffn_out + LN1(t1) (raw, in hid2) are
    // stored un-normalised together with their row statistics (accumulated by the producing epilogue); the consumers
    // apply the LayerNorm algebraically (A side) or recompute it on the residual tile.  No LayerNorm kernel runs.
    { Step st{}; st.kind = Step::ZERO_STATS; plan.steps.push_back(st); }
    auto stats = [&](size_t layer, int which) {
      return e->ln_stats.as<float2>() + (2 * layer + which) * e->ln_stats_rows;
    };
    for (size_t i = 0; i < e->layers.size(); ++i) {
      const BertLayerW& lw = e->layers[i];
      if (i == 0) {
        gemm_step(e->hid.p, lw.qkv, e->qkv.p, ACT_NONE, Epilogue{}, "bert.qkv");
      } else {
        Epilogue ep; ep.a_stats = stats(i - 1, 1); ep.ln_c1 = lw.qkv_c1.as<float>();
        gemm_step(e->hid2.p, lw.qkv_f, e->qkv.p, ACT_NONE, ep, "bert.qkv");
      }
      { Step st{}; st.kind = Step::ATTENTION; st.in = e->qkv.p; st.out = e->ctx.p; st.n = B; st.a = L; plan.steps.push_back(st); }
      {
        Epilogue ep;
        if (i == 0) {
          ep.residual = e->hid.p;
        } else {
          const BertLayerW& prev = e->layers[i - 1];
          ep.residual = e->hid2.p; ep.res_stats = stats(i - 1, 1);
          ep.res_gamma = prev.ln2_g.as<float>(); ep.res_beta = prev.ln2_b.as<float>();
        }
        ep.out_stats = stats(i, 0);
        gemm_step(e->ctx.p, lw.out, e->tmp.p, ACT_NONE, ep, "bert.attn_out");
      }
      {
        Epilogue ep; ep.a_stats = stats(i, 0); ep.ln_c1 = lw.ffn1_c1.as<float>();
        gemm_step(e->tmp.p, lw.ffn1_f, e->ffn.p, ACT_GELU, ep, "bert.ffn_in");
      }
      {
        Epilogue ep;
        ep.residual = e->tmp.p; ep.res_stats = stats(i, 0);
        ep.res_gamma = lw.ln1_g.as<float>(); ep.res_beta = lw.ln1_b.as<float>();
        ep.out_stats = stats(i, 1);
        gemm_step(e->ffn.p, lw.ffn2, e->hid2.p, ACT_NONE, ep, "bert.ffn_out");
      }
    }
    plan.final_hidden = e->hid2.p;
    plan.final_is_raw = true;
  }
  return e->bplans.emplace(key, std::move(plan)).first->second;
}

// Brackets one kernel launch with CUDA events on the launching stream when profiling is on.
struct ProfScope {
  vcg_engine* e;
  cudaStream_t s;
  ProfRec rec;
  bool on;
  ProfScope(vcg_engine* e_, cudaStream_t s_, std::string name, double flops, double bytes, bool packed = false)
      : e(e_), s(s_), on(e_->profiling) {
    ++e->launches;
    if (!on) return;
    rec.name = std::move(name); rec.flops = flops; rec.bytes = bytes; rec.packed = packed;
    rec.start = e->next_event(); rec.stop = e->next_event();
    VCG_CUDA(cudaEventRecord(rec.start, s));
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(rec.stop, s);
    e->prof.push_back(std::move(rec));
  }
};

std::string gemm_kernel_name(const ConvGemmLaunch& L) {
  return std::string("conv_gemm_") + (L.fp32 ? "tf32x3" : "bf16") + "_n" + std::to_string(L.block_n) + "|" + L.name;
}

void run_steps(vcg_engine* e, const std::vector<Step>& steps, const int64_t* mask, cudaStream_t s) {
  const double es = e->es();
  auto checksum = [&](const void* ptr, double bytes) {
    if (!e->debug_checksum || e->dbg_n >= 250) return;
    launch_checksum(ptr, static_cast<size_t>(bytes), e->dbg.as<unsigned long long>() + e->dbg_n++, s);
  };
  for (const Step& st : steps) {
    if (e->debug_checksum && mask == nullptr) {   // vision plans only
      if (st.kind == Step::CONV_GEMM) {
        const ConvGemmParams& p = st.gemm.p;
        launch_conv_gemm(st.gemm, s);
        ++e->launches;
        checksum(p.out, static_cast<double>(p.Nimg) * p.Ho * p.Wo * p.ld_out * es);
        if (p.tsm_out) checksum(p.tsm_out, static_cast<double>(p.Nimg) * p.Ho * p.Wo * p.tsm_ld * es);
        continue;
      }
      if (st.kind == Step::CONV23) {
        const ConvGemmParams& p = st.c23.q.g;
        launch_conv23(st.c23, s);
        ++e->launches;
        checksum(p.out, static_cast<double>(p.Nimg) * p.Ho * p.Wo * p.ld_out * es);
        if (p.tsm_out) checksum(p.tsm_out, static_cast<double>(p.Nimg) * p.Ho * p.Wo * p.tsm_ld * es);
        continue;
      }
      if (st.kind == Step::MAXPOOL) {
        launch_maxpool_tsm(st.in, st.n, st.out, st.out2, st.a, st.c, e->fp32, s);
        ++e->launches;
        checksum(st.out, st.n * 56.0 * 56 * 64 * es);
        if (st.out2) checksum(st.out2, st.n * 56.0 * 56 * 64 * es);
        continue;
      }
    }
    switch (st.kind) {
      case Step::CONV_GEMM: {
        ProfScope ps(e, s, gemm_kernel_name(st.gemm), st.gemm.flops, st.gemm.bytes, st.gemm.p.m_dev != nullptr);
        launch_conv_gemm(st.gemm, s);
        break;
      }
      case Step::CONV23: {
        ProfScope ps(e, s, std::string(st.c23.halo == 2 ? "conv23h2_bf16|" : st.c23.halo ? "conv23h_bf16|" : "conv23_bf16|") + st.c23.name, st.c23.flops, st.c23.bytes);
        launch_conv23(st.c23, s);
        break;
      }
      case Step::MAXPOOL: {
        ProfScope ps(e, s, "maxpool_tsm|maxpool", 0, st.n * (112.0 * 112 * 64 + (st.out2 ? 2 : 1) * 56.0 * 56 * 64) * es);
        launch_maxpool_tsm(st.in, st.n, st.out, st.out2, st.a, st.c, e->fp32, s);
        break;
      }
      case Step::LAYERNORM: {
        ProfScope ps(e, s, "layernorm768|bert.ln", 0, st.n * 768.0 * 2 * es, true);
        launch_layernorm(st.in, st.g, st.b, st.out, st.n, kBertHidden, 1e-12f, e->fp32, s, e->pk_total.as<int32_t>());
        break;
      }
      case Step::ATTENTION: {
        ProfScope ps(e, s, "bert_attention|bert.attn", 4.0 * st.n * kBertHeads * static_cast<double>(st.a) * st.a * 64, 0, true);
        launch_bert_attention(st.in, mask, e->pk_cu.as<int32_t>(), e->pk_ok.as<uint8_t>(), st.out, st.n, st.a, e->fp32, s,
                              static_cast<long>(e->Bt) * e->Lmax + 128, e->pk_items.p, e->pk_nitems.as<int32_t>());
        break;
      }
      case Step::ZERO_STATS: {
        ProfScope ps(e, s, "memset|bert.ln_stats", 0, static_cast<double>(e->ln_stats.bytes));
        VCG_CUDA(cudaMemsetAsync(e->ln_stats.p, 0, e->ln_stats.bytes, s));
        break;
      }
      default: throw Error("vcg: unknown plan step");
    }
  }
}

struct FrameSource {
  const float* img_clip = nullptr;     // [B,T,3,224,224] fp32
  const uint8_t* frames_u8 = nullptr;  // [n_frames,224,224,3]
  const int32_t* clip_start = nullptr; // [B] (device)
  const int32_t* clip_start_host = nullptr;   // the same on the host when the caller has it (pass planning, bounds checks)
  int n_frames = 0;                    // frames in frames_u8 (0 = unknown): device-side indices are clamped to it
};

// One vision pass over clips [g0, g0 + n).  stride > 0: the clips sit on a regular grid (clip g0 + i starts at frame
// f0 + i*stride, stride < T), so the stem / max-pool / layer1.0 downsample run once per DISTINCT frame.
struct VisionPass {
  int n = 0, stride = 0;
  long f0 = 0;
};

// Progress of an asynchronous host->device copy of the vision input (side stream): `ready[i]` = (units on the device
// once event i has fired, event); units are frames when clip_start_host is given, clips otherwise.
struct HostFeed {
  const int32_t* clip_start_host = nullptr;
  int T = 0;
  std::vector<std::pair<long, cudaEvent_t>> ready;
  // make `s` wait until the vision input of clips [g0, g1) is on the device
  void wait_for(cudaStream_t s, int g0, int g1) const {
    long need = g1;
    if (clip_start_host) {
      need = 0;
      for (int g = g0; g < g1; ++g) need = std::max<long>(need, static_cast<long>(clip_start_host[g]) + T);
    }
    for (const auto& r : ready)
      if (r.first >= need) {
        VCG_CUDA(cudaStreamWaitEvent(s, r.second, 0));
        return;
      }
    if (!ready.empty()) VCG_CUDA(cudaStreamWaitEvent(s, ready.back().second, 0));
  }
};

// Scores B clips; any of the sources may be used for the vision stream (or precomputed embeddings).
// Text stream for clips [b0, b0 + bt): token packing, embeddings, the encoder, BertPooler (tanh) into e->pooled and,
// for the two-stream model, the language projection (ReLU) into e->lang_out.
void run_text(vcg_engine* e, const int64_t* ids, const int64_t* mask, int b0, int bt, int L, float* lang_emb_out,
              cudaStream_t s) {
  {
    ProfScope ps(e, s, "bert_pack|bert.pack", 0, static_cast<double>(bt) * L * 13);
    launch_bert_pack(mask + static_cast<long>(b0) * L, bt, L, e->pk_cu.as<int32_t>(), e->pk_src.as<int32_t>(),
                     e->pk_ok.as<uint8_t>(), e->pk_total.as<int32_t>(), s);
  }
  e->last_bert_rows = bt * L;
  if (!e->fp32 && L <= 128) {
    ProfScope ps(e, s, "attn_items|bert.pack", 0, static_cast<double>(bt) * 12 * 8);
    launch_attention_items(e->pk_cu.as<int32_t>(), bt, e->pk_items.p, e->pk_nitems.as<int32_t>(), s);
  }
  {
    ProfScope ps(e, s, "bert_embed_ln|bert.embed", 0, static_cast<double>(bt) * L * 768 * 4 * e->es(), true);
    launch_bert_embed_ln(ids + static_cast<long>(b0) * L, bt * L, L, e->pk_src.as<int32_t>(), e->pk_total.as<int32_t>(),
                         e->vocab, e->word.p, e->pos.p, e->type.p, e->emb_g.as<float>(), e->emb_b.as<float>(), e->hid.p, e->fp32, s);
  }
  BertPlan& bp = bert_plan(e, bt, L);
  run_steps(e, bp.steps, mask + static_cast<long>(b0) * L, s);
  // ---- BertPooler (tanh) + lang projection (ReLU) for the bt clips: two small tcgen05 GEMMs over the [CLS] rows
  const bool with_proj = e->cfg.modality == VCG_MODALITY_TWO_STREAM;
  auto key = std::make_pair(bt, L);
  auto it = e->lang_tail_plans.find(key);
  if (it == e->lang_tail_plans.end()) {
    std::vector<ConvGemmLaunch> v;
    Epilogue ep1; ep1.bias = e->pooler.bias.as<float>(); ep1.act = ACT_TANH;
    v.push_back(build_gemm(e->cls.p, kBertHidden, e->pooler.w.p, e->pooled.p, kBertHidden, bt, kBertHidden, kBertHidden,
                           e->fp32, ep1, "head.pooler"));
    if (with_proj) {
      Epilogue ep2; ep2.act = ACT_RELU;
      v.push_back(build_gemm(e->pooled.p, kBertHidden, e->lang_w.p, e->lang_out.p, e->H, bt, e->H, kBertHidden, e->fp32, ep2,
                             "head.lang_proj"));
    }
    it = e->lang_tail_plans.emplace(key, std::move(v)).first;
  }
  {
    ProfScope ps(e, s, "gather_rows768|head.cls", 0, static_cast<double>(bt) * kBertHidden * 2 * e->es());
    if (bp.final_is_raw) {   // the last LayerNorm, for the pooled rows only
      const BertLayerW& last = e->layers.back();
      launch_gather_ln_rows768(bp.final_hidden, e->pk_cu.as<int32_t>(), L, bt, last.ln2_g.as<float>(),
                               last.ln2_b.as<float>(), 1e-12f, e->cls.p, s);
    } else {
      launch_gather_rows768(bp.final_hidden, e->pk_cu.as<int32_t>(), L, bt, e->cls.p, e->fp32, s);
    }
  }
  for (const ConvGemmLaunch& g : it->second) {
    ProfScope ps(e, s, gemm_kernel_name(g), g.flops, g.bytes);
    launch_conv_gemm(g, s);
  }
  if (lang_emb_out) {
    ProfScope ps(e, s, "cast_to_f32|head.lang_emb", 0, static_cast<double>(bt) * kBertHidden * (4 + e->es()));
    launch_cast_to_f32(e->pooled.p, lang_emb_out, static_cast<long>(bt) * kBertHidden, e->fp32, s);
  }
}

// Plans the next vision pass over clips [g0, g0 + remaining).  With the clip starts known on the host, every maximal
// constant-stride run of >= 8 clips (or the whole rest) becomes a shared-stem pass, and clips that are on no such run
// (the reference dataset's +1/+3 frame-file offsets at the ends of a video, data/infer_youtube_video_dataset.py:190-193,
// or arbitrary flat clips) go through the per-clip gather, cut where the next long run starts.
VisionPass next_vision_pass(const vcg_engine* e, const FrameSource& src, int g0, int remaining) {
  VisionPass vp;
  vp.n = std::min(e->Bv, remaining);
  const int32_t* h = src.clip_start_host;
  if (!src.frames_u8 || !h || remaining < 2) return vp;
  const int end = g0 + remaining;
  // length of the constant-stride run starting at clip i (not capped at Bv: a long run is cut into EQUAL passes, so a
  // 146-clip video with Bv = 64 runs as 49 + 49 + 48 clips instead of 64 + 64 + 18)
  auto run_at = [&](int i, int* stride) {
    if (i + 1 >= end) return 1;
    const int d = h[i + 1] - h[i];
    if (!shared_stem_ok(e, d)) return 1;
    int n = 2;
    while (i + n < end && h[i + n] - h[i + n - 1] == d) ++n;
    *stride = d;
    return n;
  };
  int d = 0;
  const int r = run_at(g0, &d);
  if (r >= 2 && r >= std::min(8, remaining)) {
    const int passes = (r + e->Bv - 1) / e->Bv;
    vp.n = (r + passes - 1) / passes; vp.stride = d; vp.f0 = h[g0];
    return vp;
  }
  int n = 1, d2 = 0;
  while (n < vp.n && run_at(g0 + n, &d2) < 8) ++n;
  vp.n = n;
  return vp;
}

void check_clip_starts(const int32_t* h, int B, int T, int n_frames) {
  for (int b = 0; b < B; ++b)
    if (h[b] < 0 || static_cast<long>(h[b]) + T > n_frames)
      throw Error("vcg: clip " + std::to_string(b) + " starts at frame " + std::to_string(h[b]) + ": frames [start, start+" +
                  std::to_string(T) + ") leave the buffer of " + std::to_string(n_frames) + " frames");
}

// Vision stream for clips [g0, g0 + bv) of the caller's numbering: pre-processing, ResNet-50(-TSM), average pool.
// Returns the fp32 embeddings [bv*T, 2048] (in the caller's buffer when vision_emb_out is given) and leaves the same
// values in the activation type at the fixed GEMM-operand address (e->vis_emb / e->vis_emb_act).
const float* run_vision(vcg_engine* e, const FrameSource& src, int g0, const VisionPass& vpass, float* vision_emb_out,
                        cudaStream_t s) {
  const int T = e->T, bv = vpass.n;
  if (src.img_clip) {
    ProfScope ps(e, s, "nchw_to_stem|preprocess", 0, static_cast<double>(bv) * T * kImg * kImg * 3 * (4 + e->es()));
    launch_nchw_to_stem(src.img_clip + static_cast<long>(g0) * T * 3 * kImg * kImg, bv * T, e->stem_in.p, e->fp32, s);
  } else if (e->resizing()) {
    // frames of another size: bilinear resize fused into the pre-processing (resize.cu)
    const bool shared = vpass.stride > 0;
    const int n_img = shared ? vpass.stride * (bv - 1) + T : bv * T;
    ProfScope ps(e, s, "resize_preprocess_u8|preprocess", 0,
                 static_cast<double>(n_img) * (static_cast<double>(e->frame_bytes()) + kImg * kImg * 3.0 * e->es()));
    launch_resize_preprocess_u8(src.frames_u8 + (shared ? vpass.f0 * static_cast<long>(e->frame_bytes()) : 0), nullptr,
                                shared ? nullptr : src.clip_start + g0, T, n_img, shared ? 0 : src.n_frames, e->src_h, e->src_w,
                                e->rs_xb.as<int32_t>(), e->rs_xk.as<int32_t>(), e->rs_yb.as<int32_t>(), e->rs_yk.as<int32_t>(),
                                e->stem_in.p, nullptr, e->fp32, s);
  } else if (vpass.stride > 0) {
    // overlapping clips: every unique frame of this pass is pre-processed (and run through the stem) once
    const int U = vpass.stride * (bv - 1) + T;
    ProfScope ps(e, s, "preprocess_u8|preprocess", 0, static_cast<double>(U) * kImg * kImg * 3 * (1 + e->es()));
    launch_preprocess_u8(src.frames_u8 + vpass.f0 * kImg * kImg * 3, nullptr, U, e->stem_in.p, e->fp32, s);
  } else {
    ProfScope ps(e, s, "preprocess_u8|preprocess", 0, static_cast<double>(bv) * T * kImg * kImg * 3 * (1 + e->es()));
    launch_preprocess_u8_clips(src.frames_u8, src.clip_start + g0, bv, T, e->stem_in.p, e->fp32, s, src.n_frames);
  }
  VisionPlan& vp = vision_plan(e, bv, src.frames_u8 ? vpass.stride : 0);
  if (e->debug_checksum) {
    e->dbg_n = 0;
    launch_checksum(e->stem_in.p, e->stem_in.bytes, e->dbg.as<unsigned long long>() + e->dbg_n++, s);
  }
  run_steps(e, vp.steps, nullptr, s);
  // fp32 embeddings go to the caller's buffer when asked for; in fp32 mode they are also the GEMM operand
  float* dst = vision_emb_out ? vision_emb_out + static_cast<long>(g0) * T * kVisionDim : e->vis_emb.as<float>();
  {
    ProfScope ps(e, s, "avgpool|avgpool", 0, static_cast<double>(bv) * T * kVisionDim * (49 * e->es() + 4));
    void* act_copy = (e->fp32 || e->cfg.modality == VCG_MODALITY_TWO_STREAM) ? nullptr : e->vis_emb_act.p;
    launch_avgpool(vp.final_act, bv * T, 49, kVisionDim, dst, act_copy, s, e->fp32);
  }
  if (vision_emb_out && (e->fp32 || e->cfg.modality == VCG_MODALITY_TWO_STREAM))   // keep the GEMM operand at a fixed address (cached tensor maps)
    VCG_CUDA(cudaMemcpyAsync(e->vis_emb.p, dst, static_cast<size_t>(bv) * T * kVisionDim * sizeof(float),
                             cudaMemcpyDeviceToDevice, s));
  return dst;
}

// Scores B clips; any of the sources may be used for the vision stream (or precomputed embeddings).
void score(vcg_engine* e, const FrameSource& src, const float* vision_emb_in, const int64_t* ids, const int64_t* mask,
           int B, int L, float* logits, float* probs, float* vision_emb_out, float* lang_emb_out, cudaStream_t s,
           const HostFeed* feed = nullptr) {
  VCG_REQUIRE(e->finalized, "vcg_finalize has not been called");
  VCG_REQUIRE(e->cfg.modality == VCG_MODALITY_TWO_STREAM, "engine was created for a single modality");
  VCG_REQUIRE(B >= 0 && L >= 1 && L <= e->Lmax, "token count exceeds max_tokens of the engine");
  const bool have_frames = src.img_clip || src.frames_u8;
  VCG_REQUIRE(have_frames || vision_emb_in, "no vision input given");
  if (have_frames) VCG_REQUIRE(e->cfg.vision == VCG_VISION_R50TSM, "engine was created without a vision backbone");
  const int T = e->T;
  for (int b0 = 0; b0 < B; b0 += e->Bt) {
    const int bt = std::min(e->Bt, B - b0);
    run_text(e, ids, mask, b0, bt, L, lang_emb_out ? lang_emb_out + static_cast<long>(b0) * kBertHidden : nullptr, s);
    TailParams tp{};
    tp.T = T; tp.H = e->H; tp.head_type = e->cfg.head_type;
    tp.head_w = e->head_w.as<float>(); tp.head_b = e->head_b.as<float>();
    tp.q_w_t = e->q_w_t.as<float>(); tp.q_b = e->q_b.as<float>();
    tp.k_w_t = e->k_w_t.as<float>(); tp.k_b = e->k_b.as<float>();
    tp.v_w_t = e->v_w_t.as<float>(); tp.v_b = e->v_b.as<float>();
    tp.proj_w = e->proj_w.as<float>(); tp.proj_b = e->proj_b.as<float>();
    // ---- vision stream + head in sub-chunks
    for (int c0 = 0, bv = 0; c0 < bt; c0 += bv) {
      const int g0 = b0 + c0;   // first clip of this sub-chunk in the caller's numbering
      VisionPass vpass;
      vpass.n = bt - c0;
      if (have_frames) vpass = next_vision_pass(e, src, g0, bt - c0);
      bv = vpass.n;
      if (feed) feed->wait_for(s, g0, g0 + bv);
      if (have_frames) {
        run_vision(e, src, g0, vpass, vision_emb_out, s);
      } else {
        const float* vis = vision_emb_in + static_cast<long>(g0) * T * kVisionDim;
        if (vision_emb_out)
          VCG_CUDA(cudaMemcpyAsync(vision_emb_out + static_cast<long>(g0) * T * kVisionDim, vis,
                                   static_cast<size_t>(bv) * T * kVisionDim * sizeof(float), cudaMemcpyDeviceToDevice, s));
        // caller's fp32 embeddings -> the fixed operand address of the projection GEMM
        ProfScope ps(e, s, "memcpy|head.vision_emb_in", 0, static_cast<double>(bv) * T * kVisionDim * 8);
        VCG_CUDA(cudaMemcpyAsync(e->vis_emb.p, vis, static_cast<size_t>(bv) * T * kVisionDim * sizeof(float),
                                 cudaMemcpyDeviceToDevice, s));
      }
      {
        auto it = e->vis_proj_plans.find(bv * T);
        if (it == e->vis_proj_plans.end()) {
          Epilogue ep; ep.act = ACT_RELU;
          it = e->vis_proj_plans.emplace(bv * T, build_gemm(e->vis_emb.p, kVisionDim, e->vis_w.p, e->vis_out.p, e->H, bv * T, e->H,
                                                            kVisionDim, /*fp32=*/true, ep, "head.vision_proj")).first;
        }
        ProfScope ps(e, s, gemm_kernel_name(it->second), it->second.flops, it->second.bytes);
        launch_conv_gemm(it->second, s);
      }
      TailParams hp = tp;
      hp.vis_out = e->vis_out.p;
      hp.lang_out = static_cast<const uint8_t*>(e->lang_out.p) + static_cast<size_t>(c0) * e->H * e->es();
      hp.logits = logits + static_cast<long>(g0) * 2;
      hp.probs = probs + static_cast<long>(g0) * 2;
      {
        ProfScope ps(e, s, "head_final|head.final", 2.0 * bv * (T + 1) * 128 * 2, 0);
        launch_head_final(hp, bv, e->fp32, s);
      }
    }
  }
}

// Single-modality models of the reference (--data_mode image / text): backbone embedding -> nn.Linear(D, 2) -> softmax.
//   Resnet50TSM.forward / Resnet50.forward (model/vision/resnet50_tsm.py:68-77, resnet50.py:64-73): D = T*2048
//   BertHugface.forward, pretrain_stage=False (model/lang/bert_hugface.py:98-132): D = 768 on the pooler output
void score_vision_only(vcg_engine* e, const FrameSource& src, int B, float* logits, float* probs, float* vision_emb_out,
                       cudaStream_t s) {
  VCG_REQUIRE(e->finalized, "vcg_finalize has not been called");
  VCG_REQUIRE(e->cfg.modality == VCG_MODALITY_VISION, "engine was not created for the image-only model");
  for (int g0 = 0, bv = 0; g0 < B; g0 += bv) {
    const VisionPass vpass = next_vision_pass(e, src, g0, B - g0);
    bv = vpass.n;
    const float* emb = run_vision(e, src, g0, vpass, vision_emb_out, s);
    ProfScope ps(e, s, "linear_head2|head.final", 2.0 * bv * e->T * kVisionDim * 2, 0);
    launch_linear_head2(emb, e->head_w.as<float>(), e->head_b.as<float>(), bv, e->T * kVisionDim, logits + g0 * 2L,
                        probs + g0 * 2L, s);
  }
}

void score_text_only(vcg_engine* e, const int64_t* ids, const int64_t* mask, int B, int L, float* logits, float* probs,
                     float* lang_emb_out, cudaStream_t s) {
  VCG_REQUIRE(e->finalized, "vcg_finalize has not been called");
  VCG_REQUIRE(e->cfg.modality == VCG_MODALITY_TEXT, "engine was not created for the text-only model");
  VCG_REQUIRE(B >= 0 && L >= 1 && L <= e->Lmax, "token count exceeds max_tokens of the engine");
  for (int b0 = 0; b0 < B; b0 += e->Bt) {
    const int bt = std::min(e->Bt, B - b0);
    float* pooled32 = lang_emb_out ? lang_emb_out + static_cast<long>(b0) * kBertHidden : e->pooled32.as<float>();
    run_text(e, ids, mask, b0, bt, L, pooled32, s);
    ProfScope ps(e, s, "linear_head2|head.final", 2.0 * bt * kBertHidden * 2, 0);
    launch_linear_head2(pooled32, e->head_w.as<float>(), e->head_b.as<float>(), bt, kBertHidden, logits + b0 * 2L,
                        probs + b0 * 2L, s);
  }
}

template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& ex) {
    g_last_error = ex.what();
    cudaGetLastError();   // clear sticky-free errors so the next call starts clean
    return 1;
  } catch (...) {
    g_last_error = "vcg: unknown error";
    return 1;
  }
}

}  // namespace

// =================================================================================================== C ABI
extern "C" {

const char* vcg_last_error(void) { return g_last_error.c_str(); }
const char* vcg_version(void) { return "vcg_b200 0.1 (sm_100a)"; }

int vcg_create(const vcg_config* cfg, vcg_engine** out) {
  return guarded([&] {
    VCG_REQUIRE(cfg && out, "null argument");
    VCG_REQUIRE(cfg->clip_frames >= 1 && cfg->clip_frames <= 39, "clip_frames must be in [1,39]");
    VCG_REQUIRE(cfg->max_tokens >= 1 && cfg->max_tokens <= 512, "max_tokens must be in [1,512]");
    VCG_REQUIRE(cfg->hidden_size == 128, "hidden_size must be 128");
    if (cfg->head_type != VCG_HEAD_MLP && cfg->head_type != VCG_HEAD_ATTN)
      throw Error("Unknown head_type " + std::to_string(cfg->head_type));   // two_stream.py:68
    VCG_REQUIRE(cfg->precision == VCG_PREC_BF16 || cfg->precision == VCG_PREC_FP32, "unknown precision");
    VCG_REQUIRE(cfg->modality >= VCG_MODALITY_TWO_STREAM && cfg->modality <= VCG_MODALITY_EMBED, "unknown modality");
    if (cfg->modality == VCG_MODALITY_VISION) VCG_REQUIRE(cfg->vision == VCG_VISION_R50TSM, "the image-only model needs a vision backbone");
    VCG_REQUIRE(cfg->max_batch >= 1, "max_batch must be positive");
    VCG_REQUIRE(cfg->shift_div == 8 || cfg->shift_div == 4 || cfg->shift_div == 0, "shift_div must be 8, 4 or 0 (no shift)");
    int dev = 0, major = 0, minor = 0;
    VCG_CUDA(cudaGetDevice(&dev));
    VCG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    VCG_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    VCG_REQUIRE(major == 10, "libvcg_b200 needs a Blackwell (sm_100a) GPU; there is no fallback path");
    auto e = std::make_unique<vcg_engine>();
    e->cfg = *cfg;
    e->fp32 = cfg->precision == VCG_PREC_FP32;
    e->T = cfg->clip_frames; e->Lmax = cfg->max_tokens; e->H = cfg->hidden_size;
    e->Bv = cfg->max_batch;
    e->Bt = std::max(cfg->max_batch, 256);
    e->tsm = cfg->shift_div != 0;
    if (const char* v = getenv("VCG_DEBUG_CHECKSUM")) {
      e->debug_checksum = atoi(v) != 0;
      if (e->debug_checksum) e->dbg.alloc(256 * sizeof(unsigned long long), /*zero=*/true);
    }
    *out = e.release();
  });
}

void vcg_destroy(vcg_engine* e) { delete e; }

int vcg_load_tensor(vcg_engine* e, const char* key, const void* dev_ptr, const int64_t* shape, int32_t ndim,
                    int32_t dtype, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && key && dev_ptr && (shape || ndim == 0), "null argument");
    RawTensor t;
    t.shape.assign(shape, shape + ndim);
    t.dtype = dtype;
    const size_t bytes = static_cast<size_t>(t.numel()) * (dtype == VCG_DTYPE_I64 ? 8 : 4);
    t.buf = std::make_unique<DevBuf>();
    t.buf->alloc(bytes);
    VCG_CUDA(cudaMemcpyAsync(t.buf->p, dev_ptr, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    e->raw[key] = std::move(t);
    e->finalized = false;
  });
}

int vcg_finalize(vcg_engine* e, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e, "null engine");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    e->vplans.clear();
    e->vplans_shared.clear();
    e->bplans.clear();
    e->lang_tail_plans.clear();
    e->vis_proj_plans.clear();
    if (e->cfg.vision == VCG_VISION_R50TSM && e->cfg.modality != VCG_MODALITY_TEXT) finalize_vision(e, s);
    if (e->cfg.modality != VCG_MODALITY_VISION) finalize_text(e, s);
    finalize_head(e, s);
    VCG_CUDA(cudaStreamSynchronize(s));
    e->raw.clear();   // packed copies are all the kernels read
    e->finalized = true;
  });
}

int vcg_forward(vcg_engine* e, const float* img_clip, const float* vision_emb, const int64_t* text_ids,
                const int64_t* attention_mask, int32_t B, int32_t L, float* logits, float* probs,
                float* vision_emb_out, float* lang_emb_out, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && text_ids && attention_mask && logits && probs, "null argument");
    FrameSource src;
    src.img_clip = img_clip;
    score(e, src, vision_emb, text_ids, attention_mask, B, L, logits, probs, vision_emb_out, lang_emb_out,
          static_cast<cudaStream_t>(stream));
  });
}

int vcg_forward_vision(vcg_engine* e, const float* img_clip, int32_t B, float* logits, float* probs, float* vision_emb_out,
                       void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && img_clip && logits && probs, "null argument");
    FrameSource src;
    src.img_clip = img_clip;
    score_vision_only(e, src, B, logits, probs, vision_emb_out, static_cast<cudaStream_t>(stream));
  });
}

int vcg_forward_text(vcg_engine* e, const int64_t* text_ids, const int64_t* attention_mask, int32_t B, int32_t L,
                     float* logits, float* probs, float* lang_emb_out, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && text_ids && attention_mask && logits && probs, "null argument");
    score_text_only(e, text_ids, attention_mask, B, L, logits, probs, lang_emb_out, static_cast<cudaStream_t>(stream));
  });
}

int vcg_embed(vcg_engine* e, const float* img_clip, const int64_t* text_ids, const int64_t* attention_mask, int32_t B,
              int32_t L, float* vision_emb_out, float* lang_emb_out, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && img_clip && text_ids && attention_mask && vision_emb_out && lang_emb_out, "null argument");
    VCG_REQUIRE(e->finalized, "vcg_finalize has not been called");
    VCG_REQUIRE(e->cfg.modality == VCG_MODALITY_EMBED && e->cfg.vision == VCG_VISION_R50TSM, "engine was not created for embeddings");
    VCG_REQUIRE(B >= 0 && L >= 1 && L <= e->Lmax, "token count exceeds max_tokens of the engine");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int b0 = 0; b0 < B; b0 += e->Bt)
      run_text(e, text_ids, attention_mask, b0, std::min(e->Bt, B - b0), L, lang_emb_out + static_cast<long>(b0) * kBertHidden, s);
    FrameSource src;
    src.img_clip = img_clip;
    for (int g0 = 0; g0 < B;) {
      const VisionPass vpass = next_vision_pass(e, src, g0, B - g0);
      run_vision(e, src, g0, vpass, vision_emb_out, s);
      g0 += vpass.n;
    }
  });
}

int vcg_embed_u8(vcg_engine* e, const uint8_t* frames_u8, int32_t n_frames, const int32_t* clip_start, int32_t first_start,
                 int32_t clip_stride, const int64_t* text_ids, const int64_t* attention_mask, int32_t B, int32_t L,
                 float* vision_emb_out, float* lang_emb_out, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && frames_u8 && text_ids && attention_mask && vision_emb_out && lang_emb_out, "null argument");
    VCG_REQUIRE(e->finalized, "vcg_finalize has not been called");
    VCG_REQUIRE(e->cfg.modality == VCG_MODALITY_EMBED && e->cfg.vision == VCG_VISION_R50TSM, "engine was not created for embeddings");
    VCG_REQUIRE(B >= 0 && L >= 1 && L <= e->Lmax, "token count exceeds max_tokens of the engine");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    FrameSource src;
    src.frames_u8 = frames_u8;
    src.n_frames = n_frames;
    std::vector<int32_t> starts(B);
    if (clip_start) {
      src.clip_start = clip_start;
    } else {   // regular grid: the stem runs once per distinct frame
      VCG_REQUIRE(clip_stride >= 1 && first_start >= 0 &&
                  (B == 0 || first_start + static_cast<long>(clip_stride) * (B - 1) + e->T <= n_frames),
                  "clip grid exceeds the frame buffer");
      for (int b = 0; b < B; ++b) starts[b] = first_start + b * clip_stride;
      e->st_start.ensure(static_cast<size_t>(std::max(B, 1)) * sizeof(int32_t));
      VCG_CUDA(cudaMemcpyAsync(e->st_start.p, starts.data(), static_cast<size_t>(B) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
      VCG_CUDA(cudaStreamSynchronize(s));   // `starts` is a temporary
      src.clip_start = e->st_start.as<int32_t>();
      src.clip_start_host = starts.data();
    }
    for (int b0 = 0; b0 < B; b0 += e->Bt)
      run_text(e, text_ids, attention_mask, b0, std::min(e->Bt, B - b0), L, lang_emb_out + static_cast<long>(b0) * kBertHidden, s);
    for (int g0 = 0; g0 < B;) {
      const VisionPass vpass = next_vision_pass(e, src, g0, B - g0);
      run_vision(e, src, g0, vpass, vision_emb_out, s);
      g0 += vpass.n;
    }
  });
}

int vcg_set_frame_size(vcg_engine* e, int32_t src_h, int32_t src_w, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e, "null engine");
    VCG_REQUIRE(src_h >= 1 && src_w >= 1 && src_h <= 15 * kImg && src_w <= 15 * kImg, "source frame size out of range");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    e->src_h = src_h; e->src_w = src_w;
    if (e->resizing()) {
      e->rs_xb.alloc(kImg * 2 * sizeof(int32_t));
      e->rs_yb.alloc(kImg * 2 * sizeof(int32_t));
      e->rs_xk.alloc(static_cast<size_t>(kImg) * resize_ksize(src_w, kImg) * sizeof(int32_t));
      e->rs_yk.alloc(static_cast<size_t>(kImg) * resize_ksize(src_h, kImg) * sizeof(int32_t));
      launch_resize_coeffs(src_w, kImg, e->rs_xb.as<int32_t>(), e->rs_xk.as<int32_t>(), s);
      launch_resize_coeffs(src_h, kImg, e->rs_yb.as<int32_t>(), e->rs_yk.as<int32_t>(), s);
      VCG_CUDA(cudaStreamSynchronize(s));
    }
  });
}

int vcg_score_clips_u8(vcg_engine* e, const uint8_t* frames_u8, int32_t n_frames, const int32_t* clip_start,
                       const int64_t* text_ids, const int64_t* attention_mask, int32_t B, int32_t L, float* logits,
                       float* probs, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && frames_u8 && clip_start && text_ids && attention_mask && logits && probs, "null argument");
    VCG_REQUIRE(n_frames >= e->T || B == 0, "frame buffer shorter than one clip");
    FrameSource src;
    src.frames_u8 = frames_u8;
    src.clip_start = clip_start;   // device-side starts: the gather clamps every frame index to [0, n_frames)
    src.n_frames = n_frames;
    score(e, src, nullptr, text_ids, attention_mask, B, L, logits, probs, nullptr, nullptr,
          static_cast<cudaStream_t>(stream));
  });
}

int vcg_score_clips_u8_planned(vcg_engine* e, const uint8_t* frames_u8, int32_t n_frames, const int32_t* clip_start,
                               const int32_t* clip_start_host, const int64_t* text_ids, const int64_t* attention_mask,
                               int32_t B, int32_t L, float* logits, float* probs, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && frames_u8 && clip_start && clip_start_host && text_ids && attention_mask && logits && probs, "null argument");
    VCG_REQUIRE(e->finalized, "vcg_finalize has not been called");
    check_clip_starts(clip_start_host, B, e->T, n_frames);   // before anything is enqueued
    FrameSource src;
    src.frames_u8 = frames_u8;
    src.clip_start = clip_start;
    src.clip_start_host = clip_start_host;
    src.n_frames = n_frames;
    score(e, src, nullptr, text_ids, attention_mask, B, L, logits, probs, nullptr, nullptr,
          static_cast<cudaStream_t>(stream));
  });
}

int vcg_score_video_u8(vcg_engine* e, const uint8_t* frames_u8, int32_t n_frames, int32_t first_start, int32_t clip_stride,
                       const int64_t* text_ids, const int64_t* attention_mask, int32_t B, int32_t L, float* logits,
                       float* probs, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && frames_u8 && text_ids && attention_mask && logits && probs, "null argument");
    VCG_REQUIRE(clip_stride >= 1 && first_start >= 0 && (B == 0 || first_start + static_cast<long>(clip_stride) * (B - 1) + e->T <= n_frames),
                "clip grid exceeds the frame buffer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // device-side start table for the passes that cannot share the stem
    std::vector<int32_t> starts(B);
    for (int b = 0; b < B; ++b) starts[b] = first_start + b * clip_stride;
    e->st_start.ensure(static_cast<size_t>(std::max(B, 1)) * sizeof(int32_t));
    VCG_CUDA(cudaMemcpyAsync(e->st_start.p, starts.data(), static_cast<size_t>(B) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    VCG_CUDA(cudaStreamSynchronize(s));   // `starts` is a temporary
    FrameSource src;
    src.frames_u8 = frames_u8;
    src.clip_start = e->st_start.as<int32_t>();
    src.clip_start_host = starts.data();
    src.n_frames = n_frames;
    score(e, src, nullptr, text_ids, attention_mask, B, L, logits, probs, nullptr, nullptr, s);
  });
}

int vcg_score_clips_u8_host(vcg_engine* e, const uint8_t* frames_u8_host, int32_t n_frames,
                            const int32_t* clip_start_host, const int64_t* text_ids_host,
                            const int64_t* attention_mask_host, int32_t B, int32_t L, float* logits_host,
                            float* probs_host, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && frames_u8_host && clip_start_host && text_ids_host && attention_mask_host && logits_host && probs_host,
                "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    VCG_REQUIRE(e->finalized, "vcg_finalize has not been called");
    check_clip_starts(clip_start_host, B, e->T, n_frames);   // before anything is enqueued
    const size_t fbytes = static_cast<size_t>(n_frames) * e->frame_bytes();
    const size_t tbytes = static_cast<size_t>(B) * L * sizeof(int64_t);
    e->st_frames.ensure(fbytes);
    e->st_ids.ensure(tbytes);
    e->st_mask.ensure(tbytes);
    e->st_start.ensure(static_cast<size_t>(B) * sizeof(int32_t));
    e->st_logits.ensure(static_cast<size_t>(B) * 2 * sizeof(float));
    e->st_probs.ensure(static_cast<size_t>(B) * 2 * sizeof(float));
    VCG_CUDA(cudaMemcpyAsync(e->st_ids.p, text_ids_host, tbytes, cudaMemcpyHostToDevice, s));
    VCG_CUDA(cudaMemcpyAsync(e->st_mask.p, attention_mask_host, tbytes, cudaMemcpyHostToDevice, s));
    VCG_CUDA(cudaMemcpyAsync(e->st_start.p, clip_start_host, static_cast<size_t>(B) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    // frames go up on a side stream in pieces of 128 frames (19 MB); a vision pass waits only for the frames it reads
    cudaStream_t cs = e->get_copy_stream();
    HostFeed feed;
    feed.clip_start_host = clip_start_host;
    feed.T = e->T;
    const size_t frame_bytes = e->frame_bytes();
    const int piece = std::max(1, static_cast<int>((128ull * kImg * kImg * 3) / frame_bytes));   // ~19 MB per copy
    size_t ev_i = 0;
    for (int f0 = 0; f0 < n_frames; f0 += piece, ++ev_i) {
      const int n = std::min(piece, n_frames - f0);
      VCG_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(e->st_frames.p) + f0 * frame_bytes, frames_u8_host + f0 * frame_bytes,
                               n * frame_bytes, cudaMemcpyHostToDevice, cs));
      cudaEvent_t ev = e->copy_event(ev_i);
      VCG_CUDA(cudaEventRecord(ev, cs));
      feed.ready.emplace_back(f0 + n, ev);
    }
    FrameSource src;
    src.frames_u8 = e->st_frames.as<uint8_t>();
    src.clip_start = e->st_start.as<int32_t>();
    src.clip_start_host = clip_start_host;   // regular runs (the reference's range(0, n - T, 4)) share the stem, pass by pass
    src.n_frames = n_frames;
    score(e, src, nullptr, e->st_ids.as<int64_t>(), e->st_mask.as<int64_t>(), B, L, e->st_logits.as<float>(),
          e->st_probs.as<float>(), nullptr, nullptr, s, &feed);
    VCG_CUDA(cudaMemcpyAsync(logits_host, e->st_logits.p, static_cast<size_t>(B) * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    VCG_CUDA(cudaMemcpyAsync(probs_host, e->st_probs.p, static_cast<size_t>(B) * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    VCG_CUDA(cudaStreamSynchronize(s));
  });
}

int vcg_forward_host(vcg_engine* e, const float* img_clip_host, const float* vision_emb_host,
                     const int64_t* text_ids_host, const int64_t* attention_mask_host, int32_t B, int32_t L,
                     float* logits_host, float* probs_host, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && (img_clip_host || vision_emb_host) && text_ids_host && attention_mask_host && logits_host && probs_host,
                "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t vbytes = img_clip_host ? static_cast<size_t>(B) * e->T * 3 * kImg * kImg * sizeof(float)
                                        : static_cast<size_t>(B) * e->T * kVisionDim * sizeof(float);
    const size_t tbytes = static_cast<size_t>(B) * L * sizeof(int64_t);
    e->st_frames.ensure(vbytes);
    e->st_ids.ensure(tbytes);
    e->st_mask.ensure(tbytes);
    e->st_logits.ensure(static_cast<size_t>(B) * 2 * sizeof(float));
    e->st_probs.ensure(static_cast<size_t>(B) * 2 * sizeof(float));
    VCG_CUDA(cudaMemcpyAsync(e->st_ids.p, text_ids_host, tbytes, cudaMemcpyHostToDevice, s));
    VCG_CUDA(cudaMemcpyAsync(e->st_mask.p, attention_mask_host, tbytes, cudaMemcpyHostToDevice, s));
    // the vision input is copied on a side stream in pieces of max_batch clips: the text stream (and earlier vision
    // passes) run underneath; every vision pass waits only for its own piece
    cudaStream_t cs = e->get_copy_stream();
    HostFeed feed;
    const size_t clip_bytes = vbytes / std::max(B, 1);
    const int piece = img_clip_host ? e->Bv : std::max(B, 1);
    const uint8_t* hsrc = reinterpret_cast<const uint8_t*>(img_clip_host ? img_clip_host : vision_emb_host);
    size_t ev_i = 0;
    for (int c0 = 0; c0 < B; c0 += piece, ++ev_i) {
      const int n = std::min(piece, B - c0);
      VCG_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(e->st_frames.p) + c0 * clip_bytes, hsrc + c0 * clip_bytes, n * clip_bytes,
                               cudaMemcpyHostToDevice, cs));
      cudaEvent_t ev = e->copy_event(ev_i);
      VCG_CUDA(cudaEventRecord(ev, cs));
      feed.ready.emplace_back(c0 + n, ev);
    }
    FrameSource src;
    if (img_clip_host) src.img_clip = e->st_frames.as<float>();
    score(e, src, img_clip_host ? nullptr : e->st_frames.as<float>(), e->st_ids.as<int64_t>(), e->st_mask.as<int64_t>(), B, L,
          e->st_logits.as<float>(), e->st_probs.as<float>(), nullptr, nullptr, s, &feed);
    VCG_CUDA(cudaMemcpyAsync(logits_host, e->st_logits.p, static_cast<size_t>(B) * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    VCG_CUDA(cudaMemcpyAsync(probs_host, e->st_probs.p, static_cast<size_t>(B) * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    VCG_CUDA(cudaStreamSynchronize(s));
  });
}

int64_t vcg_launch_count(const vcg_engine* e) { return e ? e->launches : 0; }

int vcg_debug_pdl_window(int32_t from, int32_t to) {
  vcg::PdlDebugWindow& w = vcg::pdl_debug_window();
  w.idx = 0; w.from = from; w.to = to;
  return w.idx;
}
int vcg_debug_checksums(vcg_engine* e, uint64_t* out_host, int32_t max_n, int32_t* n_out, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(e && out_host && n_out, "null argument");
    VCG_REQUIRE(e->debug_checksum, "engine was not created with VCG_DEBUG_CHECKSUM=1");
    VCG_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    const int n = std::min(e->dbg_n, max_n);
    VCG_CUDA(cudaMemcpy(out_host, e->dbg.p, static_cast<size_t>(n) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    *n_out = n;
  });
}

int vcg_profile_begin(vcg_engine* e) {
  return guarded([&] {
    VCG_REQUIRE(e, "null engine");
    e->prof.clear();
    e->events_used = 0;
    e->profiling = true;
  });
}

int vcg_profile_end(vcg_engine* e, void* stream, vcg_profile_entry* out, int32_t max_entries, int32_t* n_out) {
  return guarded([&] {
    VCG_REQUIRE(e && n_out, "null argument");
    e->profiling = false;
    VCG_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    // executed work of the token-packed BERT kernels: scale the padded figures by packed rows / padded rows
    double pack_ratio = 1.0;
    if (e->last_bert_rows > 0) {
      int32_t total = 0;
      VCG_CUDA(cudaMemcpy(&total, e->pk_total.p, sizeof total, cudaMemcpyDeviceToHost));
      pack_ratio = static_cast<double>(total) / e->last_bert_rows;
    }
    std::map<std::string, vcg_profile_entry> agg;
    for (const ProfRec& r : e->prof) {
      float ms = 0.f;
      VCG_CUDA(cudaEventElapsedTime(&ms, r.start, r.stop));
      auto it = agg.find(r.name);
      if (it == agg.end()) {
        vcg_profile_entry en{};
        snprintf(en.name, sizeof en.name, "%s", r.name.c_str());
        it = agg.emplace(r.name, en).first;
      }
      it->second.launches += 1;
      it->second.ms += ms;
      it->second.flops += r.flops * (r.packed ? pack_ratio : 1.0);
      it->second.bytes += r.bytes * (r.packed ? pack_ratio : 1.0);
    }
    int n = 0;
    for (auto& kv : agg) {
      if (out && n < max_entries) out[n] = kv.second;
      ++n;
    }
    *n_out = n;
    e->prof.clear();
    e->events_used = 0;
  });
}

// --------------------------------------------------------------------------------------------- operators
int vcg_op_preprocess_u8(const uint8_t* frames_u8, const int32_t* frame_index, int32_t n, void* out_padded,
                         int32_t precision, void* stream) {
  return guarded([&] {
    launch_preprocess_u8(frames_u8, frame_index, n, out_padded, precision == VCG_PREC_FP32, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_resize_u8(const uint8_t* frames_u8, int32_t n, int32_t src_h, int32_t src_w, uint8_t* out_u8, void* out_padded,
                     int32_t precision, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(frames_u8 && (out_u8 || out_padded), "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // coefficient tables per (device, source extent), built once and kept for the life of the process
    struct Axis { DevBuf b, k; };
    static std::map<std::pair<int, int>, std::unique_ptr<Axis>> cache;
    int dev = 0;
    VCG_CUDA(cudaGetDevice(&dev));
    auto axis = [&](int in_size) -> Axis& {
      auto& slot = cache[{dev, in_size}];
      if (!slot) {
        slot = std::make_unique<Axis>();
        slot->b.alloc(kImg * 2 * sizeof(int32_t));
        slot->k.alloc(static_cast<size_t>(kImg) * resize_ksize(in_size, kImg) * sizeof(int32_t));
        launch_resize_coeffs(in_size, kImg, slot->b.as<int32_t>(), slot->k.as<int32_t>(), s);
        VCG_CUDA(cudaStreamSynchronize(s));
      }
      return *slot;
    };
    Axis& ax = axis(src_w);
    Axis& ay = axis(src_h);
    launch_resize_preprocess_u8(frames_u8, nullptr, nullptr, 1, n, 0, src_h, src_w, ax.b.as<int32_t>(), ax.k.as<int32_t>(),
                                ay.b.as<int32_t>(), ay.k.as<int32_t>(), out_padded, out_u8, precision == VCG_PREC_FP32, s);
  });
}
int vcg_op_nchw_to_stem(const float* img, int32_t n, void* out_padded, int32_t precision, void* stream) {
  return guarded([&] { launch_nchw_to_stem(img, n, out_padded, precision == VCG_PREC_FP32, static_cast<cudaStream_t>(stream)); });
}
int vcg_op_gemm(const void* A, int64_t lda, const void* W, const float* bias, const void* residual, int32_t ld_res,
                void* out, int32_t ld_out, int32_t M, int32_t N, int32_t K, int32_t act, int32_t precision,
                void* stream) {
  return guarded([&] {
    Epilogue ep;
    ep.bias = bias; ep.residual = residual; ep.ld_res = ld_res; ep.act = act;
    ConvGemmLaunch L = build_gemm(A, lda, W, out, ld_out, M, N, K, precision == VCG_PREC_FP32, ep);
    launch_conv_gemm(L, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_conv2d_nhwc(const void* in, int32_t n, int32_t H, int32_t W, int32_t Cin, const void* weight,
                       const float* bias, const void* residual, void* out, int32_t Cout, int32_t k, int32_t stride,
                       int32_t act, int32_t precision, const void* tsm_in, int32_t tsm_in_ch, void* tsm_out,
                       int32_t tsm_fold, int32_t clip_frames, void* stream) {
  return guarded([&] {
    Epilogue ep;
    ep.bias = bias; ep.residual = residual; ep.ld_res = Cout; ep.act = act;
    ep.tsm_out = tsm_out; ep.tsm_fold = tsm_fold; ep.tsm_ld = 2 * tsm_fold; ep.T = clip_frames;
    ConvGemmLaunch L = build_conv(in, n, H, W, Cin, weight, Cout, k, stride, out, precision == VCG_PREC_FP32, ep, tsm_in, tsm_in_ch);
    launch_conv_gemm(L, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_bottleneck_tail(const void* in, int32_t n, int32_t H, int32_t W, int32_t P, int32_t stride, const void* w2,
                           const float* bias2, const void* w3, const float* bias3, const void* residual, void* out,
                           void* tsm_out, int32_t tsm_fold, int32_t clip_frames, int32_t variant, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(in && w2 && bias2 && w3 && bias3 && out, "null argument");
    Epilogue ep;
    ep.bias = bias3; ep.residual = residual; ep.ld_res = 4 * P; ep.act = ACT_RELU;
    ep.tsm_out = tsm_out; ep.tsm_fold = tsm_fold; ep.tsm_ld = 2 * tsm_fold; ep.T = clip_frames;
    Conv23Launch L;
    if (variant == 1 && P == 128) {
      VCG_REQUIRE(stride == 1 && W >= 16 && H >= 16 && tsm_out == nullptr, "halo variant, P = 128: stride 1, W, H >= 16, no TSM scatter");
      L = build_conv23h(in, n, H, W, w2, bias2, w3, out, ep, "op.tail_h2", 128);
    } else if (variant == 1) {
      VCG_REQUIRE(conv23h_ok(P, stride, H, W, false) || c23h_policy() == 0, "halo variant: P = 64, stride 1, W % 8 == 0, W, H >= 16");
      L = build_conv23h(in, n, H, W, w2, bias2, w3, out, ep, "op.tail_h");
    } else {
      L = build_conv23(in, n, H, W, P, stride, w2, bias2, w3, out, ep, "op.tail");
    }
    launch_conv23(L, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_stem_conv(const void* in_padded, int32_t n, const void* weight, const float* bias, void* out,
                     int32_t precision, void* stream) {
  return guarded([&] {
    Epilogue ep;
    ep.bias = bias; ep.act = ACT_RELU;
    ConvGemmLaunch L = build_stem(in_padded, n, kStemHp, kStemWp, kStemOut, kStemOut, weight, 64, out, precision == VCG_PREC_FP32, ep);
    launch_conv_gemm(L, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_stem_conv_act(const void* in_padded, int32_t n, const void* weight, const float* bias, void* out, int32_t act,
                         int32_t precision, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(in_padded && weight && out, "null argument");
    Epilogue ep;
    ep.bias = bias; ep.act = act;
    ConvGemmLaunch L = build_stem(in_padded, n, kStemHp, kStemWp, kStemOut, kStemOut, weight, 64, out, precision == VCG_PREC_FP32, ep);
    launch_conv_gemm(L, static_cast<cudaStream_t>(stream));
  });
}
int32_t vcg_op_bn_partials(int64_t rows, int32_t C) {
  if (rows < 1 || C < 64 || C > 2048 || C % 8 != 0) return 0;
  return bn_stats_partials(rows, C);
}
int vcg_op_bn_batch_stats(const void* x, int64_t rows, int32_t C, float eps, double* partial, float* mean, float* rstd,
                          int32_t precision, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(x && partial && mean && rstd, "null argument");
    launch_bn_batch_stats(x, rows, C, eps, partial, mean, rstd, precision == VCG_PREC_FP32, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_bn_apply(const void* x, int64_t rows, int32_t C, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, const void* residual, int32_t relu, void* out, int32_t precision, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(x && mean && rstd && gamma && beta && out, "null argument");
    launch_bn_apply(x, rows, C, mean, rstd, gamma, beta, residual, relu != 0, out, precision == VCG_PREC_FP32,
                    static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_tsm_shift(const void* x, int64_t n, int32_t hw, int32_t C, int32_t clip_frames, int32_t fold, void* out,
                     int32_t precision, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(x && out, "null argument");
    launch_tsm_shift(x, n, hw, C, clip_frames, fold, out, precision == VCG_PREC_FP32, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_avgpool(const void* x, int32_t n, int32_t hw, int32_t C, float* out, int32_t precision, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(x && out && C % 8 == 0 && hw >= 1, "bad argument");
    launch_avgpool(x, n, hw, C, out, nullptr, static_cast<cudaStream_t>(stream), precision == VCG_PREC_FP32);
  });
}
int vcg_op_maxpool_tsm(const void* in, int32_t n, void* out, void* out_shifted, int32_t clip_frames,
                       int32_t shift_div, int32_t precision, void* stream) {
  return guarded([&] {
    launch_maxpool_tsm(in, n, out, out_shifted, clip_frames, shift_div > 0 ? 64 / shift_div : 0, precision == VCG_PREC_FP32,
                       static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_bert_attention(const void* qkv, const int64_t* attention_mask, void* ctx, int32_t B, int32_t L,
                          int32_t precision, void* stream) {
  return guarded([&] { launch_bert_attention(qkv, attention_mask, nullptr, nullptr, ctx, B, L, precision == VCG_PREC_FP32, static_cast<cudaStream_t>(stream)); });
}
int vcg_op_bert_attention_packed(const void* qkv, const int32_t* cu, const uint8_t* key_ok, void* ctx, int32_t B,
                                 int32_t max_len, int64_t rows, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(qkv && cu && key_ok && ctx, "null argument");
    launch_bert_attention(qkv, nullptr, cu, key_ok, ctx, B, max_len, false, static_cast<cudaStream_t>(stream), rows);
  });
}
int vcg_op_cut_points(const float* logits, const int32_t* video_offsets, int32_t n_videos, int32_t clip_frames,
                      int32_t max_offset, int32_t cap, int32_t* labels_out, int32_t* cut_points, int32_t* counts,
                      void* stream) {
  return guarded([&] {
    VCG_REQUIRE(logits && video_offsets && cut_points && counts && cap >= 1 && clip_frames >= 1 && max_offset >= 1,
                "bad argument");
    launch_cut_points(logits, video_offsets, n_videos, clip_frames, max_offset, cap, labels_out, cut_points, counts,
                      static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_pr_hits(const int32_t* gt, const int32_t* gt_offsets, const int32_t* pred, const int32_t* pred_offsets,
                   int32_t n_videos, int32_t* hits, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(gt && gt_offsets && pred && pred_offsets && hits, "null argument");
    launch_pr_hits(gt, gt_offsets, pred, pred_offsets, n_videos, hits, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_auc_ap(const float* scores, const int32_t* labels, const int32_t* video_offsets, int32_t n_videos, double* auc,
                  double* ap, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(scores && labels && video_offsets && auc && ap, "null argument");
    launch_auc_ap(scores, labels, video_offsets, n_videos, auc, ap, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_mlp_chain(const float* x0, int32_t dim0, int64_t stride0, const float* x1, int32_t dim1, int64_t stride1,
                     int32_t rows, const vcg_mlp_op* ops, int32_t n_ops, float* out, int64_t out_stride, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(x0 && ops && out && (dim1 == 0 || x1), "null argument");
    launch_mlp_chain(x0, dim0, stride0, x1, dim1, stride1, rows, ops, n_ops, out, out_stride, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_cross_attention(const vcg_cross_attn_params* p, const float* lang, const float* vision, int32_t B, int32_t T,
                           float* out, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(p && lang && vision && out, "null argument");
    launch_cross_attention(*p, lang, vision, B, T, out, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_self_attention_first(const vcg_self_attn_params* p, const float* vision, const float* lang, int32_t B, int32_t T,
                                float* out, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(p && lang && vision && out, "null argument");
    launch_self_attention_first(*p, vision, lang, B, T, out, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_bilinear_contract(const float* y, const float* x1, const float* bias, int32_t rows, int32_t in1,
                             int32_t out_features, float* out, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(y && x1 && out, "null argument");
    launch_bilinear_contract(y, x1, bias, rows, in1, out_features, out, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_center_attention(const vcg_center_attn_params* p, const float* x, int32_t B, int32_t W, float* out, void* stream) {
  return guarded([&] {
    VCG_REQUIRE(p && x && out, "null argument");
    launch_center_attention(*p, x, B, W, out, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_window_stack(const vcg_window_stack_params* p, const float* x, int32_t B, int32_t W, float* logits, float* probs,
                        void* stream) {
  return guarded([&] {
    VCG_REQUIRE(p && x && logits && probs, "null argument");
    launch_window_stack(*p, x, B, W, logits, probs, static_cast<cudaStream_t>(stream));
  });
}
int vcg_op_layernorm(const void* x, const float* gamma, const float* beta, void* y, int32_t rows, int32_t cols,
                     float eps, int32_t precision, void* stream) {
  return guarded([&] { launch_layernorm(x, gamma, beta, y, rows, cols, eps, precision == VCG_PREC_FP32, static_cast<cudaStream_t>(stream)); });
}

}  // extern "C"
