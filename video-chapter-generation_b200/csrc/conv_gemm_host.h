// Host-side builders for conv_gemm_kernel launches: tensor maps, patch geometry, tap tables.
#pragma once
#include "conv_gemm.cuh"
#include "conv23.cuh"
#include "conv23h.cuh"
#include "conv23h2.cuh"
#include "conv23t.cuh"
#include "tensormap.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace vcg {

struct ConvGemmLaunch {
  ConvGemmParams p;
  int block_n = 0;
  bool fp32 = false;
  bool cg2 = false;   // CTA-pair kernel (cta_group::2): 256-row tiles, each CTA stages half of the B tile
  bool lnf = false;   // epilogue with the LayerNorm terms
  int grid = 0;
  double flops = 0;   // 2*M*N*K of useful work (for reporting)
  double bytes = 0;   // algorithmic HBM bytes: activations in + weights + out (+ residual, + shifted copy)
  const char* name = "";
};

int sm_count();   // conv_gemm.cu

// CTA-pair policy: VCG_CG2=0 never, =1 whenever possible, unset -> per-layer heuristic (use_cg2)
int cg2_policy();   // conv_gemm.cu

inline int elem_size(bool fp32) { return fp32 ? 4 : 2; }
inline int block_k(bool fp32) { return fp32 ? 32 : 64; }

inline int pick_block_n(int N, bool fp32) {
  const int cap = fp32 ? 128 : 256;
  if (N <= 64) return 64;
  if (N <= 128 || cap == 128) return 128;
  // N = 768 (BERT hidden): four 192-wide N tiles instead of three 256-wide ones -> 1/3 more tiles of 3/4 the size,
  // which fills the 148 SMs' last wave much better (330 tiles = 2.2 waves vs 440 tiles = 2.97 waves at 14 k rows)
  if (N == 768) return 192;
  return 256;
}

inline int pow2_divisor(int x, int cap) {
  int d = 1;
  while (d * 2 <= cap && x % (d * 2) == 0) d *= 2;
  return d;
}

// Output patch (bw, bh, nf) covered by one 128-row M tile.
inline void pick_patch(int Wo, int Ho, int Nimg, int& bw, int& bh, int& nf) {
  if (Ho == 1 && Nimg == 1) { bw = 128; bh = 1; nf = 1; return; }   // plain GEMM: W axis = rows
  if (Wo % 2 == 1 && Wo * Ho <= 64 && Nimg % 128 != 0) {            // small odd maps (7x7), small batches
    bw = Wo; bh = Ho; nf = 128 / (Wo * Ho);
    return;
  }
  bw = pow2_divisor(Wo, 16);
  bh = pow2_divisor(Ho, 128 / bw);
  nf = 128 / (bw * bh);
}

// The pair kernel pays when the K loop is long enough to be limited by L2->SM operand traffic (the single-CTA kernel
// saturates it at ~900 TFLOP/s with 128x256 tiles); short-K / residual layers are HBM-bound either way.
inline bool use_cg2(int N, int num_kb, long m_tiles, bool fp32) {
  if (fp32) return false;
  const int pol = cg2_policy();
  if (pol == 0) return false;
  if (m_tiles < 2) return false;
  if (pol == 1) return true;
  // measured on B200 (tools/bench_layer.py, VCG_CG2=0 vs 1): +8..13 % for K loops of >= 16 blocks (3x3 convs of
  // layer3/4, FFN-out, conv1 of layer3/4) and for the wide 12-block GEMMs (QKV, FFN-in); -5..-14 % for the short-K,
  // HBM-bound layers (conv3, downsample, conv1 of layer1/2)
  return num_kb >= 16 || (num_kb >= 12 && N >= 1024);
}

inline void finish_launch(ConvGemmLaunch& L, int Wo, int Ho, int Nimg, int N, bool fp32) {
  ConvGemmParams& p = L.p;
  p.Wo = Wo; p.Ho = Ho; p.Nimg = Nimg; p.N = N;
  p.tiles_w = (Wo + p.bw - 1) / p.bw;
  p.tiles_h = (Ho + p.bh - 1) / p.bh;
  p.tiles_n = (Nimg + p.nf - 1) / p.nf;
  L.block_n = pick_block_n(N, fp32);
  L.fp32 = fp32;
  p.n_tiles = (N + L.block_n - 1) / L.block_n;
  const long m_tiles = static_cast<long>(p.tiles_w) * p.tiles_h * p.tiles_n;
  L.cg2 = use_cg2(N, p.n_taps * p.cpt, m_tiles, fp32);
  p.a_bytes = static_cast<uint32_t>(p.bw * p.bh * p.nf) * 128u;
  p.b_bytes = static_cast<uint32_t>(L.cg2 ? L.block_n / 2 : L.block_n) * 128u;
  if (L.cg2) {
    const long pairs = (m_tiles + 1) / 2 * p.n_tiles;
    L.grid = 2 * static_cast<int>(std::min<long>(pairs, sm_count() / 2));
  } else {
    L.grid = static_cast<int>(std::min<long>(m_tiles * p.n_tiles, sm_count()));
  }
  VCG_REQUIRE(N % 8 == 0, "output channels must be a multiple of 8");
  VCG_REQUIRE(p.bw * p.bh * p.nf <= 128, "patch larger than the M tile");
}

struct Epilogue {
  const float* bias = nullptr;
  const void* residual = nullptr;
  int ld_res = 0;
  int act = ACT_NONE;
  void* tsm_out = nullptr;   // shifted copy of channels [0, 2*fold) for the next bottleneck
  int tsm_ld = 0, tsm_fold = 0, T = 1;
  // residual given per UNIQUE frame and shared by overlapping clips: image (clip b, frame t) reads frame b*stride + t
  int res_clip_T = 0, res_clip_stride = 0;
  // LayerNorm folded into the GEMM (bf16 plain GEMMs; ConvGemmParams)
  const float2* a_stats = nullptr;
  const float* ln_c1 = nullptr;
  const float2* res_stats = nullptr;
  const float* res_gamma = nullptr;
  const float* res_beta = nullptr;
  float2* out_stats = nullptr;
  int ln_dim = 768;
  float ln_eps = 1e-12f;
  bool ln_fused() const { return a_stats || res_stats || out_stats; }
};

// (C, W, H, T, clip) view of a per-unique-frame NHWC tensor: clip b, frame t -> unique frame b*clip_stride + t
inline CUtensorMap clip_view_map(const void* ptr, int C, int W, int H, int T, int n_clips, int clip_stride, int box_c,
                                 int bw, int bh, int nf, bool fp32, int nf_clips = 1) {
  const uint64_t es = elem_size(fp32);
  const uint64_t img = static_cast<uint64_t>(H) * W * C * es;
  const uint64_t dims[5] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                            static_cast<uint64_t>(T), static_cast<uint64_t>(n_clips)};
  const uint64_t str[4] = {static_cast<uint64_t>(C) * es, static_cast<uint64_t>(W) * C * es, img, img * clip_stride};
  const uint32_t box[5] = {static_cast<uint32_t>(box_c), static_cast<uint32_t>(bw), static_cast<uint32_t>(bh),
                           static_cast<uint32_t>(nf), static_cast<uint32_t>(nf_clips)};
  return make_tensor_map(ptr, fp32, 5, dims, str, box);
}

// [Nimg, Ho, Wo, ld] bf16 tensor seen as (C, W, H, 1, N) with a (64 x bw x bh x 1 x nf) box: the C tiles of the epilogue
inline CUtensorMap c_tile_map(const void* ptr, long ld, const ConvGemmParams& p) {
  const uint64_t dims[5] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.Wo), static_cast<uint64_t>(p.Ho), 1,
                            static_cast<uint64_t>(p.Nimg)};
  const uint64_t row = static_cast<uint64_t>(ld) * 2;
  const uint64_t img = row * p.Wo * p.Ho;
  const uint64_t str[4] = {row, row * p.Wo, img, img};
  const uint32_t box[5] = {64, static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), 1, static_cast<uint32_t>(p.nf)};
  return make_tensor_map(ptr, false, 5, dims, str, box);
}

inline void set_epilogue(ConvGemmLaunch& L, void* out, int ld_out, const Epilogue& e) {
  ConvGemmParams& p = L.p;
  p.out = out; p.ld_out = ld_out;
  p.bias = e.bias; p.residual = e.residual; p.ld_res = e.ld_res; p.act = e.act;
  p.tsm_out = e.tsm_out; p.tsm_ld = e.tsm_ld; p.tsm_fold = e.tsm_fold; p.T = e.T > 0 ? e.T : 1;
  if (e.tsm_out) VCG_REQUIRE(e.tsm_fold % 32 == 0, "TSM fold must be a multiple of 32 channels");
  L.lnf = e.ln_fused();
  if (L.lnf) {
    VCG_REQUIRE(!L.fp32 && p.Ho == 1 && p.Nimg == 1, "LayerNorm-fused epilogue: bf16 plain GEMMs only");
    VCG_REQUIRE((e.a_stats == nullptr) == (e.ln_c1 == nullptr), "a_stats and ln_c1 go together");
    VCG_REQUIRE(e.res_stats == nullptr || (e.residual && e.res_gamma && e.res_beta), "res_stats needs residual, gamma, beta");
    p.a_stats = e.a_stats; p.ln_c1 = e.ln_c1; p.res_stats = e.res_stats; p.res_gamma = e.res_gamma; p.res_beta = e.res_beta;
    p.out_stats = e.out_stats; p.ln_inv_dim = 1.0f / static_cast<float>(e.ln_dim); p.ln_eps = e.ln_eps;
  }
  if (!L.fp32) {
    // smem split: residual layers are HBM-bound -> short K loops get few A/B stages and a deep residual-prefetch ring
    const int stage_bytes = kBlockM * 128 + static_cast<int>(p.b_bytes);
    const int budget = 224 * 1024;
    // C ring: a multiple of the sub-tiles per output tile (k_sub = BLOCK_N / 64), so that ring slot s always belongs to the
    // same epilogue group (conv_gemm.cuh, C producer) - a ring of e.g. 5 slots for 4 groups lets a fast group run past
    // another group's use of its slot.  Layers with a residual want a second round of slots so that the next tile's
    // residual prefetch starts early; everything else goes to A/B stages: deeper rings are what the long-K GEMMs need
    // (measured at 14 k rows: 3 / 4 / 5 stages -> FFN-out 69 / 62 / 57 us)
    const int k_sub = L.block_n / 64;
    const int min_slots = std::max(kMinCSlots, k_sub) + ((std::max(kMinCSlots, k_sub) % k_sub) ? k_sub - std::max(kMinCSlots, k_sub) % k_sub : 0);
    const int max_stages = std::min(kMaxStages, (budget - min_slots * kCBytes) / stage_bytes);
    const int num_kb = p.n_taps * p.cpt;
    p.n_stages = e.residual ? std::max(2, std::min(max_stages, num_kb + 1)) : max_stages;
    // the GELU epilogue is the long one: one A/B stage less buys two more C slots, so a group's next tile does not wait
    // for its previous TMA store to finish reading (FFN-in at 25 600 rows: 111.1 -> 108.9 us; QKV / FFN-out prefer depth)
    if (!e.residual && e.act == ACT_GELU && p.n_stages > 4) p.n_stages -= 1;
    if (const char* v = getenv("VCG_STAGES")) {   // tuning knob (tools/bench_layer.py)
      const int forced = atoi(v);
      if (forced >= 2 && (budget - forced * stage_bytes) / kCBytes >= min_slots) p.n_stages = std::min(forced, kMaxStages);
    }
    p.n_cslots = std::min(kMaxCSlots, (budget - p.n_stages * stage_bytes) / kCBytes);
    p.n_cslots -= p.n_cslots % k_sub;
    VCG_REQUIRE(p.n_stages >= 2 && p.n_cslots >= min_slots && p.n_cslots % k_sub == 0, "shared-memory split failed");
    VCG_REQUIRE(p.N % 32 == 0, "bf16 path: output channels must be a multiple of 32");
    VCG_REQUIRE(ld_out % 8 == 0 && (e.residual == nullptr || e.ld_res % 8 == 0), "row strides must be multiples of 16 bytes");
    p.out_map = c_tile_map(out, ld_out, p);
    p.res_map = e.residual ? c_tile_map(e.residual, e.ld_res, p) : p.out_map;
    if (e.residual && e.res_clip_T > 0) {
      VCG_REQUIRE(e.res_clip_T % p.nf == 0 && p.Nimg % e.res_clip_T == 0 && e.ld_res == p.N, "clip-view residual: tile frames must stay inside a clip");
      p.res_map = clip_view_map(e.residual, p.N, p.Wo, p.Ho, e.res_clip_T, p.Nimg / e.res_clip_T, e.res_clip_stride, 64, p.bw,
                                p.bh, p.nf, false);
      p.res_clip_T = e.res_clip_T;
    }
  }
}

inline CUtensorMap weight_map(const void* W, int N, int K, int block_n, bool fp32) {
  const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
  const uint64_t str[1] = {static_cast<uint64_t>(K) * elem_size(fp32)};
  const uint32_t box[2] = {static_cast<uint32_t>(block_k(fp32)), static_cast<uint32_t>(block_n)};
  return make_tensor_map(W, fp32, 2, dims, str, box);
}

// Plain GEMM: out[M,N] = act(A[M,K] * W[N,K]^T + bias (+ residual)); lda in elements.
inline ConvGemmLaunch build_gemm(const void* A, long lda, const void* W, void* out, int ld_out, int M, int N, int K,
                                 bool fp32, const Epilogue& e, const char* name = "gemm") {
  ConvGemmLaunch L;
  memset(&L.p, 0, sizeof L.p);
  L.name = name;
  const int es = elem_size(fp32), bk = block_k(fp32);
  VCG_REQUIRE(K % bk == 0, "GEMM K must be a multiple of the K block");
  VCG_REQUIRE((lda * es) % 16 == 0, "GEMM row stride must be a multiple of 16 bytes");
  ConvGemmParams& p = L.p;
  p.bw = 128; p.bh = 1; p.nf = 1;
  const uint64_t dims[5] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M), 1, 1, 1};
  const uint64_t big = static_cast<uint64_t>(lda) * es * static_cast<uint64_t>(M);
  const uint64_t plane = (big + 15) / 16 * 16;
  const uint64_t str[4] = {static_cast<uint64_t>(lda) * es, plane, plane, plane};
  const uint32_t box[5] = {static_cast<uint32_t>(bk), 128, 1, 1, 1};
  p.a_map[0] = make_tensor_map(A, fp32, 5, dims, str, box);
  for (int i = 1; i < 4; ++i) p.a_map[i] = p.a_map[0];
  p.n_taps = 1; p.cpt = K / bk;
  p.taps[0] = TapDesc{0, 0, 0, 0, 0};
  p.tsm_split_cb = 0; p.tsm_map = 0;
  finish_launch(L, /*Wo=*/M, /*Ho=*/1, /*Nimg=*/1, N, fp32);
  p.b_map = weight_map(W, N, K, L.cg2 ? L.block_n / 2 : L.block_n, fp32);
  set_epilogue(L, out, ld_out, e);
  L.flops = 2.0 * M * N * K;
  L.bytes = static_cast<double>(es) * (static_cast<double>(M) * K + static_cast<double>(N) * K +
                                       static_cast<double>(M) * N * (e.residual ? 2.0 : 1.0));
  return L;
}

// NHWC convolution, kernel k x k (1 or 3), stride 1 or 2, pad = k/2.  Weights packed [Cout][kh][kw][Cin].
// tsm_in: optional shifted buffer [Nimg*H*W, tsm_in_ch] that replaces channels [0, tsm_in_ch) of the input
// (temporal shift folded into the A-operand load; 1x1 stride-1 only).
inline ConvGemmLaunch build_conv(const void* in, int Nimg, int H, int W, int Cin, const void* Wp, int Cout, int k,
                                 int stride, void* out, bool fp32, const Epilogue& e, const void* tsm_in = nullptr,
                                 int tsm_in_ch = 0, const char* name = "conv") {
  ConvGemmLaunch L;
  memset(&L.p, 0, sizeof L.p);
  L.name = name;
  const int es = elem_size(fp32), bk = block_k(fp32);
  VCG_REQUIRE(Cin % bk == 0, "conv input channels must be a multiple of the K block");
  VCG_REQUIRE(k == 1 || k == 3, "conv kernel must be 1x1 or 3x3");
  VCG_REQUIRE(stride == 1 || stride == 2, "conv stride must be 1 or 2");
  const int pad = k / 2;
  const int Ho = H / stride, Wo = W / stride;
  if (stride == 2) VCG_REQUIRE(H % 2 == 0 && W % 2 == 0, "stride-2 conv needs even input size");
  ConvGemmParams& p = L.p;
  pick_patch(Wo, Ho, Nimg, p.bw, p.bh, p.nf);
  const uint32_t box[5] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), 1,
                           static_cast<uint32_t>(p.nf)};
  const uint64_t img = static_cast<uint64_t>(H) * W * Cin * es;
  if (stride == 1) {
    const uint64_t dims[5] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H), 1,
                              static_cast<uint64_t>(Nimg)};
    const uint64_t str[4] = {static_cast<uint64_t>(Cin) * es, static_cast<uint64_t>(W) * Cin * es, img, img};
    p.a_map[0] = make_tensor_map(in, fp32, 5, dims, str, box);
    for (int i = 1; i < 4; ++i) p.a_map[i] = p.a_map[0];
  } else {
    // four parity views: (row parity, col parity) -> every tap of a stride-2 conv is a dense box in one view
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const uint8_t* base = static_cast<const uint8_t*>(in) + (static_cast<uint64_t>(ph) * W + pw) * Cin * es;
        const uint64_t dims[5] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W / 2),
                                  static_cast<uint64_t>(H / 2), 1, static_cast<uint64_t>(Nimg)};
        const uint64_t str[4] = {2ull * Cin * es, 2ull * W * Cin * es, img, img};
        p.a_map[ph * 2 + pw] = make_tensor_map(base, fp32, 5, dims, str, box);
      }
  }
  p.n_taps = k * k;
  p.cpt = Cin / bk;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw) {
      TapDesc t{};
      const int oh = kh - pad, ow = kw - pad;
      if (stride == 1) {
        t.dh = static_cast<int16_t>(oh); t.dw = static_cast<int16_t>(ow); t.map = 0;
      } else {
        const int ph = oh & 1, pw = ow & 1;
        t.dh = static_cast<int16_t>((oh - ph) / 2); t.dw = static_cast<int16_t>((ow - pw) / 2);
        t.map = static_cast<int8_t>(ph * 2 + pw);
      }
      t.c_off = 0; t.plane = 0;
      p.taps[kh * k + kw] = t;
    }
  p.tsm_split_cb = 0; p.tsm_map = 0;
  if (tsm_in) {
    VCG_REQUIRE(k == 1 && stride == 1, "temporal shift is folded into 1x1 stride-1 convs only");
    VCG_REQUIRE(tsm_in_ch % bk == 0 && tsm_in_ch <= Cin, "shifted channel count must be a multiple of the K block");
    const uint64_t simg = static_cast<uint64_t>(H) * W * tsm_in_ch * es;
    const uint64_t dims[5] = {static_cast<uint64_t>(tsm_in_ch), static_cast<uint64_t>(W), static_cast<uint64_t>(H), 1,
                              static_cast<uint64_t>(Nimg)};
    const uint64_t str[4] = {static_cast<uint64_t>(tsm_in_ch) * es, static_cast<uint64_t>(W) * tsm_in_ch * es, simg,
                             simg};
    p.a_map[1] = make_tensor_map(tsm_in, fp32, 5, dims, str, box);
    p.tsm_map = 1;
    p.tsm_split_cb = tsm_in_ch / bk;
  }
  finish_launch(L, Wo, Ho, Nimg, Cout, fp32);
  p.b_map = weight_map(Wp, Cout, k * k * Cin, L.cg2 ? L.block_n / 2 : L.block_n, fp32);
  set_epilogue(L, out, Cout, e);
  L.flops = 2.0 * Nimg * Ho * Wo * static_cast<double>(Cout) * k * k * Cin;
  // a stride-2 1x1 conv touches only the pixels it keeps
  const double in_px = (k == 1) ? static_cast<double>(Ho) * Wo : static_cast<double>(H) * W;
  L.bytes = static_cast<double>(es) * (Nimg * in_px * Cin + static_cast<double>(Cout) * k * k * Cin +
                                       static_cast<double>(Nimg) * Ho * Wo * (Cout * (e.residual ? 2.0 : 1.0) +
                                                                              (e.tsm_out ? 2.0 * e.tsm_fold : 0.0)));
  return L;
}

// ResNet stem: 7x7 stride-2 pad-3 conv over the zero-padded stem input (kernels.cuh: pixel (0,0) of the buffer is input
// pixel (-3,-3)).  The filter rows are walked by TMA boxes over *overlapping* windows: the tensor map's W dimension
// has a 2-pixel stride while a window is 8 pixels wide.
//   bf16: the image is stored with ROW PAIRS interleaved per pixel, [N][Hp/2][Wp][2 rows][4 ch], so that a window of
//         8 pixels x 2 rows x 4 channels is 128 contiguous bytes = one K block of 64; the stride-2 row walk becomes a
//         unit walk over row pairs: map (64, Wo [32 B], Hp/2, 1, N), 4 K blocks (K = 256, 147 useful);
//         weights packed [Cout][4 row pairs][8 px][2][4], zero for kh = 7, kw = 7, c = 3.
//   fp32: plain NHWC4 [N][Hp][Wp][4]; a window of 8 pixels is one K block of 32: map (32, Wo [32 B], Hp/2, parity, N),
//         7 K blocks (K = 224); weights packed [Cout][7][8 px][4].
inline ConvGemmLaunch build_stem(const void* in_padded, int Nimg, int Hp, int Wp, int Ho, int Wo, const void* Wp_packed,
                                 int Cout, void* out, bool fp32, const Epilogue& e) {
  ConvGemmLaunch L;
  memset(&L.p, 0, sizeof L.p);
  L.name = "stem";
  VCG_REQUIRE(Hp % 2 == 0, "padded stem height must be even");
  ConvGemmParams& p = L.p;
  pick_patch(Wo, Ho, Nimg, p.bw, p.bh, p.nf);
  const uint32_t box[5] = {static_cast<uint32_t>(block_k(fp32)), static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh), 1,
                           static_cast<uint32_t>(p.nf)};
  if (!fp32) {
    const uint64_t pair_pitch = static_cast<uint64_t>(Wp) * 8 * 2;   // one row pair: Wp pixels x (2 rows x 4 ch) bf16
    const uint64_t img = pair_pitch * (Hp / 2);
    const uint64_t dims[5] = {64, static_cast<uint64_t>(Wo), static_cast<uint64_t>(Hp / 2), 1, static_cast<uint64_t>(Nimg)};
    const uint64_t str[4] = {2 * 16, pair_pitch, img, img};            // W step: 2 pixels x 16 B
    p.a_map[0] = make_tensor_map(in_padded, false, 5, dims, str, box);
    p.n_taps = 4; p.cpt = 1;
    for (int j = 0; j < 4; ++j) {
      TapDesc t{};
      t.dh = static_cast<int16_t>(j);
      p.taps[j] = t;
    }
  } else {
    const uint64_t pitch = static_cast<uint64_t>(Wp) * 4 * 4;
    const uint64_t dims[5] = {32, static_cast<uint64_t>(Wo), static_cast<uint64_t>(Hp / 2), 2, static_cast<uint64_t>(Nimg)};
    const uint64_t str[4] = {2 * 16, 2 * pitch, pitch, pitch * Hp};
    p.a_map[0] = make_tensor_map(in_padded, true, 5, dims, str, box);
    p.n_taps = 7; p.cpt = 1;
    for (int kh = 0; kh < 7; ++kh) {
      TapDesc t{};
      t.dh = static_cast<int16_t>(kh / 2); t.plane = static_cast<int8_t>(kh & 1);
      p.taps[kh] = t;
    }
  }
  for (int i = 1; i < 4; ++i) p.a_map[i] = p.a_map[0];
  finish_launch(L, Wo, Ho, Nimg, Cout, fp32);
  p.b_map = weight_map(Wp_packed, Cout, fp32 ? 7 * 32 : 4 * 64, L.cg2 ? L.block_n / 2 : L.block_n, fp32);
  set_epilogue(L, out, Cout, e);
  L.flops = 2.0 * Nimg * Ho * Wo * static_cast<double>(Cout) * 147;
  L.bytes = static_cast<double>(elem_size(fp32)) * Nimg * (static_cast<double>(Hp) * Wp * 4 + static_cast<double>(Ho) * Wo * Cout);
  return L;
}

// Can conv1 of a TSM bottleneck read its temporally shifted channels straight from the block input?  The two shifted
// groups are `fold` = Cin / shift_div channels each and a K block is 64 channels, so fold must be a multiple of 64
// (Cin >= 512 at shift_div 8), and a tile's frames must be whole clips or divide one.
inline bool conv1_tsm_direct_ok(int W, int H, int Nimg, int Cin, int fold, int T, bool fp32) {
  if (fp32 || fold <= 0 || fold % 64 != 0 || Cin % fold != 0 || Cin / fold > 16 || T <= 0 || Nimg % T != 0) return false;
  int bw, bh, nf;
  pick_patch(W, H, Nimg, bw, bh, nf);
  return nf <= T ? (T % nf == 0) : (nf % T == 0);
}

// conv1 (1x1) of a TSM bottleneck with the temporal shift folded into the TMA coordinates: the input is addressed as
// (C, W, H, t, clip); K blocks of channels [0, fold) come from frame t+1, [fold, 2 fold) from frame t-1, the rest from
// frame t, and frames outside the clip are TMA zero fill — no shifted copy of the activations exists anywhere
// (ops/temporal_shift.py:34-51).  One "tap" per `fold` channels.  bf16.
inline ConvGemmLaunch build_conv1_tsm_direct(const void* x, int Nimg, int T, int H, int W, int Cin, int fold, const void* Wp,
                                             int Cout, void* out, const Epilogue& e, const char* name) {
  ConvGemmLaunch L;
  memset(&L.p, 0, sizeof L.p);
  L.name = name;
  ConvGemmParams& p = L.p;
  pick_patch(W, H, Nimg, p.bw, p.bh, p.nf);
  const int nf_t = std::min(p.nf, T), nf_c = p.nf / nf_t;
  p.a_map[0] = clip_view_map(x, Cin, W, H, T, Nimg / T, T, 64, p.bw, p.bh, nf_t, false, nf_c);
  for (int i = 1; i < 4; ++i) p.a_map[i] = p.a_map[0];
  p.a_clip_T = T;
  p.n_taps = Cin / fold; p.cpt = fold / 64;
  for (int i = 0; i < p.n_taps; ++i) {
    TapDesc t{};
    t.c_off = static_cast<int16_t>(i * fold);
    t.plane = static_cast<int8_t>(i == 0 ? 1 : i == 1 ? -1 : 0);
    p.taps[i] = t;
  }
  finish_launch(L, W, H, Nimg, Cout, false);
  p.b_map = weight_map(Wp, Cout, Cin, L.cg2 ? L.block_n / 2 : L.block_n, false);
  set_epilogue(L, out, Cout, e);
  L.flops = 2.0 * Nimg * H * W * static_cast<double>(Cout) * Cin;
  L.bytes = 2.0 * (static_cast<double>(Nimg) * H * W * (Cin + Cout) + static_cast<double>(Cout) * Cin);
  return L;
}

// layer1.0.conv1 over a per-unique-frame input x0u [U, H, W, Cin] shared by overlapping clips (clip b = frames
// b*clip_stride .. +T-1), with the temporal shift expressed as three temporal taps with channel-masked weights:
//   tap dt=0 -> channels [2f, Cin), dt=+1 -> channels [0, f), dt=-1 -> channels [f, 2f)   (f = Cin / shift_div)
// weights packed [Cout][3][Cin] in that tap order; frames outside the clip are TMA zero fill.  bf16 only.
inline ConvGemmLaunch build_conv1_shared(const void* x0u, int n_clips, int T, int clip_stride, int H, int W, int Cin,
                                         const void* Wp, int Cout, void* out, const Epilogue& e, const char* name) {
  ConvGemmLaunch L;
  memset(&L.p, 0, sizeof L.p);
  L.name = name;
  VCG_REQUIRE(Cin % 64 == 0, "conv input channels must be a multiple of the K block");
  ConvGemmParams& p = L.p;
  const int Nimg = n_clips * T;
  pick_patch(W, H, Nimg, p.bw, p.bh, p.nf);
  VCG_REQUIRE(T % p.nf == 0, "shared-stem conv1: the frames of one tile must belong to one clip");
  p.a_map[0] = clip_view_map(x0u, Cin, W, H, T, n_clips, clip_stride, 64, p.bw, p.bh, p.nf, false);
  for (int i = 1; i < 4; ++i) p.a_map[i] = p.a_map[0];
  p.a_clip_T = T;
  p.n_taps = 3; p.cpt = Cin / 64;
  const int dts[3] = {0, 1, -1};
  for (int i = 0; i < 3; ++i) {
    TapDesc t{};
    t.plane = static_cast<int8_t>(dts[i]);
    p.taps[i] = t;
  }
  finish_launch(L, W, H, Nimg, Cout, false);
  p.b_map = weight_map(Wp, Cout, 3 * Cin, L.cg2 ? L.block_n / 2 : L.block_n, false);
  set_epilogue(L, out, Cout, e);
  L.flops = 2.0 * Nimg * H * W * static_cast<double>(Cout) * Cin;
  L.bytes = 2.0 * (static_cast<double>(Nimg) * H * W * (Cin + Cout) + 3.0 * Cout * Cin);
  return L;
}

void launch_conv_gemm(const ConvGemmLaunch& L, cudaStream_t stream);   // conv_gemm.cu

// ---- fused conv2 (3x3) + conv3 (1x1 + residual) of a bottleneck, planes P = 64 / 128 (conv23.cuh) -------------------
struct Conv23Launch {
  Conv23Params q;
  int halo = 0;        // 1: conv23h_kernel (P = 64), 2: conv23h2_kernel (P = 128): conv2 input as one halo patch per tile
  int grid = 0;
  double flops = 0;
  double bytes = 0;    // algorithmic HBM bytes: conv2 input + residual + output + shifted copy (bf16)
  const char* name = "";
};

// VCG_FUSE23=0 keeps the two separate kernels
int fuse23_policy();   // conv_gemm.cu

inline bool conv23_ok(int P, bool fp32) { return !fp32 && (P == 64 || P == 128) && fuse23_policy() != 0; }

// in: conv2's input [Nimg, H, W, P]; W2 [P][3][3][P], bias2; W3 [4P][P]; e3 = conv3's epilogue (bias, residual, act, TSM)
inline Conv23Launch build_conv23(const void* in, int Nimg, int H, int W, int P, int stride, const void* W2, const float* bias2,
                                 const void* W3, void* out, const Epilogue& e3, const char* name) {
  VCG_REQUIRE(P == 64 || P == 128, "fused conv2+conv3: planes must be 64 or 128");
  Epilogue none;
  ConvGemmLaunch g2 = build_conv(in, Nimg, H, W, P, W2, P, 3, stride, out, false, none, nullptr, 0, name);   // conv2 geometry
  Conv23Launch L;
  memset(&L.q, 0, sizeof L.q);
  L.name = name;
  ConvGemmParams& p = L.q.g;
  p = g2.p;
  const int Cout = 4 * P;
  p.N = Cout;
  p.n_tiles = Cout / 256;
  p.b_map = weight_map(W2, P, 9 * P, P, false);
  p.b_bytes = static_cast<uint32_t>(P) * 128u;
  p.out = out; p.ld_out = Cout;
  p.bias = e3.bias; p.residual = e3.residual; p.ld_res = Cout; p.act = e3.act;
  p.tsm_out = e3.tsm_out; p.tsm_ld = e3.tsm_ld; p.tsm_fold = e3.tsm_fold; p.T = e3.T > 0 ? e3.T : 1;
  if (e3.tsm_out) VCG_REQUIRE(e3.tsm_fold % 32 == 0, "TSM fold must be a multiple of 32 channels");
  p.out_map = c_tile_map(out, Cout, p);
  p.res_map = e3.residual ? c_tile_map(e3.residual, Cout, p) : p.out_map;
  p.res_clip_T = 0;
  if (e3.residual && e3.res_clip_T > 0) {
    VCG_REQUIRE(e3.res_clip_T % p.nf == 0 && p.Nimg % e3.res_clip_T == 0, "clip-view residual: tile frames must stay inside a clip");
    p.res_map = clip_view_map(e3.residual, Cout, p.Wo, p.Ho, e3.res_clip_T, p.Nimg / e3.res_clip_T, e3.res_clip_stride, 64, p.bw,
                              p.bh, p.nf, false);
    p.res_clip_T = e3.res_clip_T;
  }
  L.q.w3_map = weight_map(W3, Cout, P, 256, false);
  L.q.bias2 = bias2;
  L.q.P = P;
  L.q.n2 = Cout / 256;
  // shared memory: 32 KB stages | A2 double buffer (2 x P/64 x 16 KB) | C ring (16 KB slots, >= 4)
  const int a2 = 2 * (P / 64) * kCBytes;
  int stages = P == 64 ? 4 : 3;   // P = 64: 4 stages + 4 C slots; P = 128: 3 stages + 4 C slots (224 KB)
  int cslots = (kC23Budget - a2 - stages * kC23StageBytes) / kCBytes;
  if (const char* v = getenv("VCG_STAGES23")) {
    const int f = atoi(v);
    if (f >= 2 && (kC23Budget - a2 - f * kC23StageBytes) / kCBytes >= 4) { stages = f; cslots = (kC23Budget - a2 - f * kC23StageBytes) / kCBytes; }
  }
  L.q.n_stages = std::min(stages, kMaxStages);
  L.q.n_cslots = std::min(cslots, kMaxCSlots);
  L.q.n_cslots -= L.q.n_cslots % 4;   // four epilogue groups per sub-tile: a ring slot must always come back to the same group
  {
    static const int early = [] { const char* v = getenv("VCG_C23_EARLY"); return v ? atoi(v) : 1; }();
    L.q.early_release = early;
  }
  VCG_REQUIRE(L.q.n_stages >= 2 && L.q.n_cslots >= 4, "fused conv2+conv3: shared-memory split failed");
  {
    static const int pf = [] { const char* v = getenv("VCG_C23T_PF"); return v ? atoi(v) : 0; }();
    L.q.prefetch_tiles = pf;   // conv23t: L2 prefetch distance of the residual boxes (tiles); measured 230 / 239 / 247 us at 0 / 1 / 2
  }
  const long m_tiles = static_cast<long>(p.tiles_w) * p.tiles_h * p.tiles_n;
  L.grid = static_cast<int>(std::min<long>(m_tiles, sm_count()));
  L.flops = 2.0 * Nimg * p.Ho * p.Wo * (static_cast<double>(P) * 9 * P + static_cast<double>(Cout) * P);
  L.bytes = 2.0 * Nimg * (static_cast<double>(H) * W * P + static_cast<double>(p.Ho) * p.Wo * (Cout * (e3.residual ? 2.0 : 1.0) +
                                                                                          (e3.tsm_out ? 2.0 * e3.tsm_fold : 0.0)));
  return L;
}

void launch_conv23(const Conv23Launch& L, cudaStream_t stream);   // conv_gemm.cu

// ---- the same for layer1 (P = 64, stride 1) with the weights and the conv2 halo patch resident in shared memory
//      (conv23h.cuh).  VCG_C23H=0 falls back to conv23_kernel.
int c23h_policy();   // conv_gemm.cu
inline bool conv23h_ok(int P, int stride, int H, int W, bool fp32) {
  return !fp32 && P == 64 && stride == 1 && W % 8 == 0 && W >= 16 && H >= 16 && c23h_policy() != 0;
}

// layer2 (P = 128, stride 1, no TSM scatter): conv23h2.cuh.  OPT-IN (VCG_C23H2=1): measured 3 % SLOWER than
// conv23_kernel<128> on the 28 x 28 maps (8-pixel-wide tile rows cover 28 = 3.5 x 8 pixels: 8 tiles per frame instead of
// 6.125, and with N = 128 MMAs the tensor / shared-memory pipe is already 50 % busy), see DESIGN.md.
int c23h2_policy();   // conv_gemm.cu
inline bool conv23h2_ok(int P, int stride, int H, int W, bool fp32, const void* tsm_out) {
  return !fp32 && P == 128 && stride == 1 && W >= 16 && H >= 16 && tsm_out == nullptr && c23h2_policy() == 1;
}

inline Conv23Launch build_conv23h(const void* in, int Nimg, int H, int W, const void* W2, const float* bias2, const void* W3,
                                  void* out, const Epilogue& e3, const char* name, int P = 64) {
  const int Cout = 4 * P;
  VCG_REQUIRE(P == 64 || P == 128, "halo variant: planes must be 64 or 128");
  Conv23Launch L;
  memset(&L.q, 0, sizeof L.q);
  L.name = name;
  L.halo = P == 64 ? 1 : 2;
  ConvGemmParams& p = L.q.g;
  p.bw = 8; p.bh = 16; p.nf = 1;
  p.Wo = W; p.Ho = H; p.Nimg = Nimg; p.N = Cout;
  p.tiles_w = (W + p.bw - 1) / p.bw;
  p.tiles_h = (H + p.bh - 1) / p.bh;
  p.tiles_n = Nimg;
  p.n_tiles = 1;
  p.n_taps = 9; p.cpt = 1;
  {   // conv2 input [Nimg, H, W, 64] as (C, W, H, 1, N); one box = the (8+2) x (16+2) halo patch of a tile
    const uint64_t img = static_cast<uint64_t>(H) * W * P * 2;
    const uint64_t dims[5] = {static_cast<uint64_t>(P), static_cast<uint64_t>(W), static_cast<uint64_t>(H), 1,
                              static_cast<uint64_t>(Nimg)};
    const uint64_t str[4] = {static_cast<uint64_t>(P) * 2, static_cast<uint64_t>(W) * P * 2, img, img};
    const uint32_t box[5] = {64, kC23hHaloW, kC23hHaloH, 1, 1};
    p.a_map[0] = make_tensor_map(in, false, 5, dims, str, box);
    for (int i = 1; i < 4; ++i) p.a_map[i] = p.a_map[0];
  }
  p.a_bytes = kC23hHaloBytes;
  p.b_map = weight_map(W2, P, 9 * P, P, false);
  p.b_bytes = static_cast<uint32_t>(P) * 128u;
  p.out = out; p.ld_out = Cout;
  p.bias = e3.bias; p.residual = e3.residual; p.ld_res = Cout; p.act = e3.act;
  p.tsm_out = e3.tsm_out; p.tsm_ld = e3.tsm_ld; p.tsm_fold = e3.tsm_fold; p.T = e3.T > 0 ? e3.T : 1;
  if (e3.tsm_out) VCG_REQUIRE(e3.tsm_fold % 32 == 0, "TSM fold must be a multiple of 32 channels");
  p.out_map = c_tile_map(out, Cout, p);
  p.res_map = e3.residual ? c_tile_map(e3.residual, Cout, p) : p.out_map;
  p.res_clip_T = 0;
  if (e3.residual && e3.res_clip_T > 0) {
    VCG_REQUIRE(p.Nimg % e3.res_clip_T == 0, "clip-view residual: whole clips only");
    p.res_map = clip_view_map(e3.residual, Cout, p.Wo, p.Ho, e3.res_clip_T, p.Nimg / e3.res_clip_T, e3.res_clip_stride, 64, p.bw,
                              p.bh, p.nf, false);
    p.res_clip_T = e3.res_clip_T;
  }
  L.q.w3_map = weight_map(W3, Cout, P, P == 64 ? 256 : 128, false);
  L.q.bias2 = bias2;
  L.q.P = P;
  L.q.n2 = Cout / 256;
  L.q.n_stages = kC23hHaloStages;
  L.q.n_cslots = kC23hCSlots;
  if (P == 128) VCG_REQUIRE(e3.tsm_out == nullptr, "halo variant, P = 128: no TSM scatter");
  VCG_REQUIRE(e3.act == ACT_RELU, "halo variant: ReLU epilogue only");
  {
    static const int pf = [] { const char* v = getenv("VCG_C23H_PF"); return v ? atoi(v) : 1; }();
    L.q.prefetch_tiles = pf;
  }
  L.q.magic_tpi = static_cast<uint32_t>((1ull << 32) / static_cast<uint64_t>(p.tiles_w * p.tiles_h)) + 1u;
  L.q.magic_tw = static_cast<uint32_t>((1ull << 32) / static_cast<uint64_t>(p.tiles_w)) + 1u;
  L.q.early_release = 1;
  const long m_tiles = static_cast<long>(p.tiles_w) * p.tiles_h * p.tiles_n;
  L.grid = static_cast<int>(std::min<long>(m_tiles, sm_count()));
  L.flops = 2.0 * Nimg * H * W * (static_cast<double>(P) * 9 * P + static_cast<double>(Cout) * P);
  L.bytes = 2.0 * Nimg * H * W * (P + Cout * (e3.residual ? 2.0 : 1.0) + (e3.tsm_out ? 2.0 * e3.tsm_fold : 0.0));
  return L;
}

}  // namespace vcg
