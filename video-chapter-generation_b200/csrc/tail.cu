// Fused tail of TwoStream.forward, one CTA per clip, all arithmetic in fp32 (the argmax of the two logits decides
// the chapter timestamps, so nothing here is rounded to bf16):
//   BertPooler        lang_emb = tanh(W_p h[CLS] + b_p)                      modeling_bert.py:456-468
//   ChapterHead       relu(W_l lang_emb), relu(W_v vision_emb[t])            two_stream.py:79-86
//     head_type mlp   logits = W_h concat(vision_out[0..T-1], lang_out) + b  two_stream.py:88-93
//     head_type attn  4-head self-attention over the T+1 tokens, token 0 only two_stream.py:31-48
//   softmax over the two logits                                               two_stream.py:189
// Weights are stored transposed ([in][out]) so that consecutive threads read consecutive addresses; they are
// ~3.5 MB in total and stay L2-resident across the clips of a batch.
#include "kernels.cuh"
#include "tensormap.h"
#include <cuda_bf16.h>
#include <algorithm>

namespace vcg {

namespace {

constexpr int kTailThreads = 256;
constexpr int kMaxTok = 40;   // T + 1 <= 40

template <bool FP32>
__global__ void __launch_bounds__(kTailThreads) tail_kernel(const TailParams p) {
  extern __shared__ float sm[];
  const int H = p.H, T = p.T, tid = threadIdx.x, b = blockIdx.x;
  float* s_h0 = sm;                          // [768]
  float* s_pool = s_h0 + kBertHidden;        // [768]
  float* s_tok = s_pool + kBertHidden;       // [(T+1)][H]   fused tokens: vision_out[0..T-1], lang_out
  float* s_vis = s_tok + (T + 1) * H;        // [T][256]     K-chunk of this clip's vision embeddings
  const int vis_floats = max(T * 256, 2 * (T + 1) * H);   // attn head parks k and v of all tokens here
  float* s_red = s_vis + vis_floats;         // [2*H + 64 + 4*kMaxTok] scratch

  // ---- [CLS] hidden state
  for (int i = tid; i < kBertHidden; i += kTailThreads) {
    if constexpr (FP32) s_h0[i] = static_cast<const float*>(p.hidden)[static_cast<long>(b) * p.L * kBertHidden + i];
    else s_h0[i] = __bfloat162float(static_cast<const __nv_bfloat16*>(p.hidden)[static_cast<long>(b) * p.L * kBertHidden + i]);
  }
  __syncthreads();
  // ---- pooler: thread j owns outputs j, j+256, j+512
  {
    float acc[3] = {p.pool_b[tid], p.pool_b[tid + 256], p.pool_b[tid + 512]};
    for (int k = 0; k < kBertHidden; ++k) {
      const float x = s_h0[k];
      const float* w = p.pool_w_t + static_cast<long>(k) * kBertHidden + tid;
      acc[0] = fmaf(__ldg(w), x, acc[0]);
      acc[1] = fmaf(__ldg(w + 256), x, acc[1]);
      acc[2] = fmaf(__ldg(w + 512), x, acc[2]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float v = tanhf(acc[i]);
      s_pool[tid + i * 256] = v;
      if (p.lang_emb) p.lang_emb[static_cast<long>(b) * kBertHidden + tid + i * 256] = v;
    }
  }
  __syncthreads();
  // ---- lang projection (bias-free) + ReLU: thread = (output j, K half)
  {
    const int j = tid % H, half = tid / H;   // H = 128 -> two halves over 256 threads
    const int halves = kTailThreads / H;
    float acc = 0.f;
    for (int k = half; k < kBertHidden; k += halves) acc = fmaf(__ldg(p.lang_w_t + static_cast<long>(k) * H + j), s_pool[k], acc);
    s_red[tid] = acc;
    __syncthreads();
    if (tid < H) {
      float v = 0.f;
      for (int h = 0; h < halves; ++h) v += s_red[tid + h * H];
      s_tok[T * H + tid] = fmaxf(v, 0.f);
    }
    __syncthreads();
  }
  // ---- vision projection + ReLU for the clip's T frames; K streamed through smem in chunks of 256
  {
    const int j = tid % H, half = tid / H;
    const int halves = kTailThreads / H;
    float acc[kMaxTok];
#pragma unroll
    for (int t = 0; t < kMaxTok; ++t) acc[t] = 0.f;
    const float* vis = p.vision + static_cast<long>(b) * T * kVisionDim;
    for (int k0 = 0; k0 < kVisionDim; k0 += 256) {
      __syncthreads();
      for (int i = tid; i < T * 256; i += kTailThreads) s_vis[i] = vis[static_cast<long>(i / 256) * kVisionDim + k0 + (i % 256)];
      __syncthreads();
      for (int kk = half; kk < 256; kk += halves) {
        const float w = __ldg(p.vis_w_t + static_cast<long>(k0 + kk) * H + j);
#pragma unroll
        for (int t = 0; t < kMaxTok; ++t)
          if (t < T) acc[t] = fmaf(w, s_vis[t * 256 + kk], acc[t]);
      }
    }
    // reduce the K halves through smem (token buffer is [T+1][H])
    for (int h = 0; h < halves; ++h) {
      __syncthreads();
      if (half == h) {
#pragma unroll
        for (int t = 0; t < kMaxTok; ++t)
          if (t < T) {
            if (h == 0) s_tok[t * H + j] = acc[t];
            else s_tok[t * H + j] += acc[t];
          }
      }
    }
    __syncthreads();
    for (int i = tid; i < T * H; i += kTailThreads) s_tok[i] = fmaxf(s_tok[i], 0.f);
    __syncthreads();
  }

  float logit[2] = {0.f, 0.f};
  if (p.head_type == 0) {
    // ---- mlp head: two dot products of length (T+1)*H, block reduction
    const int n = (T + 1) * H;
    float a0 = 0.f, a1 = 0.f;
    for (int i = tid; i < n; i += kTailThreads) {
      const float x = s_tok[i];
      a0 = fmaf(__ldg(p.head_w + i), x, a0);
      a1 = fmaf(__ldg(p.head_w + n + i), x, a1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if ((tid & 31) == 0) { s_red[(tid >> 5) * 2] = a0; s_red[(tid >> 5) * 2 + 1] = a1; }
    __syncthreads();
    if (tid == 0) {
      logit[0] = p.head_b[0]; logit[1] = p.head_b[1];
      for (int w = 0; w < kTailThreads / 32; ++w) { logit[0] += s_red[w * 2]; logit[1] += s_red[w * 2 + 1]; }
    }
  } else {
    // ---- attn head: only query token 0 (the first frame) reaches the output
    const int nh = 4, hd = H / nh, ntok = T + 1;
    float* s_q = s_red;              // [H]
    float* s_att = s_red + H;        // [nh][ntok] (<= 4*40)
    float* s_k = s_vis;              // [ntok][H]
    float* s_v = s_vis + ntok * H;   // [ntok][H]
    // q0, and k/v of every token: thread = (output j, token parity)
    if (tid < H) {
      float a = p.q_b[tid];
      for (int k = 0; k < H; ++k) a = fmaf(__ldg(p.q_w_t + k * H + tid), s_tok[k], a);
      s_q[tid] = a;
    }
    for (int i = tid; i < ntok * H; i += kTailThreads) {
      const int t = i / H, j = i % H;
      float ak = p.k_b[j], av = p.v_b[j];
      for (int k = 0; k < H; ++k) {
        const float x = s_tok[t * H + k];
        ak = fmaf(__ldg(p.k_w_t + k * H + j), x, ak);
        av = fmaf(__ldg(p.v_w_t + k * H + j), x, av);
      }
      s_k[i] = ak;
      s_v[i] = av;
    }
    __syncthreads();
    if (tid < nh * ntok) {
      const int h = tid / ntok, t = tid % ntok;
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(s_q[h * hd + d], s_k[t * H + h * hd + d], a);
      s_att[tid] = a * (1.0f / sqrtf(static_cast<float>(hd)));
    }
    __syncthreads();
    if (tid < nh) {
      float mx = -INFINITY, sum = 0.f;
      for (int t = 0; t < ntok; ++t) mx = fmaxf(mx, s_att[tid * ntok + t]);
      for (int t = 0; t < ntok; ++t) { const float e = expf(s_att[tid * ntok + t] - mx); s_att[tid * ntok + t] = e; sum += e; }
      for (int t = 0; t < ntok; ++t) s_att[tid * ntok + t] /= sum;
    }
    __syncthreads();
    if (tid < H) {
      const int h = tid / hd;
      float y = 0.f;
      for (int t = 0; t < ntok; ++t) y = fmaf(s_att[h * ntok + t], s_v[t * H + tid], y);
      s_q[tid] = y;   // reuse as y0
    }
    __syncthreads();
    if (tid == 0) {
      logit[0] = p.proj_b[0]; logit[1] = p.proj_b[1];
      for (int k = 0; k < H; ++k) { logit[0] = fmaf(p.proj_w[k], s_q[k], logit[0]); logit[1] = fmaf(p.proj_w[H + k], s_q[k], logit[1]); }
    }
  }
  if (tid == 0) {
    p.logits[b * 2] = logit[0];
    p.logits[b * 2 + 1] = logit[1];
    const float mx = fmaxf(logit[0], logit[1]);
    const float e0 = expf(logit[0] - mx), e1 = expf(logit[1] - mx);
    p.probs[b * 2] = e0 / (e0 + e1);
    p.probs[b * 2 + 1] = e1 / (e0 + e1);
  }
}

}  // namespace

void launch_tail(const TailParams& p, int B, bool fp32, cudaStream_t s) {
  if (B == 0) return;
  VCG_REQUIRE(p.H == 128, "ChapterHead hidden size must be 128");
  VCG_REQUIRE(p.T + 1 <= kMaxTok, "clip_frame_num + 1 must be <= 40");
  const int vis_floats = std::max(p.T * 256, 2 * (p.T + 1) * p.H);
  const size_t smem = sizeof(float) * (2 * kBertHidden + (p.T + 1) * p.H + vis_floats + 2 * p.H + 64 + 4 * kMaxTok);
  static size_t configured_a = 0, configured_b = 0;
  if (fp32) {
    if (smem > configured_a) {
      VCG_CUDA(cudaFuncSetAttribute(tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      configured_a = smem;
    }
    tail_kernel<true><<<B, kTailThreads, smem, s>>>(p);
  } else {
    if (smem > configured_b) {
      VCG_CUDA(cudaFuncSetAttribute(tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      configured_b = smem;
    }
    tail_kernel<false><<<B, kTailThreads, smem, s>>>(p);
  }
  VCG_CUDA(cudaGetLastError());
}

}  // namespace vcg
