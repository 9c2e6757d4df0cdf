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
constexpr int kMaxTok = 40;       // T + 1 <= 40
constexpr int kProjFrames = 4;    // frames per CTA in the vision projection

// ---- (1) BertPooler + lang projection, one CTA per clip -------------------------------------------------------
template <bool FP32>
__global__ void __launch_bounds__(kTailThreads) lang_tail_kernel(const TailParams p) {
  __shared__ float s_h0[kBertHidden];
  __shared__ float s_pool[kBertHidden];
  __shared__ float s_red[kTailThreads];
  const int H = p.H, tid = threadIdx.x, b = blockIdx.x;
  for (int i = tid; i < kBertHidden; i += kTailThreads) {
    if constexpr (FP32) s_h0[i] = static_cast<const float*>(p.hidden)[static_cast<long>(b) * p.L * kBertHidden + i];
    else s_h0[i] = __bfloat162float(static_cast<const __nv_bfloat16*>(p.hidden)[static_cast<long>(b) * p.L * kBertHidden + i]);
  }
  __syncthreads();
  {   // pooler: thread j owns outputs j, j+256, j+512; transposed weights -> consecutive threads, consecutive addresses
    float acc[3] = {p.pool_b[tid], p.pool_b[tid + 256], p.pool_b[tid + 512]};
#pragma unroll 4
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
  {   // lang projection (bias-free) + ReLU: thread = (output j, K half)
    const int j = tid % H, half = tid / H, halves = kTailThreads / H;
    float acc = 0.f;
    for (int k = half; k < kBertHidden; k += halves) acc = fmaf(__ldg(p.lang_w_t + static_cast<long>(k) * H + j), s_pool[k], acc);
    s_red[tid] = acc;
    __syncthreads();
    if (tid < H) {
      float v = 0.f;
      for (int h = 0; h < halves; ++h) v += s_red[tid + h * H];
      p.lang_out[static_cast<long>(b) * H + tid] = fmaxf(v, 0.f);
    }
  }
}

// ---- (2) vision projection + ReLU: kProjFrames frames per CTA, thread = (output j, K half) ------------------------
__global__ void __launch_bounds__(kTailThreads) vision_proj_kernel(const float* __restrict__ vision,
                                                                  const float* __restrict__ vis_w_t,
                                                                  float* __restrict__ vis_out, int n_frames, int H) {
  extern __shared__ float s_x[];   // [kProjFrames][2048]
  __shared__ float s_part[kProjFrames][kTailThreads];
  const int tid = threadIdx.x, f0 = blockIdx.x * kProjFrames;
  const int nf = min(kProjFrames, n_frames - f0);
  for (int i = tid; i < kProjFrames * kVisionDim / 4; i += kTailThreads) {
    const int f = i / (kVisionDim / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < nf) v = __ldg(reinterpret_cast<const float4*>(vision + static_cast<long>(f0) * kVisionDim) + i);
    reinterpret_cast<float4*>(s_x)[i] = v;
  }
  __syncthreads();
  const int j = tid % H, half = tid / H, halves = kTailThreads / H;
  float acc[kProjFrames];
#pragma unroll
  for (int f = 0; f < kProjFrames; ++f) acc[f] = 0.f;
#pragma unroll 4
  for (int k = half; k < kVisionDim; k += halves) {
    const float w = __ldg(vis_w_t + static_cast<long>(k) * H + j);
#pragma unroll
    for (int f = 0; f < kProjFrames; ++f) acc[f] = fmaf(w, s_x[f * kVisionDim + k], acc[f]);
  }
#pragma unroll
  for (int f = 0; f < kProjFrames; ++f) s_part[f][tid] = acc[f];
  __syncthreads();
  for (int i = tid; i < kProjFrames * H; i += kTailThreads) {
    const int f = i / H, jj = i % H;
    if (f < nf) {
      float v = 0.f;
      for (int h = 0; h < halves; ++h) v += s_part[f][jj + h * H];
      vis_out[static_cast<long>(f0 + f) * H + jj] = fmaxf(v, 0.f);
    }
  }
}

// ---- (3) head on the fused tokens [vision_out[0..T-1], lang_out] + 2-way softmax, one CTA per clip -------------------
__global__ void __launch_bounds__(128) head_final_kernel(const TailParams p) {
  extern __shared__ float sm[];
  const int H = p.H, T = p.T, tid = threadIdx.x, b = blockIdx.x, ntok = T + 1;
  float* s_tok = sm;                       // [(T+1)][H]
  float* s_k = s_tok + ntok * H;           // attn only: [(T+1)][H]
  float* s_v = s_k + ntok * H;             // attn only
  float* s_q = s_v + ntok * H;             // [H]
  float* s_att = s_q + H;                  // [4][ntok]
  __shared__ float s_red[8];
  for (int i = tid; i < T * H; i += 128) s_tok[i] = p.vis_out[static_cast<long>(b) * T * H + i];
  for (int i = tid; i < H; i += 128) s_tok[T * H + i] = p.lang_out[static_cast<long>(b) * H + i];
  __syncthreads();
  float logit[2] = {0.f, 0.f};
  if (p.head_type == 0) {
    const int n = ntok * H;
    float a0 = 0.f, a1 = 0.f;
    for (int i = tid; i < n; i += 128) {
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
      for (int w = 0; w < 4; ++w) { logit[0] += s_red[w * 2]; logit[1] += s_red[w * 2 + 1]; }
    }
  } else {
    // attn head: only query token 0 (the first frame) reaches the output (two_stream.py:46)
    const int nh = 4, hd = H / nh;
    if (tid < H) {
      float a = p.q_b[tid];
      for (int k = 0; k < H; ++k) a = fmaf(__ldg(p.q_w_t + k * H + tid), s_tok[k], a);
      s_q[tid] = a;
    }
    for (int i = tid; i < ntok * H; i += 128) {
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
    for (int i = tid; i < nh * ntok; i += 128) {
      const int h = i / ntok, t = i % ntok;
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(s_q[h * hd + d], s_k[t * H + h * hd + d], a);
      s_att[i] = a * (1.0f / sqrtf(static_cast<float>(hd)));
    }
    __syncthreads();
    if (tid < nh) {
      float mx = -INFINITY, sum = 0.f;
      for (int t = 0; t < ntok; ++t) mx = fmaxf(mx, s_att[tid * ntok + t]);
      for (int t = 0; t < ntok; ++t) { const float e = expf(s_att[tid * ntok + t] - mx); s_att[tid * ntok + t] = e; sum += e; }
      for (int t = 0; t < ntok; ++t) s_att[tid * ntok + t] /= sum;
    }
    __syncthreads();
    float y = 0.f;
    if (tid < H) {
      const int h = tid / hd;
      for (int t = 0; t < ntok; ++t) y = fmaf(s_att[h * ntok + t], s_v[t * H + tid], y);
    }
    __syncthreads();
    if (tid < H) s_q[tid] = y;   // reuse as y0
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

void launch_lang_tail(const TailParams& p, int B, bool fp32, cudaStream_t s) {
  if (B == 0) return;
  VCG_REQUIRE(p.H == 128, "ChapterHead hidden size must be 128");
  if (fp32) lang_tail_kernel<true><<<B, kTailThreads, 0, s>>>(p);
  else lang_tail_kernel<false><<<B, kTailThreads, 0, s>>>(p);
  VCG_CUDA(cudaGetLastError());
}

void launch_vision_proj(const float* vision, const float* vis_w_t, float* vis_out, int n_frames, int H, cudaStream_t s) {
  if (n_frames == 0) return;
  VCG_REQUIRE(H == 128, "ChapterHead hidden size must be 128");
  const size_t smem = sizeof(float) * kProjFrames * kVisionDim;
  static bool configured = false;
  if (!configured) {
    VCG_CUDA(cudaFuncSetAttribute(vision_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  vision_proj_kernel<<<(n_frames + kProjFrames - 1) / kProjFrames, kTailThreads, smem, s>>>(vision, vis_w_t, vis_out, n_frames, H);
  VCG_CUDA(cudaGetLastError());
}

void launch_head_final(const TailParams& p, int B, cudaStream_t s) {
  if (B == 0) return;
  VCG_REQUIRE(p.H == 128, "ChapterHead hidden size must be 128");
  VCG_REQUIRE(p.T + 1 <= kMaxTok, "clip_frame_num + 1 must be <= 40");
  const int ntok = p.T + 1;
  const size_t smem = sizeof(float) * ((p.head_type == 0 ? 1 : 3) * ntok * p.H + p.H + 4 * ntok);
  static size_t configured = 0;
  if (smem > configured) {
    VCG_CUDA(cudaFuncSetAttribute(head_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  head_final_kernel<<<B, 128, smem, s>>>(p);
  VCG_CUDA(cudaGetLastError());
}

}  // namespace vcg
