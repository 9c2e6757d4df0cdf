// Final stage of TwoStream.forward (two_stream.py:88-95, :189): the head over the fused tokens
// [relu(W_v vision_emb[0..T-1]), relu(W_l lang_emb)] and the 2-way softmax, one CTA per clip, all arithmetic in fp32
// (the argmax of the two logits decides the chapter timestamps):
//     head_type mlp   logits = W_h concat(tokens) + b                           two_stream.py:88-93
//     head_type attn  4-head self-attention over the T+1 tokens, token 0 only   two_stream.py:31-48
// The pooler (tanh) and the two bias-free projections (ReLU) run on the tcgen05 GEMM kernel (engine.cu).
#include "kernels.cuh"
#include "launch.cuh"
#include "tensormap.h"
#include <cuda_bf16.h>
#include <algorithm>
#include <type_traits>

namespace vcg {

namespace {

constexpr int kMaxTok = 40;       // T + 1 <= 40

// ---- (3) head on the fused tokens [vision_out[0..T-1], lang_out] + 2-way softmax, one CTA per clip -------------------
template <bool FP32>
__global__ void __launch_bounds__(128) head_final_kernel(const TailParams p) {
  using in_t = typename std::conditional<FP32, float, __nv_bfloat16>::type;
  extern __shared__ float sm[];
  pdl_enter();
  const int H = p.H, T = p.T, tid = threadIdx.x, b = blockIdx.x, ntok = T + 1;
  float* s_tok = sm;                       // [(T+1)][H]
  float* s_k = s_tok + ntok * H;           // attn only: [(T+1)][H]
  float* s_v = s_k + ntok * H;             // attn only
  float* s_q = s_v + ntok * H;             // [H]
  float* s_att = s_q + H;                  // [4][ntok]
  __shared__ float s_red[8];
  const float* vis_out = static_cast<const float*>(p.vis_out);   // fp32 in both precisions (3xTF32 projection)
  const in_t* lang_out = static_cast<const in_t*>(p.lang_out);
  for (int i = tid; i < T * H; i += 128) s_tok[i] = vis_out[static_cast<long>(b) * T * H + i];
  for (int i = tid; i < H; i += 128) s_tok[T * H + i] = static_cast<float>(lang_out[static_cast<long>(b) * H + i]);
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

// nn.Linear(D, 2) + 2-way softmax on fp32 embeddings (single-modality heads): one CTA per clip.
__global__ void __launch_bounds__(256) linear_head2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ b, int D, float* __restrict__ logits,
                                                           float* __restrict__ probs) {
  pdl_enter();
  const float* xr = x + static_cast<long>(blockIdx.x) * D;
  float a0 = 0.f, a1 = 0.f;
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {   // D % 4 == 0 (768, T*2048)
    const float4 v = *reinterpret_cast<const float4*>(xr + i);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + i));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + D + i));
    a0 += v.x * w0.x + v.y * w0.y + v.z * w0.z + v.w * w0.w;
    a1 += v.x * w1.x + v.y * w1.y + v.z * w1.z + v.w * w1.w;
  }
  __shared__ float red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a0; red[1][threadIdx.x >> 5] = a1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l0 = b[0], l1 = b[1];
    for (int i = 0; i < 8; ++i) { l0 += red[0][i]; l1 += red[1][i]; }
    const float m = fmaxf(l0, l1), e0 = expf(l0 - m), e1 = expf(l1 - m), inv = 1.f / (e0 + e1);
    logits[blockIdx.x * 2] = l0; logits[blockIdx.x * 2 + 1] = l1;
    probs[blockIdx.x * 2] = e0 * inv; probs[blockIdx.x * 2 + 1] = e1 * inv;
  }
}

void launch_linear_head2(const float* x, const float* w, const float* b, int B, int D, float* logits, float* probs,
                         cudaStream_t s) {
  if (B == 0) return;
  VCG_REQUIRE(D % 4 == 0, "linear head: feature size must be a multiple of 4");
  launch_pdl(linear_head2_kernel, B, 256, 0, s, x, w, b, D, logits, probs);
}

void launch_head_final(const TailParams& p, int B, bool fp32, cudaStream_t s) {
  if (B == 0) return;
  VCG_REQUIRE(p.H == 128, "ChapterHead hidden size must be 128");
  VCG_REQUIRE(p.T + 1 <= kMaxTok, "clip_frame_num + 1 must be <= 40");
  const int ntok = p.T + 1;
  const size_t smem = sizeof(float) * ((p.head_type == 0 ? 1 : 3) * ntok * p.H + p.H + 4 * ntok);
  static PerDeviceMax configured[2];
  if (configured[fp32].raise(smem)) {
    if (fp32) VCG_CUDA(cudaFuncSetAttribute(head_final_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    else VCG_CUDA(cudaFuncSetAttribute(head_final_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  if (fp32) launch_pdl(head_final_kernel<true>, B, 128, smem, s, p);
  else launch_pdl(head_final_kernel<false>, B, 128, smem, s, p);
  VCG_CUDA(cudaGetLastError());
}

}  // namespace vcg
