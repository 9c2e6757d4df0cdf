// Post-backbone part of the reference's WINDOW ("update") model, model/fusion/two_stream_window.py +
// model/fusion/stacked_window_self_attention.py.  Everything here is a few MFLOP per clip on
// [rows, <= 2176] fp32 vectors (the backbones are > 99.9 % of the work and run in the tcgen05 kernels), so the kernels are
// plain fp32 SIMT, one CTA per row / batch item, weights read through L2:
//   mlp_chain_kernel        a short program of Linear / LayerNorm / ReLU / GELU steps over one row (the per-position
//                           projection heads, the "mlp" fusion head): ChapterHead.forward, two_stream_window.py:252-290
//   cross_attention_kernel  CrossAttention.forward, two_stream_window.py:53-88 (head_type "cross_attn", the default of
//                           test_video_segment_update.py:43)
//   window_stack_kernel     StackedVideoChapterAttention.forward, stacked_window_self_attention.py:203-223: six pre-LN
//                           blocks over the 2w+1 clip tokens, final LayerNorm, middle token, classifier, softmax
#include "../../include/vcg.h"
#include "kernels.cuh"
#include "launch.cuh"

namespace vcg {

namespace {

constexpr int kMaxSaved = 1024;  // widest row a SAVE step keeps
constexpr int kMaxDim = 4352;   // widest row of any step (2 * T*128 with T <= 17 for the "multiplication" head)

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                      // red may still be read from a previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// y[n] = sum_k x[k] W[n,k] + b[n]; x, y in shared memory, one warp per output neuron (coalesced weight rows)
__device__ __forceinline__ void row_linear(const float* x, int K, const float* __restrict__ W, const float* __restrict__ b,
                                           float* y, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int n = warp; n < N; n += nw) {
    const float* w = W + static_cast<long>(n) * K;
    float acc = 0.f;
    if ((K & 3) == 0) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
        acc = fmaf(x[k], wv.x, acc); acc = fmaf(x[k + 1], wv.y, acc);
        acc = fmaf(x[k + 2], wv.z, acc); acc = fmaf(x[k + 3], wv.w, acc);
      }
    } else {
      for (int k = lane; k < K; k += 32) acc = fmaf(x[k], __ldg(w + k), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[n] = acc + (b ? b[n] : 0.f);
  }
  __syncthreads();
}

// Several rows against the same weights: y[r][n] = b[n] + sum_k x[r][k] W[n,k] for r < R, rows in shared memory (16-byte
// aligned, strides xs / ys floats, K % 4 == 0).  One thread per output neuron: it walks its weight row once (float4 loads,
// the lines stay in L1 between the 8 consecutive k steps that share them) and feeds up to 8 row accumulators from
// broadcast shared-memory reads — the T frame vectors / W window tokens of one item cost one pass over the weights instead
// of one latency-bound warp-per-neuron pass per row.
__device__ __forceinline__ void rows_linear(const float* x, int xs, int R, int K, const float* __restrict__ W,
                                            const float* __restrict__ b, float* y, int ys, int N) {
  for (int r0 = 0; r0 < R; r0 += 8) {
    const int rc = min(8, R - r0);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      const float4* w4 = reinterpret_cast<const float4*>(W + static_cast<long>(n) * K);
      for (int k4 = 0; k4 < (K >> 2); ++k4) {
        const float4 w = __ldg(w4 + k4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < rc) {
            const float4 xv = *reinterpret_cast<const float4*>(x + (r0 + i) * xs + (k4 << 2));
            acc[i] = fmaf(xv.x, w.x, acc[i]); acc[i] = fmaf(xv.y, w.y, acc[i]);
            acc[i] = fmaf(xv.z, w.z, acc[i]); acc[i] = fmaf(xv.w, w.w, acc[i]);
          }
        }
      }
      const float bb = b ? b[n] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < rc) y[(r0 + i) * ys + n] = acc[i] + bb;
    }
  }
  __syncthreads();
}

// nn.LayerNorm over a row in shared memory (two-pass statistics, in place)
__device__ __forceinline__ void row_layernorm(float* x, int N, const float* __restrict__ g, const float* __restrict__ b,
                                              float eps, float* red) {
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += x[i];
  const float mean = block_sum(s, red) / N;
  float q = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) { const float d = x[i] - mean; q += d * d; }
  const float rstd = rsqrtf(block_sum(q, red) / N + eps);
  for (int i = threadIdx.x; i < N; i += blockDim.x) x[i] = (x[i] - mean) * rstd * g[i] + b[i];
  __syncthreads();
}

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

struct MlpProgram {
  vcg_mlp_op ops[16];
  int n_ops;
};

__global__ void __launch_bounds__(256) mlp_chain_kernel(const float* __restrict__ x0, int dim0, long stride0,
                                                        const float* __restrict__ x1, int dim1, long stride1,
                                                        const __grid_constant__ MlpProgram prog, float* __restrict__ out,
                                                        long out_stride) {
  pdl_enter();
  __shared__ float buf[2][kMaxDim];
  __shared__ float saved[kMaxSaved];
  __shared__ float red[8];
  const long r = blockIdx.x;
  for (int i = threadIdx.x; i < dim0; i += blockDim.x) buf[0][i] = x0[r * stride0 + i];
  for (int i = threadIdx.x; i < dim1; i += blockDim.x) buf[0][dim0 + i] = x1[r * stride1 + i];
  __syncthreads();
  int cur = 0, dim = dim0 + dim1;
  for (int s = 0; s < prog.n_ops; ++s) {
    const vcg_mlp_op op = prog.ops[s];
    if (op.type == VCG_MLP_LINEAR) {
      row_linear(buf[cur], op.in_dim, static_cast<const float*>(op.w), static_cast<const float*>(op.b), buf[cur ^ 1], op.out_dim);
      cur ^= 1;
      dim = op.out_dim;
    } else if (op.type == VCG_MLP_LAYERNORM) {
      row_layernorm(buf[cur], dim, static_cast<const float*>(op.w), static_cast<const float*>(op.b), op.eps, red);
    } else if (op.type == VCG_MLP_MULHALVES) {
      dim >>= 1;
      for (int i = threadIdx.x; i < dim; i += blockDim.x) buf[cur][i] *= buf[cur][dim + i];
      __syncthreads();
    } else if (op.type == VCG_MLP_MEANGROUPS) {
      const int g = dim / op.out_dim;
      for (int i = threadIdx.x; i < op.out_dim; i += blockDim.x) {
        float a = 0.f;
        for (int j = 0; j < g; ++j) a += buf[cur][j * op.out_dim + i];
        buf[cur ^ 1][i] = a / static_cast<float>(g);
      }
      cur ^= 1;
      dim = op.out_dim;
      __syncthreads();
    } else if (op.type == VCG_MLP_SOFTMAX) {
      if (threadIdx.x == 0) {          // rows here are a handful of class logits
        float m = -INFINITY, sum = 0.f;
        for (int i = 0; i < dim; ++i) m = fmaxf(m, buf[cur][i]);
        for (int i = 0; i < dim; ++i) { buf[cur][i] = expf(buf[cur][i] - m); sum += buf[cur][i]; }
        for (int i = 0; i < dim; ++i) buf[cur][i] /= sum;
      }
      __syncthreads();
    } else if (op.type == VCG_MLP_SAVE) {
      for (int i = threadIdx.x; i < dim; i += blockDim.x) saved[i] = buf[cur][i];
      __syncthreads();
    } else if (op.type == VCG_MLP_ADDSAVED) {
      for (int i = threadIdx.x; i < dim; i += blockDim.x) buf[cur][i] += saved[i];
      __syncthreads();
    } else {
      for (int i = threadIdx.x; i < dim; i += blockDim.x)
        buf[cur][i] = op.type == VCG_MLP_RELU ? fmaxf(buf[cur][i], 0.f) : gelu_erf(buf[cur][i]);
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < dim; i += blockDim.x) out[r * out_stride + i] = buf[cur][i];
}

// CrossAttention.forward (two_stream_window.py:53-88), hidden size H = 128, nh heads: the language vector queries the T
// frame vectors of its clip.  One CTA per clip.
__global__ void __launch_bounds__(128) cross_attention_kernel(const __grid_constant__ vcg_cross_attn_params p,
                                                              const float* __restrict__ lang, const float* __restrict__ vision,
                                                              int T, float* __restrict__ out) {
  pdl_enter();
  constexpr int H = 128, kMaxT = 40;
  extern __shared__ __align__(16) float dyn[];        // sv | sk | sval, T rows of H each
  float (*sv)[H] = reinterpret_cast<float (*)[H]>(dyn);
  float (*sk)[H] = sv + T;
  float (*sval)[H] = sk + T;
  __shared__ float sl[H], sq[H], sctx[H], sp[16][kMaxT], red[8];
  const long b = blockIdx.x;
  const int nh = p.num_heads, hd = H / nh;
  for (int i = threadIdx.x; i < H; i += blockDim.x) sl[i] = lang[b * H + i];
  for (int i = threadIdx.x; i < T * H; i += blockDim.x) sv[i / H][i % H] = vision[b * T * H + i];
  __syncthreads();
  row_layernorm(sl, H, p.lang_norm_w, p.lang_norm_b, 1e-5f, red);
  for (int t = 0; t < T; ++t) {
    row_layernorm(sv[t], H, p.vision_norm_w, p.vision_norm_b, 1e-5f, red);
    const float pos = static_cast<float>(t) / static_cast<float>(T - 1);   // get_relative_positions: t / (T - 1)
    for (int i = threadIdx.x; i < H; i += blockDim.x) sv[t][i] += pos * p.pos_w[i] + p.pos_b[i];   // Linear(1, H)
    __syncthreads();
  }
  row_linear(sl, H, p.q_w, p.q_b, sq, H);
  rows_linear(&sv[0][0], H, T, H, p.k_w, p.k_b, &sk[0][0], H, H);
  rows_linear(&sv[0][0], H, T, H, p.v_w, p.v_b, &sval[0][0], H, H);
  const float scale = rsqrtf(static_cast<float>(hd));
  for (int i = threadIdx.x; i < nh * T; i += blockDim.x) {
    const int h = i / T, t = i % T;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(sq[h * hd + d], sk[t][h * hd + d], s);
    sp[h][t] = s * scale;
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < nh) {
    const int h = threadIdx.x;
    float m = -INFINITY, sum = 0.f;
    for (int t = 0; t < T; ++t) m = fmaxf(m, sp[h][t]);
    for (int t = 0; t < T; ++t) { sp[h][t] = expf(sp[h][t] - m); sum += sp[h][t]; }
    for (int t = 0; t < T; ++t) sp[h][t] /= sum;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const int h = i / hd;
    float c = 0.f;
    for (int t = 0; t < T; ++t) c = fmaf(sp[h][t], sval[t][i], c);
    sctx[i] = c;
  }
  __syncthreads();
  row_linear(sctx, H, p.o_w, p.o_b, sq, H);
  for (int i = threadIdx.x; i < H; i += blockDim.x) out[b * H + i] = sq[i];
}

// head_type "self_attn": SelfAttention.forward (two_stream_window.py:114-131) over the T frame vectors + the language
// vector of one clip; only token 0's output row is used (:129), so only its query is formed.  One CTA per clip.
__global__ void __launch_bounds__(128) self_attention_first_kernel(const __grid_constant__ vcg_self_attn_params p,
                                                                   const float* __restrict__ vision,
                                                                   const float* __restrict__ lang, int T,
                                                                   float* __restrict__ out) {
  pdl_enter();
  constexpr int H = 128, kMaxN = 41;
  extern __shared__ __align__(16) float dyn[];        // sx | sk | sval, N = T + 1 rows of H each
  const int N = T + 1;
  float (*sx)[H] = reinterpret_cast<float (*)[H]>(dyn);
  float (*sk)[H] = sx + N;
  float (*sval)[H] = sk + N;
  __shared__ float sq[H], sctx[H], sp[16][kMaxN];
  const long b = blockIdx.x;
  const int nh = p.num_heads, hd = H / nh;
  for (int i = threadIdx.x; i < T * H; i += blockDim.x) sx[i / H][i % H] = vision[b * T * H + i];
  for (int i = threadIdx.x; i < H; i += blockDim.x) sx[T][i] = lang[b * H + i];
  __syncthreads();
  row_linear(sx[0], H, p.q_w, p.q_b, sq, H);
  rows_linear(&sx[0][0], H, N, H, p.k_w, p.k_b, &sk[0][0], H, H);
  rows_linear(&sx[0][0], H, N, H, p.v_w, p.v_b, &sval[0][0], H, H);
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  for (int i = threadIdx.x; i < nh * N; i += blockDim.x) {
    const int h = i / N, t = i % N;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(sq[h * hd + d], sk[t][h * hd + d], s);
    sp[h][t] = s * scale;
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < nh) {
    const int h = threadIdx.x;
    float m = -INFINITY, sum = 0.f;
    for (int t = 0; t < N; ++t) m = fmaxf(m, sp[h][t]);
    for (int t = 0; t < N; ++t) { sp[h][t] = expf(sp[h][t] - m); sum += sp[h][t]; }
    for (int t = 0; t < N; ++t) sp[h][t] /= sum;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const int h = i / hd;
    float c = 0.f;
    for (int t = 0; t < N; ++t) c = fmaf(sp[h][t], sval[t][i], c);
    sctx[i] = c;
  }
  __syncthreads();
  row_linear(sctx, H, p.o_w, p.o_b, sq, H);
  for (int i = threadIdx.x; i < H; i += blockDim.x) out[b * H + i] = sq[i];
}

// nn.Bilinear's contraction with x1 after the GEMM over x2: out[r,o] = bias[o] + sum_i x1[r,i] y[r, o*in1 + i].
// One warp per (row, output): coalesced reads of y.
__global__ void __launch_bounds__(256) bilinear_contract_kernel(const float* __restrict__ y, const float* __restrict__ x1,
                                                                const float* __restrict__ bias, long n_out, int in1,
                                                                int out_features, float* __restrict__ out) {
  pdl_enter();
  const long w = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_out) return;
  const long r = w / out_features;
  const int o = static_cast<int>(w % out_features);
  const float* yr = y + (r * out_features + o) * in1;
  const float* xr = x1 + r * in1;
  float acc = 0.f;
  for (int i = lane; i < in1; i += 32) acc = fmaf(xr[i], yr[i], acc);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) out[w] = acc + (bias ? bias[o] : 0.f);
}

// Window attention with the centre clip as the only query (see vcg_center_attn_params).  One CTA per batch item.
__global__ void __launch_bounds__(128) center_attention_kernel(const __grid_constant__ vcg_center_attn_params p,
                                                               const float* __restrict__ x, int W, float* __restrict__ out) {
  pdl_enter();
  constexpr int H = 128, kMaxW = 9;
  __shared__ __align__(16) float sx[kMaxW][H], sk[kMaxW][H], sval[kMaxW][H];
  __shared__ float spos[H], sq[H], sctx[H], sp[16][kMaxW], red[8];
  const long b = blockIdx.x;
  const int nh = p.num_heads, hd = H / nh, mid = W / 2;
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) sx[i / H][i % H] = x[b * W * H + i];
  __syncthreads();
  for (int t = 0; t < W; ++t) {
    if (p.pre_norm_w) row_layernorm(sx[t], H, p.pre_norm_w, p.pre_norm_b, 1e-5f, red);
    const float pos = static_cast<float>(t - mid) / (static_cast<float>(mid) + 1e-6f);
    for (int i = threadIdx.x; i < H; i += blockDim.x) spos[i] = pos * p.pos_w[i] + p.pos_b[i];     // Linear(1, H)
    __syncthreads();
    row_layernorm(spos, H, p.pos_norm_w, p.pos_norm_b, 1e-5f, red);
    for (int i = threadIdx.x; i < H; i += blockDim.x) sx[t][i] += spos[i];
    __syncthreads();
    if (p.post_norm_w) row_layernorm(sx[t], H, p.post_norm_w, p.post_norm_b, 1e-5f, red);
  }
  rows_linear(&sx[0][0], H, W, H, p.k_w, p.k_b, &sk[0][0], H, H);
  rows_linear(&sx[0][0], H, W, H, p.v_w, p.v_b, &sval[0][0], H, H);
  row_linear(sx[mid], H, p.q_w, p.q_b, sq, H);
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  for (int i = threadIdx.x; i < nh * W; i += blockDim.x) {
    const int h = i / W, t = i % W;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(sq[h * hd + d], sk[t][h * hd + d], s);
    sp[h][t] = s * scale + p.pos_bias[h * p.bias_head_stride + p.bias_offset + t];
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < nh) {
    const int h = threadIdx.x;
    float m = -INFINITY, sum = 0.f;
    for (int t = 0; t < W; ++t) m = fmaxf(m, sp[h][t]);
    for (int t = 0; t < W; ++t) { sp[h][t] = expf(sp[h][t] - m); sum += sp[h][t]; }
    for (int t = 0; t < W; ++t) sp[h][t] /= sum;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const int h = i / hd;
    float c = 0.f;
    for (int t = 0; t < W; ++t) c = fmaf(sp[h][t], sval[t][i], c);
    sctx[i] = c;
  }
  __syncthreads();
  const float* res = sctx;
  if (p.o_w) {
    row_linear(sctx, H, p.o_w, p.o_b, sq, H);
    res = sq;
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x)
    out[b * H + i] = res[i] + (p.add_residual ? x[(b * W + mid) * H + i] : 0.f);
}

// StackedVideoChapterAttention.forward: x [B, W, 128] -> logits, probs [B, 2].  One CTA per batch item; the W tokens go
// through every Linear together (rows_linear).  Dynamic shared memory: h | n | q | k | v [W][H], f1 | f2 [W][4H],
// sc [NH][W][W].
__global__ void __launch_bounds__(256) window_stack_kernel(const __grid_constant__ vcg_window_stack_params p,
                                                           const float* __restrict__ x, int W, float* __restrict__ logits,
                                                           float* __restrict__ probs) {
  pdl_enter();
  constexpr int H = 128, NH = 16, HD = 8;
  extern __shared__ __align__(16) float dyn[];
  float (*h)[H] = reinterpret_cast<float (*)[H]>(dyn);
  float (*n)[H] = h + W;
  float (*q)[H] = n + W;
  float (*k)[H] = q + W;
  float (*v)[H] = k + W;
  float* f1 = &v[W][0];                   // [W][4H]
  float* f2 = f1 + W * 4 * H;             // [W][4H]
  float* sc = f2 + W * 4 * H;             // [NH][W][W]
  __shared__ float red[8];
  const long b = blockIdx.x;
  for (int i = threadIdx.x; i < W * H; i += blockDim.x) h[i / H][i % H] = x[b * W * H + i];
  __syncthreads();
  const int mid = W / 2;
  for (int l = 0; l < p.num_layers; ++l) {
    const vcg_window_layer& L = p.layers[l];
    // ---- attention: pre-LN, + Linear(1,H)((t - mid) / (mid + 1e-6)), 16 heads x 8, + window_pos_bias[head, key]
    for (int t = 0; t < W; ++t) {
      for (int i = threadIdx.x; i < H; i += blockDim.x) n[t][i] = h[t][i];
      __syncthreads();
      row_layernorm(n[t], H, L.attn_norm_w, L.attn_norm_b, 1e-5f, red);
      const float pos = static_cast<float>(t - mid) / (static_cast<float>(mid) + 1e-6f);
      for (int i = threadIdx.x; i < H; i += blockDim.x) n[t][i] += pos * L.pos_w[i] + L.pos_b[i];
      __syncthreads();
    }
    rows_linear(&n[0][0], H, W, H, L.q_w, L.q_b, &q[0][0], H, H);
    rows_linear(&n[0][0], H, W, H, L.k_w, L.k_b, &k[0][0], H, H);
    rows_linear(&n[0][0], H, W, H, L.v_w, L.v_b, &v[0][0], H, H);
    for (int i = threadIdx.x; i < NH * W * W; i += blockDim.x) {
      const int hh = i / (W * W), tq = (i / W) % W, tk = i % W;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(q[tq][hh * HD + d], k[tk][hh * HD + d], s);
      sc[i] = s * 0.35355339059327373f + L.pos_bias[hh * p.pos_bias_stride + tk];   // / sqrt(8)
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NH * W; i += blockDim.x) {
      float* row = sc + i * W;
      float m = -INFINITY, sum = 0.f;
      for (int t = 0; t < W; ++t) m = fmaxf(m, row[t]);
      for (int t = 0; t < W; ++t) { row[t] = expf(row[t] - m); sum += row[t]; }
      for (int t = 0; t < W; ++t) row[t] /= sum;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < W * H; i += blockDim.x) {     // context, written over n
      const int t = i / H, c = i % H, hh = c / HD;
      float a = 0.f;
      for (int tk = 0; tk < W; ++tk) a = fmaf(sc[(hh * W + t) * W + tk], v[tk][c], a);
      n[t][c] = a;
    }
    __syncthreads();
    rows_linear(&n[0][0], H, W, H, L.o_w, L.o_b, &q[0][0], H, H);
    for (int i = threadIdx.x; i < W * H; i += blockDim.x) h[i / H][i % H] += q[i / H][i % H];     // residual
    __syncthreads();
    // ---- FFN: pre-LN, 128 -> 256 -> 512 -> 256 -> 128 with GELU between, residual
    for (int t = 0; t < W; ++t) {
      for (int i = threadIdx.x; i < H; i += blockDim.x) n[t][i] = h[t][i];
      __syncthreads();
      row_layernorm(n[t], H, L.ffn_norm_w, L.ffn_norm_b, 1e-5f, red);
    }
    rows_linear(&n[0][0], H, W, H, L.f0_w, L.f0_b, f1, 4 * H, 2 * H);
    for (int i = threadIdx.x; i < W * 2 * H; i += blockDim.x) { float& e = f1[(i / (2 * H)) * 4 * H + i % (2 * H)]; e = gelu_erf(e); }
    __syncthreads();
    rows_linear(f1, 4 * H, W, 2 * H, L.f1_w, L.f1_b, f2, 4 * H, 4 * H);
    for (int i = threadIdx.x; i < W * 4 * H; i += blockDim.x) f2[i] = gelu_erf(f2[i]);
    __syncthreads();
    rows_linear(f2, 4 * H, W, 4 * H, L.f2_w, L.f2_b, f1, 4 * H, 2 * H);
    for (int i = threadIdx.x; i < W * 2 * H; i += blockDim.x) { float& e = f1[(i / (2 * H)) * 4 * H + i % (2 * H)]; e = gelu_erf(e); }
    __syncthreads();
    rows_linear(f1, 4 * H, W, 2 * H, L.f3_w, L.f3_b, f2, 4 * H, H);
    for (int i = threadIdx.x; i < W * H; i += blockDim.x) h[i / H][i % H] += f2[(i / H) * 4 * H + i % H];
    __syncthreads();
  }
  // ---- final LayerNorm, middle (target) clip, classifier: 4 x (Linear, LayerNorm, GELU), Linear(32, 2), softmax
  float* t0 = h[mid];
  row_layernorm(t0, H, p.final_norm_w, p.final_norm_b, 1e-5f, red);
  const int dims[5] = {H, H, H, H / 2, H / 4};
  float* a = t0;
  float* o = f1;
  for (int i = 0; i < 4; ++i) {
    row_linear(a, dims[i], p.cls_w[i], p.cls_b[i], o, dims[i + 1]);
    row_layernorm(o, dims[i + 1], p.cls_norm_w[i], p.cls_norm_b[i], 1e-5f, red);
    for (int j = threadIdx.x; j < dims[i + 1]; j += blockDim.x) o[j] = gelu_erf(o[j]);
    __syncthreads();
    float* tmp = a; a = o; o = (tmp == t0) ? f2 : tmp;
  }
  row_linear(a, H / 4, p.cls_w[4], p.cls_b[4], o, 2);
  if (threadIdx.x == 0) {
    const float l0 = o[0], l1 = o[1], m = fmaxf(l0, l1), e0 = expf(l0 - m), e1 = expf(l1 - m);
    logits[b * 2] = l0; logits[b * 2 + 1] = l1;
    probs[b * 2] = e0 / (e0 + e1); probs[b * 2 + 1] = e1 / (e0 + e1);
  }
}

}  // namespace

void launch_mlp_chain(const float* x0, int dim0, long stride0, const float* x1, int dim1, long stride1, int rows,
                      const vcg_mlp_op* ops, int n_ops, float* out, long out_stride, cudaStream_t s) {
  if (rows == 0) return;
  VCG_REQUIRE(n_ops >= 1 && n_ops <= 16, "mlp chain: 1..16 steps");
  VCG_REQUIRE(dim0 + dim1 <= kMaxDim, "mlp chain: input row too wide");
  MlpProgram prog{};
  prog.n_ops = n_ops;
  int dim = dim0 + dim1, saved_dim = -1;
  for (int i = 0; i < n_ops; ++i) {
    prog.ops[i] = ops[i];
    if (ops[i].type == VCG_MLP_LINEAR) {
      VCG_REQUIRE(ops[i].in_dim == dim && ops[i].out_dim >= 1 && ops[i].out_dim <= kMaxDim && ops[i].w, "mlp chain: bad linear step");
      dim = ops[i].out_dim;
    } else if (ops[i].type == VCG_MLP_LAYERNORM) {
      VCG_REQUIRE(ops[i].w && ops[i].b, "mlp chain: LayerNorm needs weight and bias");
    } else if (ops[i].type == VCG_MLP_MULHALVES) {
      VCG_REQUIRE((dim & 1) == 0, "mlp chain: MULHALVES needs an even row");
      dim >>= 1;
    } else if (ops[i].type == VCG_MLP_MEANGROUPS) {
      VCG_REQUIRE(ops[i].out_dim >= 1 && dim % ops[i].out_dim == 0, "mlp chain: MEANGROUPS needs a row of g * out_dim");
      dim = ops[i].out_dim;
    } else if (ops[i].type == VCG_MLP_SOFTMAX) {
      VCG_REQUIRE(dim <= 64, "mlp chain: SOFTMAX is for short rows of class logits");
    } else if (ops[i].type == VCG_MLP_SAVE) {
      VCG_REQUIRE(dim <= kMaxSaved, "mlp chain: SAVE row too wide");
      saved_dim = dim;
    } else if (ops[i].type == VCG_MLP_ADDSAVED) {
      VCG_REQUIRE(saved_dim == dim, "mlp chain: ADDSAVED needs a saved row of the same width");
    } else {
      VCG_REQUIRE(ops[i].type == VCG_MLP_RELU || ops[i].type == VCG_MLP_GELU, "mlp chain: unknown step");
    }
  }
  launch_pdl(mlp_chain_kernel, rows, 256, 0, s, x0, dim0, stride0, x1, dim1, stride1, prog, out, out_stride);
}

void launch_cross_attention(const vcg_cross_attn_params& p, const float* lang, const float* vision, int B, int T, float* out,
                            cudaStream_t s) {
  if (B == 0) return;
  VCG_REQUIRE(T >= 2 && T <= 40, "cross attention: 2..40 frames per clip");
  VCG_REQUIRE(p.num_heads >= 1 && p.num_heads <= 16 && 128 % p.num_heads == 0, "cross attention: bad head count");
  const size_t smem = static_cast<size_t>(3) * T * 128 * sizeof(float);
  static PerDeviceMax configured;
  if (configured.raise(smem)) {
    VCG_CUDA(cudaFuncSetAttribute(cross_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  launch_pdl(cross_attention_kernel, B, 128, smem, s, p, lang, vision, T, out);
}

void launch_self_attention_first(const vcg_self_attn_params& p, const float* vision, const float* lang, int B, int T,
                                 float* out, cudaStream_t s) {
  if (B == 0) return;
  VCG_REQUIRE(T >= 1 && T <= 40, "self attention: 1..40 frames per clip");
  VCG_REQUIRE(p.num_heads >= 1 && p.num_heads <= 16 && 128 % p.num_heads == 0, "self attention: bad head count");
  const size_t smem = static_cast<size_t>(3) * (T + 1) * 128 * sizeof(float);
  static PerDeviceMax configured;
  if (configured.raise(smem)) {
    VCG_CUDA(cudaFuncSetAttribute(self_attention_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  launch_pdl(self_attention_first_kernel, B, 128, smem, s, p, vision, lang, T, out);
}

void launch_bilinear_contract(const float* y, const float* x1, const float* bias, int rows, int in1, int out_features,
                              float* out, cudaStream_t s) {
  if (rows == 0) return;
  VCG_REQUIRE(in1 >= 1 && out_features >= 1, "bilinear contract: bad sizes");
  const long n_out = static_cast<long>(rows) * out_features;
  const long blocks = (n_out * 32 + 255) / 256;
  launch_pdl(bilinear_contract_kernel, static_cast<unsigned>(blocks), 256, 0, s, y, x1, bias, n_out, in1, out_features, out);
}

void launch_center_attention(const vcg_center_attn_params& p, const float* x, int B, int W, float* out, cudaStream_t s) {
  if (B == 0) return;
  VCG_REQUIRE(W >= 1 && W <= 9 && (W & 1), "centre attention: odd window of at most 9 clips");
  VCG_REQUIRE(p.num_heads >= 1 && p.num_heads <= 16 && 128 % p.num_heads == 0, "centre attention: bad head count");
  VCG_REQUIRE(p.pos_w && p.pos_b && p.pos_norm_w && p.pos_norm_b && p.pos_bias && p.q_w && p.k_w && p.v_w,
              "centre attention: missing parameter");
  launch_pdl(center_attention_kernel, B, 128, 0, s, p, x, W, out);
}

void launch_window_stack(const vcg_window_stack_params& p, const float* x, int B, int W, float* logits, float* probs,
                         cudaStream_t s) {
  if (B == 0) return;
  VCG_REQUIRE(W >= 1 && W <= 9 && (W & 1), "window stack: odd window of at most 9 clips");
  VCG_REQUIRE(p.num_layers >= 0 && p.num_layers <= 8 && p.pos_bias_stride >= W, "window stack: bad parameters");
  const size_t smem = (static_cast<size_t>(5) * W * 128 + static_cast<size_t>(2) * W * 512 + static_cast<size_t>(16) * W * W) * sizeof(float);
  static PerDeviceMax configured;
  if (configured.raise(smem)) {
    VCG_CUDA(cudaFuncSetAttribute(window_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  launch_pdl(window_stack_kernel, B, 256, smem, s, p, x, W, logits, probs);
}

}  // namespace vcg
