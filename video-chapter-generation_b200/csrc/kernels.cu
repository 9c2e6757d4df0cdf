// HBM-bound kernels of the scoring path: frame preprocessing, max/avg pooling (with the temporal shift folded in),
// BERT embedding + LayerNorm, LayerNorm, and weight packing.  All use 8/16-byte vector accesses with consecutive
// threads on consecutive addresses; grids are sized from the element count (>= several waves at bench sizes).
#include "kernels.cuh"
#include "launch.cuh"
#include "tensormap.h"
#include <cuda_bf16.h>
#include <type_traits>

namespace vcg {

namespace {

template <bool FP32>
using elem_t = typename std::conditional<FP32, float, __nv_bfloat16>::type;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

// 8 consecutive elements as fp32
template <bool FP32>
__device__ __forceinline__ void load8(const elem_t<FP32>* p, float (&v)[8]) {
  if constexpr (FP32) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    float2 f;
    f = unpack_bf16x2(q.x); v[0] = f.x; v[1] = f.y;
    f = unpack_bf16x2(q.y); v[2] = f.x; v[3] = f.y;
    f = unpack_bf16x2(q.z); v[4] = f.x; v[5] = f.y;
    f = unpack_bf16x2(q.w); v[6] = f.x; v[7] = f.y;
  }
}
template <bool FP32>
__device__ __forceinline__ void store8(elem_t<FP32>* p, const float (&v)[8]) {
  if constexpr (FP32) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
    q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = q;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Element offset of padded pixel (hp, wp) of image n in the stem input.  fp32: plain NHWC4 [n][Hp][Wp][4];
// bf16: row pairs interleaved per pixel, [n][Hp/2][Wp][2][4] (one stem K block = 8 pixels x 2 rows, conv_gemm_host.h).
template <bool FP32>
__device__ __forceinline__ long stem_offset(long n, int hp, int wp) {
  if constexpr (FP32) return ((n * kStemHp + hp) * kStemWp + wp) * 4L;
  else return (((n * (kStemHp / 2) + (hp >> 1)) * kStemWp + wp) * 2L + (hp & 1)) * 4L;
}

// ------------------------------------------------------------------------------------------------ preprocess
// ToTensor + Normalize (test_video_segment_point.py:142-145): (u8/255 - mean)/std, HWC -> zero-padded NHWC4.
// HBM-bound: 150 528 B read + 401 408 B written per frame (bf16).  One CTA = kPrePairs padded row pairs of one frame:
//   1. the 2*kPrePairs source rows (672 B each, 16-byte aligned) come in as coalesced 16-byte loads into shared memory;
//   2. (x/255 - mean)/std is a 3 x 256 table in shared memory, computed with exactly the expression above;
//   3. bf16: the stem input interleaves each padded row pair per pixel ([n][Hp/2][Wp][2][4], stem_offset), so one thread
//      = one pixel column of a pair = ONE 16-byte store, consecutive threads on consecutive addresses (512 B per warp);
//      fp32: two float4 stores per thread, each a 512-byte run per warp.
// Rows of the pair outside the image are the zero border (bf16: written as zeros, which they already are).
__constant__ float c_mean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float c_std[3] = {0.229f, 0.224f, 0.225f};
constexpr int kPrePairs = 4;                                   // padded row pairs per CTA
constexpr int kPreChunks = (kImg / 2 + 1 + kPrePairs - 1) / kPrePairs;   // 113 pairs hold image rows -> 29 CTAs per frame
constexpr int kPreRowBytes = kImg * 3;                         // 672

template <bool FP32>
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ frames,
                                                            const int32_t* __restrict__ frame_index,
                                                            const int32_t* __restrict__ clip_start, int T, int n_frames,
                                                            elem_t<FP32>* __restrict__ out) {
  __shared__ __align__(16) uint8_t s_rows[kPrePairs * 2][kPreRowBytes];
  __shared__ float s_lut[3][256];
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  for (int i = tid; i < 768; i += 256) {
    const int c = i >> 8, v = i & 255;
    s_lut[c][v] = (static_cast<float>(v) / 255.0f - c_mean[c]) / c_std[c];
  }
  const long n = blockIdx.x / kPreChunks;                       // destination image
  const int q0 = 1 + (blockIdx.x % kPreChunks) * kPrePairs;     // first padded row pair (pair q = padded rows 2q, 2q+1)
  pdl_wait();
  long f = n;
  if (clip_start) f = static_cast<long>(__ldg(clip_start + n / T)) + (n % T);
  else if (frame_index) f = __ldg(frame_index + n);
  if (n_frames > 0) f = min(max(f, 0L), static_cast<long>(n_frames) - 1);   // never read outside the frame buffer
  const uint8_t* src = frames + f * (static_cast<long>(kImg) * kPreRowBytes);
  constexpr int V = kPreRowBytes / 16;                          // 42 16-byte vectors per row
  for (int i = tid; i < kPrePairs * 2 * V; i += 256) {
    const int r = i / V, v = i % V;
    const int h = 2 * (q0 + (r >> 1)) + (r & 1) - kStemPad;     // image row of padded row 2q + (r & 1)
    if (h >= 0 && h < kImg)
      *reinterpret_cast<uint4*>(&s_rows[r][v * 16]) = __ldg(reinterpret_cast<const uint4*>(src + h * kPreRowBytes) + v);
  }
  __syncthreads();
  for (int i = tid; i < kPrePairs * kImg; i += 256) {
    const int pr = i / kImg, w = i % kImg;
    const int q = q0 + pr;
    if (q > kImg / 2 + 1) break;                                // past the last pair that holds an image row
    float v[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int h = 2 * q + r - kStemPad;
      const bool in = h >= 0 && h < kImg;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[r][c] = in ? s_lut[c][s_rows[pr * 2 + r][w * 3 + c]] : 0.f;
    }
    if constexpr (FP32) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int h = 2 * q + r - kStemPad;
        if (h >= 0 && h < kImg)
          *reinterpret_cast<float4*>(out + stem_offset<true>(n, 2 * q + r, w + kStemPad)) = make_float4(v[r][0], v[r][1], v[r][2], 0.f);
      }
    } else {
      uint4 o;
      o.x = pack_bf16x2(v[0][0], v[0][1]); o.y = pack_bf16x2(v[0][2], 0.f);
      o.z = pack_bf16x2(v[1][0], v[1][1]); o.w = pack_bf16x2(v[1][2], 0.f);
      *reinterpret_cast<uint4*>(out + stem_offset<false>(n, 2 * q, w + kStemPad)) = o;
    }
  }
}

// fp32 NCHW (already normalised, the reference's img_clip) -> padded NHWC4
template <bool FP32>
__global__ void nchw_to_stem_kernel(const float* __restrict__ img, long total, elem_t<FP32>* __restrict__ out) {
  pdl_enter();
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int w = static_cast<int>(idx % kImg);
  const int h = static_cast<int>((idx / kImg) % kImg);
  const long n = idx / (kImg * kImg);
  const float* src = img + n * 3L * kImg * kImg + h * kImg + w;
  const float r = __ldg(src), g = __ldg(src + kImg * kImg), b = __ldg(src + 2 * kImg * kImg);
  elem_t<FP32>* dst = out + stem_offset<FP32>(n, h + kStemPad, w + kStemPad);
  if constexpr (FP32) {
    *reinterpret_cast<float4*>(dst) = make_float4(r, g, b, 0.f);
  } else {
    uint2 q;
    q.x = pack_bf16x2(r, g);
    q.y = pack_bf16x2(b, 0.f);
    *reinterpret_cast<uint2*>(dst) = q;
  }
}

// ------------------------------------------------------------------------------------------------ max pool + TSM
// MaxPool2d(3, stride 2, pad 1) over NHWC [n,112,112,64]; writes x and the temporally shifted copy that layer1.0.conv1
// consumes (ops/temporal_shift.py:34-51): channels [0,fold) of frame t land in frame t-1, [fold,2*fold) in frame t+1.
// One thread = 8 channels of one output pixel.
template <bool FP32>
__global__ void maxpool_tsm_kernel(const elem_t<FP32>* __restrict__ in, long total, elem_t<FP32>* __restrict__ out,
                                   elem_t<FP32>* __restrict__ shifted, int T, int fold) {
  pdl_enter();
  constexpr int C = 64, HI = kStemOut, HO = kStemOut / 2;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % (C / 8));
  const int wo = static_cast<int>((idx / (C / 8)) % HO);
  const int ho = static_cast<int>((idx / (C / 8 * HO)) % HO);
  const long n = idx / (C / 8 * HO * HO);
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int y = 2 * ho + dy;
    if (y < 0 || y >= HI) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = 2 * wo + dx;
      if (x < 0 || x >= HI) continue;
      float v[8];
      load8<FP32>(in + ((n * HI + y) * HI + x) * C + c8 * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
    }
  }
  const long pix = (n * HO + ho) * HO + wo;
  store8<FP32>(out + pix * C + c8 * 8, m);
  if (shifted) {
    const int c0 = c8 * 8;
    const int t = static_cast<int>(n % T);
    long dpix = pix;
    bool ok = true;
    if (c0 < fold) { ok = t >= 1; dpix = pix - HO * HO; }
    else if (c0 < 2 * fold) { ok = t + 1 < T; dpix = pix + HO * HO; }
    if (ok) store8<FP32>(shifted + dpix * C + c0, m);
  }
}

// ------------------------------------------------------------------------------------------------ avg pool
// AdaptiveAvgPool2d(1) over NHWC [n, hw, C] -> fp32 [n, C]
template <bool FP32>
__global__ void avgpool_kernel(const elem_t<FP32>* __restrict__ in, long total, int hw, int C,
                               float* __restrict__ out, elem_t<FP32>* __restrict__ out_act) {
  pdl_enter();
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % (C / 8));
  const long n = idx / (C / 8);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = 0; p < hw; ++p) {
    float v[8];
    load8<FP32>(in + (n * hw + p) * C + c8 * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] += v[e];
  }
  const float d = static_cast<float>(hw);
  float* o = out + n * C + c8 * 8;
  *reinterpret_cast<float4*>(o) = make_float4(s[0] / d, s[1] / d, s[2] / d, s[3] / d);
  *reinterpret_cast<float4*>(o + 4) = make_float4(s[4] / d, s[5] / d, s[6] / d, s[7] / d);
  if (out_act) {
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = s[e] / d;
    store8<FP32>(out_act + n * C + c8 * 8, m);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row of 768: 3 chunks of 8 elements per lane, statistics in fp32 with warp-shuffle reductions.
template <bool FP32>
__device__ __forceinline__ void ln_row_768(float (&x)[3][8], const float* __restrict__ gamma,
                                           const float* __restrict__ beta, float eps, elem_t<FP32>* __restrict__ y,
                                           int lane) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) s += x[c][e];
  const float mean = warp_sum(s) * (1.0f / 768.0f);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float d = x[c][e] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / 768.0f) + eps);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = c * 256 + lane * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (x[c][e] - mean) * rstd * g[e] + b[e];
    store8<FP32>(y + col, o);
  }
}

// Two rows per warp: both rows' loads are in flight together and the two dependent shuffle reductions of each row
// interleave (the kernel is latency-bound: 3 KB per row, everything L2-resident right after the producing GEMM);
// gamma / beta are fetched once for both rows.
template <bool FP32>
__global__ void layernorm768_kernel(const elem_t<FP32>* __restrict__ x, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, elem_t<FP32>* __restrict__ y, int rows,
                                    const int32_t* __restrict__ rows_dev, float eps) {
  pdl_enter();
  const int row0 = 2 * (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  const int lane = threadIdx.x & 31;
  if (row0 >= rows) return;                      // `rows` = allocated rows: every load below stays inside the buffer
  // CTAs beyond the first wave (8 CTAs x 148 SMs) start late anyway: they look at the device-side row count first and
  // leave without touching memory when their rows do not exist (with token packing that is the usual case)
  if (rows_dev && blockIdx.x >= 8 * 148) {
    rows = min(rows, ld_chain_i32(rows_dev));
    if (row0 >= rows) return;
  }
  const bool two_alloc = row0 + 1 < rows;
  float v[2][3][8];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (r == 0 || two_alloc) load8<FP32>(x + (row0 + r) * 768L + c * 256 + lane * 8, v[r][c]);
      else
#pragma unroll
        for (int e = 0; e < 8; ++e) v[r][c][e] = 0.f;
    }
  // token-packed BERT: only the packed rows exist.  Read the device-side count AFTER issuing the row loads (it is a
  // dependent L2 round trip that would otherwise sit in front of them).
  if (rows_dev) rows = min(rows, ld_chain_i32(rows_dev));
  if (row0 >= rows) return;
  const bool two = row0 + 1 < rows;
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) s[r] += v[r][c][e];
  float mean[2], q[2] = {0.f, 0.f}, rstd[2];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
    s[1] += __shfl_xor_sync(0xffffffffu, s[1], o);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mean[r] = s[r] * (1.0f / 768.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[r][c][e] - mean[r]; q[r] += d * d; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    q[0] += __shfl_xor_sync(0xffffffffu, q[0], o);
    q[1] += __shfl_xor_sync(0xffffffffu, q[1], o);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) rstd[r] = 1.0f / sqrtf(q[r] * (1.0f / 768.0f) + eps);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = c * 256 + lane * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r == 1 && !two) break;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[r][c][e] - mean[r]) * rstd[r] * g[e] + b[e];
      store8<FP32>(y + (row0 + r) * 768L + col, o);
    }
  }
}

// BertEmbeddings (modeling_bert.py:72-113): word[id] + position[pos] + token_type[0], LayerNorm(eps 1e-12)
template <bool FP32>
__global__ void bert_embed_ln_kernel(const int64_t* __restrict__ ids, int rows, int L,
                                     const int32_t* __restrict__ tok_src, const int32_t* __restrict__ rows_dev, int vocab,
                                     const elem_t<FP32>* __restrict__ word, const elem_t<FP32>* __restrict__ pos,
                                     const elem_t<FP32>* __restrict__ type, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, elem_t<FP32>* __restrict__ out) {
  pdl_enter();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (rows_dev) rows = min(rows, ld_chain_i32(rows_dev));
  if (row >= rows) return;
  const int src = tok_src ? tok_src[row] : row;   // packed row -> b*L + j
  const long id = min(max(ids[src], 0L), static_cast<long>(vocab) - 1);   // never index outside the embedding table
  const int p = src % L;
  float v[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = c * 256 + lane * 8;
    float a[8], b[8], t[8];
    load8<FP32>(word + id * 768L + col, a);
    load8<FP32>(pos + p * 768L + col, b);
    load8<FP32>(type + col, t);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[c][e] = (a[e] + t[e]) + b[e];   // HF order: (inputs + token_type) + position
  }
  ln_row_768<FP32>(v, gamma, beta, 1e-12f, out + row * 768L, lane);
}

// ------------------------------------------------------------------------------------------------ token packing
// Variable-length BERT: padded positions are masked as keys in every layer and only h[:,0] is consumed, so tokens with
// attention_mask == 0 never influence the output and are dropped (token 0 is always kept as a query).  One CTA:
//   cu[b]      first packed row of clip b (cu[B] = number of packed rows, also written to *total)
//   tok_src[m] b*L + j of packed row m;  key_ok[m] = attention_mask[b, j] != 0
__global__ void __launch_bounds__(1024) bert_pack_kernel(const int64_t* __restrict__ mask, int B, int L,
                                                          int32_t* __restrict__ cu, int32_t* __restrict__ tok_src,
                                                          uint8_t* __restrict__ key_ok, int32_t* __restrict__ total) {
  pdl_enter();
  extern __shared__ int s_cnt[];   // [B + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // (the kernel is one CTA and purely latency-bound: the four 32-token groups of a step are loaded together)
  for (int b = warp; b < B; b += nwarps) {
    int c = 0;
    for (int j0 = 0; j0 < L; j0 += 128) {
      bool v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * 32 + lane;
        v[u] = j < L && (mask[static_cast<long>(b) * L + j] != 0 || j == 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) c += __popc(__ballot_sync(0xffffffffu, v[u]));
    }
    if (lane == 0) s_cnt[b] = c;
  }
  __syncthreads();
  if (warp == 0) {                 // exclusive scan of the per-clip counts: 32 clips per step, carry in a register
    int carry = 0;
    for (int b0 = 0; b0 < B; b0 += 32) {
      const int b = b0 + lane;
      const int c = b < B ? s_cnt[b] : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (b < B) { s_cnt[b] = carry + incl - c; cu[b] = carry + incl - c; }
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) { s_cnt[B] = carry; cu[B] = carry; *total = carry; }
  }
  __syncthreads();
  for (int b = warp; b < B; b += nwarps) {
    int base = s_cnt[b];
    for (int j0 = 0; j0 < L; j0 += 128) {
      bool m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * 32 + lane;
        m[u] = j < L && mask[static_cast<long>(b) * L + j] != 0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * 32 + lane;
        const bool v = m[u] || j == 0;
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        if (v) {
          const int dst = base + __popc(bal & ((1u << lane) - 1));
          tok_src[dst] = b * L + j;
          key_ok[dst] = m[u] ? 1 : 0;
        }
        base += __popc(bal);
      }
    }
  }
}

// [CLS] rows of the packed hidden states -> dense [B, 768] (the pooler GEMM's A operand)
// CLS rows of a RAW pre-LayerNorm matrix -> LayerNorm -> out (bf16 path with LayerNorm folded into the GEMMs: the final
// LayerNorm is only ever needed for the B pooled rows).  One warp per clip, exact two-pass statistics.
__global__ void gather_ln_rows768_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ row_of, int stride,
                                         int B, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                         __nv_bfloat16* __restrict__ out) {
  pdl_enter();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const long row = row_of ? row_of[b] : static_cast<long>(b) * stride;
  float v[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) load8<false>(x + row * 768L + c * 256 + lane * 8, v[c]);
  ln_row_768<false>(v, gamma, beta, eps, out + b * 768L, lane);
}

// Linear layer that consumes LayerNorm(x) re-expressed on the raw x (conv_gemm.cuh, LNF epilogue):
//   LN(x) W^T + b = rstd * (x (W*gamma)^T - mean * c1) + c2b,   c1[n] = sum_k W'[n,k],  c2b[n] = b[n] + sum_k beta[k] W[n,k]
// W' = bf16(W * gamma) is what the tensor core multiplies, so c1 is summed over the ROUNDED values (the mean term then
// cancels exactly).  One warp per output row n.
__global__ void fold_ln_linear_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ bias, int N, int K,
                                      __nv_bfloat16* __restrict__ w_out, float* __restrict__ c1, float* __restrict__ c2b) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float s1 = 0.f, s2 = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = w[static_cast<long>(n) * K + k];
    const __nv_bfloat16 wb = __float2bfloat16_rn(wv * gamma[k]);
    w_out[static_cast<long>(n) * K + k] = wb;
    s1 += __bfloat162float(wb);
    s2 = fmaf(beta[k], wv, s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    c1[n] = s1;
    c2b[n] = s2 + bias[n];
  }
}

template <bool FP32>
__global__ void gather_rows768_kernel(const elem_t<FP32>* __restrict__ x, const int32_t* __restrict__ row_of, int stride,
                                      int B, elem_t<FP32>* __restrict__ out) {
  pdl_enter();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const long src = row_of ? row_of[b] : static_cast<long>(b) * stride;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[8];
    load8<FP32>(x + src * 768 + c * 256 + lane * 8, v);
    store8<FP32>(out + static_cast<long>(b) * 768 + c * 256 + lane * 8, v);
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
// conv weight [Cout,Cin,k,k] fp32 + eval-mode BatchNorm -> [Cout][k][k][Cin] (scaled) and bias[Cout]
template <bool FP32>
__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                 const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                 const float* __restrict__ bn_var, float eps, int Cout, int Cin, int k,
                                 elem_t<FP32>* __restrict__ w_out, float* __restrict__ bias_out) {
  const long total = static_cast<long>(Cout) * Cin * k * k;
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int ci = static_cast<int>(idx % Cin);
  const int kk = static_cast<int>((idx / Cin) % (k * k));
  const int co = static_cast<int>(idx / (static_cast<long>(Cin) * k * k));
  const float scale = bn_w[co] / sqrtf(bn_var[co] + eps);
  const float v = w[(static_cast<long>(co) * Cin + ci) * k * k + kk] * scale;
  if constexpr (FP32) w_out[idx] = v; else w_out[idx] = __float2bfloat16_rn(v);
  if (ci == 0 && kk == 0) bias_out[co] = bn_b[co] - bn_mean[co] * scale;
}

// layer1.0.conv1 for the shared-stem path: [Cout,Cin,1,1] + BN -> [Cout][3 temporal taps][Cin] with channel masks
// (tap 0 = same frame: channels >= 2f; tap 1 = frame t+1: channels < f; tap 2 = frame t-1: channels [f, 2f)), bf16
__global__ void pack_conv1_shared_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                         const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                         const float* __restrict__ bn_var, float eps, int Cout, int Cin, int fold,
                                         __nv_bfloat16* __restrict__ w_out, float* __restrict__ bias_out) {
  const int total = Cout * 3 * Cin;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ci = idx % Cin, tap = (idx / Cin) % 3, co = idx / (3 * Cin);
  const float scale = bn_w[co] / sqrtf(bn_var[co] + eps);
  const bool keep = tap == 0 ? ci >= 2 * fold : tap == 1 ? ci < fold : (ci >= fold && ci < 2 * fold);
  w_out[idx] = __float2bfloat16_rn(keep ? w[co * Cin + ci] * scale : 0.f);
  if (ci == 0 && tap == 0) bias_out[co] = bn_b[co] - bn_mean[co] * scale;
}

// stem weight [64,3,7,7] -> fp32: [64][7][8 px][4]; bf16: [64][4 row pairs][8 px][2 rows][4]; zero where kh = 7, kw = 7, c = 3
template <bool FP32>
__global__ void pack_stem_kernel(const float* __restrict__ w, const float* __restrict__ bn_w,
                                 const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                 const float* __restrict__ bn_var, float eps, elem_t<FP32>* __restrict__ w_out,
                                 float* __restrict__ bias_out) {
  constexpr int per_co = FP32 ? 7 * 32 : 4 * 64;
  const int total = 64 * per_co;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = idx / per_co, k = idx % per_co;
  int c, kw, kh;
  if constexpr (FP32) {
    c = k % 4; kw = (k / 4) % 8; kh = k / 32;
  } else {
    c = k % 4; kh = 2 * (k / 64) + ((k / 4) % 2); kw = (k / 8) % 8;
  }
  const float scale = bn_w[co] / sqrtf(bn_var[co] + eps);
  float v = 0.f;
  if (c < 3 && kw < 7 && kh < 7) v = w[((co * 3 + c) * 7 + kh) * 7 + kw] * scale;
  if constexpr (FP32) w_out[idx] = v; else w_out[idx] = __float2bfloat16_rn(v);
  if (k == 0) bias_out[co] = bn_b[co] - bn_mean[co] * scale;
}

template <bool FP32>
__global__ void convert_kernel(const float* __restrict__ in, elem_t<FP32>* __restrict__ out, long n) {
  pdl_enter();
  // eight elements per thread (two 16-byte loads, one 16-byte store in the bf16 case); scalar tail / unaligned fallback
  const long i8 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) * 8;
  if (i8 >= n) return;
  const bool vec = i8 + 8 <= n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    float v[8];
    load8<true>(in + i8, v);
    store8<FP32>(out + i8, v);
  } else {
    for (long i = i8; i < min(i8 + 8, n); ++i) {
      if constexpr (FP32) out[i] = in[i]; else out[i] = __float2bfloat16_rn(in[i]);
    }
  }
}
template <bool FP32>
__global__ void cast_to_f32_kernel(const elem_t<FP32>* __restrict__ in, float* __restrict__ out, long n) {
  pdl_enter();
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  if constexpr (FP32) out[idx] = in[idx]; else out[idx] = __bfloat162float(in[idx]);
}
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (x < cols && y0 + j < rows) tile[j][threadIdx.x] = in[static_cast<long>(y0 + j) * cols + x];
  __syncthreads();
  const int ox = blockIdx.y * 32 + threadIdx.x;   // row index of input = col of output
  const int oy0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (ox < rows && oy0 + j < cols) out[static_cast<long>(oy0 + j) * rows + ox] = tile[threadIdx.x][j];
}

inline unsigned blocks_for(long total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

}  // namespace

#define VCG_DISPATCH(fp32, ...)                 \
  do {                                          \
    if (fp32) { constexpr bool FP = true; __VA_ARGS__; } \
    else { constexpr bool FP = false; __VA_ARGS__; }     \
  } while (0)

void launch_preprocess_u8(const uint8_t* frames, const int32_t* frame_index, int n, void* out, bool fp32,
                          cudaStream_t s, int n_frames) {
  if (n == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(preprocess_u8_kernel<FP>, static_cast<unsigned>(n) * kPreChunks, 256, 0, s, frames, frame_index,
                                 nullptr, 1, n_frames, static_cast<elem_t<FP>*>(out))));
  VCG_CUDA(cudaGetLastError());
}
void launch_preprocess_u8_clips(const uint8_t* frames, const int32_t* clip_start, int B, int T, void* out, bool fp32,
                                cudaStream_t s, int n_frames) {
  if (B * T == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(preprocess_u8_kernel<FP>, static_cast<unsigned>(B * T) * kPreChunks, 256, 0, s, frames, nullptr,
                                 clip_start, T, n_frames, static_cast<elem_t<FP>*>(out))));
  VCG_CUDA(cudaGetLastError());
}
void launch_nchw_to_stem(const float* img, int n, void* out, bool fp32, cudaStream_t s) {
  const long total = static_cast<long>(n) * kImg * kImg;
  if (total == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(nchw_to_stem_kernel<FP>, blocks_for(total, 256), 256, 0, s, img, total, static_cast<elem_t<FP>*>(out))));
  VCG_CUDA(cudaGetLastError());
}
void launch_maxpool_tsm(const void* in, int n, void* out, void* out_shifted, int T, int fold, bool fp32,
                        cudaStream_t s) {
  VCG_REQUIRE(out_shifted == nullptr || (fold % 8 == 0 && fold > 0), "TSM fold of the stem output must be a multiple of 8");
  const long total = static_cast<long>(n) * 56 * 56 * 8;
  if (total == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(maxpool_tsm_kernel<FP>, blocks_for(total, 256), 256, 0, s, static_cast<const elem_t<FP>*>(in), total, static_cast<elem_t<FP>*>(out),
                         static_cast<elem_t<FP>*>(out_shifted), T, fold)));
  VCG_CUDA(cudaGetLastError());
}
void launch_avgpool(const void* in, int n, int hw, int C, float* out, void* out_act, cudaStream_t s, bool fp32) {
  const long total = static_cast<long>(n) * (C / 8);
  if (total == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(avgpool_kernel<FP>, blocks_for(total, 128), 128, 0, s, static_cast<const elem_t<FP>*>(in), total, hw, C, out, static_cast<elem_t<FP>*>(out_act))));
  VCG_CUDA(cudaGetLastError());
}
void launch_bert_pack(const int64_t* mask, int B, int L, int32_t* cu, int32_t* tok_src, uint8_t* key_ok, int32_t* total,
                      cudaStream_t s) {
  if (B == 0) return;
  launch_pdl(bert_pack_kernel, 1, 1024, (B + 1) * sizeof(int), s, mask, B, L, cu, tok_src, key_ok, total);
  VCG_CUDA(cudaGetLastError());
}
void launch_gather_rows768(const void* x, const int32_t* row_of, int stride, int B, void* out, bool fp32, cudaStream_t s) {
  if (B == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(gather_rows768_kernel<FP>, blocks_for(B, 8), 256, 0, s, static_cast<const elem_t<FP>*>(x), row_of, stride, B, static_cast<elem_t<FP>*>(out))));
  VCG_CUDA(cudaGetLastError());
}
void launch_bert_embed_ln(const int64_t* ids, int rows, int L, const int32_t* tok_src, const int32_t* rows_dev, int vocab,
                          const void* word, const void* pos, const void* type, const float* gamma, const float* beta,
                          void* out, bool fp32, cudaStream_t s) {
  if (rows == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(bert_embed_ln_kernel<FP>, blocks_for(rows, 8), 256, 0, s, ids, rows, L, tok_src, rows_dev, vocab, static_cast<const elem_t<FP>*>(word), static_cast<const elem_t<FP>*>(pos),
                         static_cast<const elem_t<FP>*>(type), gamma, beta, static_cast<elem_t<FP>*>(out))));
  VCG_CUDA(cudaGetLastError());
}
void launch_layernorm(const void* x, const float* gamma, const float* beta, void* y, int rows, int cols, float eps,
                      bool fp32, cudaStream_t s, const int32_t* rows_dev) {
  VCG_REQUIRE(cols == 768, "LayerNorm kernel is specialised for 768 columns");
  if (rows == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(layernorm768_kernel<FP>, blocks_for(rows, 16), 256, 0, s, static_cast<const elem_t<FP>*>(x), gamma, beta, static_cast<elem_t<FP>*>(y), rows, rows_dev, eps)));
  VCG_CUDA(cudaGetLastError());
}
void launch_gather_ln_rows768(const void* x, const int32_t* row_of, int stride, int B, const float* gamma, const float* beta,
                              float eps, void* out, cudaStream_t s) {
  if (B == 0) return;
  launch_pdl(gather_ln_rows768_kernel, blocks_for(B, 8), 256, 0, s, static_cast<const __nv_bfloat16*>(x), row_of, stride, B,
             gamma, beta, eps, static_cast<__nv_bfloat16*>(out));
}
void launch_fold_ln_linear(const float* w, const float* gamma, const float* beta, const float* bias, int N, int K, void* w_out,
                           float* c1, float* c2b, cudaStream_t s) {
  fold_ln_linear_kernel<<<blocks_for(N, 8), 256, 0, s>>>(w, gamma, beta, bias, N, K, static_cast<__nv_bfloat16*>(w_out), c1, c2b);
  VCG_CUDA(cudaGetLastError());
}
void launch_pack_conv(const float* w, const float* bn_w, const float* bn_b, const float* bn_mean, const float* bn_var,
                      float eps, int Cout, int Cin, int k, void* w_out, float* bias_out, bool fp32, cudaStream_t s) {
  const long total = static_cast<long>(Cout) * Cin * k * k;
  VCG_DISPATCH(fp32, (pack_conv_kernel<FP><<<blocks_for(total, 256), 256, 0, s>>>(
                         w, bn_w, bn_b, bn_mean, bn_var, eps, Cout, Cin, k, static_cast<elem_t<FP>*>(w_out),
                         bias_out)));
  VCG_CUDA(cudaGetLastError());
}
void launch_pack_conv1_shared(const float* w, const float* bn_w, const float* bn_b, const float* bn_mean,
                              const float* bn_var, float eps, int Cout, int Cin, int fold, void* w_out, float* bias_out,
                              cudaStream_t s) {
  const int total = Cout * 3 * Cin;
  pack_conv1_shared_kernel<<<blocks_for(total, 256), 256, 0, s>>>(w, bn_w, bn_b, bn_mean, bn_var, eps, Cout, Cin, fold,
                                                                  static_cast<__nv_bfloat16*>(w_out), bias_out);
  VCG_CUDA(cudaGetLastError());
}
void launch_pack_stem(const float* w, const float* bn_w, const float* bn_b, const float* bn_mean, const float* bn_var,
                      float eps, void* w_out, float* bias_out, bool fp32, cudaStream_t s) {
  const int total = 64 * (fp32 ? 7 * 32 : 4 * 64);
  VCG_DISPATCH(fp32, (pack_stem_kernel<FP><<<blocks_for(total, 256), 256, 0, s>>>(
                         w, bn_w, bn_b, bn_mean, bn_var, eps, static_cast<elem_t<FP>*>(w_out), bias_out)));
  VCG_CUDA(cudaGetLastError());
}
// debug: order-independent 64-bit checksum of a buffer (sum of word * position weight), accumulated into *out
// ------------------------------------------------------------------------------------------------ batch-statistics BatchNorm
// Opt-in compatibility mode for reference caller #1, which nulls the running statistics of every BatchNorm2d after
// .eval() (test_video_segment_point.py:116-122): F.batch_norm then normalises with the statistics of the batch it is
// given.  Three small HBM-bound kernels over NHWC [rows, C] activations (rows = frames * H * W of ONE forward call):
//   bn_partial_stats: per-CTA double sums of x and x^2 per channel (fixed grid, fixed order -> deterministic)
//   bn_finish_stats : mean and 1/sqrt(biased var + eps) per channel
//   bn_apply        : (x - mean) * rstd * gamma + beta (+ residual) (ReLU)
constexpr int kBnThreads = 256;
template <bool FP32>
__global__ void __launch_bounds__(kBnThreads) bn_partial_stats_kernel(const elem_t<FP32>* __restrict__ x, long rows, int C,
                                                                      double* __restrict__ partial /*[grid][2][C]*/) {
  __shared__ double sh[kBnThreads][17];                    // 16 sums per thread (+1: bank spread)
  const int groups = C / 8;                                // 8 .. 256 channel groups of 8
  const int lanes = kBnThreads / groups;                   // rows in flight per CTA
  const int g = threadIdx.x % groups, lane = threadIdx.x / groups;
  double sum[8], sq[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sum[e] = sq[e] = 0.0;
  for (long r = static_cast<long>(blockIdx.x) * lanes + lane; r < rows; r += static_cast<long>(gridDim.x) * lanes) {
    float v[8];
    load8<FP32>(x + r * C + g * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) { sum[e] += v[e]; sq[e] += static_cast<double>(v[e]) * v[e]; }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { sh[threadIdx.x][e] = sum[e]; sh[threadIdx.x][8 + e] = sq[e]; }
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int e = 0; e < 8; ++e) { sum[e] += sh[l * groups + g][e]; sq[e] += sh[l * groups + g][8 + e]; }
    double* o = partial + static_cast<long>(blockIdx.x) * 2 * C;
#pragma unroll
    for (int e = 0; e < 8; ++e) { o[g * 8 + e] = sum[e]; o[C + g * 8 + e] = sq[e]; }
  }
}
__global__ void bn_finish_stats_kernel(const double* __restrict__ partial, int n_partial, int C, long rows, float eps,
                                       float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < n_partial; ++b) { s += partial[static_cast<long>(b) * 2 * C + c]; q += partial[static_cast<long>(b) * 2 * C + C + c]; }
  const double m = s / static_cast<double>(rows);
  double var = q / static_cast<double>(rows) - m * m;      // biased variance, as F.batch_norm normalises with
  if (var < 0.0) var = 0.0;
  mean[c] = static_cast<float>(m);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}
template <bool FP32>
__global__ void bn_apply_kernel(const elem_t<FP32>* __restrict__ x, long total /*rows * C / 8*/, int C,
                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const elem_t<FP32>* __restrict__ residual, int relu, elem_t<FP32>* __restrict__ out) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c0 = static_cast<int>(idx % (C / 8)) * 8;
  float v[8], r[8];
  load8<FP32>(x + idx * 8, v);
  if (residual) load8<FP32>(residual + idx * 8, r);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float y = (v[e] - __ldg(mean + c0 + e)) * __ldg(rstd + c0 + e) * __ldg(gamma + c0 + e) + __ldg(beta + c0 + e);
    if (residual) y += r[e];
    v[e] = relu ? fmaxf(y, 0.0f) : y;
  }
  store8<FP32>(out + idx * 8, v);
}
// TemporalShift.shift (ops/temporal_shift.py:34-51) as a stand-alone gather: frame t takes channels [0, fold) from frame
// t + 1 and [fold, 2 fold) from frame t - 1 of its own clip (zeros at the clip ends), the rest from itself.
template <bool FP32>
__global__ void tsm_shift_kernel(const elem_t<FP32>* __restrict__ x, long total /*n * hw * C / 8*/, int hw, int C, int T,
                                 int fold, elem_t<FP32>* __restrict__ out) {
  const long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c0 = static_cast<int>(idx % (C / 8)) * 8;
  const long row = idx / (C / 8);
  const int t = static_cast<int>((row / hw) % T);
  long src = row;
  bool ok = true;
  if (c0 < fold) { ok = t + 1 < T; src = row + hw; }
  else if (c0 < 2 * fold) { ok = t >= 1; src = row - hw; }
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ok) load8<FP32>(x + src * C + c0, v);
  store8<FP32>(out + idx * 8, v);
}
__global__ void checksum_kernel(const uint32_t* __restrict__ p, long n_words, unsigned long long* out) {
  unsigned long long s = 0;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n_words; i += static_cast<long>(gridDim.x) * blockDim.x)
    s += static_cast<unsigned long long>(p[i]) * static_cast<unsigned long long>((i & 1023) + 1);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}
void launch_checksum(const void* p, size_t bytes, unsigned long long* out, cudaStream_t s) {
  VCG_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long), s));
  if (bytes < 4) return;
  checksum_kernel<<<592, 256, 0, s>>>(static_cast<const uint32_t*>(p), static_cast<long>(bytes / 4), out);
  VCG_CUDA(cudaGetLastError());
}
int bn_stats_partials(long rows, int C) {
  const long per_cta = kBnThreads / (C / 8);
  return static_cast<int>(std::min<long>(592, (rows + per_cta - 1) / per_cta));
}
void launch_bn_batch_stats(const void* x, long rows, int C, float eps, double* partial, float* mean, float* rstd, bool fp32,
                           cudaStream_t s) {
  VCG_REQUIRE(C % 8 == 0 && C >= 64 && C <= 2048 && kBnThreads % (C / 8) == 0, "BatchNorm statistics: C must be 64 * 2^k, at most 2048");
  VCG_REQUIRE(rows >= 1, "BatchNorm statistics of an empty batch");
  const int grid = bn_stats_partials(rows, C);
  VCG_DISPATCH(fp32, (bn_partial_stats_kernel<FP><<<grid, kBnThreads, 0, s>>>(static_cast<const elem_t<FP>*>(x), rows, C, partial)));
  VCG_CUDA(cudaGetLastError());
  bn_finish_stats_kernel<<<(C + 127) / 128, 128, 0, s>>>(partial, grid, C, rows, eps, mean, rstd);
  VCG_CUDA(cudaGetLastError());
}
void launch_bn_apply(const void* x, long rows, int C, const float* mean, const float* rstd, const float* gamma,
                     const float* beta, const void* residual, bool relu, void* out, bool fp32, cudaStream_t s) {
  VCG_REQUIRE(C % 8 == 0, "BatchNorm: C must be a multiple of 8");
  const long total = rows * (C / 8);
  if (total == 0) return;
  VCG_DISPATCH(fp32, (bn_apply_kernel<FP><<<blocks_for(total, 256), 256, 0, s>>>(static_cast<const elem_t<FP>*>(x), total, C, mean, rstd, gamma, beta,
                                                                                static_cast<const elem_t<FP>*>(residual), relu ? 1 : 0,
                                                                                static_cast<elem_t<FP>*>(out))));
  VCG_CUDA(cudaGetLastError());
}
void launch_tsm_shift(const void* x, long n, int hw, int C, int T, int fold, void* out, bool fp32, cudaStream_t s) {
  VCG_REQUIRE(C % 8 == 0 && fold % 8 == 0 && fold >= 0 && 2 * fold <= C && T >= 1 && n % T == 0, "temporal shift: bad shape");
  const long total = n * hw * (C / 8);
  if (total == 0) return;
  VCG_DISPATCH(fp32, (tsm_shift_kernel<FP><<<blocks_for(total, 256), 256, 0, s>>>(static_cast<const elem_t<FP>*>(x), total, hw, C, T, fold,
                                                                                 static_cast<elem_t<FP>*>(out))));
  VCG_CUDA(cudaGetLastError());
}
void launch_convert(const float* in, void* out, long n, bool fp32, cudaStream_t s) {
  if (n == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(convert_kernel<FP>, blocks_for((n + 7) / 8, 256), 256, 0, s, in, static_cast<elem_t<FP>*>(out), n)));
  VCG_CUDA(cudaGetLastError());
}
void launch_cast_to_f32(const void* in, float* out, long n, bool fp32, cudaStream_t s) {
  if (n == 0) return;
  VCG_DISPATCH(fp32, (launch_pdl(cast_to_f32_kernel<FP>, blocks_for(n, 256), 256, 0, s, static_cast<const elem_t<FP>*>(in),
                                                                                 out, n)));
  VCG_CUDA(cudaGetLastError());
}
void launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t s) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, s>>>(in, out, rows, cols);
  VCG_CUDA(cudaGetLastError());
}

}  // namespace vcg
