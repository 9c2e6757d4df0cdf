// K1 with resize: uint8 HWC frames of ANY size -> bilinear resize to 224 x 224 -> (x/255 - mean)/std -> bf16 / fp32 stem
// input (or the resized uint8 frames), one kernel.
//
// The resize is the one the reference applies to PIL frames: GroupScale (data/transforms.py:79-92) =
// torchvision.transforms.Resize(size, BILINEAR) = PIL.Image.resize(.., BILINEAR), i.e. Pillow's ImagingResample
// (src/libImaging/Resample.c, un-vendored dependency; restated and pinned against PIL in oracle/resize_oracle.py):
// separable, horizontal pass first, triangle filter whose support grows with the down-scale factor (antialiasing),
// weights normalised in double precision and converted to 22-bit fixed point, each pass rounded to uint8.  The result is
// BIT-IDENTICAL to PIL's, so the integer arithmetic below must not be "improved".
//
// One CTA = the 8 output rows of four padded row pairs of one frame (the chunking of preprocess_u8_kernel):
//   1. the source rows those output rows depend on are streamed through shared memory one at a time (16-byte coalesced
//      loads when the row pitch allows it), each immediately reduced by the horizontal pass to 224 x 3 uint8;
//   2. the vertical pass combines those rows, the normalisation is a 3 x 256 table, and the store pattern is the one of
//      preprocess_u8_kernel (bf16: one 16-byte store per pixel column of a row pair).
// HBM-bound: Hs*Ws*3 B read (rows shared by neighbouring bands come from L2) + 401 408 B written per frame (bf16).
#include "kernels.cuh"
#include "launch.cuh"
#include <cuda_bf16.h>

namespace vcg {

namespace {

constexpr int kPrecBits = 32 - 8 - 2;
constexpr int kRsPairs = 4;                                   // padded row pairs per CTA
constexpr int kRsChunks = (kImg / 2 + 1 + kRsPairs - 1) / kRsPairs;
constexpr int kOutRowBytes = kImg * 3;                        // 672

__constant__ float c_rs_mean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float c_rs_std[3] = {0.229f, 0.224f, 0.225f};

// Resample.c precompute_coeffs + normalize_coeffs_8bpc, one thread per output index (IEEE double arithmetic, same order)
__global__ void resize_coeffs_kernel(int in_size, int out_size, int ksize, int32_t* __restrict__ bounds, int32_t* __restrict__ kk) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_size) return;
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const double ss = 1.0 / filterscale;
  const double center = (xx + 0.5) * scale;
  int xmin = static_cast<int>(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    double t = (x + xmin - center + 0.5) * ss;
    if (t < 0) t = -t;
    ww += t < 1.0 ? 1.0 - t : 0.0;
  }
  for (int x = 0; x < ksize; ++x) {
    double w = 0.0;
    if (x < xmax) {
      double t = (x + xmin - center + 0.5) * ss;
      if (t < 0) t = -t;
      w = t < 1.0 ? 1.0 - t : 0.0;
      if (ww != 0.0) w /= ww;
    }
    kk[xx * ksize + x] = w < 0 ? static_cast<int>(-0.5 + w * (1 << kPrecBits)) : static_cast<int>(0.5 + w * (1 << kPrecBits));
  }
  bounds[2 * xx] = xmin;
  bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ uint32_t pk_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(lo), "f"(hi));
  return r;
}

// Element offset of padded pixel (hp, wp) of image n in the stem input (kernels.cu stem_offset)
template <bool FP32>
__device__ __forceinline__ long rs_stem_offset(long n, int hp, int wp) {
  if constexpr (FP32) return ((n * kStemHp + hp) * kStemWp + wp) * 4L;
  else return (((n * (kStemHp / 2) + (hp >> 1)) * kStemWp + wp) * 2L + (hp & 1)) * 4L;
}

// KX: horizontal taps held in registers (0 = any ksize, weights read from shared memory per tap)
template <bool FP32, int KX>
__global__ void __launch_bounds__(256) resize_preprocess_u8_kernel(
    const uint8_t* __restrict__ frames, const int32_t* __restrict__ frame_index, const int32_t* __restrict__ clip_start, int T,
    int n_frames, int Hs, int Ws, const int32_t* __restrict__ xb, const int32_t* __restrict__ xk, int kx,
    const int32_t* __restrict__ yb, const int32_t* __restrict__ yk, int ky, int max_rows, void* __restrict__ out_stem,
    uint8_t* __restrict__ out_u8) {
  extern __shared__ __align__(16) uint8_t rs_smem[];
  const int src_pitch = ((Ws * 3 + 15) & ~15) + 64;                 // + zero tail: taps with zero weight may read past the row
  uint8_t* s_src = rs_smem;                                         // [2][src_pitch] (double buffered)
  uint8_t* s_tmp = s_src + 2 * src_pitch;                           // [max_rows][672]
  float* s_lut = reinterpret_cast<float*>(s_tmp + ((max_rows * kOutRowBytes + 15) & ~15));   // [3][256]
  int32_t* s_xk = reinterpret_cast<int32_t*>(s_lut + 768);          // [224][kx]
  int32_t* s_xb = s_xk + kImg * kx;                                 // [224][2]
  int32_t* s_yk = s_xb + kImg * 2;                                  // [8 output rows of the band][ky]
  int32_t* s_yb = s_yk + 2 * kRsPairs * ky;                         // [8][2]
  pdl_launch_dependents();
  const int tid = threadIdx.x;
  for (int i = tid; i < 768; i += 256) {
    const int c = i >> 8, v = i & 255;
    s_lut[i] = (static_cast<float>(v) / 255.0f - c_rs_mean[c]) / c_rs_std[c];
  }
  for (int i = tid; i < kImg * kx; i += 256) s_xk[i] = __ldg(xk + i);
  for (int i = tid; i < kImg * 2; i += 256) s_xb[i] = __ldg(xb + i);
  const long n = blockIdx.x / kRsChunks;                            // destination image
  const int q0 = 1 + (blockIdx.x % kRsChunks) * kRsPairs;           // first padded row pair (pair q = padded rows 2q, 2q+1)
  const int h_first = max(0, 2 * q0 - kStemPad), h_last = min(kImg - 1, 2 * (q0 + kRsPairs - 1) + 1 - kStemPad);
  for (int i = tid; i < 2 * kRsPairs * ky; i += 256) {
    const int h = 2 * q0 - kStemPad + i / ky;
    s_yk[i] = (h >= 0 && h < kImg) ? __ldg(yk + h * ky + i % ky) : 0;
  }
  if (tid < 2 * kRsPairs * 2) {
    const int h = 2 * q0 - kStemPad + tid / 2;
    s_yb[tid] = (h >= 0 && h < kImg) ? __ldg(yb + 2 * h + (tid & 1)) : 0;
  }
  for (int i = tid; i < 2 * 64; i += 256) s_src[(i >> 6) * src_pitch + (src_pitch - 64) + (i & 63)] = 0;
  __syncthreads();                                                  // tables visible to every thread
  pdl_wait();
  if (h_first > h_last) return;
  long f = n;
  if (clip_start) f = static_cast<long>(__ldg(clip_start + n / T)) + (n % T);
  else if (frame_index) f = __ldg(frame_index + n);
  if (n_frames > 0) f = min(max(f, 0L), static_cast<long>(n_frames) - 1);   // never read outside the frame buffer
  const long row_bytes = static_cast<long>(Ws) * 3;
  const uint8_t* src = frames + f * (static_cast<long>(Hs) * row_bytes);
  const int ys0 = __ldg(yb + 2 * h_first), ys1 = __ldg(yb + 2 * h_last) + __ldg(yb + 2 * h_last + 1);
  const bool vec_ok = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(frames) & 15) == 0);
  auto stage_row = [&](int y, int buf) {
    const uint8_t* r = src + y * row_bytes;
    uint8_t* d = s_src + buf * src_pitch;
    if (vec_ok) {
      for (int i = tid; i < static_cast<int>(row_bytes / 16); i += 256)
        *reinterpret_cast<uint4*>(d + i * 16) = __ldg(reinterpret_cast<const uint4*>(r) + i);
    } else {
      for (int i = tid; i < static_cast<int>(row_bytes); i += 256) d[i] = __ldg(r + i);
    }
  };
  // ---- horizontal pass, one source row at a time.  A thread owns up to three (output pixel, channel) columns for the
  //      whole band and keeps their fixed-point weights in registers.
  constexpr int KR = KX > 0 ? KX : 1;
  int base[3];
  int32_t wreg[3][KR];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int pcol = tid + 256 * j;
    const int xo = pcol / 3, c = pcol - xo * 3;
    base[j] = pcol < kOutRowBytes ? s_xb[2 * xo] * 3 + c : -1;
    if constexpr (KX > 0) {
#pragma unroll
      for (int x = 0; x < KX; ++x) wreg[j][x] = (pcol < kOutRowBytes && x < kx) ? s_xk[xo * kx + x] : 0;
    }
  }
  stage_row(ys0, 0);
  __syncthreads();
  for (int y = ys0; y < ys1; ++y) {
    const int buf = (y - ys0) & 1;
    if (y + 1 < ys1) stage_row(y + 1, buf ^ 1);
    const uint8_t* srow = s_src + buf * src_pitch;
    uint8_t* trow = s_tmp + (y - ys0) * kOutRowBytes;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (base[j] >= 0) {
        int acc = 1 << (kPrecBits - 1);
        if constexpr (KX > 0) {
#pragma unroll
          for (int x = 0; x < KX; ++x) acc += static_cast<int>(srow[base[j] + 3 * x]) * wreg[j][x];
        } else {
          const int pcol = tid + 256 * j;
          const int xo = pcol / 3;
          const int cnt = s_xb[2 * xo + 1];
          const int32_t* k = s_xk + xo * kx;
          for (int x = 0; x < cnt; ++x) acc += static_cast<int>(srow[base[j] + 3 * x]) * k[x];
        }
        trow[tid + 256 * j] = static_cast<uint8_t>(min(max(acc >> kPrecBits, 0), 255));
      }
    }
    __syncthreads();
  }
  // ---- vertical pass + normalise + store
  for (int i = tid; i < kRsPairs * kImg; i += 256) {
    const int pr = i / kImg, w = i - pr * kImg;
    const int q = q0 + pr;
    if (q > kImg / 2 + 1) break;                                    // past the last pair that holds an image row
    uint8_t px[2][3];
    bool in_img[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int h = 2 * q + r - kStemPad;
      in_img[r] = h >= 0 && h < kImg;
#pragma unroll
      for (int c = 0; c < 3; ++c) px[r][c] = 0;
      if (in_img[r]) {
        const int hb = 2 * pr + r;                                  // row of the band
        const int ymin = s_yb[2 * hb], cnt = s_yb[2 * hb + 1];
        const int32_t* k = s_yk + hb * ky;
        int acc[3] = {1 << (kPrecBits - 1), 1 << (kPrecBits - 1), 1 << (kPrecBits - 1)};
        for (int y = 0; y < cnt; ++y) {
          const int kv = k[y];
          const uint8_t* t = s_tmp + (ymin - ys0 + y) * kOutRowBytes + w * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[c] += static_cast<int>(t[c]) * kv;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) px[r][c] = static_cast<uint8_t>(min(max(acc[c] >> kPrecBits, 0), 255));
      }
    }
    if (out_u8) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int h = 2 * q + r - kStemPad;
        if (in_img[r]) {
          uint8_t* d = out_u8 + ((n * kImg + h) * kImg + w) * 3;
          d[0] = px[r][0]; d[1] = px[r][1]; d[2] = px[r][2];
        }
      }
    }
    if (out_stem) {
      float v[2][3];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) v[r][c] = in_img[r] ? s_lut[c * 256 + px[r][c]] : 0.f;
      if constexpr (FP32) {
        float* o = static_cast<float*>(out_stem);
#pragma unroll
        for (int r = 0; r < 2; ++r)
          if (in_img[r])
            *reinterpret_cast<float4*>(o + rs_stem_offset<true>(n, 2 * q + r, w + kStemPad)) = make_float4(v[r][0], v[r][1], v[r][2], 0.f);
      } else {
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out_stem);
        uint4 pk;
        pk.x = pk_bf16x2(v[0][0], v[0][1]); pk.y = pk_bf16x2(v[0][2], 0.f);
        pk.z = pk_bf16x2(v[1][0], v[1][1]); pk.w = pk_bf16x2(v[1][2], 0.f);
        *reinterpret_cast<uint4*>(o + rs_stem_offset<false>(n, 2 * q, w + kStemPad)) = pk;
      }
    }
  }
}

}  // namespace

int resize_ksize(int in_size, int out_size) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double support = scale < 1.0 ? 1.0 : scale;
  int c = static_cast<int>(support);
  if (static_cast<double>(c) < support) ++c;   // ceil
  return c * 2 + 1;
}

// bounds [out,2] / kk [out, ksize] int32 device arrays for one axis
void launch_resize_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* kk, cudaStream_t s) {
  const int ksize = resize_ksize(in_size, out_size);
  resize_coeffs_kernel<<<(out_size + 127) / 128, 128, 0, s>>>(in_size, out_size, ksize, bounds, kk);
  VCG_CUDA(cudaGetLastError());
}

// frames: uint8 [*, Hs, Ws, 3]; image i of the n outputs reads frame frame_index[i] / clip_start[i / T] + i % T / i
void launch_resize_preprocess_u8(const uint8_t* frames, const int32_t* frame_index, const int32_t* clip_start, int T, int n,
                                 int n_frames, int Hs, int Ws, const int32_t* xb, const int32_t* xk, const int32_t* yb,
                                 const int32_t* yk, void* out_stem, uint8_t* out_u8, bool fp32, cudaStream_t s) {
  if (n == 0) return;
  VCG_REQUIRE(Hs >= 1 && Ws >= 1 && Hs <= 15 * kImg && Ws <= 15 * kImg, "source frame size out of range (<= 15x down-scaling)");
  const int kx = resize_ksize(Ws, kImg), ky = resize_ksize(Hs, kImg);
  // source rows one band of 8 output rows can depend on
  const double sy = static_cast<double>(Hs) / kImg;
  const int max_rows = static_cast<int>(8 * (sy < 1.0 ? 1.0 : sy) + 2 * (sy < 1.0 ? 1.0 : sy) + 4);
  const int src_pitch = ((Ws * 3 + 15) & ~15) + 64;
  const size_t smem = 2 * static_cast<size_t>(src_pitch) + ((static_cast<size_t>(max_rows) * kOutRowBytes + 15) & ~size_t(15)) +
                      768 * sizeof(float) + static_cast<size_t>(kImg) * kx * 4 + kImg * 2 * 4 + 2 * kRsPairs * (ky + 2) * 4;
  VCG_REQUIRE(smem <= 200 * 1024, "source frames too large for the fused resize kernel");
  const unsigned grid = static_cast<unsigned>(n) * kRsChunks;
  auto go = [&](auto kernel, PerDeviceMax& cfg) {
    if (cfg.raise(smem)) VCG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    launch_pdl(kernel, grid, 256, smem, s, frames, frame_index, clip_start, T, n_frames, Hs, Ws, xb, xk, kx, yb, yk, ky, max_rows,
               out_stem, out_u8);
  };
  static PerDeviceMax configured[2][4];
  const int kv = kx <= 3 ? 0 : kx <= 5 ? 1 : kx <= 7 ? 2 : 3;   // taps in registers: 3 / 5 / 7 (down-scaling up to 3x); else generic
  if (fp32) {
    if (kv == 0) go(resize_preprocess_u8_kernel<true, 3>, configured[1][0]);
    else if (kv == 1) go(resize_preprocess_u8_kernel<true, 5>, configured[1][1]);
    else if (kv == 2) go(resize_preprocess_u8_kernel<true, 7>, configured[1][2]);
    else go(resize_preprocess_u8_kernel<true, 0>, configured[1][3]);
  } else {
    if (kv == 0) go(resize_preprocess_u8_kernel<false, 3>, configured[0][0]);
    else if (kv == 1) go(resize_preprocess_u8_kernel<false, 5>, configured[0][1]);
    else if (kv == 2) go(resize_preprocess_u8_kernel<false, 7>, configured[0][2]);
    else go(resize_preprocess_u8_kernel<false, 0>, configured[0][3]);
  }
  VCG_CUDA(cudaGetLastError());
}

}  // namespace vcg
