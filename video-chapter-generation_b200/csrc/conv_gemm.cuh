// Persistent, warp-specialised implicit-GEMM kernel for sm_100a.
//
//   D[M, N] = act( sum_k A[M, k] * W[N, k] + bias[N] (+ residual[M, N]) )
//
// A is an NHWC activation tensor reached through up to four rank-5 TMA tensor maps (C, W, H, plane, image); one
// M-tile is a (bw x bh x nf) patch of output pixels, and the K loop walks (filter tap, channel block) pairs, every
// step being one TMA box load whose origin is the patch origin plus the tap offset.  Out-of-bounds box elements are
// zero-filled by TMA, which is exactly conv zero padding.  Plain GEMMs are the degenerate case bw=128, bh=nf=1,
// one tap.  W is [N, K] K-major.  tcgen05.mma accumulates 128 x BLOCK_N fp32 tiles in TMEM (double buffered), the
// epilogue warps drain them with tcgen05.ld and apply bias / residual / activation, and optionally scatter the
// first C/4 channels into the next bottleneck's temporally shifted input (TSM).
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warp 3 idle,
// warps 4..11 epilogue (two warps per 32-lane TMEM quarter, interleaved over 32-column chunks).
// In the TF32x3 variant (fp32 verification mode) warps 12..15 split every landed fp32 stage into
// (hi, lo) TF32 parts in shared memory and the MMA warp issues hi*hi + hi*lo + lo*hi.
#pragma once
#include "ptx.cuh"
#include <cuda_bf16.h>

namespace vcg {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_TANH = 3 };

struct TapDesc {
  int16_t dw, dh;   // offset of the box origin in W and H (may be negative: zero padding)
  int16_t c_off;    // channel offset inside the tap (elements)
  int8_t map;       // which A tensor map
  int8_t plane;     // coordinate on the "plane" dimension
};

struct ConvGemmParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  TapDesc taps[16];
  int n_taps, cpt;             // taps; channel blocks (of BLOCK_K elements) per tap
  int tsm_split_cb, tsm_map;   // channel blocks below tsm_split_cb are read through a_map[tsm_map]
  int bw, bh, nf;              // patch extents; bw*bh*nf <= 128 rows
  int Wo, Ho, Nimg;            // output geometry
  int tiles_w, tiles_h, tiles_n, n_tiles;
  uint32_t a_bytes, b_bytes;   // bytes one TMA box delivers (for expect_tx)
  int N;                       // output channels
  void* out;                   // [Nimg*Ho*Wo, ld_out]
  int ld_out;
  const float* bias;           // [N] or nullptr
  const void* residual;        // same geometry as out, or nullptr
  int ld_res;
  int act;
  void* tsm_out;               // [Nimg*Ho*Wo, tsm_ld] shifted copy of channels [0, 2*tsm_fold), or nullptr
  int tsm_ld, tsm_fold, T;
};

constexpr int kBlockM = 128;
constexpr int kNumEpiWarps = 8;
constexpr int kFirstEpiWarp = 4;

template <int BLOCK_N, bool TF32X3>
struct ConvGemmCfg {
  static constexpr int kElem = TF32X3 ? 4 : 2;
  static constexpr int kBlockK = 128 / kElem;                 // elements per 128-byte swizzled row
  static constexpr int kUmmaK = 32 / kElem;                   // elements per tcgen05.mma (32 bytes of K)
  static constexpr int kABytes = kBlockM * 128;
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = (kABytes + kBBytes) * (TF32X3 ? 2 : 1);
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  // bf16: two accumulators (tile i+1 accumulates while tile i drains).  TF32x3: the tensor core truncates on every
  // accumulate, so long fp32 sums pick up a bias ~ (#MMAs) * ulp/2; the 512 TMEM columns are used as 512/BLOCK_N
  // partial accumulators instead (number 0 takes the small lo*hi + hi*lo terms, the others take hi*hi round-robin
  // over K blocks) and the epilogue adds them with round-to-nearest fp32 adds.
  static constexpr int kNumAcc = TF32X3 ? 512 / BLOCK_N : 2;
  static constexpr int kTmemCols = TF32X3 ? 512 : (2 * BLOCK_N <= 32 ? 32 : 2 * BLOCK_N <= 64 ? 64 : 2 * BLOCK_N <= 128 ? 128
                                   : 2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int kThreads = TF32X3 ? 512 : 384;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

template <int BLOCK_N, bool TF32X3>
__global__ void __launch_bounds__(ConvGemmCfg<BLOCK_N, TF32X3>::kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BLOCK_N, TF32X3>;
  using OutT = typename std::conditional<TF32X3, float, __nv_bfloat16>::type;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                  // [stage][128 rows][128 B]
  uint8_t* sB = sA + kStages * Cfg::kABytes;           // [stage][BLOCK_N rows][128 B]
  uint8_t* sA_lo = sB + kStages * Cfg::kBBytes;        // TF32X3 only
  uint8_t* sB_lo = sA_lo + kStages * Cfg::kABytes;     // TF32X3 only
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                  // TMA -> (splitter | MMA)
  uint64_t* empty_bar = bars + kStages;       // MMA -> TMA
  uint64_t* split_bar = bars + 2 * kStages;   // splitter -> MMA (TF32X3)
  uint64_t* tmem_full = bars + 3 * kStages;   // MMA -> epilogue   [2]
  uint64_t* tmem_empty = tmem_full + 2;       // epilogue -> MMA   [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kNumEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.n_taps * p.cpt;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = m_tiles * p.n_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.n_tiles, n_blk = tile - m_blk * p.n_tiles;
        const int iw = m_blk % p.tiles_w;
        const int ih = (m_blk / p.tiles_w) % p.tiles_h;
        const int in = m_blk / (p.tiles_w * p.tiles_h);
        const int w0 = iw * p.bw, h0 = ih * p.bh, n0 = in * p.nf;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
          const TapDesc t = p.taps[tap];
          const int map = (cb < p.tsm_split_cb) ? p.tsm_map : t.map;
          tma_load_5d(sA + stage * Cfg::kABytes, &p.a_map[map], &full_bar[stage], cb * Cfg::kBlockK + t.c_off,
                      w0 + t.dw, h0 + t.dh, t.plane, n0);
          tma_load_2d(sB + stage * Cfg::kBBytes, &p.b_map, &full_bar[stage], kb * Cfg::kBlockK, n_blk * BLOCK_N);
          if (++cb == p.cpt) { cb = 0; ++tap; }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(TF32X3 ? 2u : 1u, kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = TF32X3 ? 0 : (it & 1);
        mbar_wait(&tmem_empty[acc], (TF32X3 ? (it & 1) : ((it >> 1) & 1)) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(TF32X3 ? &split_bar[stage] : &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < Cfg::kBlockK / Cfg::kUmmaK; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * 32);
            const uint64_t db = umma_desc_sw128(b_addr + k * 32);
            if constexpr (TF32X3) {
              const uint64_t da_lo = umma_desc_sw128(smem_u32(sA_lo + stage * Cfg::kABytes) + k * 32);
              const uint64_t db_lo = umma_desc_sw128(smem_u32(sB_lo + stage * Cfg::kBBytes) + k * 32);
              constexpr int kHiAcc = Cfg::kNumAcc - 1;
              const uint32_t d_hi = tmem_base + (1 + kb % kHiAcc) * BLOCK_N;
              umma_tf32(tmem_base, da_lo, db, idesc, (kb | k) != 0);      // accumulator 0: correction terms
              umma_tf32(tmem_base, da, db_lo, idesc, 1);
              umma_tf32(d_hi, da, db, idesc, (kb >= kHiAcc) || (k != 0));  // hi*hi partial sums
            } else {
              umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp >= kFirstEpiWarp && warp < kFirstEpiWarp + kNumEpiWarps) {
    // ------------------------------------------------------------ epilogue
    const int quarter = warp & 3;                       // TMEM lanes [32*quarter, +32)
    const int half = (warp - kFirstEpiWarp) >> 2;       // which 32-column chunks (even / odd)
    const int row = quarter * 32 + lane;
    const int dw = row % p.bw;
    const int dh = (row / p.bw) % p.bh;
    const int dn = row / (p.bw * p.bh);
    const int HW = p.Ho * p.Wo;
    OutT* out = reinterpret_cast<OutT*>(p.out);
    const OutT* res = reinterpret_cast<const OutT*>(p.residual);
    OutT* tsm = reinterpret_cast<OutT*>(p.tsm_out);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / p.n_tiles, n_blk = tile - m_blk * p.n_tiles;
      const int iw = m_blk % p.tiles_w;
      const int ih = (m_blk / p.tiles_w) % p.tiles_h;
      const int in = m_blk / (p.tiles_w * p.tiles_h);
      const int w = iw * p.bw + dw, h = ih * p.bh + dh, n = in * p.nf + dn;
      const bool row_ok = (dn < p.nf) && (w < p.Wo) && (h < p.Ho) && (n < p.Nimg);
      const long grow = (static_cast<long>(n) * p.Ho + h) * p.Wo + w;
      const int acc = TF32X3 ? 0 : (it & 1);
      mbar_wait(&tmem_full[acc], TF32X3 ? (it & 1) : ((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int chunk = half; chunk < BLOCK_N / 32; chunk += 2) {
        uint32_t r[32];
        float v[32];
        if constexpr (TF32X3) {
          // sum the hi*hi partial accumulators (round-to-nearest adds), then the correction accumulator
          const int n_hi = min(Cfg::kNumAcc - 1, num_kb);
          tmem_ld_32x32(taddr + BLOCK_N + chunk * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          for (int a = 2; a <= n_hi; ++a) {
            tmem_ld_32x32(taddr + a * BLOCK_N + chunk * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
          }
          tmem_ld_32x32(taddr + chunk * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
        } else {
          tmem_ld_32x32(taddr + chunk * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
        const int col0 = n_blk * BLOCK_N + chunk * 32;
        if (row_ok && col0 < p.N) {
          const int ncols = min(32, p.N - col0);   // multiple of 8 by contract
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < ncols) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
          }
          if (res) {
            const OutT* rp = res + grow * p.ld_res + col0;
            if constexpr (TF32X3) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (j < ncols) {
                  const float4 q = *reinterpret_cast<const float4*>(rp + j);
                  v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (j < ncols) {
                  const uint4 q = *reinterpret_cast<const uint4*>(rp + j);
                  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    v[j + 2 * e] += f.x; v[j + 2 * e + 1] += f.y;
                  }
                }
              }
            }
          }
          if (p.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
          }
          // destination rows: the output itself, plus (TSM) the neighbouring frame's slot in the shifted buffer
          OutT* dst = out + grow * p.ld_out + col0;
          OutT* dst2 = nullptr;
          if (tsm && col0 < 2 * p.tsm_fold) {
            const int t = static_cast<int>((grow / HW) % p.T);
            if (col0 < p.tsm_fold) {             // out[t-1, c] = x[t, c]   (shift left in time)
              if (t >= 1) dst2 = tsm + (grow - HW) * p.tsm_ld + col0;
            } else {                             // out[t+1, c] = x[t, c]   (shift right in time)
              if (t + 1 < p.T) dst2 = tsm + (grow + HW) * p.tsm_ld + col0;
            }
          }
          if constexpr (TF32X3) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < ncols) {
                const float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                *reinterpret_cast<float4*>(dst + j) = o;
                if (dst2) *reinterpret_cast<float4*>(dst2 + j) = o;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (j < ncols) {
                uint4 o;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[j + 2 * e], v[j + 2 * e + 1]);
                *reinterpret_cast<uint4*>(dst + j) = o;
                if (dst2) *reinterpret_cast<uint4*>(dst2 + j) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  } else if (TF32X3 && warp >= 12) {
    // ------------------------------------------------------------ TF32 hi/lo splitter (fp32 verification mode)
    // hi keeps the top 19 bits (exactly representable in TF32); lo = x - hi is exact in fp32.
    const int tid = threadIdx.x - 12 * 32;   // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        float4* a_hi = reinterpret_cast<float4*>(sA + stage * Cfg::kABytes);
        float4* a_lo = reinterpret_cast<float4*>(sA_lo + stage * Cfg::kABytes);
        float4* b_hi = reinterpret_cast<float4*>(sB + stage * Cfg::kBBytes);
        float4* b_lo = reinterpret_cast<float4*>(sB_lo + stage * Cfg::kBBytes);
        auto split4 = [](float4* hi, float4* lo, int i) {
          float4 x = hi[i];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
          hi[i] = h;
          lo[i] = l;
        };
        for (int i = tid; i < Cfg::kABytes / 16; i += 128) split4(a_hi, a_lo, i);
        for (int i = tid; i < Cfg::kBBytes / 16; i += 128) split4(b_hi, b_lo, i);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace vcg
