// Persistent, warp-specialised implicit-GEMM kernel for sm_100a.
//
//   D[M, N] = act( sum_k A[M, k] * W[N, k] + bias[N] (+ residual[M, N]) )
//
// A is an NHWC activation tensor reached through up to four rank-5 TMA tensor maps (C, W, H, plane, image); one
// M-tile is a (bw x bh x nf) patch of output pixels, and the K loop walks (filter tap, channel block) pairs, every
// step being one TMA box load whose origin is the patch origin plus the tap offset.  Out-of-bounds box elements are
// zero-filled by TMA, which is exactly conv zero padding.  Plain GEMMs are the degenerate case bw=128, bh=nf=1,
// one tap.  W is [N, K] K-major.  tcgen05.mma accumulates 128 x BLOCK_N fp32 tiles in TMEM.
//
// bf16 variant (the product path), 640 threads:
//   warp 0      TMA producer of the A/B stage ring
//   warp 1      MMA issuer (one elected lane), two TMEM accumulators so tile i+1 runs under tile i's epilogue
//   warp 2      TMEM allocator
//   warp 3      "C" producer: a ring of four 128x64 bf16 smem tiles; for layers with a residual it prefetches the
//               residual tile by TMA (the residual layers are HBM-bound: the ring keeps 64 KB per SM in flight)
//   warps 4-19  epilogue, four groups of four warps (one per TMEM lane quarter) that split the tile's columns and run
//               concurrently: tcgen05.ld -> +bias (+residual from smem) -> activation -> bf16 into the same smem tile
//               (TMA 128-byte swizzle, conflict-free) -> one TMA store per 128x64 tile; the first C/4 channels are
//               additionally scattered into the next bottleneck's temporally shifted input (TSM).
// TF32x3 variant (fp32 verification mode), 512 threads: warps 12-15 split every landed fp32 stage into (hi, lo) TF32
// parts in shared memory, the MMA warp issues hi*hi + hi*lo + lo*hi into several partial accumulators, and the
// epilogue uses plain row-per-lane global accesses.
#pragma once
#include "ptx.cuh"
#include "launch.cuh"
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
  CUtensorMap out_map;         // bf16 path: output [Nimg, Ho, Wo, ld_out] as (C, W, H, 1, N), box (64, bw, bh, 1, nf)
  CUtensorMap res_map;         // bf16 path: residual, same geometry
  TapDesc taps[16];
  int n_taps, cpt;             // taps; channel blocks (of BLOCK_K elements) per tap
  int tsm_split_cb, tsm_map;   // channel blocks below tsm_split_cb are read through a_map[tsm_map]
  int n_stages, n_cslots;      // bf16 path: split of the smem budget between the A/B stage ring and the C-tile ring
  const int* m_dev;            // plain GEMM only: number of valid rows lives on the device (token-packed BERT)
  // "clip view" of a per-unique-frame tensor (frames shared by overlapping clips): map dims (C, W, H, T, clip) with
  // box (.., nf, 1); image index n = clip*T + t.  0 = ordinary (C, W, H, plane, image) addressing.
  int a_clip_T, res_clip_T;
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
  // ---- LayerNorm folded into the GEMMs around it (bf16 plain GEMMs, LNF instantiations; see "LayerNorm" in DESIGN.md)
  // row statistics are (sum, sum of squares) over the ln_dim columns of a pre-LayerNorm matrix
  const float2* a_stats;       // A is a raw pre-LN matrix, W was packed as W*gamma: D = rstd*(acc - mean*ln_c1) + bias'
  const float* ln_c1;          // [N] sum_k W'[n,k]  (bias' = bias + sum_k beta_k W[n,k] arrives through `bias`)
  const float2* res_stats;     // the residual is a raw pre-LN matrix: residual = (t - mean)*rstd*res_gamma + res_beta
  const float* res_gamma;      // [N]
  const float* res_beta;       // [N]
  float2* out_stats;           // accumulate (sum, sum of squares) of the bf16-rounded outputs per row (atomics)
  float ln_inv_dim, ln_eps;
};

constexpr int kBlockM = 128;
constexpr int kFirstEpiWarp = 4;
constexpr int kEpiWarpsBf16 = 16;        // bf16: four warps per TMEM lane quarter, 16 columns of a 64-column C tile each
constexpr int kEpiWarpsTf32 = 8;
constexpr int kMaxStages = 8;
constexpr int kMaxCSlots = 12;           // smem C-tile ring (bf16 path); the host picks n_stages / n_cslots per layer
constexpr int kMinCSlots = 4;
constexpr int kCBytes = kBlockM * 128;   // one 128 x 64 bf16 tile

template <int BLOCK_N, bool TF32X3>
struct ConvGemmCfg {
  static constexpr int kElem = TF32X3 ? 4 : 2;
  static constexpr int kBlockK = 128 / kElem;                 // elements per 128-byte swizzled row
  static constexpr int kUmmaK = 32 / kElem;                   // elements per tcgen05.mma (32 bytes of K)
  static constexpr int kABytes = kBlockM * 128;
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = (kABytes + kBBytes) * (TF32X3 ? 2 : 1);
  static constexpr int kCRingBytes = TF32X3 ? 0 : kMinCSlots * kCBytes;
  static constexpr int kBudget = 224 * 1024;   // stage ring + C ring
  // default split: as many stages as fit next to the minimum C ring
  static constexpr int kStages = (kBudget - kCRingBytes) / kStageBytes > kMaxStages ? kMaxStages : (kBudget - kCRingBytes) / kStageBytes;
  // bf16: two accumulators (tile i+1 accumulates while tile i drains).  TF32x3: the tensor core truncates on every
  // accumulate, so long fp32 sums pick up a bias ~ (#MMAs) * ulp/2; the 512 TMEM columns are used as 512/BLOCK_N
  // partial accumulators instead (number 0 takes the small lo*hi + hi*lo terms, the others take hi*hi round-robin
  // over K blocks) and the epilogue adds them with round-to-nearest fp32 adds.
  static constexpr int kNumAcc = TF32X3 ? 512 / BLOCK_N : 2;
  static constexpr int kTmemCols = TF32X3 ? 512 : (2 * BLOCK_N <= 32 ? 32 : 2 * BLOCK_N <= 64 ? 64 : 2 * BLOCK_N <= 128 ? 128
                                   : 2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int kNumEpiWarps = TF32X3 ? kEpiWarpsTf32 : kEpiWarpsBf16;
  static constexpr int kThreads = TF32X3 ? 512 : (kFirstEpiWarp + kEpiWarpsBf16) * 32;
  static constexpr int kSmemBytes = kBudget + 1024 /*align*/ + 512 /*barriers*/;
  static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
  static_assert(kStages >= 2, "need at least two pipeline stages");
};

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): one MUFU.RCP, one MUFU.EX2 and a 5-term Horner chain — a
// third of the instructions of erff(), which matters because the FFN-in epilogue is ALU-bound.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-ax * ax);
  return copysignf(e, x);
}

// Activation over a register tile; the switch is hoisted out of the element loop.
template <bool EXACT>
__device__ __forceinline__ void apply_act32(float (&v)[32], int act) {
  if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if constexpr (EXACT) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752440f));
      else v[j] = 0.5f * v[j] * (1.f + erf_fast(v[j] * 0.70710678118654752440f));
    }
  } else if (act == ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
  }
}

// GELU for the bf16 epilogue: x * Phi(x) with Phi(x) - 1/2 = u * P(u^2), u = clamp(x, +-3.75), P a degree-6 minimax
// polynomial (max |error| of the GELU value 2.5e-4, an eighth of a bf16 ulp at 1; the fp32 verification mode uses
// erff).  No MUFU, and two elements per instruction on the packed fp32x2 pipe (FMUL2 / FFMA2): the FFN-in epilogue is
// issue-bound, the A&S erf it replaces cost 13 FMA-pipe + 2 MUFU instructions per element.
__device__ __forceinline__ float2 gelu_poly2(float2 x) {
  constexpr float c = 3.75f;
  const float2 u = make_float2(fminf(fmaxf(x.x, -c), c), fminf(fmaxf(x.y, -c), c));
  const float2 t = __fmul2_rn(u, u);
  float2 p = __ffma2_rn(t, make_float2(3.419050747872096e-08f, 3.419050747872096e-08f),
                        make_float2(-2.1325091891693936e-06f, -2.1325091891693936e-06f));
  p = __ffma2_rn(p, t, make_float2(5.764379563294139e-05f, 5.764379563294139e-05f));
  p = __ffma2_rn(p, t, make_float2(-0.0008995933840409692f, -0.0008995933840409692f));
  p = __ffma2_rn(p, t, make_float2(0.009149351282721695f, 0.009149351282721695f));
  p = __ffma2_rn(p, t, make_float2(-0.06531965073372963f, -0.06531965073372963f));
  p = __ffma2_rn(p, t, make_float2(0.3983374174621377f, 0.3983374174621377f));
  return __fmul2_rn(x, __ffma2_rn(u, p, make_float2(0.5f, 0.5f)));
}

// Activation over eight fp32 pairs (the bf16 epilogue works on packed pairs throughout)
__device__ __forceinline__ void apply_act8x2(float2 (&v)[8], int act) {
  if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = make_float2(fmaxf(v[j].x, 0.f), fmaxf(v[j].y, 0.f));
  } else if (act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gelu_poly2(v[j]);
  } else if (act == ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = make_float2(tanhf(v[j].x), tanhf(v[j].y));
  }
}

// shared-memory accesses through 32-bit shared-window addresses (the generic-pointer forms cost a 64-bit address chain each)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// two fp32 -> packed bf16x2 (round to nearest even), and the same followed by ReLU on the packed pair
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("{\n\t.reg .b32 t;\n\tcvt.rn.bf16x2.f32 t, %2, %1;\n\tmax.bf16x2 %0, t, %3;\n\t}" : "=r"(r) : "f"(lo), "f"(hi), "r"(0u));
  return r;
}
// bf16x2 -> two fp32 (exact)
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
}

// CG2 = CTA-pair mode (bf16 only): the two CTAs of a (2,1,1) cluster compute one 256 x BLOCK_N output tile with
// tcgen05.mma.cta_group::2.  Each CTA loads its own 128-row A patch and HALF of the B tile (BLOCK_N/2 weight rows), the
// leader CTA's MMA thread issues for both, and every CTA drains the 128 accumulator rows that live in its own TMEM.
// Per SM and K block that is 16 KB + BLOCK_N/2*128 B from L2 instead of 16 KB + BLOCK_N*128 B: the single-CTA kernel
// is pinned at the ~11 TB/s L2->SM limit (48 KB per 128x256x64 MACs), the pair needs a third less.
//   full_bar      lives in the leader; both CTAs' TMA loads complete_tx on it (leader expects the bytes of both)
//   empty_bar     one per CTA, signalled by the leader's multicast tcgen05.commit
//   tmem_full     one per CTA, multicast commit
//   tmem_empty    lives in the leader; the epilogue warps of both CTAs arrive on it (remote arrive from the peer)
template <int BLOCK_N, bool TF32X3, bool CG2, bool LNF = false>
__device__ __forceinline__ void conv_gemm_body(const ConvGemmParams& p) {
  using Cfg = ConvGemmCfg<BLOCK_N, TF32X3>;
  static_assert(!(CG2 && TF32X3), "CTA-pair mode is bf16 only");
  const int kStages = TF32X3 ? Cfg::kStages : p.n_stages;
  const int kCSlots = TF32X3 ? kMinCSlots : p.n_cslots;
  constexpr int kBRows = CG2 ? BLOCK_N / 2 : BLOCK_N;       // weight rows this CTA stages per K block
  constexpr int kBBytes = kBRows * 128;
  constexpr int kStageBytes = (Cfg::kABytes + kBBytes) * (TF32X3 ? 2 : 1);
  const uint32_t cta_rank = CG2 ? cluster_ctarank() : 0u;   // 0 = leader
  // tile walk: a "unit" is one CTA (one M tile) or one CTA pair (two adjacent M tiles)
  const int unit = CG2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_units = CG2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                  // [stage][128 rows][128 B]
  uint8_t* sB = sA + kStages * Cfg::kABytes;           // [stage][kBRows rows][128 B]
  uint8_t* sA_lo = sB + kStages * kBBytes;             // TF32X3 only
  uint8_t* sB_lo = sA_lo + kStages * Cfg::kABytes;     // TF32X3 only
  uint8_t* sC = smem + kStages * kStageBytes;          // bf16 only: [kCSlots][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBudget);
  uint64_t* full_bar = bars;                     // TMA -> (splitter | MMA)
  uint64_t* empty_bar = bars + kMaxStages;       // MMA -> TMA
  uint64_t* split_bar = bars + 2 * kMaxStages;   // splitter -> MMA (TF32X3)
  uint64_t* tmem_full = bars + 3 * kMaxStages;   // MMA -> epilogue   [2]
  uint64_t* tmem_empty = tmem_full + 2;          // epilogue -> MMA   [2]
  uint64_t* c_full = tmem_empty + 2;             // C producer -> epilogue [kCSlots]
  uint64_t* c_empty = c_full + kMaxCSlots;       // epilogue -> C producer [kCSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_empty + kMaxCSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();   // the next kernel's prologue may overlap this kernel's tail (launch.cuh)
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
    if (!TF32X3) {
      tma_prefetch_desc(&p.out_map);
      tma_prefetch_desc(&p.res_map);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], Cfg::kNumEpiWarps * (CG2 ? 2 : 1));
    }
    for (int i = 0; i < kCSlots; ++i) {
      mbar_init(&c_full[i], 1);
      mbar_init(&c_empty[i], TF32X3 ? 1 : (4 / (BLOCK_N / 64)) * 4);   // warps sharing one C tile
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CG2) {
      tmem_alloc_cg2(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG2) cluster_sync(); else __syncthreads();   // barrier inits visible to the peer before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                // everything above touched only shared / tensor memory and kernel parameters

  const int num_kb = p.n_taps * p.cpt;
  // rows beyond *m_dev are never needed: only the M tiles that hold valid rows are computed
  const int m_tiles = p.m_dev ? (min(ld_chain_i32(p.m_dev), p.Wo) + kBlockM - 1) / kBlockM : p.tiles_w * p.tiles_h * p.tiles_n;
  // CG2: a unit tile is (pair of M tiles, N tile); the odd CTA of a last, half-empty pair works on an M tile past the
  // end (its image coordinate is out of range: TMA zero-fills the loads and clips the stores)
  const int total_tiles = (CG2 ? (m_tiles + 1) / 2 : m_tiles) * p.n_tiles;
  auto m_of = [&](int tile) { return CG2 ? 2 * (tile / p.n_tiles) + static_cast<int>(cta_rank) : tile / p.n_tiles; };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (A/B stages)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // CG2: the bytes of both CTAs land on the leader's full barrier
      const uint32_t full0 = CG2 ? mapa_cluster(smem_u32(full_bar), 0) : smem_u32(full_bar);
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        const int m_blk = m_of(tile), n_blk = tile % p.n_tiles;
        const int iw = m_blk % p.tiles_w;
        const int ih = (m_blk / p.tiles_w) % p.tiles_h;
        const int in = m_blk / (p.tiles_w * p.tiles_h);
        const int w0 = iw * p.bw, h0 = ih * p.bh, n0 = in * p.nf;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (!CG2 || cta_rank == 0) mbar_expect_tx(&full_bar[stage], (p.a_bytes + p.b_bytes) * (CG2 ? 2u : 1u));
          const uint32_t bar = full0 + stage * 8;
          const TapDesc t = p.taps[tap];
          const int map = (cb < p.tsm_split_cb) ? p.tsm_map : t.map;
          if (p.a_clip_T == 0)
            tma_load_5d<CG2>(sA + stage * Cfg::kABytes, &p.a_map[map], bar, cb * Cfg::kBlockK + t.c_off,
                             w0 + t.dw, h0 + t.dh, t.plane, n0);
          else   // plane = temporal tap offset (frame t-1 / t / t+1 of the same clip; out of range -> zero fill)
            tma_load_5d<CG2>(sA + stage * Cfg::kABytes, &p.a_map[map], bar, cb * Cfg::kBlockK + t.c_off,
                             w0 + t.dw, h0 + t.dh, n0 % p.a_clip_T + t.plane, n0 / p.a_clip_T);
          tma_load_2d<CG2>(sB + stage * kBBytes, &p.b_map, bar, kb * Cfg::kBlockK,
                           n_blk * BLOCK_N + static_cast<int>(cta_rank) * kBRows);
          if (++cb == p.cpt) { cb = 0; ++tap; }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (CG2: the leader issues for the pair)
    if ((!CG2 || cta_rank == 0) && elect_one()) {
      constexpr uint32_t idesc = umma_idesc(TF32X3 ? 2u : 1u, CG2 ? 2 * kBlockM : kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = unit; tile < total_tiles; tile += n_units, ++it) {
        const int acc = TF32X3 ? 0 : (it & 1);
        mbar_wait(&tmem_empty[acc], (TF32X3 ? (it & 1) : ((it >> 1) & 1)) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(TF32X3 ? &split_bar[stage] : &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * kBBytes);
#pragma unroll
          for (int k = 0; k < Cfg::kBlockK / Cfg::kUmmaK; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * 32);
            const uint64_t db = umma_desc_sw128(b_addr + k * 32);
            if constexpr (TF32X3) {
              const uint64_t da_lo = umma_desc_sw128(smem_u32(sA_lo + stage * Cfg::kABytes) + k * 32);
              const uint64_t db_lo = umma_desc_sw128(smem_u32(sB_lo + stage * kBBytes) + k * 32);
              constexpr int kHiAcc = Cfg::kNumAcc - 1;
              const uint32_t d_hi = tmem_base + (1 + kb % kHiAcc) * BLOCK_N;
              umma_tf32(tmem_base, da_lo, db, idesc, (kb | k) != 0);      // accumulator 0: correction terms
              umma_tf32(tmem_base, da, db_lo, idesc, 1);
              umma_tf32(d_hi, da, db, idesc, (kb >= kHiAcc) || (k != 0));  // hi*hi partial sums
            } else if constexpr (CG2) {
              umma_bf16_cg2(d_tmem, da, db, idesc, (kb | k) != 0);
            } else {
              umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
            }
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if constexpr (CG2) umma_commit_cg2(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG2) umma_commit_cg2(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
      }
    }
  } else if (!TF32X3 && warp == 3) {
    // ------------------------------------------------------------ C producer (bf16): residual prefetch ring
    if (elect_one()) {
      const bool has_res = p.residual != nullptr;
      int c_it = 0;
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        const int m_blk = m_of(tile), n_blk = tile % p.n_tiles;
        const int iw = m_blk % p.tiles_w;
        const int ih = (m_blk / p.tiles_w) % p.tiles_h;
        const int in = m_blk / (p.tiles_w * p.tiles_h);
        const int n_sub = min(BLOCK_N / 64, (p.N - n_blk * BLOCK_N + 63) / 64);
        // EVERY tile takes BLOCK_N / 64 consecutive ring positions, and the host makes the ring a multiple of that: slot s
        // is then always consumed by the same epilogue group, in tile order.  (With a ring of 5 slots and 4 groups a fast
        // group could wait for use u of a slot whose use u-1 - another group's, one tile back - had not completed yet; the
        // parity wait then returns at once and the group works on the wrong residual: a rare, timing-dependent glitch of a
        // few rows of layer3's conv3, found with tools/stress_checksums.py.)
        for (int j = 0; j < BLOCK_N / 64; ++j, ++c_it) {
          const int slot = c_it % kCSlots;
          mbar_wait(&c_empty[slot], ((c_it / kCSlots) & 1) ^ 1);
          if (has_res && j < n_sub) {
            mbar_expect_tx(&c_full[slot], p.a_bytes);   // same box extents as the A patch: rows x 128 B
            const int n0 = in * p.nf;
            if (p.res_clip_T == 0)
              tma_load_5d(sC + slot * kCBytes, &p.res_map, &c_full[slot], n_blk * BLOCK_N + j * 64, iw * p.bw, ih * p.bh,
                          0, n0);
            else
              tma_load_5d(sC + slot * kCBytes, &p.res_map, &c_full[slot], n_blk * BLOCK_N + j * 64, iw * p.bw, ih * p.bh,
                          n0 % p.res_clip_T, n0 / p.res_clip_T);
          } else {
            mbar_arrive(&c_full[slot]);
          }
        }
      }
    }
  } else if (warp >= kFirstEpiWarp && warp < kFirstEpiWarp + Cfg::kNumEpiWarps) {
    // ------------------------------------------------------------ epilogue
    const int quarter = warp & 3;                       // TMEM lanes [32*quarter, +32)
    const int half = (warp - kFirstEpiWarp) >> 2;       // TF32x3: even/odd 32-column chunks; bf16: 16-column group 0..3
    const int row = quarter * 32 + lane;
    const int dw = row % p.bw;
    const int dh = (row / p.bw) % p.bh;
    const int dn = row / (p.bw * p.bh);
    const int HW = p.Ho * p.Wo;
    int it = 0;
    [[maybe_unused]] int c_it = 0;
    for (int tile = unit; tile < total_tiles; tile += n_units, ++it) {
      const int m_blk = m_of(tile), n_blk = tile % p.n_tiles;
      const int iw = m_blk % p.tiles_w;
      const int ih = (m_blk / p.tiles_w) % p.tiles_h;
      const int in = m_blk / (p.tiles_w * p.tiles_h);
      const int w = iw * p.bw + dw, h = ih * p.bh + dh, n = in * p.nf + dn;
      const bool row_ok = (dn < p.nf) && (w < p.Wo) && (h < p.Ho) && (n < p.Nimg);
      const long grow = row_ok ? (static_cast<long>(n) * p.Ho + h) * p.Wo + w : -1;
      const int acc = TF32X3 ? 0 : (it & 1);
      mbar_wait(&tmem_full[acc], TF32X3 ? (it & 1) : ((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      [[maybe_unused]] bool released_any = false;   // bf16: this warp already handed its TMEM buffer back

      if constexpr (TF32X3) {
        // ---- fp32 verification path: direct row-per-lane accesses
        float* out = reinterpret_cast<float*>(p.out);
        const float* res = reinterpret_cast<const float*>(p.residual);
        float* tsm = reinterpret_cast<float*>(p.tsm_out);
#pragma unroll 1
        for (int chunk = half; chunk < BLOCK_N / 32; chunk += 2) {
          uint32_t r[32];
          float v[32];
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
              const float* rp = res + grow * p.ld_res + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (j < ncols) {
                  const float4 q = *reinterpret_cast<const float4*>(rp + j);
                  v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                }
              }
            }
            apply_act32<true>(v, p.act);
            float* dst = out + grow * p.ld_out + col0;
            float* dst2 = nullptr;
            if (tsm && col0 < 2 * p.tsm_fold) {
              const int t = static_cast<int>((grow / HW) % p.T);
              if (col0 < p.tsm_fold) {             // out[t-1, c] = x[t, c]   (shift left in time)
                if (t >= 1) dst2 = tsm + (grow - HW) * p.tsm_ld + col0;
              } else {                             // out[t+1, c] = x[t, c]   (shift right in time)
                if (t + 1 < p.T) dst2 = tsm + (grow + HW) * p.tsm_ld + col0;
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (j < ncols) {
                const float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                *reinterpret_cast<float4*>(dst + j) = o;
                if (dst2) *reinterpret_cast<float4*>(dst2 + j) = o;
              }
            }
          }
        }
      } else {
        // ---- bf16 path: 128x64 C tiles through the smem ring, TMA store.  The 16 epilogue warps form four groups of
        //      four (one warp per TMEM lane quarter); the groups split the tile's BLOCK_N columns, so with
        //      BLOCK_N = 256 the four C tiles of an output tile drain concurrently (each chain of
        //      tcgen05.ld -> math -> st.shared -> fence -> barrier -> TMA store is latency-bound).
        constexpr int kSub = BLOCK_N / 64;          // C tiles (ring slots) per output tile
        constexpr int kGps = 4 / kSub;              // groups sharing one C tile (BLOCK_N = 192: three groups work, one idles)
        constexpr int kCg = 64 / kGps;              // columns per group: 64, 32 or 16
        const int group = half;                     // (warp - kFirstEpiWarp) >> 2
        const int my_sub = group / kGps;
        const int col_in_sub = (group % kGps) * kCg;
        const bool storer = (quarter == 0) && (group % kGps == 0);
        __nv_bfloat16* tsm = reinterpret_cast<__nv_bfloat16*>(p.tsm_out);
        const bool has_res = p.residual != nullptr;
        const int srow = lane >> 1, spiece = lane & 1;    // TSM scatter: rows srow + 16*i of the warp, 16-byte piece
        long g_i[2];
        int t_i[2];
        if (tsm) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            g_i[i] = __shfl_sync(0xffffffffu, grow, srow + 16 * i);
            t_i[i] = g_i[i] >= 0 ? static_cast<int>((g_i[i] / HW) % p.T) : 0;
          }
        }
        const int n_sub = min(kSub, (p.N - n_blk * BLOCK_N + 63) / 64);
        [[maybe_unused]] float mean_a = 0.f, rstd_a = 1.f, mean_r = 0.f, rstd_r = 1.f;
        [[maybe_unused]] float2 st_sum = make_float2(0.f, 0.f), st_sq = make_float2(0.f, 0.f);
        if constexpr (LNF) {
          if (my_sub < n_sub && row_ok) {
            if (p.a_stats) {
              const float2 st = __ldg(p.a_stats + grow);
              mean_a = st.x * p.ln_inv_dim;
              rstd_a = rsqrtf(fmaxf(st.y * p.ln_inv_dim - mean_a * mean_a, 0.f) + p.ln_eps);
            }
            if (p.res_stats) {
              const float2 st = __ldg(p.res_stats + grow);
              mean_r = st.x * p.ln_inv_dim;
              rstd_r = rsqrtf(fmaxf(st.y * p.ln_inv_dim - mean_r * mean_r, 0.f) + p.ln_eps);
            }
          }
        }
        if (my_sub < n_sub) {
          const int my_it = c_it + my_sub;
          const int slot = my_it % kCSlots;
          uint8_t* ctile = sC + slot * kCBytes;
          const uint32_t ctile_s = smem_u32(ctile);
          const uint32_t crow_s = ctile_s + row * 128;
          bool relu_fast = p.act == ACT_RELU;
          if constexpr (LNF) relu_fast = false;             // (the row statistics are taken from the activated fp32 values)
          // software-pipelined over the 16-column passes: the tcgen05.ld of pass p+1 is in flight during the math of
          // pass p (the first one during the wait for the C slot), and the TMEM buffer goes back to the MMA warp as soon
          // as the last load has landed
          constexpr int kPasses = kCg / 16;
          uint32_t rr[2][16];
          tmem_ld_32x16(taddr + my_sub * 64 + col_in_sub, rr[0]);
          mbar_wait(&c_full[slot], (my_it / kCSlots) & 1);    // residual landed / slot free
#pragma unroll
          for (int pass = 0; pass < kPasses; ++pass) {
            const int cs = col_in_sub + pass * 16;            // first column inside the C tile
            const int col0 = n_blk * BLOCK_N + my_sub * 64 + cs;
            tmem_ld_wait();
            if (pass + 1 < kPasses) {
              tmem_ld_32x16(taddr + my_sub * 64 + cs + 16, rr[(pass + 1) & 1]);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (CG2) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
              }
              released_any = true;
            }
            const uint32_t (&r)[16] = rr[pass & 1];
            float2 v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
            if (col0 < p.N) {
              if constexpr (LNF) {
                if (p.a_stats) {     // LayerNorm of the A rows, applied after the contraction
                  const float2 nm = make_float2(-mean_a, -mean_a), rs = make_float2(rstd_a, rstd_a);
#pragma unroll
                  for (int e = 0; e < 8; e += 2) {
                    const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.ln_c1 + col0 + 2 * e));
                    v[e] = __fmul2_rn(__ffma2_rn(nm, make_float2(c4.x, c4.y), v[e]), rs);
                    v[e + 1] = __fmul2_rn(__ffma2_rn(nm, make_float2(c4.z, c4.w), v[e + 1]), rs);
                  }
                }
              }
              if (p.bias) {
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 2 * e));   // N % 32 == 0 (host check)
                  v[e] = __fadd2_rn(v[e], make_float2(b4.x, b4.y));
                  v[e + 1] = __fadd2_rn(v[e + 1], make_float2(b4.z, b4.w));
                }
              }
              if (has_res) {
                bool ln_res = false;
                if constexpr (LNF) ln_res = p.res_stats != nullptr;
                if (ln_res) {        // residual = LayerNorm(raw tile) recomputed from the row statistics
                  const float2 nm = make_float2(-mean_r, -mean_r), rs = make_float2(rstd_r, rstd_r);
#pragma unroll
                  for (int c = 0; c < 2; ++c) {
                    const uint4 q = lds128(crow_s + ((((cs >> 3) + c) ^ (row & 7)) << 4));
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.res_gamma + col0 + c * 8));
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.res_gamma + col0 + c * 8 + 4));
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.res_beta + col0 + c * 8));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.res_beta + col0 + c * 8 + 4));
                    const float2 gg[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y),
                                          make_float2(g1.z, g1.w)};
                    const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                          make_float2(b1.z, b1.w)};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 xn = __fmul2_rn(__fadd2_rn(__bfloat1622float2(h2[e]), nm), rs);
                      v[c * 4 + e] = __fadd2_rn(v[c * 4 + e], __ffma2_rn(xn, gg[e], bb[e]));
                    }
                  }
                } else {
#pragma unroll
                  for (int c = 0; c < 2; ++c) {      // logical 16-byte chunk cs/8 + c, XOR-swizzled with row % 8
                    const uint4 q = lds128(crow_s + ((((cs >> 3) + c) ^ (row & 7)) << 4));
                    v[c * 4 + 0] = __fadd2_rn(v[c * 4 + 0], unpack_bf16x2(q.x));
                    v[c * 4 + 1] = __fadd2_rn(v[c * 4 + 1], unpack_bf16x2(q.y));
                    v[c * 4 + 2] = __fadd2_rn(v[c * 4 + 2], unpack_bf16x2(q.z));
                    v[c * 4 + 3] = __fadd2_rn(v[c * 4 + 3], unpack_bf16x2(q.w));
                  }
                }
              }
              if (!relu_fast) apply_act8x2(v, p.act);   // ReLU is applied on the packed bf16 pairs below
              if constexpr (LNF) {
                if (p.out_stats) {   // statistics of what the consumers will read: the bf16-rounded values
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const float2 r2 = __bfloat1622float2(__float22bfloat162_rn(v[e]));
                    st_sum = __fadd2_rn(st_sum, r2);
                    st_sq = __ffma2_rn(r2, r2, st_sq);
                  }
                }
              }
            }
            if (relu_fast) {
#pragma unroll
              for (int c = 0; c < 2; ++c)
                sts128(crow_s + ((((cs >> 3) + c) ^ (row & 7)) << 4), pack_relu_bf16x2(v[c * 4].x, v[c * 4].y),
                       pack_relu_bf16x2(v[c * 4 + 1].x, v[c * 4 + 1].y), pack_relu_bf16x2(v[c * 4 + 2].x, v[c * 4 + 2].y),
                       pack_relu_bf16x2(v[c * 4 + 3].x, v[c * 4 + 3].y));
            } else {
#pragma unroll
              for (int c = 0; c < 2; ++c)
                sts128(crow_s + ((((cs >> 3) + c) ^ (row & 7)) << 4), pack_bf16x2(v[c * 4].x, v[c * 4].y),
                       pack_bf16x2(v[c * 4 + 1].x, v[c * 4 + 1].y), pack_bf16x2(v[c * 4 + 2].x, v[c * 4 + 2].y),
                       pack_bf16x2(v[c * 4 + 3].x, v[c * 4 + 3].y));
            }
          }
          fence_proxy_async_smem();                       // generic-proxy writes -> visible to the TMA store
          // every warp that shares this C tile has finished it
          asm volatile("bar.sync %0, %1;" ::"r"(1 + my_sub), "n"(kGps * 128) : "memory");
          if (storer && lane == 0) {
            tma_store_5d(ctile, &p.out_map, n_blk * BLOCK_N + my_sub * 64, iw * p.bw, ih * p.bh, 0, in * p.nf);
            tma_store_commit();
          }
          // TSM: this warp's 32 rows x kCg columns, re-read with 2 lanes per row (32-byte segments)
          if (tsm) {
#pragma unroll
            for (int pass = 0; pass < kCg / 16; ++pass) {
              const int cs = col_in_sub + pass * 16;
              const int col0 = n_blk * BLOCK_N + my_sub * 64 + cs;
              const bool zone_a = col0 < p.tsm_fold;
              const bool zone_b = !zone_a && col0 < 2 * p.tsm_fold;
              if (zone_a || zone_b) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const int rr = quarter * 32 + srow + 16 * i;
                  if (g_i[i] >= 0) {
                    const uint4 o = lds128(ctile_s + rr * 128 + ((((cs >> 3) + spiece) ^ (rr & 7)) << 4));
                    if (zone_a && t_i[i] >= 1)           // out[t-1, c] = x[t, c]   (shift left in time)
                      *reinterpret_cast<uint4*>(tsm + (g_i[i] - HW) * p.tsm_ld + col0 + spiece * 8) = o;
                    if (zone_b && t_i[i] + 1 < p.T)      // out[t+1, c] = x[t, c]   (shift right in time)
                      *reinterpret_cast<uint4*>(tsm + (g_i[i] + HW) * p.tsm_ld + col0 + spiece * 8) = o;
                  }
                }
              }
            }
          }
          if constexpr (LNF) {
            if (p.out_stats && row_ok) {
              atomicAdd(&p.out_stats[grow].x, st_sum.x + st_sum.y);
              atomicAdd(&p.out_stats[grow].y, st_sq.x + st_sq.y);
            }
          }
          // release the slot; the storing warp first waits until its bulk store has finished READING the tile (the
          // TSM scatter above gave it time).  Releasing one tile late instead would chain the four groups together:
          // each group's next slot is the one its neighbour group has just used.
          __syncwarp();
          if (lane == 0) {
            if (storer) tma_store_wait_read<0>();
            mbar_arrive(&c_empty[slot]);
          }
        }
        else if (my_sub < kSub) {   // N-edge tile without this group's sub-tile: hand the ring position straight back
          const int my_it = c_it + my_sub;
          const int slot = my_it % kCSlots;
          mbar_wait(&c_full[slot], (my_it / kCSlots) & 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&c_empty[slot]);
        }
        c_it += kSub;
      }
      bool need_release = true;
      if constexpr (!TF32X3) need_release = !released_any;
      if (need_release) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG2) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
        }
      }
    }
    if constexpr (!TF32X3) {
      if (lane == 0) tma_store_wait_all();   // (storing warps) stores complete before the CTA exits
    }
  } else if (TF32X3 && warp >= 12) {
    // ------------------------------------------------------------ TF32 hi/lo splitter (fp32 verification mode)
    // hi keeps the top 19 bits (exactly representable in TF32); lo = x - hi is exact in fp32.
    const int tid = threadIdx.x - 12 * 32;   // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = unit; tile < total_tiles; tile += n_units) {
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
  if constexpr (CG2) cluster_sync(); else __syncthreads();   // the peer's barriers / smem stay valid until both are done
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG2) tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// LNF: epilogue with the LayerNorm terms (BERT GEMMs of the bf16 path)
template <int BLOCK_N, bool TF32X3, bool LNF = false>
__global__ void __launch_bounds__(ConvGemmCfg<BLOCK_N, TF32X3>::kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  conv_gemm_body<BLOCK_N, TF32X3, false, LNF>(p);
}

// CTA-pair variant: launch with an even grid; consecutive CTAs (2i, 2i+1) form the cluster.
template <int BLOCK_N, bool LNF = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ConvGemmCfg<BLOCK_N, false>::kThreads, 1)
conv_gemm_pair_kernel(const __grid_constant__ ConvGemmParams p) {
  conv_gemm_body<BLOCK_N, false, true, LNF>(p);
}

}  // namespace vcg
