// Layer-1 bottleneck tail (planes P = 64, stride 1) with EVERYTHING the tensor core reads resident in shared memory:
//   conv2 (3x3) + BN + ReLU -> conv3 (1x1) + BN + residual + ReLU (+ TSM scatter), bf16, one persistent CTA per SM.
//
// conv23_kernel (conv23.cuh) feeds conv2 as an implicit GEMM with one TMA box per filter tap: every 128-pixel tile pulls
// its input patch NINE times and both weight matrices once through L2 -> shared memory (312 KB per tile; ncu: L2 -> SM
// traffic 2.0x the DRAM bytes, tensor pipe 17 % busy).  Here
//   * W3 [256][64] (32 KB) is loaded ONCE per CTA and stays in shared memory; W2 streams through a ring of four 8 KB
//     tap tiles (72 KB per tile out of L2, which is not the bottleneck; keeping W2 resident as well was measured: it
//     leaves room for only three C slots and the epilogue then serialises on them, 7.1 us per tile instead of 5.8),
//   * the RESIDUAL ADD runs on the tensor core: the bf16 residual tile TMA drops into a C slot is exactly a K-major
//     A operand, so D[:, 64 g ..] += R_g * I_64 (four N = 64 MMAs per slot, exact: 1.0 * r in fp32) replaces the
//     epilogue's shared-memory reads, bf16 unpacking and adds.  ncu on the first version of this kernel: 54 % of the
//     issue slots busy, ~1100 instructions per warp and tile in the conv3 epilogue, tensor pipe 20 % busy, HBM at 63 %
//     and L2 -> SM traffic down to 1.17x the DRAM bytes without any gain in time - the bottleneck was instruction issue
//     in the epilogue, not memory.  The epilogue is now tcgen05.ld -> + bias (shared memory) -> cvt.bf16x2 -> max.bf16x2
//     -> st.shared through 32-bit shared addresses: ~35 instructions per 16-column pass.
//   * the conv2 input of a tile is ONE TMA box: the (8+2) x (16+2) pixel halo patch of the 8 x 16 output tile (23 KB,
//     zero-filled outside the image = conv padding), double buffered,
//   * the nine taps are nine VIEWS of that patch: a tile row is 8 consecutive pixels = 8 consecutive 128-byte rows of
//     the patch = one 8-row swizzle group of the UMMA shared-memory descriptor, consecutive tile rows are one patch row
//     (10 pixels = 1280 B) apart = the descriptor's stride byte offset, and tap (kh, kw) just moves the start address
//     by (kh*10 + kw) * 128 B.  The 128-byte swizzle is a function of the shared-memory address bits, TMA wrote the
//     patch with the same function, so any 128-byte-aligned start inside the 1024-byte-aligned patch reads back right.
// Per tile that leaves 23 KB (patch) + 64 KB (residual) + 72 KB (W2, L2 hits) of loads and 64 + 16 KB of stores, and
// six 16 KB C slots keep ~100 KB of residual loads / output stores in flight per SM: the kernel sits on HBM.
// Tile = 8 x 16 pixels of ONE frame (56 = 3.5 x 16: the last tile row of a frame is half empty, TMA clips / zero-fills).
#pragma once
#include "conv23.cuh"

namespace vcg {

constexpr int kC23hHaloW = 10, kC23hHaloH = 18;
constexpr int kC23hHaloBytes = kC23hHaloW * kC23hHaloH * 128;        // 23 040 B delivered by TMA
constexpr int kC23hHaloStride = 23 * 1024;                            // stage pitch (1024-byte aligned)
constexpr int kC23hHaloStages = 2;
constexpr int kC23hW2Stages = 5;
constexpr int kC23hTapBytes = 64 * 128;                               // one tap of W2: 64 rows x 64 channels
constexpr int kC23hW2Bytes = kC23hW2Stages * kC23hTapBytes;           // 40 960
constexpr int kC23hW3Bytes = 256 * 128;                               // 32 768
constexpr int kC23hIdentBytes = 64 * 128;                             // 64 x 64 identity (residual add on the tensor core)
constexpr int kC23hCSlots = 6;
constexpr int kC23hBiasBytes = (256 + 64) * 4;
constexpr int kC23hSmemBytes = kC23hW2Bytes + kC23hW3Bytes + kC23hHaloStages * kC23hHaloStride +
                               kC23hIdentBytes + kC23hCSlots * kCBytes + kC23hBiasBytes + 1024 /*align*/ + 512 /*barriers*/;
static_assert(kC23hSmemBytes <= 232448, "shared memory budget exceeded");

// K-major SW128 operand whose 8-row groups are `sbo` bytes apart (a view into the halo patch)
__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int kDummy = 0>
__global__ void __launch_bounds__(kC23Threads, 1) conv23h_kernel(const __grid_constant__ Conv23Params q) {
  const ConvGemmParams& p = q.g;
  constexpr int P = 64, BLOCK_N = 256, kCSlots = kC23hCSlots, kHS = kC23hHaloStages, kW2S = kC23hW2Stages;
  constexpr int kBw = 8, kBh = 16;                          // tile: 8 x 16 pixels of one frame
  constexpr uint32_t kA2Col = 128;                          // TMEM columns [128, 160): conv2's output tile as conv3's A operand

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW2 = smem;                                      // [kW2S tap stages][64 rows][128 B]
  uint8_t* sW3 = sW2 + kC23hW2Bytes;                        // [256 rows][128 B]
  uint8_t* sHalo = sW3 + kC23hW3Bytes;                      // [kHS][18][10][128 B]
  uint8_t* sIdent = sHalo + kHS * kC23hHaloStride;          // [64 rows][128 B]: I_64 as a K-major SW128 B operand
  uint8_t* sC = sIdent + kC23hIdentBytes;                   // [kCSlots][128 rows][128 B]
  float* sBias3 = reinterpret_cast<float*>(sC + kCSlots * kCBytes);   // [256] conv3 bias, then [64] conv2 bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias3) + kC23hBiasBytes);
  uint64_t* w_full = bars;                       // [1]
  uint64_t* halo_full = bars + 1;                // [kHS]
  uint64_t* halo_empty = halo_full + kHS;        // [kHS]
  uint64_t* tmem_full = halo_empty + kHS;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint64_t* a2_full = tmem_empty + 2;            // [1] epilogue A -> MMA
  uint64_t* a2_empty = a2_full + 1;              // [1] MMA (conv3 retired) -> epilogue A
  uint64_t* c_full = a2_empty + 1;               // [kCSlots] residual landed -> MMA
  uint64_t* c_empty = c_full + kCSlots;          // [kCSlots] output store has read the slot -> C producer
  uint64_t* w2_full = c_empty + kCSlots;         // [kW2S]
  uint64_t* w2_empty = w2_full + kW2S;           // [kW2S]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w2_empty + kW2S);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&q.w3_map);
    tma_prefetch_desc(&p.out_map);
    tma_prefetch_desc(&p.res_map);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < kHS; ++s) { mbar_init(&halo_full[s], 1); mbar_init(&halo_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarpsBf16); }
    mbar_init(a2_full, kEpiWarpsBf16);
    mbar_init(a2_empty, 1);
    for (int i = 0; i < kCSlots; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 4); }
    for (int i = 0; i < kW2S; ++i) { mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // constants of the forward pass (written once at load time): identity operand and the two bias vectors
  for (int i = threadIdx.x; i < kC23hIdentBytes / 16; i += blockDim.x) {
    const int n = i >> 3, chunk = i & 7;                    // physical 16-byte chunk `chunk` of row n holds logical chunk chunk ^ (n & 7)
    const int lc = chunk ^ (n & 7);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (lc == (n >> 3)) w[(n & 7) >> 1] = (n & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 at column n
    sts128(smem_u32(sIdent) + i * 16, w[0], w[1], w[2], w[3]);
  }
  for (int i = threadIdx.x; i < 256 + 64; i += blockDim.x) sBias3[i] = i < 256 ? __ldg(p.bias + i) : __ldg(q.bias2 + (i - 256));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int nM = m_tiles > static_cast<int>(blockIdx.x)
                     ? (m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  // tile j of this CTA -> (iw, ih, frame); exact magic-number division (m_blk < 2^24, divisors < 2^8)
  const uint32_t tpi = static_cast<uint32_t>(p.tiles_w * p.tiles_h);
  auto tile_at = [&](int j, int& iw, int& ih, int& in) {
    const uint32_t m_blk = blockIdx.x + static_cast<uint32_t>(j) * gridDim.x;
    const uint32_t n = __umulhi(m_blk, q.magic_tpi);
    const uint32_t rem = m_blk - n * tpi;
    const uint32_t h = __umulhi(rem, q.magic_tw);
    in = static_cast<int>(n); ih = static_cast<int>(h); iw = static_cast<int>(rem - h * static_cast<uint32_t>(p.tiles_w));
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: W3 once, then per tile one halo patch + 9 taps of W2
    if (elect_one()) {
      // the weights are constants of the forward pass: they may be fetched before the previous kernel has finished
      mbar_expect_tx(w_full, kC23hW3Bytes);
      tma_load_2d(sW3, &q.w3_map, w_full, 0, 0);
      pdl_wait();
      int ws = 0;
      uint32_t wphase = 0;
      for (int j = 0; j < nM; ++j) {
        const int st = j % kHS;
        int iw, ih, in;
        tile_at(j, iw, ih, in);
        mbar_wait(&halo_empty[st], ((j / kHS) & 1) ^ 1);
        mbar_expect_tx(&halo_full[st], kC23hHaloBytes);
        tma_load_5d(sHalo + st * kC23hHaloStride, &p.a_map[0], &halo_full[st], 0, iw * kBw - 1, ih * kBh - 1, 0, in);
        for (int t = 0; t < 9; ++t) {
          mbar_wait(&w2_empty[ws], wphase ^ 1);
          mbar_expect_tx(&w2_full[ws], kC23hTapBytes);
          tma_load_2d(sW2 + ws * kC23hTapBytes, &p.b_map, &w2_full[ws], t * 64, 0);
          if (++ws == kW2S) { ws = 0; wphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc(1u, kBlockM, P);          // conv2 and the residual add: N = 64
      constexpr uint32_t idesc2 = umma_idesc(1u, kBlockM, BLOCK_N);
      const bool has_res = p.residual != nullptr;
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t w2_addr = smem_u32(sW2), w3_addr = smem_u32(sW3), id_addr = smem_u32(sIdent);
      const uint32_t a2_tmem = tmem_base + kA2Col;          // conv3's A operand lives in tensor memory
      const uint32_t c_addr = smem_u32(sC);
      int ws = 0, c_it = 0;
      uint32_t wphase = 0;
      // conv2 of tile j into accumulator 0, conv3 (+ residual) of tile j into accumulator 1.  The W2 tap ring refills at L2
      // latency, so conv3 of the PREVIOUS tile is issued in the middle of a tile's nine taps: taps 0-4 come out of the
      // ring as prefetched, the ~0.7 us of conv3 + residual MMAs cover the refill, taps 5-8 follow.
      auto issue_taps = [&](uint32_t h_addr, int t0, int t1) {
        for (int tap = t0; tap < t1; ++tap) {
          mbar_wait(&w2_full[ws], wphase);
          tc_fence_after();
          const uint32_t a_addr = h_addr + ((tap / 3) * kC23hHaloW + (tap % 3)) * 128;
          const uint32_t b_addr = w2_addr + ws * kC23hTapBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, umma_desc_sw128_sbo(a_addr + k * 32, kC23hHaloW * 128), umma_desc_sw128(b_addr + k * 32), idesc1,
                      (tap | k) != 0);
          umma_commit(&w2_empty[ws]);
          if (++ws == kW2S) { ws = 0; wphase ^= 1; }
        }
      };
      auto issue_conv3 = [&](int j) {
        mbar_wait(&tmem_empty[1], (j & 1) ^ 1);
        mbar_wait(a2_full, j & 1);                          // the conv3 operand of tile j has been written
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + BLOCK_N;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(d_tmem, a2_tmem + k * 8, umma_desc_sw128(w3_addr + k * 32), idesc2, k != 0);
        umma_commit(a2_empty);                              // A2 may be overwritten once these MMAs retire
        // residual add on the tensor core: D[:, 64 g .. 64 g + 63] += R_g * I  (R_g = the bf16 residual tile TMA put into
        // C slot g, exactly an A operand; 1.0 * r accumulates exactly in fp32).  Four N = 64 MMAs per slot.
        for (int g = 0; g < BLOCK_N / 64; ++g, ++c_it) {
          const int slot = c_it % kCSlots;
          mbar_wait(&c_full[slot], (c_it / kCSlots) & 1);
          tc_fence_after();
          if (has_res) {
            const uint32_t r_addr = c_addr + slot * kCBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem + g * 64, umma_desc_sw128(r_addr + k * 32), umma_desc_sw128(id_addr + k * 32), idesc1, 1u);
          }
        }
        umma_commit(&tmem_full[1]);                         // also: the residual slots of this tile have been read
      };
      for (int j = 0; j < nM; ++j) {
        const int st = j % kHS;
        mbar_wait(&tmem_empty[0], (j & 1) ^ 1);
        mbar_wait(&halo_full[st], (j / kHS) & 1);
        tc_fence_after();
        const uint32_t h_addr = smem_u32(sHalo + st * kC23hHaloStride);
        issue_taps(h_addr, 0, 5);
        if (j > 0) issue_conv3(j - 1);
        issue_taps(h_addr, 5, 9);
        umma_commit(&halo_empty[st]);
        umma_commit(&tmem_full[0]);
      }
      if (nM > 0) issue_conv3(nM - 1);
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ C producer: residual prefetch for the conv3 sub-tiles
    if (elect_one()) {
      pdl_wait();
      const bool has_res = p.residual != nullptr;
      // DRAM latency is hidden in L2, not in shared memory: the residual boxes and the halo patch of tile j + kPF are
      // prefetched into L2 (no smem destination) while tile j is in flight; the five C slots then only have to cover
      // the L2 -> SM latency
      const int kPF = q.prefetch_tiles;
      auto prefetch_tile = [&](int j2) {
        if (j2 >= nM) return;
        int iw, ih, in;
        tile_at(j2, iw, ih, in);
        tma_prefetch_l2_5d(&p.a_map[0], 0, iw * kBw - 1, ih * kBh - 1, 0, in);
        if (has_res) {
          for (int jj = 0; jj < BLOCK_N / 64; ++jj) {
            if (p.res_clip_T == 0) tma_prefetch_l2_5d(&p.res_map, jj * 64, iw * kBw, ih * kBh, 0, in);
            else tma_prefetch_l2_5d(&p.res_map, jj * 64, iw * kBw, ih * kBh, in % p.res_clip_T, in / p.res_clip_T);
          }
        }
      };
      if (kPF > 0) for (int j2 = 0; j2 < kPF; ++j2) prefetch_tile(j2);
      int c_it = 0;
      for (int j = 0; j < nM; ++j) {
        int iw, ih, in;
        tile_at(j, iw, ih, in);
        if (kPF > 0) prefetch_tile(j + kPF);
        for (int jj = 0; jj < BLOCK_N / 64; ++jj, ++c_it) {
          const int slot = c_it % kCSlots;
          mbar_wait(&c_empty[slot], ((c_it / kCSlots) & 1) ^ 1);
          if (has_res) {
            mbar_expect_tx(&c_full[slot], kCBytes);
            if (p.res_clip_T == 0)
              tma_load_5d(sC + slot * kCBytes, &p.res_map, &c_full[slot], jj * 64, iw * kBw, ih * kBh, 0, in);
            else
              tma_load_5d(sC + slot * kCBytes, &p.res_map, &c_full[slot], jj * 64, iw * kBw, ih * kBh, in % p.res_clip_T,
                          in / p.res_clip_T);
          } else {
            mbar_arrive(&c_full[slot]);
          }
        }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ------------------------------------------------------------ epilogue warps (16: quarter x group)
    pdl_wait();
    const int quarter = warp & 3;
    const int group = (warp - kFirstEpiWarp) >> 2;
    const int row = quarter * 32 + lane;
    __nv_bfloat16* tsm = reinterpret_cast<__nv_bfloat16*>(p.tsm_out);
    const int srow = lane >> 1, spiece = lane & 1;
    const uint32_t rsw = static_cast<uint32_t>(row & 7);
    const uint32_t bias3_addr = smem_u32(sBias3) + group * 64 * 4;
    const uint32_t bias2_addr = smem_u32(sBias3) + (256 + group * 16) * 4;
    const uint32_t c_addr = smem_u32(sC);
    int c_it = 0;
    for (int s = 0; s < 2 * nM; ++s) {
      const int j = s >> 1;
      const bool is_a = (s & 1) == 0;                        // A(j): conv2 accumulator 0, B(j): conv3 accumulator 1
      const int acc = s & 1;
      mbar_wait(&tmem_full[acc], j & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      if (is_a) {
        // ---- A: conv2 accumulator -> + bias2 -> bf16 -> ReLU -> A2 in tensor memory; group g takes channels [16 g, 16 g + 16)
        uint32_t r[16];
        tmem_ld_32x16(taddr + group * 16, r);
        mbar_wait(a2_empty, (j & 1) ^ 1);
        tmem_ld_wait();
        uint32_t o[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 b4 = lds128f(bias2_addr + e * 16);
          o[2 * e] = pack_relu_bf16x2(__uint_as_float(r[4 * e]) + b4.x, __uint_as_float(r[4 * e + 1]) + b4.y);
          o[2 * e + 1] = pack_relu_bf16x2(__uint_as_float(r[4 * e + 2]) + b4.z, __uint_as_float(r[4 * e + 3]) + b4.w);
        }
        // conv3's A operand goes back into TENSOR memory (row = lane, 2 bf16 per column): no shared memory, no proxy fence
        tmem_st_32x8(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kA2Col + group * 8, o);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(a2_full);
          mbar_arrive(&tmem_empty[acc]);
        }
        continue;
      }
      // ---- B: conv3 + residual are in the accumulator; group g drains the g-th 64-column C tile:
      //      + bias -> bf16 -> ReLU -> slot (the residual the slot held has been consumed by the tensor core) -> TMA store
      int iw, ih, in;
      tile_at(j, iw, ih, in);
      const int my_it = c_it + group;
      const int slot = my_it % kCSlots;
      const uint32_t ctile = c_addr + slot * kCBytes;
      const uint32_t crow = ctile + row * 128;
      uint32_t rr[2][16];
      tmem_ld_32x16(taddr + group * 64, rr[0]);
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        tmem_ld_wait();
        if (pass < 3) {
          tmem_ld_32x16(taddr + group * 64 + (pass + 1) * 16, rr[(pass + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);      // accumulator handed back right after the last tcgen05.ld
        }
        const uint32_t (&r)[16] = rr[pass & 1];
        uint32_t o[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 b4 = lds128f(bias3_addr + (pass * 16 + e * 4) * 4);
          o[2 * e] = pack_relu_bf16x2(__uint_as_float(r[4 * e]) + b4.x, __uint_as_float(r[4 * e + 1]) + b4.y);
          o[2 * e + 1] = pack_relu_bf16x2(__uint_as_float(r[4 * e + 2]) + b4.z, __uint_as_float(r[4 * e + 3]) + b4.w);
        }
        sts128(crow + (((2u * pass) ^ rsw) << 4), o[0], o[1], o[2], o[3]);
        sts128(crow + (((2u * pass + 1u) ^ rsw) << 4), o[4], o[5], o[6], o[7]);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(128) : "memory");
      if (quarter == 0 && lane == 0) {
        tma_store_5d(reinterpret_cast<const void*>(sC + slot * kCBytes), &p.out_map, group * 64, iw * kBw, ih * kBh, 0, in);
        tma_store_commit();
      }
      if (tsm != nullptr && group == 0) {
        // next bottleneck's temporally shifted input: channels [0, fold) of frame t -> frame t-1, [fold, 2 fold) -> t+1
        // (ops/temporal_shift.py:34-51).  A tile lies in ONE frame, so t is uniform; 2 lanes per row, 16 bytes each.
        const int t = in % p.T;
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          const int col0 = pass * 16;
          const bool zone_a = col0 < p.tsm_fold;
          const bool zone_b = !zone_a && col0 < 2 * p.tsm_fold;
          if ((zone_a && t >= 1) || (zone_b && t + 1 < p.T)) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int r2 = quarter * 32 + srow + 16 * i;
              const int w = iw * kBw + (r2 & (kBw - 1)), h = ih * kBh + (r2 >> 3);
              if (w < p.Wo && h < p.Ho) {
                const uint4 o = lds128(ctile + r2 * 128 + ((((col0 >> 3) + spiece) ^ (r2 & 7)) << 4));
                const long g = (static_cast<long>(in + (zone_a ? -1 : 1)) * p.Ho + h) * p.Wo + w;
                *reinterpret_cast<uint4*>(tsm + g * p.tsm_ld + col0 + spiece * 8) = o;
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        if (quarter == 0) tma_store_wait_read<0>();
        mbar_arrive(&c_empty[slot]);
      }
      c_it += BLOCK_N / 64;
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace vcg
