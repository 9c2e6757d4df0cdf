// Layer-2 bottleneck tail (planes P = 128, stride 1, 28 x 28 maps): the halo-patch kernel of conv23h.cuh for two K blocks.
//   conv2 (3x3, 128 -> 128) + BN + ReLU -> conv3 (1x1, 128 -> 512) + BN + residual + ReLU, bf16, one persistent CTA per SM.
//
// What bounded conv23_kernel<128> (33.7 ms per 1-hour video, 569 TFLOP/s, tensor pipe < 30 %): not L2 bandwidth (a probe
// streams > 20 TB/s of L2 hits into shared memory, tools/probes/mc_probe.cu) but the DEPTH of its TMA ring - three 32 KB
// stages ([A tap tile | W2 block]) against ~0.8 us of L2 latency feed the tensor core one 0.13 us stage every ~0.27 us.
// Here the conv2 input of a tile is two TMA boxes (the 10 x 18 halo patch, one per 64-channel block; the nine taps are
// descriptor views of it, conv23h.cuh), so the ring carries ONLY weights in 16 KB stages: W2 blocks [128 x 64] and W3
// chunks [128 output channels x 64], 26 per tile, four in flight.  Residual add on the tensor core (identity MMA), lean
// epilogue, L2 prefetch of the next tile's patch and residual as in conv23h.cuh.
//
// Tile = 8 x 16 pixels of one frame: 28 = 3.5 x 8 = 1.75 x 16, so 8 tiles per frame cover 6.125 tiles' worth of pixels
// (23 % of the MMA rows are clipped away).  MEASURED (512 frames): 268 us against 260 us for conv23_kernel<128> - per tile
// it is 20 % faster (tensor pipe 50 % busy instead of < 30 %), but there are 30 % more tiles, and at N = 128 every MMA
// already reads 128 B/clk of operands from shared memory.  The kernel is therefore OPT-IN (VCG_C23H2=1) and documents the
// design point; it is parity-tested like the others (tests/test_bottleneck_tail_gpu.py).
// Issue order per tile j (B(t, nb) = conv3 of tile t, output channels [256 nb, +256); cb = 64-channel block of conv2's input):
//   B(j-1,0) conv3 | taps 0-4 of cb 0 | B(j-1,0) residual | taps 5-8 of cb 0 | B(j-1,1) conv3 | taps 0-4 of cb 1 |
//   B(j-1,1) residual | taps 5-8 of cb 1
// so that the conv3 MMAs of the previous tile cover the ring refills and the residual boxes (whose C slots only free up when
// the previous sub-tile has been stored) get five taps of slack before the tensor core needs them.  The two channel blocks of
// the halo patch are separate buffers filled by their own producer (warp 2): block 0 of tile j+1 loads while block 1 of tile j
// is being multiplied.  TMEM: conv2 accumulator columns [0, 128), conv3 accumulator [256, 512).
#pragma once
#include "conv23h.cuh"

namespace vcg {

constexpr int kC23h2WStages = 4;
constexpr int kC23h2WBytes = 128 * 128;                               // one ring stage: 128 rows x 64 channels
constexpr int kC23h2CSlots = 4;
constexpr int kC23h2BiasBytes = (512 + 128) * 4;
constexpr int kC23h2SmemBytes = kC23h2WStages * kC23h2WBytes + 2 * kC23hHaloStride + 2 * kCBytes /*A2*/ + kC23hIdentBytes +
                                kC23h2CSlots * kCBytes + kC23h2BiasBytes + 1024 /*align*/ + 512 /*barriers*/;
static_assert(kC23h2SmemBytes <= 232448, "shared memory budget exceeded");

template <int kDummy = 0>
__global__ void __launch_bounds__(kC23Threads, 1) conv23h2_kernel(const __grid_constant__ Conv23Params q) {
  const ConvGemmParams& p = q.g;
  constexpr int kCSlots = kC23h2CSlots, kWS = kC23h2WStages;   // planes P = 128
  constexpr int kBw = 8, kBh = 16;
  constexpr uint32_t kAccA = 0, kAccB = 256;               // TMEM columns of the conv2 / conv3 accumulators

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                                       // [kWS][128 rows][128 B]
  uint8_t* sHalo = sW + kWS * kC23h2WBytes;                 // [2 channel blocks][18][10][128 B]
  uint8_t* sA2 = sHalo + 2 * kC23hHaloStride;               // [2 K blocks][128 rows][128 B]
  uint8_t* sIdent = sA2 + 2 * kCBytes;                      // [64 rows][128 B]
  uint8_t* sC = sIdent + kC23hIdentBytes;                   // [kCSlots][128 rows][128 B]
  float* sBias = reinterpret_cast<float*>(sC + kCSlots * kCBytes);   // [512] conv3 bias, then [128] conv2 bias
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + kC23h2BiasBytes);
  uint64_t* halo_full = bars;                    // [2] one per channel block
  uint64_t* halo_empty = bars + 2;               // [2]
  uint64_t* ta_full = bars + 4;                  // conv2 accumulator ready
  uint64_t* ta_empty = bars + 5;
  uint64_t* tb_full = bars + 6;                  // conv3 accumulator ready
  uint64_t* tb_empty = bars + 7;
  uint64_t* a2_full = bars + 8;                  // epilogue A -> MMA
  uint64_t* a2_empty = bars + 9;                 // MMA (conv3 retired) -> epilogue A
  uint64_t* c_full = bars + 10;                  // [kCSlots]
  uint64_t* c_empty = c_full + kCSlots;
  uint64_t* w_full = c_empty + kCSlots;          // [kWS]
  uint64_t* w_empty = w_full + kWS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + kWS);

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
    for (int i = 0; i < 2; ++i) { mbar_init(&halo_full[i], 1); mbar_init(&halo_empty[i], 1); }
    mbar_init(ta_full, 1); mbar_init(ta_empty, kEpiWarpsBf16);
    mbar_init(tb_full, 1); mbar_init(tb_empty, kEpiWarpsBf16);
    mbar_init(a2_full, kEpiWarpsBf16); mbar_init(a2_empty, 1);
    for (int i = 0; i < kCSlots; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 4); }
    for (int i = 0; i < kWS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kC23hIdentBytes / 16; i += blockDim.x) {
    const int n = i >> 3, chunk = i & 7;
    const int lc = chunk ^ (n & 7);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (lc == (n >> 3)) w[(n & 7) >> 1] = (n & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 at column n
    sts128(smem_u32(sIdent) + i * 16, w[0], w[1], w[2], w[3]);
  }
  for (int i = threadIdx.x; i < 512 + 128; i += blockDim.x) sBias[i] = i < 512 ? __ldg(p.bias + i) : __ldg(q.bias2 + (i - 512));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int nM = m_tiles > static_cast<int>(blockIdx.x)
                     ? (m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  const uint32_t tpi = static_cast<uint32_t>(p.tiles_w * p.tiles_h);
  auto tile_at = [&](int j, int& iw, int& ih, int& in) {
    const uint32_t m_blk = blockIdx.x + static_cast<uint32_t>(j) * gridDim.x;
    const uint32_t n = __umulhi(m_blk, q.magic_tpi);
    const uint32_t rem = m_blk - n * tpi;
    const uint32_t h = __umulhi(rem, q.magic_tw);
    in = static_cast<int>(n); ih = static_cast<int>(h); iw = static_cast<int>(rem - h * static_cast<uint32_t>(p.tiles_w));
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: the weight ring, in MMA issue order
    if (elect_one()) {
      pdl_wait();
      int ws = 0;
      uint32_t wphase = 0;
      auto ring_load = [&](const CUtensorMap* map, int k, int row) {
        mbar_wait(&w_empty[ws], wphase ^ 1);
        mbar_expect_tx(&w_full[ws], kC23h2WBytes);
        tma_load_2d(sW + ws * kC23h2WBytes, map, &w_full[ws], k, row);
        if (++ws == kWS) { ws = 0; wphase ^= 1; }
      };
      auto load_b = [&](int nb) {                           // W3 chunks of conv3 sub-tile nb: (K block, half) -> 128 rows x 64
        for (int kb = 0; kb < 2; ++kb)
          for (int h = 0; h < 2; ++h) ring_load(&q.w3_map, kb * 64, nb * 256 + h * 128);
      };
      for (int j = 0; j < nM; ++j) {
        if (j > 0) load_b(0);
        for (int tap = 0; tap < 9; ++tap) ring_load(&p.b_map, tap * 128, 0);          // channel block 0
        if (j > 0) load_b(1);
        for (int tap = 0; tap < 9; ++tap) ring_load(&p.b_map, tap * 128 + 64, 0);     // channel block 1
      }
      if (nM > 0) { load_b(0); load_b(1); }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_n128 = umma_idesc(1u, kBlockM, 128);
      constexpr uint32_t idesc_n64 = umma_idesc(1u, kBlockM, 64);
      const bool has_res = p.residual != nullptr;
      const uint32_t w_addr = smem_u32(sW), a2_addr = smem_u32(sA2), id_addr = smem_u32(sIdent), c_addr = smem_u32(sC);
      const uint32_t h_addr = smem_u32(sHalo);
      int ws = 0, c_it = 0, nb_count = 0;
      uint32_t wphase = 0;
      auto issue_taps = [&](int cb, int t0, int t1) {       // conv2 K steps of channel block cb: taps [t0, t1)
        for (int tap = t0; tap < t1; ++tap) {
          mbar_wait(&w_full[ws], wphase);
          tc_fence_after();
          const uint32_t a_addr = h_addr + cb * kC23hHaloStride + ((tap / 3) * kC23hHaloW + (tap % 3)) * 128;
          const uint32_t b_addr = w_addr + ws * kC23h2WBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + kAccA, umma_desc_sw128_sbo(a_addr + k * 32, kC23hHaloW * 128), umma_desc_sw128(b_addr + k * 32),
                      idesc_n128, (cb | tap | k) != 0);
          umma_commit(&w_empty[ws]);
          if (++ws == kWS) { ws = 0; wphase ^= 1; }
        }
      };
      auto b_main = [&](int j, int nb) {                    // conv3 of tile j, output channels [256 nb, +256)
        mbar_wait(tb_empty, (nb_count & 1) ^ 1);
        if (nb == 0) mbar_wait(a2_full, j & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + kAccB;
        for (int kb = 0; kb < 2; ++kb)
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&w_full[ws], wphase);
            tc_fence_after();
            const uint32_t b_addr = w_addr + ws * kC23h2WBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem + h * 128, umma_desc_sw128(a2_addr + kb * kCBytes + k * 32), umma_desc_sw128(b_addr + k * 32),
                        idesc_n128, (kb | k) != 0);
            umma_commit(&w_empty[ws]);
            if (++ws == kWS) { ws = 0; wphase ^= 1; }
          }
        if (nb == 1) umma_commit(a2_empty);                 // the conv3 operand of tile j may be overwritten
      };
      auto b_res = [&]() {                                  // residual add of the sub-tile started last: D[:, 64 g ..] += R_g * I
        const uint32_t d_tmem = tmem_base + kAccB;
        for (int g = 0; g < 4; ++g, ++c_it) {
          const int slot = c_it % kCSlots;
          mbar_wait(&c_full[slot], (c_it / kCSlots) & 1);
          tc_fence_after();
          if (has_res) {
            const uint32_t r_addr = c_addr + slot * kCBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem + g * 64, umma_desc_sw128(r_addr + k * 32), umma_desc_sw128(id_addr + k * 32), idesc_n64, 1u);
          }
        }
        umma_commit(tb_full);
        ++nb_count;
      };
      for (int j = 0; j < nM; ++j) {
        if (j > 0) b_main(j - 1, 0);
        mbar_wait(ta_empty, (j & 1) ^ 1);
        mbar_wait(&halo_full[0], j & 1);
        tc_fence_after();
        issue_taps(0, 0, 5);
        if (j > 0) b_res();
        issue_taps(0, 5, 9);
        umma_commit(&halo_empty[0]);
        if (j > 0) b_main(j - 1, 1);
        mbar_wait(&halo_full[1], j & 1);
        tc_fence_after();
        issue_taps(1, 0, 5);
        if (j > 0) b_res();
        issue_taps(1, 5, 9);
        umma_commit(&halo_empty[1]);
        umma_commit(ta_full);
      }
      if (nM > 0) { b_main(nM - 1, 0); b_res(); b_main(nM - 1, 1); b_res(); }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ halo producer: channel block cb of tile j into buffer cb
    if (elect_one()) {
      pdl_wait();
      for (int j = 0; j < nM; ++j) {
        int iw, ih, in;
        tile_at(j, iw, ih, in);
        for (int cb = 0; cb < 2; ++cb) {
          mbar_wait(&halo_empty[cb], (j & 1) ^ 1);
          mbar_expect_tx(&halo_full[cb], kC23hHaloBytes);
          tma_load_5d(sHalo + cb * kC23hHaloStride, &p.a_map[0], &halo_full[cb], cb * 64, iw * kBw - 1, ih * kBh - 1, 0, in);
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ C producer: residual boxes in B order, L2 prefetch ahead
    if (elect_one()) {
      pdl_wait();
      const bool has_res = p.residual != nullptr;
      const int kPF = q.prefetch_tiles;
      auto prefetch_tile = [&](int j2) {
        if (j2 >= nM) return;
        int iw, ih, in;
        tile_at(j2, iw, ih, in);
        tma_prefetch_l2_5d(&p.a_map[0], 0, iw * kBw - 1, ih * kBh - 1, 0, in);
        tma_prefetch_l2_5d(&p.a_map[0], 64, iw * kBw - 1, ih * kBh - 1, 0, in);
        if (has_res)
          for (int jj = 0; jj < 8; ++jj) tma_prefetch_l2_5d(&p.res_map, jj * 64, iw * kBw, ih * kBh, 0, in);
      };
      if (kPF > 0) for (int j2 = 0; j2 < kPF; ++j2) prefetch_tile(j2);
      int c_it = 0;
      for (int j = 0; j < nM; ++j) {
        int iw, ih, in;
        tile_at(j, iw, ih, in);
        if (kPF > 0) prefetch_tile(j + kPF);
        for (int jj = 0; jj < 8; ++jj, ++c_it) {            // nb = jj / 4, group = jj % 4
          const int slot = c_it % kCSlots;
          mbar_wait(&c_empty[slot], ((c_it / kCSlots) & 1) ^ 1);
          if (has_res) {
            mbar_expect_tx(&c_full[slot], kCBytes);
            tma_load_5d(sC + slot * kCBytes, &p.res_map, &c_full[slot], jj * 64, iw * kBw, ih * kBh, 0, in);
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
    const uint32_t rsw = static_cast<uint32_t>(row & 7);
    const uint32_t a2_row = smem_u32(sA2) + (group >> 1) * kCBytes + row * 128;   // group g: conv2 channels [32 g, 32 g + 32)
    const uint32_t bias3_addr = smem_u32(sBias);
    const uint32_t bias2_addr = smem_u32(sBias) + (512 + group * 32) * 4;
    const uint32_t c_addr = smem_u32(sC);
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int c_it = 0, nb_count = 0;
    auto epi_b = [&](int j, int nb) {
      mbar_wait(tb_full, nb_count & 1);
      ++nb_count;
      tc_fence_after();
      int iw, ih, in;
      tile_at(j, iw, ih, in);
      const int my_it = c_it + group;
      const int slot = my_it % kCSlots;
      const uint32_t crow = c_addr + slot * kCBytes + row * 128;
      const uint32_t taddr = lane_addr + kAccB + group * 64;
      const uint32_t bias = bias3_addr + (nb * 256 + group * 64) * 4;
      uint32_t rr[2][16];
      tmem_ld_32x16(taddr, rr[0]);
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        tmem_ld_wait();
        if (pass < 3) {
          tmem_ld_32x16(taddr + (pass + 1) * 16, rr[(pass + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tb_empty);              // accumulator handed back right after the last tcgen05.ld
        }
        const uint32_t (&r)[16] = rr[pass & 1];
        uint32_t o[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 b4 = lds128f(bias + (pass * 16 + e * 4) * 4);
          o[2 * e] = pack_relu_bf16x2(__uint_as_float(r[4 * e]) + b4.x, __uint_as_float(r[4 * e + 1]) + b4.y);
          o[2 * e + 1] = pack_relu_bf16x2(__uint_as_float(r[4 * e + 2]) + b4.z, __uint_as_float(r[4 * e + 3]) + b4.w);
        }
        sts128(crow + (((2u * pass) ^ rsw) << 4), o[0], o[1], o[2], o[3]);
        sts128(crow + (((2u * pass + 1u) ^ rsw) << 4), o[4], o[5], o[6], o[7]);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(128) : "memory");
      if (quarter == 0 && lane == 0) {
        tma_store_5d(reinterpret_cast<const void*>(sC + slot * kCBytes), &p.out_map, nb * 256 + group * 64, iw * kBw, ih * kBh, 0, in);
        tma_store_commit();
      }
      __syncwarp();
      if (lane == 0) {
        if (quarter == 0) tma_store_wait_read<0>();
        mbar_arrive(&c_empty[slot]);
      }
      c_it += 4;
    };
    for (int j = 0; j < nM; ++j) {
      if (j > 0) { epi_b(j - 1, 0); epi_b(j - 1, 1); }
      // ---- A(j): conv2 accumulator -> + bias2 -> bf16 -> ReLU -> A2 (two K-major 128-byte-swizzled blocks)
      mbar_wait(ta_full, j & 1);
      tc_fence_after();
      uint32_t r[2][16];
      tmem_ld_32x16(lane_addr + kAccA + group * 32, r[0]);
      tmem_ld_32x16(lane_addr + kAccA + group * 32 + 16, r[1]);
      mbar_wait(a2_empty, (j & 1) ^ 1);
      tmem_ld_wait();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 b4 = lds128f(bias2_addr + (hh * 16 + e * 4) * 4);
          o[2 * e] = pack_relu_bf16x2(__uint_as_float(r[hh][4 * e]) + b4.x, __uint_as_float(r[hh][4 * e + 1]) + b4.y);
          o[2 * e + 1] = pack_relu_bf16x2(__uint_as_float(r[hh][4 * e + 2]) + b4.z, __uint_as_float(r[hh][4 * e + 3]) + b4.w);
        }
        const uint32_t chunk0 = static_cast<uint32_t>((group & 1) * 4 + hh * 2);     // 16-byte chunk inside the 64-channel block
        sts128(a2_row + (((chunk0) ^ rsw) << 4), o[0], o[1], o[2], o[3]);
        sts128(a2_row + (((chunk0 + 1u) ^ rsw) << 4), o[4], o[5], o[6], o[7]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a2_full);
        mbar_arrive(ta_empty);
      }
    }
    if (nM > 0) { epi_b(nM - 1, 0); epi_b(nM - 1, 1); }
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
