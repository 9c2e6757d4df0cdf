// Layer-1 bottleneck tail (planes P = 64, stride 1) with EVERYTHING the tensor core reads resident in shared memory:
//   conv2 (3x3) + BN + ReLU -> conv3 (1x1) + BN + residual + ReLU (+ TSM scatter), bf16, one persistent CTA per SM.
//
// conv23_kernel (conv23.cuh) feeds conv2 as an implicit GEMM with one TMA box per filter tap: every 128-pixel tile pulls
// its input patch NINE times and both weight matrices once through L2 -> shared memory (312 KB per tile; ncu: L2 -> SM
// traffic 2.0x the DRAM bytes, tensor pipe 17 % busy).  Here
//   * W2 [64][9*64] (72 KB) and W3 [256][64] (32 KB) are loaded ONCE per CTA and stay in shared memory,
//   * the conv2 input of a tile is ONE TMA box: the (8+2) x (16+2) pixel halo patch of the 8 x 16 output tile (23 KB,
//     zero-filled outside the image = conv padding), double buffered,
//   * the nine taps are nine VIEWS of that patch: a tile row is 8 consecutive pixels = 8 consecutive 128-byte rows of
//     the patch = one 8-row swizzle group of the UMMA shared-memory descriptor, consecutive tile rows are one patch row
//     (10 pixels = 1280 B) apart = the descriptor's stride byte offset, and tap (kh, kw) just moves the start address
//     by (kh*10 + kw) * 128 B.  The 128-byte swizzle is a function of the shared-memory address bits, TMA wrote the
//     patch with the same function, so any 128-byte-aligned start inside the 1024-byte-aligned patch reads back right.
// Per tile that leaves 23 KB (patch) + 64 KB (residual) of loads and 64 + 16 KB of stores: the kernel sits on HBM.
// Tile = 8 x 16 pixels of ONE frame (56 = 3.5 x 16: the last tile row of a frame is half empty, TMA clips / zero-fills).
#pragma once
#include "conv23.cuh"

namespace vcg {

constexpr int kC23hHaloW = 10, kC23hHaloH = 18;
constexpr int kC23hHaloBytes = kC23hHaloW * kC23hHaloH * 128;        // 23 040 B delivered by TMA
constexpr int kC23hHaloStride = 23 * 1024;                            // stage pitch (1024-byte aligned)
constexpr int kC23hHaloStages = 2;
constexpr int kC23hW2Bytes = 9 * 64 * 128;                            // 73 728
constexpr int kC23hW3Bytes = 256 * 128;                               // 32 768
constexpr int kC23hCSlots = 3;
constexpr int kC23hSmemBytes = kC23hW2Bytes + kC23hW3Bytes + kC23hHaloStages * kC23hHaloStride + kCBytes /*A2*/ +
                               kC23hCSlots * kCBytes + 1024 /*align*/ + 512 /*barriers*/;
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
  constexpr int P = 64, BLOCK_N = 256, kCSlots = kC23hCSlots, kHS = kC23hHaloStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW2 = smem;                                      // [9 taps][64 rows][128 B]
  uint8_t* sW3 = sW2 + kC23hW2Bytes;                        // [256 rows][128 B]
  uint8_t* sHalo = sW3 + kC23hW3Bytes;                      // [kHS][18][10][128 B]
  uint8_t* sA2 = sHalo + kHS * kC23hHaloStride;             // [128 rows][128 B]
  uint8_t* sC = sA2 + kCBytes;                              // [kCSlots][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + kCSlots * kCBytes);
  uint64_t* w_full = bars;                       // [1]
  uint64_t* halo_full = bars + 1;                // [kHS]
  uint64_t* halo_empty = halo_full + kHS;        // [kHS]
  uint64_t* tmem_full = halo_empty + kHS;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint64_t* a2_full = tmem_empty + 2;            // [1] epilogue A -> MMA
  uint64_t* a2_empty = a2_full + 1;              // [1] MMA (conv3 retired) -> epilogue A
  uint64_t* c_full = a2_empty + 1;               // [kCSlots]
  uint64_t* c_empty = c_full + kCSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_empty + kCSlots);

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
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int nM = m_tiles > static_cast<int>(blockIdx.x)
                     ? (m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  const int n_sub = 2 * nM;
  // s-th sub-tile of this CTA: A(0), A(1), B(0), A(2), B(1), ..., B(nM-1)   (A = conv2, B = conv3 of a tile)
  auto sub_at = [&](int s, int& j, bool& is_a) {
    if (s == 0) { j = 0; is_a = true; return; }
    const int t = s - 1, blk = t >> 1;
    if (blk < nM - 1) {
      if ((t & 1) == 0) { j = blk + 1; is_a = true; } else { j = blk; is_a = false; }
    } else {
      j = nM - 1; is_a = false;
    }
  };
  auto m_blk_of = [&](int j) { return static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x); };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: weights once, then one halo patch per tile
    if (elect_one()) {
      // the weights are constants of the forward pass: they may be fetched before the previous kernel has finished
      mbar_expect_tx(w_full, kC23hW2Bytes + kC23hW3Bytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(sW2 + t * 64 * 128, &p.b_map, w_full, t * 64, 0);
      tma_load_2d(sW3, &q.w3_map, w_full, 0, 0);
      pdl_wait();
      for (int j = 0; j < nM; ++j) {
        const int st = j % kHS;
        const int m_blk = m_blk_of(j);
        const int iw = m_blk % p.tiles_w, ih = (m_blk / p.tiles_w) % p.tiles_h, in = m_blk / (p.tiles_w * p.tiles_h);
        mbar_wait(&halo_empty[st], ((j / kHS) & 1) ^ 1);
        mbar_expect_tx(&halo_full[st], kC23hHaloBytes);
        tma_load_5d(sHalo + st * kC23hHaloStride, &p.a_map[0], &halo_full[st], 0, iw * p.bw - 1, ih * p.bh - 1, 0, in);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc(1u, kBlockM, P);
      constexpr uint32_t idesc2 = umma_idesc(1u, kBlockM, BLOCK_N);
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t w2_addr = smem_u32(sW2), w3_addr = smem_u32(sW3), a2_addr = smem_u32(sA2);
      for (int s = 0; s < n_sub; ++s) {
        int j;
        bool is_a;
        sub_at(s, j, is_a);
        const int acc = s & 1;
        mbar_wait(&tmem_empty[acc], ((s >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        if (is_a) {
          const int st = j % kHS;
          mbar_wait(&halo_full[st], (j / kHS) & 1);
          tc_fence_after();
          const uint32_t h_addr = smem_u32(sHalo + st * kC23hHaloStride);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_addr = h_addr + ((tap / 3) * kC23hHaloW + (tap % 3)) * 128;
            const uint32_t b_addr = w2_addr + tap * 64 * 128;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, umma_desc_sw128_sbo(a_addr + k * 32, kC23hHaloW * 128), umma_desc_sw128(b_addr + k * 32), idesc1,
                        (tap | k) != 0);
          }
          umma_commit(&halo_empty[st]);
        } else {
          mbar_wait(a2_full, j & 1);                        // the conv3 operand of tile j has been written
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, umma_desc_sw128(a2_addr + k * 32), umma_desc_sw128(w3_addr + k * 32), idesc2, k != 0);
          umma_commit(a2_empty);                            // A2 may be overwritten once these MMAs retire
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ C producer: residual prefetch for the conv3 sub-tiles
    if (elect_one()) {
      pdl_wait();
      const bool has_res = p.residual != nullptr;
      int c_it = 0;
      for (int j = 0; j < nM; ++j) {
        const int m_blk = m_blk_of(j);
        const int iw = m_blk % p.tiles_w, ih = (m_blk / p.tiles_w) % p.tiles_h, in = m_blk / (p.tiles_w * p.tiles_h);
        for (int jj = 0; jj < BLOCK_N / 64; ++jj, ++c_it) {
          const int slot = c_it % kCSlots;
          mbar_wait(&c_empty[slot], ((c_it / kCSlots) & 1) ^ 1);
          if (has_res) {
            mbar_expect_tx(&c_full[slot], kCBytes);
            const int n0 = in * p.nf;
            if (p.res_clip_T == 0)
              tma_load_5d(sC + slot * kCBytes, &p.res_map, &c_full[slot], jj * 64, iw * p.bw, ih * p.bh, 0, n0);
            else
              tma_load_5d(sC + slot * kCBytes, &p.res_map, &c_full[slot], jj * 64, iw * p.bw, ih * p.bh, n0 % p.res_clip_T,
                          n0 / p.res_clip_T);
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
    const int dw = row % p.bw, dh = (row / p.bw) % p.bh, dn = row / (p.bw * p.bh);
    const int HW = p.Ho * p.Wo;
    __nv_bfloat16* tsm = reinterpret_cast<__nv_bfloat16*>(p.tsm_out);
    const bool has_res = p.residual != nullptr;
    const int srow = lane >> 1, spiece = lane & 1;
    int c_it = 0;
    for (int s = 0; s < n_sub; ++s) {
      int j;
      bool is_a;
      sub_at(s, j, is_a);
      const int acc = s & 1;
      mbar_wait(&tmem_full[acc], (s >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      if (is_a) {
        // ---- A: conv2 accumulator -> + bias2 -> ReLU -> bf16 -> A2 (K-major, 128-byte swizzle); group g takes columns
        //      [16 g, 16 g + 16)
        mbar_wait(a2_empty, (j & 1) ^ 1);
        const int c0 = group * 16;
        uint32_t r[16];
        tmem_ld_32x16(taddr + c0, r);
        tmem_ld_wait();
        float2 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(q.bias2 + c0 + 2 * e));
          v[e] = __fadd2_rn(v[e], make_float2(b4.x, b4.y));
          v[e + 1] = __fadd2_rn(v[e + 1], make_float2(b4.z, b4.w));
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = make_float2(fmaxf(v[e].x, 0.f), fmaxf(v[e].y, 0.f));
        uint8_t* arow = sA2 + row * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 o;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __float22bfloat162_rn(v[c * 4 + e]);
          *reinterpret_cast<uint4*>(arow + ((((c0 >> 3) + c) ^ (row & 7)) << 4)) = o;
        }
        fence_proxy_async_smem();                              // generic-proxy writes -> visible to the MMA operand fetch
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(a2_full);
          mbar_arrive(&tmem_empty[acc]);
        }
        continue;
      }
      // ---- B: the conv3 epilogue (group g drains the g-th 64-column C tile)
      const int m_blk = m_blk_of(j);
      const int iw = m_blk % p.tiles_w, ih = (m_blk / p.tiles_w) % p.tiles_h, in = m_blk / (p.tiles_w * p.tiles_h);
      const int w = iw * p.bw + dw, h = ih * p.bh + dh, n = in * p.nf + dn;
      const bool row_ok = (dn < p.nf) && (w < p.Wo) && (h < p.Ho) && (n < p.Nimg);
      const long grow = row_ok ? (static_cast<long>(n) * p.Ho + h) * p.Wo + w : -1;
      long g_i[2];
      int t_i[2];
      if (tsm) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          g_i[i] = __shfl_sync(0xffffffffu, grow, srow + 16 * i);
          t_i[i] = g_i[i] >= 0 ? static_cast<int>((g_i[i] / HW) % p.T) : 0;
        }
      }
      const int my_sub = group;
      const int my_it = c_it + my_sub;
      const int slot = my_it % kCSlots;
      uint8_t* ctile = sC + slot * kCBytes;
      uint8_t* crow = ctile + row * 128;
      uint32_t rr[2][16];
      tmem_ld_32x16(taddr + my_sub * 64, rr[0]);
      mbar_wait(&c_full[slot], (my_it / kCSlots) & 1);
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int cs = pass * 16;
        const int col0 = my_sub * 64 + cs;
        tmem_ld_wait();
        if (pass < 3) {
          tmem_ld_32x16(taddr + my_sub * 64 + cs + 16, rr[(pass + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);      // accumulator handed back right after the last tcgen05.ld
        }
        const uint32_t (&r)[16] = rr[pass & 1];
        float2 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 2 * e));
          v[e] = __fadd2_rn(v[e], make_float2(b4.x, b4.y));
          v[e + 1] = __fadd2_rn(v[e + 1], make_float2(b4.z, b4.w));
        }
        if (has_res) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint4 qq = *reinterpret_cast<const uint4*>(crow + ((((cs >> 3) + c) ^ (row & 7)) << 4));
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&qq);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[c * 4 + e] = __fadd2_rn(v[c * 4 + e], __bfloat1622float2(h2[e]));
          }
        }
        apply_act8x2(v, p.act);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 o;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __float22bfloat162_rn(v[c * 4 + e]);
          *reinterpret_cast<uint4*>(crow + ((((cs >> 3) + c) ^ (row & 7)) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + my_sub), "n"(128) : "memory");
      if (quarter == 0 && lane == 0) {
        tma_store_5d(ctile, &p.out_map, my_sub * 64, iw * p.bw, ih * p.bh, 0, in * p.nf);
        tma_store_commit();
      }
      if (tsm) {
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          const int cs = pass * 16;
          const int col0 = my_sub * 64 + cs;
          const bool zone_a = col0 < p.tsm_fold;
          const bool zone_b = !zone_a && col0 < 2 * p.tsm_fold;
          if (zone_a || zone_b) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int r2 = quarter * 32 + srow + 16 * i;
              if (g_i[i] >= 0) {
                const uint4 o = *reinterpret_cast<const uint4*>(ctile + r2 * 128 + ((((cs >> 3) + spiece) ^ (r2 & 7)) << 4));
                if (zone_a && t_i[i] >= 1)
                  *reinterpret_cast<uint4*>(tsm + (g_i[i] - HW) * p.tsm_ld + col0 + spiece * 8) = o;
                if (zone_b && t_i[i] + 1 < p.T)
                  *reinterpret_cast<uint4*>(tsm + (g_i[i] + HW) * p.tsm_ld + col0 + spiece * 8) = o;
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
