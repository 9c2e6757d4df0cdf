// Layer-2 bottleneck tail (planes P = 128, stride 1 or 2) with the conv3 operand in TENSOR MEMORY.
//   conv2 (3x3) + BN + ReLU -> conv3 (1x1, 128 -> 512) + BN + residual + ReLU, bf16, one persistent CTA per SM.
//
// conv23_kernel<128> is bound by the depth of its TMA ring: three 32 KB stages ([A tap tile | W2 block], or one 256 x 64
// W3 block) carry 704 KB per tile against ~0.9 us of L2 latency, because 64 KB of shared memory hold the double-buffered
// conv3 operand (conv2's output tile, 128 x 128 bf16).  Here that tile never touches shared memory: the conv2 epilogue
// writes it back into tensor memory with tcgen05.st (row = lane, two bf16 per 32-bit column, 64 columns) and conv3 issues
// tcgen05.mma with the A operand read FROM TENSOR MEMORY (layout verified by tools/probes/ts_probe.cu).  The ring grows to
// five stages.  TMEM map (512 columns): conv2 accumulator [0, 128), conv3 operand A2[0] [128, 192) / A2[1] [192, 256),
// conv3 accumulator [256, 512).
// Issue order per tile j:  conv2 K steps 0-8 | B(j-1, 0) | conv2 K steps 9-17 | B(j-1, 1)   (B(t, nb) = conv3 of tile t,
// output channels [256 nb, +256)); the epilogue warps follow in completion order  B(j-1,0), A(j), B(j-1,1).
// Same patch geometry, tap tables and tensor maps as conv23_kernel (conv_gemm_host.h build_conv23).
#pragma once
#include "conv23.cuh"

namespace vcg {

constexpr int kC23tStages = 5;
constexpr int kC23tCSlots = 4;
constexpr int kC23tSmemBytes = kC23tStages * kC23StageBytes + kC23tCSlots * kCBytes + 1024 /*align*/ + 512 /*barriers*/;
static_assert(kC23tSmemBytes <= 232448, "shared memory budget exceeded");

template <int kDummy = 0>
__global__ void __launch_bounds__(kC23Threads, 1) conv23t_kernel(const __grid_constant__ Conv23Params q) {
  const ConvGemmParams& p = q.g;
  constexpr int kStages = kC23tStages, kCSlots = kC23tCSlots;
  constexpr uint32_t kAccA = 0, kA2 = 128, kAccB = 256;    // TMEM columns

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStage = smem;                                   // [stage][32 KB]
  uint8_t* sC = sStage + kStages * kC23StageBytes;          // [kCSlots][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + kCSlots * kCBytes);
  uint64_t* full_bar = bars;                     // [kStages]
  uint64_t* empty_bar = bars + kStages;
  uint64_t* ta_full = empty_bar + kStages;       // conv2 accumulator ready
  uint64_t* ta_empty = ta_full + 1;
  uint64_t* tb_full = ta_full + 2;               // conv3 accumulator ready
  uint64_t* tb_empty = ta_full + 3;
  uint64_t* a2_full = ta_full + 4;               // [2] epilogue A -> MMA: conv3 operand j & 1 written to tensor memory
  uint64_t* a2_empty = ta_full + 6;              // [2] MMA (conv3 retired) -> epilogue A
  uint64_t* c_full = ta_full + 8;                // [kCSlots]
  uint64_t* c_empty = c_full + kCSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_empty + kCSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&q.w3_map);
    tma_prefetch_desc(&p.out_map);
    tma_prefetch_desc(&p.res_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(ta_full, 1); mbar_init(ta_empty, kEpiWarpsBf16);
    mbar_init(tb_full, 1); mbar_init(tb_empty, kEpiWarpsBf16);
    for (int i = 0; i < 2; ++i) { mbar_init(&a2_full[i], kEpiWarpsBf16); mbar_init(&a2_empty[i], 1); }
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
  pdl_wait();

  const int num_kb1 = p.n_taps * p.cpt;                     // 18 conv2 K steps
  const int half1 = (num_kb1 + 1) / 2;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int nM = m_tiles > static_cast<int>(blockIdx.x)
                     ? (m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  auto m_blk_of = [&](int j) { return static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x); };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer, in MMA issue order
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      auto load_taps = [&](int j, int k0, int k1) {
        const int m_blk = m_blk_of(j);
        const int iw = m_blk % p.tiles_w, ih = (m_blk / p.tiles_w) % p.tiles_h, in = m_blk / (p.tiles_w * p.tiles_h);
        const int w0 = iw * p.bw, h0 = ih * p.bh, n0 = in * p.nf;
        for (int kb = k0; kb < k1; ++kb) {
          const int tap = kb / p.cpt, cb = kb - tap * p.cpt;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
          const TapDesc t = p.taps[tap];
          uint8_t* st = sStage + stage * kC23StageBytes;
          tma_load_5d(st, &p.a_map[t.map], &full_bar[stage], cb * 64 + t.c_off, w0 + t.dw, h0 + t.dh, t.plane, n0);
          tma_load_2d(st + kCBytes, &p.b_map, &full_bar[stage], kb * 64, 0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      };
      auto load_b = [&](int nb) {                           // the two 256 x 64 W3 blocks of conv3 sub-tile nb
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], 256u * 128u);
          tma_load_2d(sStage + stage * kC23StageBytes, &q.w3_map, &full_bar[stage], kb * 64, nb * 256);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      };
      for (int j = 0; j < nM; ++j) {
        load_taps(j, 0, half1);
        if (j > 0) load_b(0);
        load_taps(j, half1, num_kb1);
        if (j > 0) load_b(1);
      }
      if (nM > 0) { load_b(0); load_b(1); }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc1 = umma_idesc(1u, kBlockM, 128);
      constexpr uint32_t idesc2 = umma_idesc(1u, kBlockM, 256);
      int stage = 0, nb_count = 0;
      uint32_t phase = 0;
      auto issue_taps = [&](int k0, int k1) {
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sStage + stage * kC23StageBytes);
          const uint32_t b_addr = a_addr + kCBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + kAccA, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc1, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      };
      auto issue_b = [&](int j, int nb) {                   // conv3 of tile j, output channels [256 nb, +256); A from TMEM
        mbar_wait(tb_empty, (nb_count & 1) ^ 1);
        if (nb == 0) mbar_wait(&a2_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t a_tmem = tmem_base + kA2 + (j & 1) * 64;
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(sStage + stage * kC23StageBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ts(tmem_base + kAccB, a_tmem + kb * 32 + k * 8, umma_desc_sw128(b_addr + k * 32), idesc2, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (nb == 1) umma_commit(&a2_empty[j & 1]);         // the conv3 operand of tile j may be overwritten
        umma_commit(tb_full);
        ++nb_count;
      };
      for (int j = 0; j < nM; ++j) {
        mbar_wait(ta_empty, (j & 1) ^ 1);
        tc_fence_after();
        issue_taps(0, half1);
        if (j > 0) issue_b(j - 1, 0);
        issue_taps(half1, num_kb1);
        umma_commit(ta_full);
        if (j > 0) issue_b(j - 1, 1);
      }
      if (nM > 0) { issue_b(nM - 1, 0); issue_b(nM - 1, 1); }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ C producer: residual prefetch for the conv3 sub-tiles
    if (elect_one()) {
      const bool has_res = p.residual != nullptr;
      // the residual boxes of the NEXT tile are prefetched into L2 (no shared-memory destination): a C slot only frees up
      // when the previous sub-tile has been stored, so without it every sub-tile exposes a full DRAM round trip while the
      // epilogue holds the conv3 accumulator
      const int kPF = q.prefetch_tiles;
      auto prefetch_tile = [&](int j2) {
        if (!has_res || j2 >= nM) return;
        const int m_blk = m_blk_of(j2);
        const int iw = m_blk % p.tiles_w, ih = (m_blk / p.tiles_w) % p.tiles_h, in = m_blk / (p.tiles_w * p.tiles_h);
        const int n0 = in * p.nf;
        for (int jj = 0; jj < 8; ++jj) {
          if (p.res_clip_T == 0) tma_prefetch_l2_5d(&p.res_map, jj * 64, iw * p.bw, ih * p.bh, 0, n0);
          else tma_prefetch_l2_5d(&p.res_map, jj * 64, iw * p.bw, ih * p.bh, n0 % p.res_clip_T, n0 / p.res_clip_T);
        }
      };
      for (int j2 = 0; j2 < kPF; ++j2) prefetch_tile(j2);
      int c_it = 0;
      for (int j = 0; j < nM; ++j) {
        if (kPF > 0) prefetch_tile(j + kPF);
        const int m_blk = m_blk_of(j);
        const int iw = m_blk % p.tiles_w, ih = (m_blk / p.tiles_w) % p.tiles_h, in = m_blk / (p.tiles_w * p.tiles_h);
        for (int jj = 0; jj < 8; ++jj, ++c_it) {            // nb = jj / 4, 64-column C tile jj % 4
          const int slot = c_it % kCSlots;
          mbar_wait(&c_empty[slot], ((c_it / kCSlots) & 1) ^ 1);
          if (has_res) {
            mbar_expect_tx(&c_full[slot], p.a_bytes);
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
    const int quarter = warp & 3;
    const int group = (warp - kFirstEpiWarp) >> 2;
    const int row = quarter * 32 + lane;
    const bool has_res = p.residual != nullptr;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t c_addr = smem_u32(sC);
    int c_it = 0, nb_count = 0;
    auto epi_b = [&](int j, int nb) {
      mbar_wait(tb_full, nb_count & 1);
      ++nb_count;
      tc_fence_after();
      const int m_blk = m_blk_of(j);
      const int iw = m_blk % p.tiles_w, ih = (m_blk / p.tiles_w) % p.tiles_h, in = m_blk / (p.tiles_w * p.tiles_h);
      const int my_it = c_it + group;
      const int slot = my_it % kCSlots;
      const uint32_t ctile_s = c_addr + slot * kCBytes;
      const uint32_t crow_s = ctile_s + row * 128;
      const uint32_t taddr = lane_addr + kAccB + group * 64;
      uint32_t rr[2][16];
      tmem_ld_32x16(taddr, rr[0]);
      mbar_wait(&c_full[slot], (my_it / kCSlots) & 1);
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int cs = pass * 16;
        const int col0 = nb * 256 + group * 64 + cs;
        tmem_ld_wait();
        if (pass < 3) {
          tmem_ld_32x16(taddr + cs + 16, rr[(pass + 1) & 1]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tb_empty);              // accumulator handed back right after the last tcgen05.ld
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
            const uint4 qq = lds128(crow_s + ((((cs >> 3) + c) ^ (row & 7)) << 4));
            v[c * 4 + 0] = __fadd2_rn(v[c * 4 + 0], unpack_bf16x2(qq.x));
            v[c * 4 + 1] = __fadd2_rn(v[c * 4 + 1], unpack_bf16x2(qq.y));
            v[c * 4 + 2] = __fadd2_rn(v[c * 4 + 2], unpack_bf16x2(qq.z));
            v[c * 4 + 3] = __fadd2_rn(v[c * 4 + 3], unpack_bf16x2(qq.w));
          }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
          sts128(crow_s + ((((cs >> 3) + c) ^ (row & 7)) << 4), pack_relu_bf16x2(v[c * 4].x, v[c * 4].y),
                 pack_relu_bf16x2(v[c * 4 + 1].x, v[c * 4 + 1].y), pack_relu_bf16x2(v[c * 4 + 2].x, v[c * 4 + 2].y),
                 pack_relu_bf16x2(v[c * 4 + 3].x, v[c * 4 + 3].y));
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(128) : "memory");
      if (quarter == 0 && lane == 0) {
        tma_store_5d(reinterpret_cast<const void*>(sC + slot * kCBytes), &p.out_map, nb * 256 + group * 64, iw * p.bw, ih * p.bh, 0,
                     in * p.nf);
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
      if (j > 0) epi_b(j - 1, 0);
      // ---- A(j): conv2 accumulator -> + bias2 -> bf16 -> ReLU -> tensor memory (conv3's A operand): group g takes channels
      //      [32 g, 32 g + 32) = A2 columns [16 g, 16 g + 16)
      mbar_wait(ta_full, j & 1);
      tc_fence_after();
      uint32_t r[2][16];
      tmem_ld_32x16(lane_addr + kAccA + group * 32, r[0]);
      tmem_ld_32x16(lane_addr + kAccA + group * 32 + 16, r[1]);
      mbar_wait(&a2_empty[j & 1], ((j >> 1) & 1) ^ 1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ta_empty);                  // conv2 accumulator drained
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(q.bias2 + group * 32 + hh * 16 + e * 4));
          o[2 * e] = pack_relu_bf16x2(__uint_as_float(r[hh][4 * e]) + b4.x, __uint_as_float(r[hh][4 * e + 1]) + b4.y);
          o[2 * e + 1] = pack_relu_bf16x2(__uint_as_float(r[hh][4 * e + 2]) + b4.z, __uint_as_float(r[hh][4 * e + 3]) + b4.w);
        }
        tmem_st_32x8(lane_addr + kA2 + (j & 1) * 64 + group * 16 + hh * 8, o);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a2_full[j & 1]);
      if (j > 0) epi_b(j - 1, 1);
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
