// Fused bottleneck tail: conv2 (3x3, stride 1/2) + BN + ReLU  ->  conv3 (1x1) + BN + residual + ReLU (+ TSM scatter)
// in ONE persistent tcgen05 kernel (bf16, planes P = 64 or 128: ResNet layer1 / layer2).
//
// Why: conv3 of layer1/2 is HBM-bound (residual read + output write, 4 MB per frame and block) while conv2 is
// tensor-bound, and conv2's output tile is exactly conv3's A operand.  Fused, the 128 x P tile never leaves the SM
// (-0.8 MB of HBM traffic per frame and block) and, more importantly, conv3's memory traffic overlaps conv2's math.
//
// One M tile (128 output pixels, the patch geometry of conv_gemm.cuh) becomes 1 + n2 "sub-tiles":
//   A(j)      GEMM1: 9 taps x P/64 K blocks of the conv2 input patch against W2 [P, 9P] into a TMEM accumulator (P columns);
//             the epilogue adds the folded BN bias, applies ReLU and writes bf16 into the shared-memory tile A2[j & 1] in
//             the swizzled K-major operand layout (what a TMA load of conv2's output would have produced)
//   B(j, n)   GEMM2: P/64 K blocks of A2[j & 1] against W3 [4P, P], 256 output channels per sub-tile; the epilogue is the
//             ordinary conv3 one: + bias + residual (TMA-prefetched into the C ring) -> ReLU -> bf16 -> TMA store, and the
//             first channels scattered into the next bottleneck's temporally shifted input
// issued as  A(0), A(1), B(0,*), A(2), B(1,*), ...  so that the tensor core works on the next tile's conv2 while the
// epilogue warps turn the current one into the conv3 operand.  Two TMEM accumulators of 256 columns alternate over
// the sub-tiles; stages are 32 KB ([A patch 16 KB | W2 block] or one 256 x 64 W3 block).
#pragma once
#include "conv_gemm.cuh"

namespace vcg {

struct Conv23Params {
  ConvGemmParams g;        // conv2 geometry / maps (b_map = W2, box {64, P}); out / residual / tsm / bias = conv3's
  CUtensorMap w3_map;      // W3 [4P, P], box {64, 256}
  const float* bias2;      // conv2's folded BN bias [P]
  int P;                   // planes (64 or 128)
  int n2;                  // 4P / 256 output-channel sub-tiles of conv3
  int n_stages, n_cslots;
  int early_release;       // hand the TMEM buffer back right after the last tcgen05.ld of a B sub-tile (VCG_C23_EARLY)
  int prefetch_tiles;      // conv23h: tiles of L2 prefetch distance (VCG_C23H_PF, default 1; 0 = off)
  uint32_t magic_tpi, magic_tw;   // conv23h: floor(2^32 / d) + 1 for d = tiles per image, tiles per row (exact __umulhi division)
};

constexpr int kC23Threads = (kFirstEpiWarp + kEpiWarpsBf16) * 32;
constexpr int kC23StageBytes = 32 * 1024;
constexpr int kC23Budget = 224 * 1024;
constexpr int kC23SmemBytes = kC23Budget + 1024 + 512;

template <int P>
__global__ void __launch_bounds__(kC23Threads, 1) conv23_kernel(const __grid_constant__ Conv23Params q) {
  const ConvGemmParams& p = q.g;
  constexpr int BLOCK_N = 256;
  const int kStages = q.n_stages, kCSlots = q.n_cslots;
  constexpr int kb2 = P / 64;                               // K blocks of GEMM2
  constexpr int a2_bytes = kb2 * kCBytes;
  const int n2 = q.n2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStage = smem;                                   // [stage][32 KB]
  uint8_t* sA2 = sStage + kStages * kC23StageBytes;         // [2][kb2][128 rows][128 B]
  uint8_t* sC = sA2 + 2 * a2_bytes;                         // [kCSlots][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kC23Budget);
  uint64_t* full_bar = bars;                     // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint64_t* a2_full = tmem_empty + 2;            // [2] epilogue A -> MMA
  uint64_t* a2_empty = a2_full + 2;              // [2] MMA (GEMM2 retired) -> epilogue A
  uint64_t* c_full = a2_empty + 2;               // [kCSlots]
  uint64_t* c_empty = c_full + kMaxCSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c_empty + kMaxCSlots);

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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarpsBf16);
      mbar_init(&a2_full[i], kEpiWarpsBf16); mbar_init(&a2_empty[i], 1);
    }
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

  const int num_kb1 = p.n_taps * p.cpt;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  const int nM = m_tiles > static_cast<int>(blockIdx.x)
                     ? (m_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  const int n_sub = nM * (1 + n2);
  // s-th sub-tile of this CTA: A(0); A(j), B(j-1, 0..n2-1) for j = 1..nM-1; B(nM-1, 0..n2-1)
  auto sub_at = [&](int s, int& j, int& nb) {     // nb < 0: A sub-tile
    if (s == 0) { j = 0; nb = -1; return; }
    const int t = s - 1, blk = t / (1 + n2), r = t - blk * (1 + n2);
    if (blk < nM - 1) {
      if (r == 0) { j = blk + 1; nb = -1; } else { j = blk; nb = r - 1; }
    } else {
      j = nM - 1; nb = t - (nM - 1) * (1 + n2);
    }
  };
  auto m_blk_of = [&](int j) { return static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x); };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < n_sub; ++s) {
        int j, nb;
        sub_at(s, j, nb);
        if (nb < 0) {
          const int m_blk = m_blk_of(j);
          const int iw = m_blk % p.tiles_w, ih = (m_blk / p.tiles_w) % p.tiles_h, in = m_blk / (p.tiles_w * p.tiles_h);
          const int w0 = iw * p.bw, h0 = ih * p.bh, n0 = in * p.nf;
          int tap = 0, cb = 0;
          for (int kb = 0; kb < num_kb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], p.a_bytes + p.b_bytes);
            const TapDesc t = p.taps[tap];
            uint8_t* st = sStage + stage * kC23StageBytes;
            tma_load_5d(st, &p.a_map[t.map], &full_bar[stage], cb * 64 + t.c_off, w0 + t.dw, h0 + t.dh, t.plane, n0);
            tma_load_2d(st + kCBytes, &p.b_map, &full_bar[stage], kb * 64, 0);
            if (++cb == p.cpt) { cb = 0; ++tap; }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        } else {
          for (int kb = 0; kb < kb2; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], BLOCK_N * 128u);
            tma_load_2d(sStage + stage * kC23StageBytes, &q.w3_map, &full_bar[stage], kb * 64, nb * BLOCK_N);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc1 = umma_idesc(1u, kBlockM, 0) | (static_cast<uint32_t>(P >> 3) << 17);
      constexpr uint32_t idesc2 = umma_idesc(1u, kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < n_sub; ++s) {
        int j, nb;
        sub_at(s, j, nb);
        const int acc = s & 1;
        mbar_wait(&tmem_empty[acc], ((s >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        if (nb < 0) {
          for (int kb = 0; kb < num_kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sStage + stage * kC23StageBytes);
            const uint32_t b_addr = a_addr + kCBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc1, (kb | k) != 0);
            umma_commit(&empty_bar[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        } else {
          if (nb == 0) {                                    // the conv3 operand of tile j has been written
            mbar_wait(&a2_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
          }
          for (int kb = 0; kb < kb2; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA2 + (j & 1) * a2_bytes + kb * kCBytes);
            const uint32_t b_addr = smem_u32(sStage + stage * kC23StageBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc2, (kb | k) != 0);
            umma_commit(&empty_bar[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          if (nb == n2 - 1) umma_commit(&a2_empty[j & 1]);   // A2[j & 1] may be overwritten once these MMAs retire
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ C producer: residual prefetch for the B sub-tiles
    if (elect_one()) {
      const bool has_res = p.residual != nullptr;
      int c_it = 0;
      for (int s = 0; s < n_sub; ++s) {
        int j, nb;
        sub_at(s, j, nb);
        if (nb < 0) continue;
        const int m_blk = m_blk_of(j);
        const int iw = m_blk % p.tiles_w, ih = (m_blk / p.tiles_w) % p.tiles_h, in = m_blk / (p.tiles_w * p.tiles_h);
        for (int jj = 0; jj < BLOCK_N / 64; ++jj, ++c_it) {
          const int slot = c_it % kCSlots;
          mbar_wait(&c_empty[slot], ((c_it / kCSlots) & 1) ^ 1);
          if (has_res) {
            mbar_expect_tx(&c_full[slot], p.a_bytes);
            const int n0 = in * p.nf;
            if (p.res_clip_T == 0)
              tma_load_5d(sC + slot * kCBytes, &p.res_map, &c_full[slot], nb * BLOCK_N + jj * 64, iw * p.bw, ih * p.bh, 0, n0);
            else
              tma_load_5d(sC + slot * kCBytes, &p.res_map, &c_full[slot], nb * BLOCK_N + jj * 64, iw * p.bw, ih * p.bh,
                          n0 % p.res_clip_T, n0 / p.res_clip_T);
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
    const int dw = row % p.bw, dh = (row / p.bw) % p.bh, dn = row / (p.bw * p.bh);
    const int HW = p.Ho * p.Wo;
    __nv_bfloat16* tsm = reinterpret_cast<__nv_bfloat16*>(p.tsm_out);
    const bool has_res = p.residual != nullptr;
    const int srow = lane >> 1, spiece = lane & 1;
    int c_it = 0;
    for (int s = 0; s < n_sub; ++s) {
      int j, nb;
      sub_at(s, j, nb);
      const int acc = s & 1;
      mbar_wait(&tmem_full[acc], (s >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      if (nb < 0) {
        // ---- A: conv2 accumulator -> + bias2 -> ReLU -> bf16 -> A2[j & 1] (K-major, 128-byte swizzle); the 16 warps split
        //      the P columns (P/4 per group)
        mbar_wait(&a2_empty[j & 1], ((j >> 1) & 1) ^ 1);
        const uint32_t a2 = smem_u32(sA2 + (j & 1) * a2_bytes);
        constexpr int cpg = P / 4;                               // 16 or 32 columns per group
        for (int c0 = group * cpg; c0 < (group + 1) * cpg; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x16(taddr + c0, r);
          tmem_ld_wait();
          uint32_t o[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(q.bias2 + c0 + 4 * e));
            o[2 * e] = pack_relu_bf16x2(__uint_as_float(r[4 * e]) + b4.x, __uint_as_float(r[4 * e + 1]) + b4.y);
            o[2 * e + 1] = pack_relu_bf16x2(__uint_as_float(r[4 * e + 2]) + b4.z, __uint_as_float(r[4 * e + 3]) + b4.w);
          }
          const uint32_t arow = a2 + (c0 >> 6) * kCBytes + row * 128;
          const uint32_t ch = static_cast<uint32_t>((c0 & 63) >> 3);
          sts128(arow + (((ch) ^ (row & 7)) << 4), o[0], o[1], o[2], o[3]);
          sts128(arow + (((ch + 1u) ^ (row & 7)) << 4), o[4], o[5], o[6], o[7]);
        }
        fence_proxy_async_smem();                              // generic-proxy writes -> visible to the MMA operand fetch
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a2_full[j & 1]);
          mbar_arrive(&tmem_empty[acc]);
        }
        continue;
      }
      // ---- B: the conv3 epilogue (conv_gemm.cuh, BLOCK_N = 256: group g drains the g-th 64-column C tile)
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
      const uint32_t ctile_s = smem_u32(ctile);
      const uint32_t crow_s = ctile_s + row * 128;
      const bool relu_fast = p.act == ACT_RELU;
      // passes software-pipelined as in conv_gemm.cuh: next tcgen05.ld in flight during the math, TMEM handed back to
      // the MMA warp right after the last load
      uint32_t rr[2][16];
      tmem_ld_32x16(taddr + my_sub * 64, rr[0]);
      mbar_wait(&c_full[slot], (my_it / kCSlots) & 1);
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int cs = pass * 16;
        const int col0 = nb * BLOCK_N + my_sub * 64 + cs;
        tmem_ld_wait();
        if (pass < 3) {
          tmem_ld_32x16(taddr + my_sub * 64 + cs + 16, rr[(pass + 1) & 1]);
        } else if (q.early_release) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
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
        if (relu_fast) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            sts128(crow_s + ((((cs >> 3) + c) ^ (row & 7)) << 4), pack_relu_bf16x2(v[c * 4].x, v[c * 4].y),
                   pack_relu_bf16x2(v[c * 4 + 1].x, v[c * 4 + 1].y), pack_relu_bf16x2(v[c * 4 + 2].x, v[c * 4 + 2].y),
                   pack_relu_bf16x2(v[c * 4 + 3].x, v[c * 4 + 3].y));
        } else {
          apply_act8x2(v, p.act);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            sts128(crow_s + ((((cs >> 3) + c) ^ (row & 7)) << 4), pack_bf16x2(v[c * 4].x, v[c * 4].y),
                   pack_bf16x2(v[c * 4 + 1].x, v[c * 4 + 1].y), pack_bf16x2(v[c * 4 + 2].x, v[c * 4 + 2].y),
                   pack_bf16x2(v[c * 4 + 3].x, v[c * 4 + 3].y));
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + my_sub), "n"(128) : "memory");
      if (quarter == 0 && lane == 0) {
        tma_store_5d(ctile, &p.out_map, nb * BLOCK_N + my_sub * 64, iw * p.bw, ih * p.bh, 0, in * p.nf);
        tma_store_commit();
      }
      if (tsm) {
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          const int cs = pass * 16;
          const int col0 = nb * BLOCK_N + my_sub * 64 + cs;
          const bool zone_a = col0 < p.tsm_fold;
          const bool zone_b = !zone_a && col0 < 2 * p.tsm_fold;
          if (zone_a || zone_b) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int rr = quarter * 32 + srow + 16 * i;
              if (g_i[i] >= 0) {
                const uint4 o = lds128(ctile_s + rr * 128 + ((((cs >> 3) + spiece) ^ (rr & 7)) << 4));
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
      if (!q.early_release) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      }
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
