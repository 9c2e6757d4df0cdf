// BERT self-attention on the 5th-generation tensor cores (bf16, token-packed layout, at most 128 tokens per clip).
//
//   ctx = softmax(Q K^T / 8 + key_mask) V        per (clip, head); modeling_bert.py:115-140 (12 heads x 64)
//
// A work item is one (clip, head): at L <= 128 all its queries and keys fit one 128 x 128 score tile, so there is no
// online-softmax loop.  The kernel is persistent (one CTA per SM walks items blockIdx.x, +gridDim.x, ...) and
// warp-specialised, so the latencies that bound the mma.sync kernel (global -> shared staging, ldmatrix chains,
// 3 CTAs per SM) are hidden by a pipeline instead of by occupancy:
//   warp 0      TMA producer: Q, K, V rows of the item (32-row boxes of the packed [rows, 2304] QKV matrix, only
//               ceil(L/32) of them) into a 3-stage ring; publishes (first row, L) next to the stage
//   warp 1      MMA issuer: S = Q K^T (tcgen05.mma 128 x Nk x 64, Nk = L rounded up to 16) into one of two TMEM score
//               buffers, and, once the softmax warps have written P, O = P V (128 x 64 x Nk) with V consumed in place
//               as an MN-major operand (rows of 128 B, exactly what TMA delivered); S of item i+1 is issued before
//               P V of item i
//   warp 2      TMEM allocator
//   warps 4-7 / 8-11   two softmax groups (even / odd items), one thread per query row: tcgen05.ld the score row twice
//               (row max, then exp2 / row sum / bf16 P into the swizzled K-major A-operand layout), later tcgen05.ld the
//               O row, scale by 1/sum and store 128 contiguous bytes of ctx.  A thread owns its row: no shuffles.
// Keys beyond L and masked keys are excluded by SELECT (not by adding -inf), so stale shared-memory rows can never
// poison a valid row; V rows in [L, Nk) come from global memory (other clips' rows or the zero-initialised slack) and
// are multiplied by p = 0 exactly.
#include "kernels.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include "tensormap.h"
#include <cuda_bf16.h>
#include <map>

namespace vcg {

namespace {

constexpr int kStages = 3;
constexpr int kTileBytes = 128 * 128;             // 128 rows x 64 bf16
constexpr int kStageBytes = 3 * kTileBytes;       // Q, K, V
constexpr int kPBytes = 2 * kTileBytes;           // P: 128 rows x 128 keys bf16 = two 64-key K blocks
constexpr int kSmemTiles = kStages * kStageBytes + 2 * kPBytes;
constexpr int kSmemBytes = kSmemTiles + 1024 /*align*/ + 2 * 128 * 4 /*masks*/ + 256 /*barriers, info*/;
constexpr int kThreads = 12 * 32;
constexpr uint32_t kTmemCols = 512;               // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

struct ItemInfo {
  int row_base, L;
};

struct AttnParams {
  CUtensorMap qkv_map;        // [rows, 2304] bf16, box {64, 32}
  const int32_t* cu;          // [B + 1]
  const uint8_t* key_ok;      // [rows]
  __nv_bfloat16* ctx;         // [rows, 768]
  int n_items;                // B * 12
};

__global__ void __launch_bounds__(kThreads, 1) bert_attention_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStage = smem;                                   // [stage][Q | K | V][128 rows][128 B]
  uint8_t* sP = smem + kStages * kStageBytes;               // [buf][2 K blocks][128 rows][128 B]
  float* sMask = reinterpret_cast<float*>(smem + kSmemTiles);   // per softmax group: 128 key bits (1 = may be attended)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMask + 2 * 128);
  uint64_t* full = bars;                 // [kStages] TMA -> MMA / softmax
  uint64_t* empty = full + kStages;      // [kStages] MMA -> TMA
  uint64_t* s_full = empty + kStages;    // [2] MMA -> softmax
  uint64_t* p_full = s_full + 2;         // [2] softmax -> MMA
  uint64_t* o_full = p_full + 2;         // [2] MMA -> softmax
  uint64_t* t_empty = o_full + 2;        // [2] softmax -> MMA (S and O buffers drained)
  ItemInfo* info = reinterpret_cast<ItemInfo*>(t_empty + 2);   // [kStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(info + kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) tma_prefetch_desc(&p.qkv_map);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&t_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int n_my = (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      for (int i = 0; i < n_my; ++i) {
        const int item = blockIdx.x + i * gridDim.x;
        const int b = item / kBertHeads, h = item - b * kBertHeads;
        const int stage = i % kStages;
        const int row_base = __ldg(p.cu + b);
        const int L = min(__ldg(p.cu + b + 1) - row_base, 128);
        mbar_wait(&empty[stage], ((i / kStages) & 1) ^ 1);
        info[stage] = ItemInfo{row_base, L};
        const int nch = (L + 31) >> 5;
        mbar_expect_tx(&full[stage], static_cast<uint32_t>(3 * nch) * 4096u);
        uint8_t* dst = sStage + stage * kStageBytes;
        for (int c = 0; c < nch; ++c) {
#pragma unroll
          for (int op = 0; op < 3; ++op)
            tma_load_2d(dst + op * kTileBytes + c * 4096, &p.qkv_map, &full[stage], op * kBertHidden + h * 64, row_base + c * 32);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_base = umma_idesc(1u, 128, 0);
      auto issue_s = [&](int i) {
        const int stage = i % kStages, buf = i & 1;
        mbar_wait(&full[stage], (i / kStages) & 1);
        mbar_wait(&t_empty[buf], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const int Nk = (info[stage].L + 15) & ~15;
        const uint32_t idesc = idesc_base | (static_cast<uint32_t>(Nk >> 3) << 17);
        const uint32_t q_addr = smem_u32(sStage + stage * kStageBytes);
        const uint32_t k_addr = q_addr + kTileBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + buf * 128, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idesc, k != 0);
        umma_commit(&s_full[buf]);
      };
      if (n_my > 0) issue_s(0);
      for (int i = 0; i < n_my; ++i) {
        if (i + 1 < n_my) issue_s(i + 1);
        const int stage = i % kStages, buf = i & 1;
        const int Nk = (info[stage].L + 15) & ~15;
        mbar_wait(&p_full[buf], (i >> 1) & 1);
        tc_fence_after();
        // O = P V: A = P (K-major, 64-key K blocks), B = V in place, MN-major (d contiguous), 8-key atoms 1024 B apart
        constexpr uint32_t idesc_pv = umma_idesc(1u, 128, 64) | (1u << 16);
        const uint32_t p_addr = smem_u32(sP + buf * kPBytes);
        const uint32_t v_addr = smem_u32(sStage + stage * kStageBytes + 2 * kTileBytes);
        for (int j = 0; j < Nk / 16; ++j)
          umma_bf16(tmem_base + 256 + buf * 64, umma_desc_sw128(p_addr + (j >> 2) * kTileBytes + (j & 3) * 32),
                    umma_desc_sw128(v_addr + j * 2048), idesc_pv, j != 0);
        umma_commit(&empty[stage]);     // Q, K, V of this stage are free once these MMAs retire
        umma_commit(&o_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax + output, one thread per query row
    const int group = (warp - 4) >> 2;            // even / odd items
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32)
    const int row = quarter * 32 + lane;
    const float sl2 = 0.125f * 1.4426950408889634f;        // 1/sqrt(64) * log2(e)
    uint32_t* mbits = reinterpret_cast<uint32_t*>(sMask) + group * 4;   // 128 key bits of the group's current item
    for (int i = group; i < n_my; i += 2) {
      const int item = blockIdx.x + i * gridDim.x;
      const int h = item % kBertHeads;
      const int stage = i % kStages, buf = i & 1;           // buf == group
      mbar_wait(&full[stage], (i / kStages) & 1);            // (row_base, L) published by the producer
      const ItemInfo it = info[stage];
      const int Nk = (it.L + 15) & ~15;
      {   // key bits: warp `quarter` of the group covers keys [32*quarter, +32)
        const int j = quarter * 32 + lane;
        const uint32_t bits = __ballot_sync(0xffffffffu, j < it.L && p.key_ok[it.row_base + j] != 0);
        if (lane == 0) mbits[quarter] = bits;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
      const uint4 kb4 = *reinterpret_cast<const uint4*>(mbits);
      const uint32_t kbits[4] = {kb4.x, kb4.y, kb4.z, kb4.w};
      mbar_wait(&s_full[buf], (i >> 1) & 1);
      tc_fence_after();
      const bool warp_active = quarter * 32 < it.L;          // (warp-uniform) at least one valid row
      float l = 0.f;
      if (warp_active) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * 128;
        uint8_t* prow = sP + buf * kPBytes + row * 128;
        // 64 score columns per TMEM round trip (four tcgen05.ld, one wait); columns >= Nk are never touched
        auto load64 = [&](int c0, uint32_t (&r)[4][16]) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (c0 + q * 16 < Nk) tmem_ld_32x16(taddr + c0 + q * 16, r[q]);
          tmem_ld_wait();
        };
        auto row_max = [&](int c0, const uint32_t (&r)[4][16], float m) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (c0 + q * 16 < Nk) {
              const uint32_t w = kbits[(c0 + q * 16) >> 5] >> ((c0 + q * 16) & 31);   // 16 key bits of this chunk
              if ((w & 0xffffu) == 0xffffu) {
#pragma unroll
                for (int e = 0; e < 16; ++e) m = fmaxf(m, __uint_as_float(r[q][e]));
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) m = fmaxf(m, (w >> e) & 1u ? __uint_as_float(r[q][e]) : -INFINITY);
              }
            }
          }
          return m;
        };
        auto exp_store = [&](int c0, const uint32_t (&r)[4][16], float msub) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * 16;
            if (c < Nk) {
              const uint32_t w = kbits[c >> 5] >> (c & 31);
              float pv[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                float x;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(fmaf(__uint_as_float(r[q][e]), sl2, -msub)));
                pv[e] = (w >> e) & 1u ? x : 0.f;      // select, not add: stale rows beyond L may hold anything
                l += pv[e];
              }
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                uint4 o;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(pv[t * 8 + 2 * e], pv[t * 8 + 2 * e + 1]);
                const int chunk = ((c & 63) >> 3) + t;     // 16-byte chunk inside the 128-byte row of the K block
                *reinterpret_cast<uint4*>(prow + (c >> 6) * kTileBytes + ((chunk ^ (row & 7)) << 4)) = o;
              }
            }
          }
        };
        uint32_t r[4][16];
        load64(0, r);
        float m = row_max(0, r, -INFINITY);
        if (Nk > 64) {
          load64(64, r);
          m = row_max(64, r, m);
        }
        const float msub = (m == -INFINITY) ? 0.f : m * sl2;
        if (Nk > 64) {
          exp_store(64, r, msub);       // the upper half is still in registers
          load64(0, r);
        }
        exp_store(0, r, msub);
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[buf]);
      // ---- output
      mbar_wait(&o_full[buf], (i >> 1) & 1);
      tc_fence_after();
      if (warp_active) {
        const uint32_t oaddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256 + buf * 64;
        const float inv = 1.0f / l;
        __nv_bfloat16* dst = p.ctx + static_cast<long>(it.row_base + row) * kBertHidden + h * 64;
        uint32_t r[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_32x16(oaddr + q * 16, r[q]);
        tmem_ld_wait();
        if (row < it.L) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              uint4 o;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e)
                h2[e] = __floats2bfloat162_rn(__uint_as_float(r[q][t * 8 + 2 * e]) * inv, __uint_as_float(r[q][t * 8 + 2 * e + 1]) * inv);
              *reinterpret_cast<uint4*>(dst + q * 16 + t * 8) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// Packed layout only (cu, key_ok), bf16, every clip at most 128 tokens.  `rows` = rows of the qkv allocation that may be
// read (at least cu[B] + 31; rows beyond the packed ones must hold finite values).
void launch_bert_attention_tc(const void* qkv, const int32_t* cu, const uint8_t* key_ok, void* ctx, int B, long rows,
                              cudaStream_t s) {
  if (B == 0) return;
  static bool configured = false;
  if (!configured) {
    VCG_CUDA(cudaFuncSetAttribute(bert_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  static std::map<std::pair<const void*, long>, CUtensorMap> maps;
  auto key = std::make_pair(qkv, rows);
  auto it = maps.find(key);
  if (it == maps.end()) {
    const uint64_t dims[2] = {3 * kBertHidden, static_cast<uint64_t>(rows)};
    const uint64_t str[1] = {3 * kBertHidden * 2};
    const uint32_t box[2] = {64, 32};
    it = maps.emplace(key, make_tensor_map(qkv, false, 2, dims, str, box)).first;
  }
  AttnParams p;
  p.qkv_map = it->second;
  p.cu = cu; p.key_ok = key_ok; p.ctx = static_cast<__nv_bfloat16*>(ctx);
  p.n_items = B * kBertHeads;
  int dev = 0, sms = 0;
  VCG_CUDA(cudaGetDevice(&dev));
  VCG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(p.n_items, sms);
  launch_pdl(bert_attention_tc_kernel, grid, kThreads, kSmemBytes, s, p);
}

}  // namespace vcg
