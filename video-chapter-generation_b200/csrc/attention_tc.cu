// BERT self-attention on the 5th-generation tensor cores (bf16, token-packed layout, at most 128 tokens per clip).
//
//   ctx = softmax(Q K^T / 8 + key_mask) V        per (clip, head); modeling_bert.py:115-140 (12 heads x 64)
//
// A work item is one clip and 1, 2 or 4 of its heads (attn_items_kernel): at L <= 128 all queries and keys of a head
// fit one 128 x 128 score tile, so there is no online-softmax loop, and short clips stack several heads in one tile.  An item is tiny (~2 MFLOP); what bounds the kernel is the chain of hand-offs
//   TMA -> S = Q K^T -> softmax -> O = P V -> output        (about 3 k cycles of barrier / MMA-completion latency)
// so the kernel is persistent (one CTA per SM walks items blockIdx.x, +gridDim.x, ...), warp-specialised and keeps
// THREE items in compute plus one in flight from memory:
//   warp 0      TMA producer: Q, K, V rows of the item (32-row boxes of the packed [rows, 2304] QKV matrix, only
//               ceil(L/32) of them) into one of four 48 KB slots; its lanes also gather the item's key-mask bits
//   warp 1      S issuer: tcgen05.mma 128 x Nk x 64 (Nk = L rounded up to 16) into the slot's 128 TMEM columns
//   warp 2      TMEM allocator, then O issuer: once the softmax warps have written P, O = P V (128 x 64 x Nk) with V
//               consumed in place as an MN-major operand (rows of 128 B, exactly what TMA delivered).  O re-uses the
//               first 64 TMEM columns of S, and P re-uses the shared memory of Q and K (both dead by then)
//   warps 4-15  three softmax groups (items i = g, g+3, ...), one thread per query row: tcgen05.ld the score row twice
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

constexpr int kSlots = 4;                         // items resident in shared / tensor memory
constexpr int kGroups = 3;                        // softmax groups = items in compute
constexpr int kTileBytes = 128 * 128;             // 128 rows x 64 bf16
constexpr int kSlotBytes = 3 * kTileBytes;        // Q | K | V; P (128 rows x 128 keys = two 64-key K blocks) overwrites Q | K
constexpr int kSmemTiles = kSlots * kSlotBytes;
constexpr int kSmemBytes = kSmemTiles + 1024 /*align*/ + 512 /*barriers, item info*/;
constexpr int kThreads = (4 + 4 * kGroups) * 32;
constexpr uint32_t kTmemCols = 512;               // slot s: S in columns [128 s, +Nk), later O in [128 s, +64)

struct ItemInfo {
  int row_base, L;
  uint32_t key_bits[4];   // bit j: key j of the clip may be attended (j < L and key_ok)
  int head0, G;           // the item covers heads head0 .. head0 + G - 1 of the clip (G = 1, 2 or 4)
};

struct AttnParams {
  CUtensorMap qkv_map;        // [rows, 2304] bf16, box {64, 32}
  const int32_t* cu;          // [B + 1]
  const uint8_t* key_ok;      // [rows]
  __nv_bfloat16* ctx;         // [rows, 768]
  const int2* items;          // work list: x = clip, y = head0 | G << 8
  const int32_t* n_items;     // its length (device)
};

// Work list: a clip of L <= 32 tokens packs FOUR heads into one 128-row tile (rows / keys [32 g, 32 g + L) belong to head
// head0 + g), 32 < L <= 64 packs TWO, longer clips one.  S = Q K^T of the stacked tile has the wanted per-head score
// blocks on its diagonal (the cross-head blocks are never read), P is written block-diagonal (zeros elsewhere) and
// O = P V then yields every head's output in its own rows: G heads for the barrier / MMA latency chain of one.
// The list is sorted by clip length, longest first (counting sort), and CTA k takes items k, k + gridDim.x, ...: every
// CTA gets the same mix of expensive and cheap items.  (In clip order the static split left the slowest of 148 CTAs
// with ~1.5x the mean work at the U{10..100} length mix.)  One CTA.
__global__ void __launch_bounds__(256) attn_items_kernel(const int32_t* __restrict__ cu, int B, int2* __restrict__ items,
                                                         int32_t* __restrict__ n_items) {
  pdl_enter();
  __shared__ int s_off[130];       // s_off[l]: next free item slot for clips of length l (clamped to 128)
  for (int i = threadIdx.x; i < 130; i += blockDim.x) s_off[i] = 0;
  __syncthreads();
  auto heads_per_item = [](int L) { return L <= 32 ? 4 : L <= 64 ? 2 : 1; };
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int L = min(cu[b + 1] - cu[b], 128);
    atomicAdd(&s_off[L], kBertHeads / heads_per_item(L));
  }
  __syncthreads();
  if (threadIdx.x == 0) {          // exclusive prefix over descending length
    int acc = 0;
    for (int l = 128; l >= 0; --l) { const int c = s_off[l]; s_off[l] = acc; acc += c; }
    *n_items = acc;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int L = min(cu[b + 1] - cu[b], 128);
    const int G = heads_per_item(L), cnt = kBertHeads / G;
    const int base = atomicAdd(&s_off[L], cnt);
    for (int i = 0; i < cnt; ++i) items[base + i] = make_int2(b, (i * G) | (G << 8));
  }
}

__global__ void __launch_bounds__(kThreads, 1) bert_attention_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sSlot = smem;                                    // [slot][Q | K | V][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemTiles);
  uint64_t* full = bars;                 // [kSlots] TMA -> S issuer / softmax
  uint64_t* empty = full + kSlots;       // [kSlots] O MMAs retired -> TMA (the slot's shared memory is free)
  uint64_t* s_full = empty + kSlots;     // [kSlots] S MMAs retired -> softmax
  uint64_t* p_full = s_full + kSlots;    // [kSlots] softmax -> O issuer (P written, S consumed)
  uint64_t* o_full = p_full + kSlots;    // [kSlots] O MMAs retired -> softmax
  uint64_t* t_empty = o_full + kSlots;   // [kSlots] softmax -> S issuer (O read: the slot's TMEM columns are free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + kSlots);
  ItemInfo* info = reinterpret_cast<ItemInfo*>(bars + 32);     // [kSlots], 32 bytes each, 16-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) tma_prefetch_desc(&p.qkv_map);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&t_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  // V tiles start out as zeros: with several heads per tile the O MMA also walks V rows that no TMA box of the item
  // covered; they meet p = 0 exactly, which is only harmless while they hold finite values (later: older V rows)
  for (int i = threadIdx.x; i < kSlots * (kTileBytes / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(sSlot + (i / (kTileBytes / 16)) * kSlotBytes + 2 * kTileBytes)[i % (kTileBytes / 16)] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int n_items = ld_chain_i32(p.n_items);   // written by attn_items_kernel: not before the wait (launch.cuh)
  const int n_my = n_items > static_cast<int>(blockIdx.x)
                       ? (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp: the lanes gather the key bits)
    // Everything the producer reads from global memory is fetched ahead of use: a dependent items -> cu -> key_ok chain per
    // item (three L2 round trips) would otherwise bound the whole kernel.
    auto fetch_item = [&](int i, int& row_base, int& L, int& hg) {     // lane-parallel: lane j holds item (i0 + j)
      const int2 it = __ldg(p.items + blockIdx.x + static_cast<long>(i) * gridDim.x);
      row_base = __ldg(p.cu + it.x);
      L = min(__ldg(p.cu + it.x + 1) - row_base, 128);
      hg = it.y;
    };
    int my_rb = 0, my_L = 0, my_hg = 0;
    if (lane < n_my) fetch_item(lane, my_rb, my_L, my_hg);
    auto load_keys = [&](int rb, int L, uint8_t (&k)[4]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = q * 32 + lane;
        k[q] = j < L ? __ldg(p.key_ok + rb + j) : uint8_t(0);
      }
    };
    uint8_t keys[4] = {0, 0, 0, 0};
    int rb_cur = __shfl_sync(0xffffffffu, my_rb, 0), L_cur = __shfl_sync(0xffffffffu, my_L, 0);
    int hg_cur = __shfl_sync(0xffffffffu, my_hg, 0);
    if (n_my > 0) load_keys(rb_cur, L_cur, keys);
    for (int i = 0; i < n_my; ++i) {
      const int slot = i % kSlots;
      const int row_base = rb_cur, L = L_cur, head0 = hg_cur & 0xff, G = hg_cur >> 8;
      uint32_t bits[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) bits[q] = __ballot_sync(0xffffffffu, keys[q] != 0);
      if (i + 1 < n_my) {     // prefetch the next item (a new batch of 32 items every 32 iterations) and its key bytes
        if (((i + 1) & 31) == 0 && i + 1 + lane < n_my) fetch_item(i + 1 + lane, my_rb, my_L, my_hg);
        rb_cur = __shfl_sync(0xffffffffu, my_rb, (i + 1) & 31);
        L_cur = __shfl_sync(0xffffffffu, my_L, (i + 1) & 31);
        hg_cur = __shfl_sync(0xffffffffu, my_hg, (i + 1) & 31);
        load_keys(rb_cur, L_cur, keys);
      }
      mbar_wait(&empty[slot], ((i / kSlots) & 1) ^ 1);
      if (lane == 0) {
        ItemInfo inf;
        inf.row_base = row_base; inf.L = L;
        inf.key_bits[0] = bits[0]; inf.key_bits[1] = bits[1]; inf.key_bits[2] = bits[2]; inf.key_bits[3] = bits[3];
        inf.head0 = head0; inf.G = G;
        info[slot] = inf;
        const int nch = (L + 31) >> 5;                 // 32-row chunks per head (<= 4 / G)
        const int cpb = 4 / G;                         // chunks per head block of the tile
        mbar_expect_tx(&full[slot], static_cast<uint32_t>(3 * G * nch) * 4096u);
        uint8_t* dst = sSlot + slot * kSlotBytes;
        for (int g = 0; g < G; ++g) {
          const int col = (head0 + g) * 64;
          for (int c = 0; c < nch; ++c) {
            const int blk = g * cpb + c;
            // one head per tile: Q chunk c lands in 32-row block (c + i) % 4, because TMEM lanes [32 w, +32) can only be
            // read by warps with warp % 4 == w, which all sit on scheduler w -- without the rotation scheduler 0 would run
            // the softmax of every item.  Several heads per tile spread the rows by themselves.
            const int qblk = G == 1 ? ((c + i) & 3) : blk;
            tma_load_2d(dst + qblk * 4096, &p.qkv_map, &full[slot], col, row_base + c * 32);
            tma_load_2d(dst + kTileBytes + blk * 4096, &p.qkv_map, &full[slot], kBertHidden + col, row_base + c * 32);
            tma_load_2d(dst + 2 * kTileBytes + blk * 4096, &p.qkv_map, &full[slot], 2 * kBertHidden + col, row_base + c * 32);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ S issuer
    if (elect_one()) {
      constexpr uint32_t idesc_base = umma_idesc(1u, 128, 0);
      for (int i = 0; i < n_my; ++i) {
        const int slot = i % kSlots;
        const uint32_t ph = (i / kSlots) & 1;
        mbar_wait(&full[slot], ph);
        mbar_wait(&t_empty[slot], ph ^ 1);      // the previous item of this slot has read its O
        tc_fence_after();
        const int G = info[slot].G;
        const int Nk = (G - 1) * (128 / G) + ((info[slot].L + 15) & ~15);   // keys up to the last head's last one
        const uint32_t idesc = idesc_base | (static_cast<uint32_t>(Nk >> 3) << 17);
        const uint32_t q_addr = smem_u32(sSlot + slot * kSlotBytes);
        const uint32_t k_addr = q_addr + kTileBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + slot * 128, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), idesc, k != 0);
        umma_commit(&s_full[slot]);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ O issuer
    if (elect_one()) {
      constexpr uint32_t idesc_pv = umma_idesc(1u, 128, 64) | (1u << 16);   // B (= V) is MN-major
      for (int i = 0; i < n_my; ++i) {
        const int slot = i % kSlots;
        mbar_wait(&p_full[slot], (i / kSlots) & 1);
        tc_fence_after();
        const int G = info[slot].G;
        const int Nk = (G - 1) * (128 / G) + ((info[slot].L + 15) & ~15);
        // O = P V: A = P (K-major, 64-key K blocks, where Q | K were), B = V in place (8-key atoms 1024 B apart)
        const uint32_t p_addr = smem_u32(sSlot + slot * kSlotBytes);
        const uint32_t v_addr = p_addr + 2 * kTileBytes;
        for (int j = 0; j < Nk / 16; ++j)
          umma_bf16(tmem_base + slot * 128, umma_desc_sw128(p_addr + (j >> 2) * kTileBytes + (j & 3) * 32),
                    umma_desc_sw128(v_addr + j * 2048), idesc_pv, j != 0);
        umma_commit(&empty[slot]);     // the slot's shared memory is free once these MMAs retire
        umma_commit(&o_full[slot]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax + output, one thread per tile row
    const int group = (warp - 4) >> 2;
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32)
    const int row = quarter * 32 + lane;
    const float sl2 = 0.125f * 1.4426950408889634f;        // 1/sqrt(64) * log2(e)
    for (int i = group; i < n_my; i += kGroups) {
      const int slot = i % kSlots;
      const uint32_t ph = (i / kSlots) & 1;
      mbar_wait(&full[slot], ph);                            // item info published by the producer
      const int row_base = info[slot].row_base, L = info[slot].L, G = info[slot].G, head0 = info[slot].head0;
      const uint4 kb4 = *reinterpret_cast<const uint4*>(info[slot].key_bits);
      const uint32_t kbits[4] = {kb4.x, kb4.y, kb4.z, kb4.w};
      const int bs = 128 / G;                                // rows (= keys) per head block
      const int Nh = (L + 15) & ~15;                         // score columns of one head
      const int Nk = (G - 1) * bs + Nh;                      // columns the O MMA walks
      // tile row -> (head block g, query q).  One head per tile: Q blocks are rotated by i (see the producer)
      const int g = G == 1 ? 0 : row / bs;
      const int q = G == 1 ? ((((quarter - i) & 3) << 5) + lane) : row % bs;
      const int c_base = g * bs;                             // first score column of this row's head
      mbar_wait(&s_full[slot], ph);
      tc_fence_after();
      const bool warp_active = (q - lane) < L;               // (warp-uniform) at least one valid row
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + slot * 128;
      float l = 0.f;
      if (warp_active) {
        uint8_t* prow = sSlot + slot * kSlotBytes + row * 128;      // P overwrites Q | K (S has retired)
        // 64 score columns per TMEM round trip (four tcgen05.ld, one wait); c0 is relative to the head's first column
        auto load64 = [&](int c0, uint32_t (&r)[4][16]) {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (c0 + t * 16 < Nh) tmem_ld_32x16(taddr + c_base + c0 + t * 16, r[t]);
          tmem_ld_wait();
        };
        auto row_max = [&](int c0, const uint32_t (&r)[4][16], float m) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (c0 + t * 16 < Nh) {
              const uint32_t w = kbits[(c0 + t * 16) >> 5] >> ((c0 + t * 16) & 31);   // 16 key bits of this chunk
              float m4[4] = {m, -INFINITY, -INFINITY, -INFINITY};
              if ((w & 0xffffu) == 0xffffu) {
#pragma unroll
                for (int e = 0; e < 16; ++e) m4[e & 3] = fmaxf(m4[e & 3], __uint_as_float(r[t][e]));
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) m4[e & 3] = fmaxf(m4[e & 3], (w >> e) & 1u ? __uint_as_float(r[t][e]) : -INFINITY);
              }
              m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            }
          }
          return m;
        };
        auto store16 = [&](int c, const float (&pv)[16]) {          // c: absolute P column (multiple of 16)
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            uint4 o;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(pv[t * 8 + 2 * e], pv[t * 8 + 2 * e + 1]);
            const int chunk = ((c & 63) >> 3) + t;     // 16-byte chunk inside the 128-byte row of the K block
            *reinterpret_cast<uint4*>(prow + (c >> 6) * kTileBytes + ((chunk ^ (row & 7)) << 4)) = o;
          }
        };
        auto exp_store = [&](int c0, const uint32_t (&r)[4][16], float msub) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int c = c0 + t * 16;
            if (c < Nh) {
              const uint32_t w = kbits[c >> 5] >> (c & 31);
              float pv[16];
#pragma unroll
              for (int e = 0; e < 16; ++e)
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv[e]) : "f"(fmaf(__uint_as_float(r[t][e]), sl2, -msub)));
              if ((w & 0xffffu) != 0xffffu) {           // select, not add: stale rows beyond L may hold anything
#pragma unroll
                for (int e = 0; e < 16; ++e) pv[e] = (w >> e) & 1u ? pv[e] : 0.f;
              }
              l += ((pv[0] + pv[1]) + (pv[2] + pv[3])) + ((pv[4] + pv[5]) + (pv[6] + pv[7])) +
                   (((pv[8] + pv[9]) + (pv[10] + pv[11])) + ((pv[12] + pv[13]) + (pv[14] + pv[15])));
              store16(c_base + c, pv);
            }
          }
        };
        uint32_t r[4][16];
        load64(0, r);
        float m = row_max(0, r, -INFINITY);
        if (Nh > 64) {
          load64(64, r);
          m = row_max(64, r, m);
        }
        const float msub = (m == -INFINITY) ? 0.f : m * sl2;
        if (Nh > 64) {
          exp_store(64, r, msub);       // the upper half is still in registers
          load64(0, r);
        }
        exp_store(0, r, msub);
        if (G > 1) {                    // block-diagonal P: zeros in the other heads' key columns
          const float z[16] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int c = 0; c < Nk; c += 16)
            if (c < c_base || c >= c_base + Nh) store16(c, z);
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[slot]);
      // ---- output
      mbar_wait(&o_full[slot], ph);
      tc_fence_after();
      if (warp_active) {
        const float inv = l > 0.f ? 1.0f / l : 0.f;   // no attendable key (fully masked clip): zero context, as torch's sdpa
        __nv_bfloat16* dst = p.ctx + static_cast<long>(row_base + q) * kBertHidden + (head0 + g) * 64;
        uint32_t r[4][16];
#pragma unroll
        for (int t = 0; t < 4; ++t) tmem_ld_32x16(taddr + t * 16, r[t]);
        tmem_ld_wait();
        if (q < L) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              uint4 o;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e)
                h2[e] = __floats2bfloat162_rn(__uint_as_float(r[t][u * 8 + 2 * e]) * inv, __uint_as_float(r[t][u * 8 + 2 * e + 1]) * inv);
              *reinterpret_cast<uint4*>(dst + t * 16 + u * 8) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[slot]);
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

// Work list of the kernel above for B clips: items [12 * B] (int2), n_items [1]; depends only on cu.
void launch_attention_items(const int32_t* cu, int B, void* items, int32_t* n_items, cudaStream_t s) {
  if (B == 0) return;
  launch_pdl(attn_items_kernel, 1, 256, 0, s, cu, B, static_cast<int2*>(items), n_items);
}

// Packed layout only (cu, key_ok), bf16, every clip at most 128 tokens.  `rows` = rows of the qkv allocation that may be
// read (at least cu[B] + 31; rows beyond the packed ones must hold finite values).  items / n_items: work list from
// launch_attention_items (nullptr: built here into a cached scratch buffer).
void launch_bert_attention_tc(const void* qkv, const int32_t* cu, const uint8_t* key_ok, void* ctx, int B, long rows,
                              cudaStream_t s, const void* items, const int32_t* n_items) {
  if (B == 0) return;
  static PerDeviceOnce configured;
  if (configured.first()) {
    VCG_CUDA(cudaFuncSetAttribute(bert_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
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
  if (!items) {   // stand-alone operator: scratch work list (grown on demand, per device; not thread-safe like the engine)
    static void* scratch = nullptr;
    static int scratch_B = 0;
    if (B > scratch_B) {
      if (scratch) VCG_CUDA(cudaFree(scratch));
      VCG_CUDA(cudaMalloc(&scratch, static_cast<size_t>(B) * kBertHeads * sizeof(int2) + 16));
      scratch_B = B;
    }
    int32_t* cnt = reinterpret_cast<int32_t*>(static_cast<uint8_t*>(scratch) + static_cast<size_t>(scratch_B) * kBertHeads * sizeof(int2));
    launch_attention_items(cu, B, scratch, cnt, s);
    items = scratch;
    n_items = cnt;
  }
  AttnParams p;
  p.qkv_map = it->second;
  p.cu = cu; p.key_ok = key_ok; p.ctx = static_cast<__nv_bfloat16*>(ctx);
  p.items = static_cast<const int2*>(items); p.n_items = n_items;
  int dev = 0, sms = 0;
  VCG_CUDA(cudaGetDevice(&dev));
  VCG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = std::min(B * kBertHeads, sms);
  launch_pdl(bert_attention_tc_kernel, grid, kThreads, kSmemBytes, s, p);
}

}  // namespace vcg
