// BERT self-attention  ctx = softmax(Q K^T / sqrt(64) + key_mask) V   (modeling_bert.py:115-140, 12 heads x 64).
//
// bf16 path: one CTA (8 warps) per (clip, head, 128-query block); K, V and the Q block are staged in shared memory by
// cp.async in two groups (Q+K, then V, so QK^T starts while V is still in flight; rows padded to 144 B so ldmatrix is
// bank-conflict free), each warp owns 16 query rows and walks the keys in blocks of 64 with an online softmax (fp32
// statistics, quad shuffles for the row reductions); QK^T and PV run on mma.sync m16n8k16.  At L <= 128 one CTA sees
// all queries of its (clip, head), so K and V are read exactly once.  Attention is 2 % of the path's FLOPs at L=100
// (10 % at L=512) and latency/L2-bound at these sizes, so it stays on the legacy tensor path.
// fp32 path (verification mode): straightforward SIMT kernel, one warp per query row.
#include "kernels.cuh"
#include "launch.cuh"
#include "tensormap.h"
#include <cuda_bf16.h>
#include <cstdlib>

namespace vcg {

namespace {

constexpr int kHeadDim = 64;
constexpr int kRowPad = 72;   // bf16 elements per padded smem row (144 B)
constexpr int kQBlock = 128;   // queries per CTA (8 warps x 16 rows)

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Packed (variable-length) mode: cu[b] .. cu[b+1] are the rows of clip b in the token-packed activation matrices and
// key_ok[row] says whether that token may be attended to; cu == nullptr: rows b*L .. b*L+L-1 and the int64 mask.
__global__ void __launch_bounds__(256, 3) bert_attention_bf16_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                  const int64_t* __restrict__ mask,
                                                                  const int32_t* __restrict__ cu,
                                                                  const uint8_t* __restrict__ key_ok,
                                                                  __nv_bfloat16* __restrict__ ctx, int Lmax, int Lp_max) {
  extern __shared__ __align__(16) uint8_t smem[];
  pdl_enter();
  const int head = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * kQBlock;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long row_base = cu ? cu[b] : static_cast<long>(b) * Lmax;
  const int L = cu ? (cu[b + 1] - cu[b]) : Lmax;
  if (q0 >= L) return;                              // (uniform) nothing to do for this query block
  const int Lp = (L + 63) / 64 * 64;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sV = sK + Lp_max * kRowPad;
  __nv_bfloat16* sQ = sV + Lp_max * kRowPad;
  float* sMask = reinterpret_cast<float*>(sQ + kQBlock * kRowPad);
  const int ld = 3 * kBertHidden;
  const int Lk = (L + 15) / 16 * 16;                // keys in steps of 16: tiles beyond Lk are skipped (p = 0 exactly)
  const int nq = min(kQBlock, L - q0);              // query rows of this block
  const int nqp = (nq + 15) / 16 * 16;

  // group 0: Q block + K; group 1: V.  Rows beyond L are zero (their probabilities are exactly 0).
  const __nv_bfloat16* src0 = qkv + row_base * ld + head * kHeadDim;
  for (int i = tid; i < nqp * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    if (r < nq) cp_async16(sQ + r * kRowPad + c, src0 + static_cast<long>(q0 + r) * ld + c);
    else *reinterpret_cast<uint4*>(sQ + r * kRowPad + c) = make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < Lp * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    if (r < L) cp_async16(sK + r * kRowPad + c, src0 + static_cast<long>(r) * ld + kBertHidden + c);
    else *reinterpret_cast<uint4*>(sK + r * kRowPad + c) = make_uint4(0, 0, 0, 0);
  }
  cp_async_commit();
  for (int i = tid; i < Lp * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    if (r < L) cp_async16(sV + r * kRowPad + c, src0 + static_cast<long>(r) * ld + 2 * kBertHidden + c);
    else *reinterpret_cast<uint4*>(sV + r * kRowPad + c) = make_uint4(0, 0, 0, 0);
  }
  cp_async_commit();
  for (int j = tid; j < Lp; j += blockDim.x) {
    bool ok = j < L;
    if (ok) ok = cu ? (key_ok[row_base + j] != 0) : (mask[row_base + j] != 0);
    sMask[j] = ok ? 0.f : -INFINITY;
  }
  cp_async_wait<1>();
  __syncthreads();          // Q, K and the mask are in shared memory

  const int qrow = warp * 16;                 // this warp's 16 query rows within the block
  const bool active = q0 + qrow < L;          // (warp-uniform) idle warps only take part in the barrier below

  const float sl2 = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  for (int kb = 0; kb < Lp; kb += 64) {
    uint32_t pf[4][4];   // P as A fragments: 4 k-steps of 16 keys
    float corr[2] = {1.f, 1.f};
    if (active) {
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t qf[4];    // Q fragment of this k-step (re-read per key block: cheaper than 16 live registers)
        ldmatrix_x4(qf, sQ + (qrow + (lane & 7) + ((lane >> 3) & 1) * 8) * kRowPad + ks * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int np = 0; np < 4; ++np) {   // pairs of 8-key tiles
          if (kb + np * 16 >= Lk) break;   // (uniform) nothing but padding from here on
          uint32_t kf[4];
          ldmatrix_x4(kf, sK + (kb + np * 16 + (lane & 7) + (lane >> 4) * 8) * kRowPad + ks * 16 + ((lane >> 3) & 1) * 8);
          mma_bf16_16816(s[2 * np], qf, kf[0], kf[1]);
          mma_bf16_16816(s[2 * np + 1], qf, kf[2], kf[3]);
        }
      }
      // scale, mask, block row-max
      float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = kb + nt * 8 + (lane & 3) * 2;
        const float mk0 = sMask[key], mk1 = sMask[key + 1];
        s[nt][0] = s[nt][0] * sl2 + mk0; s[nt][1] = s[nt][1] * sl2 + mk1;
        s[nt][2] = s[nt][2] * sl2 + mk0; s[nt][3] = s[nt][3] * sl2 + mk1;
        bm[0] = fmaxf(bm[0], fmaxf(s[nt][0], s[nt][1]));
        bm[1] = fmaxf(bm[1], fmaxf(s[nt][2], s[nt][3]));
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        bm[r] = fmaxf(bm[r], __shfl_xor_sync(0xffffffffu, bm[r], 1));
        bm[r] = fmaxf(bm[r], __shfl_xor_sync(0xffffffffu, bm[r], 2));
      }
      float mnew[2], msub[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mnew[r] = fmaxf(m[r], bm[r]);
        msub[r] = (mnew[r] == -INFINITY) ? 0.f : mnew[r];
        corr[r] = exp2f(m[r] - msub[r]);   // m = -inf -> 0
        m[r] = mnew[r];
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = exp2f(s[nt][0] - msub[0]), p1 = exp2f(s[nt][1] - msub[0]);
        const float p2 = exp2f(s[nt][2] - msub[1]), p3 = exp2f(s[nt][3] - msub[1]);
        rs[0] += p0 + p1; rs[1] += p2 + p3;
        pf[nt >> 1][(nt & 1) * 2] = pack2(p0, p1);
        pf[nt >> 1][(nt & 1) * 2 + 1] = pack2(p2, p3);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
        rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
        l[r] = l[r] * corr[r] + rs[r];
      }
    }
    if (kb == 0) {            // (CTA-uniform) V has landed
      cp_async_wait<0>();
      __syncthreads();
    }
    if (active) {
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        o[dt][0] *= corr[0]; o[dt][1] *= corr[0]; o[dt][2] *= corr[1]; o[dt][3] *= corr[1];
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {       // 16 keys per step
        if (kb + ks * 16 >= Lk) break;       // (uniform) padding keys: p = 0
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {     // pairs of 8-wide d tiles
          uint32_t vf[4];
          ldmatrix_x4_trans(vf, sV + (kb + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kRowPad + dp * 16 + (lane >> 4) * 8);
          mma_bf16_16816(o[2 * dp], pf[ks], vf[0], vf[1]);
          mma_bf16_16816(o[2 * dp + 1], pf[ks], vf[2], vf[3]);
        }
      }
    }
  }
  if (!active) return;

  // normalise and write: row g -> regs 0,1 ; row g+8 -> regs 2,3
  const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int q = q0 + qrow + g + r * 8;
    if (q < L) {
      const float inv = l[r] > 0.f ? 1.0f / l[r] : 0.f;   // fully masked row: zero context (torch sdpa's safe softmax)
      __nv_bfloat16* dst = ctx + (row_base + q) * kBertHidden + head * kHeadDim + t2;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt)
        *reinterpret_cast<uint32_t*>(dst + dt * 8) = pack2(o[dt][r * 2] * inv, o[dt][r * 2 + 1] * inv);
    }
  }
}

// fp32 verification path: one warp per query row, keys/values straight from L2.
__global__ void __launch_bounds__(128) bert_attention_fp32_kernel(const float* __restrict__ qkv,
                                                                  const int64_t* __restrict__ mask,
                                                                  const int32_t* __restrict__ cu,
                                                                  const uint8_t* __restrict__ key_ok,
                                                                  float* __restrict__ ctx, int Lmax) {
  extern __shared__ float fsm[];
  pdl_enter();
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sq = fsm + warp * (kHeadDim + Lmax);
  float* sp = sq + kHeadDim;
  const long row_base = cu ? cu[b] : static_cast<long>(b) * Lmax;
  const int L = cu ? (cu[b + 1] - cu[b]) : Lmax;
  const int ld = 3 * kBertHidden;
  for (int q = blockIdx.z * 4 + warp; q < L; q += gridDim.z * 4) {
    const float* qp = qkv + (row_base + q) * ld + head * kHeadDim;
    sq[lane] = qp[lane];
    sq[lane + 32] = qp[lane + 32];
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) {
      const float* kp = qkv + (row_base + j) * ld + kBertHidden + head * kHeadDim;
      float acc = 0.f;
#pragma unroll 16
      for (int d = 0; d < kHeadDim; ++d) acc = fmaf(sq[d], kp[d], acc);
      const bool ok = cu ? (key_ok[row_base + j] != 0) : (mask[row_base + j] != 0);
      acc = acc * 0.125f + (ok ? 0.f : -INFINITY);
      sp[j] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (mx == -INFINITY) mx = 0.f;
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float p = expf(sp[j] - mx);
      sp[j] = p;
      sum += p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < L; ++j) {
      const float* vp = qkv + (row_base + j) * ld + 2 * kBertHidden + head * kHeadDim;
      const float p = sp[j];
      a0 = fmaf(p, vp[lane], a0);
      a1 = fmaf(p, vp[lane + 32], a1);
    }
    float* dst = ctx + (row_base + q) * kBertHidden + head * kHeadDim;
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;   // fully masked row: zero context (torch sdpa's safe softmax)
    dst[lane] = a0 * inv;
    dst[lane + 32] = a1 * inv;
    __syncwarp();
  }
}

}  // namespace

void launch_bert_attention(const void* qkv, const int64_t* mask, const int32_t* cu, const uint8_t* key_ok, void* ctx, int B,
                           int L, bool fp32, cudaStream_t s, long qkv_rows, const void* items, const int32_t* n_items) {
  if (B == 0) return;
  VCG_REQUIRE(L >= 1 && L <= 512, "BERT sequence length must be in [1, 512]");
  static int tc_policy = -1;   // VCG_ATTN_TC=0 keeps the mma.sync kernel everywhere
  if (tc_policy < 0) {
    const char* v = getenv("VCG_ATTN_TC");
    tc_policy = (v && atoi(v) == 0) ? 0 : 1;
  }
  if (!fp32 && cu && key_ok && L <= 128 && qkv_rows > 0 && tc_policy) {   // packed bf16, one score tile per (clip, head)
    launch_bert_attention_tc(qkv, cu, key_ok, ctx, B, qkv_rows, s, items, n_items);
    return;
  }
  if (!fp32) {
    const int Lp = (L + 63) / 64 * 64;
    const size_t smem = static_cast<size_t>(2 * Lp + kQBlock) * kRowPad * sizeof(__nv_bfloat16) + Lp * sizeof(float);
    static PerDeviceMax configured;
    if (configured.raise(smem)) {
      VCG_CUDA(cudaFuncSetAttribute(bert_attention_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    }
    dim3 grid(kBertHeads, B, (L + kQBlock - 1) / kQBlock);
    launch_pdl(bert_attention_bf16_kernel, grid, 256, smem, s, static_cast<const __nv_bfloat16*>(qkv), mask, cu, key_ok,
               static_cast<__nv_bfloat16*>(ctx), L, Lp);
  } else {
    const size_t smem = static_cast<size_t>(4) * (kHeadDim + L) * sizeof(float);
    dim3 grid(kBertHeads, B, (L + 31) / 32);
    launch_pdl(bert_attention_fp32_kernel, grid, 128, smem, s, static_cast<const float*>(qkv), mask, cu, key_ok,
               static_cast<float*>(ctx), L);
  }
  VCG_CUDA(cudaGetLastError());
}

}  // namespace vcg
