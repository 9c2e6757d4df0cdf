// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st), B in shared memory (K-major SW128).
// D[128 x 64] = A[128 x 64] * B[64 x 64]^T, bf16 in, fp32 out; checks the layout assumption "row = lane, two bf16 per
// 32-bit column, K ascending with the column" against a host reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I video-chapter-generation_b200/csrc tools/probes/ts_probe.cu -o tools/probes/ts_probe
#include "ptx.cuh"
#include "tensormap.h"
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

using namespace vcg;

__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
  __shared__ __align__(1024) uint8_t sB[64 * 128];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 128); tmem_relinquish(); }
  // B [n][k] -> K-major SW128 rows of 128 B
  for (int i = tid; i < 64 * 8; i += 128) {
    const int n = i >> 3, chunk = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(B + n * 64 + chunk * 8);
    *reinterpret_cast<uint4*>(sB + n * 128 + ((chunk ^ (n & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  // A row `tid` -> TMEM lane tid, columns [64, 96): 32 columns x 2 bf16
  const uint32_t lane_addr = tbase + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < 4; ++c) {
    uint32_t r[8];
    const uint4 v0 = *reinterpret_cast<const uint4*>(A + tid * 64 + c * 16);
    const uint4 v1 = *reinterpret_cast<const uint4*>(A + tid * 64 + c * 16 + 8);
    r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
    tmem_st_32x8(lane_addr + 64 + c * 8, r);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(1u, 128, 64);
    for (int k = 0; k < 4; ++k) umma_bf16_ts(tbase, tbase + 64 + k * 8, umma_desc_sw128(smem_u32(sB) + k * 32), idesc, k != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    uint32_t r[16];
    tmem_ld_32x16(lane_addr + c * 16, r);
    tmem_ld_wait();
    for (int e = 0; e < 16; ++e) D[tid * 64 + c * 16 + e] = __uint_as_float(r[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 128);
}

int main() {
  std::vector<__nv_bfloat16> A(128 * 64), B(64 * 64);
  srand(1);
  for (auto& x : A) x = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (auto& x : B) x = __float2bfloat16((rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  probe<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> D(128 * 64);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += (double)__bfloat162float(A[m * 64 + k]) * (double)__bfloat162float(B[n * 64 + k]);
      maxerr = fmax(maxerr, fabs(ref - D[m * 64 + n]));
    }
  printf("A-in-TMEM MMA: max |D - ref| = %g  (%s)\n", maxerr, maxerr < 1e-3 ? "layout assumption holds" : "MISMATCH");
  return 0;
}
