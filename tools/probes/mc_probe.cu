// Probe: is the ~10 TB/s L2 -> SM operand ceiling on the L2 / crossbar side or at the SM's ingress?
// Every CTA of a (2,1,1) cluster streams 32 KB tiles (256 rows x 128 B, SWIZZLE_128B) from an L2-resident matrix into a
// 4-stage shared-memory ring and throws them away.
//   mode 0: each CTA loads its own 32 KB per stage                         (per-SM ingress 32 KB, L2 reads 32 KB per CTA)
//   mode 1: each CTA loads HALF a tile and multicasts it to both CTAs      (per-SM ingress 32 KB, L2 reads 16 KB per CTA)
//   mode 2: each CTA loads half a tile for itself only (control)           (per-SM ingress 16 KB, L2 reads 16 KB per CTA)
// Delivered bytes per second per SM tell which side limits.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
//   -I video-chapter-generation_b200/csrc tools/probes/mc_probe.cu -o tools/probes/mc_probe
#include "ptx.cuh"
#include "tensormap.h"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace vcg;

constexpr int kStages = 4, kTileRows = 256, kTileBytes = kTileRows * 128;

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1)
probe_kernel(const __grid_constant__ CUtensorMap full_map, const __grid_constant__ CUtensorMap half_map, int mode, int iters,
             int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kTileBytes);
  uint64_t* empty = full + kStages;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], mode == 1 ? 2 : 1); }
    fence_mbar_init();
  }
  cluster_sync();
  const int pair = blockIdx.x >> 1;
  if (threadIdx.x == 0) {          // producer
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[s], ph ^ 1);
      const int tile = (mode == 3 ? (static_cast<int>(blockIdx.x) * 13 + i * 5) : (pair * 7 + i)) % n_tiles;
      if (mode == 0 || mode == 3) {
        mbar_expect_tx(&full[s], kTileBytes);
        tma_load_2d(smem + s * kTileBytes, &full_map, &full[s], 0, tile * kTileRows);
      } else if (mode == 1) {
        mbar_expect_tx(&full[s], kTileBytes);          // own half + the peer's half
        tma_load_2d_mc(smem + s * kTileBytes + rank * (kTileBytes / 2), &half_map, &full[s], 0,
                       tile * kTileRows + rank * (kTileRows / 2), static_cast<uint16_t>(3));
      } else {
        mbar_expect_tx(&full[s], kTileBytes / 2);
        tma_load_2d(smem + s * kTileBytes, &half_map, &full[s], 0, tile * kTileRows + rank * (kTileRows / 2));
      }
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {   // consumer: release the stage (in both CTAs when the peer writes into it)
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[s], ph);
      if (mode == 1) { mbar_arrive_cluster(&empty[s], 0); mbar_arrive_cluster(&empty[s], 1); }
      else mbar_arrive(&empty[s]);
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
  }
  cluster_sync();
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  const int n_tiles = argc > 2 ? atoi(argv[2]) : 512;        // 512 tiles = 16 MB working set: L2 resident
  void* buf = nullptr;
  VCG_CUDA(cudaMalloc(&buf, static_cast<size_t>(n_tiles) * kTileBytes));
  VCG_CUDA(cudaMemset(buf, 1, static_cast<size_t>(n_tiles) * kTileBytes));
  const uint64_t dims[2] = {64, static_cast<uint64_t>(n_tiles) * kTileRows};
  const uint64_t str[1] = {128};
  const uint32_t box_full[2] = {64, kTileRows}, box_half[2] = {64, kTileRows / 2};
  CUtensorMap full_map = make_tensor_map(buf, false, 2, dims, str, box_full);
  CUtensorMap half_map = make_tensor_map(buf, false, 2, dims, str, box_half);
  const size_t smem = kStages * kTileBytes + 1024 + 256;
  VCG_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int sms = 0;
  VCG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = sms / 2 * 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[4] = {"own full tile (pair shares the tile)", "half tile multicast to the pair", "own half tile (control)", "own full tile, distinct tile per CTA"};
  for (int mode = 0; mode < 4; ++mode) {
    probe_kernel<<<grid, 64, smem>>>(full_map, half_map, mode, 200, n_tiles);
    VCG_CUDA(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    probe_kernel<<<grid, 64, smem>>>(full_map, half_map, mode, iters, n_tiles);
    cudaEventRecord(e1);
    VCG_CUDA(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double delivered = static_cast<double>(grid) * iters * (mode == 2 ? kTileBytes / 2 : kTileBytes);
    const double l2_read = static_cast<double>(grid) * iters * ((mode == 0 || mode == 3) ? kTileBytes : kTileBytes / 2);
    printf("mode %d (%s): %.3f ms, delivered to shared memory %.2f TB/s (%.1f GB/s per SM), read from L2 %.2f TB/s\n", mode,
           names[mode], ms, delivered / ms / 1e9, delivered / ms / 1e6 / grid, l2_read / ms / 1e9);
  }
  return 0;
}
