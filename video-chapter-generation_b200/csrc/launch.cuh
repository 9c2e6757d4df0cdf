// Programmatic dependent launch (PDL) for the forward path.
//
// A scoring pass is a chain of ~100 short kernels on one stream.  Launched the ordinary way, kernel i+1 is only
// scheduled after kernel i has drained, so every boundary costs launch latency plus the prologue of the next kernel
// (barrier init, TMEM allocation, tensor-map prefetch).  With PDL every forward kernel
//   * calls griddepcontrol.launch_dependents first thing, so its successor's CTAs become resident as soon as SMs free up,
//   * and executes griddepcontrol.wait before its first access to global memory that a predecessor may still be
//     reading or writing (the wait returns when ALL earlier grids of the chain have completed and flushed).
// Both instructions are no-ops for a kernel launched without the attribute.  VCG_PDL=0 disables the attribute.
#pragma once
#include "tensormap.h"
#include <cstdlib>

namespace vcg {

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// for kernels without a prologue worth overlapping
__device__ __forceinline__ void pdl_enter() {
  pdl_launch_dependents();
  pdl_wait();
}

// A value that an EARLIER KERNEL OF THE CHAIN wrote (device-side row / item counts) must be loaded after griddepcontrol.wait.
// __ldg and loads through const __restrict__ pointers are invariant loads to the compiler, which is free to hoist them
// above the wait's asm statement: the tcgen05 attention kernel read its item count before the wait (LDG.E.CONSTANT ahead of
// ACQBULK in the SASS), i.e. possibly before attn_items_kernel had written it, and used the count of the PREVIOUS pass (zero
// on a fresh engine) whenever its CTAs became resident early enough.  A volatile asm load keeps its place behind the wait.
__device__ __forceinline__ int ld_chain_i32(const int* p) {
  int v;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Function attributes (dynamic shared memory opt-in) and the SM count are PER DEVICE: a process may hold engines on several
// GPUs (Engine(device=...)), so one-time setup is keyed by the current device, not by the process.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    VCG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || done[dev]) return dev < 0 || dev >= 64;
    done[dev] = true;
    return true;
  }
};

// the same for a dynamic-shared-memory limit that grows with the problem: true when `bytes` exceeds what was set so far
struct PerDeviceMax {
  size_t cur[64] = {};
  bool raise(size_t bytes) {
    int dev = 0;
    VCG_CUDA(cudaGetDevice(&dev));
    const int slot = (dev >= 0 && dev < 64) ? dev : 0;
    if (bytes <= cur[slot]) return false;
    cur[slot] = bytes;
    return true;
  }
};

inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* v = getenv("VCG_PDL");
    on = (v && atoi(v) == 0) ? 0 : 1;
  }
  return on != 0;
}

// debug (vcg_debug_pdl_window): launches number [from, to) since the window was set go out WITHOUT the attribute, to bisect
// an ordering problem to one kernel boundary
struct PdlDebugWindow { int idx = 0, from = 0, to = 0; };
PdlDebugWindow& pdl_debug_window();   // defined in engine.cu

template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  PdlDebugWindow& dw = pdl_debug_window();
  const bool windowed_off = dw.idx >= dw.from && dw.idx < dw.to;
  ++dw.idx;
  attr[0].val.programmaticStreamSerializationAllowed = (pdl_enabled() && !windowed_off) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VCG_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

}  // namespace vcg
