// Instantiations and launcher of the tcgen05 implicit-GEMM kernel.
#include "conv_gemm_host.h"
#include <cstdlib>

namespace vcg {

int sm_count() {   // of the CURRENT device (plans are built per engine, i.e. per device)
  static int n[64] = {};
  int dev = 0;
  VCG_CUDA(cudaGetDevice(&dev));
  const int slot = (dev >= 0 && dev < 64) ? dev : 0;
  if (!n[slot]) VCG_CUDA(cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, dev));
  return n[slot];
}

int cg2_policy() {
  static int pol = -2;
  if (pol == -2) {
    const char* v = getenv("VCG_CG2");
    pol = v ? atoi(v) : -1;
  }
  return pol;
}

int fuse23_policy() {
  static int pol = -2;
  if (pol == -2) {
    const char* v = getenv("VCG_FUSE23");
    pol = v ? atoi(v) : -1;
  }
  return pol;
}

int c23h_policy() {
  static int pol = -2;
  if (pol == -2) {
    const char* v = getenv("VCG_C23H");
    pol = v ? atoi(v) : -1;
  }
  return pol;
}

int c23h2_policy() {
  static int pol = -2;
  if (pol == -2) {
    const char* v = getenv("VCG_C23H2");
    pol = v ? atoi(v) : -1;
  }
  return pol;
}

void launch_conv23(const Conv23Launch& L, cudaStream_t stream) {
  if (L.grid <= 0) return;
  if (L.halo == 2) {
    static PerDeviceOnce configured_h2;
    if (configured_h2.first())
      VCG_CUDA(cudaFuncSetAttribute(conv23h2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC23h2SmemBytes));
    launch_pdl(conv23h2_kernel<0>, L.grid, kC23Threads, kC23h2SmemBytes, stream, L.q);
    return;
  }
  if (L.halo) {
    static PerDeviceOnce configured_h;
    if (configured_h.first())
      VCG_CUDA(cudaFuncSetAttribute(conv23h_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC23hSmemBytes));
    launch_pdl(conv23h_kernel<0>, L.grid, kC23Threads, kC23hSmemBytes, stream, L.q);
    return;
  }
  // P = 128 without a TSM scatter: conv3 operand in tensor memory, five ring stages (conv23t.cuh); VCG_C23T=0 -> conv23_kernel
  static const bool c23t_on = !(getenv("VCG_C23T") && atoi(getenv("VCG_C23T")) == 0);
  if (c23t_on && L.q.P == 128 && L.q.g.tsm_out == nullptr && L.q.g.act == ACT_RELU && L.q.g.n_taps * L.q.g.cpt == 18) {
    static PerDeviceOnce configured_t;
    if (configured_t.first())
      VCG_CUDA(cudaFuncSetAttribute(conv23t_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC23tSmemBytes));
    launch_pdl(conv23t_kernel<0>, L.grid, kC23Threads, kC23tSmemBytes, stream, L.q);
    return;
  }
  static PerDeviceOnce configured;
  if (configured.first()) {
    VCG_CUDA(cudaFuncSetAttribute(conv23_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC23SmemBytes));
    VCG_CUDA(cudaFuncSetAttribute(conv23_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC23SmemBytes));
  }
  if (L.q.P == 64) launch_pdl(conv23_kernel<64>, L.grid, kC23Threads, kC23SmemBytes, stream, L.q);
  else launch_pdl(conv23_kernel<128>, L.grid, kC23Threads, kC23SmemBytes, stream, L.q);
}

template <int BLOCK_N, bool LNF = false>
static void launch_pair_t(const ConvGemmLaunch& L, cudaStream_t stream) {
  using Cfg = ConvGemmCfg<BLOCK_N, false>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    VCG_CUDA(cudaFuncSetAttribute(conv_gemm_pair_kernel<BLOCK_N, LNF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
  }
  launch_pdl(conv_gemm_pair_kernel<BLOCK_N, LNF>, L.grid, Cfg::kThreads, Cfg::kSmemBytes, stream, L.p);   // __cluster_dims__(2,1,1)
}

template <int BLOCK_N, bool TF32X3, bool LNF = false>
static void launch_t(const ConvGemmLaunch& L, cudaStream_t stream) {
  using Cfg = ConvGemmCfg<BLOCK_N, TF32X3>;
  static PerDeviceOnce configured;
  if (configured.first()) {
    VCG_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, TF32X3, LNF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
  }
  launch_pdl(conv_gemm_kernel<BLOCK_N, TF32X3, LNF>, L.grid, Cfg::kThreads, Cfg::kSmemBytes, stream, L.p);
}

void launch_conv_gemm(const ConvGemmLaunch& L, cudaStream_t stream) {
  if (L.grid <= 0) return;
  if (L.lnf) {   // BERT GEMMs with the LayerNorm terms: N = 768 (192-wide tiles), 2304 / 3072 (256-wide)
    VCG_REQUIRE(!L.fp32 && (L.block_n == 192 || L.block_n == 256), "LayerNorm-fused epilogue: unsupported tile");
    if (L.cg2) {
      if (L.block_n == 192) launch_pair_t<192, true>(L, stream); else launch_pair_t<256, true>(L, stream);
    } else {
      if (L.block_n == 192) launch_t<192, false, true>(L, stream); else launch_t<256, false, true>(L, stream);
    }
  } else if (L.cg2) {
    VCG_REQUIRE(!L.fp32 && L.grid % 2 == 0, "CTA-pair launch needs bf16 and an even grid");
    switch (L.block_n) {
      case 64: launch_pair_t<64>(L, stream); break;
      case 128: launch_pair_t<128>(L, stream); break;
      case 192: launch_pair_t<192>(L, stream); break;
      case 256: launch_pair_t<256>(L, stream); break;
      default: throw Error("vcg: unsupported BLOCK_N");
    }
  } else if (!L.fp32) {
    switch (L.block_n) {
      case 64: launch_t<64, false>(L, stream); break;
      case 128: launch_t<128, false>(L, stream); break;
      case 192: launch_t<192, false>(L, stream); break;
      case 256: launch_t<256, false>(L, stream); break;
      default: throw Error("vcg: unsupported BLOCK_N");
    }
  } else {
    switch (L.block_n) {
      case 64: launch_t<64, true>(L, stream); break;
      case 128: launch_t<128, true>(L, stream); break;
      default: throw Error("vcg: unsupported BLOCK_N (fp32 mode)");
    }
  }
}

}  // namespace vcg
