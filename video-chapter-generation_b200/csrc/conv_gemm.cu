// Instantiations and launcher of the tcgen05 implicit-GEMM kernel.
#include "conv_gemm_host.h"

namespace vcg {

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    VCG_CUDA(cudaGetDevice(&dev));
    VCG_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

template <int BLOCK_N, bool TF32X3>
static void launch_t(const ConvGemmLaunch& L, cudaStream_t stream) {
  using Cfg = ConvGemmCfg<BLOCK_N, TF32X3>;
  static bool configured = false;
  if (!configured) {
    VCG_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, TF32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    configured = true;
  }
  conv_gemm_kernel<BLOCK_N, TF32X3><<<L.grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(L.p);
  VCG_CUDA(cudaGetLastError());
}

void launch_conv_gemm(const ConvGemmLaunch& L, cudaStream_t stream) {
  if (L.grid <= 0) return;
  if (!L.fp32) {
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
