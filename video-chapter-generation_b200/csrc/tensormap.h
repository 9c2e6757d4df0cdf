// Host-side helpers: CUDA error handling and TMA tensor-map encoding (driver entry point fetched at run time so
// the shared library has no link-time dependency on libcuda and loads on GPU-less machines).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace vcg {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    std::string msg = std::string(what) + " failed: " + cudaGetErrorString(e);
    if (e == cudaErrorMemoryAllocation) msg = "CUDA out of memory. " + msg;   // keeps reference OOM handlers working
    msg += " (" + std::string(file) + ":" + std::to_string(line) + ")";
    throw Error(msg);
  }
}
#define VCG_CUDA(expr) ::vcg::cuda_check((expr), #expr, __FILE__, __LINE__)
#define VCG_REQUIRE(cond, msg)                                                                      \
  do {                                                                                              \
    if (!(cond)) throw ::vcg::Error(std::string("vcg: ") + (msg) + " [" #cond "]");                 \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    VCG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    VCG_REQUIRE(q == cudaDriverEntryPointSuccess && p != nullptr, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of dims[i+1].
// 128-byte swizzle, zero fill out of bounds.
inline CUtensorMap make_tensor_map(const void* base, bool fp32, int rank, const uint64_t* dims,
                                   const uint64_t* strides_bytes, const uint32_t* box) {
  CUtensorMap m;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = get_encode_tiled()(&m, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                                  static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[512];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] strides [%llu %llu "
             "%llu %llu] box [%u %u %u %u %u] base %p",
             static_cast<int>(r), rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
             (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
             (unsigned long long)(rank > 4 ? gdim[4] : 0), (unsigned long long)(rank > 1 ? gstr[0] : 0),
             (unsigned long long)(rank > 2 ? gstr[1] : 0), (unsigned long long)(rank > 3 ? gstr[2] : 0),
             (unsigned long long)(rank > 4 ? gstr[3] : 0), bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0,
             rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0, base);
    throw Error(buf);
  }
  return m;
}

}  // namespace vcg
