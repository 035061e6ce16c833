// cuTensorMapEncodeTiled through the runtime's driver entry point: the library links against
// cudart only (no libcuda dependency at load time: the CPU-only test box loads it too).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace hrf {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
static inline PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<PFN_tmapEncodeTiled>(f);
  }();
  return fn;
}

}  // namespace hrf
