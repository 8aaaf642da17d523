// Launch shims of the conv_gemm_kernel instantiations (spread over several translation units to build in parallel).
#pragma once
#include "umma_gemm.cuh"

namespace sdm {
template <int BN, int MT, int MODE, bool UPS2>
void conv_gemm_launch(const ConvGemmParams& p, int grid, cudaStream_t st);

#define SDM_DEFINE_CONV_GEMM_LAUNCH(BN, MT, MODE, UPS2)                                                                      \
  template <>                                                                                                                \
  void conv_gemm_launch<BN, MT, MODE, UPS2>(const ConvGemmParams& p, int grid, cudaStream_t st) {                            \
    using Cfg = ConvGemmCfg<BN, MT>;                                                                                         \
    static bool attr = false;                                                                                                \
    if (!attr) {                                                                                                             \
      SDM_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, MT, MODE, UPS2>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                       Cfg::kSmemBytes));                                                                    \
      attr = true;                                                                                                           \
    }                                                                                                                        \
    conv_gemm_kernel<BN, MT, MODE, UPS2><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(p);                                   \
    SDM_CUDA_OK(cudaGetLastError());                                                                                         \
  }
}  // namespace sdm
