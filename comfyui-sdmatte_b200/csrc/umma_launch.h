// Launch shims of the conv_gemm_kernel instantiations (spread over several translation units to build in parallel).
#pragma once
#include "umma_gemm.cuh"

namespace sdm {
template <int BN, int MT, int MODE, bool UPS2, bool LIGHT = false>
void conv_gemm_launch(const ConvGemmParams& p, int grid, cudaStream_t st);

#define SDM_DEFINE_CONV_GEMM_LAUNCH(BN, MT, MODE, UPS2) SDM_DEFINE_CONV_GEMM_LAUNCH_L(BN, MT, MODE, UPS2, false)
#define SDM_DEFINE_CONV_GEMM_LAUNCH_L(BN, MT, MODE, UPS2, LIGHT)                                                                  \
  template <>                                                                                                                \
  void conv_gemm_launch<BN, MT, MODE, UPS2, LIGHT>(const ConvGemmParams& p, int grid, cudaStream_t st) {                            \
    using Cfg = ConvGemmCfg<BN, MT, LIGHT>;                                                                                       \
    static bool attr = false;                                                                                                \
    if (!attr) {                                                                                                             \
      SDM_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, MT, MODE, UPS2, LIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                       Cfg::kSmemBytes));                                                                    \
      attr = true;                                                                                                           \
    }                                                                                                                        \
    conv_gemm_kernel<BN, MT, MODE, UPS2, LIGHT><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(p);                                   \
    SDM_CUDA_OK(cudaGetLastError());                                                                                         \
  }
}  // namespace sdm
