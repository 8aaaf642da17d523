// Launch shims of the conv_gemm_kernel instantiations (spread over several translation units to build in parallel).
#pragma once
#include "umma_gemm.cuh"

namespace sdm {
template <int BN, int MT, int MODE, bool UPS2, bool LIGHT = false, int EWG = 1>
void conv_gemm_launch(const ConvGemmParams& p, int grid, cudaStream_t st);

#define SDM_DEFINE_CONV_GEMM_LAUNCH(BN, MT, MODE, UPS2) SDM_DEFINE_CONV_GEMM_LAUNCH_E(BN, MT, MODE, UPS2, false, 1)
#define SDM_DEFINE_CONV_GEMM_LAUNCH_L(BN, MT, MODE, UPS2, LIGHT) SDM_DEFINE_CONV_GEMM_LAUNCH_E(BN, MT, MODE, UPS2, LIGHT, 1)
#define SDM_DEFINE_CONV_GEMM_LAUNCH_E(BN, MT, MODE, UPS2, LIGHT, EWG)                                                                  \
  template <>                                                                                                                \
  void conv_gemm_launch<BN, MT, MODE, UPS2, LIGHT, EWG>(const ConvGemmParams& p, int grid, cudaStream_t st) {                            \
    using Cfg = ConvGemmCfg<BN, MT, LIGHT, EWG>;                                                                                       \
    static bool attr = false;                                                                                                \
    if (!attr) {                                                                                                             \
      SDM_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, MT, MODE, UPS2, LIGHT, EWG>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                       Cfg::kSmemBytes));                                                                    \
      attr = true;                                                                                                           \
    }                                                                                                                        \
    conv_gemm_kernel<BN, MT, MODE, UPS2, LIGHT, EWG><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(p);                                   \
    SDM_CUDA_OK(cudaGetLastError());                                                                                         \
  }
}  // namespace sdm
