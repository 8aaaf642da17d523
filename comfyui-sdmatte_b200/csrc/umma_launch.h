// Launch shims of the conv_gemm_kernel instantiations (spread over several translation units to build in parallel).
#pragma once
#include "umma_gemm.cuh"

namespace sdm {
template <int BN, int MT, int MODE, bool UPS2, int EWG = 1>
void conv_gemm_launch(const ConvGemmParams& p, int grid, cudaStream_t st);

#define SDM_DEFINE_CONV_GEMM_LAUNCH(BN, MT, MODE, UPS2) SDM_DEFINE_CONV_GEMM_LAUNCH_E(BN, MT, MODE, UPS2, 1)
#define SDM_DEFINE_CONV_GEMM_LAUNCH_E(BN, MT, MODE, UPS2, EWG)                                                               \
  template <>                                                                                                                \
  void conv_gemm_launch<BN, MT, MODE, UPS2, EWG>(const ConvGemmParams& p, int grid, cudaStream_t st) {                       \
    using Cfg = ConvGemmCfg<BN, MT, EWG>;                                                                                    \
    auto kern = conv_gemm_kernel<BN, MT, MODE, UPS2, EWG>;                                                                   \
    static PerDeviceOnce attr;                                                                                               \
    attr([&] { SDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes)); });    \
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(p);                                                                   \
    SDM_CUDA_OK(cudaGetLastError());                                                                                         \
  }

// resident-halo variant for the 3x3 stride-1 convs (EPI_F16)
template <int BN, int MT, bool UPS2, int EWG>
void conv_gemm_launch_halo(const ConvGemmParams& p, int grid, cudaStream_t st);
#define SDM_DEFINE_CONV_GEMM_LAUNCH_HALO(BN, MT, UPS2, EWG)                                                                  \
  template <>                                                                                                                \
  void conv_gemm_launch_halo<BN, MT, UPS2, EWG>(const ConvGemmParams& p, int grid, cudaStream_t st) {                         \
    using Cfg = ConvGemmCfg<BN, MT, EWG, true>;                                                                              \
    auto kern = conv_gemm_kernel<BN, MT, EPI_F16, UPS2, EWG, true>;                                                          \
    static PerDeviceOnce attr;                                                                                               \
    attr([&] { SDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes)); });    \
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(p);                                                                   \
    SDM_CUDA_OK(cudaGetLastError());                                                                                         \
  }
}  // namespace sdm
