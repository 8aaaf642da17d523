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
    static PerDeviceOnce attr;                                                                                               \
    attr([] {                                                                                                                \
      SDM_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BN, MT, MODE, UPS2, LIGHT, EWG>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                       Cfg::kSmemBytes));                                                                    \
    });                                                                                                                      \
    conv_gemm_kernel<BN, MT, MODE, UPS2, LIGHT, EWG><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(p);                                   \
    SDM_CUDA_OK(cudaGetLastError());                                                                                         \
  }

// resident-halo variant for the 3x3 stride-1 convs (EPI_F16)
template <int BN, int MT, bool UPS2, int EWG>
void conv_gemm_launch_halo(const ConvGemmParams& p, int grid, cudaStream_t st);
#define SDM_DEFINE_CONV_GEMM_LAUNCH_HALO(BN, MT, UPS2, EWG)                                                                  \
  template <>                                                                                                                \
  void conv_gemm_launch_halo<BN, MT, UPS2, EWG>(const ConvGemmParams& p, int grid, cudaStream_t st) {                         \
    using Cfg = ConvGemmCfg<BN, MT, false, EWG, false, true>;                                                                \
    auto kern = conv_gemm_kernel<BN, MT, EPI_F16, UPS2, false, EWG, false, true>;                                            \
    static PerDeviceOnce attr;                                                                                               \
    attr([&] { SDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes)); });    \
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(p);                                                                   \
    SDM_CUDA_OK(cudaGetLastError());                                                                                         \
  }

// CTA-pair (cta_group::2) variant: cluster of 2 CTAs, grid = an even number of CTAs
template <int BN, int MODE>
void conv_gemm_launch_pair(const ConvGemmParams& p, int grid, cudaStream_t st);
#define SDM_DEFINE_CONV_GEMM_LAUNCH_PAIR(BN, MODE)                                                                           \
  template <>                                                                                                                \
  void conv_gemm_launch_pair<BN, MODE>(const ConvGemmParams& p, int grid, cudaStream_t st) {                                 \
    using Cfg = ConvGemmCfg<BN, 1, false, 1, true>;                                                                          \
    auto kern = conv_gemm_kernel<BN, 1, MODE, false, false, 1, true>;                                                        \
    static PerDeviceOnce attr;                                                                                               \
    attr([&] { SDM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes)); });    \
    cudaLaunchConfig_t cfg = {};                                                                                             \
    cfg.gridDim = dim3((unsigned)grid);                                                                                      \
    cfg.blockDim = dim3(Cfg::kThreads);                                                                                      \
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;                                                                                  \
    cfg.stream = st;                                                                                                         \
    cudaLaunchAttribute at[1];                                                                                               \
    at[0].id = cudaLaunchAttributeClusterDimension;                                                                          \
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;                                      \
    cfg.attrs = at;                                                                                                          \
    cfg.numAttrs = 1;                                                                                                        \
    SDM_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));                                                                          \
  }
}  // namespace sdm
