#include "umma_launch.h"
namespace sdm {
SDM_DEFINE_CONV_GEMM_LAUNCH_E(256, 1, EPI_F16, true, 2)
SDM_DEFINE_CONV_GEMM_LAUNCH_E(160, 1, EPI_F16, true, 2)
SDM_DEFINE_CONV_GEMM_LAUNCH_E(128, 1, EPI_F16, true, 2)
}  // namespace sdm
