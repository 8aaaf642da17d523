#include "umma_launch.h"
namespace sdm {
SDM_DEFINE_CONV_GEMM_LAUNCH_E(128, 2, EPI_F16, true, 2)
SDM_DEFINE_CONV_GEMM_LAUNCH_E(64, 1, EPI_F16, true, 2)
SDM_DEFINE_CONV_GEMM_LAUNCH_E(256, 1, EPI_F32, false, 2)
}  // namespace sdm
