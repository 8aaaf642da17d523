#include "umma_launch.h"
namespace sdm {
SDM_DEFINE_CONV_GEMM_LAUNCH_L(128, 1, EPI_F16, false, true)
SDM_DEFINE_CONV_GEMM_LAUNCH_L(128, 1, EPI_F16_T, false, true)
SDM_DEFINE_CONV_GEMM_LAUNCH_L(128, 1, EPI_F32, false, true)
}  // namespace sdm
