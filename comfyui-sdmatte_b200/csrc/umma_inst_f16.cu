#include "umma_launch.h"
namespace sdm {
SDM_DEFINE_CONV_GEMM_LAUNCH(256, 1, EPI_F16, false)
SDM_DEFINE_CONV_GEMM_LAUNCH(160, 1, EPI_F16, false)
SDM_DEFINE_CONV_GEMM_LAUNCH(128, 1, EPI_F16, false)
}  // namespace sdm
