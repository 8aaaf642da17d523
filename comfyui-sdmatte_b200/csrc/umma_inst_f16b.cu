#include "umma_launch.h"
namespace sdm {
SDM_DEFINE_CONV_GEMM_LAUNCH(128, 2, EPI_F16, false)
SDM_DEFINE_CONV_GEMM_LAUNCH(64, 1, EPI_F16, false)
SDM_DEFINE_CONV_GEMM_LAUNCH(16, 1, EPI_F16, false)
}  // namespace sdm
