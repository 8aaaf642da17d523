#include "umma_launch.h"
namespace sdm {
SDM_DEFINE_CONV_GEMM_LAUNCH_PAIR(128, EPI_F16)
}  // namespace sdm
