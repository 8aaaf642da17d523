#include "umma_launch.h"
namespace sdm {
SDM_DEFINE_CONV_GEMM_LAUNCH_HALO(160, 1, false, 2)
SDM_DEFINE_CONV_GEMM_LAUNCH_HALO(160, 1, true, 2)
}  // namespace sdm
