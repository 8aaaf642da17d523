#include "umma_launch.h"
namespace sdm {
SDM_DEFINE_CONV_GEMM_LAUNCH_HALO(256, 1, false, 1)
SDM_DEFINE_CONV_GEMM_LAUNCH_HALO(256, 1, true, 1)
}  // namespace sdm
