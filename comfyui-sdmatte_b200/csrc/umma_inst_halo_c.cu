#include "umma_launch.h"
namespace sdm {
SDM_DEFINE_CONV_GEMM_LAUNCH_HALO(128, 2, false, 1)
}  // namespace sdm
