#include "tmap.h"
#include "common.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace sdm {

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static void resolve_encode() {
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
}

void make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box) {
  resolve_encode();
  SDM_CHECK(g_encode != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  SDM_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    SDM_CHECK(box[i] >= 1 && box[i] <= 256, "TMA box dim out of range");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    SDM_CHECK(strides_bytes[i] % 16 == 0, "TMA stride must be a multiple of 16 bytes");
  }
  SDM_CHECK(box[0] * 2 == 128, "inner box must be 128 bytes for SWIZZLE_128B operands");
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    throw Error{"cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (rank " + std::to_string(rank) +
                ", dims " + std::to_string(dims[0]) + "," + std::to_string(rank > 1 ? dims[1] : 0) + "," +
                std::to_string(rank > 2 ? dims[2] : 0) + "," + std::to_string(rank > 3 ? dims[3] : 0) + ")"};
  }
}

}  // namespace sdm
