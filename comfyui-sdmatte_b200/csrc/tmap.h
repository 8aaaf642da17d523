// TMA tensor-map construction (driver entry point resolved at run time; no link against libcuda).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace sdm {
// fp16 tensor, 128-byte swizzle, zero fill for out-of-bounds elements.
// dims[0] is the contiguous dimension; strides (bytes) are given for dims[1..rank-1].
void make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box);
}  // namespace sdm
