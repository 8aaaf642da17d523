// Hardware probe (test infrastructure, not on the product path): can a tcgen05 A-operand descriptor address a SHIFTED
// 16 x 8 pixel window inside a (16+2) x (8+2) halo box that TMA wrote with the 128-byte swizzle?  The window's 8-pixel rows
// are 10 pixels (1280 B) apart (SBO = 1280, not the dense 1024) and start at (dy*10 + dx) * 128 B, i.e. not on a 1024-byte
// swizzle-atom boundary (descriptor base_offset = (start >> 7) & 7, or 0).  A = window, B = 64 x 64 identity, D = A.
// This decides whether the 3x3 convolutions can keep ONE halo tile per 64-channel slice resident in shared memory and issue
// all nine taps from it (9x less activation fill traffic) instead of one TMA box per tap.
#include "common.cuh"
#include "kernels.h"
#include "tmap.h"

namespace sdm {

struct alignas(64) ProbeParams {
  CUtensorMap a_map, i_map;
  float* out;
  int dy, dx, mode;
};

__device__ __forceinline__ uint64_t umma_desc_k128_ex(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128) halo_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 24 * 1024, bar = base + 32 * 1024, bar2 = bar + 8, slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) tmem_alloc<64>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 64 * 2 * 10 * 18 + 64 * 128);
    tma_load_4d(a_base, &p.a_map, bar, 0, -1, -1, 0);  // halo box: x in [-1, 9), y in [-1, 17), zero filled outside
    tma_load_2d(b_base, &p.i_map, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = a_base + (uint32_t)(p.dy * 10 + p.dx) * 128u;
    const uint32_t bo = p.mode == 1 ? ((a_addr >> 7) & 7u) : 0u;
    const uint64_t ad = umma_desc_k128_ex(a_addr, 1280, bo);
    const uint64_t bd = umma_desc_k128(b_base);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16(tmem, ad + 2 * k, bd + 2 * k, umma_idesc_f16(64), k != 0);
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  uint32_t r[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 64; c += 32) {
    __syncwarp();
    tmem_ld32(taddr + c, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) p.out[(size_t)threadIdx.x * 64 + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<64>(tmem);
  }
}

void probe_halo_run(const __half* x, const __half* eye, float* out, int dy, int dx, int mode, cudaStream_t st) {
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  {
    const uint64_t dims[4] = {64, 8, 16, 1};
    const uint64_t str[3] = {64 * 2, 8 * 64 * 2, 16 * 8 * 64 * 2};
    const uint32_t box[4] = {64, 10, 18, 1};
    make_tmap(&p.a_map, x, 4, dims, str, box);
  }
  {
    const uint64_t dims[2] = {64, 64};
    const uint64_t str[1] = {64 * 2};
    const uint32_t box[2] = {64, 64};
    make_tmap(&p.i_map, eye, 2, dims, str, box);
  }
  p.out = out; p.dy = dy; p.dx = dx; p.mode = mode;
  const int smem = 34 * 1024 + 1024;
  SDM_CUDA_OK(cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  halo_probe_kernel<<<1, 128, smem, st>>>(p);
  SDM_CUDA_OK(cudaGetLastError());
}

}  // namespace sdm
