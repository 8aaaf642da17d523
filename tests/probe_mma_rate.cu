// Hardware probe (test infrastructure, not on the product path): how fast does the tensor core execute back-to-back tcgen05 MMAs
// whose operands sit in shared memory, for the shapes the 3x3 convolution kernels use?
//   ss1  : cta_group::1, M128 x N256 x K16, A (4 KB) and B (8 KB) from shared memory      <- conv_swap_halo_kernel today
//   ss1n : cta_group::1, M128 x N128 x K16
//   pair : cta_group::2, M256 x N256 x K16: per CTA 128 rows of A (4 KB) + 128 rows of B (4 KB)   <- a CTA pair sharing the pixel tile
// SDM_GEMM_PROF on the real kernel (profiles/r2p_swh_prof.txt) shows its MMA issuer blocked on ISSUE 90 % of the time at 189-217 clk
// per M128xN256xK16 MMA (131 clk at the nominal 8 kFLOP/clk/SM): the operand fetch, not the producer, is the limit.  This probe
// decides whether a pair kernel could escape it.   Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I comfyui-sdmatte_b200/csrc -I include tests/probe_mma_rate.cu -o /tmp/probe_mma_rate && /tmp/probe_mma_rate
#include "common.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace sdm {
void set_last_error(const std::string&) {}
}
using namespace sdm;

__device__ __forceinline__ uint32_t cluster_ctarank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_() { asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
template <int NCOLS> __device__ __forceinline__ void tmem_alloc_pair_(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_dealloc_pair_(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_f16_pair_(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(adesc),
               "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void umma_commit_pair_(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_f16_m256_(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24); }

constexpr int kWTiles = 6, kWBytes = 16 * 1024, kXBytes = 44 * 1024;  // the real kernel's rings
constexpr int kSmem = kWTiles * kWBytes + 2 * kXBytes + 1024 + 256;

// MODE 0: ss1 (N = 256), 1: ss1n (N = 128), 2: pair
template <int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* clk_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_base = base + kWTiles * kWBytes;
  const uint32_t bar = base + kWTiles * kWBytes + 2 * kXBytes, slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = MODE == 2 ? cluster_ctarank_() : 0u;
  for (uint32_t i = threadIdx.x; i < (kWTiles * kWBytes + 2 * kXBytes) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    if (MODE == 2) tmem_alloc_pair_<512>(slot);
    else tmem_alloc<512>(slot);
  }
  tc_fence_before();
  if (MODE == 2) cluster_sync_(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0 && rank == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {          // one "slice": nine taps of four K16 MMAs, shifted windows of the halo tile
      const uint32_t x_addr = x_base + (it & 1) * kXBytes;
      for (int tap = 0; tap < 9; ++tap) {
        const uint64_t adesc = umma_desc_k128(base + ((it * 9 + tap) % kWTiles) * kWBytes);
        const uint64_t bdesc = umma_desc_k128_sbo(x_addr + (uint32_t)((tap / 3) * 10 + tap % 3) * 128u, 1280);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (MODE == 2) umma_f16_pair_(tmem, adesc + 2 * k, bdesc + 2 * k, umma_idesc_f16_m256_(256), (it | tap | k) != 0);
          else umma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, umma_idesc_f16(MODE == 1 ? 128 : 256), (it | tap | k) != 0);
        }
      }
    }
    if (MODE == 2) umma_commit_pair_(bar); else umma_commit(bar);
    mbar_wait(bar, 0);
    clk_out[blockIdx.x] = clock64() - t0;
  } else if (MODE == 2 && threadIdx.x == 0) {
    mbar_wait(bar, 0);  // the multicast commit arrives here too
  }
  tc_fence_before();
  if (MODE == 2) cluster_sync_(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (MODE == 2) tmem_dealloc_pair_<512>(tmem); else tmem_dealloc<512>(tmem);
  }
}

template <int MODE>
static void run(const char* name, int grid, int iters, double flop_per_mma_per_sm) {
  long long* d = nullptr;
  cudaMalloc(&d, sizeof(long long) * grid);
  cudaMemset(d, 0, sizeof(long long) * grid);
  cudaFuncSetAttribute(mma_rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (MODE == 2) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = kSmem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, mma_rate_kernel<MODE>, iters, d);
    } else {
      mma_rate_kernel<MODE><<<grid, 128, kSmem>>>(iters, d);
    }
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  const double mmas = 36.0 * iters;
  printf("%-6s grid %3d: %8.1f clk per MMA instruction (issuing CTA), kernel %.3f ms -> %7.1f TFLOP/s on %d SMs\n", name, grid, mx / mmas, best,
         mmas * flop_per_mma_per_sm * grid / (best * 1e-3) / 1e12, grid);
  cudaFree(d);
}

int main() {
  const int iters = 2000;
  run<0>("ss1", 148, iters, 2.0 * 128 * 256 * 16);
  run<1>("ss1n", 148, iters, 2.0 * 128 * 128 * 16);
  run<2>("pair", 148, iters, 2.0 * 128 * 256 * 16);  // per SM: 128 of the 256 rows
  run<0>("ss1", 1, iters, 2.0 * 128 * 256 * 16);
  run<2>("pair", 2, iters, 2.0 * 128 * 256 * 16);
  return 0;
}
