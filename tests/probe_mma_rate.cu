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
// `bg_gap` > 0: warps 1..3 write 512 B of shared memory each every ~bg_gap clocks while the MMAs run (stand-in for the TMA fill, the
// GroupNorm transform and the epilogue staging of the real kernel); bg_bytes_out receives the bytes they wrote.
template <int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* clk_out, int bg_gap = 0, long long* bg_bytes_out = nullptr, int random_data = 0, int commit_mode = 0) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_base = base + kWTiles * kWBytes;
  const uint32_t bar = base + kWTiles * kWBytes + 2 * kXBytes, slot = bar + 32;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = MODE == 2 ? cluster_ctarank_() : 0u;
  for (uint32_t i = threadIdx.x; i < (kWTiles * kWBytes + 2 * kXBytes) / 4; i += blockDim.x)
    {
      uint32_t v = 0x3c003c00u;  // fp16 1.0
      if (random_data) {         // two pseudo-random fp16 values in (-2, 2): realistic operand toggling (power), finite accumulators
        uint32_t h = (i + 1u + blockIdx.x * 7919u) * 2654435761u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        v = (h & 0x83ff83ffu) | 0x3c003c00u;
      }
      reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = v;
    }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);   // scratch barrier for the per-tap commits of commit_mode (nobody waits on it)
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
        // the real kernel's per-tap protocol: commit the weight stage back to the producer (1), + a fence as after an mbarrier wait (2),
        // + an actual (already satisfied) mbarrier probe (3)
        if (commit_mode >= 1) { if (MODE == 2) umma_commit_pair_(bar + 8); else umma_commit(bar + 8); }
        if (commit_mode >= 2) tc_fence_after();
        if (commit_mode >= 3) (void)mbar_test(bar + 8, 0);
      }
    }
    if (MODE == 2) umma_commit_pair_(bar); else umma_commit(bar);
    mbar_wait(bar, 0);
    clk_out[blockIdx.x] = clock64() - t0;
  } else if (MODE == 2 && threadIdx.x == 0) {
    mbar_wait(bar, 0);  // the multicast commit arrives here too
  } else if (bg_gap > 0 && warp >= 1) {
    // background shared-memory writes into the second halo slot's tail (never read by the MMAs of this probe: it only uses windows
    // of the first 40 KB of each slot) until the MMAs are done
    uint8_t* dst = smem_raw + (x_base + kXBytes + 41 * 1024 - smem_u32(smem_raw)) + (warp - 1) * 512 + (threadIdx.x & 31) * 16;
    long long n = 0;
    const long long t0 = clock64();
    while (!mbar_test(bar, 0)) {
      for (int j = 0; j < 64; ++j) {  // 64 stores per barrier probe; one store every bg_gap clocks
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(smem_u32(dst)), "r"((uint32_t)n), "r"(0u) : "memory");
        ++n;
        const long long until = t0 + n * bg_gap;
        while (clock64() < until) {}
      }
    }
    if ((threadIdx.x & 31) == 0 && bg_bytes_out) atomicAdd((unsigned long long*)bg_bytes_out, (unsigned long long)(n * 512));
  }
  tc_fence_before();
  if (MODE == 2) cluster_sync_(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (MODE == 2) tmem_dealloc_pair_<512>(tmem); else tmem_dealloc<512>(tmem);
  }
}

// The real kernel's operand pipeline without its epilogue / transform: lane 0 of warp 1 streams weight tiles (16 KB per tap) and
// lane 1 the pixel halo tiles (43.5 KB per slice) from global memory through TMA into the rings, warp 0 issues the MMAs behind the
// full / empty mbarriers exactly like conv_swap_halo_kernel.  `what`: 3 = both streams, 1 = weights only, 2 = halo tiles only.
struct PipeParams { CUtensorMap w_map, x_map; };
// `what` & 4: warps 2..5 read the OTHER accumulator buffer (TMEM columns 256..511) with tcgen05.ld once per two slices, like the
// epilogue of the real kernel; & 8: they also push the values through a shared-memory staging tile and store 64 KB to global memory
__global__ void __launch_bounds__(192, 1) mma_pipe_kernel(const __grid_constant__ PipeParams p, int iters, int what, long long* clk_out, uint4* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_base = base + kWTiles * kWBytes;
  const uint32_t bar = base + kWTiles * kWBytes + 2 * kXBytes, slot = bar + 8 * 24;
  auto wfull = [&](int i) { return bar + 8u * i; };
  auto wempty = [&](int i) { return bar + 8u * (kWTiles + i); };
  auto xfull = [&](int i) { return bar + 8u * (2 * kWTiles + i); };
  auto xempty = [&](int i) { return bar + 8u * (2 * kWTiles + 2 + i); };
  const uint32_t done = bar + 8u * (2 * kWTiles + 4);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (kWTiles * kWBytes + 2 * kXBytes) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWTiles; ++i) { mbar_init(wfull(i), 1); mbar_init(wempty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(xfull(i), 1); mbar_init(xempty(i), 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) tmem_alloc<512>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  const bool use_w = what & 1, use_x = what & 2;
  if (warp == 1 && lane == 0 && use_w) {
    int ws = 0; uint32_t wph = 0;
    for (int it = 0; it < iters; ++it)
      for (int tap = 0; tap < 9; ++tap) {
        mbar_wait(wempty(ws), wph ^ 1u);
        mbar_expect_tx(wfull(ws), kWBytes);
        tma_load_2d(base + ws * kWBytes, &p.w_map, wfull(ws), ((it * 9 + tap) % 64) * 64, 0);
        if (++ws == kWTiles) { ws = 0; wph ^= 1u; }
      }
  } else if (warp == 1 && lane == 1 && use_x) {
    int xs = 0; uint32_t xph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(xempty(xs), xph ^ 1u);
      mbar_expect_tx(xfull(xs), 340 * 128);
      tma_load_2d(x_base + xs * kXBytes, &p.x_map, xfull(xs), (it % 64) * 64, (blockIdx.x * 340) % 32768);
      tma_load_2d(x_base + xs * kXBytes + 170 * 128, &p.x_map, xfull(xs), (it % 64) * 64, (blockIdx.x * 340) % 32768 + 170);
      if (++xs == 2) { xs = 0; xph ^= 1u; }
    }
  } else if (threadIdx.x == 0) {
    int ws = 0, xs = 0; uint32_t wph = 0, xph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (use_x) mbar_wait(xfull(xs), xph);
      const uint32_t x_addr = x_base + xs * kXBytes;
      for (int tap = 0; tap < 9; ++tap) {
        if (use_w) mbar_wait(wfull(ws), wph);
        tc_fence_after();
        const uint64_t adesc = umma_desc_k128(base + ws * kWBytes);
        const uint64_t bdesc = umma_desc_k128_sbo(x_addr + (uint32_t)((tap / 3) * 10 + tap % 3) * 128u, 1280);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem, adesc + 2 * k, bdesc + 2 * k, umma_idesc_f16(256), (it | tap | k) != 0);
        umma_commit(wempty(ws));
        if (++ws == kWTiles) { ws = 0; wph ^= 1u; }
      }
      umma_commit(xempty(xs));
      if (++xs == 2) { xs = 0; xph ^= 1u; }
    }
    umma_commit(done);
    mbar_wait(done, 0);
    clk_out[blockIdx.x] = clock64() - t0;
  } else if (warp >= 2 && (what & 4)) {
    const int quad = warp & 3;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + 256;
    uint8_t* stg = smem_raw + (x_base + kXBytes + 41 * 1024 - smem_u32(smem_raw)) + quad * 512;
    uint32_t acc = 0;
    long long n = 0;
    const long long t0 = clock64();
    while (!mbar_test(done, 0)) {
      // one "tile" of epilogue work: 256 accumulator columns per lane
      for (int c = 0; c < 256; c += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(taddr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += r[i];
        if (what & 8) {
          *reinterpret_cast<uint4*>(stg + (threadIdx.x & 31) * 16) = make_uint4(r[0], r[1], r[2], r[3]);
          __syncwarp();
          const uint4 v = *reinterpret_cast<const uint4*>(stg + ((threadIdx.x + 1) & 31) * 16);
          for (int j = 0; j < 4; ++j) sink[((size_t)blockIdx.x * 4 + quad) * 1024 + ((n * 8 + c / 32) * 4 + j) % 1024 * 1 + 0] = v;
        }
      }
      ++n;
      // pace: the real epilogue drains one accumulator per tile = 72 MMAs x ~130-190 clk
      const long long until = t0 + n * 9400;
      while (clock64() < until && !mbar_test(done, 0)) {}
    }
    if (acc == 0x12345678u) sink[0] = make_uint4(acc, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

#include "tmap.h"
__global__ void fill_random_kernel(uint32_t* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)(i + 1u) * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = (h & 0x83ff83ffu) | 0x3c003c00u;  // two fp16 values in +-[1, 2)
  }
}
static void run_pipe(int what, int iters, int random_data = 0) {
  const int grid = 148;
  __half *wsrc = nullptr, *xsrc = nullptr;
  cudaMalloc(&wsrc, (size_t)128 * 4096 * 2); cudaMemset(wsrc, 0x3c, (size_t)128 * 4096 * 2);
  cudaMalloc(&xsrc, (size_t)65536 * 4096 * 2); cudaMemset(xsrc, 0x3c, (size_t)65536 * 4096 * 2);  // 512 MB: the halo tiles miss L2
  if (random_data) {
    fill_random_kernel<<<1024, 256>>>((uint32_t*)wsrc, (size_t)128 * 4096 / 2);
    fill_random_kernel<<<4096, 256>>>((uint32_t*)xsrc, (size_t)65536 * 4096 / 2);
    cudaDeviceSynchronize();
  }
  PipeParams p;
  { const uint64_t dims[2] = {4096, 128}; const uint64_t str[1] = {4096 * 2}; const uint32_t box[2] = {64, 128}; make_tmap(&p.w_map, wsrc, 2, dims, str, box); }
  { const uint64_t dims[2] = {4096, 65536}; const uint64_t str[1] = {4096 * 2}; const uint32_t box[2] = {64, 170}; make_tmap(&p.x_map, xsrc, 2, dims, str, box); }
  long long* d = nullptr;
  cudaMalloc(&d, sizeof(long long) * grid); cudaMemset(d, 0, sizeof(long long) * grid);
  uint4* sink = nullptr;
  cudaMalloc(&sink, (size_t)grid * 4 * 1024 * 16 + 64);
  cudaFuncSetAttribute(mma_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    mma_pipe_kernel<<<grid, 192, kSmem>>>(p, iters, what, d, sink);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("pipe: %s\n", cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  printf("pipe   %s TMA streams %s%s%s%s: %8.1f clk per MMA instruction, kernel %.3f ms -> %7.1f TFLOP/s\n", random_data ? "random-data" : "all-ones   ", (what & 1) ? "weights " : "", (what & 2) ? "halo-tiles " : "", (what & 4) ? "+tcgen05.ld of the other accumulator " : "",
         (what & 8) ? "+staging+STG" : "", mx / (36.0 * iters), best, 36.0 * iters * 2.0 * 128 * 256 * 16 * grid / (best * 1e-3) / 1e12);
  cudaFree(d); cudaFree(wsrc); cudaFree(xsrc);
}

template <int MODE>
static void run(const char* name, int grid, int iters, double flop_per_mma_per_sm, int bg_gap = 0, int random_data = 0, int commit_mode = 0) {
  long long* d = nullptr;
  cudaMalloc(&d, sizeof(long long) * (grid + 1));
  cudaMemset(d, 0, sizeof(long long) * (grid + 1));
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
      cudaLaunchKernelEx(&cfg, mma_rate_kernel<MODE>, iters, d, bg_gap, d + grid, random_data, commit_mode);
    } else {
      cudaMemset(d + grid, 0, sizeof(long long));
      mma_rate_kernel<MODE><<<grid, 128, kSmem>>>(iters, d, bg_gap, d + grid, random_data, commit_mode);
    }
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  std::vector<long long> h(grid + 1);
  cudaMemcpy(h.data(), d, sizeof(long long) * (grid + 1), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double mmas = 36.0 * iters;
  printf("%-6s %s commit_mode %d grid %3d bg_gap %4d: %8.1f clk per MMA instruction (issuing CTA), kernel %.3f ms -> %7.1f TFLOP/s on %d SMs; background stores %.1f B/clk/SM\n",
         name, random_data ? "random-data" : "all-ones   ", commit_mode, grid, bg_gap, mx / mmas, best, mmas * flop_per_mma_per_sm * grid / (best * 1e-3) / 1e12, grid, mx ? (double)h[grid] / grid / mx : 0.0);
  cudaFree(d);
}

int main() {
  const int iters = 2000;
  run<0>("ss1", 148, iters, 2.0 * 128 * 256 * 16);
  run<1>("ss1n", 148, iters, 2.0 * 128 * 128 * 16);
  run<2>("pair", 148, iters, 2.0 * 128 * 256 * 16);  // per SM: 128 of the 256 rows
  run<0>("ss1", 1, iters, 2.0 * 128 * 256 * 16);
  run<2>("pair", 2, iters, 2.0 * 128 * 256 * 16);
  for (int it : {2000, 20000}) {  // 5 ms and 50 ms kernels: does the power cap stretch the MMAs (in SM clocks)?
    run<0>("ss1", 148, it, 2.0 * 128 * 256 * 16, 0, 1);
    run<2>("pair", 148, it, 2.0 * 128 * 256 * 16, 0, 1);
  }
  run_pipe(3, iters); run_pipe(7, iters); run_pipe(15, iters);
  run_pipe(3, iters, 1); run_pipe(15, iters, 1); run_pipe(15, 10 * iters, 1);
  for (int cm : {1}) {
    run<0>("ss1", 148, iters, 2.0 * 128 * 256 * 16, 0, 0, cm);
    run<2>("pair", 148, iters, 2.0 * 128 * 256 * 16, 0, 0, cm);
  }
  return 0;
}
