// 3x3 stride-1 convolution with 128 output channels per tile, OPERANDS SWAPPED: D^T[channel][pixel] = W[channel][k] . X[pixel][k]^T.
//
// Why (measured on B200, SDM_GEMM_PROF r1s): a tcgen05 SS-MMA fetches both operands from shared memory for every instruction,
// and the fetch sustains ~75 B/clk/SM.  With pixels on M (128 rows) and the 128 output channels on N, one M128 x N128 x K16 MMA
// needs 4 KB + 4 KB for 64 clk of math -> 109 clk measured, 59 % of the tensor peak (the VAE's 128-channel convs at 1024^2,
// 47 ms of a 290 ms step).  Neither fewer fill bytes (resident halo tile: 125 clk / MMA) nor a CTA pair (cta_group::2: 129 clk)
// changes that ratio.  Swapping the roles does: weights on M (128 channels), 256 PIXELS on N -> one M128 x N256 x K16 MMA needs
// 4 KB + 8 KB for 128 clk of math = the ratio of the 256-channel layers, which run at 80 % (about the cuBLAS sustained figure).
//
// Accumulator: TMEM lane = output channel, column = pixel of a 16 x 16 patch; two 256-column stages (512 columns).
// Epilogue: thread = one channel.  Bias is a per-thread scalar, the GroupNorm partial sums of the STORED fp16 values are plain
// per-thread accumulators (no cross-thread reduction at all); the [channel][pixel] -> NHWC transposition goes through a
// 32 pixel x 32 channel fp16 staging tile per warp: 64-byte channel runs per pixel, written with 16-byte stores.
// Residual: K steps with A = 128 x 64 slice of the identity matrix, B = the residual tile (as in conv_gemm_kernel).
// Replaces the same reference ops as conv_gemm_kernel (ResnetBlock2D conv1 / conv2 of the VAE).
#include "umma_gemm.cuh"

namespace sdm {

namespace sw {
constexpr int kWBytes = 128 * 128;   // 128 channels x 64 k (fp16)
constexpr int kXBytes = 256 * 128;   // 256 pixels x 64 k
constexpr int kStageBytes = kWBytes + kXBytes;
constexpr int kStages = 4;
constexpr int kStgBytes = 4 * 2048;  // per epilogue warp: 32 pixels x 32 channels fp16
constexpr int kSmem = kStages * kStageBytes + 1024 + 256 + kStgBytes;
constexpr int kThreads = 192;
}  // namespace sw

__global__ void __launch_bounds__(sw::kThreads, 1) conv_swap_kernel(const __grid_constant__ ConvGemmParams p) {
  using namespace sw;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // tile = (16 x 16 pixel patch, 128-channel slice); p.m_tiles = patches over all samples, p.n_tiles = N / 128
  const int nres = p.has_res ? 2 : 0;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      tma_prefetch_desc(&p.a_map[0]); tma_prefetch_desc(&p.a_map[1]); tma_prefetch_desc(&p.b_map);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles) * 128;
        const int mt = tile / p.n_tiles;
        const int x0 = (mt % p.tiles_x) * 16, y0 = ((mt / p.tiles_x) % p.tiles_y) * 16, b = mt / (p.tiles_x * p.tiles_y);
        for (int tap = 0; tap < 9; ++tap) {
          int koff = tap * p.cin_total;
          for (int s = 0; s < p.nsrc; ++s) {
            for (int c0 = 0; c0 < p.src_c[s]; c0 += 64) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t w_dst = smem_base + stage * kStageBytes;
              mbar_expect_tx(full_bar(stage), kStageBytes);
              tma_load_2d(w_dst, &p.b_map, full_bar(stage), koff + c0, n0);
              tma_load_4d(w_dst + kWBytes, &p.a_map[s], full_bar(stage), c0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            koff += p.src_c[s];
          }
        }
        for (int i = 0; i < nres; ++i) {  // D^T[c][pix] += I[c][64 i + k] . R[pix][n0 + 64 i + k]
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t w_dst = smem_base + stage * kStageBytes;
          mbar_expect_tx(full_bar(stage), kStageBytes);
          tma_load_2d(w_dst, &p.i_map, full_bar(stage), 64 * i, 0);
          tma_load_4d(w_dst + kWBytes, &p.r_map, full_bar(stage), n0 + 64 * i, x0, y0, b);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (whole warp, one elected lane issues: see conv_swap_halo.cu) ==============
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_f16(256);
      int chunks = 0;
      for (int s = 0; s < p.nsrc; ++s) chunks += p.src_c[s] >> 6;
      const int ksteps = 9 * chunks + nres;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t w_addr = smem_base + stage * kStageBytes;
          const uint64_t adesc = umma_desc_k128(w_addr), bdesc = umma_desc_k128(w_addr + kWBytes);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | k) != 0);
            umma_commit(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (leader) umma_commit(tfull_bar(acc));
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue: thread = output channel ==============================
    const int quad = warp & 3;
    const int m = quad * 32 + lane;  // channel within the 128-channel slice == TMEM lane
    uint8_t* stg = smem_raw + (bar_base - smem_u32(smem_raw)) + 256 + (warp - 2) * 2048;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int per_image = p.tiles_x * p.tiles_y;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int n0 = (tile % p.n_tiles) * 128;
      const int mt = tile / p.n_tiles;
      const int t_img = mt % per_image;
      const int x0 = (t_img % p.tiles_x) * 16, y0 = (t_img / p.tiles_x) * 16, b = mt / per_image;
      const float bias = p.bias ? p.bias[(p.bias_sel ? (long long)p.bias_sel[b] * p.N : 0) + n0 + m] : 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256;
      __half* obase = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride + n0 + quad * 32;
      // this lane's share of the transposed store: 4 x (pixel, 16-byte piece of the warp's 64-byte channel run)
      float sum = 0.f, sq = 0.f;
      uint32_t ra[32], rb[32];
      // one block = 32 pixels (patch rows 2 blk, 2 blk + 1) held in r; `nxt` receives the following block meanwhile
      auto block = [&](const uint32_t (&r)[32], uint32_t (&nxt)[32], int blk) {
        tmem_ld_wait();
        __syncwarp();  // the previous block's staging reads are done
        if (blk + 1 < 8) tmem_ld32(taddr + (blk + 1) * 32, nxt);
        const int ya = y0 + 2 * blk;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const __half h = __float2half_rn(fmaf(__uint_as_float(r[i]), p.scale, bias));
          const bool ok = (x0 + (i & 15) < p.W) && (ya + (i >> 4) < p.H);
          const float f = ok ? __half2float(h) : 0.f;
          sum += f;
          sq = fmaf(f, f, sq);
          *reinterpret_cast<__half*>(stg + i * 64 + lane * 2) = h;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int idx = lane + 32 * t;
          const int px = idx >> 2, piece = idx & 3;
          const int xx = x0 + (px & 15), yy = ya + (px >> 4);
          const uint4 v = *reinterpret_cast<const uint4*>(stg + px * 64 + piece * 16);
          if (xx < p.W && yy < p.H) *reinterpret_cast<uint4*>(obase + ((long long)yy * p.W + xx) * p.out_ld + piece * 8) = v;
        }
        if (p.stats && (blk & 3) == 3) {  // 128 pixels done: one GroupNorm-partials slot (two per 16 x 16 patch)
          const long long slot = (long long)b * (2 * per_image) + 2 * t_img + (blk >> 2);
          *reinterpret_cast<float2*>(p.stats + (slot * p.N + n0 + m) * 2) = make_float2(sum, sq);
          sum = 0.f;
          sq = 0.f;
        }
      };
      __syncwarp();
      tmem_ld32(taddr, ra);
#pragma unroll 1
      for (int b2 = 0; b2 < 4; ++b2) {
        block(ra, rb, 2 * b2);
        block(rb, ra, 2 * b2 + 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

void conv_swap_launch(const ConvGemmParams& p, int grid, cudaStream_t st) {
  static PerDeviceOnce attr;
  attr([] { SDM_CUDA_OK(cudaFuncSetAttribute(conv_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sw::kSmem)); });
  conv_swap_kernel<<<grid, sw::kThreads, sw::kSmem, st>>>(p);
  SDM_CUDA_OK(cudaGetLastError());
}

}  // namespace sdm
