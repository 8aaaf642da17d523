// Implicit-GEMM convolution / linear kernel on tcgen05 (sm_100a).
//
// One kernel template covers every dense contraction of the matte path except the d=64 attention:
//   * Conv2d 3x3 stride 1 (pad 1), stride 2 (UNet pad 1 / VAE asymmetric pad), 1x1 shortcut convs
//     (reference: diffusers ResnetBlock2D / Downsample2D / Upsample2D built at
//      /root/reference/src/utils/replace.py:239,268,321)
//   * nn.Linear (proj_in/out, to_q/k/v/out, GEGLU, FF-out; reference Attention/FeedForward)
//   * the VAE mid-block attention GEMMs (QK^T -> fp32 scores, P·V), the skinny output convs and the alpha head
//
// A operand ("activation"): NHWC fp16 tensor(s) read through TMA tensor maps (C, W, H, B).
//   A tile of 128 output pixels is a (tw x th) patch of one image; for tap (dy,dx) the producer
//   issues ONE 4-D TMA box load at (c0, x0+dx, y0+dy, b): out-of-image coordinates are zero-filled
//   by the TMA unit, which *is* the convolution padding.  Box rows land in shared memory as
//   128 rows x 128 B with the 128-byte swizzle, exactly the K-major layout tcgen05.mma expects.
// B operand ("weight"): [N][Ktot] fp16, K contiguous, Ktot ordered (tap, cin); 2-D map (or 3-D batched).
// Residual: "+ x" of ResnetBlock2D / Attention / FeedForward / Transformer2DModel is folded into the contraction as
//   extra K steps  A = residual tile (TMA, 64 channels),  B = a 64x64 identity block, issued as N=64 MMAs into the matching
//   64 accumulator columns: the add costs about one K step per tile and no epilogue work or uncoalesced loads
//   (fp32 accumulate of fp16 * 1.0 is exact).
// D: MT x (128 x BLOCK_N) fp32 accumulators in TMEM, double buffered so the epilogue of tile i overlaps
//   the main loop of tile i+1.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM
//   alloc), warps 2..5 = epilogue (TMEM -> registers -> bias/activation -> smem staging -> row-contiguous stores).
// The epilogue flavour is a template parameter: one instantiation contains one code path (the all-in-one version
// had grown to 135 KB of SASS per kernel and thrashed the instruction cache, r1d).
#pragma once
#include "common.cuh"

#include <type_traits>

namespace sdm {

enum EpiMode : int {
  EPI_F16 = 0,     // out[pixel][n] fp16  (+bias, optional 2x nearest-upsample scatter)
  EPI_F16_T = 1,   // out[b][n][pixel] fp16 (transposed; used for V^T)
  EPI_GEGLU = 2,   // out[pixel][n/2] = fp16(v) * gelu(fp16(g)), tile = [BLOCK_N/2 value | BLOCK_N/2 gate]
  EPI_F32 = 3,     // out[pixel][n] fp32 = scale * acc
  EPI_ALPHA = 4,   // N>=3: alpha[pixel] = (clip(mean(fp16(c0..c2)), -1, 1) + 1) / 2 ; out2[pixel] = pre-clip mean (meta_arch.py:258-260)
  EPI_SKINNY = 5,  // out[pixel][0..n_store) fp16, n_store < 8 (scalar stores), optional fp16 division
};

struct alignas(64) ConvGemmParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  CUtensorMap r_map;  // residual activation (same geometry as a_map[0]) or unused
  CUtensorMap i_map;  // identity matrix [NI][NI] fp16 (B operand of the residual K steps)
  int B, H, W;        // output pixel grid (per batch element)
  int tw, th;         // tile patch, tw*th == 128
  int tiles_x, tiles_y;
  int N;              // GEMM N (output channels before GEGLU halving)
  int n_tiles;
  int total_tiles;
  int m_tiles;        // number of 128-pixel M tiles (all batch elements)
  int ntaps, nsrc;
  int src_c[2];       // channels per A source (multiples of 64)
  int cin_total;      // weight K per tap
  signed char tap_map[9], tap_dx[9], tap_dy[9];
  int b_batched;      // weight map has a batch coordinate
  int has_res;        // residual K steps present
  int prof;           // SDM_GEMM_PROF=1: CTAs 0/1 print the cycles their producer / MMA / epilogue threads spent waiting
  int mode;
  int ups2;           // EPI_F16 only: write each pixel to the 2x2 block of a (2H,2W) output
  int poly;           // conv_swap_halo only: 0, or 1 + 2 py + px = polyphase component (py, px) of "nearest x2 upsample -> 3x3 conv"
  int stats_bslots, stats_slot0;  // conv_swap_halo: GroupNorm-partials slots per sample / first slot of this launch
  int res_mix;        // conv_swap_halo: interleave the residual K slices with the input slices (SDM_SWH_MIX=0: A/B switch, residual last)
  void* out;
  long long out_ld;       // elements between consecutive pixels (EPI_F16/F32/GEGLU) or row length (EPI_F16_T)
  long long out_bstride;  // elements per batch element
  const float* bias;      // [nsel][N] fp32 or null
  const int* bias_sel;    // per-batch row selector into bias, or null
  float scale;
  float post_div;   // EPI_SKINNY only: fp16(result) / post_div, rounded again (label_latent / scaling_factor); 1 = off
  int n_store;      // EPI_SKINNY: number of output columns stored
  void* out2;       // EPI_ALPHA: optional pre-clip mean
  const float* gn_ab;  // fused input GroupNorm (conv_swap_halo_kernel<true>): [B][cin_total][2] (scale, shift), or null
  int gn_silu;
  float* stats;     // EPI_F16 (no ups2): per-(M tile, channel) partial (sum, sumsq) of the STORED fp16 values for the next GroupNorm:
                    // stats[((b * tiles_per_image + tile_in_image) * N + col) * 2 + {0,1}]  (deterministic: one writer per slot)
};

// MT = number of 128-row M sub-tiles a CTA processes against ONE B tile (MT=2 halves the weight traffic per FLOP:
// the N=128 VAE convs were L2->SM bandwidth bound at 128 B/clk/SM with MT=1, r1c: 545-885 TFLOP/s vs 1300-1440 for N=256)
// EWG: number of epilogue warpgroups.  With 2, tile i of a CTA is drained by warpgroup i%2 (MT=1) or the two M sub-tiles
// of a tile are drained concurrently (MT=2): twice the epilogues in flight for the GEMMs whose short K loop cannot hide
// one (r1i: linears 350-600 TFLOP/s, the N=128 VAE convs ~1000 vs ~1400 for long-K layers).
// (Round 1 also carried a two-CTAs-per-SM "light" config, an L2 prefetch of the next tile and a CTA-pair (tcgen05 cta_group::2,
// M = 256) variant of this kernel; all three were measured slower than the configurations below — profiles/r1r_*, r1s_*,
// DESIGN.md section 3 — and were removed in round 2.)
// HALO (r1s): 3x3 stride-1 convolutions keep ONE (8+2) x (16+2)-pixel halo tile per M sub-tile and 64-channel slice resident in
// shared memory and issue all NINE taps from it: the A operand of tap (dy, dx) is the descriptor of the same tile started
// (dy*10 + dx) pixel rows later with SBO = 10 rows (1280 B) — the 128-byte swizzle is a function of the absolute shared-memory
// address, so shifted windows of a TMA-written tile are valid operands (tests/probe_halo.py, all 9 windows exact on B200).
// Measured motivation (SDM_GEMM_PROF, r1s): the 128->128 convs at 1024^2 moved 51 B/clk/SM from L2 into shared memory (the nine
// taps re-fetch the activation tile nine times) and each MMA took 109 clk instead of 64: shared-memory bandwidth = 8 KB operand
// reads + 6 KB TMA writes per MMA.  With the halo tile the fill traffic per 64-channel slice drops from 9 x 16 KB to 23 KB per
// sub-tile.  Shared memory = 2 halo slots (x MT) + a ring of weight tiles.
template <int BLOCK_N, int MT = 1, int EWG = 1, bool HALO = false>
struct ConvGemmCfg {
  static constexpr int kABytes = 128 * 128;          // 128 rows x 64 fp16 (per M sub-tile)
  static constexpr int kBBytes = BLOCK_N * 128;      // B rows x 64 fp16
  static constexpr int kHaloBytes = 23 * 1024;       // 10 x 18 pixel rows of 128 B (23 040 B) rounded up to the 1024-byte swizzle atom
  static constexpr int kHaloTx = 10 * 18 * 128;      // bytes one halo box delivers
  static constexpr int kASlots = 2;
  static constexpr int kEpiBytes = EWG * (4 * 4096 /*staging*/ + 2 * 2048 /*GroupNorm partials, double buffered*/);
  static constexpr int kBudget = 232448 - 1024 - 256 - kEpiBytes - 512;
  static constexpr int kStageBytes = HALO ? kBBytes : MT * kABytes + kBBytes;
  static constexpr int kRingBudget = HALO ? kBudget - kASlots * MT * kHaloBytes : kBudget;
  static constexpr int kStages = (kRingBudget / kStageBytes) > 8 ? 8 : (kRingBudget / kStageBytes);
  static constexpr int kSubStride = BLOCK_N < 32 ? 32 : BLOCK_N;  // TMEM columns per M sub-tile accumulator
  static constexpr int kAccStride = MT * kSubStride;              // TMEM columns between the two accumulator stages
  static constexpr int kTmemCols = (2 * kAccStride <= 64) ? 64 : (2 * kAccStride <= 128) ? 128 : (2 * kAccStride <= 256) ? 256 : 512;
  static constexpr int kPipeBytes = kStages * kStageBytes + (HALO ? kASlots * MT * kHaloBytes : 0);  // ring first, halo slots behind it
  static constexpr int kSmemBytes = kPipeBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiBytes;
  static constexpr int kThreads = 64 + 128 * EWG;
  static_assert(kStages >= 2, "pipeline depth");
};

// residual K steps of the tile starting at output column n0: one per 64-channel slice of the tile's own columns.
// Step i multiplies the residual slice [n0+64i, n0+64i+64) with the 64x64 identity block and accumulates into the
// accumulator columns [64i, 64i+64) only (an N=64 — or N=32 for the ragged end of a 160-wide tile — MMA), so the whole
// residual costs about ONE full K step per tile regardless of BLOCK_N.
template <int BLOCK_N>
__device__ __forceinline__ int residual_steps(int n0, int N) {
  return (min(BLOCK_N, N - n0) + 63) >> 6;
}

template <int BLOCK_N, int MT, int MODE, bool UPS2, int EWG = 1, bool HALO = false>
__global__ void __launch_bounds__(64 + 128 * EWG, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BLOCK_N, MT, EWG, HALO>;
  static_assert(!HALO || MODE == EPI_F16, "halo configuration");
  static_assert(EWG == 1 || EWG == 2, "one or two epilogue warpgroups");
  const int tile_first = (int)blockIdx.x;
  const int tile_step = (int)gridDim.x;
  constexpr int SUBS = MT;  // M sub-tiles per tile
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kPipeBytes;
  const uint32_t halo_base = smem_base + kStages * Cfg::kStageBytes;  // HALO: kASlots x MT halo tiles
  // barrier layout (8 B each): full[kStages], empty[kStages], tfull[2], tempty[2], then tmem ptr (4 B), then (HALO) afull[2], aempty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  auto afull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 5 + a); };
  auto aempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 7 + a); };
  static_assert(8 * (2 * 8 + 9) <= 256, "barrier area");
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), (EWG == 2 && MT == 2) ? 8 : 4);
      if (HALO) { mbar_init(afull_bar(a), 1); mbar_init(aempty_bar(a), 1); }
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int chunks_per_tap = 0;
  for (int s = 0; s < p.nsrc; ++s) chunks_per_tap += p.src_c[s] >> 6;
  const int num_ksteps = p.ntaps * chunks_per_tap;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
      tma_prefetch_desc(&p.b_map);
      int stage = 0;
      uint32_t phase = 0;
      int aslot = 0;        // HALO: halo-slot ring
      uint32_t aphase = 0;
      long long prof_wait = 0;
      const long long prof_t0 = p.prof ? clock64() : 0;
      for (int tile = tile_first; tile < p.total_tiles; tile += tile_step) {
        const int nt = tile % p.n_tiles;
        const int n0 = nt * BLOCK_N;
        int x0[MT], y0[MT], bb[MT];
#pragma unroll
        for (int u = 0; u < MT; ++u) {
          const int mt = (tile / p.n_tiles) * SUBS + u;
          const int tx = mt % p.tiles_x;
          const int ty = (mt / p.tiles_x) % p.tiles_y;
          x0[u] = tx * p.tw; y0[u] = ty * p.th;
          bb[u] = mt < p.m_tiles ? mt / (p.tiles_x * p.tiles_y) : p.B;  // past-the-end sub-tile: batch index out of range -> zero fill
        }
        if constexpr (HALO) {
          // per 64-channel slice: ONE halo box per M sub-tile (x0-1 .. x0+8, y0-1 .. y0+16; zero fill = conv padding), then the
          // nine weight tiles of the slice through the ring
          int coff = 0;
          for (int s = 0; s < p.nsrc; ++s) {
            for (int c0 = 0; c0 < p.src_c[s]; c0 += 64) {
              mbar_wait(aempty_bar(aslot), aphase ^ 1u);
              mbar_expect_tx(afull_bar(aslot), MT * Cfg::kHaloTx);
#pragma unroll
              for (int u = 0; u < MT; ++u)
                tma_load_4d(halo_base + (aslot * MT + u) * Cfg::kHaloBytes, &p.a_map[s], afull_bar(aslot), c0, x0[u] - 1, y0[u] - 1, bb[u]);
              if (++aslot == Cfg::kASlots) { aslot = 0; aphase ^= 1u; }
              for (int tap = 0; tap < 9; ++tap) {
                if (p.prof) { const long long t = clock64(); mbar_wait(empty_bar(stage), phase ^ 1u); prof_wait += clock64() - t; }
                else mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_expect_tx(full_bar(stage), Cfg::kBBytes);
                const uint32_t b_dst = smem_base + stage * Cfg::kStageBytes;
                if (p.b_batched) tma_load_3d(b_dst, &p.b_map, full_bar(stage), tap * p.cin_total + coff + c0, n0, bb[0]);
                else tma_load_2d(b_dst, &p.b_map, full_bar(stage), tap * p.cin_total + coff + c0, n0);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
              }
            }
            coff += p.src_c[s];
          }
          if (p.has_res) {  // residual K steps: dense 8 x 16 residual boxes in a halo slot, identity block in a ring stage
            const int nres = residual_steps<BLOCK_N>(n0, p.N);
            for (int i = 0; i < nres; ++i) {
              mbar_wait(aempty_bar(aslot), aphase ^ 1u);
              mbar_expect_tx(afull_bar(aslot), MT * Cfg::kABytes);
#pragma unroll
              for (int u = 0; u < MT; ++u)
                tma_load_4d(halo_base + (aslot * MT + u) * Cfg::kHaloBytes, &p.r_map, afull_bar(aslot), n0 + i * 64, x0[u], y0[u], bb[u]);
              if (++aslot == Cfg::kASlots) { aslot = 0; aphase ^= 1u; }
              mbar_wait(empty_bar(stage), phase ^ 1u);
              mbar_expect_tx(full_bar(stage), 64 * 128);
              tma_load_2d(smem_base + stage * Cfg::kStageBytes, &p.i_map, full_bar(stage), 0, 0);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
          }
        } else {
        for (int tap = 0; tap < p.ntaps; ++tap) {
          int koff = tap * p.cin_total;
          for (int s = 0; s < p.nsrc; ++s) {
            const CUtensorMap* amap = &p.a_map[p.tap_map[tap] + s];
            for (int c0 = 0; c0 < p.src_c[s]; c0 += 64) {
              if (p.prof) { const long long t = clock64(); mbar_wait(empty_bar(stage), phase ^ 1u); prof_wait += clock64() - t; }
              else mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
              const uint32_t b_dst = a_dst + MT * Cfg::kABytes;
              mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
#pragma unroll
              for (int u = 0; u < MT; ++u)
                tma_load_4d(a_dst + u * Cfg::kABytes, amap, full_bar(stage), c0, x0[u] + p.tap_dx[tap], y0[u] + p.tap_dy[tap], bb[u]);
              if (p.b_batched)
                tma_load_3d(b_dst, &p.b_map, full_bar(stage), koff + c0, n0, bb[0]);
              else
                tma_load_2d(b_dst, &p.b_map, full_bar(stage), koff + c0, n0);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            koff += p.src_c[s];
          }
        }
        if (p.has_res) {  // residual as K steps against the 64x64 identity block
          const int nres = residual_steps<BLOCK_N>(n0, p.N);
          for (int i = 0; i < nres; ++i) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
            const uint32_t b_dst = a_dst + MT * Cfg::kABytes;
            mbar_expect_tx(full_bar(stage), MT * Cfg::kABytes + 64 * 128);
#pragma unroll
            for (int u = 0; u < MT; ++u) tma_load_4d(a_dst + u * Cfg::kABytes, &p.r_map, full_bar(stage), n0 + i * 64, x0[u], y0[u], bb[u]);
            tma_load_2d(b_dst, &p.i_map, full_bar(stage), 0, 0);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
        }  // !HALO
      }
      if (p.prof && blockIdx.x < 2)
        printf("sdm prof: cta %d producer   total %lld clk, waiting for a free stage %lld\n", blockIdx.x, clock64() - prof_t0, prof_wait);
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // The whole warp runs the loop convergently, one elected lane issues (see conv_swap_halo.cu: inside an `if (lane == 0)` every
    // tcgen05.mma was wrapped in an ELECT / R2UR.BROADCAST loop by the compiler and the issuing thread became the limit).
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc = umma_idesc_f16(BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int aslot = 0;        // HALO: halo-slot ring
      uint32_t aphase = 0;
      long long prof_full = 0, prof_tempty = 0;
      const long long prof_t0 = p.prof ? clock64() : 0;
      for (int tile = tile_first; tile < p.total_tiles; tile += tile_step) {
        const int n0 = (tile % p.n_tiles) * BLOCK_N;
        const int nres = p.has_res ? residual_steps<BLOCK_N>(n0, p.N) : 0;
        if (p.prof) { const long long t = clock64(); mbar_wait(tempty_bar(acc), acc_phase ^ 1u); prof_tempty += clock64() - t; }
        else mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        if constexpr (HALO) {
          // 64-channel slices: wait for the slice's halo tiles once, then nine taps x MT sub-tiles x four K16 steps, each tap
          // reading the SAME halo tile through a descriptor started (dy*10 + dx) pixel rows later (rows of a window are 10 rows apart)
          const int nslices = chunks_per_tap;
          for (int sl = 0; sl < nslices + nres; ++sl) {
            if (p.prof) { const long long t = clock64(); mbar_wait(afull_bar(aslot), aphase); prof_full += clock64() - t; }
            else mbar_wait(afull_bar(aslot), aphase);
            const uint32_t h_addr = halo_base + aslot * MT * Cfg::kHaloBytes;
            const int ri = sl - nslices;  // >= 0: residual slice index (dense 8 x 16 box, identity weight tile)
            const int ntap = ri < 0 ? 9 : 1;
            const uint32_t id = ri < 0 ? idesc : umma_idesc_f16(min(64, BLOCK_N - ri * 64));
            const uint32_t dcol = ri < 0 ? 0u : (uint32_t)(ri * 64);
            for (int tap = 0; tap < ntap; ++tap) {
              if (p.prof) { const long long t = clock64(); mbar_wait(full_bar(stage), phase); prof_full += clock64() - t; }
              else mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint64_t bdesc = umma_desc_k128(smem_base + stage * Cfg::kStageBytes);
              const uint32_t woff = ri < 0 ? (uint32_t)((tap / 3) * 10 + tap % 3) * 128u : 0u;
              if (leader) {
#pragma unroll
                for (int u = 0; u < MT; ++u) {
                  const uint64_t adesc = ri < 0 ? umma_desc_k128_sbo(h_addr + u * Cfg::kHaloBytes + woff, 1280) : umma_desc_k128(h_addr + u * Cfg::kHaloBytes);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16(d_tmem + u * Cfg::kSubStride + dcol, adesc + 2 * k, bdesc + 2 * k, id, (sl | tap | k) != 0);
                }
                umma_commit(empty_bar(stage));
              }
              __syncwarp();
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            if (leader) umma_commit(aempty_bar(aslot));
            __syncwarp();
            if (++aslot == Cfg::kASlots) { aslot = 0; aphase ^= 1u; }
          }
        } else
        for (int ks = 0; ks < num_ksteps + nres; ++ks) {
          if (p.prof) { const long long t = clock64(); mbar_wait(full_bar(stage), phase); prof_full += clock64() - t; }
          else mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          const uint64_t bdesc = umma_desc_k128(a_addr + MT * Cfg::kABytes);
          const int ri = ks - num_ksteps;  // >= 0: residual slice index
          const uint32_t id = ri < 0 ? idesc : umma_idesc_f16(min(64, BLOCK_N - ri * 64));
          const uint32_t dcol = ri < 0 ? 0u : (uint32_t)(ri * 64);
          if (leader) {
#pragma unroll
            for (int u = 0; u < MT; ++u) {
              const uint64_t adesc = umma_desc_k128(a_addr + u * Cfg::kABytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // +32 B per 16-element K step (start-address field is in 16-byte units)
                umma_f16(d_tmem + u * Cfg::kSubStride + dcol, adesc + 2 * k, bdesc + 2 * k, id, (ks | k) != 0);
              }
            }
            umma_commit(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (leader) umma_commit(tfull_bar(acc));
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      if (p.prof && blockIdx.x < 2 && leader)
        printf("sdm prof: cta %d MMA issuer total %lld clk, waiting for operands %lld, for a free accumulator %lld\n", blockIdx.x,
               clock64() - prof_t0, prof_full, prof_tempty);
    }
  } else {
    // ============================== epilogue (4 warps, one TMEM lane quadrant each) ==============
    // Thread `lane` of quadrant `quad` owns accumulator row quad*32+lane (TMEM lane).  Global stores are made
    // row-contiguous through a per-warp 32 x 128 B staging tile in shared memory (XOR-swizzled 16-byte pieces):
    // a warp-wide 16-byte store then covers 4 rows x 128 contiguous bytes (4 L2 lines) instead of 32 rows x 16 bytes.
    const int quad = warp & 3;
    const int ewg = (warp - 2) >> 2;   // epilogue warpgroup of this warp (0 or 1)
    const int row = quad * 32 + lane;  // row of the 128-row tile == TMEM lane
    uint8_t* epi_base = smem_raw + (bar_base - smem_u32(smem_raw)) + 256;
    uint8_t* stg = epi_base + (warp - 2) * 4096;
    float* sred_base = reinterpret_cast<float*>(epi_base + EWG * 4 * 4096 + ewg * 4096);
    int sred_sel = 0;  // the partials buffer alternates per slab: one named barrier per slab instead of two
    const int t_row0 = lane >> 3, t_piece = lane & 7;
    int ltw = 0;
    while ((1 << ltw) < p.tw) ++ltw;
    // tile -> warpgroup assignment: MT=1 with two warpgroups alternates tiles (warpgroup e drains accumulator stage e);
    // otherwise every warpgroup sees every tile (and, for MT=2, owns M sub-tile `ewg`)
    constexpr bool kAlternate = (EWG == 2 && MT == 1);
    int acc = kAlternate ? ewg : 0;
    uint32_t acc_phase = 0;
    long long prof_tfull = 0;
    const long long prof_t0 = p.prof ? clock64() : 0;
    for (int tile = tile_first + (kAlternate ? ewg * tile_step : 0); tile < p.total_tiles;
         tile += (kAlternate ? 2 : 1) * tile_step) {
      if (p.prof) { const long long t = clock64(); mbar_wait(tfull_bar(acc), acc_phase); prof_tfull += clock64() - t; }
      else mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int nt = tile % p.n_tiles;
      const int n0 = nt * BLOCK_N;
#pragma unroll 1
      for (int u = (EWG == 2 && MT == 2) ? ewg : 0; u < ((EWG == 2 && MT == 2) ? ewg + 1 : MT); ++u) {
        const int mt = (tile / p.n_tiles) * SUBS + u;
        if (mt >= p.m_tiles) break;  // warp-uniform
        const int tx = mt % p.tiles_x;
        const int ty = (mt / p.tiles_x) % p.tiles_y;
        const int b = mt / (p.tiles_x * p.tiles_y);
        const float* bias = p.bias ? p.bias + (p.bias_sel ? (long long)p.bias_sel[b] * p.N : 0) : nullptr;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * Cfg::kAccStride + u * Cfg::kSubStride;

        if constexpr (MODE == EPI_F16) {
          // ---- fp16 output, staged row-contiguous stores (+ GroupNorm partials) ---------------------------------
          // r1l ncu (source view): the epilogue warps of the N=128 / short-K GEMMs were busy 85 % of the time at an IPC of
          // 0.1-0.3 (~950 instructions per 64-column slab, every tcgen05.ld and bias load waited for in place) and were the
          // critical resource (tensor pipe 59 % on the 128->128 3x3 convs, 4 % on K=64 GEMMs).  This version keeps the next
          // 32 columns' tcgen05.ld in flight while the current ones are processed, issues the bias loads before the wait,
          // does scale+bias as one FFMA2 per column pair, the statistics as FADD2/FFMA2, and hoists the 64-bit address math.
          constexpr int SLAB = 64;                            // output columns per 128-byte staging row
          const int o_lim = min(p.N, n0 + BLOCK_N);
          __half* obase = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride;
          // element offsets (within this sample: < 2^31) of the 8 rows this lane serves in the transposed phase; -1 = outside
          int toff[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rg = quad * 32 + i * 4 + t_row0;
            const int xx = tx * p.tw + (rg & (p.tw - 1)), yy = ty * p.th + (rg >> ltw);
            const int pix = UPS2 ? (2 * yy) * (2 * p.W) + 2 * xx : yy * p.W + xx;
            toff[i] = (xx < p.W && yy < p.H) ? pix * (int)p.out_ld : -1;
          }
          const uint64_t sc2 = pack_f2(p.scale, p.scale);
          // scale * acc + bias for 32 accumulator columns -> this thread's staging row, 16-byte pieces q0 .. q0+3
          auto half_tile = [&](const uint32_t (&r)[32], const float4 (&bq)[8], int q0) {
            uint32_t w[16];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const uint64_t lo = fma_f2(pack_f2(__uint_as_float(r[g * 4 + 0]), __uint_as_float(r[g * 4 + 1])), sc2, pack_f2(bq[g].x, bq[g].y));
              const uint64_t hi = fma_f2(pack_f2(__uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3])), sc2, pack_f2(bq[g].z, bq[g].w));
              float v0, v1, v2, v3;
              unpack_f2(lo, v0, v1);
              unpack_f2(hi, v2, v3);
              w[g * 2] = pack_h2(v0, v1);
              w[g * 2 + 1] = pack_h2(v2, v3);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(stg + lane * 128 + (((q0 + g) ^ (lane & 7)) << 4)) = make_uint4(w[g * 4], w[g * 4 + 1], w[g * 4 + 2], w[g * 4 + 3]);
          };
          auto load_bias = [&](int cc, float4 (&bq)[8]) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              bq[g] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bias && n0 + cc + g * 4 < p.N) bq[g] = __ldg(reinterpret_cast<const float4*>(bias + n0 + cc + g * 4));
            }
          };
          uint32_t ra[32], rb[32];
          __syncwarp();
          tmem_ld32(taddr, ra);
#pragma unroll 1
          for (int c = 0; c < BLOCK_N; c += SLAB) {
            const bool second = (BLOCK_N % SLAB == 0) || (c + 32 < BLOCK_N);  // the last slab of a 160-wide tile has one half
            float4 bq[8];
            load_bias(c, bq);
            tmem_ld_wait();                                   // ra = columns [c, c+32)
            __syncwarp();                                     // (also: the previous slab's staging reads are done)
            if (second) tmem_ld32(taddr + c + 32, rb);        // in flight while ra is processed
            half_tile(ra, bq, 0);
            if (second) {
              load_bias(c + 32, bq);
              tmem_ld_wait();                                 // rb = columns [c+32, c+64)
              __syncwarp();
              if (c + SLAB < BLOCK_N) tmem_ld32(taddr + c + SLAB, ra);  // next slab's first half flies during the stores
              half_tile(rb, bq, 4);
            } else {
#pragma unroll
              for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + g) ^ (lane & 7)) << 4)) = make_uint4(0, 0, 0, 0);
            }
            // staging -> row-contiguous global stores
            __syncwarp();
            const int col = n0 + c + t_piece * 8;
            const bool col_ok = col < o_lim;
            const bool do_stats = (!UPS2) && (p.stats != nullptr);
            __half* ocol = obase + col;
            uint64_t ssum[4], ssq[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { ssum[e] = 0ull; ssq[e] = 0ull; }
            // one staged row -> global (row-contiguous: a warp store covers 4 rows x 128 B); WITH_STATS adds the row to the
            // per-column (sum, sum of squares) of the STORED fp16 values.  Rows/columns outside the tensor are zeroed instead
            // of branched around, so the accumulators stay in place (the branchy form cost 16 register moves per row).
            auto store_rows = [&](auto with_stats) {
              constexpr bool WITH_STATS = decltype(with_stats)::value;
              (void)WITH_STATS;  // unused in the UPS2 instantiations
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rl = i * 4 + t_row0;
                uint4 v = *reinterpret_cast<const uint4*>(stg + rl * 128 + ((t_piece ^ (rl & 7)) << 4));
                const bool ok = toff[i] >= 0 && col_ok;
                if constexpr (!UPS2) {
                  if (ok) *reinterpret_cast<uint4*>(ocol + toff[i]) = v;
                  if constexpr (WITH_STATS) {
                    if (!ok) v = make_uint4(0, 0, 0, 0);
                    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float2 f = __half22float2(h[j]);
                      const uint64_t f2 = pack_f2(f.x, f.y);
                      ssum[j] = add_f2(ssum[j], f2);
                      ssq[j] = fma_f2(f2, f2, ssq[j]);
                    }
                  }
                } else if (ok) {
                  // nearest-neighbour 2x upsample fused into the store (reference Upsample2D: F.interpolate scale 2
                  // "nearest" followed by a conv; the conv then reads this tensor)
#pragma unroll
                  for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(ocol + toff[i] + ((q >> 1) * (2 * p.W) + (q & 1)) * (int)p.out_ld) = v;
                }
              }
            };
            if (do_stats) store_rows(std::true_type{});
            else store_rows(std::false_type{});
            if constexpr (!UPS2) {
              if (do_stats) {
                // column sums over this warp's 32 rows: lanes {l, l+8, l+16, l+24} hold the same 8 columns
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  ssum[e] = add_f2(ssum[e], __shfl_xor_sync(0xffffffffu, ssum[e], 8));
                  ssq[e] = add_f2(ssq[e], __shfl_xor_sync(0xffffffffu, ssq[e], 8));
                  ssum[e] = add_f2(ssum[e], __shfl_xor_sync(0xffffffffu, ssum[e], 16));
                  ssq[e] = add_f2(ssq[e], __shfl_xor_sync(0xffffffffu, ssq[e], 16));
                }
                // combine the four epilogue warps in a fixed order through shared memory (one global writer per slot).
                // Two buffers: warp 0 of the warpgroup reads buffer k while the others may already fill buffer k^1 for the
                // next slab; nobody can reach buffer k again before passing the next slab's barrier, i.e. after warp 0 left it.
                float* sred = sred_base + sred_sel * 512;
                sred_sel ^= 1;
                if (lane < 8) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    float s0, s1, q0, q1;
                    unpack_f2(ssum[e], s0, s1);
                    unpack_f2(ssq[e], q0, q1);
                    *reinterpret_cast<float4*>(&sred[(quad * 64 + lane * 8 + e * 2) * 2]) = make_float4(s0, q0, s1, q1);
                  }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + ewg) : "memory");
                if (quad == 0) {
                  const int cc = lane * 2;  // two columns per lane
                  const int gcol = n0 + c + cc;
                  if (gcol < o_lim) {
                    float4 o;
                    o.x = (sred[(0 * 64 + cc) * 2] + sred[(1 * 64 + cc) * 2]) + (sred[(2 * 64 + cc) * 2] + sred[(3 * 64 + cc) * 2]);
                    o.y = (sred[(0 * 64 + cc) * 2 + 1] + sred[(1 * 64 + cc) * 2 + 1]) + (sred[(2 * 64 + cc) * 2 + 1] + sred[(3 * 64 + cc) * 2 + 1]);
                    o.z = (sred[(0 * 64 + cc + 1) * 2] + sred[(1 * 64 + cc + 1) * 2]) + (sred[(2 * 64 + cc + 1) * 2] + sred[(3 * 64 + cc + 1) * 2]);
                    o.w = (sred[(0 * 64 + cc + 1) * 2 + 1] + sred[(1 * 64 + cc + 1) * 2 + 1]) + (sred[(2 * 64 + cc + 1) * 2 + 1] + sred[(3 * 64 + cc + 1) * 2 + 1]);
                    const long long slot = (long long)b * (p.tiles_x * p.tiles_y) + (mt % (p.tiles_x * p.tiles_y));
                    *reinterpret_cast<float4*>(p.stats + (slot * p.N + gcol) * 2) = o;
                  }
                }
              }
            }
          }
        } else if constexpr (MODE == EPI_GEGLU || MODE == EPI_F32) {
          // ---- staged, row-contiguous store path for the fp32 scores (VAE attention) and the GEGLU projection ------------
          constexpr int ES = (MODE == EPI_F32) ? 4 : 2;       // output element size
          constexpr int CPP = 16 / ES;                        // columns per 16-byte piece
          constexpr int SLAB = 128 / ES;                      // output columns per 128-byte staging row
          constexpr int OUT_COLS = (MODE == EPI_GEGLU) ? BLOCK_N / 2 : BLOCK_N;
          constexpr int HALF = BLOCK_N / 2;                   // GEGLU: tile = [HALF value columns | HALF gate columns]
          const int ocol0 = (MODE == EPI_GEGLU) ? (n0 >> 1) : n0;
          const int o_lim = (MODE == EPI_GEGLU) ? min(p.N >> 1, ocol0 + OUT_COLS) : min(p.N, n0 + BLOCK_N);
          uint8_t* obase = reinterpret_cast<uint8_t*>(p.out) + (long long)b * p.out_bstride * ES;
          // element offsets (within this sample) of the 8 rows this lane serves in the transposed phase; -1 = outside
          long long toff[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rg = quad * 32 + i * 4 + t_row0;
            const int xx = tx * p.tw + (rg & (p.tw - 1)), yy = ty * p.th + (rg >> ltw);
            toff[i] = (xx < p.W && yy < p.H) ? ((long long)yy * p.W + xx) * p.out_ld : -1;
          }
          // the first tcgen05.ld of the next slab is issued before this slab's stores (ra / rg hold 32 accumulator columns)
          uint32_t ra[32], rg[32];
          __syncwarp();
          tmem_ld32(taddr, ra);
          if constexpr (MODE == EPI_GEGLU) tmem_ld32(taddr + HALF, rg);
#pragma unroll 1
          for (int c = 0; c < OUT_COLS; c += SLAB) {
            if constexpr (MODE == EPI_F32) {
              // 32 fp32 columns per slab
              float bv[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) bv[i] = (bias && n0 + c + i < p.N) ? __ldg(bias + n0 + c + i) : 0.f;
              tmem_ld_wait();
              __syncwarp();
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = fmaf(__uint_as_float(ra[g * 4 + e]), p.scale, bv[g * 4 + e]);
                *reinterpret_cast<uint4*>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) =
                    make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
              }
              if (c + SLAB < OUT_COLS) tmem_ld32(taddr + c + SLAB, ra);
            } else {
              // 64 output columns per slab = two 32-column halves of value and gate
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                float bv[32], bg[32];
                if (bias) {
#pragma unroll
                  for (int g4 = 0; g4 < 8; ++g4) {
                    const float4 x4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + hh * 32 + g4 * 4));
                    const float4 y4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + HALF + c + hh * 32 + g4 * 4));
                    bv[g4 * 4] = x4.x; bv[g4 * 4 + 1] = x4.y; bv[g4 * 4 + 2] = x4.z; bv[g4 * 4 + 3] = x4.w;
                    bg[g4 * 4] = y4.x; bg[g4 * 4 + 1] = y4.y; bg[g4 * 4 + 2] = y4.z; bg[g4 * 4 + 3] = y4.w;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) { bv[i] = 0.f; bg[i] = 0.f; }
                }
                tmem_ld_wait();
                __syncwarp();
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  float o[2];
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    // reference rounding points: proj output fp16, gelu(gate) fp16, product fp16
                    const float v = __half2float(__float2half_rn(__uint_as_float(ra[2 * i + e]) + bv[2 * i + e]));
                    const float gt = __half2float(__float2half_rn(__uint_as_float(rg[2 * i + e]) + bg[2 * i + e]));
                    float ge = 0.5f * gt * (1.0f + erf_fast(gt * 0.70710678118654752f));
                    ge = __half2float(__float2half_rn(ge));
                    o[e] = v * ge;
                  }
                  w[i] = pack_h2(o[0], o[1]);
                }
                // the next 32 value / gate columns fly while this half is staged and (after the second half) stored
                const int nc = c + hh * 32 + 32;
                if (nc < OUT_COLS) {
                  tmem_ld32(taddr + nc, ra);
                  tmem_ld32(taddr + HALF + nc, rg);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
                  *reinterpret_cast<uint4*>(stg + lane * 128 + (((hh * 4 + g) ^ (lane & 7)) << 4)) = make_uint4(w[g * 4], w[g * 4 + 1], w[g * 4 + 2], w[g * 4 + 3]);
              }
            }
            // staging -> row-contiguous global stores
            __syncwarp();
            const int col = ocol0 + c + t_piece * CPP;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = i * 4 + t_row0;
              const uint4 v = *reinterpret_cast<const uint4*>(stg + rl * 128 + ((t_piece ^ (rl & 7)) << 4));
              if (toff[i] >= 0 && col < o_lim) *reinterpret_cast<uint4*>(obase + (toff[i] + col) * ES) = v;
            }
            __syncwarp();  // staging is rewritten by the next slab
          }
        } else {
          // ---- row-per-thread paths (transposed store is already coalesced along pixels; skinny outputs are tiny)
          const int x = tx * p.tw + (row & (p.tw - 1));
          const int y = ty * p.th + (row >> ltw);
          const bool valid = (x < p.W) && (y < p.H);
          const long long pix = (long long)y * p.W + x;
          if constexpr (MODE == EPI_F16_T) {
            // 32-column slabs, the next slab's tcgen05.ld and this slab's bias values in flight while the current one is stored
            // (r3c ncu of the r2z form — ld, wait, then per element a bias load, a bounds branch and the store —: issue slots
            // 38 % busy, top stalls long_scoreboard / wait, tensor pipe 21 %)
            __half* out = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride + pix;
            uint32_t ra[32], rb[32];
            auto slab = [&](const uint32_t (&r)[32], uint32_t (&nxt)[32], int c) {
              float bv[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) bv[i] = (bias && n0 + c + i < p.N) ? __ldg(bias + n0 + c + i) : 0.f;
              tmem_ld_wait();
              __syncwarp();  // tcgen05.ld is .sync.aligned: re-converge after the (divergent) store code
              if (c + 32 < BLOCK_N) tmem_ld32(taddr + c + 32, nxt);
              if (valid) {
                __half* o = out + (long long)(n0 + c) * p.out_ld;
                if (n0 + c + 32 <= p.N) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) o[(long long)i * p.out_ld] = __float2half_rn(fmaf(__uint_as_float(r[i]), p.scale, bv[i]));
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (n0 + c + i < p.N) o[(long long)i * p.out_ld] = __float2half_rn(fmaf(__uint_as_float(r[i]), p.scale, bv[i]));
                }
              }
            };
            __syncwarp();
            tmem_ld32(taddr, ra);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 64) {
              slab(ra, rb, c);
              if (c + 32 < BLOCK_N) slab(rb, ra, c + 32);
            }
          } else {
            uint32_t r[32];
            __syncwarp();
            tmem_ld32(taddr, r);
            tmem_ld_wait();
            if (valid) {
              if constexpr (MODE == EPI_ALPHA) {
                // rounding points of the reference fp16 path: conv outputs fp16, channel mean fp16, (clip+1) fp16, /2 exact
                const float c0 = __half2float(__float2half_rn(__uint_as_float(r[0]) + bias[0]));
                const float c1 = __half2float(__float2half_rn(__uint_as_float(r[1]) + bias[1]));
                const float c2 = __half2float(__float2half_rn(__uint_as_float(r[2]) + bias[2]));
                const __half m = __float2half_rn((c0 + c1 + c2) / 3.0f);
                const long long o = (long long)b * p.out_bstride + pix;
                if (p.out2) reinterpret_cast<__half*>(p.out2)[o] = m;
                const float cl = fminf(fmaxf(__half2float(m), -1.0f), 1.0f);
                const __half p1 = __float2half_rn(cl + 1.0f);
                reinterpret_cast<__half*>(p.out)[o] = __float2half_rn(__half2float(p1) * 0.5f);
              } else {  // EPI_SKINNY
                __half* out = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride + pix * p.out_ld;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if (i < p.n_store) {
                    float v = __uint_as_float(r[i]) * p.scale + (bias ? bias[i] : 0.f);
                    if (p.post_div != 1.0f) v = __half2float(__float2half_rn(v)) / p.post_div;
                    out[i] = __float2half_rn(v);
                  }
                }
              }
            }
          }
        }
      }  // M sub-tile
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (kAlternate) acc_phase ^= 1u;
      else if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (p.prof && blockIdx.x < 2 && (warp == 2 || warp == 6) && lane == 0)
      printf("sdm prof: cta %d epilogue warp %d total %lld clk, waiting for an accumulator %lld\n", blockIdx.x, warp, clock64() - prof_t0, prof_tfull);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace sdm
