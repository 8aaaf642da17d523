// Implicit-GEMM convolution / linear kernel on tcgen05 (sm_100a).
//
// One kernel covers every dense contraction of the matte path except attention:
//   * Conv2d 3x3 stride 1 (pad 1), stride 2 (UNet pad 1 / VAE asymmetric pad), 1x1 shortcut convs
//     (reference: diffusers ResnetBlock2D / Downsample2D / Upsample2D built at
//      /root/reference/src/utils/replace.py:239,268,321)
//   * nn.Linear (proj_in/out, to_q/k/v/out, GEGLU, FF-out; reference Attention/FeedForward)
//   * the VAE mid-block attention GEMMs (QK^T -> fp32 scores, P·V)
//
// A operand ("activation"): NHWC fp16 tensor(s) read through up-to-4 TMA tensor maps (C, W, H, B).
//   A tile of 128 output pixels is a (tw x th) patch of one image; for tap (dy,dx) the producer
//   issues ONE 4-D TMA box load at (c0, x0+dx, y0+dy, b): out-of-image coordinates are zero-filled
//   by the TMA unit, which *is* the convolution padding.  Box rows land in shared memory as
//   128 rows x 128 B with the 128-byte swizzle, exactly the K-major layout tcgen05.mma expects.
// B operand ("weight"): [N][Ktot] fp16, K contiguous, Ktot ordered (tap, cin); 2-D map (or 3-D batched).
// D: 128 x BLOCK_N fp32 accumulator in TMEM, double buffered so the epilogue of tile i overlaps
//   the main loop of tile i+1.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM
//   alloc), warps 2..5 = epilogue (TMEM -> registers -> bias/residual/activation -> global).
#pragma once
#include "common.cuh"

namespace sdm {

enum EpiMode : int {
  EPI_F16 = 0,    // out[pixel][n] fp16  (+bias, +residual, optional 2x nearest-upsample scatter)
  EPI_F16_T = 1,  // out[b][n][pixel] fp16 (transposed; used for V^T)
  EPI_GEGLU = 2,  // out[pixel][n/2] = fp16(v) * gelu(fp16(g)), tile = [BLOCK_N/2 value | BLOCK_N/2 gate]
  EPI_F32 = 3,    // out[pixel][n] fp32 = scale * acc
  EPI_ALPHA = 4,  // N>=3: alpha[pixel] = (clip(mean(fp16(c0..c2)), -1, 1) + 1) / 2 ; out2[pixel] = pre-clip mean (meta_arch.py:258-260)
};

struct alignas(64) ConvGemmParams {
  CUtensorMap a_map[4];
  CUtensorMap b_map;
  int B, H, W;        // output pixel grid (per batch element)
  int tw, th;         // tile patch, tw*th == 128
  int tiles_x, tiles_y;
  int N;              // GEMM N (output channels before GEGLU halving)
  int n_tiles;
  int total_tiles;
  int ntaps, nsrc;
  int src_c[2];       // channels per A source (multiples of 64)
  int cin_total;      // weight K per tap
  signed char tap_map[9], tap_dx[9], tap_dy[9];
  int b_batched;      // weight map has a batch coordinate
  int mode;
  int ups2;           // EPI_F16 only: write each pixel to the 2x2 block of a (2H,2W) output
  void* out;
  long long out_ld;       // elements between consecutive pixels (EPI_F16/F32/GEGLU) or row length (EPI_F16_T)
  long long out_bstride;  // elements per batch element
  const float* bias;      // [nsel][N] fp32 or null
  const int* bias_sel;    // per-batch row selector into bias, or null
  const __half* res;      // residual, same indexing as out (never with ups2)
  long long res_ld, res_bstride;
  float scale;
  float post_div;   // EPI_F16: fp16(result) / post_div, rounded again (label_latent / scaling_factor); 1 = off
  int n_store;      // EPI_F16: number of output columns actually stored (< 8 => scalar stores)
  void* out2;       // EPI_ALPHA: optional pre-clip mean
};

template <int BLOCK_N>
struct ConvGemmCfg {
  static constexpr int kABytes = 128 * 128;          // 128 rows x 64 fp16
  static constexpr int kBBytes = BLOCK_N * 128;      // BLOCK_N rows x 64 fp16
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
  static constexpr int kAccStride = BLOCK_N < 32 ? 32 : BLOCK_N;  // TMEM columns between the two accumulators
  static constexpr int kTmemCols = (2 * kAccStride <= 64) ? 64 : (2 * kAccStride <= 128) ? 128 : (2 * kAccStride <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kThreads = 192;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using Cfg = ConvGemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tfull[2], tempty[2], then tmem ptr (4 B)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
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
      mbar_init(tempty_bar(a), 4);
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
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        const int mt = tile / p.n_tiles;
        const int tx = mt % p.tiles_x;
        const int ty = (mt / p.tiles_x) % p.tiles_y;
        const int b = mt / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.tw, y0 = ty * p.th, n0 = nt * BLOCK_N;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          int koff = tap * p.cin_total;
          for (int s = 0; s < p.nsrc; ++s) {
            const CUtensorMap* amap = &p.a_map[p.tap_map[tap] + s];
            for (int c0 = 0; c0 < p.src_c[s]; c0 += 64) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
              const uint32_t b_dst = a_dst + Cfg::kABytes;
              mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
              tma_load_4d(a_dst, amap, full_bar(stage), c0, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], b);
              if (p.b_batched)
                tma_load_3d(b_dst, &p.b_map, full_bar(stage), koff + c0, n0, b);
              else
                tma_load_2d(b_dst, &p.b_map, full_bar(stage), koff + c0, n0);
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            koff += p.src_c[s];
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        for (int ks = 0; ks < num_ksteps; ++ks) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = umma_desc_k128(a_addr);
          const uint64_t bdesc = umma_desc_k128(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // +32 B per 16-element K step (start-address field is in 16-byte units)
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | k) != 0);
          }
          umma_commit(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue (4 warps, one TMEM lane quadrant each) ==============
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // row of the 128-row tile == TMEM lane
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int tx = mt % p.tiles_x;
      const int ty = (mt / p.tiles_x) % p.tiles_y;
      const int b = mt / (p.tiles_x * p.tiles_y);
      const int x = tx * p.tw + (row % p.tw);
      const int y = ty * p.th + (row / p.tw);
      const bool valid = (x < p.W) && (y < p.H);
      const int n0 = nt * BLOCK_N;
      const long long pix = (long long)y * p.W + x;
      const float* bias = p.bias ? p.bias + (p.bias_sel ? (long long)p.bias_sel[b] * p.N : 0) : nullptr;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * Cfg::kAccStride;

      if (p.mode == EPI_GEGLU) {
        constexpr int HALF = BLOCK_N / 2;
        __half* out = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride + pix * p.out_ld + (n0 >> 1);
#pragma unroll 1
        for (int c = 0; c < HALF; c += 32) {
          uint32_t rv[32], rg[32];
          __syncwarp();
          tmem_ld32(taddr + c, rv);
          tmem_ld32(taddr + HALF + c, rg);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (n0 + c + g * 8 < p.N) {  // N counts value+gate columns; tiles never straddle
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float o[2];
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    const int i = g * 8 + j * 2 + e;
                    float v = __uint_as_float(rv[i]);
                    float gt = __uint_as_float(rg[i]);
                    if (bias) {
                      v += bias[n0 + c + i];
                      gt += bias[n0 + HALF + c + i];
                    }
                    // reference rounding points: proj output fp16, gelu(gate) fp16, product fp16
                    v = __half2float(__float2half_rn(v));
                    gt = __half2float(__float2half_rn(gt));
                    float ge = 0.5f * gt * (1.0f + erff(gt * 0.70710678118654752f));
                    ge = __half2float(__float2half_rn(ge));
                    o[e] = v * ge;
                  }
                  w[j] = pack_h2(o[0], o[1]);
                }
                *reinterpret_cast<uint4*>(out + c + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
        }
      } else if (p.mode == EPI_ALPHA) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(taddr, r);
        tmem_ld_wait();
        if (valid) {
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
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t r[32];
          __syncwarp();  // tcgen05.ld is .sync.aligned: re-converge after the (divergent) store code
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
          if (valid && (n0 + c < p.N)) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * p.scale;
          if (bias) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (n0 + c + g * 4 < p.N) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + g * 4));
                v[g * 4 + 0] += bb.x; v[g * 4 + 1] += bb.y; v[g * 4 + 2] += bb.z; v[g * 4 + 3] += bb.w;
              }
            }
          }
          if (p.mode == EPI_F32) {
            float* out = reinterpret_cast<float*>(p.out) + (long long)b * p.out_bstride + pix * p.out_ld + n0 + c;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (n0 + c + g * 4 < p.N)
                *reinterpret_cast<float4*>(out + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
          } else if (p.mode == EPI_F16_T) {
            __half* out = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride + pix;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + c + i < p.N) out[(long long)(n0 + c + i) * p.out_ld] = __float2half_rn(v[i]);
          } else {
            if (p.res) {
              const __half* res = p.res + (long long)b * p.res_bstride + pix * p.res_ld + n0 + c;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (n0 + c + g * 8 < p.N) {
                  const uint4 rr = *reinterpret_cast<const uint4*>(res + g * 8);
                  const __half2* h = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    // reference rounds the producer output to fp16 before the residual add
                    const float2 f = __half22float2(h[j]);
                    v[g * 8 + j * 2] = __half2float(__float2half_rn(v[g * 8 + j * 2])) + f.x;
                    v[g * 8 + j * 2 + 1] = __half2float(__float2half_rn(v[g * 8 + j * 2 + 1])) + f.y;
                  }
                }
              }
            }
            if (p.post_div != 1.0f) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __half2float(__float2half_rn(v[i])) / p.post_div;
            }
            uint4 w[4];
#pragma unroll
            for (int g = 0; g < 4; ++g)
              w[g] = make_uint4(pack_h2(v[g * 8], v[g * 8 + 1]), pack_h2(v[g * 8 + 2], v[g * 8 + 3]),
                                pack_h2(v[g * 8 + 4], v[g * 8 + 5]), pack_h2(v[g * 8 + 6], v[g * 8 + 7]));
            if (p.n_store < 8) {
              __half* out = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride + pix * p.out_ld;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < p.n_store) out[i] = __float2half_rn(v[i]);
            } else if (!p.ups2) {
              __half* out = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride + pix * p.out_ld + n0 + c;
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (n0 + c + g * 8 < p.N) *reinterpret_cast<uint4*>(out + g * 8) = w[g];
            } else {
              // nearest-neighbour 2x upsample fused into the store (reference Upsample2D: F.interpolate
              // scale 2 "nearest" followed by a conv; the conv then reads this tensor)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const long long pix2 = (long long)(2 * y + (q >> 1)) * (2 * p.W) + (2 * x + (q & 1));
                __half* out = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride + pix2 * p.out_ld + n0 + c;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                  if (n0 + c + g * 8 < p.N) *reinterpret_cast<uint4*>(out + g * 8) = w[g];
              }
            }
          }
          }  // valid
        }
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
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace sdm
