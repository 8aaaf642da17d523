// Fused softmax(scale * Q K^T + key_bias) V for head dim 64 on tcgen05 (sm_100a).
//
// Replaces, for every UNet attention (self: attn1 with the trimap key bias, cross: attn2 over the 16 384
// trimap tokens), the reference chain
//   custom_prepare_attention_mask  /root/reference/src/utils/replace.py:20-72
//   custom_get_attention_scores    /root/reference/src/utils/replace.py:75-122  (baddbmm + softmax)
//   torch.bmm(probs, value) inside diffusers AttnProcessor / SlicedAttnProcessor (sdmatte_nodes.py:331-335)
// without ever materialising the L x L score matrix: S lives in TMEM, P in shared memory, O in TMEM.
//
// CTA = 256 queries (two 128-row tiles A and B) of one (batch, head); it streams 64-key tiles of K and V^T.
//   warps 0-3  : softmax warpgroup for tile A (thread r owns query row r == TMEM lane r)
//   warps 4-7  : softmax warpgroup for tile B
//   warp  8    : TMA producer (Q once; K, V^T and the per-key bias per tile; 6-stage ring)
//   warp  9    : tcgen05.mma issuer (S_X = Q_X K^T : M128 N64 K64 ; O_X += P_X V : M128 N64 K64) + TMEM alloc
// S_X and P_X are DOUBLE buffered: the tensor core computes S_X(j+1)/S_X(j+2) while the warpgroup exponentiates S_X(j),
// and P·V(j-1) still reads P_X[(j-1)&1] while P_X[j&1] is written, so a warpgroup never waits for the MMA it just
// triggered (r1f ncu: with single buffers ~20 % of the softmax warps' samples sat in barrier spin loops, MUFU 54 %).
// O_X accumulates in TMEM over all key tiles; the softmax reference max is raised lazily (only when a score exceeds it
// by more than 2^8), in which case the warp rescales its O rows in TMEM (tcgen05.ld / st).
// TMEM columns: S_A[0] 0, S_A[1] 64, S_B[0] 128, S_B[1] 192, O_A 256, O_B 320.
// The per-key bias is expected pre-multiplied by log2(e); scores are handled in the log2 domain, statistics in fp32,
// and scores are NOT rounded to fp16 before the softmax (the reference does).
#include "common.cuh"
#include "kernels.h"
#include "tmap.h"

#include <cstdlib>
#include <mutex>

namespace sdm {

struct alignas(64) AttnParams {
  CUtensorMap q_map, k_map, vt_map;
  const float* bias;
  long long bias_bstride;
  __half* out;
  long long ldo;
  int Lq, Lk, heads, n_ktiles;
  float scale;
};

constexpr int kAttnStages = 6;
constexpr int kAttnThreads = 320;
constexpr int kKeys = 64;                      // keys per tile
constexpr uint32_t kQBytes = 128 * 128;        // one 128 x 64 fp16 tile
constexpr uint32_t kPBytes = 128 * 128;        // 128 rows x 64 keys fp16
constexpr uint32_t kKBytes = kKeys * 128;      // 64 keys x 64 d
constexpr uint32_t kVBytes = 64 * 128;         // 64 d x 64 keys
constexpr uint32_t kStageBytes = kKBytes + kVBytes;
constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffP = 2 * kQBytes;                 // P[x][buf] at kOffP + (x*2+buf)*kPBytes
constexpr uint32_t kOffStage = kOffP + 4 * kPBytes;
constexpr uint32_t kOffBias = kOffStage + kAttnStages * kStageBytes;
constexpr uint32_t kOffBar = kOffBias + kAttnStages * 256;
constexpr uint32_t kAttnSmem = kOffBar + 512 + 1024;

template <bool HAS_BIAS>
__global__ void __launch_bounds__(kAttnThreads, 1) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + kOffBar;
  // barriers (8 B each)
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kAttnStages + s); };
  auto s_full = [&](int x, int bf) { return bar + 8u * (1 + 2 * kAttnStages + x * 2 + bf); };
  auto p_full = [&](int x, int bf) { return bar + 8u * (5 + 2 * kAttnStages + x * 2 + bf); };
  auto pv_done = [&](int x, int bf) { return bar + 8u * (9 + 2 * kAttnStages + x * 2 + bf); };
  const uint32_t tmem_slot = bar + 8u * (13 + 2 * kAttnStages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kOffBar + 8 * (13 + 2 * kAttnStages));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n = p.n_ktiles;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kAttnStages; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int x = 0; x < 2; ++x)
      for (int bf = 0; bf < 2; ++bf) { mbar_init(s_full(x, bf), 1); mbar_init(p_full(x, bf), 128); mbar_init(pv_done(x, bf), 1); }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 8) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      tma_prefetch_desc(&p.q_map); tma_prefetch_desc(&p.k_map); tma_prefetch_desc(&p.vt_map);
      mbar_expect_tx(q_full, 2 * kQBytes);
      tma_load_3d(base + kOffQ, &p.q_map, q_full, h * 64, q0, b);
      tma_load_3d(base + kOffQ + kQBytes, &p.q_map, q_full, h * 64, q0 + 128, b);
      for (int j = 0; j < n; ++j) {
        const int s = j % kAttnStages;
        const uint32_t f = (uint32_t)(j / kAttnStages);
        mbar_wait(kv_empty(s), (f & 1u) ^ 1u);
        const uint32_t kdst = base + kOffStage + s * kStageBytes;
        mbar_expect_tx(kv_full(s), kStageBytes + (HAS_BIAS ? 256u : 0u));
        tma_load_3d(kdst, &p.k_map, kv_full(s), h * 64, j * kKeys, b);
        tma_load_3d(kdst + kKBytes, &p.vt_map, kv_full(s), j * kKeys, h * 64, b);
        if (HAS_BIAS) bulk_load_1d(base + kOffBias + s * 256, p.bias + (long long)b * p.bias_bstride + (long long)j * kKeys, 256, kv_full(s));
      }
    }
  } else if (warp == 9) {
    // ======================================= MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(64);
      auto issue_s = [&](int x, int j) {  // S_x[j&1] = Q_x K_j^T
        const uint64_t ad = umma_desc_k128(base + kOffQ + x * kQBytes);
        const uint64_t bd = umma_desc_k128(base + kOffStage + (j % kAttnStages) * kStageBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + x * 128 + (j & 1) * 64, ad + 2 * k, bd + 2 * k, idesc, k != 0);
        umma_commit(s_full(x, j & 1));
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < 2 && j < n; ++j) {
        mbar_wait(kv_full(j), 0);
        tc_fence_after();
        issue_s(0, j);
        issue_s(1, j);
      }
      for (int j = 0; j < n; ++j) {
        const int s = j % kAttnStages;
        if (j + 2 < n) mbar_wait(kv_full((j + 2) % kAttnStages), (uint32_t)((j + 2) / kAttnStages) & 1u);
        for (int x = 0; x < 2; ++x) {
          mbar_wait(p_full(x, j & 1), (uint32_t)(j >> 1) & 1u);  // P_x(j) in smem; S_x[j&1] consumed
          tc_fence_after();
          const uint64_t ad = umma_desc_k128(base + kOffP + (x * 2 + (j & 1)) * kPBytes);
          const uint64_t bd = umma_desc_k128(base + kOffStage + s * kStageBytes + kKBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem + 256 + x * 64, ad + 2 * k, bd + 2 * k, idesc, (j | k) != 0);
          umma_commit(pv_done(x, j & 1));
          if (j + 2 < n) issue_s(x, j + 2);
        }
        umma_commit(kv_empty(s));
      }
    }
  } else if (warp < 8) {
    // ======================================= softmax warpgroups =================================
    const int x = warp >> 2;                 // 0: tile A, 1: tile B
    const int r = (warp & 3) * 32 + lane;    // row within the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem + lane_base + x * 128;
    const uint32_t t_o = tmem + lane_base + 256 + x * 64;
    uint8_t* p_row = base_ptr + kOffP + x * 2 * kPBytes + (r >> 3) * 1024 + (r & 7) * 128;
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    const float sc = p.scale * 1.4426950408889634f;
    constexpr float kTau = 8.0f;
    float m2 = -INFINITY, l = 0.f;

    // S(j) is fetched from TMEM into registers one tile ahead (the tcgen05.ld latency hides behind the P stores,
    // fences and barrier traffic of the previous tile), so the exp/convert loop below never stalls on TMEM.
    uint32_t ra[32], rb[32];
    if (HAS_BIAS) mbar_wait(kv_full(0), 0);
    mbar_wait(s_full(x, 0), 0);
    tc_fence_after();
    tmem_ld32(t_s, ra);
    tmem_ld32(t_s + 32, rb);

    for (int j = 0; j < n; ++j) {
      const int s = j % kAttnStages;
      const int bf = j & 1;
      const bool tail = (!HAS_BIAS) && (j == n - 1) && ((p.Lk & (kKeys - 1)) != 0);
      const int kbase = j * kKeys;
      const float4* bias4 = reinterpret_cast<const float4*>(base_ptr + kOffBias + s * 256);
      tmem_ld_wait();
      if (!HAS_BIAS && tail) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (kbase + i >= p.Lk) ra[i] = 0xff800000u;  // -inf
          if (kbase + 32 + i >= p.Lk) rb[i] = 0xff800000u;
        }
      }
      uint32_t P[32];   // the 64 probabilities of this row, packed fp16x2
      float rowsum, mx;
      // one pass over the 64 scores held in registers: e = 2^(x - mref), row sum, row max of x
      auto pass = [&](float mref) {
        rowsum = 0.f;
        mx = -INFINITY;
        const float neg_m = -mref;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8], e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(c == 0 ? ra[q * 8 + i] : rb[q * 8 + i]);
            if (HAS_BIAS) {
              const float4 b0 = bias4[c * 8 + q * 2], b1 = bias4[c * 8 + q * 2 + 1];
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float xs[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) xs[i] = fmaf(v[i], sc, bb[i]);
              mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3])), fmaxf(fmaxf(xs[4], xs[5]), fmaxf(xs[6], xs[7]))));
#pragma unroll
              for (int i = 0; i < 8; ++i) e[i] = ex2f(xs[i] + neg_m);
            } else {
              const float cm = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
              mx = fmaxf(mx, cm * sc);
#pragma unroll
              for (int i = 0; i < 8; ++i) e[i] = ex2f(fmaf(v[i], sc, neg_m));
            }
            rowsum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
            P[(c * 4 + q) * 4 + 0] = pack_h2(e[0], e[1]);
            P[(c * 4 + q) * 4 + 1] = pack_h2(e[2], e[3]);
            P[(c * 4 + q) * 4 + 2] = pack_h2(e[4], e[5]);
            P[(c * 4 + q) * 4 + 3] = pack_h2(e[6], e[7]);
          }
        }
      };
      pass(m2);
      if (__any_sync(0xffffffffu, mx > m2 + kTau)) {
        // raise the reference: every P·V issued so far must have landed in O before the rows are rescaled
        const float m_new = fmaxf(m2, mx);
        const float alpha = ex2f(m2 - m_new);  // 0 on the first tile (m2 = -inf)
        if (j > 0) {
          mbar_wait(pv_done(x, (j - 1) & 1), (uint32_t)((j - 1) >> 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t oo[32];
            tmem_ld32(t_o + c * 32, oo);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) oo[i] = __float_as_uint(__uint_as_float(oo[i]) * alpha);
            tmem_st32(t_o + c * 32, oo);
          }
          tmem_st_wait();
        }
        l *= alpha;
        m2 = m_new;
        pass(m2);
      }
      l += rowsum;
      // prefetch S(j+1) into registers (its MMA was issued two tiles ago)
      if (j + 1 < n) {
        if (HAS_BIAS) mbar_wait(kv_full((j + 1) % kAttnStages), (uint32_t)((j + 1) / kAttnStages) & 1u);
        mbar_wait(s_full(x, bf ^ 1), (uint32_t)((j + 1) >> 1) & 1u);
        tc_fence_after();
        tmem_ld32(t_s + (bf ^ 1) * 64, ra);
        tmem_ld32(t_s + (bf ^ 1) * 64 + 32, rb);
      }
      // P_x[bf] was last read by P·V(j-2)
      mbar_wait(pv_done(x, bf), ((uint32_t)(j >> 1) & 1u) ^ 1u);
#pragma unroll
      for (int chunk = 0; chunk < 8; ++chunk)
        *reinterpret_cast<uint4*>(p_row + bf * kPBytes + ((uint32_t)(chunk << 4) ^ rx)) =
            make_uint4(P[chunk * 4], P[chunk * 4 + 1], P[chunk * 4 + 2], P[chunk * 4 + 3]);
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full(x, bf));
    }
    // ---- normalise and store
    mbar_wait(pv_done(x, (n - 1) & 1), (uint32_t)((n - 1) >> 1) & 1u);
    tc_fence_after();
    const int q = q0 + x * 128 + r;
    const float inv = 1.0f / l;
    __half* dst = p.out + ((long long)b * p.Lq + q) * p.ldo + h * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t oo[32];
      tmem_ld32(t_o + c * 32, oo);
      tmem_ld_wait();
      if (q < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) =
              make_uint4(pack_h2(__uint_as_float(oo[g * 8 + 0]) * inv, __uint_as_float(oo[g * 8 + 1]) * inv),
                         pack_h2(__uint_as_float(oo[g * 8 + 2]) * inv, __uint_as_float(oo[g * 8 + 3]) * inv),
                         pack_h2(__uint_as_float(oo[g * 8 + 4]) * inv, __uint_as_float(oo[g * 8 + 5]) * inv),
                         pack_h2(__uint_as_float(oo[g * 8 + 6]) * inv, __uint_as_float(oo[g * 8 + 7]) * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

struct Attn5Launch;
std::shared_ptr<Attn5Launch> attn5_build(const AttnDesc& d);
void attn5_run(const Attn5Launch& l, cudaStream_t st);

struct AttnLaunch {
  AttnParams p;
  dim3 grid;
  bool has_bias;
  std::shared_ptr<Attn5Launch> v5;  // 128-key-tile variant (attention5.cu), selected with SDM_ATTN_VARIANT
};

static int attn_variant() {
  static int v = [] {
    const char* e = getenv("SDM_ATTN_VARIANT");
    return e ? atoi(e) : 5;
  }();
  return v;
}

std::shared_ptr<AttnLaunch> attn_build(const AttnDesc& d) {
  auto L = std::make_shared<AttnLaunch>();
  if (attn_variant() == 5) {
    L->v5 = attn5_build(d);
    return L;
  }
  AttnParams& p = L->p;
  memset(&p, 0, sizeof(p));
  SDM_CHECK(d.Lq > 0 && d.Lk > 0 && d.heads > 0, "attention dims");
  SDM_CHECK(d.ldq % 8 == 0 && d.ldk % 8 == 0 && d.ldvt % 8 == 0 && d.ldo % 8 == 0, "attention strides must be multiples of 8");
  {
    const uint64_t dims[3] = {(uint64_t)d.heads * 64, (uint64_t)d.Lq, (uint64_t)d.B};
    const uint64_t str[2] = {(uint64_t)d.ldq * 2, (uint64_t)d.Lq * d.ldq * 2};
    const uint32_t box[3] = {64, 128, 1};
    make_tmap(&p.q_map, d.q, 3, dims, str, box);
  }
  {
    const uint64_t dims[3] = {(uint64_t)d.heads * 64, (uint64_t)d.Lk, (uint64_t)d.B};
    const uint64_t str[2] = {(uint64_t)d.ldk * 2, (uint64_t)d.Lk * d.ldk * 2};
    const uint32_t box[3] = {64, (uint32_t)kKeys, 1};
    make_tmap(&p.k_map, d.k, 3, dims, str, box);
  }
  {
    const uint64_t dims[3] = {(uint64_t)d.Lk, (uint64_t)d.heads * 64, (uint64_t)d.B};
    const uint64_t str[2] = {(uint64_t)d.ldvt * 2, (uint64_t)d.heads * 64 * d.ldvt * 2};
    const uint32_t box[3] = {64, 64, 1};
    make_tmap(&p.vt_map, d.vt, 3, dims, str, box);
  }
  p.bias = d.bias;
  p.bias_bstride = d.bias_bstride;
  if (d.bias) SDM_CHECK(d.bias_bstride % 4 == 0 && d.bias_bstride >= ((d.Lk + kKeys - 1) / kKeys) * kKeys, "bias must be padded to a multiple of 64 keys");
  p.out = d.out;
  p.ldo = d.ldo;
  p.Lq = d.Lq; p.Lk = d.Lk; p.heads = d.heads;
  p.n_ktiles = (d.Lk + kKeys - 1) / kKeys;
  p.scale = d.scale;
  L->grid = dim3((d.Lq + 255) / 256, d.heads, d.B);
  L->has_bias = d.bias != nullptr;
  return L;
}

void attn_run(const AttnLaunch& l, cudaStream_t st) {
  if (l.v5) return attn5_run(*l.v5, st);
  static std::once_flag once;
  std::call_once(once, [] {
    SDM_CUDA_OK(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    SDM_CUDA_OK(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
  });
  if (l.has_bias) attention_kernel<true><<<l.grid, kAttnThreads, kAttnSmem, st>>>(l.p);
  else attention_kernel<false><<<l.grid, kAttnThreads, kAttnSmem, st>>>(l.p);
  SDM_CUDA_OK(cudaGetLastError());
}

}  // namespace sdm
