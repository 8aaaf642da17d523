// Fused softmax(scale * Q K^T + key_bias) V for head dim 64 on tcgen05 (sm_100a).
//
// Replaces, for every UNet attention (self: attn1 with the trimap key bias, cross: attn2 over the 16 384
// trimap tokens), the reference chain
//   custom_prepare_attention_mask  /root/reference/src/utils/replace.py:20-72
//   custom_get_attention_scores    /root/reference/src/utils/replace.py:75-122  (baddbmm + softmax)
//   torch.bmm(probs, value) inside diffusers AttnProcessor / SlicedAttnProcessor (sdmatte_nodes.py:331-335)
// without ever materialising the L x L score matrix: S lives in TMEM, P in shared memory, O in TMEM.
//
// CTA = 256 queries (two 128-row tiles A and B) of one (batch, head); it streams 128-key tiles of K and V^T.
//   warps 0-3  : softmax warpgroup for tile A (thread r owns query row r == TMEM lane r)
//   warps 4-7  : softmax warpgroup for tile B
//   warp  8    : TMA producer (Q once; K, V^T and the per-key bias per tile; 3-stage ring)
//   warp  9    : tcgen05.mma issuer (S_X = Q_X K^T : M128 N128 K64 ; O_X += P_X V : M128 N64 K128) + TMEM alloc
//   warps 10-11: idle (they complete the third warpgroup so that setmaxnreg can move its registers to the softmax warps)
// Per key tile a warpgroup pulls S(j) out of TMEM in four 32-column chunks with the next chunk's tcgen05.ld in flight
// while the current one is exponentiated, and signals `s_free` as soon as the LAST chunk has landed in registers — the
// tensor core then computes S(j+1) while chunk 3 is still being processed and P(j) written.  O_X accumulates in TMEM over
// all key tiles (accumulating MMAs); the softmax reference max is raised lazily and per chunk (only when a score exceeds
// it by more than 2^8, before any exp of that chunk): earlier chunks' probabilities (packed fp16 in registers), the
// running sum and the O rows in TMEM (tcgen05.ld / st) are rescaled by alpha = 2^(m_old - m_new); nothing is ever
// re-read from S, which is what allows the early release.
// Measured alternatives (same box, B2 h5 16384x16384, r1h/r1i): two-pass softmax with O in registers 437 TFLOP/s;
// 64-key tiles with double-buffered S and P 389-485; this variant 480-530.  At d=64 one ex2 is needed per 256 tensor
// FLOP, so the MUFU pipe (16/clk/SM) caps the kernel near 50 % of the tensor peak.
// TMEM columns: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384).
// The per-key bias is expected pre-multiplied by log2(e); scores are handled in the log2 domain, statistics in fp32,
// and scores are NOT rounded to fp16 before the softmax (the reference does).
#include "common.cuh"
#include "kernels.h"
#include "tmap.h"

#include <mutex>
#include <type_traits>

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

namespace a5 {
constexpr int kStages = 3;
constexpr int kThreads = 384;   // 8 softmax warps + TMA warp + MMA warp + 2 idle warps (a full third warpgroup for setmaxnreg)
constexpr uint32_t kQBytes = 128 * 128;        // one 128x64 fp16 tile
constexpr uint32_t kPBytes = 2 * 128 * 128;    // 128 x 128 fp16 as two 64-key blocks
constexpr uint32_t kKBytes = 128 * 128;        // 128 keys x 64 d
constexpr uint32_t kVBytes = 2 * 64 * 128;     // 64 d x 128 keys as two 64-key blocks
constexpr uint32_t kStageBytes = kKBytes + kVBytes;
constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffP = 2 * kQBytes;
constexpr uint32_t kOffStage = kOffP + 2 * kPBytes;
constexpr uint32_t kOffBias = kOffStage + kStages * kStageBytes;
constexpr uint32_t kOffBar = kOffBias + kStages * 512;
constexpr uint32_t kSmem = kOffBar + 256 + 1024;
}  // namespace a5

__device__ __forceinline__ uint32_t hmul2_u32(uint32_t a, __half2 s) {
  __half2 v = *reinterpret_cast<__half2*>(&a);
  v = __hmul2(v, s);
  return *reinterpret_cast<uint32_t*>(&v);
}

// OPT = false: "v5" chunk (max -> vote -> exp);  OPT = true: "v6" optimistic chunk (exp -> vote -> rare redo), see below.
template <bool HAS_BIAS, bool OPT>
__global__ void __launch_bounds__(a5::kThreads, 1) attention_kernel(const __grid_constant__ AttnParams p) {
  using namespace a5;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + kOffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (4 + s); };
  auto s_full = [&](int x) { return bar + 8u * (7 + x); };
  auto s_free = [&](int x) { return bar + 8u * (9 + x); };
  auto p_full = [&](int x) { return bar + 8u * (11 + x); };
  auto o_full = [&](int x) { return bar + 8u * (13 + x); };
  const uint32_t tmem_slot = bar + 8u * 15;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kOffBar + 8 * 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n = p.n_ktiles;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int x = 0; x < 2; ++x) { mbar_init(s_full(x), 1); mbar_init(s_free(x), 128); mbar_init(p_full(x), 128); mbar_init(o_full(x), 1); }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  // Register budget: ptxas caps the kernel at 168 registers (3 warps per scheduler); the softmax warps need ~200 for
  // P[64] + two S chunks + the exponentials, and spilled inside the MUFU loop.  The data-movement warpgroup hands its
  // registers over: 2 x 232 + 40 = 3 x 168 per scheduler.
  if (warp >= 8) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
  if (warp == 8) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      tma_prefetch_desc(&p.q_map); tma_prefetch_desc(&p.k_map); tma_prefetch_desc(&p.vt_map);
      mbar_expect_tx(q_full, 2 * kQBytes);
      tma_load_3d(base + kOffQ, &p.q_map, q_full, h * 64, q0, b);
      tma_load_3d(base + kOffQ + kQBytes, &p.q_map, q_full, h * 64, q0 + 128, b);
      for (int j = 0; j < n; ++j) {
        const int s = j % kStages;
        const uint32_t f = (uint32_t)(j / kStages);
        mbar_wait(kv_empty(s), (f & 1u) ^ 1u);
        const uint32_t kdst = base + kOffStage + s * kStageBytes;
        mbar_expect_tx(kv_full(s), kStageBytes + (HAS_BIAS ? 512u : 0u));
        tma_load_3d(kdst, &p.k_map, kv_full(s), h * 64, j * 128, b);
        tma_load_3d(kdst + kKBytes, &p.vt_map, kv_full(s), j * 128, h * 64, b);
        tma_load_3d(kdst + kKBytes + 64 * 128, &p.vt_map, kv_full(s), j * 128 + 64, h * 64, b);
        if (HAS_BIAS) bulk_load_1d(base + kOffBias + s * 512, p.bias + (long long)b * p.bias_bstride + (long long)j * 128, 512, kv_full(s));
      }
    }
  } else if (warp == 9) {
    // ======================================= MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128);
      constexpr uint32_t idesc_o = umma_idesc_f16(64);
      auto issue_s = [&](int x, int stage) {
        const uint64_t ad = umma_desc_k128(base + kOffQ + x * kQBytes);
        const uint64_t bd = umma_desc_k128(base + kOffStage + stage * kStageBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + x * 128, ad + 2 * k, bd + 2 * k, idesc_s, k != 0);
        umma_commit(s_full(x));
      };
      mbar_wait(q_full, 0);
      mbar_wait(kv_full(0), 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < n; ++j) {
        const int s = j % kStages;
        const int s1 = (j + 1) % kStages;
        if (j + 1 < n) {
          mbar_wait(kv_full(s1), (uint32_t)((j + 1) / kStages) & 1u);
          for (int x = 0; x < 2; ++x) {
            mbar_wait(s_free(x), (uint32_t)j & 1u);  // S_x(j) is in the warpgroup's registers
            tc_fence_after();
            issue_s(x, s1);
          }
        }
        for (int x = 0; x < 2; ++x) {
          mbar_wait(p_full(x), (uint32_t)j & 1u);  // P_x(j) in smem
          tc_fence_after();
          const uint32_t pa = base + kOffP + x * kPBytes;
          const uint32_t vb = base + kOffStage + s * kStageBytes + kKBytes;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t ad = umma_desc_k128(pa + (k >> 2) * (128 * 128)) + 2 * (k & 3);
            const uint64_t bd = umma_desc_k128(vb + (k >> 2) * (64 * 128)) + 2 * (k & 3);
            umma_f16(tmem + 256 + x * 64, ad, bd, idesc_o, (j | k) != 0);
          }
          umma_commit(o_full(x));
        }
        umma_commit(kv_empty(s));
      }
    }
  }
  } else {
    // ======================================= softmax warpgroups =================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int x = warp >> 2;                 // 0: tile A, 1: tile B
    const int r = (warp & 3) * 32 + lane;    // row within the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem + lane_base + x * 128;
    const uint32_t t_o = tmem + lane_base + 256 + x * 64;
    uint8_t* p_row = base_ptr + kOffP + x * kPBytes + (r >> 3) * 1024 + (r & 7) * 128;
    const uint32_t rx = (uint32_t)(r & 7) << 4;
    const float sc = p.scale * 1.4426950408889634f;
    constexpr float kTau = 8.0f;
    float m2 = -INFINITY, l = 0.f;

    for (int j = 0; j < n; ++j) {
      const int s = j % kStages;
      const bool tail = (!HAS_BIAS) && (j == n - 1) && ((p.Lk & 127) != 0);
      const int kbase = j * 128;
      if (HAS_BIAS) mbar_wait(kv_full(s), (uint32_t)(j / kStages) & 1u);  // bias tile visible to this thread
      mbar_wait(s_full(x), (uint32_t)j & 1u);
      tc_fence_after();
      const float4* bias4 = reinterpret_cast<const float4*>(base_ptr + kOffBias + s * 512);
      uint32_t P[64];   // the 128 probabilities of this row, packed fp16x2
      float rowsum = 0.f;
      uint32_t r0[32], r1[32];
      tmem_ld32(t_s, r0);
      tmem_ld32(t_s + 32, r1);

      // one 32-column chunk: (optionally raise the reference) then e = 2^(x - m2), row sum, packed P
      auto chunk = [&](uint32_t (&rr)[32], auto c_tag) {
        constexpr int c = decltype(c_tag)::value;
        if (!HAS_BIAS && tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + c * 32 + i >= p.Lk) rr[i] = 0xff800000u;  // -inf
        }
        float xs[32];
        float cm = -INFINITY;
        if (HAS_BIAS) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 bq = bias4[c * 8 + g];
            xs[g * 4 + 0] = fmaf(__uint_as_float(rr[g * 4 + 0]), sc, bq.x);
            xs[g * 4 + 1] = fmaf(__uint_as_float(rr[g * 4 + 1]), sc, bq.y);
            xs[g * 4 + 2] = fmaf(__uint_as_float(rr[g * 4 + 2]), sc, bq.z);
            xs[g * 4 + 3] = fmaf(__uint_as_float(rr[g * 4 + 3]), sc, bq.w);
            cm = fmaxf(cm, fmaxf(fmaxf(xs[g * 4], xs[g * 4 + 1]), fmaxf(xs[g * 4 + 2], xs[g * 4 + 3])));
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            cm = fmaxf(cm, fmaxf(fmaxf(__uint_as_float(rr[g * 4]), __uint_as_float(rr[g * 4 + 1])),
                                 fmaxf(__uint_as_float(rr[g * 4 + 2]), __uint_as_float(rr[g * 4 + 3]))));
          cm *= sc;
        }
        if (__any_sync(0xffffffffu, cm > m2 + kTau)) {
          // raise the reference BEFORE exponentiating this chunk; rescale what was accumulated with the old one
          const float m_new = fmaxf(m2, cm);
          const float alpha = ex2f(m2 - m_new);  // 0 when m2 = -inf
          if (j > 0) {
            mbar_wait(o_full(x), (uint32_t)(j - 1) & 1u);  // every P·V issued so far has landed in O
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t oo[32];
              tmem_ld32(t_o + cc * 32, oo);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) oo[i] = __float_as_uint(__uint_as_float(oo[i]) * alpha);
              tmem_st32(t_o + cc * 32, oo);
            }
            tmem_st_wait();
          }
          const __half2 a2 = __float2half2_rn(alpha);
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i < c * 16) P[i] = hmul2_u32(P[i], a2);
          rowsum *= alpha;
          l *= alpha;
          m2 = m_new;
        }
        const float neg_m = -m2;
        // all 32 exponentials first (32 independent FFMA -> MUFU chains keep the XU pipe fed), then the sums / packing
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
          e[i] = HAS_BIAS ? ex2f(xs[i] + neg_m) : ex2f(fmaf(__uint_as_float(rr[i]), sc, neg_m));
        float part[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          part[q] = ((e[q * 8 + 0] + e[q * 8 + 1]) + (e[q * 8 + 2] + e[q * 8 + 3])) + ((e[q * 8 + 4] + e[q * 8 + 5]) + (e[q * 8 + 6] + e[q * 8 + 7]));
          P[(c * 4 + q) * 4 + 0] = pack_h2(e[q * 8 + 0], e[q * 8 + 1]);
          P[(c * 4 + q) * 4 + 1] = pack_h2(e[q * 8 + 2], e[q * 8 + 3]);
          P[(c * 4 + q) * 4 + 2] = pack_h2(e[q * 8 + 4], e[q * 8 + 5]);
          P[(c * 4 + q) * 4 + 3] = pack_h2(e[q * 8 + 6], e[q * 8 + 7]);
        }
        rowsum += (part[0] + part[1]) + (part[2] + part[3]);
      };

      // v6 chunk: exponentiate OPTIMISTICALLY against the current reference m2 and look at the result afterwards.  The
      // v5 order (row max -> vote -> exp) puts a 16-deep FMNMX chain and a branch in front of every 32 MUFU ops, and with
      // only two softmax warps per scheduler running in lockstep the XU pipe idled half of the time (ncu r1k: XU 50 %,
      // ~3800 clk per key tile against 2048 clk of MUFU work).  Here the steady-state chunk is ONE basic block of
      // 32 FFMA -> 32 MUFU -> sums/packing; the chunk's sum doubles as the overflow test: csum <= 2^10 proves every
      // e <= 2^10 (fp16-safe, fp32 sums safe), anything else (a large score, +inf from m2 = -inf on the very first chunk,
      // NaN from -inf - -inf) takes the rare redo path, which raises the reference exactly like v5 and recomputes the chunk
      // from the S registers that are still live.
      auto chunk_opt = [&](uint32_t (&rr)[32], auto c_tag) {
        constexpr int c = decltype(c_tag)::value;
        constexpr float kLimit = 1024.0f;
        if (!HAS_BIAS && tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + c * 32 + i >= p.Lk) rr[i] = 0xff800000u;  // -inf
        }
        float csum;
        auto exp_pack = [&]() {
          const float neg_m = -m2;
          float e[32];
          if (HAS_BIAS) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 bq = bias4[c * 8 + g];
              e[g * 4 + 0] = ex2f(fmaf(__uint_as_float(rr[g * 4 + 0]), sc, bq.x) + neg_m);
              e[g * 4 + 1] = ex2f(fmaf(__uint_as_float(rr[g * 4 + 1]), sc, bq.y) + neg_m);
              e[g * 4 + 2] = ex2f(fmaf(__uint_as_float(rr[g * 4 + 2]), sc, bq.z) + neg_m);
              e[g * 4 + 3] = ex2f(fmaf(__uint_as_float(rr[g * 4 + 3]), sc, bq.w) + neg_m);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) e[i] = ex2f(fmaf(__uint_as_float(rr[i]), sc, neg_m));
          }
          float part[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            part[q] = ((e[q * 8 + 0] + e[q * 8 + 1]) + (e[q * 8 + 2] + e[q * 8 + 3])) + ((e[q * 8 + 4] + e[q * 8 + 5]) + (e[q * 8 + 6] + e[q * 8 + 7]));
            P[(c * 4 + q) * 4 + 0] = pack_h2(e[q * 8 + 0], e[q * 8 + 1]);
            P[(c * 4 + q) * 4 + 1] = pack_h2(e[q * 8 + 2], e[q * 8 + 3]);
            P[(c * 4 + q) * 4 + 2] = pack_h2(e[q * 8 + 4], e[q * 8 + 5]);
            P[(c * 4 + q) * 4 + 3] = pack_h2(e[q * 8 + 6], e[q * 8 + 7]);
          }
          csum = (part[0] + part[1]) + (part[2] + part[3]);
        };
        exp_pack();
        if (__any_sync(0xffffffffu, !(csum <= kLimit))) {
          // rare: raise the reference to this chunk's true maximum, rescale what was accumulated, redo the chunk
          float cm = -INFINITY;
          if (HAS_BIAS) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 bq = bias4[c * 8 + g];
              cm = fmaxf(cm, fmaxf(fmaxf(fmaf(__uint_as_float(rr[g * 4 + 0]), sc, bq.x), fmaf(__uint_as_float(rr[g * 4 + 1]), sc, bq.y)),
                                   fmaxf(fmaf(__uint_as_float(rr[g * 4 + 2]), sc, bq.z), fmaf(__uint_as_float(rr[g * 4 + 3]), sc, bq.w))));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) cm = fmaxf(cm, __uint_as_float(rr[i]));
            cm *= sc;
          }
          const float m_new = fmaxf(m2, cm);
          const float alpha = (m_new == m2) ? 1.0f : ex2f(m2 - m_new);  // 0 when m2 = -inf
          if (j > 0) {
            mbar_wait(o_full(x), (uint32_t)(j - 1) & 1u);  // every P·V issued so far has landed in O
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t oo[32];
              tmem_ld32(t_o + cc * 32, oo);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) oo[i] = __float_as_uint(__uint_as_float(oo[i]) * alpha);
              tmem_st32(t_o + cc * 32, oo);
            }
            tmem_st_wait();
          }
          const __half2 a2 = __float2half2_rn(alpha);
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i < c * 16) P[i] = hmul2_u32(P[i], a2);
          rowsum *= alpha;
          l *= alpha;
          m2 = m_new;
          exp_pack();
        }
        rowsum += csum;
      };
      auto run_chunk = [&](uint32_t (&rr)[32], auto c_tag) {
        if constexpr (OPT) chunk_opt(rr, c_tag);
        else chunk(rr, c_tag);
      };

      tmem_ld_wait();
      run_chunk(r0, std::integral_constant<int, 0>{});
      tmem_ld32(t_s + 64, r0);   // chunk 2 in flight while chunk 1 is processed
      run_chunk(r1, std::integral_constant<int, 1>{});
      tmem_ld_wait();
      tmem_ld32(t_s + 96, r1);   // chunk 3 in flight while chunk 2 is processed
      run_chunk(r0, std::integral_constant<int, 2>{});
      tmem_ld_wait();
      // all of S(j) is in registers: let the tensor core start S(j+1)
      tc_fence_before();
      mbar_arrive(s_free(x));
      run_chunk(r1, std::integral_constant<int, 3>{});
      l += rowsum;
      // P(j) overwrites the smem buffer P·V(j-1) reads
      if (j > 0) mbar_wait(o_full(x), (uint32_t)(j - 1) & 1u);
#pragma unroll
      for (int ch = 0; ch < 16; ++ch)
        *reinterpret_cast<uint4*>(p_row + (ch >> 3) * (128 * 128) + ((uint32_t)((ch & 7) << 4) ^ rx)) =
            make_uint4(P[ch * 4], P[ch * 4 + 1], P[ch * 4 + 2], P[ch * 4 + 3]);
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full(x));
    }
    // ---- normalise and store
    mbar_wait(o_full(x), (uint32_t)(n - 1) & 1u);
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

struct AttnLaunch {
  AttnParams p;
  dim3 grid;
  bool has_bias;
};

std::shared_ptr<AttnLaunch> attn_build(const AttnDesc& d) {
  auto L = std::make_shared<AttnLaunch>();
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
    const uint32_t box[3] = {64, 128, 1};
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
  if (d.bias) SDM_CHECK(d.bias_bstride % 4 == 0 && d.bias_bstride >= ((d.Lk + 127) / 128) * 128, "bias must be padded to 128 keys");
  p.out = d.out;
  p.ldo = d.ldo;
  p.Lq = d.Lq; p.Lk = d.Lk; p.heads = d.heads;
  p.n_ktiles = (d.Lk + 127) / 128;
  p.scale = d.scale;
  L->grid = dim3((d.Lq + 255) / 256, d.heads, d.B);
  L->has_bias = d.bias != nullptr;
  return L;
}

void attn_run(const AttnLaunch& l, cudaStream_t st) {
  static std::once_flag once;
  std::call_once(once, [] {
    SDM_CUDA_OK(cudaFuncSetAttribute(attention_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, a5::kSmem));
    SDM_CUDA_OK(cudaFuncSetAttribute(attention_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, a5::kSmem));
    SDM_CUDA_OK(cudaFuncSetAttribute(attention_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, a5::kSmem));
    SDM_CUDA_OK(cudaFuncSetAttribute(attention_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, a5::kSmem));
  });
  static const int variant = [] { const char* e = getenv("SDM_ATTN"); return e ? atoi(e) : 6; }();  // 5 = v5 chunk (A/B)
  if (variant == 5) {
    if (l.has_bias) attention_kernel<true, false><<<l.grid, a5::kThreads, a5::kSmem, st>>>(l.p);
    else attention_kernel<false, false><<<l.grid, a5::kThreads, a5::kSmem, st>>>(l.p);
  } else {
    if (l.has_bias) attention_kernel<true, true><<<l.grid, a5::kThreads, a5::kSmem, st>>>(l.p);
    else attention_kernel<false, true><<<l.grid, a5::kThreads, a5::kSmem, st>>>(l.p);
  }
  SDM_CUDA_OK(cudaGetLastError());
}

}  // namespace sdm
