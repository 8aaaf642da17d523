// Fused softmax(scale * Q K^T + key_bias) V for head dim 64 on tcgen05 (sm_100a).
//
// Replaces, for every UNet attention (self: attn1 with the trimap key bias, cross: attn2 over the 16 384
// trimap tokens), the reference chain
//   custom_prepare_attention_mask  /root/reference/src/utils/replace.py:20-72
//   custom_get_attention_scores    /root/reference/src/utils/replace.py:75-122  (baddbmm + softmax)
//   torch.bmm(probs, value) inside diffusers AttnProcessor / SlicedAttnProcessor (sdmatte_nodes.py:331-335)
// without ever materialising the L x L score matrix: S, P and O all live in tensor memory.
//
// CTA = 256 queries (two 128-row tiles A and B) of one (batch, head); it streams 128-key tiles of K and V^T.
//   warps 0-3  : softmax warpgroup for tile A (thread r owns query row r == TMEM lane r)
//   warps 4-7  : softmax warpgroup for tile B
//   warp  8    : TMA producer (Q once; K, V^T and the per-key bias per tile; 5-stage ring)
//   warps 9,10 : tcgen05.mma issuers, ONE PER QUERY TILE (warp 9 also owns the TMEM allocation)
//                S_X  = Q_X K^T : two M128 N64 K64 halves (keys 0-63 / 64-127), both operands from shared memory
//                O_X += P_X V   : M128 N64 K128, A = P_X read from TENSOR MEMORY ("TS" MMA), B = V^T from shared memory
// TMEM columns: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512)   (P = packed fp16x2).
//
// v7 (r1l): P no longer goes through shared memory.  ncu of v5/v6 (profiles/r1l_*) showed the softmax warps spending 18 %
// of their time in the 16 STS.128 + proxy fence + arrive that published P, and the shared-memory pipe carrying 256 KB per key
// tile (MMA operand reads 160 KB of which P 64 KB, P writes 64 KB, TMA 32 KB) = 2048 clk at 128 B/clk — as much as the MUFU
// work itself.  With P written straight from registers to TMEM (one tcgen05.st.x16 per 32-key chunk, right after the
// exponentials) the shared-memory traffic halves, the 64-register P array disappears (no setmaxnreg, no spills) and the
// per-tile publication tail is a tcgen05.wait::st.
// Per key tile a warpgroup pulls S(j) out of TMEM in four 32-column chunks, the next chunk's tcgen05.ld in flight while the
// current one is exponentiated (PIPE, the biased kernels: also across tile boundaries — chunk 0 of tile j+1 is requested before
// chunk 3 of tile j is exponentiated; the other kernels load chunks 0 and 1 together at the top of a tile).  S is produced and released in two 64-key halves: the low half is handed back to the tensor
// core as soon as chunks 0-1 sit in registers (the very start of the tile), the high half once chunks 2-3 do, so S(j+1) is
// complete long before the warpgroup needs it.  O_X accumulates in TMEM over all key tiles.
// Softmax reference (lazy, optimistic): a chunk is exponentiated against the current reference m2 FIRST; its sum doubles as
// the overflow test (csum <= 2^10 proves every e <= 2^10: fp16-safe).  Anything else (a score far above the reference,
// +inf from m2 = -inf on the very first chunk, NaN) takes the rare redo path: raise m2 to the chunk's true maximum, rescale
// the running sum, the chunks of P already stored for this tile and the O rows in TMEM (tcgen05.ld / st) by
// alpha = 2^(m2_old - m2_new), and exponentiate the chunk again from the S registers that are still live.  Steady state is
// one basic block per chunk: 16 FFMA2 -> 32 MUFU.EX2 -> 16 FADD2 / 16 F2FP -> tcgen05.st (fp32x2 packed arithmetic).
// Measured (B2 h5 16384x16384, same box): v5 (row max -> vote -> exp, P via smem) 532 TFLOP/s, v6 (optimistic, P via smem) 478.
// v8 (r1p) also issued P.V in two 64-key halves; measured neutral to -2 % in the step, removed in round 2 (history: profiles/r1q_*).
// v9 (r1q): one MMA issuer warp per query tile running the tile's fixed event sequence with BLOCKING mbarrier waits
//   S_lo(t+1) <- s_free_lo(t),  S_hi(t+1) <- s_free_hi(t),  P.V(t) <- p_full(t)
// instead of one thread polling the ~12 barriers of both tiles round-robin.  ncu r1p/r1q: the single poller was the bottleneck
// (the softmax warps waited 25 % of their time for S_hi(j), issued late): same box, B2 h5 16384 x 16384: 501 -> 622 TFLOP/s;
// in the step cross attention 54.1 -> 42.4 ms.
// An FMA-pipe polynomial exp2 (degree 4, packed fp32x2) for 4 / 8 of the 16 column pairs of a chunk was measured three times
// and removed: under the polling issuer (r1o) 519 -> 494 / 464 TFLOP/s at L0, under the sequenced issuers (r1v) 623 -> 603 /
// 550, and again with the warp-uniform issuers (r3a: 690 -> 653 / 619 / 559 for 1/4, 1/3, 1/2 of the pairs): the softmax warps
// are bound by their dependent instruction chain and issue slots, not by MUFU throughput (XU pipe 65 % in ncu r1q).
// Self-attention only streams the keys that can have a non-zero probability (p.ntiles, see key_compact_kernel).
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
  const int* ntiles;  // per-sample key-tile count (compacted keys) or null
  __half* out;
  long long ldo;
  int Lq, Lk, heads, n_ktiles;
  float scale;
};

namespace a7 {
constexpr int kStages = 5;
constexpr int kThreads = 352;  // 8 softmax warps + TMA producer + two MMA issuers (one per query tile)
constexpr uint32_t kQBytes = 128 * 128;        // one 128x64 fp16 tile
constexpr uint32_t kKBytes = 128 * 128;        // 128 keys x 64 d
constexpr uint32_t kVBytes = 2 * 64 * 128;     // 64 d x 128 keys as two 64-key blocks
constexpr uint32_t kStageBytes = kKBytes + kVBytes;
constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffStage = 2 * kQBytes;
constexpr uint32_t kOffBias = kOffStage + kStages * kStageBytes;
constexpr uint32_t kOffBar = kOffBias + kStages * 512;
constexpr uint32_t kSmem = kOffBar + 256 + 1024;
constexpr uint32_t kColS = 0, kColO = 256, kColP = 384;
}  // namespace a7

__device__ __forceinline__ uint32_t hmul2_u32(uint32_t a, __half2 s) {
  __half2 v = *reinterpret_cast<__half2*>(&a);
  v = __hmul2(v, s);
  return *reinterpret_cast<uint32_t*>(&v);
}

// TAIL: Lk is not a multiple of 128 and there is no bias to carry the -inf padding (never the case inside the engine;
// kept out of the common instantiations: even skipped, the masking code cost a BSSY/branch per chunk and i-cache misses)
template <bool HAS_BIAS, bool TAIL, bool PIPE>
__global__ void __launch_bounds__(a7::kThreads, 1) attention_kernel(const __grid_constant__ AttnParams p) {
  using namespace a7;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + kOffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kStages + s); };
  // S is produced and released in two 64-key halves (hf = 0: keys 0-63, 1: keys 64-127)
  auto s_full = [&](int x, int hf) { return bar + 8u * (1 + 2 * kStages + x * 2 + hf); };
  auto s_free = [&](int x, int hf) { return bar + 8u * (5 + 2 * kStages + x * 2 + hf); };
  // P is published and P.V committed per whole 128-key tile (only index hf = 1 is used; the per-half variant of r1p was
  // measured neutral to -2 % in the step and removed in round 2)
  auto p_full = [&](int x, int hf) { return bar + 8u * (9 + 2 * kStages + x * 2 + hf); };
  auto o_full = [&](int x, int hf) { return bar + 8u * (13 + 2 * kStages + x * 2 + hf); };
  const uint32_t tmem_slot = bar + 8u * (17 + 2 * kStages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kOffBar + 8 * (17 + 2 * kStages));
  static_assert(8 * (17 + 2 * kStages) + 4 <= 256, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n = p.ntiles ? min(__ldg(p.ntiles + b), p.n_ktiles) : p.n_ktiles;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2); }  // both issuers commit
    for (int x = 0; x < 2; ++x) {
      for (int hf = 0; hf < 2; ++hf) {
        mbar_init(s_full(x, hf), 1); mbar_init(s_free(x, hf), 128);
        mbar_init(p_full(x, hf), 128); mbar_init(o_full(x, hf), 1);
      }
    }
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
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n; ++j) {
        mbar_wait(kv_empty(s), ph ^ 1u);
        const uint32_t kdst = base + kOffStage + s * kStageBytes;
        mbar_expect_tx(kv_full(s), kStageBytes + (HAS_BIAS ? 512u : 0u));
        tma_load_3d(kdst, &p.k_map, kv_full(s), h * 64, j * 128, b);
        tma_load_3d(kdst + kKBytes, &p.vt_map, kv_full(s), j * 128, h * 64, b);
        tma_load_3d(kdst + kKBytes + 64 * 128, &p.vt_map, kv_full(s), j * 128 + 64, h * 64, b);
        if (HAS_BIAS) bulk_load_1d(base + kOffBias + s * 512, p.bias + (long long)b * p.bias_bstride + (long long)j * 128, 512, kv_full(s));
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ======================================= MMA issuer(s) ======================================
    // whole warp, one elected lane issues: inside an `if (lane == 0)` the compiler wrapped every tcgen05.mma in an ELECT /
    // R2UR.BROADCAST loop (conv_swap_halo.cu, r2w: +10 % on the convs from this change alone)
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = umma_idesc_f16(64);
      constexpr uint32_t idesc_o = umma_idesc_f16(64);
      // S_x half hf = Q_x (128 x 64) . K[hf*64 .. hf*64+63]^T  -> S columns [hf*64, hf*64+64)
      auto issue_s = [&](int x, int hf, int stage) {
        const uint64_t ad = umma_desc_k128(base + kOffQ + x * kQBytes);
        const uint64_t bd = umma_desc_k128(base + kOffStage + stage * kStageBytes + hf * (64 * 128));
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem + kColS + x * 128 + hf * 64, ad + 2 * k, bd + 2 * k, idesc_s, k != 0);
          umma_commit(s_full(x, hf));
        }
        __syncwarp();
      };
      // O_x += P_x[:, keys k0*16 .. k1*16) . V[those keys]   (K16 steps k0..k1-1 of the 8 in a 128-key tile)
      auto issue_pv = [&](int x, int stage, int j, int k0, int k1, uint32_t commit_bar) {
        const uint32_t vb = base + kOffStage + stage * kStageBytes + kKBytes;
        if (leader) {
#pragma unroll
          for (int k = k0; k < k1; ++k) {
            const uint64_t bd = umma_desc_k128(vb + (k >> 2) * (64 * 128)) + 2 * (k & 3);
            umma_f16_ts(tmem + kColO + x * 64, tmem + kColP + x * 64 + k * 8, bd, idesc_o, (j | k) != 0);
          }
          umma_commit(commit_bar);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      {
        const int x = warp - 9;  // this issuer's query tile
        auto wait_kv = [&](int t) { mbar_wait(kv_full(t % kStages), (uint32_t)(t / kStages) & 1u); };
        wait_kv(0);
        tc_fence_after();
        issue_s(x, 0, 0);
        issue_s(x, 1, 0);
        for (int t = 0; t < n; ++t) {
          const int st = t % kStages;
          const bool more = t + 1 < n;
          if (more) {
            wait_kv(t + 1);
            mbar_wait(s_free(x, 0), (uint32_t)t & 1u);
            tc_fence_after();
            issue_s(x, 0, (t + 1) % kStages);
          }
          if (more) {
            mbar_wait(s_free(x, 1), (uint32_t)t & 1u);
            tc_fence_after();
            issue_s(x, 1, (t + 1) % kStages);
          }
          mbar_wait(p_full(x, 1), (uint32_t)t & 1u);
          tc_fence_after();
          issue_pv(x, st, t, 0, 8, o_full(x, 1));
          if (leader) umma_commit(kv_empty(st));  // this tile's MMAs on the stage are done (the barrier counts both issuers)
          __syncwarp();
        }
      }
    }
  } else if (warp < 8) {
    // ======================================= softmax warpgroups =================================
    const int x = warp >> 2;                 // 0: tile A, 1: tile B
    const int r = (warp & 3) * 32 + lane;    // row within the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem + lane_base + kColS + x * 128;
    const uint32_t t_o = tmem + lane_base + kColO + x * 64;
    const uint32_t t_p = tmem + lane_base + kColP + x * 64;
    const float sc = p.scale * 1.4426950408889634f;
    const uint64_t sc2 = pack_f2(sc, sc);
    constexpr float kLimit = 1024.0f;
    float m2 = -INFINITY, l = 0.f;
    int s = 0;
    uint32_t ph = 0;

    uint32_t r0[32], r1[32];  // S chunks: even chunks in r0, odd in r1
    for (int j = 0; j < n; ++j) {
      const bool tail = TAIL && (j == n - 1);
      const int kbase = j * 128;
      if (HAS_BIAS) mbar_wait(kv_full(s), ph);  // bias tile visible to this thread
      if (!PIPE || j == 0) {
        mbar_wait(s_full(x, 0), (uint32_t)j & 1u);
        tc_fence_after();
      }
      const uint2* bias2 = reinterpret_cast<const uint2*>(base_ptr + kOffBias + s * 512);
      float rowsum = 0.f;
      if (!PIPE || j == 0) {
        tmem_ld32(t_s, r0);
        if (!PIPE) tmem_ld32(t_s + 32, r1);
      }

      // one 32-key chunk c (runtime, 0..3) held in rr: exponentiate, pack, (rarely) redo, store to P
      auto chunk = [&](uint32_t (&rr)[32], int c) {
        if (TAIL && tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + c * 32 + i >= p.Lk) rr[i] = 0xff800000u;  // -inf
        }
        uint32_t pk[16];
        float csum;
        auto exp_pack = [&]() {
          const uint64_t nm2 = pack_f2(-m2, -m2);
          float e[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            uint64_t v = pack_f2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1]));
            if (HAS_BIAS) {
              const uint2 bq = bias2[c * 16 + i];
              v = add_f2(fma_f2(v, sc2, ((uint64_t)bq.y << 32) | bq.x), nm2);
            } else {
              v = fma_f2(v, sc2, nm2);
            }
            float x0, x1;
            unpack_f2(v, x0, x1);
            e[2 * i] = ex2f(x0);
            e[2 * i + 1] = ex2f(x1);
          }
          // pairwise tree on packed lanes: 8 + 4 + 2 + 1 FADD2, then one FADD
          uint64_t t8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t8[i] = add_f2(pack_f2(e[4 * i], e[4 * i + 1]), pack_f2(e[4 * i + 2], e[4 * i + 3]));
#pragma unroll
          for (int i = 0; i < 4; ++i) t8[i] = add_f2(t8[2 * i], t8[2 * i + 1]);
          t8[0] = add_f2(add_f2(t8[0], t8[1]), add_f2(t8[2], t8[3]));
          float s0, s1;
          unpack_f2(t8[0], s0, s1);
          csum = s0 + s1;
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_h2(e[2 * i], e[2 * i + 1]);
        };
        exp_pack();
        if (__any_sync(0xffffffffu, !(csum <= kLimit))) {
          // rare: raise the reference to this chunk's true maximum, rescale what was accumulated, redo the chunk
          float cm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0 = __uint_as_float(rr[2 * i]) * sc, x1 = __uint_as_float(rr[2 * i + 1]) * sc;
            if (HAS_BIAS) {
              const uint2 bq = bias2[c * 16 + i];
              x0 += __uint_as_float(bq.x);
              x1 += __uint_as_float(bq.y);
            }
            cm = fmaxf(cm, fmaxf(x0, x1));
          }
          const float m_new = fmaxf(m2, cm);
          const float alpha = (m_new == m2) ? 1.0f : ex2f(m2 - m_new);  // 0 when m2 = -inf
          // every P.V issued so far (all of tile j-1) must have landed in O before its rows are rescaled
          if (j > 0) mbar_wait(o_full(x, 1), (uint32_t)(j - 1) & 1u);
          if (j > 0) {
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
          }
          const __half2 a2 = __float2half2_rn(alpha);
          tmem_st_wait();  // this tile's earlier P chunks are in TMEM before they are read back
          // chunks of this tile stored but not yet handed to the tensor core
          for (int cc = 0; cc < c; ++cc) {
            uint32_t pp[16];
            tmem_ld16(t_p + cc * 16, pp);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) pp[i] = hmul2_u32(pp[i], a2);
            tmem_st16(t_p + cc * 16, pp);
          }
          rowsum *= alpha;
          l *= alpha;
          m2 = m_new;
          exp_pack();
        }
        // P(j) overwrites the columns P.V(j-1) reads
        if (j > 0 && c == 0) {
          mbar_wait(o_full(x, 1), (uint32_t)(j - 1) & 1u);
          tc_fence_after();
        }
        tmem_st16(t_p + c * 16, pk);
        rowsum += csum;
      };

      if (PIPE) {
        // one S chunk always in flight, across tile boundaries too: chunk g+1 is requested before chunk g is exponentiated
        tmem_ld_wait();              // chunk 0 (requested during the previous tile's chunk 3)
        tmem_ld32(t_s + 32, r1);
        chunk(r0, 0);
        tmem_ld_wait();              // chunk 1 landed: the low half may be overwritten with S_lo(j+1)
        tc_fence_before();
        mbar_arrive(s_free(x, 0));
        mbar_wait(s_full(x, 1), (uint32_t)j & 1u);
        tc_fence_after();
        tmem_ld32(t_s + 64, r0);
        chunk(r1, 1);
        tmem_ld_wait();
        tmem_ld32(t_s + 96, r1);
        chunk(r0, 2);
        tmem_ld_wait();              // chunk 3 landed: the high half may be overwritten with S_hi(j+1)
        tc_fence_before();
        mbar_arrive(s_free(x, 1));
        if (j + 1 < n) {
          mbar_wait(s_full(x, 0), (uint32_t)(j + 1) & 1u);  // S_lo(j+1): issued when the low half was released
          tc_fence_after();
          tmem_ld32(t_s, r0);
        }
        chunk(r1, 3);
      } else {
      tmem_ld_wait();
      // chunks 0-1 are in registers: the tensor core may overwrite the low half with S_lo(j+1) already
      tc_fence_before();
      mbar_arrive(s_free(x, 0));
#pragma unroll 1
      for (int cp = 0; cp < 2; ++cp) {
        chunk(r0, 2 * cp);
        if (cp == 0) {
          mbar_wait(s_full(x, 1), (uint32_t)j & 1u);  // high half of S(j) (issued long ago)
          tc_fence_after();
          tmem_ld32(t_s + 64, r0);   // chunk 2 in flight while chunk 1 is processed
        } else {
          tmem_ld_wait();            // chunk 3 has landed: the high half may be overwritten with S_hi(j+1)
          tc_fence_before();
          mbar_arrive(s_free(x, 1));
        }
        chunk(r1, 2 * cp + 1);
        if (cp == 0) {
          tmem_ld_wait();
          tmem_ld32(t_s + 96, r1);   // chunk 3 in flight while chunk 2 is processed
        }
      }
      }
      l += rowsum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(x, 1));
      if (++s == kStages) { s = 0; ph ^= 1u; }
    }
    // ---- normalise and store (the last commit covers every earlier P.V)
    mbar_wait(o_full(x, 1), (uint32_t)(n - 1) & 1u);
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
  p.ntiles = d.ntiles;
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

template <bool HB, bool TL>
static void attn_launch(const AttnLaunch& l, cudaStream_t st) {
  // PIPE = HAS_BIAS: same box (profiles/r3b_attn_pipe.txt) the continuous chunk pipeline is +12 % on the biased self attention
  // (656 vs 584 TFLOP/s at level 0) and -1 % on cross attention, which keeps the two-chunk form
  static PerDeviceOnce attr;  // function attributes are per device: a second GPU in the same process needs them too
  attr([] { SDM_CUDA_OK(cudaFuncSetAttribute(attention_kernel<HB, TL, HB>, cudaFuncAttributeMaxDynamicSharedMemorySize, a7::kSmem)); });
  attention_kernel<HB, TL, HB><<<l.grid, a7::kThreads, a7::kSmem, st>>>(l.p);
}

void attn_run(const AttnLaunch& l, cudaStream_t st) {
  if (!l.has_bias && (l.p.Lk & 127) != 0) attn_launch<false, true>(l, st);
  else if (l.has_bias) attn_launch<true, false>(l, st);
  else attn_launch<false, false>(l, st);
  SDM_CUDA_OK(cudaGetLastError());
}

}  // namespace sdm
