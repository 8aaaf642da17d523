// Swapped-operand 3x3 convolution with a RESIDENT PIXEL HALO TILE, optionally with the input's GroupNorm(+SiLU) fused in.
// (written in round 1, first run on hardware in round 2: tests/test_kernels_gpu.py -k swapped)
//
// conv_swap_kernel (channels on the MMA's M, 256 pixels on N) with a RESIDENT PIXEL HALO TILE: per 64-channel slice ONE TMA box
// of (8+2) x (32+2) pixels is fetched, and the nine taps are nine B-operand descriptors of that tile started (dy*10 + dx) pixel
// rows later with SBO = 10 rows (1280 B) — the mechanism tests/probe_halo.py verified for A operands (the 128-byte swizzle
// is a function of absolute shared-memory address bits).  conv_swap_kernel is bound by the L2 -> shared-memory fill (48 KB per
// K step, ~51 B/clk/SM measured); here the fill per slice is 43.5 KB + 9 x 16 KB of weights instead of 9 x 48 KB.
// Tile = 8 wide x 32 tall pixel patch (N = 256 = 32 groups of 8 rows) x 128 output channels; GroupNorm partials: two slots per
// patch (128 pixels each), as in conv_swap_kernel.
//
// GNF (round 2): the conv of a ResnetBlock2D consumes GroupNorm(32) -> SiLU of its input (reference: diffusers ResnetBlock2D, built at
// /root/reference/src/utils/replace.py:239,268,321).  As a separate pass that is one full HBM read + write of the activation per
// conv (27 ms of a 280 ms step).  With the halo tile an activation slice sits in shared memory exactly ONCE per 64 channels, so
// eight TRANSFORM warps apply y = silu(a x + s) in place (per-(sample, channel) (a, s) from gn_finalize_kernel; the arithmetic of
// gn_apply_kernel with MUFU.RCP instead of its Newton reciprocal) between the TMA's arrival (xfull) and the MMA's use (xready).  Rows outside the image
// were zero-filled by the TMA unit and stay zero: the convolution pads the NORMALISED tensor.  The 128-byte swizzle is undone
// per thread: 16-byte piece j of tile row r holds channel chunk j ^ (r & 7); a thread keeps (piece, r mod 16), hence one fixed
// chunk and its 16 constants in registers.
//
// Round 3: PAIR (tcgen05 cta_group::2 over 256 output channels, half a patch per CTA: see swh::Cfg), the polyphase form of
// "nearest x2 upsample -> 3x3 conv" (p.poly: four taps, strided output; diffusers Upsample2D as reached from the up blocks of
// /root/reference/src/utils/replace.py and the VAE decoder, meta_arch.py:255-256), residual K slices interleaved with the input
// slices (seq_item).  Evidence: profiles/r3c_kbench_mc_after_fix.txt, r3d (run_r3d.sh), r3f_*, r3g_ops.csv, r3z_*.
#include "gn_math.cuh"
#include "umma_gemm.cuh"

namespace sdm {

namespace swh {
constexpr int kWBytes = 128 * 128;    // weight tile: 128 channels x 64 k (fp16)
constexpr int kStgBytes = 4 * 2048;
constexpr int kBaseThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
// GNF: the slot of a halo tile is busy for (TMA flight + transform + nine taps of MMAs) instead of (TMA flight + MMAs): a third
// slot (paid for with one weight stage) keeps the tensor core fed; 8 transform warps (r2b: 4 warps with a branch per piece made
// the fused conv 35 % slower than conv + separate apply pass, i.e. no net gain)
// PAIR (r3d): the 8 x 32 patch x 256 output channels is computed by a CTA PAIR (cluster of two, tcgen05 cta_group::2, M = 256):
// CTA r holds the weights of channel tile 2g + r, the HALF patch of rows [16 r, 16 r + 16) (its (8+2) x (16+2) halo tile) and
// the 128 x 256 accumulator of its channels.  Why: the single-CTA kernel is bound by SHARED-MEMORY BANDWIDTH (128 B/clk/SM), not by
// the tensor pipe, L2 or the issuer: per M128 x N256 x K16 MMA (128 clk) the tensor core reads 12 KB of operands (96 B/clk), the
// TMA unit writes 41 B/clk (weights 32, halo tile 9.4), the GroupNorm transform reads + writes 19 and the epilogue staging 7:
// 163 B/clk -> 163 clk per MMA predicted, 156 measured net of all barrier waits (profiles/r3c_kbench_mc_after_fix.txt; the same
// sum explains the probe's 134-137 clk with TMA streams only and the nominal rate of the no-TMA bisect variant).  In a pair each
// SM reads 4 KB of A + its 4 KB half of B per MMA (64 B/clk) and fills / transforms half a halo tile: 118 B/clk.
template <bool GNF, bool PAIR = false>
struct Cfg {
  static constexpr int kTileRows = PAIR ? 16 : 32;              // patch rows whose pixels this CTA holds
  static constexpr int kXRows = 10 * (kTileRows + 2);           // pixel rows of a halo tile (340 / 180)
  static constexpr int kXSlot = PAIR ? 23 * 1024 : 44 * 1024;   // 43 520 / 23 040 B rounded up to the 1024-byte swizzle atom
  static constexpr int kXTx = kXRows * 128;                     // bytes one halo box delivers
  static constexpr int kXDense = 8 * kTileRows * 128;           // residual: dense 8 x 32 (8 x 16) box
  static constexpr int kPieces = (kXRows + 31) / 32;            // 16-byte pieces per transform thread (11 / 6)
  static constexpr int kWStages = PAIR ? 8 : (GNF ? 5 : 6);
  static constexpr int kXSlots = (GNF || PAIR) ? 3 : 2;  // (pair: half-size slots; with four taps per slice — polyphase form — two turn over too fast)
  static constexpr int kTWarps = GNF ? 8 : 0;
  static constexpr int kThreads = kBaseThreads + 32 * kTWarps;
  static constexpr int kPipe = kWStages * kWBytes + kXSlots * kXSlot;
  static constexpr int kSmem = kPipe + 1024 + 256 + kStgBytes;
  static_assert(kSmem <= 232448, "shared memory budget");
  static_assert(8 * (2 * kWStages + 3 * kXSlots + 4) + 8 <= 256, "barrier area");
};

// one thread's share of a halo tile: pieces (r0 + 32 k, piece), k = 0..10; `exist` bit k: the row is part of the tile,
// `inside` bit k: its pixel lies inside the image (rows outside stay / become zero: the conv pads the NORMALISED tensor).
// Branch-free and unrolled four pieces deep, so that 16 independent channel-pair chains are in flight per thread.
template <bool SILU, int NP>
__device__ __forceinline__ void gn_transform_tile(uint8_t* tp, uint32_t inside, uint32_t exist, const uint64_t (&ka)[4], const uint64_t (&ks)[4]) {
#pragma unroll
  for (int g = 0; g < (NP + 3) / 4; ++g) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = g * 4 + j;
      v[j] = make_uint4(0u, 0u, 0u, 0u);
      if (k < NP && ((exist >> k) & 1u)) v[j] = *reinterpret_cast<const uint4*>(tp + k * 4096);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = g * 4 + j;
      if (k < NP) {
        uint4 o = gn_piece<SILU, true>(v[j], ka, ks);
        if (!((inside >> k) & 1u)) o = make_uint4(0u, 0u, 0u, 0u);
        if ((exist >> k) & 1u) *reinterpret_cast<uint4*>(tp + k * 4096) = o;
      }
    }
  }
}
}  // namespace swh

template <bool GNF, bool PAIR>
__global__ void __launch_bounds__(swh::Cfg<GNF, PAIR>::kThreads, 1) conv_swap_halo_kernel(const __grid_constant__ ConvGemmParams p) {
  using namespace swh;
  using C = Cfg<GNF, PAIR>;
  constexpr int kWStages = C::kWStages, kXSlots = C::kXSlots, kPipe = C::kPipe, kXSlot = C::kXSlot, kXRows = C::kXRows;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // work items: (pixel tile, channel tile) per CTA, or (pixel tile, PAIR of channel tiles) per cluster; channel tiles fastest
  const int item_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_groups = PAIR ? (p.n_tiles >> 1) : p.n_tiles;
  const int item_count = p.m_tiles * n_groups;
  auto item_n0 = [&](int item) { return ((item % n_groups) * (PAIR ? 2 : 1) + (int)rank) * 128; };  // this CTA's first output channel
  auto item_mt = [&](int item) { return item / n_groups; };
  const int yoff = PAIR ? 16 * (int)rank : 0;  // first patch row of this CTA's half
  // polyphase form of "nearest x2 upsample -> 3x3 conv" (p.poly = 1 + 2 py + px): four taps (dy, dx) in {0,1}^2 whose windows start
  // (dy + py, dx + px) halo pixels into the tile; output pixel (y, x) of the grid is stored at (2y + py, 2x + px) of a (2H, 2W) image
  const int ppy = p.poly ? ((p.poly - 1) >> 1) : 0, ppx = p.poly ? ((p.poly - 1) & 1) : 0;
  const int ntaps = p.ntaps;  // 9, or 4 (polyphase)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t x_base = smem_base + kWStages * kWBytes;
  const uint32_t bar_base = smem_base + kPipe;
  auto wfull_bar = [&](int s) { return bar_base + 8u * s; };
  auto wempty_bar = [&](int s) { return bar_base + 8u * (kWStages + s); };
  auto xfull_bar = [&](int s) { return bar_base + 8u * (2 * kWStages + s); };
  auto xempty_bar = [&](int s) { return bar_base + 8u * (2 * kWStages + kXSlots + s); };
  auto xready_bar = [&](int s) { return bar_base + 8u * (2 * kWStages + 2 * kXSlots + s); };  // GNF: the slot's tile has been normalised
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kWStages + 3 * kXSlots + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kWStages + 3 * kXSlots + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kWStages + 3 * kXSlots + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(wfull_bar(s), 1); mbar_init(wempty_bar(s), 1); }
    for (int s = 0; s < kXSlots; ++s) { mbar_init(xfull_bar(s), 1); mbar_init(xempty_bar(s), 1); }
    // PAIR: the leader's MMA thread waits for the epilogue / transform warps of BOTH CTAs
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), PAIR ? 8 : 4); }
    if (GNF) for (int s = 0; s < kXSlots; ++s) mbar_init(xready_bar(s), (PAIR ? 2 : 1) * C::kTWarps);  // one arrive per transform warp
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // measurement aids (tests/bench_kernels.py "bisect" cases; the output is garbage with any of them set), bits of gn_silu:
  //   32: the halo tile always comes from the first patch of the tensor (L2-hot, no zero fill)
  //   128: no operand traffic at all: the producers stay idle, the MMA issuer and the transform warps do not wait for operands
  const bool dbg_hot_x = p.gn_silu & 32, dbg_no_tma = p.gn_silu & 128, dbg_no_wait = dbg_no_tma;
  const int nres = p.has_res ? (PAIR ? 4 : 2) : 0;  // residual K slices: 64 of the tile's (pair's) output channels each
  int nslices = 0;
  for (int s = 0; s < p.nsrc; ++s) nslices += p.src_c[s] >> 6;
  const int per_image = p.tiles_x * p.tiles_y;
  // K sequence of a tile: the residual K slices are INTERLEAVED with the first input slices (s0 r0 s1 r1 s2 s3 ...).  A residual box
  // feeds 4 MMAs, an input slice 36: behind the input slices (r3f and earlier) two of the three halo slots held residual boxes
  // while the last slice was consumed, the next tile's first slice could only be requested when that finished, and its flight +
  // GroupNorm transform (~3.5 k clk) was covered by 1 k clk of residual MMAs — 128 -> 128 convs with a residual ran at 760-860
  // TFLOP/s against 1150 without (r3f_ops.csv).  Interleaved, every input slice is requested a full slice ahead (r3g: 864 -> 948,
  // 761 -> 882).  Only for an ODD number of channel tiles (N = 128, 384, 640): with N % 256 == 0 the same conv runs as CTA pairs or
  // single CTAs depending on the batch size, a pair adds FOUR residual slices (two of them zeros for either CTA) and its interleaved
  // order would differ from the single CTA's — a sample must give the same bits alone and in a batch, so those keep the residual last.
  const int nseq = nslices + nres, nmix = ((p.n_tiles & 1) && p.res_mix) ? min(nslices, nres) : 0;
  auto seq_item = [&](int q, bool& resid) {  // q-th K item of a tile -> (residual slice i | input slice sl)
    if (q < 2 * nmix) { resid = (q & 1) != 0; return q >> 1; }
    const int r = q - 2 * nmix, rem = nslices - nmix;  // behind the interleaved prefix: the other input slices, then the other residual slices
    resid = r >= rem;
    return nmix + (resid ? r - rem : r);
  };

  if (warp == 0) {
    // ============================== TMA producers ==============================
    // lane 0 streams the weight tiles, lane 1 the pixel halo tiles, each throttled by its own ring only.  One in-order producer
    // (round 1) issued halo tile i+1 behind the nine weight taps of slice i, i.e. about half a slice ahead of its use: enough
    // when the MMA reads the tile as it lands, but with the GroupNorm transform in between (r2g: T + M serialised, the fused conv
    // 30 % slower than the barrier hop alone) the tile has to arrive a full slice earlier.
    if (lane == 0 && !dbg_no_tma) {
      tma_prefetch_desc(&p.b_map);
      int ws = 0;
      uint32_t wph = 0;
      // PAIR: every transaction byte of both CTAs' weight tiles is accounted on the LEADER's full barrier
      auto load_w = [&](const CUtensorMap* map, int c0, int c1) {
        mbar_wait(wempty_bar(ws), wph ^ 1u);
        if constexpr (PAIR) {
          if (rank == 0) mbar_expect_tx(wfull_bar(ws), 2 * kWBytes);
          tma_load_2d_pair(smem_base + ws * kWBytes, map, mapa_shared(wfull_bar(ws), 0), c0, c1);
        } else {
          mbar_expect_tx(wfull_bar(ws), kWBytes);
          tma_load_2d(smem_base + ws * kWBytes, map, wfull_bar(ws), c0, c1);
        }
        if (++ws == kWStages) { ws = 0; wph ^= 1u; }
      };
      for (int item = item_first; item < item_count; item += item_step) {
        const int n0 = item_n0(item);
        for (int q = 0; q < nseq; ++q) {
          bool resid;
          const int i = seq_item(q, resid);
          if (!resid) {  // input slice i = K columns [64 i, 64 i + 64) of every tap
            for (int tap = 0; tap < ntaps; ++tap) load_w(&p.b_map, tap * p.cin_total + i * 64, n0);
          } else {  // D^T[c][pix] += I[c][64 i + k] . R[pix][g0 + 64 i + k], g0 = first channel of the tile (pair): rows 128 rank .. of the identity
            load_w(&p.i_map, 64 * i, 128 * (int)rank);
          }
        }
      }
    } else if (lane == 1 && !dbg_no_tma) {
      tma_prefetch_desc(&p.a_map[0]); tma_prefetch_desc(&p.a_map[1]);
      int xs = 0;
      uint32_t xph = 0;
      for (int item = item_first; item < item_count; item += item_step) {
        const int g0 = item_n0(item) - 128 * (int)rank;  // first output channel of the tile (pair)
        const int mt = item_mt(item);
        const int t_img = mt % per_image;
        const int x0 = (t_img % p.tiles_x) * 8, y0 = (t_img / p.tiles_x) * 32 + yoff, b = mt / per_image;
        const int s0_slices = p.src_c[0] >> 6;  // input slices of the first source (concatenated inputs: the second follows)
        for (int q = 0; q < nseq; ++q) {
          bool resid;
          const int i = seq_item(q, resid);
          mbar_wait(xempty_bar(xs), xph ^ 1u);
          if (!resid) {
            const int s = i < s0_slices ? 0 : 1, c0 = (i < s0_slices ? i : i - s0_slices) * 64;
            if constexpr (PAIR && !GNF) {  // no transform warps in between: the leader's MMA thread waits for both halves directly
              if (rank == 0) mbar_expect_tx(xfull_bar(xs), 2 * C::kXTx);
              tma_load_4d_pair(x_base + xs * kXSlot, &p.a_map[s], mapa_shared(xfull_bar(xs), 0), c0, x0 - 1, y0 - 1, b);
            } else {
              mbar_expect_tx(xfull_bar(xs), C::kXTx);
              if (dbg_hot_x) tma_load_4d(x_base + xs * kXSlot, &p.a_map[s], xfull_bar(xs), c0, 0, 0, 0);
              else tma_load_4d(x_base + xs * kXSlot, &p.a_map[s], xfull_bar(xs), c0, x0 - 1, y0 - 1, b);  // zero fill = conv padding
            }
          } else {  // dense residual box in a halo slot
            if constexpr (PAIR && !GNF) {
              if (rank == 0) mbar_expect_tx(xfull_bar(xs), 2 * C::kXDense);
              tma_load_4d_pair(x_base + xs * kXSlot, &p.r_map, mapa_shared(xfull_bar(xs), 0), g0 + 64 * i, x0, y0, b);
            } else {
              mbar_expect_tx(xfull_bar(xs), C::kXDense);
              tma_load_4d(x_base + xs * kXSlot, &p.r_map, xfull_bar(xs), g0 + 64 * i, x0, y0, b);
            }
          }
          if (++xs == kXSlots) { xs = 0; xph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (PAIR: the leader CTA's only) ==============================
    // The WHOLE warp runs this loop convergently and one elected lane issues the tcgen05 instructions.  With the loop inside an
    // `if (lane == 0)` (rounds 1 / 2 up to r2u) the descriptors lived in per-thread registers of a divergent region and the
    // compiler wrapped EVERY tcgen05.mma in an ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY loop (SASS): ~17 dependent instructions per
    // MMA on top of the per-tap wait, fence and descriptor arithmetic — the issuing thread, not the tensor core or the operand
    // supply, was the limit (r2u: with the operand waits removed the same kernel ran 1.3x faster; tests/probe_mma_rate.cu).
    if (!PAIR || rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc = PAIR ? umma_idesc_f16_m256(256) : umma_idesc_f16(256);
      int ws = 0, xs = 0, acc = 0;
      uint32_t wph = 0, xph = 0, acc_phase = 0;
      // SDM_GEMM_PROF=1 (measurement aid): cycles the issuing thread waited for a free accumulator / a (normalised) pixel tile /
      // a weight tile
      long long pw_acc = 0, pw_x = 0, pw_w = 0;
      const long long prof_t0 = p.prof ? clock64() : 0;
      auto timed_wait = [&](uint32_t bar, uint32_t ph, long long& accum) {
        const long long t = p.prof ? clock64() : 0;
        if constexpr (PAIR) mbar_wait_cluster(bar, ph);  // arrivals (and the shared-memory writes behind them) come from the peer CTA too
        else mbar_wait(bar, ph);
        if (p.prof) accum += clock64() - t;
      };
      for (int item = item_first; item < item_count; item += item_step) {
        timed_wait(tempty_bar(acc), acc_phase ^ 1u, pw_acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int sl = 0; sl < nseq; ++sl) {  // sl = position in the tile's K sequence (seq_item)
          if (!dbg_no_wait) timed_wait(GNF ? xready_bar(xs) : xfull_bar(xs), xph, pw_x);
          const uint32_t x_addr = x_base + xs * kXSlot;
          bool resid;
          (void)seq_item(sl, resid);
          const int ntap = resid ? 1 : ntaps;
          for (int tap = 0; tap < ntap; ++tap) {
            if (!dbg_no_wait) timed_wait(wfull_bar(ws), wph, pw_w);
            tc_fence_after();
            const uint64_t adesc = umma_desc_k128(smem_base + ws * kWBytes);
            const int win = p.poly ? ((tap >> 1) + ppy) * 10 + (tap & 1) + ppx : (tap / 3) * 10 + tap % 3;  // first halo pixel of the tap's window
            const uint64_t bdesc = resid ? umma_desc_k128(x_addr) : umma_desc_k128_sbo(x_addr + (uint32_t)win * 128u, 1280);
            if (leader) {
              if constexpr (PAIR) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (sl | tap | k) != 0);
                umma_commit_pair(wempty_bar(ws));  // frees the stage in both CTAs
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (sl | tap | k) != 0);
                umma_commit(wempty_bar(ws));
              }
            }
            __syncwarp();
            if (++ws == kWStages) { ws = 0; wph ^= 1u; }
          }
          if (leader) { if constexpr (PAIR) umma_commit_pair(xempty_bar(xs)); else umma_commit(xempty_bar(xs)); }
          __syncwarp();
          if (++xs == kXSlots) { xs = 0; xph ^= 1u; }
        }
        if (leader) { if constexpr (PAIR) umma_commit_pair(tfull_bar(acc)); else umma_commit(tfull_bar(acc)); }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      if (p.prof && blockIdx.x < 2 && leader)
        printf("sdm prof: swap_halo cta %d MMA issuer total %lld clk, waiting: accumulator %lld, pixel tile %lld, weight tile %lld\n", blockIdx.x,
               clock64() - prof_t0, pw_acc, pw_x, pw_w);
    }
  } else if (GNF && warp >= 6) {
    // ============================== GroupNorm transform (8 warps) ==============================
    const int t = threadIdx.x - kBaseThreads;  // 0..255
    const int piece = t & 7, r0 = t >> 3;      // rows r0, r0 + 32, ... of the tile; r & 7 == r0 & 7 for all of them
    const int chunk = piece ^ (r0 & 7);        // channel chunk (8 channels) this thread's pieces hold
    constexpr int NP = C::kPieces;
    const uint32_t exist = r0 < kXRows - 32 * (NP - 1) ? (1u << NP) - 1u : (1u << (NP - 1)) - 1u;  // the last piece row exists for r0 < 20
    auto publish = [&](int slot) {  // this warp's share of the slot's tile is final (PAIR: tell the leader CTA's MMA thread)
      if constexpr (PAIR) mbar_arrive_cluster_release(mapa_shared(xready_bar(slot), 0));
      else mbar_arrive(xready_bar(slot));
    };
    int xs = 0;
    uint32_t xph = 0;
    // (scale, shift) of this thread's 8 channels, one slice ahead: ncu r2f — fetched at the top of every slice the constants cost
    // one exposed L2 round trip per slice (30 % of the kernel's stall samples sat on their first use)
    auto load_raw = [&](int item_, int slice, float4 (&raw)[4]) {
      const int b_ = item_mt(item_) / per_image;
      const float4* src = reinterpret_cast<const float4*>(p.gn_ab + ((size_t)b_ * p.cin_total + slice * 64 + chunk * 8) * 2);
#pragma unroll
      for (int j = 0; j < 4; ++j) raw[j] = __ldg(src + j);
    };
    float4 raw[4];
    if (item_first < item_count) load_raw(item_first, 0, raw);
    for (int item = item_first; item < item_count; item += item_step) {
      const int mt = item_mt(item);
      const int t_img = mt % per_image;
      const int x0 = (t_img % p.tiles_x) * 8, y0 = (t_img / p.tiles_x) * 32 + yoff;
      // validity of this thread's rows depends on the tile position only: bit k = row r0 + 32 k lies inside the image
      uint32_t inside = 0;
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int r = r0 + 32 * k;
        const int px = x0 - 1 + r % 10, py = y0 - 1 + r / 10;
        if (px >= 0 && px < p.W && py >= 0 && py < p.H) inside |= 1u << k;
      }
      inside &= exist;
      for (int q = 0; q < nseq; ++q) {
        bool resid;
        const int sl = seq_item(q, resid);  // input slice sl = channels [64 sl, 64 sl + 64) of the concatenated input
        if (resid) {  // residual boxes pass through untouched
          mbar_wait(xfull_bar(xs), xph);
          __syncwarp();
          if (lane == 0) publish(xs);
          if (++xs == kXSlots) { xs = 0; xph ^= 1u; }
          continue;
        }
        uint64_t ka[4], ks[4];
        if (p.gn_silu == 0) gn_consts_from_raw<false>(raw, ka, ks);
        else gn_consts_from_raw<true>(raw, ka, ks);
        // next slice's constants (or the first slice of this CTA's next tile) fly during the transform
        if (sl + 1 < nslices) load_raw(item, sl + 1, raw);
        else if (item + item_step < item_count) load_raw(item + item_step, 0, raw);
        if (!dbg_no_tma) mbar_wait(xfull_bar(xs), xph);
        uint8_t* tp = smem_raw + (x_base + xs * kXSlot - smem_u32(smem_raw)) + r0 * 128 + piece * 16;
        if (p.gn_silu == 1) gn_transform_tile<true, NP>(tp, inside, exist, ka, ks);
        else if (p.gn_silu == 0) gn_transform_tile<false, NP>(tp, inside, exist, ka, ks);
        else if (p.gn_silu == 4) {  // measurement aid (tests/bench_kernels.py): shared-memory traffic of the transform without its math
#pragma unroll
          for (int k = 0; k < NP; ++k)
            if ((exist >> k) & 1u) { uint4* q = reinterpret_cast<uint4*>(tp + k * 4096); uint4 v = *q; v.x ^= inside; *q = v; }
        }  // gn_silu == 2 (measurement aid): no transform at all, only the extra barrier hop
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane == 0) publish(xs);
        if (++xs == kXSlots) { xs = 0; xph ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue: thread = output channel; column j = pixel (y = j / 8, x = j % 8) ==============
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    uint8_t* stg = smem_raw + (bar_base - smem_u32(smem_raw)) + 256 + (warp - 2) * 2048;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = item_first; item < item_count; item += item_step) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int n0 = item_n0(item);
      const int mt = item_mt(item);
      const int t_img = mt % per_image;
      const int x0 = (t_img % p.tiles_x) * 8, y0 = (t_img / p.tiles_x) * 32, b = mt / per_image;
      const float bias = p.bias ? p.bias[(p.bias_sel ? (long long)p.bias_sel[b] * p.N : 0) + n0 + m] : 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * 256;
      __half* obase = reinterpret_cast<__half*>(p.out) + (long long)b * p.out_bstride + n0 + quad * 32;
      const int osc = p.poly ? 2 : 1;  // polyphase: this launch writes the pixels (2y + ppy, 2x + ppx) of the (2H, 2W) output
      float sum = 0.f, sq = 0.f;
      uint32_t ra[32], rb[32];
      // one block = 32 pixels = patch rows 4 blk .. 4 blk + 3, held in r; `nxt` receives the following block meanwhile
      auto block = [&](const uint32_t (&r)[32], uint32_t (&nxt)[32], int blk) {
        tmem_ld_wait();
        __syncwarp();
        if (blk + 1 < 8) tmem_ld32(taddr + (blk + 1) * 32, nxt);
        const int ya = y0 + 4 * blk;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const __half h = __float2half_rn(fmaf(__uint_as_float(r[i]), p.scale, bias));
          const bool ok = (x0 + (i & 7) < p.W) && (ya + (i >> 3) < p.H);
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
          const int xx = x0 + (px & 7), yy = ya + (px >> 3);
          const uint4 v = *reinterpret_cast<const uint4*>(stg + px * 64 + piece * 16);
          if (xx < p.W && yy < p.H && !(p.gn_silu & 16))  // (bit 4: measurement aid — no global stores)
            *reinterpret_cast<uint4*>(obase + ((long long)(yy * osc + ppy) * (p.W * osc) + xx * osc + ppx) * p.out_ld + piece * 8) = v;
        }
        if (p.stats && (blk & 3) == 3) {
          const long long slot = (long long)b * p.stats_bslots + p.stats_slot0 + 2 * t_img + (blk >> 2);
          *reinterpret_cast<float2*>(p.stats + (slot * p.N + n0 + m) * 2) = make_float2(sum, sq);
          sum = 0.f;
          sq = 0.f;
        }
      };
      __syncwarp();
      if (!(p.gn_silu & 8)) {  // (bit 3 of gn_silu: measurement aid of tests/bench_kernels.py — the tile's epilogue is skipped)
        tmem_ld32(taddr, ra);
#pragma unroll 1
        for (int b2 = 0; b2 < 4; ++b2) {
          block(ra, rb, 2 * b2);
          block(rb, ra, 2 * b2 + 1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));  // the leader's MMA thread waits for both CTAs
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA frees tensor memory (or exits) while the pair's MMAs / remote arrives are in flight
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

template <bool GNF, bool PAIR>
static void swap_halo_launch_t(const ConvGemmParams& p, int grid, cudaStream_t st) {
  using C = swh::Cfg<GNF, PAIR>;
  static PerDeviceOnce attr;
  attr([] { SDM_CUDA_OK(cudaFuncSetAttribute(conv_swap_halo_kernel<GNF, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem)); });
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SDM_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_swap_halo_kernel<GNF, PAIR>, p));
}

// pair: CTA pairs (cta_group::2) over 256 output channels; needs N % 256 == 0, an even grid and the half-patch tensor-map
// boxes (conv_gemm_build)
void conv_swap_halo_launch(const ConvGemmParams& p, int grid, bool pair, cudaStream_t st) {
  if (pair) {
    SDM_CHECK((p.n_tiles & 1) == 0 && (grid & 1) == 0, "pair launch preconditions");
    if (p.gn_ab) swap_halo_launch_t<true, true>(p, grid, st);
    else swap_halo_launch_t<false, true>(p, grid, st);
  } else if (p.gn_ab) {
    swap_halo_launch_t<true, false>(p, grid, st);
  } else {
    swap_halo_launch_t<false, false>(p, grid, st);
  }
}

}  // namespace sdm
