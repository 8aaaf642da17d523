// Host side of the tcgen05 implicit-GEMM kernel: tensor-map construction, tile-shape selection, launch.
#include "kernels.h"
#include "umma_launch.h"
#include "tmap.h"

#include <cstdlib>
#include <map>

#include <mutex>

namespace sdm {

void conv_swap_launch(const ConvGemmParams& p, int grid, cudaStream_t st);       // conv_swap.cu
void conv_swap_halo_launch(const ConvGemmParams& p, int grid, bool pair, cudaStream_t st);  // conv_swap_halo.cu

struct ConvGemmLaunch {
  ConvGemmParams p;
  int block_n = 0;
  int mt = 1;
  int ewg = 1;
  bool halo = false;  // 3x3 stride-1 conv with a resident halo tile per 64-channel slice
  bool swap = false;  // conv_swap_kernel: channels on M, 256 pixels on N (128-channel 3x3 convs)
  bool swap_halo = false;  // ... with a resident 8 x 32 pixel halo tile (conv_swap_halo.cu; carries the fused GroupNorm)
  bool swap_pair = false;  // ... computed by CTA pairs (cta_group::2) over 256 output channels, half a patch per CTA
  int grid = 0;
  double flops = 0;
};

static void pick_patch(int H, int W, int& tw, int& th) {
  // 128-pixel patch tw x th minimising wasted (out-of-image) pixels; prefer long rows on ties
  long long best = -1;
  for (int cand = 128; cand >= 8; cand >>= 1) {
    const int ch = 128 / cand;
    if (ch > 1 && H == 1) continue;
    const long long tiles = (long long)((W + cand - 1) / cand) * ((H + ch - 1) / ch);
    if (best < 0 || tiles < best) {
      best = tiles;
      tw = cand;
      th = ch;
    }
  }
}

static int pick_block_n(int N, int mode, long long m_tiles, int num_sms) {
  if (mode == EPI_GEGLU) return 256;
  if (N <= 16) return 16;
  if (mode == EPI_F32) return (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  int bn;
  if (N % 256 == 0) bn = 256;
  else if (N % 160 == 0) bn = 160;
  else if (N % 128 == 0) bn = 128;
  else if (N > 128) bn = (N % 64 == 0) ? 64 : 128;
  else bn = (N > 64) ? 128 : 64;
  // small problems: trade tile size for parallelism (one wave should cover the SMs)
  while (bn == 256 && m_tiles * (N / bn) < num_sms && N % 128 == 0) bn = 128;
  if (bn == 128 && m_tiles * (N / bn) < num_sms && N % 64 == 0) bn = 64;
  return bn;
}

// identity matrix [kIdentityN][kIdentityN] fp16, one per device: B operand of the residual K steps
constexpr int kIdentityN = 2048;
__global__ void identity_init_kernel(__half* m, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) m[(size_t)i * n + i] = __float2half_rn(1.0f);
}
static const __half* identity_matrix() {
  static std::mutex mu;
  static std::map<int, __half*> per_device;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  SDM_CUDA_OK(cudaGetDevice(&dev));
  auto it = per_device.find(dev);
  if (it != per_device.end()) return it->second;
  __half* m = nullptr;
  SDM_CUDA_OK(cudaMalloc(&m, (size_t)kIdentityN * kIdentityN * 2));
  SDM_CUDA_OK(cudaMemset(m, 0, (size_t)kIdentityN * kIdentityN * 2));
  identity_init_kernel<<<(kIdentityN + 255) / 256, 256>>>(m, kIdentityN);
  SDM_CUDA_OK(cudaGetLastError());
  SDM_CUDA_OK(cudaDeviceSynchronize());
  per_device[dev] = m;
  return m;
}

// Kernel variant of a 3x3 stride-1 conv, decided from the per-sample GEOMETRY only: neither the batch size nor the SM count is
// an argument.  The variants differ in K order (halo: slice-major with the taps inner; the others tap-major) and in how pixels
// are grouped into GroupNorm-partials slots, and a sample must produce the same bits alone and inside a batch
// (test_batch_independence_and_determinism on the GPU, test_conv_variant_is_a_function_of_geometry in the CPU suite).
struct ConvVariant {
  bool swap_can = false, swap = false;      // conv_swap_kernel possible / selected
  bool swap_halo_geom = false, swap_halo = false;  // its resident-halo form (8 x 32 patches) possible / selected
  bool halo_geom = false, halo_auto = false;  // resident-halo patches fit / selected without a force flag
  int halo_bn = 0;                          // tile width of the halo kernel by N alone (256 | 160 | 0 = none)
};
static ConvVariant conv3x3_variant(int ksize, int stride, int mode, int ups2, bool batched_w, int N, int n_store, bool has_res,
                                   int Hout, int Wout, int force_swap, int force_halo, int force_block_n, int force_mt,
                                   bool want_gn = false) {
  ConvVariant v;
  const bool conv3 = ksize == 3 && stride == 1 && mode == EPI_F16 && (n_store == 0 || n_store == N);
  int tw = 128, th = 1;
  pick_patch(Hout, Wout, tw, th);
  const long long tiles_img = (long long)((Wout + tw - 1) / tw) * ((Hout + th - 1) / th);  // default patch, per sample
  // swapped operands (conv_swap.cu; force_swap = 1 / 2 / -1 from the tests): 16 x 16 pixel patches, two
  // GroupNorm-partials slots per patch -> only where that equals the default slot count
  v.swap_can = conv3 && !ups2 && !batched_w && N % 128 == 0 && Hout >= 16 && Wout >= 16 &&
               2ll * ((Wout + 15) / 16) * ((Hout + 15) / 16) == tiles_img && (!has_res || N <= kIdentityN);
  // resident-halo form: 8 x 32 pixel patches, two GroupNorm-partials slots per patch -> only where that equals the default count
  v.swap_halo_geom = v.swap_can && Hout >= 32 && Wout >= 8 && 2ll * ((Wout + 7) / 8) * ((Hout + 31) / 32) == tiles_img;
  v.swap = v.swap_can && force_halo != 1 &&
           (force_swap >= 1 || (want_gn && v.swap_halo_geom) ||
            (force_swap == 0 && N == 128 && force_block_n == 0 && force_mt == 0));
  // the resident-halo form wherever its patches fit (kbench r2h: 895 vs 823 TFLOP/s on the 128 -> 128 conv at 1024^2, and it
  // carries the fused GroupNorm); force_swap = 1 keeps the one-box-per-tap form for the kernel tests
  v.swap_halo = v.swap && v.swap_halo_geom && force_swap != 1;
  // resident halo tile (force_halo = 1 / -1 from the tests): 8 x 16 pixel patches, used where that
  // gives the default number of M tiles (the slot count conv_gemm_tiles_per_image must not depend on the kernel choice).
  // Tile width by N alone: 256 | 160 (the small-problem narrowing of pick_block_n depends on the batch size).
  v.halo_geom = conv3 && !v.swap && Hout >= 16 && Wout >= 8 && (long long)((Wout + 7) / 8) * ((Hout + 15) / 16) == tiles_img;
  v.halo_bn = N % 256 == 0 ? 256 : (N % 160 == 0 ? 160 : 0);
  v.halo_auto = v.halo_geom && v.halo_bn != 0 && force_halo == 0 && force_block_n == 0 && force_mt == 0;
  return v;
}
bool conv_gemm_can_poly(int N, int Hin, int Win) {
  static const int on = [] { const char* e = getenv("SDM_CONV_POLY"); return e ? atoi(e) : 1; }();  // A/B switch
  if (!on) return false;
  const ConvVariant v = conv3x3_variant(3, 1, EPI_F16, 0, false, N, 0, false, Hin, Win, 2, 0, 0, 0, false);
  return v.swap && v.swap_halo;
}
// 0 = one TMA box per tap (conv_gemm_kernel), 1 / 2 = resident halo with 256- / 160-wide tiles, 3 = swapped operands
// GroupNorm fusion is available where the resident-halo swapped-operand kernel applies.  Every 128-channel N tile of a pixel tile
// re-normalises the same input slice, so the transform work grows with N / 128 while the apply pass it replaces does not.
// SDM_GN_FUSE (A/B switch of round 2): 0 = never, 1 = N == 128 only, 2 = every N % 128 == 0 conv, 3 = N <= 256
bool conv_gemm_can_fuse_gn(int ksize, int stride, int mode, int ups2, int N, int has_res, int Hout, int Wout) {
  static const int level = [] { const char* e = getenv("SDM_GN_FUSE"); return e ? atoi(e) : 2; }();
  if (level <= 0 || (level == 1 && N != 128) || (level == 3 && N > 256)) return false;
  const ConvVariant v = conv3x3_variant(ksize, stride, mode, ups2, false, N, 0, has_res != 0, Hout, Wout, 0, 0, 0, 0, true);
  return v.swap && v.swap_halo;
}
int conv_gemm_variant_code(int ksize, int stride, int mode, int ups2, int N, int has_res, int Hout, int Wout) {
  const ConvVariant v = conv3x3_variant(ksize, stride, mode, ups2, false, N, 0, has_res != 0, Hout, Wout, 0, 0, 0, 0);
  return v.swap ? 3 : (v.halo_auto ? (v.halo_bn == 256 ? 1 : 2) : 0);
}

std::shared_ptr<ConvGemmLaunch> conv_gemm_build(const ConvGemmDesc& d, int num_sms) {
  auto L = std::make_shared<ConvGemmLaunch>();
  ConvGemmParams& p = L->p;
  memset(&p, 0, sizeof(p));
  SDM_CHECK(d.ksize == 1 || d.ksize == 3, "ksize");
  SDM_CHECK(d.stride == 1 || (d.stride == 2 && d.ksize == 3 && d.nsrc == 1), "stride");
  SDM_CHECK(d.nsrc == 1 || d.nsrc == 2, "nsrc");
  SDM_CHECK(d.N % 8 == 0, "N must be a multiple of 8");
  int cin_total = 0;
  for (int s = 0; s < d.nsrc; ++s) {
    SDM_CHECK(d.src[s].C % 64 == 0 && d.src[s].C > 0, "source channels must be a multiple of 64");
    SDM_CHECK(d.src[s].ld % 8 == 0, "pixel stride must be a multiple of 8 elements");
    cin_total += d.src[s].C;
  }
  const int Hout = d.Hin / d.stride, Wout = d.Win / d.stride;
  if (d.stride == 2) SDM_CHECK(d.Hin % 2 == 0 && d.Win % 2 == 0, "stride-2 conv needs even input dims");
  p.B = d.B; p.H = Hout; p.W = Wout;
  pick_patch(Hout, Wout, p.tw, p.th);
  p.tiles_x = (Wout + p.tw - 1) / p.tw;
  p.tiles_y = (Hout + p.th - 1) / p.th;
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * d.B;
  p.N = d.N;
  // ---- kernel variant of the 3x3 stride-1 convs: conv3x3_variant() sees the per-sample geometry only
  if (d.poly) {
    SDM_CHECK(d.poly >= 1 && d.poly <= 4 && d.ksize == 3 && d.stride == 1 && d.mode == EPI_F16 && !d.ups2 && !d.res && d.nsrc == 1 && !d.gn_ab &&
                  !d.w_bstride && (d.n_store == 0 || d.n_store == d.N),
              "polyphase conv: plain 3x3-geometry stride-1 fp16 conv of one source only");
    SDM_CHECK(conv_gemm_can_poly(d.N, Hout, Wout), "polyphase conv: geometry not supported (conv_gemm_can_poly)");
  }
  const ConvVariant cv = conv3x3_variant(d.ksize, d.stride, d.mode, d.ups2, d.w_bstride != 0, d.N, d.n_store, d.res != nullptr, Hout, Wout,
                                         d.poly ? 2 : d.force_swap, d.force_halo, d.force_block_n, d.force_mt, d.gn_ab != nullptr);
  if (d.poly) SDM_CHECK(cv.swap && cv.swap_halo, "polyphase conv needs the resident-halo swapped-operand kernel");
  if (d.force_swap >= 1) SDM_CHECK(cv.swap_can, "force_swap: configuration not supported by the swapped-operand kernel");
  if (d.force_swap == 2) SDM_CHECK(cv.swap_halo, "force_swap = 2: geometry not supported by the resident-halo swapped-operand kernel");
  if (d.gn_ab) SDM_CHECK(cv.swap && cv.swap_halo, "fused GroupNorm needs the resident-halo swapped-operand kernel (conv_gemm_can_fuse_gn)");
  L->swap = cv.swap;
  const bool halo_geom = cv.halo_geom, halo_auto = cv.halo_auto;
  const int halo_bn = cv.halo_bn;
  const int bn = halo_auto ? halo_bn : (d.force_block_n ? d.force_block_n : pick_block_n(d.N, d.mode, m_tiles, num_sms));
  if (d.mode == EPI_GEGLU) SDM_CHECK(bn == 256 && d.N % 256 == 0, "GEGLU needs N % 256 == 0");
  L->block_n = bn;
  p.n_tiles = (d.N + bn - 1) / bn;
  // 256 x 128 CTA tiles (two M sub-tiles per B tile) when there is enough work to keep every SM busy
  L->mt = (bn == 128 && d.mode == EPI_F16 && d.force_mt != 1 && (d.force_mt == 2 || (m_tiles / 2) * p.n_tiles >= 2 * num_sms)) ? 2 : 1;
  {
    const bool can = halo_geom && (bn == 256 || bn == 160 || (bn == 128 && L->mt == 2));
    if (d.force_halo == 1) SDM_CHECK(can, "force_halo: configuration not supported by the halo kernel");
    L->halo = can && (d.force_halo == 1 || halo_auto);
    if (L->halo) {
      p.tw = 8; p.th = 16;
      p.tiles_x = (Wout + 7) / 8;
      p.tiles_y = (Hout + 15) / 16;
    }
  }
  long long m_tiles_eff = m_tiles;
  if (L->swap) {
    L->swap_halo = cv.swap_halo;  // resident 8 x 32 halo tile (conv_swap_halo.cu)
    p.tw = L->swap_halo ? 8 : 16;
    p.th = L->swap_halo ? 32 : 16;
    p.tiles_x = (Wout + p.tw - 1) / p.tw;
    p.tiles_y = (Hout + p.th - 1) / p.th;
    m_tiles_eff = (long long)p.tiles_x * p.tiles_y * d.B;
    L->mt = 1;
    L->block_n = 128;
    p.n_tiles = d.N / 128;
    // CTA pairs: fused-GroupNorm form, an even number of channel tiles, at least one item per cluster (SDM_SWH_PAIR=0: A/B switch)
    static const int env_pair = [] { const char* e = getenv("SDM_SWH_PAIR"); return e ? atoi(e) : 1; }();
    L->swap_pair = L->swap_halo && d.N % 256 == 0 && env_pair != 0 && num_sms % 2 == 0 &&
                   m_tiles_eff * (p.n_tiles / 2) >= num_sms / 2;
  }
  p.m_tiles = (int)m_tiles_eff;
  const long long total = ((m_tiles_eff + L->mt - 1) / L->mt) * p.n_tiles;
  SDM_CHECK(total < (1ll << 31), "too many tiles");
  p.total_tiles = (int)total;
  p.ntaps = d.poly ? 4 : d.ksize * d.ksize;  // polyphase: taps (dy, dx) in {0,1}^2, weights [N][4][cin]
  p.poly = d.poly;
  {
    static const int env_mix = [] { const char* e = getenv("SDM_SWH_MIX"); return e ? atoi(e) : 1; }();
    p.res_mix = env_mix;
  }
  p.nsrc = d.nsrc;
  p.cin_total = cin_total;
  for (int s = 0; s < d.nsrc; ++s) p.src_c[s] = d.src[s].C;

  // ---- A tensor maps
  const uint32_t box[4] = {64u, (uint32_t)p.tw, (uint32_t)p.th, 1u};
  if (d.stride == 1) {
    for (int s = 0; s < d.nsrc; ++s) {
      const long long bs = d.src_bstride[s] ? d.src_bstride[s] : (long long)d.Hin * d.Win * d.src[s].ld;
      const uint64_t dims[4] = {(uint64_t)d.src[s].C, (uint64_t)d.Win, (uint64_t)d.Hin, (uint64_t)d.B};
      const uint64_t strides[3] = {(uint64_t)d.src[s].ld * 2, (uint64_t)d.Win * d.src[s].ld * 2, (uint64_t)bs * 2};
      // halo: (8+2) x (16+2) pixels, loaded at (x0-1, y0-1); swapped kernel: (8+2) x (32+2), or per CTA of a pair (8+2) x (16+2)
      const uint32_t hbox[4] = {64u, (uint32_t)p.tw + 2u, (uint32_t)(L->swap_pair ? p.th / 2 : p.th) + 2u, 1u};
      make_tmap(&p.a_map[s], d.src[s].ptr, 4, dims, strides, (L->halo || L->swap_halo) ? hbox : box);
    }
    for (int s = d.nsrc; s < 4; ++s) p.a_map[s] = p.a_map[0];
    for (int t = 0; t < p.ntaps; ++t) {
      p.tap_map[t] = 0;
      p.tap_dx[t] = (d.ksize == 3) ? (t % 3) - 1 : 0;
      p.tap_dy[t] = (d.ksize == 3) ? (t / 3) - 1 : 0;
    }
  } else {
    // stride 2: four parity views of the input (py, px); each is a plain strided 4-D tensor
    const long long ld = d.src[0].ld;
    const long long bs = d.src_bstride[0] ? d.src_bstride[0] : (long long)d.Hin * d.Win * ld;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const uint64_t dims[4] = {(uint64_t)d.src[0].C, (uint64_t)(d.Win / 2), (uint64_t)(d.Hin / 2), (uint64_t)d.B};
        const uint64_t strides[3] = {(uint64_t)ld * 2 * 2, (uint64_t)d.Win * ld * 2 * 2, (uint64_t)bs * 2};
        make_tmap(&p.a_map[py * 2 + px], d.src[0].ptr + ((long long)py * d.Win + px) * ld, 4, dims, strides, box);
      }
    for (int t = 0; t < 9; ++t) {
      const int ky = t / 3, kx = t % 3;
      int py, dy, px, dx;
      if (d.pad == PAD_SAME) {  // iy = 2*oy + ky - 1
        py = (ky == 1) ? 0 : 1; dy = (ky == 0) ? -1 : 0;
        px = (kx == 1) ? 0 : 1; dx = (kx == 0) ? -1 : 0;
      } else {  // iy = 2*oy + ky, zero row/col past the bottom/right edge
        py = (ky == 1) ? 1 : 0; dy = (ky == 2) ? 1 : 0;
        px = (kx == 1) ? 1 : 0; dx = (kx == 2) ? 1 : 0;
      }
      p.tap_map[t] = (signed char)(py * 2 + px);
      p.tap_dx[t] = (signed char)dx;
      p.tap_dy[t] = (signed char)dy;
    }
  }
  // ---- B tensor map
  const long long ktot = (long long)p.ntaps * cin_total;
  if (d.w_bstride) {
    SDM_CHECK(d.w_bstride % 8 == 0, "weight batch stride alignment");
    const uint64_t dims[3] = {(uint64_t)ktot, (uint64_t)d.N, (uint64_t)d.B};
    const uint64_t strides[2] = {(uint64_t)ktot * 2, (uint64_t)d.w_bstride * 2};
    const uint32_t bbox[3] = {64u, (uint32_t)bn, 1u};
    make_tmap(&p.b_map, d.w, 3, dims, strides, bbox);
    p.b_batched = 1;
  } else {
    const uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)d.N};
    const uint64_t strides[1] = {(uint64_t)ktot * 2};
    const uint32_t bbox[2] = {64u, (uint32_t)(L->swap ? 128 : bn)};
    make_tmap(&p.b_map, d.w, 2, dims, strides, bbox);
  }
  // ---- residual: extra K steps  A = residual tile, B = identity columns
  if (d.res) {
    SDM_CHECK(d.mode == EPI_F16, "residual only with EPI_F16");
    SDM_CHECK(d.N <= kIdentityN && d.res_ld % 8 == 0, "residual constraints");
    const long long bs = d.res_bstride ? d.res_bstride : (long long)Hout * Wout * d.res_ld;
    const uint64_t dims[4] = {(uint64_t)d.N, (uint64_t)Wout, (uint64_t)Hout, (uint64_t)d.B};
    const uint64_t strides[3] = {(uint64_t)d.res_ld * 2, (uint64_t)Wout * d.res_ld * 2, (uint64_t)bs * 2};
    const uint32_t rbox[4] = {64u, (uint32_t)p.tw, (uint32_t)(L->swap_pair ? p.th / 2 : p.th), 1u};  // pair: half a patch per CTA
    make_tmap(&p.r_map, d.res, 4, dims, strides, rbox);
    const uint64_t idims[2] = {(uint64_t)kIdentityN, (uint64_t)kIdentityN};
    const uint64_t istr[1] = {(uint64_t)kIdentityN * 2};
    const uint32_t ibox[2] = {64u, L->swap ? 128u : 64u};  // swap: 128 channel rows x one 64-column slice
    make_tmap(&p.i_map, identity_matrix(), 2, idims, istr, ibox);
    p.has_res = 1;
  } else {
    p.r_map = p.a_map[0];
    p.i_map = p.b_map;
  }
  // ---- epilogue
  p.mode = d.mode;
  p.ups2 = d.ups2;
  SDM_CHECK(!(d.ups2 && (d.res || d.mode != EPI_F16)) || d.mode == EPI_F16, "ups2 only with EPI_F16");
  p.out = d.out;
  p.out_ld = d.out_ld;
  p.out_bstride = d.out_bstride;
  p.bias = d.bias;
  p.bias_sel = d.bias_sel;
  p.scale = d.scale;
  p.post_div = d.post_div;
  p.n_store = d.n_store > 0 ? d.n_store : d.N;
  if (p.mode == EPI_F16 && p.n_store < 8) p.mode = EPI_SKINNY;
  SDM_CHECK(p.post_div == 1.0f || p.mode == EPI_SKINNY, "post_div is only implemented for skinny (n_store < 8) outputs");
  if (p.mode == EPI_SKINNY) SDM_CHECK(bn == 16 && !d.res && !d.ups2, "skinny output needs N <= 16 and no residual");
  p.out2 = d.out2;
  p.gn_ab = d.gn_ab;
  p.gn_silu = d.gn_silu;
  p.stats = (d.mode == EPI_F16 && !d.ups2) ? d.stats : nullptr;
  {  // GroupNorm-partials slots of the swapped kernels: two per 8 x 32 (16 x 16) patch; a polyphase launch fills its quarter
    const int per = 2 * p.tiles_x * p.tiles_y;
    p.stats_bslots = d.poly ? 4 * per : per;
    p.stats_slot0 = d.poly ? (d.poly - 1) * per : 0;
  }
  if (d.mode == EPI_ALPHA) SDM_CHECK(d.N >= 3 && d.N <= 16 && d.bias != nullptr, "EPI_ALPHA needs 3..16 columns and a bias");
  {
    // epilogue warpgroups: two for the GEMMs whose K loop is too short to hide one epilogue
    // measured (A/B on one box): 2 warpgroups help short-K tiles (1x1 conv K=128: 170 -> 294 TFLOP/s), cost a pipeline stage on the
    // 256-wide and 256x128 tiles (1337 -> 1299) -> only where the tile is <= 160 wide and single
    const long long ksteps = (long long)p.ntaps * (cin_total / 64);
    // r1p: 256x128 tiles with ONE or a few K steps (im2col conv_in: 1, VAE 1x1 shortcuts: 4) are pure epilogue: drain the two
    // M sub-tiles concurrently
    int ewg = (p.mode == EPI_GEGLU || p.mode == EPI_F32 || p.mode == EPI_F16_T ||
               (p.mode == EPI_F16 && L->mt == 1 && (bn <= 160 || ksteps <= 16)) ||
               (p.mode == EPI_F16 && L->mt == 2 && ksteps <= 4)) ? 2 : 1;
    if (bn == 16 || p.mode == EPI_ALPHA || p.mode == EPI_SKINNY) ewg = 1;
    L->ewg = ewg;
  }
  {
    static const int env_prof = [] { const char* e = getenv("SDM_GEMM_PROF"); return e ? atoi(e) : 0; }();
    p.prof = env_prof;  // measurement aid: in-kernel cycle counters of the producer / MMA / epilogue threads
  }
  L->grid = (int)std::min<long long>(total, num_sms);
  if (L->swap_pair) L->grid = 2 * (int)std::min<long long>(m_tiles_eff * (p.n_tiles / 2), num_sms / 2);  // whole clusters
  L->flops = 2.0 * (double)d.B * Hout * Wout * (double)d.N * (double)ktot;  // (polyphase: the work actually done, 4/9 of the 3x3 form's per output pixel)
  return L;
}

void conv_gemm_run(const ConvGemmLaunch& l, cudaStream_t st) {
  const ConvGemmParams& p = l.p;
  const int bn = l.block_n, mt = l.mt, g = l.grid;
#define SDM_GO(BN, MT, MODE, UPS2)                                                  \
  do {                                                                              \
    if (l.ewg == 2) return conv_gemm_launch<BN, MT, MODE, UPS2, 2>(p, g, st); \
    return conv_gemm_launch<BN, MT, MODE, UPS2, 1>(p, g, st);                 \
  } while (0)
  if (l.swap) return l.swap_halo ? conv_swap_halo_launch(p, g, l.swap_pair, st) : conv_swap_launch(p, g, st);
  if (l.halo) {
    if (bn == 256 && !p.ups2) return conv_gemm_launch_halo<256, 1, false, 1>(p, g, st);
    if (bn == 256 && p.ups2) return conv_gemm_launch_halo<256, 1, true, 1>(p, g, st);
    if (bn == 160 && !p.ups2) return conv_gemm_launch_halo<160, 1, false, 2>(p, g, st);
    if (bn == 160 && p.ups2) return conv_gemm_launch_halo<160, 1, true, 2>(p, g, st);
    if (bn == 128 && mt == 2 && !p.ups2) return conv_gemm_launch_halo<128, 2, false, 1>(p, g, st);
    throw Error{"conv_gemm: no halo instantiation for this configuration"};
  }
  switch (p.mode) {
    case EPI_F16:
      if (!p.ups2) {
        if (bn == 256) SDM_GO(256, 1, EPI_F16, false);
        if (bn == 160) SDM_GO(160, 1, EPI_F16, false);
        if (bn == 128 && mt == 2) SDM_GO(128, 2, EPI_F16, false);
        if (bn == 128) SDM_GO(128, 1, EPI_F16, false);
        if (bn == 64) SDM_GO(64, 1, EPI_F16, false);
        if (bn == 16) return conv_gemm_launch<16, 1, EPI_F16, false, 1>(p, g, st);
      } else {
        if (bn == 256) SDM_GO(256, 1, EPI_F16, true);
        if (bn == 160) SDM_GO(160, 1, EPI_F16, true);
        if (bn == 128 && mt == 2) SDM_GO(128, 2, EPI_F16, true);
        if (bn == 128) SDM_GO(128, 1, EPI_F16, true);
        if (bn == 64) SDM_GO(64, 1, EPI_F16, true);
      }
      break;
    case EPI_F16_T:
      if (bn == 256) SDM_GO(256, 1, EPI_F16_T, false);
      if (bn == 160) SDM_GO(160, 1, EPI_F16_T, false);
      if (bn == 128) SDM_GO(128, 1, EPI_F16_T, false);
      if (bn == 64) SDM_GO(64, 1, EPI_F16_T, false);
      break;
    case EPI_GEGLU:
      if (bn == 256) SDM_GO(256, 1, EPI_GEGLU, false);
      break;
    case EPI_F32:
      if (bn == 256) SDM_GO(256, 1, EPI_F32, false);
      if (bn == 128) SDM_GO(128, 1, EPI_F32, false);
      if (bn == 64) SDM_GO(64, 1, EPI_F32, false);
      break;
    case EPI_ALPHA:
      if (bn == 16) return conv_gemm_launch<16, 1, EPI_ALPHA, false, 1>(p, g, st);
      break;
    case EPI_SKINNY:
      if (bn == 16) return conv_gemm_launch<16, 1, EPI_SKINNY, false, 1>(p, g, st);
      break;
  }
#undef SDM_GO
  throw Error{"conv_gemm: no kernel instantiation for mode " + std::to_string(p.mode) + " BLOCK_N " + std::to_string(bn) + " MT " +
              std::to_string(mt) + (p.ups2 ? " ups2" : "")};
}
double conv_gemm_flops(const ConvGemmLaunch& l) { return l.flops; }
int conv_gemm_tiles_per_image(int Hout, int Wout) {
  int tw = 128, th = 1;
  pick_patch(Hout, Wout, tw, th);
  return ((Wout + tw - 1) / tw) * ((Hout + th - 1) / th);
}
void conv_gemm_set_outputs(ConvGemmLaunch& l, void* out, void* out2) {
  l.p.out = out;
  l.p.out2 = out2;
}

}  // namespace sdm
