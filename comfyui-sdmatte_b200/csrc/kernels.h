// Host-callable launchers for the sm_100a kernels of the SDMatte matte path.
// Plain C++ (no torch). Every launcher enqueues on the given stream and never synchronises.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <memory>

namespace sdm {

// ------------------------------------------------------------------ tcgen05 conv / GEMM
struct ActSrc {
  const __half* ptr = nullptr;  // NHWC activation (or [B][L][C] tokens)
  int C = 0;                    // channels used from this source (multiple of 64)
  long long ld = 0;             // elements between consecutive pixels (>= C)
};

enum ConvPad { PAD_SAME = 0, PAD_VAE_DOWN = 1 };  // PAD_VAE_DOWN: F.pad(0,1,0,1) + stride 2 + pad 0

struct ConvGemmDesc {
  int B = 1, Hin = 1, Win = 1;  // input spatial dims (tokens: Hin = 1, Win = L)
  int nsrc = 1;
  ActSrc src[2];
  long long src_bstride[2] = {0, 0};  // elements per batch element (0 -> Hin*Win*ld)
  int ksize = 1;                      // 1 or 3
  int stride = 1;                     // 1 or 2 (2 only with ksize 3, nsrc 1)
  int pad = PAD_SAME;
  const __half* w = nullptr;  // [N][ksize*ksize*cin_total], K contiguous
  int N = 0;
  long long w_bstride = 0;  // != 0: per-batch weights (batched GEMM), elements
  int mode = 0;             // EpiMode
  int ups2 = 0;
  void* out = nullptr;
  long long out_ld = 0, out_bstride = 0;
  const float* bias = nullptr;
  const int* bias_sel = nullptr;
  const __half* res = nullptr;
  long long res_ld = 0, res_bstride = 0;
  float scale = 1.0f;
  float post_div = 1.0f;   // EPI_F16: fp16(result) / post_div (rounded again)
  int n_store = 0;         // EPI_F16: store only the first n_store (< 8) columns; 0 = all N
  void* out2 = nullptr;    // EPI_ALPHA: optional pre-clip channel mean
  float* stats = nullptr;  // EPI_F16 without ups2: GroupNorm partials [B][tiles_per_image][N][2] (see conv_gemm_tiles_per_image)
  int force_block_n = 0;  // tests only
  int force_mt = 0;       // tests only: 1 / 2 = force the number of M sub-tiles per CTA tile
  int force_halo = 0;     // tests only: 1 = force the resident-halo 3x3 kernel, -1 = forbid it, 0 = auto
  int force_swap = 0;     // tests only: 1 = force the swapped-operand 3x3 kernel (channels on M), 2 = its resident-halo form, -1 = forbid, 0 = auto
  // fused GroupNorm(+SiLU) of the INPUT: the conv reads the raw producer output and normalises its resident halo tile in shared
  // memory with this [B][cin_total][2] (scale, shift) table (groupnorm_ab); only where conv_gemm_can_fuse_gn() says so
  const float* gn_ab = nullptr;
  int gn_silu = 0;
  // polyphase component of "nearest x2 upsample -> 3x3 conv" (0 = off, 1 + 2 py + px): a 2x2-tap conv over the LOW-resolution
  // input (Hin x Win) with the pre-combined weights [N][4][cin] of that parity (conv_poly_weights), stored at the pixels
  // (2y + py, 2x + px) of a (2 Hin, 2 Win) output; out_bstride / stats describe the full-resolution output: `stats` holds
  // 4 x conv_gemm_tiles_per_image(Hin, Win) slots per sample, launch (py, px) fills slots [q, q + 1) x that count, q = 2 py + px.
  // Only where conv_gemm_can_poly() says so (resident-halo swapped-operand kernel).
  int poly = 0;
};

struct ConvGemmLaunch;  // opaque: prebuilt tensor maps + params
std::shared_ptr<ConvGemmLaunch> conv_gemm_build(const ConvGemmDesc& d, int num_sms);
void conv_gemm_run(const ConvGemmLaunch& l, cudaStream_t st);
void conv_gemm_set_outputs(ConvGemmLaunch& l, void* out, void* out2);  // run-time output slots (EPI_ALPHA)
double conv_gemm_flops(const ConvGemmLaunch& l);
int conv_gemm_tiles_per_image(int Hout, int Wout);  // number of 128-pixel M tiles per batch element
// kernel variant the engine uses for a conv of this per-sample geometry (no batch size in the signature, on purpose):
// 0 = one TMA box per tap, 1 / 2 = resident halo tile with 256- / 160-wide tiles, 3 = swapped operands
int conv_gemm_variant_code(int ksize, int stride, int mode, int ups2, int N, int has_res, int Hout, int Wout);
// can a conv of this per-sample geometry take its input GroupNorm fused (ConvGemmDesc::gn_ab)?  Geometry only, like the variant.
bool conv_gemm_can_fuse_gn(int ksize, int stride, int mode, int ups2, int N, int has_res, int Hout, int Wout);
// can "nearest x2 upsample -> 3x3 conv to N channels" of a Hin x Win input run as four polyphase 2x2-tap convs (ConvGemmDesc::poly)?
bool conv_gemm_can_poly(int N, int Hin, int Win);

// ------------------------------------------------------------------ flash attention (d = 64)
struct AttnDesc {
  int B = 1, heads = 1, Lq = 0, Lk = 0;
  const __half* q = nullptr;   // [B][Lq][ldq], head h at columns h*64..
  long long ldq = 0;
  const __half* k = nullptr;   // [B][Lk][ldk]
  long long ldk = 0;
  const __half* vt = nullptr;  // [B][heads*64][Lk_ld]  (V transposed: keys contiguous)
  long long ldvt = 0;          // row length (>= Lk, multiple of 8)
  const float* bias = nullptr; // [B][Lk_pad] additive per-key bias * log2(e) (padded with -inf to a multiple of 128) or null
  long long bias_bstride = 0;
  const int* ntiles = nullptr; // optional [B]: 128-key tiles to stream for sample b (compacted keys); null = ceil(Lk/128)
  __half* out = nullptr;       // [B][Lq][ldo]
  long long ldo = 0;
  float scale = 0.125f;
};
struct AttnLaunch;
std::shared_ptr<AttnLaunch> attn_build(const AttnDesc& d);
void attn_run(const AttnLaunch& l, cudaStream_t st);

// ------------------------------------------------------------------ norms / softmax
// GroupNorm(32 groups) over NHWC input given as up to two channel-concatenated sources.
// stats: per-(b, channel) partial sums -> per-(b, group) mean/rstd; apply: y = silu?((x-mean)*rstd*gamma+beta) fp16.
struct GroupNormDesc {
  int B = 1, HW = 0;
  int nsrc = 1;
  const __half* src[2] = {nullptr, nullptr};
  int C[2] = {0, 0};
  long long ld[2] = {0, 0};
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float eps = 1e-5f;
  int silu = 1;
  __half* out = nullptr;   // [B][HW][C0+C1]; null = statistics + finalize only (the (scale, shift) table stays in scratch: groupnorm_ab)
  float* scratch = nullptr;  // >= groupnorm_scratch_floats(...) floats
  // optional: per-(image tile, channel) partial sums written by the producing conv's epilogue (one buffer per source)
  const float* pre_partial[2] = {nullptr, nullptr};
  int pre_slots = 0;
};
size_t groupnorm_scratch_floats(int B, int HW, int Ctot);
void groupnorm_run(const GroupNormDesc& d, cudaStream_t st);
size_t groupnorm_ab_offset_floats(int B, int HW, int Ctot);
const float* groupnorm_ab(const GroupNormDesc& d);  // [B][C0+C1][2] (scale, shift) inside d.scratch
int groupnorm_num_launches(const GroupNormDesc& d);  // kernels groupnorm_run will launch for this descriptor (2..4)

void layernorm_run(const __half* x, __half* y, const float* gamma, const float* beta, long long rows, int C, float eps,
                   cudaStream_t st);
// rows of fp32 scores -> fp16 probabilities (softmax over the last dim)
void softmax_rows_run(const float* s, __half* p, long long rows, int L, cudaStream_t st);

// ------------------------------------------------------------------ per-sample conditioning for point / bbox / mask prompts (cond_embed.cu)
struct CondEmbedDesc {
  int B = 0;
  const int* is_trans = nullptr;     // [B] device
  const float* coords = nullptr;     // [B][ncoords] device
  int ncoords = 0, npad = 0, dim = 0, Kc = 0;  // coordinates, padded count, sinusoid channels per coordinate, npad * dim
  const float *te_w1 = nullptr, *te_b1 = nullptr, *te_w2 = nullptr, *te_b2 = nullptr;  // time_embedding (320 -> 1280 -> 1280)
  const float *ce_w1 = nullptr, *ce_b1 = nullptr, *ce_w2 = nullptr, *ce_b2 = nullptr;  // bbox_ / point_embedding (Kc -> 1280 -> 1280)
  const __half* tp_w = nullptr;      // [rows][1280]: time_emb_proj of all 22 resnets, stacked
  const float* tp_b = nullptr;       // [rows]: time_emb_proj.bias + conv1.bias
  int rows = 0;
  const int4* row_map = nullptr;     // [rows]: (table offset of the row's resnet, its channel count, channel)
  float *xt = nullptr, *xc = nullptr, *ht = nullptr, *hc = nullptr, *semb = nullptr;  // scratch: [B][320], [B][Kc], [B][1280] x 3
  float* tables = nullptr;           // out: per resnet [B][C] bias rows
};
void cond_embed_run(const CondEmbedDesc& d, cudaStream_t st);  // 5 launches

// ------------------------------------------------------------------ small direct convs (Cin <= 8 or Cout <= 8)
struct DirectConvDesc {
  int B = 1, H = 0, W = 0;          // stride 1, pad (k-1)/2
  int Cin = 0, Cout = 0, ksize = 3;
  const __half* x = nullptr; long long x_ld = 0;   // NHWC, pixel stride
  const __half* w = nullptr;   // [Cout][k*k][Cin] fp16
  const float* bias = nullptr; // [Cout]
  __half* out = nullptr; long long out_ld = 0; int out_coff = 0;  // writes channels [out_coff, out_coff+Cout)
  float out_scale = 1.0f;      // applied after fp16 rounding of conv+bias (fp16 multiply), 1 = none
  __half* out2 = nullptr; long long out2_ld = 0; int out2_coff = 0;  // optional second copy
  int cout_limit = 0;          // if >0 only the first cout_limit channels are stored (VAE "mean" half)
  float out_div = 1.0f;        // small-Cout path only: fp16(out) / out_div, rounded to fp16 (latent / scaling_factor)
};
void direct_conv_run(const DirectConvDesc& d, cudaStream_t st);

// decoder head: conv3x3 (C->3) on an already normalised/activated input, then mean over the 3 channels,
// clip(-1,1), (x+1)/2 -> alpha fp16 [B][H][W]   (reference meta_arch.py:256-260)
void alpha_head_run(const __half* x, long long x_ld, int B, int H, int W, int Cin, const __half* w, const float* bias,
                    __half* alpha, __half* premean /*nullable: pre-clip mean, fp16*/, cudaStream_t st);

// alpha head, second half: y [B][H][W][32] fp32 = per-tap partial products (column tap*3 + c) of the decoder conv_out;
// alpha[b][y][x] = (clip(mean_c fp16(bias_c + sum_tap y[(y+ky-1, x+kx-1)][tap*3+c]), -1, 1) + 1) / 2 with the reference's fp16
// rounding points (meta_arch.py:256-260); premean (nullable) = the pre-clip mean
void alpha_col2im_run(const float* y, const float* bias, int B, int H, int W, __half* alpha, __half* premean, cudaStream_t st);

// ------------------------------------------------------------------ elementwise
// im2col of the VAE conv_in input: out[2B*R*R][64], k = tap*4 + channel (zero beyond 36); first B images = (x-0.5)/0.5 of the
// fp32 [B][R][R][3] image, last B = trimap*2-1 replicated to 3 channels (sdmatte_nodes.py:343,351, meta_arch.py:141)
void prep_inputs_run(const float* image, const float* trimap, __half* out /*[2B][R][R][64]*/, int ldc, int B, int R,
                     cudaStream_t st);
// key-bias vectors for the 4 UNet levels: bias_k[b][i*s+j] = (1 - trimap[b][8*2^k*i][8*2^k*j]) * -10000,
// padded to lpad[k] (multiple of 128) with -inf.   (reference meta_arch.py:200-204, replace.py:56-63,401-403)
void key_bias_run(const float* trimap, int B, int R, float* bias0, float* bias1, float* bias2, float* bias3,
                  const int* lpad, cudaStream_t st);
// self-attention key compaction (small_ops.cu): per sample and level, the keys whose bias is within 2500 of the sample's
// maximum, gathered in order: idx[b][i] (padded to a multiple of 128 with a valid index), cbias[b][i] (padding = -inf),
// ntiles[b] = padded count / 128.  All arrays use the batch stride lpad[level].
void key_compact_run(const float* const* bias, float* const* cbias, int* const* idx, int* const* ntiles, const int* lpad, int B, int S,
                     cudaStream_t st);
void key_compact_level_run(const float* bias, float* cbias, int* idx, int* ntiles, int B, int L, int lpad, cudaStream_t st);  // one level
// dst[b][i][:] = src[b][idx[b][i]][:] for i < 128 * ntiles[b]; src/dst are [B][L][C] fp16
void gather_rows_run(const __half* src, __half* dst, const int* idx, const int* ntiles, int B, int L, int C, int idx_bstride,
                     cudaStream_t st);

// ------------------------------------------------------------------ node pre/post-processing (prepost.cu)
// antialiased bilinear resize of image [B][H][W][3] / trimap [B][H][W] (fp32) to R x R (sdmatte_nodes.py:204-214,343,349)
void preprocess_run(const float* image, const float* trimap, int B, int H, int W, int R, float* image_out, float* trimap_out,
                    cudaStream_t st);
// resize alpha [B][R][R] fp16 back to (H, W), clamp, mask_refine, output_mode composition (sdmatte_nodes.py:362-397)
void postprocess_run(const __half* alpha, int B, int R, int H, int W, const float* image, const float* trimap, int mask_refine,
                     double trimap_constraint, int output_mode, __half* alpha_out, float* matted_out, cudaStream_t st);

// hardware probe (probe.cu): shifted window of a swizzled halo tile as a tcgen05 A operand; out [128][64] fp32
void probe_halo_run(const __half* x /*[16][8][64]*/, const __half* eye /*[64][64]*/, float* out, int dy, int dx, int mode, cudaStream_t st);

int device_sm_count();
}  // namespace sdm
