/* sdmatte_b200 — C ABI of the B200-native SDMatte matte engine (libsdmatte_b200.so).
 *
 * The reference (flybirdxx/ComfyUI-SDMatte) has no FFI boundary of its own: the path sits behind the
 * ComfyUI node protocol (sdmatte_nodes.py:217-405).  This header is the boundary the replacement
 * exports; each entry point cites the reference code it replaces.  Conventions:
 *   - return 0 = OK, non-zero = error; sdm_last_error() gives the message (thread-local).
 *   - every pointer is caller-owned memory; "dev" = device pointer on the handle's device.
 *   - nothing here synchronises the stream unless stated; `stream` is a cudaStream_t cast to uintptr_t.
 *   - no torch / C++ types cross this boundary.
 */
#ifndef SDMATTE_B200_H
#define SDMATTE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdm_handle sdm_handle;

typedef struct {
  const char* name;      /* state-dict key with the reference's names, e.g. "unet.conv_in.weight"  */
  int dtype;             /* 0 = fp32, 1 = fp16, 2 = bf16 */
  int ndim;
  int64_t shape[4];
  const void* data;      /* HOST pointer, contiguous, borrowed for the duration of the call */
} sdm_tensor_desc;

int sdm_version(void);
const char* sdm_last_error(void);

/* Replaces SDMatteCore(...) construction, sdmatte_nodes.py:286-296 (architecture is fixed: SD-2.1 CustomUNet + SD VAE). */
int sdm_create(sdm_handle** out, int device);
void sdm_destroy(sdm_handle* h);

/* Replaces load_state_dict(strict=False) + .to(device), sdmatte_nodes.py:298-323.  Repacks to fp16 device buffers,
 * constant-folds the time/opacity/bbox embeddings (replace.py:430-459).  Fails loudly on missing keys;
 * unexpected keys are counted (sdm_load_report). */
int sdm_load_weights(sdm_handle* h, const sdm_tensor_desc* tensors, int n);
int sdm_load_report(sdm_handle* h, int* n_used, int* n_unexpected);

/* Native .safetensors reader: replaces the safe_open / get_tensor loop of sdmatte_nodes.py:298-304.
 * The file is mapped read-only; sdm_safetensors_entry fills a descriptor whose `data` points INTO the mapping (valid until
 * sdm_safetensors_close), ready to be passed to sdm_load_weights.  dtype: 0 = F32, 1 = F16, 2 = BF16, -1 = anything else
 * (not a weight of this model); ndim is the tensor's true rank (shape[] holds at most the first four extents). */
typedef struct sdm_safetensors sdm_safetensors;
int sdm_safetensors_open(const char* path, sdm_safetensors** out);
int sdm_safetensors_count(const sdm_safetensors* f);
int sdm_safetensors_entry(sdm_safetensors* f, int i, sdm_tensor_desc* out);
void sdm_safetensors_close(sdm_safetensors* f);

/* Workspace the caller must provide for a (B, R) forward; R in {64k}, multiples of 64. */
size_t sdm_workspace_bytes(sdm_handle* h, int B, int R);

/* Replaces SDMatte.forward (meta_arch.py:127-261) incl. CustomUNet.forward (replace.py:379-549) and the
 * pre-processing normalisation of sdmatte_nodes.py:343,351 for inputs already at R x R.
 *   image_dev  : [B][R][R][3] fp32 in [0,1]   (ComfyUI IMAGE layout)
 *   trimap_dev : [B][R][R]    fp32 in [0,1]   (ComfyUI MASK layout)
 *   is_trans   : HOST int32[B] (sdmatte_nodes.py:345)
 *   alpha_dev  : [B][R][R] fp16 in [0,1] (reference returns fp16 on its CUDA path)
 *   premean_dev: optional [B][R][R] fp16, decoder channel-mean before clip (parity diagnostics), may be NULL */
int sdm_forward(sdm_handle* h, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                void* alpha_dev, void* premean_dev, void* workspace_dev, size_t workspace_bytes, uintptr_t stream);

/* The other visual prompts of the SDMatte model (SDMatte.forward meta_arch.py:130-197, CustomUNet.forward
 * replace.py:446-455) — not reachable from the reference's node, which fixes aux_input="trimap" and coords [0,0,1,1]
 * (sdmatte_nodes.py:286-296,353): the auxiliary image `aux_dev` [B][R][R] fp32 in [0,1] (it takes the trimap's place: VAE latent,
 * cross-attention context, attention-mask source) is
 *   prompt_kind 0: a mask / bbox mask (or a trimap with real coordinates): coords_host [B][4] -> 4 x 320 sinusoids -> bbox_embedding
 *   prompt_kind 1: a point mask: coords_host [B][ncoords] -> zero-padded to the first divisor of 1680 -> point_embedding
 * The coordinate -> emb -> time_emb_proj chain, folded into constants for sdm_forward, runs per sample at the head of the plan
 * (csrc/cond_embed.cu).  Workspace: sdm_workspace_bytes_prompt. */
size_t sdm_workspace_bytes_prompt(sdm_handle* h, int B, int R, int prompt_kind, int ncoords);
int sdm_forward_prompt(sdm_handle* h, const float* image_dev, const float* aux_dev, int B, int R, const int32_t* is_trans, int prompt_kind,
                       const float* coords_host, int ncoords, void* alpha_dev, void* premean_dev, void* workspace_dev,
                       size_t workspace_bytes, uintptr_t stream);

/* Same, with HOST buffers (pinned or pageable) already at R x R: H2D of image/trimap and D2H of alpha on `stream`, then a
 * stream sync (replaces the .to(device) / .cpu() pair at sdmatte_nodes.py:342,349,363 for pre-sized inputs; bench / tests). */
int sdm_forward_host(sdm_handle* h, const float* image_host, const float* trimap_host, int B, int R, const int32_t* is_trans,
                     void* alpha_host_f16, void* workspace_dev, size_t workspace_bytes, uintptr_t stream);

/* THE call the `Apply SDMatte` node makes — everything between the argument checks and the return of
 * SDMatteApply.apply_matte (sdmatte_nodes.py:339-397) for host tensors of ANY size:
 *   image_host [B][H][W][3] fp32 / trimap_host [B][H][W] fp32 (pageable ComfyUI tensors are fine: they are staged through a
 *   page-locked buffer owned by the handle, by several host threads, chunk by chunk, each chunk's H2D enqueued behind its memcpy)
 *   -> antialiased resize to R x R (only if (H, W) != (R, R)) -> sdm_forward -> resize back + clamp + mask_refine + output_mode
 *   composition -> D2H -> alpha_out_host_f16 [B][H][W] fp16 and matted_out_host [B][H][W][3|4] fp32 (NULL / untouched for
 *   output_mode 0 "alpha_only": the caller returns zeros, sdmatte_nodes.py:384-385).  Synchronises `stream` before returning.
 * The device-side staging lives behind the plan arena in the caller's workspace (sdm_node_workspace_bytes), at fixed offsets,
 * so repeated calls of one geometry replay the plan's CUDA graph. */
size_t sdm_node_workspace_bytes(sdm_handle* h, int B, int H, int W, int R, int output_mode);
int sdm_apply_matte_host(sdm_handle* h, const float* image_host, const float* trimap_host, int B, int H, int W, int R,
                         const int32_t* is_trans, int mask_refine, double trimap_constraint, int output_mode, void* alpha_out_host_f16,
                         float* matted_out_host, void* workspace_dev, size_t workspace_bytes, uintptr_t stream);

/* Node-side pre/post-processing on the device, all pointers are device pointers.
 * sdm_preprocess replaces torchvision Resize(antialias=True) of image and trimap to R x R
 *   (_resize_norm_image_bchw / _resize_mask_b1hw, sdmatte_nodes.py:204-214, called at :343 and :349; the affine
 *   normalisations are folded into sdm_forward):
 *   image_dev [B][H][W][3] fp32, trimap_dev [B][H][W] fp32 -> image_out_dev [B][R][R][3], trimap_out_dev [B][R][R] fp32.
 * sdm_postprocess replaces sdmatte_nodes.py:362-397: Resize((H, W)) of the fp16 alpha, clamp(0,1), mask_refine with the
 *   ORIGINAL trimap and `trimap_constraint`, and the output_mode composition:
 *   output_mode 0 = "alpha_only" (matted_out untouched: the caller returns zeros), 1 = "matted_rgba" (matted_out
 *   [B][H][W][4] fp32), 2 = "matted_rgb" ([B][H][W][3] fp32), 3 = image * alpha (the reference's unreachable else branch).
 *   alpha_out_dev [B][H][W] fp16.  image_dev / trimap_dev are the caller's original-size tensors. */
int sdm_preprocess(const float* image_dev, const float* trimap_dev, int B, int H, int W, int R, float* image_out_dev,
                   float* trimap_out_dev, uintptr_t stream);
int sdm_postprocess(const void* alpha_dev_f16, int B, int R, int H, int W, const float* image_dev, const float* trimap_dev,
                    int mask_refine, double trimap_constraint, int output_mode, void* alpha_out_dev_f16, float* matted_out_dev,
                    uintptr_t stream);

/* Measurement aid (bench.py roofline): one forward with a CUDA-event pair around every op of the plan (synchronises).
 * sdm_profile_entry returns, per op: kind ("tc:conv3x3", "tc:attention_self", "groupnorm", ...), device ms,
 * algorithmic FLOPs and algorithmic HBM bytes. */
int sdm_forward_profiled(sdm_handle* h, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                         void* alpha_dev, void* workspace_dev, size_t workspace_bytes, uintptr_t stream);
int sdm_profile_count(sdm_handle* h);
int sdm_profile_entry(sdm_handle* h, int i, char* kind, int kind_len, float* ms, double* flops, double* bytes);

/* Counters for the last forward: number of kernel launches, algorithmic tensor FLOPs issued by the tcgen05 kernels. */
int sdm_last_forward_stats(sdm_handle* h, int* n_launches, double* tensor_flops);

/* Intermediate taps for block-level parity tests: copies the named activation of the LAST forward into dst (device).
 * Always available: "unet_in" [B,S,S,8] (rgb latent | trimap latent), "ctx" [B,S,S,1024] (trimap tokens),
 * "unet_out_scaled" [B,S,S,4] (UNet output / scaling_factor).  With sdm_set_option(h, "keep_taps", 1) one tap per block of the
 * graph (SDMatte.forward meta_arch.py:127-261 / CustomUNet.forward replace.py:462-544 / the VAE, in graph order, enumerated by
 * sdm_debug_tensor_count / _name): "enc.conv_in", "enc.down0..3", "enc.mid_attn", "enc.mid" (batch 2B: rgb samples, then
 * trimap samples), "unet.conv_in", "unet.down{i}.{j}", "unet.down{i}.ds", "unet.mid", "unet.up{i}.{j}" (the last block of an
 * up stage is tapped after its fused nearest-x2 store), "unet.up{i}.us", "dec.conv_in", "dec.mid", "dec.up0..3".
 * shape4 receives (B,H,W,C); dtype 1 = fp16.  dst_dev == NULL: shape query only. */
int sdm_debug_tensor(sdm_handle* h, const char* name, void* dst_dev, size_t dst_bytes, int64_t* shape4, int* dtype);
int sdm_debug_tensor_count(sdm_handle* h);
int sdm_debug_tensor_name(sdm_handle* h, int i, char* name, int name_len);
/* Engine options (changing one drops the cached plan): "keep_taps" 0/1 (diagnostics) — tapped block outputs are not recycled
 * by the workspace arena, so sdm_workspace_bytes grows.  "cuda_graph" 0/1 (default 1): when a forward is called again with
 * the same (B, R, workspace, image/trimap/alpha pointers) the ~700 launches of the plan are replayed as ONE CUDA graph
 * (captured on an internal stream, launched on the caller's); sdm_forward_host always qualifies (it stages through fixed
 * buffers in the workspace).  sdm_graph_stats: graphs captured / graph launches so far. */
int sdm_set_option(sdm_handle* h, const char* name, int value);
int sdm_graph_stats(sdm_handle* h, int* captures, int* launches);
/* Host-side wall-clock split of the LAST sdm_apply_matte_host call, milliseconds: [0] staging memcpy + H2D enqueue, [1] kernel /
 * D2H enqueue, [2] waiting for the GPU, [3] copy-out of the results (measurement aid for bench.py's e2e line).
 * Option "copy_threads" (1..64, default 8): host threads of the staging copies. */
int sdm_node_call_timing(sdm_handle* h, double* ms4);

/* ---- single-kernel entry points (parity tests at the kernel level; all pointers are device pointers) ---- */
typedef struct {
  int B, Hin, Win;
  int nsrc;
  const void* src0; int c0; int64_t ld0;
  const void* src1; int c1; int64_t ld1;
  int ksize, stride, pad;       /* pad: 0 = same, 1 = VAE down (0,1,0,1) */
  const void* w; int N; int64_t w_bstride;
  int mode; int ups2;           /* mode: 0 f16, 1 f16 transposed, 2 GEGLU, 3 f32, 4 alpha head */
  void* out; int64_t out_ld; int64_t out_bstride;
  const float* bias; const int32_t* bias_sel;
  const void* res; int64_t res_ld; int64_t res_bstride;
  float scale;
  int force_block_n;
  float post_div;               /* skinny outputs (n_store < 8) only: fp16(result) / post_div, rounded again; 0 or 1 = off */
  int n_store;                  /* mode 0: store only the first n_store (< 8) columns; 0 = all */
  void* out2;                   /* mode 4 (alpha head): optional pre-clip mean */
  int force_mt;                 /* tests: 1 / 2 = force M sub-tiles per CTA tile (BLOCK_N 128 only), 0 = auto */
  float* stats;                 /* mode 0 without ups2: GroupNorm partials [B][sdm_k_conv_tiles_per_image][N][2] (sum, sumsq), or NULL */
  int force_halo;               /* tests: 1 = force the resident-halo 3x3 kernel (one halo tile per 64 channels, nine taps from it), -1 = forbid */
  int force_swap;               /* tests: 1 = force the swapped-operand 3x3 kernel (channels on M, 256 pixels on N), 2 = its resident-halo form, -1 = forbid it */
  const float* gn_ab;           /* fused GroupNorm(+SiLU) of the INPUT: [B][c0+c1][2] (scale, shift) as sdm_k_groupnorm leaves them in its scratch
                                   (offset sdm_k_groupnorm_ab_offset); the conv then reads the raw tensor.  Only where sdm_k_conv_can_fuse_gn() != 0 */
  int gn_silu;
  int poly;                     /* 0, or 1 + 2 py + px: polyphase component (py, px) of "nearest x2 upsample -> 3x3 conv": a 2x2-tap conv over the
                                   LOW-resolution input with the pre-combined weights [N][4][c0] of that parity, stored at the pixels (2y+py, 2x+px) of
                                   the (2 Hin, 2 Win) output (out_bstride / stats describe that output; stats: 4 x sdm_k_conv_tiles_per_image(Hin, Win)
                                   slots per sample, this launch fills quarter 2 py + px).  Only where sdm_k_conv_can_poly() != 0 */
} sdm_conv_gemm_args;
int sdm_k_conv_gemm(const sdm_conv_gemm_args* a, uintptr_t stream);
int sdm_k_conv_tiles_per_image(int Hout, int Wout);
/* Kernel variant the engine picks for a conv of this PER-SAMPLE geometry (host logic only; the batch size is not an argument
   on purpose: a sample must give the same bits alone and in a batch).  mode as in sdm_conv_gemm_args.
   0 = one TMA box per tap, 1 / 2 = resident halo tile with 256- / 160-wide tiles, 3 = swapped operands (conv_swap_kernel) */
int sdm_k_conv_variant(int ksize, int stride, int mode, int ups2, int N, int has_res, int Hout, int Wout);
/* 1 if a conv of this per-sample geometry can take its input GroupNorm fused (sdm_conv_gemm_args.gn_ab) */
int sdm_k_conv_can_fuse_gn(int ksize, int stride, int mode, int ups2, int N, int has_res, int Hout, int Wout);
/* 1 if "nearest x2 upsample -> 3x3 conv to N channels" of a Hin x Win input can run as four polyphase launches (sdm_conv_gemm_args.poly) */
int sdm_k_conv_can_poly(int N, int Hin, int Win);

typedef struct {
  int B, heads, Lq, Lk;
  const void* q; int64_t ldq;
  const void* k; int64_t ldk;
  const void* vt; int64_t ldvt;
  const float* bias; int64_t bias_bstride;   /* additive per-key bias * log2(e), padded with -inf to 128 keys */
  void* out; int64_t ldo;
  float scale;
  const int32_t* ntiles;                     /* optional [B]: number of 128-key tiles to stream for sample b (compacted keys) */
} sdm_attn_args;
int sdm_k_attention(const sdm_attn_args* a, uintptr_t stream);
/* Self-attention key compaction for ONE UNet level (replaces nothing in the reference: it removes the keys whose
   probability is exactly 0 under the -10000 trimap bias of replace.py:401-403 / :100-106).
   bias [B][lpad] (log2 domain, -inf padded) -> idx [B][lpad] kept key indices in order (padded to a multiple of 128 with a
   valid index), cbias [B][lpad] their biases (-inf padding), ntiles [B] = padded count / 128.  L = real key count. */
int sdm_k_key_compact(const float* bias, float* cbias, int32_t* idx, int32_t* ntiles, int B, int L, int lpad, uintptr_t stream);
/* Additive attn1 key bias of the four UNet levels from the trimap at R x R (replaces meta_arch.py:200-204: (tri+1)/2, nearest / 8,
   flatten; replace.py:401-403: (1 - mask) * -10000; replace.py:56-63: nearest resize to each level's grid — the three compose to
   strided sampling of the trimap): bias_l [B][lpad4[l]] fp32 in the LOG2 domain (x log2 e), -inf beyond the level's (R/8 >> l)^2 keys. */
int sdm_k_key_bias(const float* trimap, int B, int R, float* bias0, float* bias1, float* bias2, float* bias3, const int32_t* lpad4,
                   uintptr_t stream);
/* Hardware probe (tests/probe_halo.py): D = shifted 16x8 window of a TMA-swizzled (18x10) halo tile, as a tcgen05 A operand
   with SBO = 1280 B and an unaligned start; x [16][8][64] fp16, eye [64][64] fp16 identity, out [128][64] fp32;
   mode 0: descriptor base_offset 0, mode 1: base_offset = (start >> 7) & 7 */
int sdm_k_probe_halo(const void* x, const void* eye, float* out, int dy, int dx, int mode, uintptr_t stream);
/* dst[b][i][:] = src[b][idx[b][i]][:] for i < 128 * ntiles[b]; src/dst [B][L][C] fp16 */
int sdm_k_gather_rows(const void* src, void* dst, const int32_t* idx, const int32_t* ntiles, int B, int L, int C, int idx_bstride,
                      uintptr_t stream);

typedef struct {
  int B, HW, nsrc;
  const void* src0; int c0; int64_t ld0;
  const void* src1; int c1; int64_t ld1;
  const float* gamma; const float* beta; float eps; int silu;
  void* out; float* scratch; size_t scratch_floats;
  const float* pre0; const float* pre1; int pre_slots;   /* optional partial statistics from the producing convs' epilogues */
} sdm_groupnorm_args;
size_t sdm_k_groupnorm_scratch_floats(int B, int HW, int C);
/* float offset of the [B][C][2] (scale, shift) table inside the scratch buffer; out == NULL in sdm_groupnorm_args: statistics +
   finalize only (no apply pass) */
size_t sdm_k_groupnorm_ab_offset(int B, int HW, int C);
int sdm_k_groupnorm(const sdm_groupnorm_args* a, uintptr_t stream);
int sdm_k_layernorm(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int C, float eps, uintptr_t stream);
int sdm_k_softmax_rows(const float* s, void* p, int64_t rows, int L, uintptr_t stream);

typedef struct {
  int B, H, W, Cin, Cout, ksize;
  const void* x; int64_t x_ld;
  const void* w; const float* bias;
  void* out; int64_t out_ld; int out_coff; float out_scale; int cout_limit;
} sdm_direct_conv_args;
int sdm_k_direct_conv(const sdm_direct_conv_args* a, uintptr_t stream);

#ifdef __cplusplus
}
#endif
#endif
