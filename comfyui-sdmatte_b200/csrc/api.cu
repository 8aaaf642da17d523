// extern "C" surface of libsdmatte_b200.so (see include/sdmatte_b200.h).
#include "sdmatte_b200.h"

#include "common.cuh"
#include "engine.h"
#include "kernels.h"

#include <string>

namespace sdm {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
int device_sm_count() {
  int dev = 0, n = 0;
  SDM_CUDA_OK(cudaGetDevice(&dev));
  SDM_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}
}  // namespace sdm

#define SDM_API_BEGIN try {
#define SDM_API_END                          \
  return 0;                                  \
  }                                          \
  catch (const sdm::Error& e) {              \
    sdm::set_last_error(e.msg);              \
    return 1;                                \
  }                                          \
  catch (const std::exception& e) {          \
    sdm::set_last_error(e.what());           \
    return 2;                                \
  }

extern "C" {

int sdm_version(void) { return 201; }
const char* sdm_last_error(void) { return sdm::g_last_error.c_str(); }

int sdm_create(sdm_handle** out, int device) {
  SDM_API_BEGIN
  *out = reinterpret_cast<sdm_handle*>(sdm::engine_create(device));
  SDM_API_END
}
void sdm_destroy(sdm_handle* h) { sdm::engine_destroy(reinterpret_cast<sdm::Engine*>(h)); }

int sdm_load_weights(sdm_handle* h, const sdm_tensor_desc* tensors, int n) {
  SDM_API_BEGIN
  sdm::engine_load(reinterpret_cast<sdm::Engine*>(h), tensors, n);
  SDM_API_END
}
int sdm_load_report(sdm_handle* h, int* n_used, int* n_unexpected) {
  SDM_API_BEGIN
  sdm::engine_load_report(reinterpret_cast<sdm::Engine*>(h), n_used, n_unexpected);
  SDM_API_END
}
size_t sdm_workspace_bytes(sdm_handle* h, int B, int R) {
  try {
    return sdm::engine_workspace_bytes(reinterpret_cast<sdm::Engine*>(h), B, R);
  } catch (const sdm::Error& e) {
    sdm::set_last_error(e.msg);
    return 0;
  }
}
int sdm_forward(sdm_handle* h, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                void* alpha_dev, void* premean_dev, void* workspace_dev, size_t workspace_bytes, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::engine_forward(reinterpret_cast<sdm::Engine*>(h), image_dev, trimap_dev, B, R, is_trans, alpha_dev, premean_dev,
                      workspace_dev, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
size_t sdm_workspace_bytes_prompt(sdm_handle* h, int B, int R, int prompt_kind, int ncoords) {
  try {
    return sdm::engine_workspace_bytes_prompt(reinterpret_cast<sdm::Engine*>(h), B, R, prompt_kind, ncoords);
  } catch (const sdm::Error& e) {
    sdm::set_last_error(e.msg);
    return 0;
  }
}
int sdm_forward_prompt(sdm_handle* h, const float* image_dev, const float* aux_dev, int B, int R, const int32_t* is_trans, int prompt_kind,
                       const float* coords_host, int ncoords, void* alpha_dev, void* premean_dev, void* workspace_dev,
                       size_t workspace_bytes, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::engine_forward_prompt(reinterpret_cast<sdm::Engine*>(h), image_dev, aux_dev, B, R, is_trans, prompt_kind, coords_host, ncoords,
                             alpha_dev, premean_dev, workspace_dev, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
int sdm_forward_host(sdm_handle* h, const float* image_host, const float* trimap_host, int B, int R, const int32_t* is_trans,
                     void* alpha_host_f16, void* workspace_dev, size_t workspace_bytes, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::engine_forward_host(reinterpret_cast<sdm::Engine*>(h), image_host, trimap_host, B, R, is_trans, alpha_host_f16,
                           workspace_dev, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
size_t sdm_node_workspace_bytes(sdm_handle* h, int B, int H, int W, int R, int output_mode) {
  try {
    return sdm::engine_node_workspace_bytes(reinterpret_cast<sdm::Engine*>(h), B, H, W, R, output_mode);
  } catch (const sdm::Error& e) {
    sdm::set_last_error(e.msg);
    return 0;
  }
}
int sdm_apply_matte_host(sdm_handle* h, const float* image_host, const float* trimap_host, int B, int H, int W, int R,
                         const int32_t* is_trans, int mask_refine, double trimap_constraint, int output_mode, void* alpha_out_host_f16,
                         float* matted_out_host, void* workspace_dev, size_t workspace_bytes, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::engine_apply_host(reinterpret_cast<sdm::Engine*>(h), image_host, trimap_host, B, H, W, R, is_trans, mask_refine, trimap_constraint,
                         output_mode, alpha_out_host_f16, matted_out_host, workspace_dev, workspace_bytes,
                         reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
int sdm_preprocess(const float* image_dev, const float* trimap_dev, int B, int H, int W, int R, float* image_out_dev,
                   float* trimap_out_dev, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::preprocess_run(image_dev, trimap_dev, B, H, W, R, image_out_dev, trimap_out_dev, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
int sdm_postprocess(const void* alpha_dev_f16, int B, int R, int H, int W, const float* image_dev, const float* trimap_dev,
                    int mask_refine, double trimap_constraint, int output_mode, void* alpha_out_dev_f16, float* matted_out_dev,
                    uintptr_t stream) {
  SDM_API_BEGIN
  sdm::postprocess_run(reinterpret_cast<const __half*>(alpha_dev_f16), B, R, H, W, image_dev, trimap_dev, mask_refine,
                       trimap_constraint, output_mode, reinterpret_cast<__half*>(alpha_out_dev_f16), matted_out_dev,
                       reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
int sdm_forward_profiled(sdm_handle* h, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                         void* alpha_dev, void* workspace_dev, size_t workspace_bytes, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::engine_forward_profiled(reinterpret_cast<sdm::Engine*>(h), image_dev, trimap_dev, B, R, is_trans, alpha_dev, workspace_dev,
                               workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
int sdm_profile_count(sdm_handle* h) { return sdm::engine_profile_count(reinterpret_cast<sdm::Engine*>(h)); }
int sdm_profile_entry(sdm_handle* h, int i, char* kind, int kind_len, float* ms, double* flops, double* bytes) {
  SDM_API_BEGIN
  sdm::engine_profile_entry(reinterpret_cast<sdm::Engine*>(h), i, kind, kind_len, ms, flops, bytes);
  SDM_API_END
}
int sdm_last_forward_stats(sdm_handle* h, int* n_launches, double* tensor_flops) {
  SDM_API_BEGIN
  sdm::engine_stats(reinterpret_cast<sdm::Engine*>(h), n_launches, tensor_flops);
  SDM_API_END
}
int sdm_debug_tensor(sdm_handle* h, const char* name, void* dst_dev, size_t dst_bytes, int64_t* shape4, int* dtype) {
  SDM_API_BEGIN
  sdm::engine_debug_tensor(reinterpret_cast<sdm::Engine*>(h), name, dst_dev, dst_bytes, shape4, dtype);
  SDM_API_END
}

int sdm_debug_tensor_count(sdm_handle* h) { return sdm::engine_debug_tensor_count(reinterpret_cast<sdm::Engine*>(h)); }
int sdm_debug_tensor_name(sdm_handle* h, int i, char* name, int name_len) {
  SDM_API_BEGIN
  snprintf(name, name_len, "%s", sdm::engine_debug_tensor_name(reinterpret_cast<sdm::Engine*>(h), i));
  SDM_API_END
}
int sdm_graph_stats(sdm_handle* h, int* captures, int* launches) {
  SDM_API_BEGIN
  sdm::engine_graph_stats(reinterpret_cast<sdm::Engine*>(h), captures, launches);
  SDM_API_END
}
int sdm_node_call_timing(sdm_handle* h, double* ms4) {
  SDM_API_BEGIN
  sdm::engine_node_timing(reinterpret_cast<sdm::Engine*>(h), ms4);
  SDM_API_END
}
int sdm_set_option(sdm_handle* h, const char* name, int value) {
  SDM_API_BEGIN
  sdm::engine_set_option(reinterpret_cast<sdm::Engine*>(h), name, value);
  SDM_API_END
}

// ---------------------------------------------------------------- kernel-level entry points
int sdm_k_conv_gemm(const sdm_conv_gemm_args* a, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::ConvGemmDesc d;
  d.B = a->B; d.Hin = a->Hin; d.Win = a->Win; d.nsrc = a->nsrc;
  d.src[0] = {reinterpret_cast<const __half*>(a->src0), a->c0, a->ld0};
  d.src[1] = {reinterpret_cast<const __half*>(a->src1), a->c1, a->ld1};
  d.ksize = a->ksize; d.stride = a->stride; d.pad = a->pad;
  d.w = reinterpret_cast<const __half*>(a->w); d.N = a->N; d.w_bstride = a->w_bstride;
  d.mode = a->mode; d.ups2 = a->ups2;
  d.out = a->out; d.out_ld = a->out_ld; d.out_bstride = a->out_bstride;
  d.bias = a->bias; d.bias_sel = a->bias_sel;
  d.res = reinterpret_cast<const __half*>(a->res); d.res_ld = a->res_ld; d.res_bstride = a->res_bstride;
  d.scale = a->scale; d.force_block_n = a->force_block_n;
  d.post_div = a->post_div == 0.f ? 1.f : a->post_div; d.n_store = a->n_store; d.out2 = a->out2; d.force_mt = a->force_mt; d.stats = a->stats; d.force_halo = a->force_halo; d.force_swap = a->force_swap;
  d.gn_ab = a->gn_ab; d.gn_silu = a->gn_silu;
  d.poly = a->poly;
  auto l = sdm::conv_gemm_build(d, sdm::device_sm_count());
  sdm::conv_gemm_run(*l, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}

int sdm_k_conv_tiles_per_image(int Hout, int Wout) { return sdm::conv_gemm_tiles_per_image(Hout, Wout); }
int sdm_k_conv_variant(int ksize, int stride, int mode, int ups2, int N, int has_res, int Hout, int Wout) {
  return sdm::conv_gemm_variant_code(ksize, stride, mode, ups2, N, has_res, Hout, Wout);
}

int sdm_k_conv_can_fuse_gn(int ksize, int stride, int mode, int ups2, int N, int has_res, int Hout, int Wout) {
  return sdm::conv_gemm_can_fuse_gn(ksize, stride, mode, ups2, N, has_res, Hout, Wout) ? 1 : 0;
}

int sdm_k_conv_can_poly(int N, int Hin, int Win) { return sdm::conv_gemm_can_poly(N, Hin, Win) ? 1 : 0; }

int sdm_k_attention(const sdm_attn_args* a, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::AttnDesc d;
  d.B = a->B; d.heads = a->heads; d.Lq = a->Lq; d.Lk = a->Lk;
  d.q = reinterpret_cast<const __half*>(a->q); d.ldq = a->ldq;
  d.k = reinterpret_cast<const __half*>(a->k); d.ldk = a->ldk;
  d.vt = reinterpret_cast<const __half*>(a->vt); d.ldvt = a->ldvt;
  d.bias = a->bias; d.bias_bstride = a->bias_bstride;
  d.out = reinterpret_cast<__half*>(a->out); d.ldo = a->ldo; d.scale = a->scale;
  d.ntiles = a->ntiles;
  auto l = sdm::attn_build(d);
  sdm::attn_run(*l, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}

int sdm_k_key_compact(const float* bias, float* cbias, int32_t* idx, int32_t* ntiles, int B, int L, int lpad, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::key_compact_level_run(bias, cbias, idx, ntiles, B, L, lpad, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}

int sdm_k_key_bias(const float* trimap, int B, int R, float* bias0, float* bias1, float* bias2, float* bias3, const int32_t* lpad4,
                   uintptr_t stream) {
  SDM_API_BEGIN
  SDM_CHECK(R % 64 == 0 && R >= 64, "R must be a multiple of 64");
  const int lp[4] = {lpad4[0], lpad4[1], lpad4[2], lpad4[3]};
  sdm::key_bias_run(trimap, B, R, bias0, bias1, bias2, bias3, lp, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}

int sdm_k_probe_halo(const void* x, const void* eye, float* out, int dy, int dx, int mode, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::probe_halo_run(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(eye), out, dy, dx, mode,
                      reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}

int sdm_k_gather_rows(const void* src, void* dst, const int32_t* idx, const int32_t* ntiles, int B, int L, int C, int idx_bstride,
                      uintptr_t stream) {
  SDM_API_BEGIN
  sdm::gather_rows_run(reinterpret_cast<const __half*>(src), reinterpret_cast<__half*>(dst), idx, ntiles, B, L, C, idx_bstride,
                       reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}

size_t sdm_k_groupnorm_scratch_floats(int B, int HW, int C) { return sdm::groupnorm_scratch_floats(B, HW, C); }
size_t sdm_k_groupnorm_ab_offset(int B, int HW, int C) { return sdm::groupnorm_ab_offset_floats(B, HW, C); }
int sdm_k_groupnorm(const sdm_groupnorm_args* a, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::GroupNormDesc d;
  d.B = a->B; d.HW = a->HW; d.nsrc = a->nsrc;
  d.src[0] = reinterpret_cast<const __half*>(a->src0); d.C[0] = a->c0; d.ld[0] = a->ld0;
  d.src[1] = reinterpret_cast<const __half*>(a->src1); d.C[1] = a->c1; d.ld[1] = a->ld1;
  d.gamma = a->gamma; d.beta = a->beta; d.eps = a->eps; d.silu = a->silu;
  d.out = reinterpret_cast<__half*>(a->out); d.scratch = a->scratch;
  d.pre_partial[0] = a->pre0; d.pre_partial[1] = a->pre1; d.pre_slots = a->pre_slots;
  SDM_CHECK(a->scratch_floats >= sdm::groupnorm_scratch_floats(a->B, a->HW, a->c0 + (a->nsrc > 1 ? a->c1 : 0)), "scratch too small");
  sdm::groupnorm_run(d, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
int sdm_k_layernorm(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int C, float eps,
                    uintptr_t stream) {
  SDM_API_BEGIN
  sdm::layernorm_run(reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), gamma, beta, rows, C, eps,
                     reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
int sdm_k_softmax_rows(const float* s, void* p, int64_t rows, int L, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::softmax_rows_run(s, reinterpret_cast<__half*>(p), rows, L, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}
int sdm_k_direct_conv(const sdm_direct_conv_args* a, uintptr_t stream) {
  SDM_API_BEGIN
  sdm::DirectConvDesc d;
  d.B = a->B; d.H = a->H; d.W = a->W; d.Cin = a->Cin; d.Cout = a->Cout; d.ksize = a->ksize;
  d.x = reinterpret_cast<const __half*>(a->x); d.x_ld = a->x_ld;
  d.w = reinterpret_cast<const __half*>(a->w); d.bias = a->bias;
  d.out = reinterpret_cast<__half*>(a->out); d.out_ld = a->out_ld; d.out_coff = a->out_coff;
  d.out_scale = a->out_scale; d.cout_limit = a->cout_limit;
  sdm::direct_conv_run(d, reinterpret_cast<cudaStream_t>(stream));
  SDM_API_END
}

}  // extern "C"
