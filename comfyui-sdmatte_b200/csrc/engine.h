// Engine: weight store + execution plan for the SDMatte single-pass matte path.
#pragma once
#include "sdmatte_b200.h"

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace sdm {
struct Engine;
Engine* engine_create(int device);
void engine_destroy(Engine* e);
void engine_load(Engine* e, const sdm_tensor_desc* tensors, int n);
void engine_load_report(Engine* e, int* n_used, int* n_unexpected);
size_t engine_workspace_bytes(Engine* e, int B, int R);
void engine_forward(Engine* e, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                    void* alpha_dev, void* premean_dev, void* ws, size_t ws_bytes, cudaStream_t st);
void engine_forward_prompt(Engine* e, const float* image_dev, const float* aux_dev, int B, int R, const int32_t* is_trans, int prompt_kind,
                           const float* coords_host, int ncoords, void* alpha_dev, void* premean_dev, void* ws, size_t ws_bytes,
                           cudaStream_t st);
size_t engine_workspace_bytes_prompt(Engine* e, int B, int R, int prompt_kind, int ncoords);
void engine_forward_host(Engine* e, const float* image_host, const float* trimap_host, int B, int R, const int32_t* is_trans,
                         void* alpha_host_f16, void* ws, size_t ws_bytes, cudaStream_t st);
size_t engine_node_workspace_bytes(Engine* e, int B, int H, int W, int R, int output_mode);
void engine_apply_host(Engine* e, const float* image_host, const float* trimap_host, int B, int H, int W, int R, const int32_t* is_trans,
                       int mask_refine, double trimap_constraint, int output_mode, void* alpha_out_host_f16, float* matted_out_host,
                       void* ws, size_t ws_bytes, cudaStream_t st);
void engine_forward_profiled(Engine* e, const float* image_dev, const float* trimap_dev, int B, int R, const int32_t* is_trans,
                             void* alpha_dev, void* ws, size_t ws_bytes, cudaStream_t st);
int engine_profile_count(Engine* e);
void engine_profile_entry(Engine* e, int i, char* kind, int kind_len, float* ms, double* flops, double* bytes);
void engine_stats(Engine* e, int* n_launches, double* tensor_flops);
void engine_debug_tensor(Engine* e, const char* name, void* dst_dev, size_t dst_bytes, int64_t* shape4, int* dtype);
int engine_debug_tensor_count(Engine* e);
const char* engine_debug_tensor_name(Engine* e, int i);
void engine_set_option(Engine* e, const char* name, int value);
void engine_graph_stats(Engine* e, int* captures, int* launches);
void engine_node_timing(Engine* e, double* ms4);
}  // namespace sdm
