// Node-side pre/post-processing on the GPU: everything the reference does either side of the model
// call in SDMatteApply.apply_matte, so that one H2D of the caller's tensors and one D2H of the results bracket the path.
//
//   pre  : torchvision Resize(antialias=True) of image and trimap to R x R
//            /root/reference/sdmatte_nodes.py:204-214 (_resize_norm_image_bchw / _resize_mask_b1hw), :343, :349-351
//          (the (x-0.5)/0.5 and *2-1 affine maps live in prep_inputs_kernel, small_ops.cu)
//   post : Resize back to (orig_h, orig_w) -> clamp(0,1)                       sdmatte_nodes.py:362-363
//          mask_refine (bg -> 0, fg -> clamp(1.2 a), unknown & a < 0.3 -> 0)    sdmatte_nodes.py:365-380
//          output_mode composition (zeros / cat[img, a] / img * keep)           sdmatte_nodes.py:384-397
//
// The resize is torch's `upsample_bilinear2d_aa` (align_corners=False, sizes given): a separable triangle filter whose
// support grows with the down-scaling factor; per output index i
//   scale = in/out, support = max(scale, 1), center = scale*(i+0.5), invscale = 1/max(scale, 1)
//   xmin = max(int(center - support + 0.5), 0), xsize = min(int(center + support + 0.5), in) - xmin
//   w_j = tri((j + xmin - center + 0.5) * invscale) / sum_j(...)
// and out = sum_y wy * (sum_x wx * src), all in fp32 (the fp16 alpha is promoted, the result rounded back to fp16, as
// torchvision does for half tensors).  HBM-bound: one thread per output pixel, x fastest.
#include "common.cuh"
#include "kernels.h"

namespace sdm {

struct AaSpan {
  int lo, n;
  float center, invscale, total;
};

__device__ __forceinline__ float aa_tri(float x) {
  x = fabsf(x);
  return x < 1.0f ? 1.0f - x : 0.0f;
}

__device__ __forceinline__ AaSpan aa_span(int i, int in_size, float scale) {
  AaSpan s;
  const float support = scale >= 1.0f ? scale : 1.0f;
  s.center = scale * ((float)i + 0.5f);
  s.invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  s.lo = max((int)(s.center - support + 0.5f), 0);
  s.n = min((int)(s.center + support + 0.5f), in_size) - s.lo;
  float t = 0.f;
  for (int j = 0; j < s.n; ++j) t += aa_tri(((float)j + ((float)s.lo - s.center) + 0.5f) * s.invscale);
  s.total = t;
  return s;
}

__device__ __forceinline__ float aa_weight(const AaSpan& s, int j) {
  const float w = aa_tri(((float)j + ((float)s.lo - s.center) + 0.5f) * s.invscale);
  return s.total != 0.f ? w / s.total : w;
}

__device__ __forceinline__ float ld_f(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_f(const __half* p) { return __half2float(__ldg(p)); }

// one output pixel (oy, ox) of a [Hin][Win][C] image
template <typename Tin, int C>
__device__ __forceinline__ void aa_pixel(const Tin* __restrict__ src, int Hin, int Win, float sy, float sx, int oy, int ox,
                                         float (&out)[C]) {
  const AaSpan ys = aa_span(oy, Hin, sy);
  const AaSpan xs = aa_span(ox, Win, sx);
#pragma unroll
  for (int c = 0; c < C; ++c) out[c] = 0.f;
  for (int jy = 0; jy < ys.n; ++jy) {
    const float wy = aa_weight(ys, jy);
    const Tin* row = src + ((long long)(ys.lo + jy) * Win + xs.lo) * C;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int jx = 0; jx < xs.n; ++jx) {
      const float wx = aa_weight(xs, jx);
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += ld_f(row + jx * C + c) * wx;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] += acc[c] * wy;
  }
}

template <int C>
__global__ void __launch_bounds__(256) resize_aa_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int Hin, int Win,
                                                           int Hout, int Wout, float sy, float sx) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  const int b = blockIdx.z;
  if (ox >= Wout) return;
  float v[C];
  aa_pixel<float, C>(src + (long long)b * Hin * Win * C, Hin, Win, sy, sx, oy, ox, v);
  float* o = dst + (((long long)b * Hout + oy) * Wout + ox) * C;
#pragma unroll
  for (int c = 0; c < C; ++c) o[c] = v[c];
}

// alpha [B][R][R] fp16 -> alpha_out [B][H][W] fp16 and matted [B][H][W][3|4] fp32
__global__ void __launch_bounds__(256) postprocess_kernel(const __half* __restrict__ alpha, int R, const float* __restrict__ image,
                                                         const float* __restrict__ trimap, int H, int W, float sy, float sx,
                                                         int identity, int refine, float c_fg, float c_bg, int mode,
                                                         __half* __restrict__ alpha_out, float* __restrict__ matted) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  const int b = blockIdx.z;
  if (ox >= W) return;
  const long long pix = ((long long)b * H + oy) * W + ox;
  __half ah;
  if (identity) {
    ah = alpha[pix];
  } else {
    float v[1];
    aa_pixel<__half, 1>(alpha + (long long)b * R * R, R, R, sy, sx, oy, ox, v);
    ah = __float2half_rn(v[0]);  // torchvision: half -> float -> interpolate -> half
  }
  float a = fminf(fmaxf(__half2float(ah), 0.0f), 1.0f);  // .clamp(0, 1) on the fp16 tensor (0 and 1 are fp16 values)
  const float t = trimap ? __ldg(trimap + pix) : 0.f;
  if (refine) {
    const bool fg = t > c_fg, bg = t < c_bg;
    if (bg) a = 0.0f;
    if (fg) a = fminf(fmaxf(__half2float(__float2half_rn(a * 1.2f)), 0.0f), 1.0f);  // fp16 tensor * 1.2 -> fp16, then clamp
    if (a < 0.3f && !(fg || bg)) a = 0.0f;  // a is an fp16 value: a < 0.3 (real) <=> a < half(0.3)
  }
  alpha_out[pix] = __float2half_rn(a);  // exact: a is already representable
  if (mode == 0 || matted == nullptr) return;  // alpha_only: the caller materialises the zeros (sdmatte_nodes.py:384-385)
  const float* ip = image + pix * 3;
  const float r = __ldg(ip), g = __ldg(ip + 1), bl = __ldg(ip + 2);
  if (mode == 1) {  // matted_rgba: cat([image, alpha]) (fp32 <- fp16 promotion)
    float4* o = reinterpret_cast<float4*>(matted) + pix;
    *o = make_float4(r, g, bl, a);
  } else {
    // matted_rgb: image * ((trimap > 0.2) & (alpha > 0.1));  anything else: image * alpha (unreachable from the schema)
    const float k = (mode == 2) ? ((t > 0.2f && a > 0.1f) ? 1.0f : 0.0f) : a;
    float* o = matted + pix * 3;
    o[0] = r * k; o[1] = g * k; o[2] = bl * k;
  }
}

void preprocess_run(const float* image, const float* trimap, int B, int H, int W, int R, float* image_out, float* trimap_out,
                    cudaStream_t st) {
  SDM_CHECK(B > 0 && H > 0 && W > 0 && R > 0, "preprocess dims");
  const float sy = (float)H / (float)R, sx = (float)W / (float)R;
  const dim3 grid((R + 255) / 256, R, B);
  resize_aa_f32_kernel<3><<<grid, 256, 0, st>>>(image, image_out, H, W, R, R, sy, sx);
  SDM_CUDA_OK(cudaGetLastError());
  resize_aa_f32_kernel<1><<<grid, 256, 0, st>>>(trimap, trimap_out, H, W, R, R, sy, sx);
  SDM_CUDA_OK(cudaGetLastError());
}

void postprocess_run(const __half* alpha, int B, int R, int H, int W, const float* image, const float* trimap, int mask_refine,
                     double trimap_constraint, int output_mode, __half* alpha_out, float* matted_out, cudaStream_t st) {
  SDM_CHECK(B > 0 && H > 0 && W > 0 && R > 0, "postprocess dims");
  SDM_CHECK(output_mode >= 0 && output_mode <= 3, "output_mode: 0 alpha_only, 1 matted_rgba, 2 matted_rgb, 3 image*alpha");
  SDM_CHECK(!(mask_refine || output_mode == 2) || trimap != nullptr, "trimap needed for mask_refine / matted_rgb");
  SDM_CHECK(output_mode == 0 || (image != nullptr && matted_out != nullptr), "image and matted_out needed for this output_mode");
  const float sy = (float)R / (float)H, sx = (float)R / (float)W;
  // python: trimap > c  and  trimap < (1.0 - c): the double scalars are cast to the tensor's float32
  const float c_fg = (float)trimap_constraint, c_bg = (float)(1.0 - trimap_constraint);
  const dim3 grid((W + 255) / 256, H, B);
  postprocess_kernel<<<grid, 256, 0, st>>>(alpha, R, image, trimap, H, W, sy, sx, (H == R && W == R) ? 1 : 0, mask_refine, c_fg, c_bg,
                                            output_mode, alpha_out, matted_out);
  SDM_CUDA_OK(cudaGetLastError());
}

}  // namespace sdm
