// HBM-bound "skinny" ops of the matte path: convs with <= 8 input or output channels, input preparation,
// per-level attention key-bias vectors, the alpha head.  (a1, a4, a5, a7, a14, a16.)
#include "common.cuh"
#include "kernels.h"

#include <algorithm>

namespace sdm {

// ------------------------------------------------------------------------------------------------
// conv with small Cin (4 or 8): thread -> (pixel, 8 output channels)
//   VAE conv_in (3->128, input padded to 4 ch), UNet conv_in (8->320), aux_conv_in (4->1024, utils.py:33-41),
//   decoder conv_in (4->512), quant/post_quant 1x1 convs.
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void conv_small_cin_kernel(const __half* __restrict__ x, long long x_ld, const __half* __restrict__ w,
                                      const float* __restrict__ bias, __half* __restrict__ out, long long out_ld, int out_coff,
                                      float out_scale, int cout_store, int B, int H, int W, int Cout, int ksize) {
  const int cg = Cout >> 3;
  const long long total = (long long)B * H * W * cg;
  const int taps = ksize * ksize;
  const int r = ksize >> 1;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    const long long pix = idx / cg;
    const int xx = (int)(pix % W);
    const int yy = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bias ? bias[g * 8 + i] : 0.f;
    for (int t = 0; t < taps; ++t) {
      const int iy = yy + t / ksize - r, ix = xx + t % ksize - r;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const __half* xp = x + (((long long)b * H + iy) * W + ix) * x_ld;
      float xv[CIN];
      if (CIN == 4) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(xp));
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
        const float2 a = __half22float2(h[0]), c = __half22float2(h[1]);
        xv[0] = a.x; xv[1] = a.y; xv[2] = c.x; xv[3] = c.y;
      } else {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(xp));
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); xv[2 * j] = f.x; xv[2 * j + 1] = f.y; }
      }
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const __half* wp = w + ((long long)(g * 8 + o) * taps + t) * CIN;
        if (CIN == 4) {
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(wp));
          const __half2* h = reinterpret_cast<const __half2*>(&raw);
          const float2 a = __half22float2(h[0]), c = __half22float2(h[1]);
          acc[o] += xv[0] * a.x + xv[1] * a.y + xv[2] * c.x + xv[3] * c.y;
        } else {
          const uint4 raw = __ldg(reinterpret_cast<const uint4*>(wp));
          const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); acc[o] += xv[2 * j] * f.x + xv[2 * j + 1] * f.y; }
        }
      }
    }
    __half hv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __half h = __float2half_rn(acc[i]);
      if (out_scale != 1.0f) h = __float2half_rn(__half2float(h) * out_scale);  // fp16 * python scalar -> fp16
      hv[i] = h;
    }
    __half* op = out + pix * out_ld + out_coff + g * 8;
    if (cout_store >= (g + 1) * 8) {
      *reinterpret_cast<uint4*>(op) = *reinterpret_cast<const uint4*>(hv);
    } else {
      for (int i = 0; i < 8; ++i)
        if (g * 8 + i < cout_store) op[i] = hv[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 conv with small Cin, tiled: CTA = 16x16 output pixels x 64 output channels.  Input halo and the
// [36 or 72][64] fp32 weight slab live in shared memory; each thread owns one pixel and 64 fp32 accumulators, the
// weight reads are warp-broadcast LDS.128 (4 FMAs per LDS).  Replaces the per-thread global weight loads of
// conv_small_cin_kernel for the heavy cases (VAE conv_in 3->128 @1024^2 was 8.7 ms per launch, ncu r1a).
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(128) conv_small_cin_tiled_kernel(const __half* __restrict__ x, long long x_ld,
                                                                   const __half* __restrict__ w, const float* __restrict__ bias,
                                                                   __half* __restrict__ out, long long out_ld, int out_coff, int B,
                                                                   int H, int W, int Cout, int tiles_x, int tiles_y) {
  constexpr int K = 9 * CIN;
  __shared__ __align__(16) float sW[K][64];
  __shared__ __align__(16) __half sIn[18 * 18][CIN];
  const int tile = blockIdx.x;
  const int tx0 = (tile % tiles_x) * 16, ty0 = ((tile / tiles_x) % tiles_y) * 16, b = tile / (tiles_x * tiles_y);
  const int c0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  for (int i = tid; i < K * 64; i += 128) {
    const int k = i / 64, c = i % 64;
    sW[k][c] = (c0 + c < Cout) ? __half2float(w[(long long)(c0 + c) * K + k]) : 0.f;
  }
  for (int i = tid; i < 18 * 18; i += 128) {
    const int yy = ty0 + i / 18 - 1, xx = tx0 + i % 18 - 1;
    const bool in = (yy >= 0 && yy < H && xx >= 0 && xx < W);
    const __half* src = x + (((long long)b * H + (in ? yy : 0)) * W + (in ? xx : 0)) * x_ld;
    if (CIN == 4) {
      uint2 v = in ? __ldg(reinterpret_cast<const uint2*>(src)) : make_uint2(0, 0);
      *reinterpret_cast<uint2*>(&sIn[i][0]) = v;
    } else {
      uint4 v = in ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(&sIn[i][0]) = v;
    }
  }
  __syncthreads();
  // each thread: 2 pixels (rows py and py+8 of the tile) x 64 output channels -> every weight LDS.128 feeds 8 FMAs
  const int px = tid & 15, py = tid >> 4;
  float acc[2][64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[0][c] = acc[1][c] = (bias && c0 + c < Cout) ? bias[c0 + c] : 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    float xv[2][CIN];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const __half* ip = &sIn[(py + u * 8 + t / 3) * 18 + px + t % 3][0];
#pragma unroll
      for (int j = 0; j < CIN / 2; ++j) {
        const float2 f = __half22float2(reinterpret_cast<const __half2*>(ip)[j]);
        xv[u][2 * j] = f.x; xv[u][2 * j + 1] = f.y;
      }
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float4* wr = reinterpret_cast<const float4*>(&sW[t * CIN + ci][0]);
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 w4 = wr[c4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          acc[u][c4 * 4 + 0] = fmaf(xv[u][ci], w4.x, acc[u][c4 * 4 + 0]);
          acc[u][c4 * 4 + 1] = fmaf(xv[u][ci], w4.y, acc[u][c4 * 4 + 1]);
          acc[u][c4 * 4 + 2] = fmaf(xv[u][ci], w4.z, acc[u][c4 * 4 + 2]);
          acc[u][c4 * 4 + 3] = fmaf(xv[u][ci], w4.w, acc[u][c4 * 4 + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int ox = tx0 + px, oy = ty0 + py + u * 8;
    if (ox < W && oy < H) {
      __half* op = out + (((long long)b * H + oy) * W + ox) * out_ld + out_coff + c0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (c0 + g * 8 < Cout)
          *reinterpret_cast<uint4*>(op + g * 8) =
              make_uint4(pack_h2(acc[u][g * 8 + 0], acc[u][g * 8 + 1]), pack_h2(acc[u][g * 8 + 2], acc[u][g * 8 + 3]),
                         pack_h2(acc[u][g * 8 + 4], acc[u][g * 8 + 5]), pack_h2(acc[u][g * 8 + 6], acc[u][g * 8 + 7]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// conv 3x3 with small Cout (<= 8): thread -> one pixel, all outputs.  Weights [COUT][9][Cin].
//   UNet conv_out (320->4), VAE encoder conv_out (512->8), decoder conv_out (128->3, via alpha head).
// ------------------------------------------------------------------------------------------------
template <int COUT>
__device__ __forceinline__ void small_cout_accumulate(const __half* __restrict__ x, long long x_ld, const __half* __restrict__ w,
                                                      int b, int yy, int xx, int H, int W, int Cin, float (&acc)[COUT]) {
  for (int t = 0; t < 9; ++t) {
    const int iy = yy + t / 3 - 1, ix = xx + t % 3 - 1;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    const uint4* xp = reinterpret_cast<const uint4*>(x + (((long long)b * H + iy) * W + ix) * x_ld);
    for (int c8 = 0; c8 < (Cin >> 3); ++c8) {
      const uint4 raw = __ldg(xp + c8);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
      float xv[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); xv[2 * j] = f.x; xv[2 * j + 1] = f.y; }
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const uint4 wr = __ldg(reinterpret_cast<const uint4*>(w + ((long long)o * 9 + t) * Cin) + c8);
        const __half2* wh = reinterpret_cast<const __half2*>(&wr);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(wh[j]); acc[o] += xv[2 * j] * f.x + xv[2 * j + 1] * f.y; }
      }
    }
  }
}

template <int COUT>
__global__ void conv_small_cout_kernel(const __half* __restrict__ x, long long x_ld, const __half* __restrict__ w,
                                       const float* __restrict__ bias, __half* __restrict__ out, long long out_ld, int out_coff,
                                       float out_div, int B, int H, int W, int Cin) {
  const long long total = (long long)B * H * W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(pix % W);
    const int yy = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = bias ? bias[o] : 0.f;
    small_cout_accumulate<COUT>(x, x_ld, w, b, yy, xx, H, W, Cin, acc);
    __half* op = out + pix * out_ld + out_coff;
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      __half hv = __float2half_rn(acc[o]);
      if (out_div != 1.0f) hv = __float2half_rn(__half2float(hv) / out_div);  // fp16 / python scalar -> fp16
      op[o] = hv;
    }
  }
}

void direct_conv_run(const DirectConvDesc& d, cudaStream_t st) {
  SDM_CHECK(d.ksize == 1 || d.ksize == 3, "direct conv ksize");
  const long long npix = (long long)d.B * d.H * d.W;
  if (d.Cin <= 8) {
    SDM_CHECK((d.Cin == 4 || d.Cin == 8) && d.Cout % 8 == 0, "small-Cin conv: Cin in {4,8}, Cout % 8 == 0");
    const int store = d.cout_limit > 0 ? d.cout_limit : d.Cout;
    if (d.ksize == 3 && d.Cout >= 64 && d.cout_limit == 0 && d.out_scale == 1.0f && (d.out_coff % 8) == 0 && (d.out_ld % 8) == 0) {
      const int tiles_x = (d.W + 15) / 16, tiles_y = (d.H + 15) / 16;
      const dim3 grid((unsigned)(tiles_x * tiles_y * d.B), (unsigned)((d.Cout + 63) / 64));
      if (d.Cin == 4)
        conv_small_cin_tiled_kernel<4><<<grid, 128, 0, st>>>(d.x, d.x_ld, d.w, d.bias, d.out, d.out_ld, d.out_coff, d.B, d.H, d.W, d.Cout, tiles_x, tiles_y);
      else
        conv_small_cin_tiled_kernel<8><<<grid, 128, 0, st>>>(d.x, d.x_ld, d.w, d.bias, d.out, d.out_ld, d.out_coff, d.B, d.H, d.W, d.Cout, tiles_x, tiles_y);
      SDM_CUDA_OK(cudaGetLastError());
      return;
    }
    const long long total = npix * (d.Cout / 8);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148ll * 32);
    if (d.Cin == 4)
      conv_small_cin_kernel<4><<<blocks, 256, 0, st>>>(d.x, d.x_ld, d.w, d.bias, d.out, d.out_ld, d.out_coff, d.out_scale, store,
                                                      d.B, d.H, d.W, d.Cout, d.ksize);
    else
      conv_small_cin_kernel<8><<<blocks, 256, 0, st>>>(d.x, d.x_ld, d.w, d.bias, d.out, d.out_ld, d.out_coff, d.out_scale, store,
                                                      d.B, d.H, d.W, d.Cout, d.ksize);
  } else {
    SDM_CHECK(d.ksize == 3 && d.Cin % 8 == 0 && (d.Cout == 4 || d.Cout == 8), "small-Cout conv: 3x3, Cout in {4,8}");
    SDM_CHECK(d.out_scale == 1.0f && d.cout_limit == 0, "small-Cout conv has no scale/limit");
    const int blocks = (int)std::min<long long>((npix + 127) / 128, 148ll * 32);
    if (d.Cout == 4)
      conv_small_cout_kernel<4><<<blocks, 128, 0, st>>>(d.x, d.x_ld, d.w, d.bias, d.out, d.out_ld, d.out_coff, d.out_div, d.B, d.H, d.W, d.Cin);
    else
      conv_small_cout_kernel<8><<<blocks, 128, 0, st>>>(d.x, d.x_ld, d.w, d.bias, d.out, d.out_ld, d.out_coff, d.out_div, d.B, d.H, d.W, d.Cin);
  }
  SDM_CUDA_OK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// alpha head: decoder conv_out (Cin->3) + mean over channels + clip + (x+1)/2   (meta_arch.py:256-260)
// rounding points as on the reference fp16 path: conv outputs fp16, mean fp16, (clip(x)+1) fp16, /2 exact
// ------------------------------------------------------------------------------------------------
__global__ void alpha_head_kernel(const __half* __restrict__ x, long long x_ld, const __half* __restrict__ w,
                                  const float* __restrict__ bias, __half* __restrict__ alpha, __half* __restrict__ premean, int B,
                                  int H, int W, int Cin) {
  const long long total = (long long)B * H * W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(pix % W);
    const int yy = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float acc[3] = {bias[0], bias[1], bias[2]};
    small_cout_accumulate<3>(x, x_ld, w, b, yy, xx, H, W, Cin, acc);
    const float c0 = __half2float(__float2half_rn(acc[0]));
    const float c1 = __half2float(__float2half_rn(acc[1]));
    const float c2 = __half2float(__float2half_rn(acc[2]));
    const __half m = __float2half_rn((c0 + c1 + c2) / 3.0f);
    if (premean) premean[pix] = m;
    const float cl = fminf(fmaxf(__half2float(m), -1.0f), 1.0f);
    const __half p1 = __float2half_rn(cl + 1.0f);
    alpha[pix] = __float2half_rn(__half2float(p1) * 0.5f);
  }
}

void alpha_head_run(const __half* x, long long x_ld, int B, int H, int W, int Cin, const __half* w, const float* bias,
                    __half* alpha, __half* premean, cudaStream_t st) {
  SDM_CHECK(Cin % 8 == 0, "alpha head Cin");
  const long long npix = (long long)B * H * W;
  const int blocks = (int)std::min<long long>((npix + 127) / 128, 148ll * 32);
  alpha_head_kernel<<<blocks, 128, 0, st>>>(x, x_ld, w, bias, alpha, premean, B, H, W, Cin);
  SDM_CUDA_OK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// alpha head as GEMM + col2im (r1p).  The tensor-core GEMM leaves, for every INPUT pixel q, the 27 partial products
// y[q][tap*3 + c] = sum_ch w[c][tap][ch] * x[q][ch]; the 3x3 conv output at p is sum_tap y[p + off(tap)][tap*3 + c].
// CTA = 32 x 8 output pixels; the (34 x 10) halo of y rows is staged in shared memory with fully coalesced 128-byte row
// reads (pixel stride 33 floats: conflict-free column reads), zero outside the image (= the conv's zero padding).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) alpha_col2im_kernel(const float* __restrict__ y, const float* __restrict__ bias, int H, int W,
                                                           __half* __restrict__ alpha, __half* __restrict__ premean) {
  __shared__ float sm[10 * 34 * 33];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  const float* yb = y + (size_t)b * H * W * 32;
  for (int i = threadIdx.x; i < 340 * 8; i += 256) {
    const int px = i >> 3, q = i & 7;
    const int hy = px / 34, hx = px - hy * 34;
    const int gy = y0 + hy - 1, gx = x0 + hx - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(reinterpret_cast<const float4*>(yb + ((size_t)gy * W + gx) * 32) + q);
    float* d = sm + px * 33 + q * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx >= W || gy >= H) return;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float* s = sm + ((ty + t / 3) * 34 + tx + t % 3) * 33 + t * 3;
    acc[0] += s[0]; acc[1] += s[1]; acc[2] += s[2];
  }
  // rounding points of the reference fp16 path: conv outputs fp16, channel mean fp16, (clip+1) fp16, /2 exact
  const float c0 = __half2float(__float2half_rn(acc[0] + bias[0]));
  const float c1 = __half2float(__float2half_rn(acc[1] + bias[1]));
  const float c2 = __half2float(__float2half_rn(acc[2] + bias[2]));
  const __half m = __float2half_rn((c0 + c1 + c2) / 3.0f);
  const size_t o = ((size_t)b * H + gy) * W + gx;
  if (premean) premean[o] = m;
  const float cl = fminf(fmaxf(__half2float(m), -1.0f), 1.0f);
  const __half p1 = __float2half_rn(cl + 1.0f);
  alpha[o] = __float2half_rn(__half2float(p1) * 0.5f);
}
void alpha_col2im_run(const float* y, const float* bias, int B, int H, int W, __half* alpha, __half* premean, cudaStream_t st) {
  alpha_col2im_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, B), 256, 0, st>>>(y, bias, H, W, alpha, premean);
  SDM_CUDA_OK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// input preparation (sdmatte_nodes.py:343,351 at native resolution; meta_arch.py:141)
// ------------------------------------------------------------------------------------------------
// Writes the im2col matrix of the VAE conv_in (3x3, 3 input channels): out[pixel][k], k = tap*4 + channel (k >= 36 zero),
// so that conv_in becomes ONE 64-wide K step of the tcgen05 GEMM instead of a SIMT direct convolution (9 ms -> ~2 ms, r1e).
// rows [0, B*R*R): normalised image (x-0.5)/0.5; rows [B*R*R, 2*B*R*R): trimap*2-1 replicated to 3 channels.
__global__ void prep_inputs_kernel(const float* __restrict__ image, const float* __restrict__ trimap, __half* __restrict__ out,
                                   int R, long long npix_b /* B*R*R */) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * npix_b; i += (long long)gridDim.x * blockDim.x) {
    const bool is_img = i < npix_b;
    const long long q = is_img ? i : i - npix_b;
    const int x = (int)(q % R), y = (int)((q / R) % R);
    const long long img0 = q - ((long long)y * R + x);  // first pixel of this image
    uint32_t w[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) w[k] = 0u;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      float a = 0.f, b = 0.f, c = 0.f;
      if (yy >= 0 && yy < R && xx >= 0 && xx < R) {
        const long long s = img0 + (long long)yy * R + xx;
        if (is_img) {
          a = (image[s * 3] - 0.5f) / 0.5f;
          b = (image[s * 3 + 1] - 0.5f) / 0.5f;
          c = (image[s * 3 + 2] - 0.5f) / 0.5f;
        } else {
          a = b = c = trimap[s] * 2.0f - 1.0f;
        }
      }
      w[t * 2] = pack_h2(a, b);
      w[t * 2 + 1] = pack_h2(c, 0.f);
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 64);
#pragma unroll
    for (int g = 0; g < 8; ++g) o[g] = make_uint4(w[g * 4], w[g * 4 + 1], w[g * 4 + 2], w[g * 4 + 3]);
  }
}
void prep_inputs_run(const float* image, const float* trimap, __half* out, int ldc, int B, int R, cudaStream_t st) {
  SDM_CHECK(ldc == 64, "prep writes 64-wide im2col rows");
  const long long n = (long long)B * R * R;
  const int blocks = (int)std::min<long long>((2 * n + 127) / 128, 148ll * 64);
  prep_inputs_kernel<<<blocks, 128, 0, st>>>(image, trimap, out, R, n);
  SDM_CUDA_OK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// additive key bias per UNet level (meta_arch.py:200-204 -> replace.py:401-403 -> replace.py:56-63)
// ------------------------------------------------------------------------------------------------
struct KeyBiasArgs {
  float* dst[4];
  int lpad[4];
};
__global__ void key_bias_kernel(const float* __restrict__ trimap, int B, int R, KeyBiasArgs a) {
  const int S = R >> 3;
  const int level = blockIdx.y;
  const int s = S >> level;
  const int step = 8 << level;
  const int lp = a.lpad[level];
  const long long total = (long long)B * lp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / lp);
    const int k = (int)(i % lp);
    float v = -INFINITY;
    if (k < s * s) {
      const int yy = (k / s) * step, xx = (k % s) * step;
      const float t = trimap[((long long)b * R + yy) * R + xx];
      const float tri = t * 2.0f - 1.0f;
      const float m = (tri + 1.0f) / 2.0f;
      v = ((1.0f - m) * -10000.0f) * 1.4426950408889634f;  // stored in the log2 domain for the attention kernel
    }
    a.dst[level][i] = v;
  }
}
void key_bias_run(const float* trimap, int B, int R, float* bias0, float* bias1, float* bias2, float* bias3, const int* lpad,
                  cudaStream_t st) {
  KeyBiasArgs a;
  a.dst[0] = bias0; a.dst[1] = bias1; a.dst[2] = bias2; a.dst[3] = bias3;
  for (int i = 0; i < 4; ++i) a.lpad[i] = lpad[i];
  const long long n = (long long)B * lpad[0];
  const int blocks = (int)std::min<long long>((n + 255) / 256, 1024);
  key_bias_kernel<<<dim3(blocks, 4), 256, 0, st>>>(trimap, B, R, a);
  SDM_CUDA_OK(cudaGetLastError());
}


// ------------------------------------------------------------------------------------------------
// Self-attention key compaction.
// The additive key bias of attn1 is (1 - mask) * -10000 (replace.py:401-403): as soon as a sample has ONE key with mask 1
// (definite foreground), every key with a smaller mask gets exp(-5000) or exp(-10000) = exactly 0 in the reference's fp32
// softmax.  Those keys contribute nothing to either the row sum or P.V, so the engine drops them: the kept keys
// {k : bias_k >= max_k(bias) - kDropMargin} are gathered (in order) in front, padded to a multiple of 128 with bias -inf,
// and the attention kernel streams ntiles[b] = ceil(kept/128) key tiles instead of L/128.
// Exact whenever |scale * q.k| < 1190 for all pairs (then a dropped key's probability is < exp(2*1190 - 2500) = 2^-173 < the
// smallest fp32 denormal relative to the kept maximum; LayerNorm-ed activations give |scale q.k| of order 10).
// One CTA per (sample, level); deterministic ordered compaction (block scan), per-sample only => batch-invariant.
// ------------------------------------------------------------------------------------------------
struct KeyCompactArgs {
  const float* bias[4];
  float* cbias[4];
  int* idx[4];
  int* ntiles[4];
  int lpad[4];
  int L[4];
};
constexpr float kDropMarginLog2 = 2500.0f * 1.4426950408889634f;

__global__ void __launch_bounds__(1024) key_compact_kernel(KeyCompactArgs a) {
  __shared__ float redf[32];
  __shared__ int redi[32];
  __shared__ int total_s;
  const int level = blockIdx.y, b = blockIdx.x;
  const int L = a.L[level], lp = a.lpad[level];
  const float* bias = a.bias[level] + (size_t)b * lp;
  float* cb = a.cbias[level] + (size_t)b * lp;
  int* idx = a.idx[level] + (size_t)b * lp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -INFINITY;
  for (int k = tid; k < L; k += 1024) mx = fmaxf(mx, bias[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) redf[warp] = mx;
  __syncthreads();
  mx = redf[0];
#pragma unroll
  for (int w = 1; w < 32; ++w) mx = fmaxf(mx, redf[w]);
  const float thr = mx - kDropMarginLog2;
  // thread t owns the contiguous run [t*per, (t+1)*per): ordered compaction = exclusive scan of the per-run counts
  const int per = (L + 1023) / 1024;
  const int k0 = tid * per, k1 = min(L, k0 + per);
  int cnt = 0;
  for (int k = k0; k < k1; ++k) cnt += (bias[k] >= thr) ? 1 : 0;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) redi[warp] = incl;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += redi[w];
  if (tid == 1023) total_s = wbase + incl;
  int pos = wbase + incl - cnt;
  for (int k = k0; k < k1; ++k) {
    const float v = bias[k];
    if (v >= thr) { idx[pos] = k; cb[pos] = v; ++pos; }
  }
  __syncthreads();
  const int total = total_s;                 // >= 1: the maximum itself is kept
  const int padded = (total + 127) & ~127;   // <= lp
  const int first = idx[0];
  for (int i = total + tid; i < padded; i += 1024) { idx[i] = first; cb[i] = -INFINITY; }  // valid row, probability 0
  if (tid == 0) a.ntiles[level][b] = padded >> 7;
}
void key_compact_run(const float* const* bias, float* const* cbias, int* const* idx, int* const* ntiles, const int* lpad, int B, int S,
                     cudaStream_t st) {
  KeyCompactArgs a;
  for (int i = 0; i < 4; ++i) {
    a.bias[i] = bias[i]; a.cbias[i] = cbias[i]; a.idx[i] = idx[i]; a.ntiles[i] = ntiles[i]; a.lpad[i] = lpad[i];
    a.L[i] = (S >> i) * (S >> i);
  }
  key_compact_kernel<<<dim3(B, 4), 1024, 0, st>>>(a);
  SDM_CUDA_OK(cudaGetLastError());
}
void key_compact_level_run(const float* bias, float* cbias, int* idx, int* ntiles, int B, int L, int lpad, cudaStream_t st) {
  SDM_CHECK(lpad % 128 == 0 && lpad >= L && L > 0, "key_compact: lpad must be a multiple of 128 and >= L");
  KeyCompactArgs a{};
  a.bias[0] = bias; a.cbias[0] = cbias; a.idx[0] = idx; a.ntiles[0] = ntiles; a.lpad[0] = lpad; a.L[0] = L;
  key_compact_kernel<<<dim3(B, 1), 1024, 0, st>>>(a);
  SDM_CUDA_OK(cudaGetLastError());
}

// dst[b][i][:] = src[b][idx[b][i]][:] for i < 128 * ntiles[b]   (rows of C fp16, C % 8 == 0); one warp per row
__global__ void __launch_bounds__(256) gather_rows_kernel(const __half* __restrict__ src, __half* __restrict__ dst, const int* __restrict__ idx,
                                                          const int* __restrict__ ntiles, int L, int C, int idx_bstride) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= (ntiles[b] << 7) || i >= L) return;
  const int k = idx[(size_t)b * idx_bstride + i];
  const uint4* s = reinterpret_cast<const uint4*>(src + ((size_t)b * L + k) * C);
  uint4* d = reinterpret_cast<uint4*>(dst + ((size_t)b * L + i) * C);
  for (int v = threadIdx.x & 31; v < (C >> 3); v += 32) d[v] = __ldg(s + v);
}
void gather_rows_run(const __half* src, __half* dst, const int* idx, const int* ntiles, int B, int L, int C, int idx_bstride,
                     cudaStream_t st) {
  SDM_CHECK(C % 8 == 0, "gather_rows: C must be a multiple of 8");
  gather_rows_kernel<<<dim3((L + 7) / 8, B), 256, 0, st>>>(src, dst, idx, ntiles, L, C, idx_bstride);
  SDM_CUDA_OK(cudaGetLastError());
}

}  // namespace sdm
