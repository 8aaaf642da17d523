// HBM-bound normalisation kernels: GroupNorm(32)(+SiLU), LayerNorm, row softmax.
// Reference ops: nn.GroupNorm inside diffusers ResnetBlock2D / Transformer2DModel / VAE (eps 1e-5 / 1e-6),
// nn.LayerNorm in BasicTransformerBlock, softmax in the VAE mid-block attention.
// All statistics are fp32/fp64; inputs and outputs are fp16 NHWC (the reference's autocast rounding points, A.6).
#include "common.cuh"
#include "gn_math.cuh"
#include "kernels.h"

#include <algorithm>

namespace sdm {

// ------------------------------------------------------------------------------------------------
// GroupNorm
// ------------------------------------------------------------------------------------------------
static int gn_ny(int nvec) { return std::max(1, 256 / nvec); }
static int gn_nslab(int B, int HW, int Ctot) {
  // The slab partition must NOT depend on the batch size: the fp32 partial sums of a sample are then identical
  // whether it runs alone or inside a batch (bit-exact batch sharding across GPUs).
  (void)B;
  const int ny = gn_ny(Ctot / 8);
  const int by_rows = (HW + ny * 64 - 1) / (ny * 64);  // >= 64 pixels per thread
  return std::max(1, std::min(by_rows, 512));
}
size_t groupnorm_scratch_floats(int B, int HW, int Ctot) {
  return (size_t)B * gn_nslab(B, HW, Ctot) * Ctot * 2 + (size_t)B * Ctot * 2;
}

// partial[b][slab][c][2] = (sum, sumsq) over the slab's pixels
__global__ void gn_stats_kernel(const __half* __restrict__ s0, const __half* __restrict__ s1, int C0, int Ctot, long long ld0,
                                long long ld1, int HW, int pix_per_slab, float* __restrict__ partial) {
  extern __shared__ float red[];  // [ny][nvec*16]
  const int nvec = Ctot >> 3;
  const int v = threadIdx.x, y = threadIdx.y, ny = blockDim.y;
  const int b = blockIdx.y, slab = blockIdx.x;
  const int p0 = slab * pix_per_slab;
  const int p1 = min(HW, p0 + pix_per_slab);
  const int c = v * 8;
  const __half* base;
  long long ld;
  if (c < C0) { base = s0 + (long long)b * HW * ld0 + c; ld = ld0; }
  else { base = s1 + (long long)b * HW * ld1 + (c - C0); ld = ld1; }
  float sum[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sum[i] = 0.f; sq[i] = 0.f; }
  for (int p = p0 + y; p < p1; p += ny) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(base + (long long)p * ld));
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      sum[2 * j] += f.x; sq[2 * j] += f.x * f.x;
      sum[2 * j + 1] += f.y; sq[2 * j + 1] += f.y * f.y;
    }
  }
  float* mine = red + ((size_t)y * nvec + v) * 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) { mine[i] = sum[i]; mine[8 + i] = sq[i]; }
  __syncthreads();
  if (y == 0) {
    for (int yy = 1; yy < ny; ++yy) {
      const float* o = red + ((size_t)yy * nvec + v) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) { sum[i] += o[i]; sq[i] += o[8 + i]; }
    }
    float* dst = partial + (((size_t)b * gridDim.x + slab) * Ctot + c) * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dst[2 * i] = sum[i]; dst[2 * i + 1] = sq[i]; }
  }
}

// Pre-reduction of conv-epilogue partials: in [B][slots][C][2] -> out [B][nchunk][C][2], chunk k = slots [k*spc, (k+1)*spc).
// A 1024^2 tensor has 8192 slots per sample; gn_finalize (one CTA per (sample, group), 32-byte reads 1 KB apart) took
// 175 us on them (ncu r1k: 0.19 TB/s) — 40 % of the whole GroupNorm time.  Here every slot row (C x 8 bytes) is read
// fully coalesced by C/2 threads (float4 = two channels), accumulated in fp64 in a fixed order (deterministic, and the
// chunking depends only on the slot count, never on B: batch-invariant bits).
__global__ void __launch_bounds__(256) gn_reduce_partials_kernel(const float* __restrict__ in, float* __restrict__ out, int slots, int C,
                                                                 int spc) {
  extern __shared__ double red2[];  // [rows_par][C/2][4]
  const int b = blockIdx.y, k = blockIdx.x, nchunk = gridDim.x;
  const int tpr = C >> 1;               // threads per slot row
  const int rows_par = blockDim.x / tpr;
  const int t = threadIdx.x % tpr, rp = threadIdx.x / tpr;
  const int s0 = k * spc, s1 = min(slots, s0 + spc);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (rp < rows_par) {
    const float4* src = reinterpret_cast<const float4*>(in + ((size_t)b * slots) * C * 2) + t;
    for (int sl = s0 + rp; sl < s1; sl += rows_par) {
      const float4 v = __ldg(src + (size_t)sl * tpr);
      a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
    }
    double* mine = red2 + ((size_t)rp * tpr + t) * 4;
    mine[0] = a0; mine[1] = a1; mine[2] = a2; mine[3] = a3;
  }
  __syncthreads();
  if (rp == 0) {
    for (int r = 1; r < rows_par; ++r) {  // fixed order
      const double* o = red2 + ((size_t)r * tpr + t) * 4;
      a0 += o[0]; a1 += o[1]; a2 += o[2]; a3 += o[3];
    }
    reinterpret_cast<float4*>(out + (((size_t)b * nchunk + k) * C) * 2)[t] = make_float4((float)a0, (float)a1, (float)a2, (float)a3);
  }
}

// one 128-thread CTA per (b, group): reduce the slab partials in fp64 (fixed order => deterministic), emit per-channel
// scale/shift
template <int NT>
__global__ void __launch_bounds__(NT) gn_finalize_kernel(const float* __restrict__ part0, const float* __restrict__ part1, int C0,
                                                          int nslab, int Ctot, int HW, float eps, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ ab) {
  __shared__ double red[2][NT / 32];
  const int b = blockIdx.x >> 5, g = blockIdx.x & 31;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gs = Ctot >> 5;
  const int C1 = Ctot - C0;
  double s = 0.0, q = 0.0;
  const int n = nslab * gs;
  for (int i = threadIdx.x; i < n; i += NT) {
    const int slab = i / gs, c = g * gs + i % gs;
    const float* src = (c < C0) ? part0 + (((size_t)b * nslab + slab) * C0 + c) * 2
                                : part1 + (((size_t)b * nslab + slab) * C1 + (c - C0)) * 2;
    const float2 v = *reinterpret_cast<const float2*>(src);
    s += (double)v.x;
    q += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) { red[0][warp] = s; red[1][warp] = q; }
  __syncthreads();
  s = 0.0;
  q = 0.0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) { s += red[0][w]; q += red[1][w]; }  // fixed order: deterministic
  const double cnt = (double)HW * gs;
  const double mean = s / cnt;
  double var = q / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float fmean = (float)mean;
  for (int cc = threadIdx.x; cc < gs; cc += NT) {
    const int c = g * gs + cc;
    const float a = rstd * gamma[c];
    ab[((size_t)b * Ctot + c) * 2] = a;
    ab[((size_t)b * Ctot + c) * 2 + 1] = beta[c] - fmean * a;
  }
}

// apply: same (channel-vector, pixel-row) thread layout as the stats kernel, so each thread keeps the scale/shift of its
// 8 channels in registers and streams pixels with 4 independent 16-byte loads in flight (r1c ncu: the grid-stride version
// was latency/issue-bound at 3.4 TB/s: one load in flight per thread, 64 B of scale/shift re-fetched per vector).
// r1p: packed fp32x2 arithmetic (FFMA2/FMUL2/FADD2) and sign-folded constants.  Inside a step the SM clock sits at
// 1.3-1.4 GHz (power cap) and this kernel was issue-bound there (ncu r1k: 59 % issue-active at 5.4 TB/s unthrottled,
// 4.0 TB/s in the step): 12.5 -> 8.5 instructions per element.  SiLU(y) = y / (1 + 2^(-y log2 e)) with ONE MUFU op: the
// reciprocal of d = 1 + e is the bit-trick guess (negated for free through the magic constant) + two Newton steps,
//   n0 = -r0,  p1 = n0 (2 + d n0) = -r1,  p2 = p1 (2 + d p1) = -r2,  result = (-y) p2,
// with -y produced directly by the scale/shift FMA (negated per-channel constants).
// The arithmetic of pixel 0 is made to depend on all four loads of the iteration (a never-taken trap on the AND of their
// first words): ptxas otherwise sinks loads 2 and 3 below the arithmetic of pixel 0 — two 16-byte loads in flight per
// thread instead of four (volatile accesses did not help: it then parks them right before the first store).
// Measured r1p/r1q at 1024^2 x 128 ch (isolated, GB/s algorithmic): scalar v0 5.1-5.2, packed without the load join 4.5,
// packed + joined loads 5.5, scalar + joined loads 5.5; in the step 4.0-4.2 -> 4.5-4.7 TB/s.
// MAXT: launch bound (256 for up to 2048 channels at 8 per thread, 1024 beyond)
template <bool SILU, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? 3 : 1) gn_apply_kernel(const __half* __restrict__ s0, const __half* __restrict__ s1, int C0, int Ctot,
                                                          long long ld0, long long ld1, int HW, int pix_per_slab,
                                                          const float* __restrict__ ab, __half* __restrict__ out) {
  const int v = threadIdx.x, y = threadIdx.y, ny = blockDim.y;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_slab;
  const int p1 = min(HW, p0 + pix_per_slab);
  const int c = v * 8;
  const __half* base;
  long long ld;
  if (c < C0) { base = s0 + (long long)b * HW * ld0 + c; ld = ld0; }
  else { base = s1 + (long long)b * HW * ld1 + (c - C0); ld = ld1; }
  __half* obase = out + (long long)b * HW * Ctot + c;
  uint64_t ka[4], ks[4];  // channel pairs (2j, 2j+1); negated when SILU
  gn_load_consts<SILU>(ab + ((size_t)b * Ctot + c) * 2, ka, ks);
  auto emit = [&](const uint4& raw, __half* dst) { *reinterpret_cast<uint4*>(dst) = gn_piece<SILU>(raw, ka, ks); };
  // pointer-increment addressing: the 64-bit (pixel * ld) products were 7 instructions per 16-byte load
  int p = p0 + y;
  const __half* src = base + (long long)p * ld;
  __half* dst = obase + (long long)p * Ctot;
  const long long ss = (long long)ny * ld, ds = (long long)ny * Ctot;
#pragma unroll 1
  for (; p + 3 * ny < p1; p += 4 * ny) {
    uint4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = __ldg(reinterpret_cast<const uint4*>(src + u * ss));
    // four fp16 NaN payloads that a finite tensor can never hold all at once
    if (((r[0].x & r[1].x & r[2].x & r[3].x) & 0x7fff7fffu) == 0x7fff7fffu) __trap();
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(r[u], dst + u * ds);
    src += 4 * ss;
    dst += 4 * ds;
  }
#pragma unroll 1
  for (; p < p1; p += ny, src += ss, dst += ds) emit(__ldg(reinterpret_cast<const uint4*>(src)), dst);
}

// how the statistics of this GroupNorm are obtained: 0 = own stats pass, 1 = finalize the conv-epilogue partials directly,
// 2 = coalesced pre-reduction of the partials first (many slots)
static int gn_stats_path(const GroupNormDesc& d) {
  const int C0 = d.C[0];
  const int Ctot = d.C[0] + (d.nsrc > 1 ? d.C[1] : 0);
  if (!(d.pre_partial[0] && (d.nsrc == 1 || d.pre_partial[1]))) return 0;
  const int nchunk = std::min(64, gn_nslab(d.B, d.HW, Ctot));
  if (d.pre_slots >= 8 * nchunk && C0 % 2 == 0 && (Ctot - C0) % 2 == 0 && C0 <= 512 && (Ctot - C0) <= 512) return 2;
  return 1;
}
int groupnorm_num_launches(const GroupNormDesc& d) {
  const int path = gn_stats_path(d);
  return (path == 0 ? 3 : path == 1 ? 2 : 2 + d.nsrc) - (d.out ? 0 : 1);
}
// per-(sample, channel) (scale, shift) table [B][Ctot][2] inside the scratch buffer, valid after groupnorm_run
size_t groupnorm_ab_offset_floats(int B, int HW, int Ctot) { return (size_t)B * gn_nslab(B, HW, Ctot) * Ctot * 2; }
const float* groupnorm_ab(const GroupNormDesc& d) {
  return d.scratch + groupnorm_ab_offset_floats(d.B, d.HW, d.C[0] + (d.nsrc > 1 ? d.C[1] : 0));
}

void groupnorm_run(const GroupNormDesc& d, cudaStream_t st) {
  const int C0 = d.C[0];
  const int Ctot = d.C[0] + (d.nsrc > 1 ? d.C[1] : 0);
  SDM_CHECK(Ctot % 32 == 0 && C0 % 8 == 0 && Ctot % 8 == 0, "GroupNorm channel constraints");
  const int nvec = Ctot / 8;
  SDM_CHECK(nvec <= 1024, "GroupNorm: too many channels");
  const int ny = gn_ny(nvec);
  const int nslab = gn_nslab(d.B, d.HW, Ctot);
  const int pps = (d.HW + nslab - 1) / nslab;
  float* partial = d.scratch;
  float* ab = d.scratch + (size_t)d.B * nslab * Ctot * 2;
  const __half* s1 = d.nsrc > 1 ? d.src[1] : d.src[0];
  const long long ld1 = d.nsrc > 1 ? d.ld[1] : d.ld[0];
  const size_t smem = (size_t)ny * nvec * 16 * sizeof(float);
  static PerDeviceOnce attr_set;
  attr_set([] { SDM_CUDA_OK(cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); });
  const int path = gn_stats_path(d);
  if (path != 0) {
    // statistics were produced by the epilogue of the conv(s) that wrote the input: only reduce them
    // the thread count only depends on the per-sample slot count, never on B (batch-invariant reduction order)
    const int nchunk = std::min(64, nslab);
    if (path == 2) {
      // two-stage: coalesced pre-reduction of each source's partials into `nchunk` rows, then the per-group finalize
      const int spc = (d.pre_slots + nchunk - 1) / nchunk;
      float* red0 = partial;
      float* red1 = partial + (size_t)d.B * nchunk * C0 * 2;
      auto reduce = [&](const float* in, float* out, int C) {
        const int tpr = C / 2;
        const int threads = std::max(tpr, (256 / tpr) * tpr);
        const int rows_par = threads / tpr;
        gn_reduce_partials_kernel<<<dim3(nchunk, d.B), threads, (size_t)rows_par * tpr * 4 * sizeof(double), st>>>(in, out, d.pre_slots, C, spc);
        SDM_CUDA_OK(cudaGetLastError());
      };
      reduce(d.pre_partial[0], red0, C0);
      if (d.nsrc > 1) reduce(d.pre_partial[1], red1, Ctot - C0);
      gn_finalize_kernel<128><<<d.B * 32, 128, 0, st>>>(red0, d.nsrc > 1 ? red1 : red0, C0, nchunk, Ctot, d.HW, d.eps, d.gamma, d.beta, ab);
    } else if ((long long)d.pre_slots * (Ctot / 32) > 4096)
      gn_finalize_kernel<512><<<d.B * 32, 512, 0, st>>>(d.pre_partial[0], d.nsrc > 1 ? d.pre_partial[1] : d.pre_partial[0], C0, d.pre_slots,
                                                        Ctot, d.HW, d.eps, d.gamma, d.beta, ab);
    else
      gn_finalize_kernel<128><<<d.B * 32, 128, 0, st>>>(d.pre_partial[0], d.nsrc > 1 ? d.pre_partial[1] : d.pre_partial[0], C0, d.pre_slots,
                                                        Ctot, d.HW, d.eps, d.gamma, d.beta, ab);
  } else {
    gn_stats_kernel<<<dim3(nslab, d.B), dim3(nvec, ny), smem, st>>>(d.src[0], s1, C0, Ctot, d.ld[0], ld1, d.HW, pps, partial);
    SDM_CUDA_OK(cudaGetLastError());
    gn_finalize_kernel<128><<<d.B * 32, 128, 0, st>>>(partial, partial, Ctot, nslab, Ctot, d.HW, d.eps, d.gamma, d.beta, ab);
  }
  SDM_CUDA_OK(cudaGetLastError());
  if (!d.out) return;  // finalize only: the consuming conv normalises its resident input tile itself (groupnorm_ab)
  const int app_pps = ny * 16;  // 16 pixels per thread
  const int app_slabs = (d.HW + app_pps - 1) / app_pps;
  const dim3 ag(app_slabs, d.B), ab_(nvec, ny);
#define SDM_GN_APPLY_ARGS d.src[0], s1, C0, Ctot, d.ld[0], ld1, d.HW, app_pps, ab
  if (nvec * ny <= 256) {
    if (d.silu) gn_apply_kernel<true, 256><<<ag, ab_, 0, st>>>(SDM_GN_APPLY_ARGS, d.out);
    else gn_apply_kernel<false, 256><<<ag, ab_, 0, st>>>(SDM_GN_APPLY_ARGS, d.out);
  } else {
    if (d.silu) gn_apply_kernel<true, 1024><<<ag, ab_, 0, st>>>(SDM_GN_APPLY_ARGS, d.out);
    else gn_apply_kernel<false, 1024><<<ag, ab_, 0, st>>>(SDM_GN_APPLY_ARGS, d.out);
  }
#undef SDM_GN_APPLY_ARGS
  SDM_CUDA_OK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the channel dim: one warp per token row, two-pass from registers
// ------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, long long rows, int C, float eps) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = C >> 3;
  const uint4* src = reinterpret_cast<const uint4*>(x + row * C);
  float v[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 raw = __ldg(src + vi);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        v[i][2 * j] = f.x; v[i][2 * j + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float dlt = v[i][j] - mean; sq += dlt * dlt; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  uint4* dst = reinterpret_cast<uint4*>(y + row * C);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float4* g = reinterpret_cast<const float4*>(gamma + vi * 8);
      const float4* bt = reinterpret_cast<const float4*>(beta + vi * 8);
      const float4 g0 = __ldg(g), g1 = __ldg(g + 1), b0 = __ldg(bt), b1 = __ldg(bt + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        w[j] = pack_h2((v[i][2 * j] - mean) * rstd * gg[2 * j] + bb[2 * j],
                       (v[i][2 * j + 1] - mean) * rstd * gg[2 * j + 1] + bb[2 * j + 1]);
      dst[vi] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

void layernorm_run(const __half* x, __half* y, const float* gamma, const float* beta, long long rows, int C, float eps,
                   cudaStream_t st) {
  SDM_CHECK(C % 8 == 0 && C <= 8 * 32 * 5, "LayerNorm: C must be a multiple of 8 and <= 1280");
  const int wpb = 8;
  const unsigned blocks = (unsigned)((rows + wpb - 1) / wpb);
  const int nvec = C / 8;
  if (nvec <= 64) layernorm_kernel<2><<<blocks, wpb * 32, 0, st>>>(x, y, gamma, beta, rows, C, eps);
  else if (nvec <= 96) layernorm_kernel<3><<<blocks, wpb * 32, 0, st>>>(x, y, gamma, beta, rows, C, eps);
  else layernorm_kernel<5><<<blocks, wpb * 32, 0, st>>>(x, y, gamma, beta, rows, C, eps);
  SDM_CUDA_OK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// Row softmax fp32 -> fp16 (VAE mid-block attention probabilities). One CTA per row, row kept in registers.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, __half* __restrict__ p, int L) {
  __shared__ float red[8];
  __shared__ float bcast;
  const long long row = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(s + row * L);
  const int nv = L >> 2;
  float4 v[16];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nv) {
      v[i] = __ldg(src + vi);
      mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  const float l2e = 1.4426950408889634f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nv) {
      v[i].x = ex2f((v[i].x - mx) * l2e); v[i].y = ex2f((v[i].y - mx) * l2e);
      v[i].z = ex2f((v[i].z - mx) * l2e); v[i].w = ex2f((v[i].w - mx) * l2e);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = bcast;
  uint2* dst = reinterpret_cast<uint2*>(p + row * L);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nv) dst[vi] = make_uint2(pack_h2(v[i].x * inv, v[i].y * inv), pack_h2(v[i].z * inv, v[i].w * inv));
  }
}

void softmax_rows_run(const float* s, __half* p, long long rows, int L, cudaStream_t st) {
  SDM_CHECK(L % 4 == 0 && L <= 16384, "softmax_rows: L must be a multiple of 4 and <= 16384");
  SDM_CHECK(rows < (1ll << 31), "softmax_rows: too many rows");
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(s, p, L);
  SDM_CUDA_OK(cudaGetLastError());
}

}  // namespace sdm
