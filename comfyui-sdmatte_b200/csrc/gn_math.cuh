// GroupNorm(+SiLU) arithmetic on one 16-byte piece (8 channels of one pixel), shared by the stand-alone apply pass (norms.cu)
// and by the transform warps that normalise a convolution's resident input tile in shared memory (conv_swap_halo.cu): both
// produce the same fp16 bits, so fusing the normalisation into the consuming conv does not move the result.
// Reference op: nn.GroupNorm(32) -> SiLU in diffusers ResnetBlock2D (fp32 under autocast, re-cast to fp16 by the conv).
#pragma once
#include "common.cuh"

namespace sdm {

// per-thread constants of 8 consecutive channels: ab = (scale, shift) pairs as gn_finalize_kernel writes them.
// ka[j] / ks[j] hold channel pair (2j, 2j+1); negated when SILU (the scale/shift FMA then yields -y directly).
template <bool SILU>
__device__ __forceinline__ void gn_load_consts(const float* __restrict__ ab8, uint64_t (&ka)[4], uint64_t (&ks)[4]) {
  const float4* abp = reinterpret_cast<const float4*>(ab8);
  const float sg = SILU ? -1.0f : 1.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 k = __ldg(abp + j);  // (a0, s0, a1, s1)
    ka[j] = pack_f2(sg * k.x, sg * k.z);
    ks[j] = pack_f2(sg * k.y, sg * k.w);
  }
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// the same constants from four raw (a0, s0, a1, s1) quads loaded earlier (software-pipelined loads: the transform warps of the
// fused conv fetch the next slice's constants while they normalise the current one)
template <bool SILU>
__device__ __forceinline__ void gn_consts_from_raw(const float4 (&raw)[4], uint64_t (&ka)[4], uint64_t (&ks)[4]) {
  const float sg = SILU ? -1.0f : 1.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ka[j] = pack_f2(sg * raw[j].x, sg * raw[j].z);
    ks[j] = pack_f2(sg * raw[j].y, sg * raw[j].w);
  }
}

// y = a x + s, optionally SiLU(y) = y / (1 + 2^(-y log2 e)).
// MUFU_RCP = false (the HBM-bound apply pass, which two MUFU ops per element made XU-bound, ncu r1a): ONE MUFU op per element,
//   the reciprocal of d = 1 + e is the bit-trick guess (negated for free through the magic constant) + two Newton steps,
//   n0 = -r0,  p1 = n0 (2 + d n0) = -r1,  p2 = p1 (2 + d p1) = -r2,  result = (-y) p2.
// MUFU_RCP = true (the transform warps of the fused conv, which are ISSUE-bound, kbench r2d: the arithmetic alone took the fused
//   128 -> 128 conv from 1.15 to 1.44 ms): MUFU.RCP instead of the Newton chain and no overflow clamp (e = +inf gives d = +inf,
//   1/d = 0, result -0: the correct limit) — 5.5 instead of 8.5 issue slots per element.  Both forms are accurate to < 1e-5
//   relative, far below the fp16 rounding of the result; they are not bit-identical to each other.
template <bool SILU, bool MUFU_RCP = false>
__device__ __forceinline__ uint4 gn_piece(const uint4& raw, const uint64_t (&ka)[4], const uint64_t (&ks)[4]) {
  const uint64_t kLog2e = pack_f2(1.4426950408889634f, 1.4426950408889634f);
  const uint64_t kOne = pack_f2(1.0f, 1.0f), kTwo = pack_f2(2.0f, 2.0f);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    uint64_t r = fma_f2(pack_f2(f.x, f.y), ka[j], ks[j]);  // y, or -y when SILU
    if (SILU && MUFU_RCP) {
      float t0, t1;
      unpack_f2(mul_f2(r, kLog2e), t0, t1);                // -y log2(e)
      float d0, d1;
      unpack_f2(add_f2(pack_f2(ex2f(t0), ex2f(t1)), kOne), d0, d1);
      r = mul_f2(r, pack_f2(-rcp_approx(d0), -rcp_approx(d1)));
    } else if (SILU) {
      float t0, t1;
      unpack_f2(mul_f2(r, kLog2e), t0, t1);                // -y log2(e)
      const uint64_t d = add_f2(pack_f2(ex2f(fminf(t0, 80.0f)), ex2f(fminf(t1, 80.0f))), kOne);
      float d0, d1;
      unpack_f2(d, d0, d1);
      uint64_t n = pack_f2(__int_as_float(0xFEF311C7 - __float_as_int(d0)), __int_as_float(0xFEF311C7 - __float_as_int(d1)));
      n = mul_f2(n, fma_f2(d, n, kTwo));
      n = mul_f2(n, fma_f2(d, n, kTwo));
      r = mul_f2(r, n);
    }
    float y0, y1;
    unpack_f2(r, y0, y1);
    w[j] = pack_h2(y0, y1);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace sdm
