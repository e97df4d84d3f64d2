// threefry2x32 + JAX's uniform / normal float32 conversions, shared by the stand-alone draws
// (mulan_rng.cu) and the in-kernel draws of mulan_fwd_pre_keyed (mulan_fwd_pre.cu).
// See mulan_rng.cu for the provenance of every formula (jax._src.prng / jax._src.random, XLA's
// float32 ErfInv).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace mulan {

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0,
                                             uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
#define MULAN_TF_ROUND(r)              \
  x0 += x1;                            \
  x1 = __funnelshift_l(x1, x1, (r));   \
  x1 ^= x0;
#define MULAN_TF_A MULAN_TF_ROUND(13) MULAN_TF_ROUND(15) MULAN_TF_ROUND(26) MULAN_TF_ROUND(6)
#define MULAN_TF_B MULAN_TF_ROUND(17) MULAN_TF_ROUND(29) MULAN_TF_ROUND(16) MULAN_TF_ROUND(24)
  x0 += k0; x1 += k1;
  MULAN_TF_A  x0 += k1; x1 += k2 + 1u;
  MULAN_TF_B  x0 += k2; x1 += k0 + 2u;
  MULAN_TF_A  x0 += k0; x1 += k1 + 3u;
  MULAN_TF_B  x0 += k1; x1 += k2 + 4u;
  MULAN_TF_A  x0 += k2; x1 += k0 + 5u;
#undef MULAN_TF_A
#undef MULAN_TF_B
#undef MULAN_TF_ROUND
}

__device__ __forceinline__ float bits_to_uniform(uint32_t bits, float minval, float span) {
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  return fmaxf(minval, f * span + minval);
}

// XLA ErfInv (float32), coefficients of Giles' single-precision approximation.
__device__ __forceinline__ float erf_inv_xla(float x) {
  float w = -log1pf(-x * x);
  const bool lt = w < 5.0f;
  w = lt ? w - 2.5f : sqrtf(w) - 3.0f;
  float p = lt ? 2.81022636e-08f : -0.000200214257f;
  p = (lt ? 3.43273939e-07f : 0.000100950558f) + p * w;
  p = (lt ? -3.5233877e-06f : 0.00134934322f) + p * w;
  p = (lt ? -4.39150654e-06f : -0.00367342844f) + p * w;
  p = (lt ? 0.00021858087f : 0.00573950773f) + p * w;
  p = (lt ? -0.00125372503f : -0.0076224613f) + p * w;
  p = (lt ? -0.00417768164f : 0.00943887047f) + p * w;
  p = (lt ? 0.246640727f : 1.00167406f) + p * w;
  p = (lt ? 1.50140941f : 2.83297682f) + p * w;
  return fabsf(x) == 1.0f ? x * INFINITY : p * x;
}

// jax.random.normal for one 32-bit word: sqrt(2) erf_inv(uniform on [nextafter(-1, 0), 1))
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;                       // nextafterf(-1, 0)
  return 1.41421354f * erf_inv_xla(bits_to_uniform(bits, lo, 1.0f - lo));
}

}  // namespace mulan
