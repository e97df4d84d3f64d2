// Shared device helpers for the MuLAN schedule + ELBO kernels (sm_100a).
//
// Compiled with -fmad=false: every `a*b+c` written with plain operators rounds twice,
// exactly like the reference's op-by-op float32 evaluation (and the CPU oracle).  Fused
// multiply-adds appear only where written explicitly as fmaf(), i.e. where the extra
// rounding of the reference is pseudo-random per sub-pixel and averages out of the
// per-example sums (the polynomial), never where a systematic or amplified error would
// show (1 - sigmoid(g), the reconstruction logits, the prior-KL summand).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mulan {

constexpr int kThreads = 256;            // one CTA per example row
constexpr int kWarps = kThreads / 32;
constexpr float kThird = 0.333333343267440796f;  // RN(1/3)
constexpr float kFifth = 0.200000002980232239f;  // RN(1/5)

// ---------------------------------------------------------------------------------------
// Memory access: 16-byte vector loads/stores on the read-only / streaming paths.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* __restrict__ p, int i4) {
  return __ldg(reinterpret_cast<const float4*>(p) + i4);
}
__device__ __forceinline__ void st4(float* __restrict__ p, int i4, float4 v) {
  reinterpret_cast<float4*>(p)[i4] = v;
}
__device__ __forceinline__ uchar4 ldx4(const uint8_t* __restrict__ p, int i4) {
  return __ldg(reinterpret_cast<const uchar4*>(p) + i4);
}
__device__ __forceinline__ float get(const float4& v, int j) {
  return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}
__device__ __forceinline__ void put(float4& v, int j, float s) {
  if (j == 0) v.x = s; else if (j == 1) v.y = s; else if (j == 2) v.z = s; else v.w = s;
}
__device__ __forceinline__ int getx(const uchar4& v, int j) {
  return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}

// ---------------------------------------------------------------------------------------
// Reference-faithful scalar math (float32, IEEE-rounded basic ops).
// ---------------------------------------------------------------------------------------
// jax.nn.sigmoid == 1/(1+exp(-x)).
__device__ __forceinline__ float sigmoid_ref(float x) {
  return __fdiv_rn(1.0f, 1.0f + expf(-x));
}
// EncDec.encode (ldm/model_vdm.py:274-280): 2*((x+.5)/vocab) - 1 ; exact for vocab = 2^k.
__device__ __forceinline__ float encode_ref(int x, float vocab) {
  return 2.0f * __fdiv_rn((float)x + 0.5f, vocab) - 1.0f;
}

// ---------------------------------------------------------------------------------------
// Fast scalar math: bare MUFU + Newton, no slow-path branches.  Each is within ~1 ulp of the
// IEEE result on the stated domain; the residual differs from the reference's rounding
// pseudo-randomly per sub-pixel, so per-example sums are unaffected at the 1e-7 level
// (tests/test_gpu_kernels.py holds them to 1e-5).
// ---------------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2_approx(float x) {
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
// 1/x, x normal and finite: MUFU.RCP + one Newton step.
__device__ __forceinline__ float rcp_nr(float x) {
  const float r = rcp_approx(x);
  return fmaf(fmaf(-x, r, 1.0f), r, r);
}
// 1/S for the polynomial scale: Newton on the normal range, IEEE otherwise (S == 0, denormal,
// huge, NaN are possible for adversarial coefficients and must behave like the reference).
__device__ __forceinline__ bool scale_in_range(float S) {
  const float aS = fabsf(S);
  return aS > 1e-30f && aS < 1e30f;
}
__device__ __forceinline__ float rcp_scale(float S) {
  return scale_in_range(S) ? rcp_nr(S) : __frcp_rn(S);
}
// sqrt(x) = x * rsqrt(x) straight off the MUFU (~2^-22 relative): for alpha/sigma, whose
// rounding in the reference is itself pseudo-random per sub-pixel.  x > 0.
__device__ __forceinline__ float sqrt_fast(float x) { return x * rsqrt_approx(x); }
// same, with sqrt(0) = 0 (1 - sigmoid can round to 0)
__device__ __forceinline__ float sqrt_fast0(float x) {
  const float s = x * rsqrt_approx(x);
  return x > 0.0f ? s : 0.0f;
}
__device__ __forceinline__ float exp_fast(float x) { return ex2_approx(x * kLog2e); }
// sigmoid with the reference's structure 1/(1+exp(-x)): for x >~ 3 (where 1 - sigmoid is
// amplified) the rounding of 1+e dominates and the result equals the IEEE one.
__device__ __forceinline__ float sigmoid_fast(float x) {
  return rcp_nr(1.0f + exp_fast(fminf(-x, 80.0f)));
}
// log(s) for s = 1 + d in [1, 4): series for tiny d (MUFU.LG2 has ~2^-22 ABSOLUTE error,
// useless next to log(1+d) ~ d), MUFU.LG2 otherwise.
__device__ __forceinline__ float log_1p_sum(float s) {
  const float d = s - 1.0f;
  const float small = d * fmaf(d, fmaf(d, kThird, -0.5f), 1.0f);
  const float big = lg2_approx(s) * kLn2;
  return d < 0.0078125f ? small : big;
}

// ---------------------------------------------------------------------------------------
// The epilogue of _compute_coefficients (ldm/model_mulan_epsilon.py:537) for callers that hand
// over the PRE-ACTIVATION of dense_out_c (MULAN_FLAG_C_RAW):
//   c = 1e-3 + softplus(r),  softplus(r) = logaddexp(r, 0) = max(r, 0) + log1p(exp(-|r|)).
// log1p(d), d = exp(-|r|) in (0, 1]: five-term series below 1/16 (truncation < 2e-7 relative),
// MUFU.LG2 above (absolute error <= 2^-22 against a result >= 0.06: < 4e-6 relative); the
// residual is pseudo-random per sub-pixel like every other fast-math rounding here.
// dc/dr = sigmoid(r) for the backward pass.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float log1p_unit(float d) {
  const float ser = d * fmaf(d, fmaf(d, fmaf(d, fmaf(d, kFifth, -0.25f), kThird), -0.5f), 1.0f);
  const float big = lg2_approx(1.0f + d) * kLn2;
  return d < 0.0625f ? ser : big;
}
__device__ __forceinline__ float c_from_raw(float r) {
  const float d = ex2_approx(-fabsf(r) * kLog2e);
  return 1e-3f + (fmaxf(r, 0.0f) + log1p_unit(d));
}
// c and dc/dr = sigmoid(r) from one exponential (backward pass)
__device__ __forceinline__ float c_from_raw_grad(float r, float* dcdr) {
  const float d = ex2_approx(-fabsf(r) * kLog2e);
  const float inv = rcp_nr(1.0f + d);
  *dcdr = r >= 0.0f ? inv : d * inv;
  return 1e-3f + (fmaxf(r, 0.0f) + log1p_unit(d));
}
__device__ __forceinline__ float4 c_from_raw4(float4 r) {
  return make_float4(c_from_raw(r.x), c_from_raw(r.y), c_from_raw(r.z), c_from_raw(r.w));
}

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+).  Every kernel of the path releases its dependents as
// soon as it starts and, before touching global memory, waits for the grid it depends on to
// have completed and flushed.  Both are no-ops for a plain launch; with MULAN_FLAG_PDL the host
// sets cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs are
// already resident (prologue done) when the previous one drains.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_release_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait_for_primary() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Per-example time powers, staged once per row in shared memory.  Integer powers follow
// XLA's integer_pow lowering (binary exponentiation), see oracle.integer_pow.
struct RowT {
  float t, t2, t3, t4, t5;   // t^n
  float t5_5, t3_3, t4_2;    // t^5/5, t^3/3, t^4/2 (fast polynomial)
};
__device__ __forceinline__ RowT make_row_t(float t) {
  RowT r;
  r.t = t;
  r.t2 = t * t;
  r.t3 = t * r.t2;
  r.t4 = r.t2 * r.t2;
  r.t5 = t * r.t4;
  r.t5_5 = r.t5 * kFifth;
  r.t3_3 = r.t3 * kThird;
  r.t4_2 = r.t4 * 0.5f;
  return r;
}

// Discrete time (sm_n_timesteps = T > 0): differences of the time powers between t and
// s = t - 1/T in factored form, so that gamma(t) - gamma(s) = Delta (P(t) - P(s)) / S keeps full
// float32 precision (the reference subtracts two rounded gammas and loses ~T ulps; being
// closer to exact arithmetic than the reference's own float32 is within every tolerance).
struct RowD {
  float d1, d2, d3_3, d4_2, d5_5;   // (t^n - s^n)/{1,1,3,2,5}
};
__device__ __forceinline__ RowD make_row_d(float t, float s) {
  RowD r;
  const float d = t - s, p = t + s, t2 = t * t, s2 = s * s, ts = t * s;
  r.d1 = d;
  r.d2 = d * p;
  r.d3_3 = d * (t2 + ts + s2) * kThird;
  r.d4_2 = d * p * (t2 + s2) * 0.5f;
  r.d5_5 = d * (t2 * t2 + t2 * ts + t2 * s2 + ts * s2 + s2 * s2) * kFifth;
  return r;
}

// NoiseSchedule_polynomial_fixedend._eval_polynomial (ldm/model_mulan_epsilon.py:514-529)
// for one sub-pixel: P(t) = int_0^t (a s^2 + b s + c)^2 ds, S = P(1), q = a t^2 + b t + c.
// FMA form (see file header): differs from the reference's rounding by O(1 ulp) per term,
// pseudo-randomly per sub-pixel.
struct Poly {
  float P, S, q;
  float a2, b2c, ab, bc, c2;  // a^2, b^2+2ac, ab, bc, c^2
};
__device__ __forceinline__ Poly poly_eval(float a, float b, float c, const RowT& r) {
  Poly o;
  o.a2 = a * a;
  o.ab = a * b;
  o.bc = b * c;
  o.c2 = c * c;
  o.b2c = fmaf(a + a, c, b * b);
  o.S = fmaf(o.a2, kFifth, fmaf(o.b2c, kThird, fmaf(o.ab, 0.5f, o.bc + o.c2)));
  o.P = fmaf(o.a2, r.t5_5, fmaf(o.b2c, r.t3_3, fmaf(o.ab, r.t4_2, fmaf(o.bc, r.t2, o.c2 * r.t))));
  o.q = fmaf(fmaf(a, r.t, b), r.t, c);
  return o;
}

// ---------------------------------------------------------------------------------------
// Deterministic per-row reductions: fixed per-thread order, shuffle tree, fixed warp order.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum N per-thread accumulators over a CTA of NW warps; result valid in thread 0.
template <int N, int NW = kWarps>
__device__ __forceinline__ void block_sum(float (&acc)[N], float (*smem)[N]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) smem[warp][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      float s = smem[0][k];
#pragma unroll
      for (int w = 1; w < NW; ++w) s += smem[w][k];
      acc[k] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier: one elected thread moves a whole
// 4 KB slab per operand array; no per-thread address arithmetic, no registers in flight.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace mulan
