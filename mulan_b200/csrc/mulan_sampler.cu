// Ancestral sampler step + final decode (SURVEY.md 8f "next" row 3): the schedule math that
// runs T = 1000 times per generated batch around the denoiser.
//
// Reference statements:
//   VDM.sample / conditional_sample   ldm/model_mulan_epsilon.py:377-438
//                                     ldm/model_mulan_velocity.py:281-347
//       g_t, g_s = gamma(emb, t), gamma(emb, s)            t = (T-i)/T, s = (T-i-1)/T
//       net      = score_model(z_t, mean(g_t) | g_t, cond)           [framework path]
//       a = sigmoid(-g_s); b = sigmoid(-g_t); c = -expm1(g_s - g_t); sigma_t = sqrt(sigmoid(g_t))
//       (velocity: alpha_t = sqrt(sigmoid(-g_t)); eps_hat = v_hat alpha_t + sigma_t z_t)
//       z_s = sqrt(a/b) (z_t - sigma_t c eps_hat) + sqrt((1-a) c) eps
//   VDM.generate_x                    ldm/model_mulan_epsilon.py:440-457 (velocity.py:349-366)
//       g_0 = gamma(emb, 0); z = z_0 / sqrt(1 - sigmoid(g_0)); x = argmax_k decode(z, g_0)
//       (sample_softmax=False in both shipped configs)
//   Experiment_VDM.sample_fn          ldm/experiment_vdm.py:80-110  (the T-step loop; host)
//
// The coefficient arrays may be ONE row broadcast over the batch (abc_rows == 1): the
// unconditional sampler uses the same deterministic embedding for every example
// (_get_deterministic_embedding), so a, b, c stay L2-resident for all 1000 steps.
//
// Numerics: g_s - g_t uses the factored power differences (full float32 precision); the step
// itself has a fast default form and the reference's op-for-op form (see sample_step_kernel).
#include <stdlib.h>

#include "mulan_kernels.h"

namespace mulan {

// ---- noise-level input of the denoiser: per-row mean (or per-pixel) gamma_t ---------------
template <int GT>
__global__ void __launch_bounds__(kThreads)
sample_gamma_kernel(const SamplerParams p) {
  __shared__ RowT s_rt;
  __shared__ float red[kWarps][1];
  const int row = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) s_rt = make_row_t(__ldg(p.t + row));
  __syncthreads();
  const RowT rt = s_rt;
  const size_t cbase = p.abc_rows == 1 ? 0 : (size_t)row * p.dim4;
  const size_t base4 = (size_t)row * p.dim4;
  float acc[1] = {0.f};
  for (int i4 = tid; i4 < p.dim4; i4 += kThreads) {
    const float4 A = ld4(p.a, cbase + i4), Bv = ld4(p.b, cbase + i4), C = ld4(p.c, cbase + i4);
    float4 G;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const Poly po = poly_eval(get(A, j), get(Bv, j), get(C, j), rt);
      const float gt = p.gmin + (p.delta * po.P) * rcp_scale(po.S);
      put(G, j, gt);
      acc[0] += gt;
    }
    if (GT == MULAN_GT_PIXEL) st4(p.g_net, base4 + i4, G);
  }
  if (GT == MULAN_GT_MEAN) {
    block_sum<1>(acc, red);
    if (tid == 0) p.g_net[row] = __fdiv_rn(acc[0], (float)(p.dim4 * 4));
  }
}

// ---- one ancestral step -------------------------------------------------------------------
// 1 - e^{-x}, x >= 0 (= -expm1(-x)): alternating series below 1/2 (truncation x^8/8! < 1e-7
// relative), MUFU.EX2 above (absolute error 2e-7 e^{-x} on a value >= 0.39).
__device__ __forceinline__ float one_minus_exp_neg(float x) {
  const float ser = x * fmaf(x, fmaf(x, fmaf(x, fmaf(x, fmaf(x, fmaf(x, 1.0f / 5040.0f,
                    -1.0f / 720.0f), 1.0f / 120.0f), -1.0f / 24.0f), 1.0f / 6.0f), -0.5f), 1.0f);
  const float big = 1.0f - ex2_approx(-x * kLog2e);
  return x < 0.5f ? ser : big;
}

// The step is linear in (z_t, net, eps) with per-sub-pixel factors that depend on the row only
// through (t, s):   z_s = m1 z_t + m2 net + m3 eps
//   eps model:  m1 = sqrt(a/b)                      m2 = -sqrt(a/b) sigma_t c
//   v model:    m1 = sqrt(a/b) (1 - sigma_t^2 c)    m2 = -sqrt(a/b) sigma_t c alpha_t
//               (eps_hat = v_hat alpha_t + sigma_t z_t substituted)
//   both:       m3 = sqrt((1-a) c)
// Fast form of the factors, from e^{gamma/2} and MUFU reciprocal square roots:
//   a = 1/(1+e^{g_s}),  b = 1/(1+e^{g_t}),  sqrt(a/b) = sqrt(1+e^{g_t}) rsqrt(1+e^{g_s}),
//   1 - a = e^{g_s}/(1+e^{g_s})  (NOT formed as 1 - fl(a): near gamma_min that difference is
//   1.7e-6 quantised in units of 6e-8 in the reference's float32, a 3.5 % error in its own noise
//   scale -- this form is the value exact arithmetic gives, and tests/test_sampler.py bounds the
//   kernel by the reference's own float32-to-float64 distance there),
//   sigma_t = e^{g_t/2} rsqrt(1+e^{g_t}),  alpha_t = rsqrt(1+e^{g_t}),  c = 1 - e^{-(g_t-g_s)}
//   with g_t - g_s from the factored power differences (full float32 precision; the reference's
//   own float32 subtracts two rounded gammas and gets c <= 0 -> NaN where gamma is locally flat).
// Each factor is within ~3e-7 relative of exact.
struct StepF { float m1, m2, m3; };
template <int PARAM>
__device__ __forceinline__ StepF step_factors(float a, float b, float c, const RowT& rt,
                                              const RowT& rs, const RowD& rd, float gmin,
                                              float delta) {
  const Poly po = poly_eval(a, b, c, rt);
  const float rS = rcp_scale(po.S);
  const float Ps = fmaf(po.a2, rs.t5_5, fmaf(po.b2c, rs.t3_3, fmaf(po.ab, rs.t4_2,
                   fmaf(po.bc, rs.t2, po.c2 * rs.t))));
  const float dP = fmaf(po.a2, rd.d5_5, fmaf(po.b2c, rd.d3_3, fmaf(po.ab, rd.d4_2,
                   fmaf(po.bc, rd.d2, po.c2 * rd.d1))));
  const float dr = delta * rS;                       // Delta / S
  const float gt = fmaf(po.P, dr, gmin), gs = fmaf(Ps, dr, gmin);
  // e^{gamma/2}; the clamp keeps 1 + e^gamma finite (gamma > 80 is sigma = 1 anyway)
  const float ht = ex2_approx(fminf(gt, 80.0f) * (0.5f * kLog2e));
  const float hs = ex2_approx(fminf(gs, 80.0f) * (0.5f * kLog2e));
  const float pt = fmaf(ht, ht, 1.0f), ps = fmaf(hs, hs, 1.0f);   // 1/b, 1/a
  const float rpt = rsqrt_approx(pt), rps = rsqrt_approx(ps);    // alpha_t, sqrt(a)
  const float sig = ht * rpt;                                     // sigma_t
  const float cv = fmaxf(one_minus_exp_neg(fmaxf(dr * dP, 0.0f)), 0.0f);   // c
  const float fm = (pt * rpt) * rps;                              // sqrt(a/b)
  const float sc = sig * cv;                                      // sigma_t c
  StepF f;
  if (PARAM == MULAN_PARAM_EPS) {
    f.m1 = fm;
    f.m2 = -(fm * sc);
  } else {
    f.m1 = fm * fmaf(-sc, sig, 1.0f);
    f.m2 = -(fm * sc) * rpt;
  }
  f.m3 = (hs * rps) * sqrt_fast0(cv);                             // sqrt((1-a) c)
  return f;
}
__device__ __forceinline__ float step_apply(const StepF& f, float z, float n, float e) {
  return fmaf(f.m3, e, fmaf(f.m2, n, f.m1 * z));
}

// IEEE = true: the reference's statements op for op (IEEE division, sqrtf, expf, expm1f) -- the
// first version of this kernel, kept for A/B (MULAN_SAMPLER_IEEE=1); it is instruction-bound at
// 39 % of the HBM roofline.  IEEE = false (default): step_factors / step_apply.
template <int PARAM, bool IEEE>
__global__ void __launch_bounds__(kThreads)
sample_step_kernel(const SamplerParams p) {
  __shared__ RowT s_rt, s_rs;
  __shared__ RowD s_rd;
  const int row = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    const float t = __ldg(p.t + row), s = __ldg(p.s + row);
    s_rt = make_row_t(t);
    s_rs = make_row_t(s);
    s_rd = make_row_d(t, s);
  }
  __syncthreads();
  const RowT rt = s_rt, rs = s_rs;
  const RowD rd = s_rd;
  const size_t cbase = p.abc_rows == 1 ? 0 : (size_t)row * p.dim4;
  const size_t base4 = (size_t)row * p.dim4;
  for (int i4 = tid; i4 < p.dim4; i4 += kThreads) {
    const float4 A = ld4(p.a, cbase + i4), Bv = ld4(p.b, cbase + i4), C = ld4(p.c, cbase + i4);
    const float4 Z = ld4(p.z_t, base4 + i4), N = ld4(p.net, base4 + i4), E = ld4(p.eps, base4 + i4);
    float4 O;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z = get(Z, j);
      float eh = get(N, j);
      if (IEEE) {
        const Poly po = poly_eval(get(A, j), get(Bv, j), get(C, j), rt);
        const float rS = rcp_scale(po.S);
        const float Ps = fmaf(po.a2, rs.t5_5, fmaf(po.b2c, rs.t3_3, fmaf(po.ab, rs.t4_2,
                         fmaf(po.bc, rs.t2, po.c2 * rs.t))));
        const float dP = fmaf(po.a2, rd.d5_5, fmaf(po.b2c, rd.d3_3, fmaf(po.ab, rd.d4_2,
                         fmaf(po.bc, rd.d2, po.c2 * rd.d1))));
        const float gt = p.gmin + (p.delta * po.P) * rS;
        const float gs = p.gmin + (p.delta * Ps) * rS;
        const float av = sigmoid_ref(-gs);                 // a
        const float bv = sigmoid_ref(-gt);                 // b
        const float cv = fmaxf(-expm1f(-(p.delta * dP) * rS), 0.0f);   // c, clamped (see above)
        const float sig = sqrtf(sigmoid_ref(gt));          // sigma_t
        if (PARAM != MULAN_PARAM_EPS) eh = eh * sqrtf(bv) + sig * z;   // v -> eps
        const float mean = sqrtf(__fdiv_rn(av, bv)) * (z - sig * cv * eh);
        put(O, j, mean + sqrtf((1.0f - av) * cv) * get(E, j));
      } else {
        const StepF f = step_factors<PARAM>(get(A, j), get(Bv, j), get(C, j), rt, rs, rd, p.gmin,
                                            p.delta);
        put(O, j, step_apply(f, z, eh, get(E, j)));
      }
    }
    st4(p.z_s, base4 + i4, O);
  }
}

// One coefficient row broadcast over the batch (abc_rows == 1: the unconditional sampler, where
// every example also shares t and s): the factors are identical for every row with the same
// (t, s), so a persistent CTA keeps them in shared memory -- 3 x dim floats, each thread reading
// back only the slots it wrote itself, hence no barrier around the table -- filled while the
// first row of a run of equal (t, s) is processed.  Every further row of the run is three FMAs
// per sub-pixel on 16 B of traffic: HBM-bound instead of issue-bound.  Same step_factors /
// step_apply as the direct kernel: bit-identical results.
template <int PARAM>
__global__ void __launch_bounds__(kThreads, 5)
sample_step_bcast_kernel(const SamplerParams p) {
  extern __shared__ float4 s_tab[];          // [3][dim4]
  __shared__ RowT s_rt, s_rs;
  __shared__ RowD s_rd;
  const int tid = threadIdx.x;
  float4* const T1 = s_tab, * const T2 = s_tab + p.dim4, * const T3 = s_tab + 2 * p.dim4;
  float ct = __int_as_float(0x7fc00000), cs = ct;    // cached (t, s): NaN = nothing cached
  for (int row = blockIdx.x; row < p.rows; row += gridDim.x) {
    const float t = __ldg(p.t + row), s = __ldg(p.s + row);
    const size_t base4 = (size_t)row * p.dim4;
    if (t == ct && s == cs) {                          // uniform over the CTA: table hit
      for (int i4 = tid; i4 < p.dim4; i4 += kThreads) {
        const float4 Z = ld4(p.z_t, base4 + i4), N = ld4(p.net, base4 + i4),
                     E = ld4(p.eps, base4 + i4);
        const float4 M1 = T1[i4], M2 = T2[i4], M3 = T3[i4];
        float4 O;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const StepF f = {get(M1, j), get(M2, j), get(M3, j)};
          put(O, j, step_apply(f, get(Z, j), get(N, j), get(E, j)));
        }
        st4(p.z_s, base4 + i4, O);
      }
      continue;
    }
    // Miss: compute this row directly; keep its factors only if the next row this CTA will
    // process shares (t, s) -- rows with individual times never pay for a table they cannot reuse.
    const int nxt = row + (int)gridDim.x;
    const bool keep = nxt < p.rows && __ldg(p.t + nxt) == t && __ldg(p.s + nxt) == s;
    __syncthreads();                                   // the previous row structs are no longer read
    if (tid == 0) {
      s_rt = make_row_t(t);
      s_rs = make_row_t(s);
      s_rd = make_row_d(t, s);
    }
    __syncthreads();
    const RowT rt = s_rt, rs = s_rs;
    const RowD rd = s_rd;
    for (int i4 = tid; i4 < p.dim4; i4 += kThreads) {
      const float4 A = ld4(p.a, i4), Bv = ld4(p.b, i4), C = ld4(p.c, i4);
      const float4 Z = ld4(p.z_t, base4 + i4), N = ld4(p.net, base4 + i4),
                   E = ld4(p.eps, base4 + i4);
      float4 M1, M2, M3, O;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const StepF f = step_factors<PARAM>(get(A, j), get(Bv, j), get(C, j), rt, rs, rd, p.gmin,
                                            p.delta);
        put(M1, j, f.m1); put(M2, j, f.m2); put(M3, j, f.m3);
        put(O, j, step_apply(f, get(Z, j), get(N, j), get(E, j)));
      }
      if (keep) { T1[i4] = M1; T2[i4] = M2; T3[i4] = M3; }
      st4(p.z_s, base4 + i4, O);
    }
    if (keep) { ct = t; cs = s; }
  }
}

// ---- final decode: x = argmax_k log p(k | z_0 / sqrt(1 - var_0)) --------------------------
__global__ void __launch_bounds__(kThreads)
generate_x_kernel(const SamplerParams p) {
  const int row = blockIdx.x, tid = threadIdx.x;
  const size_t base4 = (size_t)row * p.dim4;
  const VocabInfo vi = p.vi;
  for (int i4 = tid; i4 < p.dim4; i4 += kThreads) {
    const float4 Z = ld4(p.z_t, base4 + i4);
    uchar4 X;
    unsigned char* xo = reinterpret_cast<unsigned char*>(&X);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z = __fdiv_rn(get(Z, j), p.den0);
      // the logits -0.5((z - x_k) e^{-g0/2})^2 peak at the nearest bin centre; jnp.argmax takes
      // the FIRST maximum, so resolve the two candidates around z with the reference's logits
      float kf = floorf((z + 1.0f) * vi.half_vocab - 0.5f);
      kf = fminf(fmaxf(kf, 0.0f), vi.vocab_m1);
      const int k0 = (int)kf, k1 = min(k0 + 1, vi.vocab - 1);
      const float u0 = (z - vi.xval(k0)) * p.inv0, u1 = (z - vi.xval(k1)) * p.inv0;
      const float l0 = -0.5f * (u0 * u0), l1 = -0.5f * (u1 * u1);
      xo[j] = (unsigned char)(l1 > l0 ? k1 : k0);
    }
    reinterpret_cast<uchar4*>(p.x)[base4 + i4] = X;
  }
}

// ---- probability-flow ODE drift + the local part of its Hutchinson divergence -------------
// VDM.reverse_ode (ldm/model_mulan_epsilon.py:459-478):
//   sigma = sqrt(sigmoid(g_t))   (high_precision: exp(g_t/2) where sigmoid(g_t) <= 1e-3)
//   drift = 0.5 (-sigma x + eps_hat) sigma g_t_grad
// _get_value_div_fn (ldm/notebook_utils.py:204-216): div = sum_d v (d<drift, v>/dx).  With
// k = 0.5 sigma g_t_grad:  d<drift,v>/dx = -sigma k v  +  J_net^T (k v), so this pass also emits
// net_bar = k v (the cotangent the denoiser's backward needs) and the per-row direct part
// div_direct = sum_d -sigma k v^2.  mulan_row_dot adds <J_net^T net_bar, v> afterwards.
template <bool HP, bool HUTCH>
__global__ void __launch_bounds__(kThreads)
ode_drift_kernel(const SamplerParams p) {
  __shared__ RowT s_rt;
  __shared__ float red[kWarps][1];
  const int row = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) s_rt = make_row_t(__ldg(p.t + row));
  __syncthreads();
  const RowT rt = s_rt;
  const size_t cbase = p.abc_rows == 1 ? 0 : (size_t)row * p.dim4;
  const size_t base4 = (size_t)row * p.dim4;
  float acc[1] = {0.f};
  for (int i4 = tid; i4 < p.dim4; i4 += kThreads) {
    const float4 A = ld4(p.a, cbase + i4), Bv = ld4(p.b, cbase + i4), C = ld4(p.c, cbase + i4);
    const float4 X = ld4(p.z_t, base4 + i4), N = ld4(p.net, base4 + i4);
    float4 V = make_float4(0.f, 0.f, 0.f, 0.f), DR, NB;
    if (HUTCH) V = ld4(p.eps, base4 + i4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const Poly po = poly_eval(get(A, j), get(Bv, j), get(C, j), rt);
      const float rS = rcp_scale(po.S);
      const float gt = p.gmin + (p.delta * po.P) * rS;
      const float w = (p.delta * (po.q * po.q)) * rS;
      const float var = sigmoid_ref(gt);
      float sigma = sqrtf(var);
      if (HP && var <= 1e-3f) sigma = expf(gt / 2.0f);
      const float x = get(X, j), n = get(N, j);
      put(DR, j, 0.5f * (-sigma * x + n) * sigma * w);
      if (HUTCH) {
        const float v = get(V, j);
        const float k = 0.5f * sigma * w;
        put(NB, j, k * v);
        acc[0] += -(sigma * k) * (v * v);
      }
    }
    st4(p.z_s, base4 + i4, DR);
    if (HUTCH) st4(p.g_net, base4 + i4, NB);
  }
  if (HUTCH) {
    block_sum<1>(acc, red);
    if (tid == 0) p.div_direct[row] = acc[0];
  }
}

// out[b] = sum_d u[b,d] v[b,d] (+ add[b])
__global__ void __launch_bounds__(kThreads)
row_dot_kernel(const float* __restrict__ u, const float* __restrict__ v,
               const float* __restrict__ add, float* __restrict__ out, int dim4) {
  __shared__ float red[kWarps][1];
  const int row = blockIdx.x;
  const size_t base4 = (size_t)row * dim4;
  float acc[1] = {0.f};
  for (int i4 = threadIdx.x; i4 < dim4; i4 += kThreads) {
    const float4 U = ld4(u, base4 + i4), V = ld4(v, base4 + i4);
    acc[0] += U.x * V.x; acc[0] += U.y * V.y; acc[0] += U.z * V.z; acc[0] += U.w * V.w;
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) out[row] = add != nullptr ? acc[0] + add[row] : acc[0];
}

cudaError_t launch_ode_drift(const SamplerParams& p, bool high_precision, cudaStream_t s) {
  if (p.rows == 0) return cudaSuccess;
  const bool hutch = p.eps != nullptr;
  if (high_precision) {
    if (hutch) ode_drift_kernel<true, true><<<p.rows, kThreads, 0, s>>>(p);
    else       ode_drift_kernel<true, false><<<p.rows, kThreads, 0, s>>>(p);
  } else {
    if (hutch) ode_drift_kernel<false, true><<<p.rows, kThreads, 0, s>>>(p);
    else       ode_drift_kernel<false, false><<<p.rows, kThreads, 0, s>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_row_dot(const float* u, const float* v, const float* add, float* out, int rows,
                           int dim4, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  row_dot_kernel<<<rows, kThreads, 0, s>>>(u, v, add, out, dim4);
  return cudaGetLastError();
}

cudaError_t launch_sample_gamma(const SamplerParams& p, cudaStream_t s) {
  if (p.rows == 0) return cudaSuccess;
  if (p.gt_mode == MULAN_GT_MEAN) sample_gamma_kernel<MULAN_GT_MEAN><<<p.rows, kThreads, 0, s>>>(p);
  else                            sample_gamma_kernel<MULAN_GT_PIXEL><<<p.rows, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_sample_step(const SamplerParams& p, cudaStream_t s) {
  if (p.rows == 0) return cudaSuccess;
  static const int ieee = [] {
    const char* e = getenv("MULAN_SAMPLER_IEEE");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  const bool eps = p.param == MULAN_PARAM_EPS;
  if (ieee) {
    if (eps) sample_step_kernel<MULAN_PARAM_EPS, true><<<p.rows, kThreads, 0, s>>>(p);
    else     sample_step_kernel<MULAN_PARAM_VEL, true><<<p.rows, kThreads, 0, s>>>(p);
    return cudaGetLastError();
  }
  const size_t tab_bytes = (size_t)3 * p.dim4 * sizeof(float4);
  if (p.abc_rows == 1 && tab_bytes <= 48 * 1024) {      // factor table in shared memory
    const void* k = eps ? (const void*)sample_step_bcast_kernel<MULAN_PARAM_EPS>
                        : (const void*)sample_step_bcast_kernel<MULAN_PARAM_VEL>;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, tab_bytes) !=
            cudaSuccess || per_sm < 1)
      per_sm = 1;
    const int grid = p.rows < sms * per_sm ? p.rows : sms * per_sm;
    if (eps) sample_step_bcast_kernel<MULAN_PARAM_EPS><<<grid, kThreads, tab_bytes, s>>>(p);
    else     sample_step_bcast_kernel<MULAN_PARAM_VEL><<<grid, kThreads, tab_bytes, s>>>(p);
    return cudaGetLastError();
  }
  if (eps) sample_step_kernel<MULAN_PARAM_EPS, false><<<p.rows, kThreads, 0, s>>>(p);
  else     sample_step_kernel<MULAN_PARAM_VEL, false><<<p.rows, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_generate_x(const SamplerParams& p, cudaStream_t s) {
  if (p.rows == 0) return cudaSuccess;
  generate_x_kernel<<<p.rows, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace mulan
