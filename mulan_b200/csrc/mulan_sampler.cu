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
// Numerics: sigmoid(-g) and 1 - a are formed exactly as the reference does (IEEE division;
// at g_s -> gamma_min, 1 - a is 1.7e-6 quantised in units of 6e-8, which the reference's noise
// scale inherits).  g_s - g_t uses the factored power differences (full float32 precision).
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
template <int PARAM>
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
      // c = -expm1(g_s - g_t) >= 0 since gamma is monotone.  The reference's float32 evaluation
      // subtracts two rounded gammas and gets c <= 0 -> NaN in pixels where gamma is locally
      // flat (tests/test_sampler.py); here the difference keeps full precision and is clamped.
      const float cv = fmaxf(-expm1f(-(p.delta * dP) * rS), 0.0f);
      const float sig = sqrtf(sigmoid_ref(gt));          // sigma_t
      const float z = get(Z, j);
      float eh = get(N, j);
      if (PARAM != MULAN_PARAM_EPS) eh = eh * sqrtf(bv) + sig * z;   // v -> eps
      const float mean = sqrtf(__fdiv_rn(av, bv)) * (z - sig * cv * eh);
      put(O, j, mean + sqrtf((1.0f - av) * cv) * get(E, j));
    }
    st4(p.z_s, base4 + i4, O);
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
  if (p.param == MULAN_PARAM_EPS) sample_step_kernel<MULAN_PARAM_EPS><<<p.rows, kThreads, 0, s>>>(p);
  else                            sample_step_kernel<MULAN_PARAM_VEL><<<p.rows, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_generate_x(const SamplerParams& p, cudaStream_t s) {
  if (p.rows == 0) return cudaSuccess;
  generate_x_kernel<<<p.rows, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace mulan
