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
