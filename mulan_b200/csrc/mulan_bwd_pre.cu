// mulan_bwd_pre: cotangents of the polynomial coefficients (a, b, c), one pass over [B, D]
// (37 B/sub-pixel algorithmic for EPS: z_bar4 + x1 + a,b,c 12 + eps4 + net4 -> 12).
//
// This is the closed form of what jax.value_and_grad (ldm/experiment.py:339) derives for
// the statements of VDM.__call__ (ldm/model_mulan_epsilon.py:307-347,
// ldm/model_mulan_velocity.py:215-260) with respect to the outputs of
// NoiseSchedule_polynomial_fixedend._compute_coefficients (:531-538):
//
//   gamma = gmin + D P/S,  w = d gamma/dt = D q^2/S,  q = a t^2 + b t + c
//   gamma_bar = z_bar (d alpha/d gamma f + d sigma/d gamma eps) + g_bar/Dim | g_bar_pix + (VEL*: direct)
//   w_bar     = .5 gL r^2 (EPS) | .5 gL (1-v) r^2 (VEL*)
//   a_bar = D [gamma_bar (P_a S - P S_a) + w_bar (Q_a S - Q S_a)] / S^2, same for b, c.
// gamma(0) and gamma(1) are fixed ends: loss_recon and the prior KL contribute nothing.
#include "mulan_kernels.h"

namespace mulan {

// Arithmetic notes.  The backward pass has no reference rounding order to reproduce (what
// jax.value_and_grad emits is XLA's business; the tolerance is 1e-4 on gradients), so products
// and sums are contracted into fmaf() freely here, and the factors 1/2 of d alpha/d gamma =
// -v alpha/2, d sigma/d gamma = sigma (1-v)/2 and of w_bar are folded into per-row constants
// (gLh = gL/2) -- the kernel is instruction-issue bound in the velocity modes.  gamma_t and
// d gamma/dt are formed exactly as mulan_fwd_pre forms them (fma(P, Delta/S, gmin), Q Delta/S),
// so alpha and sigma here are the ones z_t was built with.
// CRAW (MULAN_FLAG_C_RAW): p.c is the pre-activation r of dense_out_c; c = 1e-3 + softplus(r) is
// formed here and c_bar is returned as the cotangent of r (c_bar * sigmoid(r)), so the
// framework's softplus backward (12 B/sub-pixel) disappears (ldm/model_mulan_epsilon.py:537).
// NT / MINB: 256 threads, <= 51 registers, 5 CTAs (40 warps) per SM (throughput), or 768 threads
// = one float4 column per thread for launches of at most one row per SM (latency; there every
// thread forms the row constants itself: no staging barrier before the operand loads).
template <int PARAM, int GT, bool DISC, bool POW2, bool CRAW, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
bwd_pre_kernel(const BwdPreParams p) {
  const int row = blockIdx.x, tid = threadIdx.x;
  const bool has_gL = p.gL != nullptr;
  const bool has_zb = p.z_bar != nullptr;
  const bool has_gb = p.g_bar != nullptr;
  pdl_release_dependents();
  pdl_wait_for_primary();
  // row constants: per thread in the latency shape, staged by thread 0 in the throughput shape
  // (registers are the scarce resource there: 48 at 5 CTAs per SM)
  RowT rt;
  RowD rd;
  float gLh, gbar_row;
  if constexpr (NT >= kLatencyThreads) {
    const float t_row = __ldg(p.t + row);
    rt = make_row_t(t_row);
    if (DISC) rd = make_row_d(t_row, t_row - p.inv_T);               // s = t - 1/T
    gLh = has_gL ? 0.5f * __ldg(p.gL + row) : 0.f;
    // jnp.mean backward: cotangent / D broadcast to every sub-pixel
    gbar_row = (GT == MULAN_GT_MEAN && has_gb)
                   ? __fdiv_rn(__ldg(p.g_bar + row), (float)(p.dim4 * 4)) : 0.f;
  } else {
    __shared__ RowT s_rt;
    __shared__ RowD s_rd;
    __shared__ float s_gLh, s_gbar;
    if (tid == 0) {
      const float t_row = __ldg(p.t + row);
      s_rt = make_row_t(t_row);
      if (DISC) s_rd = make_row_d(t_row, t_row - p.inv_T);
      s_gLh = has_gL ? 0.5f * __ldg(p.gL + row) : 0.f;
      s_gbar = (GT == MULAN_GT_MEAN && has_gb)
                   ? __fdiv_rn(__ldg(p.g_bar + row), (float)(p.dim4 * 4)) : 0.f;
    }
    __syncthreads();
    rt = s_rt;
    if (DISC) rd = s_rd;
    gLh = s_gLh; gbar_row = s_gbar;
  }
  const VocabInfo vi = p.vi;
  const float two_iv = vi.inv_vocab + vi.inv_vocab, off = vi.inv_vocab - 1.0f;
  // t-dependent coefficients of P_a, P_b, P_c with the factors 2 folded in (exact scalings)
  const float t = rt.t, t2 = rt.t2, t3_3 = rt.t3_3, t4_2 = rt.t4_2, t5_5 = rt.t5_5;
  const float t5_5x2 = t5_5 + t5_5, t3_3x2 = t3_3 + t3_3, tx2 = t + t;
  constexpr float kTwoFifths = 2.0f * kFifth, kTwoThirds = 2.0f * kThird;
  const size_t base4 = (size_t)row * p.dim4;
  const size_t nbase4 = (size_t)(p.noise_rows > 0 ? row % p.noise_rows : row) * p.dim4;
  const bool need_x = has_zb || (has_gL && PARAM != MULAN_PARAM_EPS);

  for (int i4 = tid; i4 < p.dim4; i4 += NT) {
    const size_t g4 = base4 + i4;
    const float4 A = ld4(p.a, g4), Bv = ld4(p.b, g4), C = ld4(p.c, g4);
    float4 E = make_float4(0.f, 0.f, 0.f, 0.f), N = E, ZB = E, GB = E;
    uchar4 X = make_uchar4(0, 0, 0, 0);
    if (has_gL || has_zb) E = ld4(p.eps, nbase4 + i4);
    if (has_gL) N = ld4(p.net, g4);
    if (has_zb) ZB = ld4(p.z_bar, g4);
    if (GT == MULAN_GT_PIXEL && has_gb) GB = ld4(p.g_bar, g4);
    if (need_x) X = ldx4(p.x, g4);
    float4 AB, BB, CB;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = get(A, j), b = get(Bv, j);
      float dcdr = 1.0f;
      const float c = CRAW ? c_from_raw_grad(get(C, j), &dcdr) : get(C, j);
      const float e = get(E, j), n = get(N, j);
      // encode(x): exact in one fma for a power-of-two vocab (as mulan_fwd_pre)
      const float f = POW2 ? fmaf((float)getx(X, j), two_iv, off) : vi.xval(getx(X, j));
      const Poly po = poly_eval(a, b, c, rt);
      const float rS = rcp_scale(po.S);
      const float dr = p.delta * rS;                   // Delta / S
      const float Q = po.q * po.q;
      const float u = po.P * rS, y = Q * rS;
      const float gt = fmaf(po.P, dr, p.gmin);         // gamma_t
      const float w = Q * dr;                          // d gamma / dt
      const float v = sigmoid_fast(gt);
      const float om = 1.0f - v;
      const float kr = rsqrt_approx(fmaxf(om, 1e-30f));   // 1/alpha = sqrt(1 + e^gamma)
      const float alpha = om * kr, sigma = sqrt_fast(v);
      const float ha = v * alpha;       // -2 d alpha / d gamma
      const float hs = sigma * om;      //  2 d sigma / d gamma

      float gbar = (GT == MULAN_GT_MEAN) ? gbar_row : get(GB, j);
      float zb = get(ZB, j);
      float wbar = 0.f;
      float gD = 0.f, du = 0.f;         // DISC: cotangent of gamma(t) - gamma(s), (P(t)-P(s))/S
      if (PARAM == MULAN_PARAM_EPS && DISC) {
        // loss = .5 T sum expm1(dg) r^2, dg = gamma(t) - gamma(s) = Delta (P(t) - P(s)) / S
        const float r = e - n;
        const float dP = fmaf(po.a2, rd.d5_5, fmaf(po.b2c, rd.d3_3, fmaf(po.ab, rd.d4_2,
                         fmaf(po.bc, rd.d2, po.c2 * rd.d1))));
        du = dP * rS;
        const float dexp = expf(p.delta * du);
        gD = dr * ((float)p.T * gLh * (r * r) * dexp);
      } else if (PARAM == MULAN_PARAM_EPS) {
        const float r = e - n;
        wbar = gLh * (r * r);
      } else {
        const float vtg = fmaf(alpha, e, -(sigma * f));
        float vhat = n, zt = 0.f;
        if (PARAM == MULAN_PARAM_VEL_FROM_EPS) {
          // e^g = v/(1-v): sqrt(1+e^g) = 1/alpha = kr, e^{g/2} = sigma kr, e^g/sqrt(1+e^g) = v kr
          zt = fmaf(alpha, f, sigma * e);
          vhat = kr * fmaf(-sigma, zt, n);
        }
        const float r = vtg - vhat;
        wbar = (gLh * (r * r)) * om;                     // .5 gL (1-v) r^2
        const float rbh = (gLh * r) * (om * w);          // .5 dL/dr
        gbar = fmaf(-(wbar * w), v, gbar);               // through (1 - var_t)
        gbar = fmaf(-rbh, fmaf(ha, e, hs * f), gbar);    // through v_target
        if (PARAM == MULAN_PARAM_VEL_FROM_EPS) {
          zb = fmaf(rbh + rbh, sigma * kr, zb);          // v_hat's direct use of z_t
          gbar = fmaf(rbh * kr, fmaf(sigma, zt, -(n * v)), gbar);  // v_hat's own gamma dependence
        }
      }
      gbar = fmaf(0.5f * zb, fmaf(hs, e, -(ha * f)), gbar);  // through z_t = alpha f + sigma eps

      const float gG = dr * gbar, gW = dr * wbar;
      const float two_q = po.q + po.q;
      // partials of P, S, Q
      const float Pa = fmaf(a, t5_5x2, fmaf(c, t3_3x2, b * t4_2));
      const float Sa = fmaf(a, kTwoFifths, fmaf(c, kTwoThirds, 0.5f * b));
      const float Pb = fmaf(b, t3_3x2, fmaf(a, t4_2, c * t2));
      const float Sb = fmaf(b, kTwoThirds, fmaf(a, 0.5f, c));
      const float Pc = fmaf(a, t3_3x2, fmaf(b, t2, c * tx2));
      const float Sc = fmaf(a, kTwoThirds, fmaf(c, 2.0f, b));
      const float Qa = two_q * t2, Qb = two_q * t, Qc = two_q;
      float ab_ = fmaf(gG, fmaf(-u, Sa, Pa), gW * fmaf(-y, Sa, Qa));
      float bb_ = fmaf(gG, fmaf(-u, Sb, Pb), gW * fmaf(-y, Sb, Qb));
      float cb_ = fmaf(gG, fmaf(-u, Sc, Pc), gW * fmaf(-y, Sc, Qc));
      if (DISC) {                               // path through gamma(t) - gamma(s)
        const float dPa = fmaf(a + a, rd.d5_5, fmaf(c + c, rd.d3_3, b * rd.d4_2));
        const float dPb = fmaf(b + b, rd.d3_3, fmaf(a, rd.d4_2, c * rd.d2));
        const float dPc = fmaf(a + a, rd.d3_3, fmaf(b, rd.d2, (c + c) * rd.d1));
        ab_ = fmaf(gD, fmaf(-du, Sa, dPa), ab_);
        bb_ = fmaf(gD, fmaf(-du, Sb, dPb), bb_);
        cb_ = fmaf(gD, fmaf(-du, Sc, dPc), cb_);
      }
      put(AB, j, ab_);
      put(BB, j, bb_);
      put(CB, j, CRAW ? cb_ * dcdr : cb_);
    }
    st4(p.a_bar, g4, AB);
    st4(p.b_bar, g4, BB);
    st4(p.c_bar, g4, CB);
  }
}

template <int PARAM, bool POW2, bool CRAW, int NT, int MINB>
static cudaError_t launch_gt(const BwdPreParams& p, cudaStream_t s) {
  const bool pdl = p.pdl != 0;
  if (PARAM == MULAN_PARAM_EPS && p.T > 0) {
    if (p.gt_mode == MULAN_GT_MEAN)
      return launch_kernel(
          bwd_pre_kernel<MULAN_PARAM_EPS, MULAN_GT_MEAN, true, POW2, CRAW, NT, MINB>, p.rows, NT,
          s, pdl, p);
    return launch_kernel(
        bwd_pre_kernel<MULAN_PARAM_EPS, MULAN_GT_PIXEL, true, POW2, CRAW, NT, MINB>, p.rows, NT, s,
        pdl, p);
  }
  if (p.gt_mode == MULAN_GT_MEAN)
    return launch_kernel(bwd_pre_kernel<PARAM, MULAN_GT_MEAN, false, POW2, CRAW, NT, MINB>, p.rows,
                         NT, s, pdl, p);
  return launch_kernel(bwd_pre_kernel<PARAM, MULAN_GT_PIXEL, false, POW2, CRAW, NT, MINB>, p.rows,
                       NT, s, pdl, p);
}

template <int PARAM, bool POW2, bool CRAW>
static cudaError_t launch_nt(const BwdPreParams& p, cudaStream_t s) {
  // at most one CTA per SM: one float4 column per thread (see latency_rows()); the shipped
  // configs' power-of-two vocabulary only (the generic vocab keeps the throughput shape)
  if (POW2 && shape_rows(p.rows) <= latency_rows() && p.dim4 <= kLatencyThreads)
    return launch_gt<PARAM, POW2, CRAW, POW2 ? kLatencyThreads : kThreads, POW2 ? 1 : 5>(p, s);
  return launch_gt<PARAM, POW2, CRAW, kThreads, 5>(p, s);
}

template <int PARAM>
static cudaError_t launch_param(const BwdPreParams& p, cudaStream_t s) {
  const bool pow2 = p.vi.pow2 != 0;
  // the pre-activation form is built for power-of-two vocabularies (both shipped configs)
  if (p.c_raw) return pow2 ? launch_nt<PARAM, true, true>(p, s) : cudaErrorNotSupported;
  return pow2 ? launch_nt<PARAM, true, false>(p, s) : launch_nt<PARAM, false, false>(p, s);
}

// w = expm1(gamma(t) - gamma(t - 1/T)): the discrete-time weight of the epsilon loss
// (ldm/model_mulan_epsilon.py:350-354), written over the d-gamma/dt that fwd_pre saved.
__global__ void __launch_bounds__(kThreads)
discrete_w_kernel(const DiscreteWParams p) {
  __shared__ RowT s_rt;
  __shared__ RowD s_rd;
  const int row = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    const float t = __ldg(p.t + row);
    s_rt = make_row_t(t);
    s_rd = make_row_d(t, t - p.inv_T);
  }
  __syncthreads();
  const RowT rt = s_rt;
  const RowD rd = s_rd;
  const size_t base4 = (size_t)row * p.dim4;
  for (int i4 = tid; i4 < p.dim4; i4 += kThreads) {
    const size_t g4 = base4 + i4;
    const float4 A = ld4(p.a, g4), Bv = ld4(p.b, g4), C = ld4(p.c, g4);
    float4 Wv;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float c = p.c_raw ? c_from_raw(get(C, j)) : get(C, j);
      const Poly po = poly_eval(get(A, j), get(Bv, j), c, rt);
      const float rS = rcp_scale(po.S);
      const float dP = fmaf(po.a2, rd.d5_5, fmaf(po.b2c, rd.d3_3, fmaf(po.ab, rd.d4_2,
                       fmaf(po.bc, rd.d2, po.c2 * rd.d1))));
      put(Wv, j, expm1f((p.delta * dP) * rS));
    }
    st4(p.w, g4, Wv);
  }
}

cudaError_t launch_discrete_w(const DiscreteWParams& p, cudaStream_t s) {
  if (p.rows == 0) return cudaSuccess;
  discrete_w_kernel<<<p.rows, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_bwd_pre(const BwdPreParams& p, cudaStream_t s) {
  if (p.rows == 0) return cudaSuccess;
  switch (p.param) {
    case MULAN_PARAM_EPS: return launch_param<MULAN_PARAM_EPS>(p, s);
    case MULAN_PARAM_VEL: return launch_param<MULAN_PARAM_VEL>(p, s);
    default: return launch_param<MULAN_PARAM_VEL_FROM_EPS>(p, s);
  }
}

}  // namespace mulan
