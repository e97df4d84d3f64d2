// mulan_fwd_pre: schedule eval + noising + reconstruction + prior KL, one pass over
// [B, D] (25 B/sub-pixel algorithmic: x1 + a,b,c 12 + eps0 4 + eps 4 -> z_t 4; +4 when the
// loss weight w is saved for the post kernels, +4 for per-pixel g_t).
//
// Reference statements fused here (ldm/model_mulan_epsilon.py; the velocity model runs the
// same lines, ldm/model_mulan_velocity.py:208-236):
//   :300      orig_f = encode(x)                      (ldm/model_vdm.py:274-280)
//   :307-309  g_0, g_1, g_t = gamma(emb, {0,1,t})     (:514-529)
//   :311-313  var_* = sigmoid(g_*)
//   :315-318  z_0_rescaled, loss_recon                (ldm/model_vdm.py:282-303)
//   :322-325  loss_klz (prior KL at t=1)
//   :327-328  z_t = sqrt(1-var_t) f + sqrt(var_t) eps
//   :273-278  _get_score_model_gt (per-row mean or per-pixel g_t)
//   :339-343  g_t_grad = d gamma/dt (saved as w for the post kernels)
//   :361-362  var_0 / var_1 partial sums
//
// Layout: one CTA (256 threads) per example row; each thread owns float4 columns
// tid, tid+256, ... of the row (coalesced 16-B accesses, uchar4 for x).  Per-row t powers
// are staged in shared memory by thread 0.  The gamma-bound constants (gamma_0 = gamma_min
// exactly for the fixed-end polynomial, so exp(+-gamma_0/2), sigmoid(gamma_0),
// sigmoid(gamma_1), log sigmoid(gamma_1) are constants) are computed once per launch on
// the host, correctly rounded, and travel in the kernel parameter block: the reconstruction
// term is sensitive to a 1-ulp change of exp(gamma_0/2) at the 1e-5 level (it decides how
// z_0 rounds), so these constants must not depend on which exp implementation evaluates
// them.  Per-row sums use a fixed-order shuffle tree (deterministic; no atomics).
//
// The kernel is instruction-issue bound, not HBM bound, unless the per-sub-pixel instruction
// count stays near ~110 (profiles/): hence the 3-bin reconstruction window in closed form
// and the branch-free MUFU+Newton math of mulan_common.cuh on the hot path.  Everything
// that decides HOW the reference rounds where it matters (z_0, u = (z_0 - x_k) e^{-g0/2},
// 1 - sigmoid, the prior-KL summand) keeps the reference's op order.
#include "mulan_kernels.h"

namespace mulan {

// ---------------------------------------------------------------------------------------
// Generic reconstruction term: log-softmax over the vocab bins evaluated on a window of
// +-W bins around the bin nearest to z (bins further away have exp(logit - max) < e^-30 and
// cannot move a float32 sum >= 1).  IEEE expf/logf.  Used when W != 1, vocab is not a power
// of two, or the fixed ends do not hold for a sub-pixel.  Returns log p(x | z).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float recon_logprob_generic(int xi, float z, float inv0, int W,
                                                       const VocabInfo& vi) {
  float kf = rintf((z + 1.0f) * vi.half_vocab - 0.5f);   // nearest bin centre (2k+1)/vocab-1
  kf = fminf(fmaxf(kf, 0.0f), vi.vocab_m1);
  const int kc = (int)kf;
  auto logit = [&](int k) {
    const float u = (z - vi.xval(k)) * inv0;
    return -0.5f * (u * u);
  };
  const float lc = logit(kc);
  const float lm = kc > 0 ? logit(kc - 1) : -INFINITY;
  const float lp = kc < vi.vocab - 1 ? logit(kc + 1) : -INFINITY;
  const float m = fmaxf(lc, fmaxf(lm, lp));
  float sum = 0.0f;
  const int k0 = max(kc - W, 0), k1 = min(kc + W, vi.vocab - 1);
  for (int k = k0; k <= k1; ++k) sum += expf(logit(k) - m);
  return (logit(xi) - m) - logf(sum);   // z NaN -> NaN, as in the reference
}

// Fast reconstruction term for W == 1 and vocab a power of two (the shipped configs:
// gamma_0 = -13.3 puts neighbouring bins 6.04 decoder-sigmas apart).
//   u_c = (z - x_c) e^{-g0/2}           reference op order (x_c exact, subtraction, product)
//   l_{c+-1} - l_c = -+ s u_c - s^2/2    closed form, s = (2/vocab) e^{-g0/2}
//   log p(x) = (l_x - l_c) - log(1 + e^{l_{c-1}-l_c} + e^{l_{c+1}-l_c})
// The centre bin is the max up to rounding ties, where log-sum-exp is shift invariant.
__device__ __forceinline__ float recon_logprob_fast(float xf, float f, float z,
                                                    const ReconFast& rc) {
  float kf = rintf(fmaf(z, rc.half_vocab, rc.half_vocab - 0.5f));
  kf = fminf(fmaxf(kf, 0.0f), rc.vocab_m1);
  const float xc = fmaf(kf, rc.two_iv, rc.off);           // exact bin centre
  const float uc = (z - xc) * rc.inv0;
  const float em = kf > 0.0f ? ex2_approx(fmaf(-rc.s2, uc, rc.c0)) : 0.0f;
  const float ep = kf < rc.vocab_m1 ? ex2_approx(fmaf(rc.s2, uc, rc.c0)) : 0.0f;
  const float sum = (1.0f + em) + ep;
  // l_x - l_c = -(u_x - u_c)(u_x + u_c)/2 with u_x - u_c = (k_c - x) s exactly; 0 when x is
  // the nearest bin (all but the |eps_0| > 3 tail)
  const float ds = (kf - xf) * rc.s;
  const float lxc = -ds * fmaf(0.5f, ds, uc);
  return lxc - log_1p_sum(sum);
}

// Rare path: S is zero / denormal / huge / NaN so gamma(0), gamma(1) are not the fixed-end
// constants.  Evaluate this sub-pixel exactly as the reference does (IEEE ops).
struct SlowPix { float gt, wt, lp, kl, v0, v1; };
__device__ __noinline__ SlowPix slow_pixel(const Poly po, float gmin, float delta, int xi,
                                           float f, float e0, const VocabInfo vi) {
  SlowPix o;
  const float S = po.S;
  o.gt = gmin + __fdiv_rn(delta * po.P, S);
  o.wt = __fdiv_rn(delta * (po.q * po.q), S);
  const float g0 = gmin + __fdiv_rn(delta * 0.0f, S);
  const float g1 = gmin + __fdiv_rn(delta * S, S);
  o.v0 = sigmoid_ref(g0);
  o.v1 = sigmoid_ref(g1);
  const float s0 = expf(0.5f * g0), inv0 = expf(-0.5f * g0);
  const float z = f + s0 * e0;
  o.lp = recon_logprob_generic(xi, z, inv0, vi.vocab, vi);  // full vocab
  o.kl = (1.0f - o.v1) * (f * f) + o.v1 - logf(o.v1) - 1.0f;
  return o;
}

// Prior-KL summand when sigmoid(gamma_1) is not the same float for the three possible
// roundings of (delta*S)/S (never the case for the shipped gamma range).
__device__ __noinline__ float2 prior_general(float S, float gmin, float delta, float f) {
  const float g1 = gmin + __fdiv_rn(delta * S, S);
  const float v1 = sigmoid_ref(g1);
  return make_float2((1.0f - v1) * (f * f) + v1 - logf(v1) - 1.0f, v1);
}

template <int GT, bool SAVEW, bool FAST>
__global__ void __launch_bounds__(kThreads, 4)
fwd_pre_kernel(const FwdPreParams p) {
  __shared__ RowT s_rt;
  __shared__ float red[kWarps][5];
  const int row = blockIdx.x;
  const int tid = threadIdx.x;

  if (tid == 0) s_rt = make_row_t(__ldg(p.t + row));
  __syncthreads();
  const RowT rt = s_rt;
  // everything below lives in the kernel parameter (constant) bank: no registers
  const float s0 = p.k.s0, inv0 = p.k.inv0, v0c = p.k.v0;
  const float v1c = p.k.v1, om1 = p.k.om1;
  const float lv1 = p.k.lv1;
  const bool v1_uniform = p.k.v1_uniform != 0;
  const VocabInfo& vi = p.vi;
  const ReconFast& rc = p.rc;
  const float gmin = p.gmin, delta = p.delta;

  const size_t base4 = (size_t)row * p.dim4;
  // logprob, klz summand, g_t, and (only off the fixed-end path) var0 / var1 corrections
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};

  for (int i4 = tid; i4 < p.dim4; i4 += kThreads) {
    const size_t g4 = base4 + i4;
    const float4 A = ld4(p.a, g4), Bv = ld4(p.b, g4), C = ld4(p.c, g4);
    const float4 E0 = ld4(p.eps0, g4), E = ld4(p.eps, g4);
    const uchar4 X = ldx4(p.x, g4);
    float4 Z, Wv, G;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = get(A, j), b = get(Bv, j), c = get(C, j);
      const float e0 = get(E0, j), e = get(E, j);
      const int xi = getx(X, j);
      const float xf = (float)xi;
      const float f = FAST ? fmaf(xf, rc.two_iv, rc.off)  // encode(x), exact for 2^k vocab
                           : vi.xval(xi);
      const Poly po = poly_eval(a, b, c, rt);
      float gt, wt;
      if (scale_in_range(po.S)) {                         // fixed ends are exact constants
        const float rSd = delta * rcp_nr(po.S);
        gt = fmaf(po.P, rSd, gmin);                       // gamma_t
        wt = (po.q * po.q) * rSd;                         // d gamma / dt
        const float z0 = f + s0 * e0;                     // z_0_rescaled (two roundings)
        acc[0] += FAST ? recon_logprob_fast(xf, f, z0, rc)
                       : recon_logprob_generic(xi, z0, inv0, p.W, vi);
        if (v1_uniform) {
          acc[1] += om1 * (f * f) + v1c - lv1 - 1.0f;     // reference op order
        } else {
          const float2 pg = prior_general(po.S, gmin, delta, f);
          acc[1] += pg.x;
          acc[4] += pg.y - v1c;
        }
      } else {
        const SlowPix sp = slow_pixel(po, gmin, delta, xi, f, e0, vi);
        gt = sp.gt; wt = sp.wt;
        acc[0] += sp.lp; acc[1] += sp.kl;
        acc[3] += sp.v0 - v0c; acc[4] += sp.v1 - v1c;
      }
      const float vt = sigmoid_fast(gt);
      const float om = 1.0f - vt;
      const float alpha = sqrt_fast0(om), sigma = sqrt_fast(vt);
      put(Z, j, alpha * f + sigma * e);                   // z_t (two products, one add)
      if (SAVEW) put(Wv, j, wt);
      if (GT == MULAN_GT_PIXEL) put(G, j, gt);
      acc[2] += gt;
    }
    st4(p.z_t, g4, Z);
    if (SAVEW) st4(p.w_save, g4, Wv);
    if (GT == MULAN_GT_PIXEL) st4(p.g_net, g4, G);
  }

  block_sum<5>(acc, red);
  if (tid == 0) {
    const float dimf = (float)(p.dim4 * 4);
    p.loss_recon[row] = -acc[0];
    p.loss_klz[row] = 0.5f * acc[1];
    if (GT == MULAN_GT_MEAN) p.g_net[row] = __fdiv_rn(acc[2], dimf);
    // sum over the row of sigmoid(g_0), sigmoid(g_1): D * constant + corrections
    p.var_sums[2 * row + 0] = dimf * v0c + acc[3];
    p.var_sums[2 * row + 1] = dimf * v1c + acc[4];
  }
}

template <int GT, bool SAVEW>
static cudaError_t launch_w(const FwdPreParams& p, cudaStream_t s) {
  dim3 grid(p.rows), block(kThreads);
  if (p.W == 1 && p.vi.pow2) fwd_pre_kernel<GT, SAVEW, true><<<grid, block, 0, s>>>(p);
  else                       fwd_pre_kernel<GT, SAVEW, false><<<grid, block, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_fwd_pre(const FwdPreParams& p, cudaStream_t s) {
  if (p.rows == 0) return cudaSuccess;
  const bool savew = p.w_save != nullptr;
  if (p.gt_mode == MULAN_GT_MEAN)
    return savew ? launch_w<MULAN_GT_MEAN, true>(p, s) : launch_w<MULAN_GT_MEAN, false>(p, s);
  return savew ? launch_w<MULAN_GT_PIXEL, true>(p, s) : launch_w<MULAN_GT_PIXEL, false>(p, s);
}

}  // namespace mulan
