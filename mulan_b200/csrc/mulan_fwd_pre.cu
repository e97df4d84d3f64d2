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
#include <stdlib.h>
#include <string.h>

#include "mulan_kernels.h"
#include "mulan_rng.cuh"

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

// One float4 column (4 sub-pixels) of one row: everything between the loads and the stores.
// acc: logprob, klz summand, g_t, and (only off the fixed-end path) var0 / var1 corrections.
// Launch constants of the hot path, read from the parameter bank ONCE per thread (before the
// slab loop) so the loop body does not re-issue constant loads for every sub-pixel.
struct PreConsts {
  float s0, inv0, v0c, v1c, om1, lv1, gmin, delta;
  ReconFast rc;
  bool v1_uniform;
};
// The shipped configuration (gamma_min = -13.3, gamma_max = 5, vocab = 256 in both
// ldm/configs/*.py) with every launch constant as a compile-time literal: the kernel then
// carries them as instruction immediates instead of re-loading ~8 constants per sub-pixel
// from the parameter bank (the kernel is instruction-issue bound).  The literals are the
// values make_end_consts / make_recon_fast produce; launch_w() uses this specialisation only
// when the run-time constants match them BIT FOR BIT, so a different gamma range, vocab or
// host libm silently takes the generic kernel.
struct Shipped {
  static constexpr float gmin = -0x1.a9999ap+3f, delta = 0x1.24ccccp+4f;
  static constexpr float s0 = 0x1.533858p-10f, inv0 = 0x1.826468p+9f, v0 = 0x1.c17e14p-20f;
  static constexpr float v1 = 0x1.fc92c2p-1f, om1 = 0x1.b69fp-8f, lv1 = -0x1.b81872p-8f;
  static constexpr float s = 0x1.826468p+2f, s2 = 0x1.16b91ap+3f, c0 = -0x1.a4b06cp+4f;
  static constexpr float two_iv = 0x1p-7f, off = -0x1.fep-1f, half_vocab = 128.0f,
                         vocab_m1 = 255.0f;
};

template <bool BAKED>
__device__ __forceinline__ PreConsts load_pre_consts(const FwdPreParams& p) {
  PreConsts k;
  if (BAKED) {
    k.s0 = Shipped::s0; k.inv0 = Shipped::inv0; k.v0c = Shipped::v0; k.v1c = Shipped::v1;
    k.om1 = Shipped::om1; k.lv1 = Shipped::lv1; k.gmin = Shipped::gmin; k.delta = Shipped::delta;
    k.rc.inv0 = Shipped::inv0; k.rc.s = Shipped::s; k.rc.s2 = Shipped::s2; k.rc.c0 = Shipped::c0;
    k.rc.two_iv = Shipped::two_iv; k.rc.off = Shipped::off;
    k.rc.half_vocab = Shipped::half_vocab; k.rc.vocab_m1 = Shipped::vocab_m1;
    k.v1_uniform = true;
  } else {
    k.s0 = p.k.s0; k.inv0 = p.k.inv0; k.v0c = p.k.v0; k.v1c = p.k.v1; k.om1 = p.k.om1;
    k.lv1 = p.k.lv1; k.gmin = p.gmin; k.delta = p.delta; k.rc = p.rc;
    k.v1_uniform = p.k.v1_uniform != 0;
  }
  return k;
}

static bool is_shipped(const FwdPreParams& p) {
  auto eq = [](float a, float b) { return memcmp(&a, &b, sizeof(float)) == 0; };
  return p.k.v1_uniform != 0 && p.W == 1 && p.vi.vocab == 256 && eq(p.gmin, Shipped::gmin) &&
         eq(p.delta, Shipped::delta) && eq(p.k.s0, Shipped::s0) && eq(p.k.inv0, Shipped::inv0) &&
         eq(p.k.v0, Shipped::v0) && eq(p.k.v1, Shipped::v1) && eq(p.k.om1, Shipped::om1) &&
         eq(p.k.lv1, Shipped::lv1) && eq(p.rc.s, Shipped::s) && eq(p.rc.s2, Shipped::s2) &&
         eq(p.rc.c0, Shipped::c0) && eq(p.rc.two_iv, Shipped::two_iv) &&
         eq(p.rc.off, Shipped::off);
}

// The general per-sub-pixel statements: the fixed-end fast path when S is in range, the IEEE
// slow path otherwise, the per-pixel prior when sigmoid(gamma_1) is not uniform.
template <bool FAST>
__device__ __forceinline__ void pre_pixel_general(const FwdPreParams& p, const PreConsts& kc,
                                                  const Poly& po, int xi, float xf, float f,
                                                  float e0, float& gt, float& wt,
                                                  float (&acc)[5]) {
  const VocabInfo& vi = p.vi;
  if (scale_in_range(po.S)) {                             // fixed ends are exact constants
    const float rSd = kc.delta * rcp_nr(po.S);
    gt = fmaf(po.P, rSd, kc.gmin);                        // gamma_t
    wt = (po.q * po.q) * rSd;                             // d gamma / dt
    const float z0 = f + kc.s0 * e0;                      // z_0_rescaled (two roundings)
    acc[0] += FAST ? recon_logprob_fast(xf, f, z0, kc.rc)
                   : recon_logprob_generic(xi, z0, kc.inv0, p.W, vi);
    if (kc.v1_uniform) {
      acc[1] += kc.om1 * (f * f) + kc.v1c - kc.lv1 - 1.0f;   // reference op order
    } else {
      const float2 pg = prior_general(po.S, kc.gmin, kc.delta, f);
      acc[1] += pg.x;
      acc[4] += pg.y - kc.v1c;
    }
  } else {
    const SlowPix sp = slow_pixel(po, kc.gmin, kc.delta, xi, f, e0, vi);
    gt = sp.gt; wt = sp.wt;
    acc[0] += sp.lp; acc[1] += sp.kl;
    acc[3] += sp.v0 - kc.v0c; acc[4] += sp.v1 - kc.v1c;
  }
}

// One float4 column (4 sub-pixels) of one row: everything between the loads and the stores.
// acc: logprob, klz summand, g_t, and (only off the fixed-end path) var0 / var1 corrections.
//
// The four sub-pixels are independent dependency chains of ~100 instructions each with six
// MUFU results in series (rcp -> ex2, ex2 -> lg2 -> ex2 -> rcp -> rsqrt, rsqrt).  Written as a
// per-sub-pixel `if (S in range) ... else slow_pixel()` they compile to four SERIAL blocks
// separated by branches (round 1: issue slots 70 % busy at 47 % occupancy -- each warp waits on
// its own MUFU latencies).  Here the range test is made once for the whole column, so the common
// case is ONE straight-line block in which the compiler interleaves the four chains; a column
// with any out-of-range S (or a non-uniform sigmoid(gamma_1)) takes the general code.  Every
// sub-pixel runs the same statements in the same per-accumulator order either way: bit-identical.
template <int GT, bool SAVEW, bool FAST, bool CRAW = false>
__device__ __forceinline__ void pre_column(const FwdPreParams& p, const PreConsts& kc,
                                           const RowT& rt, const float4 A,
                                           const float4 Bv, const float4 Cin, const float4 E0,
                                           const float4 E, const uchar4 X, size_t g4,
                                           float (&acc)[5]) {
  // MULAN_FLAG_C_RAW: c = 1e-3 + softplus(pre-activation), ldm/model_mulan_epsilon.py:537
  const float4 C = CRAW ? c_from_raw4(Cin) : Cin;
  const VocabInfo& vi = p.vi;
  const ReconFast& rc = kc.rc;
  Poly po[4];
  float xf[4], f[4], gt[4], wt[4];
  int xi[4];
  bool all_fast = FAST && kc.v1_uniform;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    xi[j] = getx(X, j);
    xf[j] = (float)xi[j];
    f[j] = FAST ? fmaf(xf[j], rc.two_iv, rc.off)          // encode(x), exact for 2^k vocab
                : vi.xval(xi[j]);
    po[j] = poly_eval(get(A, j), get(Bv, j), get(C, j), rt);
    all_fast = all_fast && scale_in_range(po[j].S);
  }
  if (all_fast) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float rSd = kc.delta * rcp_nr(po[j].S);
      gt[j] = fmaf(po[j].P, rSd, kc.gmin);                // gamma_t
      wt[j] = (po[j].q * po[j].q) * rSd;                  // d gamma / dt
      const float z0 = f[j] + kc.s0 * get(E0, j);         // z_0_rescaled (two roundings)
      acc[0] += recon_logprob_fast(xf[j], f[j], z0, rc);
      acc[1] += kc.om1 * (f[j] * f[j]) + kc.v1c - kc.lv1 - 1.0f;   // reference op order
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      pre_pixel_general<FAST>(p, kc, po[j], xi[j], xf[j], f[j], get(E0, j), gt[j], wt[j], acc);
  }
  float4 Z, Wv, G;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float vt = sigmoid_fast(gt[j]);
    const float om = 1.0f - vt;
    const float alpha = sqrt_fast0(om), sigma = sqrt_fast(vt);
    put(Z, j, alpha * f[j] + sigma * get(E, j));          // z_t (two products, one add)
    if (SAVEW) put(Wv, j, wt[j]);
    if (GT == MULAN_GT_PIXEL) put(G, j, gt[j]);
    acc[2] += gt[j];
  }
  st4(p.z_t, g4, Z);
  if (SAVEW) st4(p.w_save, g4, Wv);
  if (GT == MULAN_GT_PIXEL) st4(p.g_net, g4, G);
}

// Row epilogue: deterministic CTA-wide sums, per-example outputs written by thread 0.
template <int GT, int NW = kWarps>
__device__ __forceinline__ void pre_row_end(const FwdPreParams& p, int row, float (&acc)[5],
                                            float (*red)[5]) {
  block_sum<5, NW>(acc, red);
  if (threadIdx.x == 0) {
    const float dimf = (float)(p.dim4 * 4);
    p.loss_recon[row] = -acc[0];
    p.loss_klz[row] = 0.5f * acc[1];
    if (GT == MULAN_GT_MEAN) p.g_net[row] = __fdiv_rn(acc[2], dimf);
    // sum over the row of sigmoid(g_0), sigmoid(g_1): D * constant + corrections
    p.var_sums[2 * row + 0] = dimf * p.k.v0 + acc[3];
    p.var_sums[2 * row + 1] = dimf * p.k.v1 + acc[4];
  }
}

// ---------------------------------------------------------------------------------------
// Direct-load kernel: one CTA (NT threads) per row, operands loaded straight into registers
// (LDG.128).  Serves any dim (multiple of 4), any vocab / window, any 4-byte aligned x.
// KIND: 0 generic window, 1 closed-form 3-bin term with the launch constants in the parameter
// bank, 2 the same with the shipped configuration's constants as immediates.
//
// Shape (NT threads, MINB resident CTAs per SM), measured on B200 at 16384 rows
// (profiles/r2_fwd_pre_variants.md): the kernel needs 64 registers, so 32 warps per SM are
// resident whatever the CTA size; 128-thread CTAs (8 per SM) beat 256-thread ones (4 per SM) by
// 5 % (epsilon form) / 9 % (plain velocity) -- the per-row prologue and the row-end barrier of a
// CTA stall a quarter of the SM's warps instead of half.  Capping the registers for more CTAs
// (48 registers: 5 x 256 or 10 x 128) spills and loses 10-15 %; prefetching the next column's
// operands into registers (80 registers, 24 warps) loses 15-25 %; one warp per row with a
// shared-memory table of encode(x) / the prior summand (no barrier, -10 instructions per
// sub-pixel) and FRND / I2F-free arithmetic are no faster: at 6.2-6.4 TB/s of DRAM traffic,
// 76 % issue utilisation and 52 % XU utilisation the kernel sits against all three limits.
// ---------------------------------------------------------------------------------------
struct PreCol {
  float4 A, B, C, E0, E;
  uchar4 X;
};
__device__ __forceinline__ PreCol load_pre_col(const FwdPreParams& p, size_t g4, size_t n4) {
  PreCol c;
  c.A = ld4(p.a, g4); c.B = ld4(p.b, g4); c.C = ld4(p.c, g4);
  c.E0 = ld4(p.eps0, n4); c.E = ld4(p.eps, n4);
  c.X = ldx4(p.x, g4);
  return c;
}

template <int GT, bool SAVEW, int KIND, bool CRAW, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
fwd_pre_kernel(const FwdPreParams p) {
  constexpr bool FAST = KIND != 0, BAKED = KIND == 2;
  constexpr int NW = NT / 32;
  __shared__ float red[NW][5];
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  pdl_release_dependents();
  const PreConsts kc = load_pre_consts<BAKED>(p);
  pdl_wait_for_primary();
  // Row constants.  Latency shape (768 threads, one column each): every thread forms the t powers
  // itself -- no barrier between the CTA's start and its first operand loads.  Throughput
  // shapes: staged through shared memory by thread 0; held in registers by all 128 threads they
  // cost the 64-register kernel six more instructions per sub-pixel in spills (measured: 0.218 ->
  // 0.231 ms at 16384 rows).
  RowT rt;
  if constexpr (NT >= kLatencyThreads) {
    rt = make_row_t(__ldg(p.t + row));
  } else {
    __shared__ RowT s_rt;
    if (tid == 0) s_rt = make_row_t(__ldg(p.t + row));
    __syncthreads();
    rt = s_rt;
  }
  const size_t base4 = (size_t)row * p.dim4;
  // eps_0 / eps broadcast over the batch (dense-VLB evaluation: every image shares one key)
  const size_t nbase4 = (size_t)(p.noise_rows > 0 ? row % p.noise_rows : row) * p.dim4;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i4 = tid; i4 < p.dim4; i4 += NT) {
    const PreCol c = load_pre_col(p, base4 + i4, nbase4 + i4);
    pre_column<GT, SAVEW, FAST, CRAW>(p, kc, rt, c.A, c.B, c.C, c.E0, c.E, c.X, base4 + i4, acc);
  }
  pre_row_end<GT, NW>(p, row, acc, red);
}

// ---------------------------------------------------------------------------------------
// In-kernel draws (SURVEY.md 8f "next" row 2 as written): eps_0 and eps are GENERATED here from
// the raw threefry keys that make_rng('sample') hands to jax.random.normal(rng, f.shape)
// (ldm/model_mulan_epsilon.py:315, :327) instead of being read from HBM.
//
// JAX's counter layout pairs element e with element e + N/2 (one threefry2x32 block yields both
// words), so a CTA owns the ROW PAIR (r, r + B/2): for every float4 column it runs four blocks
// per draw and gets the four normals of row r and the four of row r + B/2 -- no round wasted,
// bit-identical to mulan_rng_normal(key, B*D) reshaped to [B, D].  eps_out (needed again by the
// post / bwd_pre kernels) and eps0_out are optional.
// Cost: ~95 thread-instructions per normal-pair-share on top of the ~105 of the ELBO arithmetic
// -- the kernel is issue bound at ~2.7x the read version's time; whether that pays depends on
// what produced the arrays it replaces (profiles/r2_variants.md: against stand-alone draws that
// write 8 B and are read back, it is on par; against draws that already exist, it loses).
// ---------------------------------------------------------------------------------------
template <int GT, bool SAVEW, bool BAKED, bool CRAW>
__global__ void __launch_bounds__(128, 4)
fwd_pre_keyed_kernel(const FwdPreKeyedParams q) {
  constexpr int NT = 128, NW = NT / 32;
  const FwdPreParams& p = q.p;
  __shared__ float red[NW][5];
  const int half_rows = p.rows / 2;
  const int tid = threadIdx.x;
  pdl_release_dependents();
  const PreConsts kc = load_pre_consts<BAKED>(p);
  pdl_wait_for_primary();
  const int row_lo = blockIdx.x, row_hi = blockIdx.x + half_rows;
  const RowT rt_lo = make_row_t(__ldg(p.t + row_lo)), rt_hi = make_row_t(__ldg(p.t + row_hi));
  const size_t b_lo = (size_t)row_lo * p.dim4, b_hi = (size_t)row_hi * p.dim4;
  const uint32_t half = (uint32_t)half_rows * (uint32_t)(p.dim4 * 4);      // N / 2
  float acc_lo[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, acc_hi[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i4 = tid; i4 < p.dim4; i4 += NT) {
    float4 E0lo, E0hi, Elo, Ehi;
    const uint32_t e = (uint32_t)row_lo * (uint32_t)(p.dim4 * 4) + 4u * (uint32_t)i4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t x0 = e + j, x1 = e + j + half;
      threefry2x32(q.k_eps0[0], q.k_eps0[1], x0, x1);
      put(E0lo, j, bits_to_normal(x0)); put(E0hi, j, bits_to_normal(x1));
      x0 = e + j; x1 = e + j + half;
      threefry2x32(q.k_eps[0], q.k_eps[1], x0, x1);
      put(Elo, j, bits_to_normal(x0)); put(Ehi, j, bits_to_normal(x1));
    }
    if (q.eps_out != nullptr) { st4(q.eps_out, b_lo + i4, Elo); st4(q.eps_out, b_hi + i4, Ehi); }
    if (q.eps0_out != nullptr) { st4(q.eps0_out, b_lo + i4, E0lo); st4(q.eps0_out, b_hi + i4, E0hi); }
    {
      const size_t g4 = b_lo + i4;
      pre_column<GT, SAVEW, true, CRAW>(p, kc, rt_lo, ld4(p.a, g4), ld4(p.b, g4), ld4(p.c, g4),
                                        E0lo, Elo, ldx4(p.x, g4), g4, acc_lo);
    }
    {
      const size_t g4 = b_hi + i4;
      pre_column<GT, SAVEW, true, CRAW>(p, kc, rt_hi, ld4(p.a, g4), ld4(p.b, g4), ld4(p.c, g4),
                                        E0hi, Ehi, ldx4(p.x, g4), g4, acc_hi);
    }
  }
  pre_row_end<GT, NW>(p, row_lo, acc_lo, red);
  __syncthreads();
  pre_row_end<GT, NW>(p, row_hi, acc_hi, red);
}

template <int GT, bool SAVEW>
static cudaError_t launch_keyed_w(const FwdPreKeyedParams& q, cudaStream_t s) {
  const FwdPreParams& p = q.p;
  const bool baked = is_shipped(p);
  const int grid = p.rows / 2;
  const bool pdl = p.pdl != 0;
#define MULAN_KEYED(BAKED, CRAW) \
  return launch_kernel(fwd_pre_keyed_kernel<GT, SAVEW, BAKED, CRAW>, grid, 128, s, pdl, q)
  if (p.c_raw) { if (baked) MULAN_KEYED(true, true); MULAN_KEYED(false, true); }
  if (baked) MULAN_KEYED(true, false);
  MULAN_KEYED(false, false);
#undef MULAN_KEYED
}

// ---------------------------------------------------------------------------------------
// TMA-pipelined kernel (the shipped configs: dim % 1024 == 0, 16-byte aligned operands).
// Persistent CTAs walk rows blockIdx.x, += gridDim.x; a row is consumed in slabs of 256 float4
// columns.  Thread 0 moves each slab global -> shared with six cp.async.bulk copies (4 KB per
// float array, 1 KB of x) that complete on a `full` mbarrier; every thread then reads its own
// 16-byte slots (conflict-free LDS.128), one lane per warp arrives on the `empty` mbarrier,
// and the slab after next is requested as soon as its buffer is free.  DRAM latency is hidden
// by a full slab of arithmetic with no registers held in flight and no per-thread address
// arithmetic for the loads.
// ---------------------------------------------------------------------------------------
struct __align__(128) PreSlab {
  float4 f[5][kThreads];   // a, b, c, eps0, eps
  uchar4 x[kThreads];
};
constexpr unsigned kPreSlabBytes = 5 * kThreads * 16 + kThreads * 4;

#ifndef MULAN_TMA_CTAS
#define MULAN_TMA_CTAS 3
#endif
template <int GT, bool SAVEW, bool BAKED>
__global__ void __launch_bounds__(kThreads, MULAN_TMA_CTAS)
fwd_pre_tma_kernel(const FwdPreParams p) {
  __shared__ PreSlab slab[2];
  __shared__ __align__(8) uint64_t full[2], empty[2];
  __shared__ float red[2][kWarps][5];
  const int tid = threadIdx.x;
  const int nslab = p.dim4 / kThreads;
  const int my_rows = (p.rows - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_rows * nslab;
  pdl_release_dependents();
  pdl_wait_for_primary();

  if (tid == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1);
    mbar_init(&empty[0], kWarps); mbar_init(&empty[1], kWarps);
    mbar_fence_init();
  }
  __syncthreads();

  // producer cursor (thread 0 only): the next slab to request
  int prow = blockIdx.x, pslab = 0;
  auto request = [&](int stage) {
    const size_t g4 = (size_t)prow * p.dim4 + (size_t)pslab * kThreads;
    mbar_expect_tx(&full[stage], kPreSlabBytes);
    bulk_g2s(&slab[stage].f[0][0], reinterpret_cast<const float4*>(p.a) + g4, kThreads * 16, &full[stage]);
    bulk_g2s(&slab[stage].f[1][0], reinterpret_cast<const float4*>(p.b) + g4, kThreads * 16, &full[stage]);
    bulk_g2s(&slab[stage].f[2][0], reinterpret_cast<const float4*>(p.c) + g4, kThreads * 16, &full[stage]);
    bulk_g2s(&slab[stage].f[3][0], reinterpret_cast<const float4*>(p.eps0) + g4, kThreads * 16, &full[stage]);
    bulk_g2s(&slab[stage].f[4][0], reinterpret_cast<const float4*>(p.eps) + g4, kThreads * 16, &full[stage]);
    bulk_g2s(&slab[stage].x[0], reinterpret_cast<const uchar4*>(p.x) + g4, kThreads * 4, &full[stage]);
    if (++pslab == nslab) { pslab = 0; prow += gridDim.x; }
  };
  if (tid == 0 && total > 0) request(0);

  const PreConsts kc = load_pre_consts<BAKED>(p);
  int row = blockIdx.x, s = 0, parity = 0;
  RowT rt;
  float acc[5];
  for (int it = 0; it < total; ++it) {
    const int stage = it & 1;
    if (tid == 0 && it + 1 < total) {
      // the other buffer held slab it-1: wait until every warp has read it, then refill
      if (it >= 1) mbar_wait(&empty[stage ^ 1], ((it - 1) >> 1) & 1);
      request(stage ^ 1);
    }
    if (s == 0) {
      rt = make_row_t(__ldg(p.t + row));                  // per-example t powers
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[k] = 0.f;
    }
    mbar_wait(&full[stage], (it >> 1) & 1);               // slab `it` has landed
    const float4 A = slab[stage].f[0][tid], Bv = slab[stage].f[1][tid], C = slab[stage].f[2][tid];
    const float4 E0 = slab[stage].f[3][tid], E = slab[stage].f[4][tid];
    const uchar4 X = slab[stage].x[tid];
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[stage]);      // this warp is done with the buffer
    const size_t g4 = (size_t)row * p.dim4 + (size_t)s * kThreads + tid;
    pre_column<GT, SAVEW, true>(p, kc, rt, A, Bv, C, E0, E, X, g4, acc);
    if (++s == nslab) {
      pre_row_end<GT>(p, row, acc, red[parity]);
      parity ^= 1;
      s = 0;
      row += gridDim.x;
    }
  }
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

// Shape of the default direct-load kernel (threads per CTA, minimum resident CTAs per SM),
// chosen by measurement on B200 (profiles/r2_fwd_pre_variants.md).
#ifndef MULAN_PRE_NT
#define MULAN_PRE_NT 128
#define MULAN_PRE_MINB 8
#endif

// MULAN_FWD_PRE_V=<n> (read per launch; A/B measurements only) selects one of the alternative
// shapes below for the shipped configuration's kernel.  Per-pixel outputs are bit-identical
// across shapes (same arithmetic); per-row sums differ by float32 summation order.
static int experimental_shape() {
  const char* e = getenv("MULAN_FWD_PRE_V");
  return (e == nullptr || e[0] == '\0') ? -1 : atoi(e);
}

template <int GT, bool SAVEW, int KIND, bool CRAW>
static cudaError_t launch_shape(const FwdPreParams& p, cudaStream_t s) {
  if constexpr (GT == MULAN_GT_MEAN && KIND == 2 && !CRAW) {
#define MULAN_SHAPE(NT, MINB) \
    return launch_kernel(fwd_pre_kernel<GT, SAVEW, KIND, CRAW, NT, MINB>, p.rows, NT, s, \
                         p.pdl != 0, p)
    switch (experimental_shape()) {
      case 0: MULAN_SHAPE(256, 4);
      case 1: MULAN_SHAPE(256, 5);
      case 5: MULAN_SHAPE(128, 8);
      case 12: MULAN_SHAPE(64, 16);
      default: break;
    }
#undef MULAN_SHAPE
  }
  // Launches that put at most one row on an SM are latency bound: one float4 column per thread
  // (768 threads for D = 3072) instead of six; up to four rows per SM: three columns per thread.
  if (KIND != 0 && shape_rows(p.rows) <= latency_rows() && p.dim4 <= kLatencyThreads)
    return launch_kernel(fwd_pre_kernel<GT, SAVEW, KIND, CRAW, kLatencyThreads, 1>, p.rows,
                         kLatencyThreads, s, p.pdl != 0, p);
  if (KIND != 0 && shape_rows(p.rows) <= 4 * latency_rows())
    return launch_kernel(fwd_pre_kernel<GT, SAVEW, KIND, CRAW, 256, 4>, p.rows, 256, s,
                         p.pdl != 0, p);
  return launch_kernel(fwd_pre_kernel<GT, SAVEW, KIND, CRAW, MULAN_PRE_NT, MULAN_PRE_MINB>,
                       p.rows, MULAN_PRE_NT, s, p.pdl != 0, p);
}

template <int GT, bool SAVEW>
static cudaError_t launch_w(const FwdPreParams& p, cudaStream_t s) {
  const bool fast = p.W == 1 && p.vi.pow2;
  // MULAN_NO_BAKED=1 forces the generic constants (A/B measurements, tests)
  static const int no_baked = [] {
    const char* e = getenv("MULAN_NO_BAKED");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  const bool baked = fast && !no_baked && is_shipped(p);
  // The TMA-pipelined kernel is opt-in (MULAN_FWD_PRE_TMA=1): measured on B200 it hides DRAM
  // latency (issue utilisation 71 % -> 79 %) but executes 144 instead of 120 instructions per
  // sub-pixel, and this kernel is issue-bound: 0.260 ms vs 0.243 ms (profiles/).
  static const int use_tma = [] {
    const char* e = getenv("MULAN_FWD_PRE_TMA");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  if (fast && use_tma && p.dim4 % kThreads == 0 && aligned16(p.x) && !p.c_raw &&
      p.noise_rows == 0) {
    // resident CTAs of this variant on the CURRENT device (queried per launch: no process-wide
    // cache, an XLA-style host calls in from one thread per device)
    const int max_ctas = resident_ctas((const void*)fwd_pre_tma_kernel<GT, SAVEW, false>);
    const int grid = p.rows < max_ctas ? p.rows : max_ctas;
    if (baked) fwd_pre_tma_kernel<GT, SAVEW, true><<<grid, kThreads, 0, s>>>(p);
    else       fwd_pre_tma_kernel<GT, SAVEW, false><<<grid, kThreads, 0, s>>>(p);
    return cudaGetLastError();
  }
  const int kind = baked ? 2 : (fast ? 1 : 0);
  if (p.c_raw) {
    switch (kind) {
      case 2: return launch_shape<GT, SAVEW, 2, true>(p, s);
      case 1: return launch_shape<GT, SAVEW, 1, true>(p, s);
      default: return launch_shape<GT, SAVEW, 0, true>(p, s);
    }
  }
  switch (kind) {
    case 2: return launch_shape<GT, SAVEW, 2, false>(p, s);
    case 1: return launch_shape<GT, SAVEW, 1, false>(p, s);
    default: return launch_shape<GT, SAVEW, 0, false>(p, s);
  }
}

// Which fwd_pre kernel launch_w() would pick for these launch constants (host-only query):
// 0 generic (windowed log-softmax), 1 closed-form 3-bin reconstruction with the constants in the
// parameter bank, 2 the same with the shipped configuration's constants as immediates.
int fwd_pre_variant(const FwdPreParams& p) {
  const bool fast = p.W == 1 && p.vi.pow2;
  if (!fast) return 0;
  return is_shipped(p) ? 2 : 1;
}

// rows even, closed-form reconstruction term with a uniform prior (the shipped configurations)
cudaError_t launch_fwd_pre_keyed(const FwdPreKeyedParams& q, cudaStream_t s) {
  const FwdPreParams& p = q.p;
  if (p.rows == 0) return cudaSuccess;
  const bool fast = p.W == 1 && p.vi.pow2;
  if (!fast || (p.rows & 1) || p.noise_rows != 0) return cudaErrorNotSupported;
  const bool savew = p.w_save != nullptr;
  if (p.gt_mode == MULAN_GT_MEAN)
    return savew ? launch_keyed_w<MULAN_GT_MEAN, true>(q, s)
                 : launch_keyed_w<MULAN_GT_MEAN, false>(q, s);
  return savew ? launch_keyed_w<MULAN_GT_PIXEL, true>(q, s)
               : launch_keyed_w<MULAN_GT_PIXEL, false>(q, s);
}

cudaError_t launch_fwd_pre(const FwdPreParams& p, cudaStream_t s) {
  if (p.rows == 0) return cudaSuccess;
  const bool savew = p.w_save != nullptr;
  if (p.gt_mode == MULAN_GT_MEAN)
    return savew ? launch_w<MULAN_GT_MEAN, true>(p, s) : launch_w<MULAN_GT_MEAN, false>(p, s);
  return savew ? launch_w<MULAN_GT_PIXEL, true>(p, s) : launch_w<MULAN_GT_PIXEL, false>(p, s);
}

}  // namespace mulan
