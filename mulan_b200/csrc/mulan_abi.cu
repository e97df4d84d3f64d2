// extern "C" surface of libmulan_b200.so (see include/mulan_b200.h): argument validation,
// parameter-block assembly, launches, and the host-buffer convenience entry.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mulan_kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail(mulan_status st, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return (int)st;
}

int cuda_fail(const char* what, cudaError_t e) {
  return fail(MULAN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

bool aligned(const void* p, size_t n) { return (reinterpret_cast<uintptr_t>(p) % n) == 0; }

int check_desc(const mulan_desc* d, const char* fn) {
  if (d == nullptr) return fail(MULAN_ERR_INVALID_ARG, "%s: desc is NULL", fn);
  if (d->rows < 0) return fail(MULAN_ERR_INVALID_ARG, "%s: rows=%d < 0", fn, d->rows);
  if (d->dim <= 0) return fail(MULAN_ERR_INVALID_ARG, "%s: dim=%d <= 0", fn, d->dim);
  if (d->dim % 4 != 0)
    return fail(MULAN_ERR_ALIGNMENT, "%s: dim=%d is not a multiple of 4", fn, d->dim);
  if (d->vocab < 2 || d->vocab > 65536)
    return fail(MULAN_ERR_INVALID_ARG, "%s: vocab=%d outside [2,65536]", fn, d->vocab);
  if (d->param < MULAN_PARAM_EPS || d->param > MULAN_PARAM_VEL_FROM_EPS)
    return fail(MULAN_ERR_INVALID_ARG, "%s: param=%d is not a mulan_param", fn, d->param);
  if (d->gt_mode != MULAN_GT_MEAN && d->gt_mode != MULAN_GT_PIXEL)
    return fail(MULAN_ERR_INVALID_ARG, "%s: gt_mode=%d is not a mulan_gt_mode", fn, d->gt_mode);
  if (d->n_timesteps < 0)
    return fail(MULAN_ERR_INVALID_ARG, "%s: n_timesteps=%d < 0", fn, d->n_timesteps);
  if (!(d->gamma_max > d->gamma_min))
    return fail(MULAN_ERR_INVALID_ARG, "%s: gamma_max must exceed gamma_min", fn);
  if (d->flags & ~(uint32_t)(MULAN_FLAG_C_RAW | MULAN_FLAG_PDL))
    return fail(MULAN_ERR_INVALID_ARG, "%s: flags=0x%x has unknown bits", fn, d->flags);
  if (d->noise_rows < 0)
    return fail(MULAN_ERR_INVALID_ARG, "%s: noise_rows=%d < 0", fn, d->noise_rows);
  if ((d->flags & MULAN_FLAG_C_RAW) && (d->vocab & (d->vocab - 1)) != 0)
    return fail(MULAN_ERR_UNSUPPORTED, "%s: MULAN_FLAG_C_RAW needs a power-of-two vocab", fn);
  if (d->n_timesteps > 0 && d->param != MULAN_PARAM_EPS)
    return fail(MULAN_ERR_UNSUPPORTED,
                "%s: discrete time (sm_n_timesteps=%d > 0) exists only for the epsilon model; the "
                "reference's velocity model asserts T == 0 (ldm/model_mulan_velocity.py:255)", fn,
                d->n_timesteps);
  return 0;
}

#define REQ_PTR(p, fn)                                                           \
  do {                                                                           \
    if ((p) == nullptr) return fail(MULAN_ERR_INVALID_ARG, "%s: %s is NULL", fn, #p); \
  } while (0)
#define REQ_VEC(p, fn)                                                           \
  do {                                                                           \
    REQ_PTR(p, fn);                                                              \
    if (!aligned((p), 16))                                                       \
      return fail(MULAN_ERR_ALIGNMENT, "%s: %s is not 16-byte aligned", fn, #p); \
  } while (0)
#define OPT_VEC(p, fn)                                                           \
  do {                                                                           \
    if ((p) != nullptr && !aligned((p), 16))                                     \
      return fail(MULAN_ERR_ALIGNMENT, "%s: %s is not 16-byte aligned", fn, #p); \
  } while (0)
#define REQ_X(p, fn)                                                             \
  do {                                                                           \
    REQ_PTR(p, fn);                                                              \
    if (!aligned((p), 4))                                                        \
      return fail(MULAN_ERR_ALIGNMENT, "%s: %s is not 4-byte aligned", fn, #p);  \
  } while (0)

int flag(const mulan_desc* d, uint32_t f) { return (d->flags & f) ? 1 : 0; }
float f32_gmin(const mulan_desc* d) { return (float)d->gamma_min; }
float f32_delta(const mulan_desc* d) { return (float)(d->gamma_max - d->gamma_min); }

// Half-width of the bin window for the reconstruction log-softmax at gamma_0 = gamma_min:
// with bin spacing s = (2/vocab) exp(-gamma_0/2) in units of the decoder's stdev, a bin j
// steps away from the nearest one has exp(logit - max) <= exp(-s^2 j (j-1) / 2).
int recon_window(const mulan_desc* d) {
  const double s = (2.0 / d->vocab) * exp(-0.5 * (double)f32_gmin(d));
  int W = 1;
  while (W < d->vocab - 1 && 0.5 * s * s * (double)W * (double)(W + 1) < 30.0) ++W;
  return W;
}

}  // namespace

namespace mulan {
thread_local int tl_shape_rows = 0;
void set_last_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

int kernel_param(int param) {
  static const int literal = [] {
    const char* e = getenv("MULAN_VFE_LITERAL");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  return (param == MULAN_PARAM_VEL_FROM_EPS && !literal) ? (int)MULAN_PARAM_EPS : param;
}
}  // namespace mulan

extern "C" {

const char* mulan_last_error(void) { return g_err; }
int mulan_abi_version(void) { return MULAN_ABI_VERSION; }
int mulan_kernel_param(int32_t param) { return mulan::kernel_param(param); }

int mulan_fwd_pre(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                  const float* c, const float* t, const float* eps0, const float* eps,
                  float* z_t, float* g_net, float* w_save, float* loss_recon,
                  float* loss_klz_prior, float* var_sums, void* stream) {
  const char* fn = "mulan_fwd_pre";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return 0;
  REQ_X(x, fn);
  REQ_VEC(a, fn); REQ_VEC(b, fn); REQ_VEC(c, fn); REQ_PTR(t, fn);
  REQ_VEC(eps0, fn); REQ_VEC(eps, fn); REQ_VEC(z_t, fn);
  REQ_PTR(g_net, fn); OPT_VEC(w_save, fn);
  if (d->gt_mode == MULAN_GT_PIXEL) REQ_VEC(g_net, fn);
  REQ_PTR(loss_recon, fn); REQ_PTR(loss_klz_prior, fn); REQ_PTR(var_sums, fn);
  mulan::FwdPreParams p;
  p.x = x; p.a = a; p.b = b; p.c = c; p.t = t; p.eps0 = eps0; p.eps = eps;
  p.z_t = z_t; p.g_net = g_net; p.w_save = w_save;
  p.loss_recon = loss_recon; p.loss_klz = loss_klz_prior; p.var_sums = var_sums;
  p.rows = d->rows; p.dim4 = d->dim / 4; p.gt_mode = d->gt_mode;
  p.c_raw = flag(d, MULAN_FLAG_C_RAW); p.pdl = flag(d, MULAN_FLAG_PDL);
  p.noise_rows = d->noise_rows;
  p.W = recon_window(d);
  p.gmin = f32_gmin(d); p.delta = f32_delta(d);
  p.k = mulan::make_end_consts(p.gmin, p.delta);
  p.vi = mulan::make_vocab(d->vocab);
  p.rc = mulan::make_recon_fast(p.k, p.vi);
  if (d->n_timesteps > 0 && w_save == nullptr)
    return fail(MULAN_ERR_INVALID_ARG, "%s: sm_n_timesteps > 0 needs w_save (the discrete-time "
                "loss weight is produced here and consumed by the post kernels)", fn);
  cudaError_t e = mulan::launch_fwd_pre(p, (cudaStream_t)stream);
  if (e == cudaSuccess && d->n_timesteps > 0) {
    mulan::DiscreteWParams q;
    q.a = a; q.b = b; q.c = c; q.t = t; q.w = w_save;
    q.rows = d->rows; q.dim4 = d->dim / 4; q.gmin = p.gmin; q.delta = p.delta;
    q.c_raw = p.c_raw;
    q.inv_T = (float)(1.0 / (double)d->n_timesteps);
    e = mulan::launch_discrete_w(q, (cudaStream_t)stream);
  }
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

// Launch constants of fwd_pre for a descriptor, optionally with caller-supplied end constants.
static void fill_pre_consts(const mulan_desc* d, const mulan_end_consts* kc, mulan::FwdPreParams* p) {
  p->W = recon_window(d);
  p->gmin = f32_gmin(d); p->delta = f32_delta(d);
  p->k = mulan::make_end_consts(p->gmin, p->delta);
  if (kc != nullptr) {
    p->k.s0 = kc->exp_half_g0; p->k.inv0 = kc->exp_neg_half_g0; p->k.v0 = kc->sigmoid_g0;
    p->k.v1 = kc->sigmoid_g1; p->k.om1 = 1.0f - kc->sigmoid_g1; p->k.lv1 = kc->log_sigmoid_g1;
  }
  p->vi = mulan::make_vocab(d->vocab);
  p->rc = mulan::make_recon_fast(p->k, p->vi);
}

int mulan_host_end_consts(const mulan_desc* d, mulan_end_consts* out) {
  const char* fn = "mulan_host_end_consts";
  if (int r = check_desc(d, fn)) return r;
  REQ_PTR(out, fn);
  const mulan::EndConsts k = mulan::make_end_consts(f32_gmin(d), f32_delta(d));
  out->exp_half_g0 = k.s0; out->exp_neg_half_g0 = k.inv0; out->sigmoid_g0 = k.v0;
  out->sigmoid_g1 = k.v1; out->log_sigmoid_g1 = k.lv1;
  return 0;
}

int mulan_fwd_pre_variant_consts(const mulan_desc* d, const mulan_end_consts* kc) {
  const char* fn = "mulan_fwd_pre_variant_consts";
  if (int r = check_desc(d, fn)) return r;
  mulan::FwdPreParams p;
  memset(&p, 0, sizeof(p));
  fill_pre_consts(d, kc, &p);
  return mulan::fwd_pre_variant(p);
}

int mulan_fwd_pre_consts(const mulan_desc* d, const mulan_end_consts* kc, const uint8_t* x,
                         const float* a, const float* b, const float* c, const float* t,
                         const float* eps0, const float* eps, float* z_t, float* g_net,
                         float* w_save, float* loss_recon, float* loss_klz_prior, float* var_sums,
                         void* stream) {
  const char* fn = "mulan_fwd_pre_consts";
  if (kc == nullptr)
    return mulan_fwd_pre(d, x, a, b, c, t, eps0, eps, z_t, g_net, w_save, loss_recon,
                         loss_klz_prior, var_sums, stream);
  if (int r = check_desc(d, fn)) return r;
  if (d->n_timesteps > 0)
    return fail(MULAN_ERR_UNSUPPORTED, "%s: use mulan_fwd_pre for sm_n_timesteps > 0", fn);
  if (!(kc->exp_half_g0 > 0.f) || !(kc->exp_neg_half_g0 > 0.f) || !(kc->sigmoid_g0 > 0.f) ||
      !(kc->sigmoid_g1 > 0.f) || !(kc->sigmoid_g1 <= 1.f) || !(kc->log_sigmoid_g1 <= 0.f))
    return fail(MULAN_ERR_INVALID_ARG, "%s: end constants out of range", fn);
  if (d->rows == 0) return 0;
  REQ_X(x, fn);
  REQ_VEC(a, fn); REQ_VEC(b, fn); REQ_VEC(c, fn); REQ_PTR(t, fn);
  REQ_VEC(eps0, fn); REQ_VEC(eps, fn); REQ_VEC(z_t, fn);
  REQ_PTR(g_net, fn); OPT_VEC(w_save, fn);
  if (d->gt_mode == MULAN_GT_PIXEL) REQ_VEC(g_net, fn);
  REQ_PTR(loss_recon, fn); REQ_PTR(loss_klz_prior, fn); REQ_PTR(var_sums, fn);
  mulan::FwdPreParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.a = a; p.b = b; p.c = c; p.t = t; p.eps0 = eps0; p.eps = eps;
  p.z_t = z_t; p.g_net = g_net; p.w_save = w_save;
  p.loss_recon = loss_recon; p.loss_klz = loss_klz_prior; p.var_sums = var_sums;
  p.rows = d->rows; p.dim4 = d->dim / 4; p.gt_mode = d->gt_mode;
  p.c_raw = flag(d, MULAN_FLAG_C_RAW); p.pdl = flag(d, MULAN_FLAG_PDL);
  p.noise_rows = d->noise_rows;
  fill_pre_consts(d, kc, &p);
  cudaError_t e = mulan::launch_fwd_pre(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_fwd_pre_keyed(const mulan_desc* d, const uint32_t* key_eps0, const uint32_t* key_eps,
                        const uint8_t* x, const float* a, const float* b, const float* c,
                        const float* t, float* z_t, float* g_net, float* w_save, float* eps0_out,
                        float* eps_out, float* loss_recon, float* loss_klz_prior, float* var_sums,
                        void* stream) {
  const char* fn = "mulan_fwd_pre_keyed";
  if (int r = check_desc(d, fn)) return r;
  if (d->n_timesteps > 0)
    return fail(MULAN_ERR_UNSUPPORTED, "%s: use mulan_fwd_pre for sm_n_timesteps > 0", fn);
  if (d->rows % 2 != 0)
    return fail(MULAN_ERR_UNSUPPORTED, "%s: rows=%d must be even (JAX pairs element e with e + "
                "N/2: a CTA draws for the row pair (r, r + rows/2))", fn, d->rows);
  if (d->noise_rows != 0)
    return fail(MULAN_ERR_UNSUPPORTED, "%s: noise_rows does not apply to in-kernel draws", fn);
  if ((int64_t)d->rows * d->dim >= 0xFFFFFFFFLL)
    return fail(MULAN_ERR_INVALID_ARG, "%s: rows * dim must be below 2^32 - 1", fn);
  if (d->rows == 0) return 0;
  REQ_PTR(key_eps0, fn); REQ_PTR(key_eps, fn);
  REQ_X(x, fn);
  REQ_VEC(a, fn); REQ_VEC(b, fn); REQ_VEC(c, fn); REQ_PTR(t, fn); REQ_VEC(z_t, fn);
  REQ_PTR(g_net, fn); OPT_VEC(w_save, fn); OPT_VEC(eps0_out, fn); OPT_VEC(eps_out, fn);
  if (d->gt_mode == MULAN_GT_PIXEL) REQ_VEC(g_net, fn);
  REQ_PTR(loss_recon, fn); REQ_PTR(loss_klz_prior, fn); REQ_PTR(var_sums, fn);
  mulan::FwdPreKeyedParams q;
  memset(&q, 0, sizeof(q));
  mulan::FwdPreParams& p = q.p;
  p.x = x; p.a = a; p.b = b; p.c = c; p.t = t;
  p.z_t = z_t; p.g_net = g_net; p.w_save = w_save;
  p.loss_recon = loss_recon; p.loss_klz = loss_klz_prior; p.var_sums = var_sums;
  p.rows = d->rows; p.dim4 = d->dim / 4; p.gt_mode = d->gt_mode;
  p.c_raw = flag(d, MULAN_FLAG_C_RAW); p.pdl = flag(d, MULAN_FLAG_PDL);
  fill_pre_consts(d, nullptr, &p);
  if (!(p.W == 1 && p.vi.pow2 && p.k.v1_uniform))
    return fail(MULAN_ERR_UNSUPPORTED, "%s: only the closed-form reconstruction term (power-of-"
                "two vocab, 3-bin window at gamma_min) with a uniform prior end", fn);
  q.k_eps0[0] = key_eps0[0]; q.k_eps0[1] = key_eps0[1];
  q.k_eps[0] = key_eps[0]; q.k_eps[1] = key_eps[1];
  q.eps0_out = eps0_out; q.eps_out = eps_out;
  cudaError_t e = mulan::launch_fwd_pre_keyed(q, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_fwd_pre_variant(const mulan_desc* d) {
  const char* fn = "mulan_fwd_pre_variant";
  if (int r = check_desc(d, fn)) return r;
  mulan::FwdPreParams p;
  memset(&p, 0, sizeof(p));
  p.W = recon_window(d);
  p.gmin = f32_gmin(d); p.delta = f32_delta(d);
  p.k = mulan::make_end_consts(p.gmin, p.delta);
  p.vi = mulan::make_vocab(d->vocab);
  p.rc = mulan::make_recon_fast(p.k, p.vi);
  return mulan::fwd_pre_variant(p);
}

static int fill_post(const char* fn, const mulan_desc* d, const uint8_t* x, const float* a,
                     const float* b, const float* c, const float* t, const float* eps,
                     const float* net, const float* w_save, mulan::PostParams* p) {
  REQ_VEC(eps, fn); REQ_VEC(net, fn); OPT_VEC(w_save, fn);
  const int kparam = mulan::kernel_param(d->param);   // v-from-eps runs the epsilon form
  const bool need_poly = !(kparam == MULAN_PARAM_EPS && w_save != nullptr);
  if (need_poly) { REQ_VEC(a, fn); REQ_VEC(b, fn); REQ_VEC(c, fn); REQ_PTR(t, fn); }
  if (kparam != MULAN_PARAM_EPS) REQ_X(x, fn);
  p->x = x; p->a = a; p->b = b; p->c = c; p->t = t; p->eps = eps; p->net = net;
  p->w_save = (kparam == MULAN_PARAM_EPS) ? w_save : nullptr;
  p->gL = nullptr; p->loss_diff = nullptr; p->n_bar = nullptr;
  p->rows = d->rows; p->dim4 = d->dim / 4; p->param = kparam;
  p->c_raw = flag(d, MULAN_FLAG_C_RAW); p->pdl = flag(d, MULAN_FLAG_PDL);
  p->noise_rows = d->noise_rows;
  memset(&p->red, 0, sizeof(p->red));
  p->gmin = f32_gmin(d); p->delta = f32_delta(d);
  // .5 * sum(...)  |  .5 * T * sum(...)   (ldm/model_mulan_epsilon.py:345, :353)
  p->scale = d->n_timesteps > 0 ? (float)(0.5 * (double)d->n_timesteps) : 0.5f;
  if (d->n_timesteps > 0 && w_save == nullptr)
    return fail(MULAN_ERR_INVALID_ARG, "%s: sm_n_timesteps > 0 needs the w_save written by "
                "mulan_fwd_pre", fn);
  p->vi = mulan::make_vocab(d->vocab);
  return 0;
}

int mulan_fwd_post(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                   const float* c, const float* t, const float* eps, const float* net,
                   const float* w_save, float* loss_diff, void* stream) {
  const char* fn = "mulan_fwd_post";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return 0;
  mulan::PostParams p;
  if (int r = fill_post(fn, d, x, a, b, c, t, eps, net, w_save, &p)) return r;
  REQ_PTR(loss_diff, fn);
  p.loss_diff = loss_diff;
  cudaError_t e = mulan::launch_fwd_post(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_bwd_post(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                   const float* c, const float* t, const float* eps, const float* net,
                   const float* w_save, const float* gL, float* n_bar, void* stream) {
  const char* fn = "mulan_bwd_post";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return 0;
  mulan::PostParams p;
  if (int r = fill_post(fn, d, x, a, b, c, t, eps, net, w_save, &p)) return r;
  REQ_PTR(gL, fn); REQ_VEC(n_bar, fn);
  p.gL = gL; p.n_bar = n_bar;
  cudaError_t e = mulan::launch_bwd_post(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_fwd_bwd_post(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                       const float* c, const float* t, const float* eps, const float* net,
                       const float* w_save, const float* gL, float* loss_diff, float* n_bar,
                       void* stream) {
  const char* fn = "mulan_fwd_bwd_post";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return 0;
  mulan::PostParams p;
  if (int r = fill_post(fn, d, x, a, b, c, t, eps, net, w_save, &p)) return r;
  REQ_PTR(gL, fn); REQ_PTR(loss_diff, fn); REQ_VEC(n_bar, fn);
  p.gL = gL; p.loss_diff = loss_diff; p.n_bar = n_bar;
  cudaError_t e = mulan::launch_fwd_bwd_post(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_scale_rows(int32_t rows, int32_t dim, float* v, const float* num, const float* den,
                     void* stream) {
  const char* fn = "mulan_scale_rows";
  if (rows < 0 || dim <= 0) return fail(MULAN_ERR_INVALID_ARG, "%s: bad shape", fn);
  if (dim % 4 != 0) return fail(MULAN_ERR_ALIGNMENT, "%s: dim=%d is not a multiple of 4", fn, dim);
  if (rows == 0) return 0;
  REQ_VEC(v, fn); REQ_PTR(num, fn); REQ_PTR(den, fn);
  cudaError_t e = mulan::launch_scale_rows(v, num, den, rows, dim / 4, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_bwd_pre(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                  const float* c, const float* t, const float* eps, const float* net,
                  const float* z_bar, const float* g_bar, const float* gL,
                  float* a_bar, float* b_bar, float* c_bar, void* stream) {
  const char* fn = "mulan_bwd_pre";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return 0;
  REQ_VEC(a, fn); REQ_VEC(b, fn); REQ_VEC(c, fn); REQ_PTR(t, fn);
  REQ_VEC(a_bar, fn); REQ_VEC(b_bar, fn); REQ_VEC(c_bar, fn);
  OPT_VEC(z_bar, fn);
  if (d->gt_mode == MULAN_GT_PIXEL) OPT_VEC(g_bar, fn);
  if (gL != nullptr) { REQ_VEC(net, fn); }
  if (gL != nullptr || z_bar != nullptr) { REQ_VEC(eps, fn); }
  const int kparam = mulan::kernel_param(d->param);   // v-from-eps runs the epsilon form
  if (z_bar != nullptr || (gL != nullptr && kparam != MULAN_PARAM_EPS)) REQ_X(x, fn);
  mulan::BwdPreParams p;
  p.x = x; p.a = a; p.b = b; p.c = c; p.t = t; p.eps = eps; p.net = net;
  p.z_bar = z_bar; p.g_bar = g_bar; p.gL = gL;
  p.a_bar = a_bar; p.b_bar = b_bar; p.c_bar = c_bar;
  p.rows = d->rows; p.dim4 = d->dim / 4; p.param = kparam; p.gt_mode = d->gt_mode;
  p.c_raw = flag(d, MULAN_FLAG_C_RAW); p.pdl = flag(d, MULAN_FLAG_PDL);
  p.noise_rows = d->noise_rows;
  p.gmin = f32_gmin(d); p.delta = f32_delta(d);
  p.T = d->n_timesteps;
  p.inv_T = d->n_timesteps > 0 ? (float)(1.0 / (double)d->n_timesteps) : 0.f;
  p.vi = mulan::make_vocab(d->vocab);
  cudaError_t e = mulan::launch_bwd_pre(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_aux_topk_fwd(int32_t rows, int32_t latent, int32_t k, const float* logits,
                       const float* gamma_draw, float* embedding, float* kl_z, void* stream) {
  const char* fn = "mulan_aux_topk_fwd";
  if (rows < 0) return fail(MULAN_ERR_INVALID_ARG, "%s: rows=%d < 0", fn, rows);
  if (latent < 1 || latent > 64)
    return fail(MULAN_ERR_INVALID_ARG, "%s: latent=%d outside [1,64]", fn, latent);
  if (k < 1 || k > latent)
    return fail(MULAN_ERR_INVALID_ARG, "%s: k=%d outside [1,latent=%d]", fn, k, latent);
  if (rows == 0) return 0;
  REQ_PTR(logits, fn); REQ_PTR(embedding, fn); REQ_PTR(kl_z, fn);
  cudaError_t e = mulan::launch_aux_topk_fwd(rows, latent, k, 0, logits, gamma_draw, embedding,
                                             kl_z, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_aux_topk_bwd(int32_t rows, int32_t latent, int32_t k, const float* logits,
                       const float* gamma_draw, const float* emb_bar, const float* klz_bar,
                       float* logits_bar, void* stream) {
  const char* fn = "mulan_aux_topk_bwd";
  if (rows < 0) return fail(MULAN_ERR_INVALID_ARG, "%s: rows=%d < 0", fn, rows);
  if (latent < 1 || latent > 64)
    return fail(MULAN_ERR_INVALID_ARG, "%s: latent=%d outside [1,64]", fn, latent);
  if (k < 1 || k > latent)
    return fail(MULAN_ERR_INVALID_ARG, "%s: k=%d outside [1,latent=%d]", fn, k, latent);
  if (rows == 0) return 0;
  REQ_PTR(logits, fn); REQ_PTR(logits_bar, fn);
  cudaError_t e = mulan::launch_aux_topk_bwd(rows, latent, k, 0, logits, gamma_draw, emb_bar,
                                             klz_bar, logits_bar, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

static int check_aux(const char* fn, int32_t rows, int32_t latent) {
  if (rows < 0) return fail(MULAN_ERR_INVALID_ARG, "%s: rows=%d < 0", fn, rows);
  if (latent < 1 || latent > 64)
    return fail(MULAN_ERR_INVALID_ARG, "%s: latent=%d outside [1,64]", fn, latent);
  return 0;
}

int mulan_aux_topk_add_fwd(int32_t rows, int32_t latent, int32_t k, const float* logits,
                           const float* noise, float* embedding, float* kl_z, void* stream) {
  const char* fn = "mulan_aux_topk_add_fwd";
  if (int r = check_aux(fn, rows, latent)) return r;
  if (k < 1 || k > latent)
    return fail(MULAN_ERR_INVALID_ARG, "%s: k=%d outside [1,latent=%d]", fn, k, latent);
  if (rows == 0) return 0;
  REQ_PTR(logits, fn); REQ_PTR(embedding, fn); REQ_PTR(kl_z, fn);
  cudaError_t e = mulan::launch_aux_topk_fwd(rows, latent, k, 1, logits, noise, embedding, kl_z,
                                             (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_aux_topk_add_bwd(int32_t rows, int32_t latent, int32_t k, const float* logits,
                           const float* noise, const float* emb_bar, const float* klz_bar,
                           float* logits_bar, void* stream) {
  const char* fn = "mulan_aux_topk_add_bwd";
  if (int r = check_aux(fn, rows, latent)) return r;
  if (k < 1 || k > latent)
    return fail(MULAN_ERR_INVALID_ARG, "%s: k=%d outside [1,latent=%d]", fn, k, latent);
  if (rows == 0) return 0;
  REQ_PTR(logits, fn); REQ_PTR(logits_bar, fn);
  cudaError_t e = mulan::launch_aux_topk_bwd(rows, latent, k, 1, logits, noise, emb_bar, klz_bar,
                                             logits_bar, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_aux_gumbel_fwd(int32_t rows, int32_t latent, double tau, const float* logits,
                         const float* gumbel_noise, float* embedding, float* kl_z, void* stream) {
  const char* fn = "mulan_aux_gumbel_fwd";
  if (int r = check_aux(fn, rows, latent)) return r;
  if (!(tau > 0)) return fail(MULAN_ERR_INVALID_ARG, "%s: tau must be positive", fn);
  if (rows == 0) return 0;
  REQ_PTR(logits, fn); REQ_PTR(embedding, fn); REQ_PTR(kl_z, fn);
  cudaError_t e = mulan::launch_aux_gumbel(false, rows, latent, (float)tau, logits, gumbel_noise,
                                           nullptr, nullptr, embedding, kl_z, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_aux_gumbel_bwd(int32_t rows, int32_t latent, double tau, const float* logits,
                         const float* gumbel_noise, const float* emb_bar, const float* klz_bar,
                         float* logits_bar, void* stream) {
  const char* fn = "mulan_aux_gumbel_bwd";
  if (int r = check_aux(fn, rows, latent)) return r;
  if (!(tau > 0)) return fail(MULAN_ERR_INVALID_ARG, "%s: tau must be positive", fn);
  if (rows == 0) return 0;
  REQ_PTR(logits, fn); REQ_PTR(logits_bar, fn);
  cudaError_t e = mulan::launch_aux_gumbel(true, rows, latent, (float)tau, logits, gumbel_noise,
                                           emb_bar, klz_bar, logits_bar, nullptr,
                                           (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_aux_gaussian_fwd(int32_t rows, int32_t latent, const float* mu, const float* var,
                           const float* eps_z, float* embedding, float* kl_z, void* stream) {
  const char* fn = "mulan_aux_gaussian_fwd";
  if (int r = check_aux(fn, rows, latent)) return r;
  if (rows == 0) return 0;
  REQ_PTR(mu, fn); REQ_PTR(var, fn); REQ_PTR(eps_z, fn); REQ_PTR(embedding, fn); REQ_PTR(kl_z, fn);
  cudaError_t e = mulan::launch_aux_gaussian(false, rows, latent, mu, var, eps_z, nullptr, nullptr,
                                             embedding, kl_z, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_aux_gaussian_bwd(int32_t rows, int32_t latent, const float* mu, const float* var,
                           const float* eps_z, const float* emb_bar, const float* klz_bar,
                           float* mu_bar, float* var_bar, void* stream) {
  const char* fn = "mulan_aux_gaussian_bwd";
  if (int r = check_aux(fn, rows, latent)) return r;
  if (rows == 0) return 0;
  REQ_PTR(mu, fn); REQ_PTR(var, fn); REQ_PTR(eps_z, fn); REQ_PTR(mu_bar, fn); REQ_PTR(var_bar, fn);
  cudaError_t e = mulan::launch_aux_gaussian(true, rows, latent, mu, var, eps_z, emb_bar, klz_bar,
                                             mu_bar, var_bar, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

static int fill_reduce(const char* fn, const mulan_desc* d, const float* loss_recon,
                       const float* loss_klz_prior, const float* kl_z, const float* loss_diff,
                       const float* var_sums, float* scalars, float* loss_klz_total,
                       void* reduce_ws, mulan::BpdReduceParams* r) {
  REQ_PTR(loss_recon, fn); REQ_PTR(loss_klz_prior, fn); REQ_PTR(var_sums, fn); REQ_PTR(scalars, fn);
  OPT_VEC(reduce_ws, fn);
  r->loss_recon = loss_recon; r->loss_klz_prior = loss_klz_prior; r->kl_z = kl_z;
  r->loss_diff = loss_diff; r->var_sums = var_sums; r->scalars = scalars;
  r->loss_klz_total = loss_klz_total; r->ws = static_cast<unsigned*>(reduce_ws);
  r->rows = d->rows; r->dim = d->dim;
  memset(&r->board, 0, sizeof(r->board));
  return 0;
}

static int fill_board(const char* fn, const mulan_scalar_board* b, mulan::ScalarBoard* out) {
  memset(out, 0, sizeof(*out));
  if (b == nullptr) return 0;
  if (b->world < 1 || b->world > 8 || b->rank < 0 || b->rank >= b->world)
    return fail(MULAN_ERR_INVALID_ARG, "%s: scalar board world=%d rank=%d", fn, b->world, b->rank);
  for (int r = 0; r < b->world; ++r) {
    if (b->boards[r] == nullptr || !aligned(b->boards[r], 16))
      return fail(MULAN_ERR_INVALID_ARG, "%s: scalar board of rank %d is NULL / misaligned", fn, r);
    out->boards[r] = b->boards[r];
  }
  out->world = b->world; out->rank = b->rank;
  return 0;
}

size_t mulan_reduce_ws_bytes(int32_t rows) { return mulan::reduce_ws_bytes(rows); }

int mulan_bpd_reduce(const mulan_desc* d, const float* loss_recon, const float* loss_klz_prior,
                     const float* kl_z, const float* loss_diff, const float* var_sums,
                     float* scalars, float* loss_klz_total, void* reduce_ws, void* stream) {
  const char* fn = "mulan_bpd_reduce";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return fail(MULAN_ERR_INVALID_ARG, "%s: rows=0 has no mean", fn);
  mulan::BpdReduceParams q;
  if (int r = fill_reduce(fn, d, loss_recon, loss_klz_prior, kl_z, loss_diff, var_sums, scalars,
                          loss_klz_total, reduce_ws, &q)) return r;
  cudaError_t e = mulan::launch_bpd_reduce(q, flag(d, MULAN_FLAG_PDL) != 0, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_post_bpd(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                   const float* c, const float* t, const float* eps, const float* net,
                   const float* w_save, const float* gL, const float* loss_recon,
                   const float* loss_klz_prior, const float* kl_z, const float* var_sums,
                   float* loss_diff, float* n_bar, float* scalars, float* loss_klz_total,
                   void* reduce_ws, void* stream) {
  return mulan_post_bpd_peer(d, x, a, b, c, t, eps, net, w_save, gL, loss_recon, loss_klz_prior,
                             kl_z, var_sums, loss_diff, n_bar, scalars, loss_klz_total, reduce_ws,
                             nullptr, stream);
}

size_t mulan_scalar_board_bytes(void) {
  return sizeof(float) * (size_t)(mulan::kBoardSlots * 8 * mulan::kBoardRow) + 16;
}

int mulan_scalar_board_read(const mulan_scalar_board* board, float* mean_out, uint32_t* epoch_out,
                            void* stream) {
  const char* fn = "mulan_scalar_board_read";
  REQ_PTR(board, fn); REQ_PTR(mean_out, fn);
  mulan::ScalarBoard b;
  if (int r = fill_board(fn, board, &b)) return r;
  cudaError_t e = mulan::launch_board_read(b, mean_out, epoch_out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_post_bpd_peer(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                        const float* c, const float* t, const float* eps, const float* net,
                        const float* w_save, const float* gL, const float* loss_recon,
                        const float* loss_klz_prior, const float* kl_z, const float* var_sums,
                        float* loss_diff, float* n_bar, float* scalars, float* loss_klz_total,
                        void* reduce_ws, const mulan_scalar_board* board, void* stream) {
  const char* fn = "mulan_post_bpd";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return fail(MULAN_ERR_INVALID_ARG, "%s: rows=0 has no mean", fn);
  mulan::PostParams p;
  if (int r = fill_post(fn, d, x, a, b, c, t, eps, net, w_save, &p)) return r;
  REQ_PTR(loss_diff, fn);
  REQ_PTR(reduce_ws, fn);
  if (gL != nullptr) REQ_VEC(n_bar, fn);
  p.gL = gL; p.loss_diff = loss_diff; p.n_bar = gL != nullptr ? n_bar : nullptr;
  if (int r = fill_reduce(fn, d, loss_recon, loss_klz_prior, kl_z, loss_diff, var_sums, scalars,
                          loss_klz_total, reduce_ws, &p.red)) return r;
  if (int r = fill_board(fn, board, &p.red.board)) return r;
  cudaError_t e = gL != nullptr ? mulan::launch_fwd_bwd_post(p, (cudaStream_t)stream)
                                : mulan::launch_fwd_post(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

static int fill_sampler(const char* fn, const mulan_desc* d, int32_t abc_rows, const float* a,
                        const float* b, const float* c, mulan::SamplerParams* p) {
  if (abc_rows != 1 && abc_rows != d->rows)
    return fail(MULAN_ERR_INVALID_ARG, "%s: abc_rows=%d must be 1 (broadcast) or rows=%d", fn,
                abc_rows, d->rows);
  REQ_VEC(a, fn); REQ_VEC(b, fn); REQ_VEC(c, fn);
  memset(p, 0, sizeof(*p));
  p->a = a; p->b = b; p->c = c;
  p->rows = d->rows; p->dim4 = d->dim / 4; p->abc_rows = abc_rows; p->param = d->param;
  p->gt_mode = d->gt_mode;
  p->gmin = f32_gmin(d); p->delta = f32_delta(d);
  p->vi = mulan::make_vocab(d->vocab);
  return 0;
}

int mulan_sample_gamma(const mulan_desc* d, int32_t abc_rows, const float* a, const float* b,
                       const float* c, const float* t, float* g_net, void* stream) {
  const char* fn = "mulan_sample_gamma";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return 0;
  mulan::SamplerParams p;
  if (int r = fill_sampler(fn, d, abc_rows, a, b, c, &p)) return r;
  REQ_PTR(t, fn); REQ_PTR(g_net, fn);
  if (d->gt_mode == MULAN_GT_PIXEL) REQ_VEC(g_net, fn);
  p.t = t; p.g_net = g_net;
  cudaError_t e = mulan::launch_sample_gamma(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_sample_step(const mulan_desc* d, int32_t abc_rows, const float* a, const float* b,
                      const float* c, const float* t, const float* s, const float* z_t,
                      const float* net, const float* eps, float* z_s, void* stream) {
  const char* fn = "mulan_sample_step";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return 0;
  mulan::SamplerParams p;
  if (int r = fill_sampler(fn, d, abc_rows, a, b, c, &p)) return r;
  REQ_PTR(t, fn); REQ_PTR(s, fn); REQ_VEC(z_t, fn); REQ_VEC(net, fn); REQ_VEC(eps, fn);
  REQ_VEC(z_s, fn);
  p.t = t; p.s = s; p.z_t = z_t; p.net = net; p.eps = eps; p.z_s = z_s;
  cudaError_t e = mulan::launch_sample_step(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_generate_x(const mulan_desc* d, const float* z_0, uint8_t* x, void* stream) {
  const char* fn = "mulan_generate_x";
  if (int r = check_desc(d, fn)) return r;
  if (d->rows == 0) return 0;
  REQ_VEC(z_0, fn); REQ_X(x, fn);
  if (d->vocab > 256)
    return fail(MULAN_ERR_UNSUPPORTED, "%s: vocab=%d does not fit the uint8 output", fn, d->vocab);
  mulan::SamplerParams p;
  memset(&p, 0, sizeof(p));
  p.z_t = z_0; p.x = x; p.rows = d->rows; p.dim4 = d->dim / 4;
  p.gmin = f32_gmin(d); p.delta = f32_delta(d);
  p.vi = mulan::make_vocab(d->vocab);
  // gamma(0) == gamma_min exactly (fixed end): var_0 and the decoder scale are constants
  const mulan::EndConsts k = mulan::make_end_consts(p.gmin, p.delta);
  p.den0 = (float)sqrt((double)(1.0f - k.v0));
  p.inv0 = k.inv0;
  cudaError_t e = mulan::launch_generate_x(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_ode_drift(const mulan_desc* d, int32_t abc_rows, const float* a, const float* b,
                    const float* c, const float* t, const float* x_t, const float* eps_hat,
                    const float* v, int32_t high_precision, float* drift, float* net_bar,
                    float* div_direct, void* stream) {
  const char* fn = "mulan_ode_drift";
  if (int r = check_desc(d, fn)) return r;
  if (d->param != MULAN_PARAM_EPS)
    return fail(MULAN_ERR_UNSUPPORTED, "%s: only the epsilon model has a working reverse_ode "
                "(ldm/model_mulan_velocity.py:393-420 returns nothing)", fn);
  if (d->rows == 0) return 0;
  mulan::SamplerParams p;
  if (int r = fill_sampler(fn, d, abc_rows, a, b, c, &p)) return r;
  REQ_PTR(t, fn); REQ_VEC(x_t, fn); REQ_VEC(eps_hat, fn); REQ_VEC(drift, fn); OPT_VEC(v, fn);
  if (v != nullptr) { REQ_VEC(net_bar, fn); REQ_PTR(div_direct, fn); }
  p.t = t; p.z_t = x_t; p.net = eps_hat; p.eps = v; p.z_s = drift; p.g_net = net_bar;
  p.div_direct = div_direct;
  cudaError_t e = mulan::launch_ode_drift(p, high_precision != 0, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_row_dot(int32_t rows, int32_t dim, const float* u, const float* v, const float* add,
                  float* out, void* stream) {
  const char* fn = "mulan_row_dot";
  if (rows < 0 || dim <= 0) return fail(MULAN_ERR_INVALID_ARG, "%s: bad shape", fn);
  if (dim % 4 != 0) return fail(MULAN_ERR_ALIGNMENT, "%s: dim=%d is not a multiple of 4", fn, dim);
  if (rows == 0) return 0;
  REQ_VEC(u, fn); REQ_VEC(v, fn); REQ_PTR(out, fn);
  cudaError_t e = mulan::launch_row_dot(u, v, add, out, rows, dim / 4, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

static int rk45_fill(const char* fn, mulan::Rk45Params& p, int64_t n, int32_t n_k,
                     const double* coef, double h, const double* y, const float* K,
                     int64_t k_stride) {
  if (n < 0 || n_k < 0 || n_k > 7) return fail(MULAN_ERR_INVALID_ARG, "%s: need n >= 0, 0 <= n_k <= 7", fn);
  if (n_k > 0 && (coef == nullptr || K == nullptr))
    return fail(MULAN_ERR_INVALID_ARG, "%s: coef / K is NULL with n_k=%d", fn, n_k);
  if (n_k > 1 && k_stride < n) return fail(MULAN_ERR_INVALID_ARG, "%s: k_stride < n", fn);
  if (y == nullptr) return fail(MULAN_ERR_INVALID_ARG, "%s: y is NULL", fn);
  if ((reinterpret_cast<uintptr_t>(y) & 7u) != 0)
    return fail(MULAN_ERR_ALIGNMENT, "%s: y is not 8-byte aligned", fn);
  p = mulan::Rk45Params{};
  p.n = n; p.k_stride = k_stride; p.n_k = n_k; p.h = h; p.y = y; p.K = K;
  for (int j = 0; j < n_k; ++j) p.coef[j] = coef[j];
  return 0;
}

int mulan_rk45_stage(int64_t n, int32_t n_k, const double* coef, double h, const double* y,
                     const float* K, int64_t k_stride, float* y_stage, double* y_out,
                     void* stream) {
  const char* fn = "mulan_rk45_stage";
  mulan::Rk45Params p;
  if (int rc = rk45_fill(fn, p, n, n_k, coef, h, y, K, k_stride)) return rc;
  if (y_stage == nullptr && y_out == nullptr)
    return fail(MULAN_ERR_INVALID_ARG, "%s: both outputs are NULL", fn);
  if (n == 0) return 0;
  p.y_stage = y_stage; p.y_out = y_out;
  cudaError_t e = mulan::launch_rk45_stage(p, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_rk45_norm(int64_t n, int32_t n_k, const double* coef, double h, double rtol,
                    double atol, const double* y, const double* y_new, const float* K,
                    int64_t k_stride, int32_t of_y, double* scratch, double* out, void* stream) {
  const char* fn = "mulan_rk45_norm";
  mulan::Rk45Params p;
  if (int rc = rk45_fill(fn, p, n, n_k, coef, h, y, K, k_stride)) return rc;
  if (n == 0) return fail(MULAN_ERR_INVALID_ARG, "%s: n = 0", fn);
  if (!of_y && n_k == 0) return fail(MULAN_ERR_INVALID_ARG, "%s: nothing to measure", fn);
  if (!(rtol >= 0.0) || !(atol >= 0.0) || (rtol == 0.0 && atol == 0.0))
    return fail(MULAN_ERR_INVALID_ARG, "%s: bad tolerances", fn);
  REQ_PTR(scratch, fn); REQ_PTR(out, fn);
  p.rtol = rtol; p.atol = atol; p.y_new = y_new; p.of_y = of_y; p.scratch = scratch;
  cudaError_t e = mulan::launch_rk45_norm(p, out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_adamw_ema(const mulan_adamw_desc* d, float* params, const float* grads, float* mu,
                    float* nu, float* ema_params, void* stream) {
  const char* fn = "mulan_adamw_ema";
  if (d == nullptr) return fail(MULAN_ERR_INVALID_ARG, "%s: desc is NULL", fn);
  if (d->n < 0 || d->n_decay < 0 || d->n_decay > d->n)
    return fail(MULAN_ERR_INVALID_ARG, "%s: need 0 <= n_decay <= n", fn);
  if (d->n % 4 != 0 || d->n_decay % 4 != 0)
    return fail(MULAN_ERR_ALIGNMENT, "%s: n and n_decay must be multiples of 4", fn);
  if (d->step < 1) return fail(MULAN_ERR_INVALID_ARG, "%s: step=%d must be >= 1", fn, d->step);
  if (!(d->clip_norm >= 0.0)) return fail(MULAN_ERR_INVALID_ARG, "%s: clip_norm < 0", fn);
  if (d->clip_norm > 0.0 && d->grad_sumsq == nullptr)
    return fail(MULAN_ERR_INVALID_ARG, "%s: clip_norm set but grad_sumsq is NULL", fn);
  if (d->n == 0) return 0;
  REQ_VEC(params, fn); REQ_VEC(grads, fn); REQ_VEC(mu, fn); REQ_VEC(nu, fn); REQ_VEC(ema_params, fn);
  cudaError_t e = mulan::launch_adamw_ema(*d, params, grads, mu, nu, ema_params,
                                          (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

static int rng_draw(const char* fn, int kind, uint32_t key0, uint32_t key1, int64_t n,
                    float minval, float maxval, void* out, void* stream) {
  if (n < 0 || n >= 0xFFFFFFFFLL)
    return fail(MULAN_ERR_INVALID_ARG, "%s: n must be in [0, 2^32 - 1)", fn);
  if (n == 0) return 0;
  if (out == nullptr) return fail(MULAN_ERR_INVALID_ARG, "%s: out is NULL", fn);
  if (!aligned(out, 4)) return fail(MULAN_ERR_ALIGNMENT, "%s: out is not 4-byte aligned", fn);
  if (kind == 1 && !(maxval >= minval))
    return fail(MULAN_ERR_INVALID_ARG, "%s: need maxval >= minval", fn);
  cudaError_t e = mulan::launch_rng_draw(kind, key0, key1, n, minval, maxval, out,
                                         (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

int mulan_rng_bits(uint32_t key0, uint32_t key1, int64_t n, uint32_t* out, void* stream) {
  return rng_draw("mulan_rng_bits", 0, key0, key1, n, 0.f, 1.f, out, stream);
}

int mulan_rng_uniform(uint32_t key0, uint32_t key1, int64_t n, float minval, float maxval,
                      float* out, void* stream) {
  return rng_draw("mulan_rng_uniform", 1, key0, key1, n, minval, maxval, out, stream);
}

int mulan_rng_normal(uint32_t key0, uint32_t key1, int64_t n, float* out, void* stream) {
  // uniform on [nextafter(-1, 0), 1): jax._src.random._normal_real
  return rng_draw("mulan_rng_normal", 2, key0, key1, n, nextafterf(-1.0f, 0.0f), 1.0f, out,
                  stream);
}

int mulan_grad_sumsq(int64_t n, const float* g, double* scratch, float* out, void* stream) {
  const char* fn = "mulan_grad_sumsq";
  if (n < 0) return fail(MULAN_ERR_INVALID_ARG, "%s: n < 0", fn);
  if (n % 4 != 0) return fail(MULAN_ERR_ALIGNMENT, "%s: n must be a multiple of 4", fn);
  REQ_PTR(scratch, fn); REQ_PTR(out, fn);
  if (n > 0) REQ_VEC(g, fn);
  cudaError_t e = mulan::launch_grad_sumsq(g, n, scratch, out, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : cuda_fail(fn, e);
}

}  // extern "C"
