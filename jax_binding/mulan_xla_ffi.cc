// XLA-FFI handlers over the C ABI of libmulan_b200.so (include/mulan_b200.h).
//
// STATUS: written against the public XLA FFI API (xla/ffi/api/ffi.h, jax >= 0.4.31:
// jax.ffi.include_dir()).  NOT compiled or run in this repository's image: JAX and its headers
// are not installable there (no wheels, no network).  The C ABI underneath is what the test
// suite exercises, through ctypes.
//
// Build (where JAX exists):
//   g++ -O2 -shared -fPIC -std=c++17 -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I../include mulan_xla_ffi.cc -L../mulan_b200 -lmulan_b200 -o libmulan_xla_ffi.so
//
// Threading: XLA calls a handler from the executor thread of the device that owns the
// buffers, with that device current; inputs are immutable, outputs pre-allocated -- exactly the
// contract of the C ABI (enqueue-only on the given stream, no allocation, no global state).
#include <cstdint>

#include "mulan_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Check(int status) {
  if (status == 0) return ffi::Error::Success();
  return ffi::Error(status == MULAN_ERR_CUDA ? ffi::ErrorCode::kInternal
                                             : ffi::ErrorCode::kInvalidArgument,
                    mulan_last_error());
}

// `flags`: mulan_flags of the C ABI (MULAN_FLAG_C_RAW: the `c` operand is the pre-activation of
// dense_out_c, ldm/model_mulan_epsilon.py:537; MULAN_FLAG_PDL).  noise_rows stays 0 here: XLA
// hands over full [B, D] draws.
static mulan_desc MakeDesc(int64_t rows, int64_t dim, int32_t vocab, int32_t param,
                           int32_t gt_mode, int32_t n_timesteps, double gmin, double gmax,
                           int32_t flags = 0) {
  mulan_desc d;
  d.rows = (int32_t)rows; d.dim = (int32_t)dim; d.vocab = vocab; d.param = param;
  d.gt_mode = gt_mode; d.n_timesteps = n_timesteps; d.gamma_min = gmin; d.gamma_max = gmax;
  d.flags = (uint32_t)flags; d.noise_rows = 0;
  return d;
}

// ---- fwd_pre: replaces ldm/model_mulan_epsilon.py:300-328 (+ :339-343, :273-278) -------------
static ffi::Error FwdPreImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> x, ffi::Buffer<ffi::F32> a,
                             ffi::Buffer<ffi::F32> b, ffi::Buffer<ffi::F32> c,
                             ffi::Buffer<ffi::F32> t, ffi::Buffer<ffi::F32> eps0,
                             ffi::Buffer<ffi::F32> eps, ffi::ResultBuffer<ffi::F32> z_t,
                             ffi::ResultBuffer<ffi::F32> g_net, ffi::ResultBuffer<ffi::F32> w,
                             ffi::ResultBuffer<ffi::F32> loss_recon,
                             ffi::ResultBuffer<ffi::F32> loss_klz,
                             ffi::ResultBuffer<ffi::F32> var_sums, int32_t vocab, int32_t param,
                             int32_t gt_mode, int32_t n_timesteps, int32_t flags,
                             double gamma_min, double gamma_max) {
  auto dims = a.dimensions();
  mulan_desc d = MakeDesc(dims[0], dims[1], vocab, param, gt_mode, n_timesteps, gamma_min, gamma_max,
                          flags);
  return Check(mulan_fwd_pre(&d, x.typed_data(), a.typed_data(), b.typed_data(), c.typed_data(),
                             t.typed_data(), eps0.typed_data(), eps.typed_data(),
                             z_t->typed_data(), g_net->typed_data(), w->typed_data(),
                             loss_recon->typed_data(), loss_klz->typed_data(),
                             var_sums->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MulanFwdPre, FwdPreImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
        .Attr<int32_t>("vocab").Attr<int32_t>("param").Attr<int32_t>("gt_mode")
        .Attr<int32_t>("n_timesteps").Attr<int32_t>("flags").Attr<double>("gamma_min")
        .Attr<double>("gamma_max"));

// ---- fwd_post / bwd_post: ldm/model_mulan_epsilon.py:338-355, velocity.py:246-260 ------------
static ffi::Error FwdPostImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> x, ffi::Buffer<ffi::F32> a,
                              ffi::Buffer<ffi::F32> b, ffi::Buffer<ffi::F32> c,
                              ffi::Buffer<ffi::F32> t, ffi::Buffer<ffi::F32> eps,
                              ffi::Buffer<ffi::F32> net, ffi::Buffer<ffi::F32> w,
                              ffi::ResultBuffer<ffi::F32> loss_diff, int32_t vocab, int32_t param,
                              int32_t gt_mode, int32_t n_timesteps, int32_t flags,
                              double gamma_min, double gamma_max) {
  auto dims = a.dimensions();
  // n_timesteps > 0: the loss is .5 * T * sum(w ...) with the discrete weight fwd_pre saved
  // (ldm/model_mulan_epsilon.py:348-355)
  mulan_desc d = MakeDesc(dims[0], dims[1], vocab, param, gt_mode, n_timesteps, gamma_min, gamma_max,
                          flags);
  return Check(mulan_fwd_post(&d, x.typed_data(), a.typed_data(), b.typed_data(), c.typed_data(),
                              t.typed_data(), eps.typed_data(), net.typed_data(),
                              mulan_kernel_param(param) == MULAN_PARAM_EPS ? w.typed_data() : nullptr,
                              loss_diff->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MulanFwdPost, FwdPostImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>()
        .Attr<int32_t>("vocab").Attr<int32_t>("param").Attr<int32_t>("gt_mode")
        .Attr<int32_t>("n_timesteps").Attr<int32_t>("flags").Attr<double>("gamma_min")
        .Attr<double>("gamma_max"));

static ffi::Error BwdPostImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> x, ffi::Buffer<ffi::F32> a,
                              ffi::Buffer<ffi::F32> b, ffi::Buffer<ffi::F32> c,
                              ffi::Buffer<ffi::F32> t, ffi::Buffer<ffi::F32> eps,
                              ffi::Buffer<ffi::F32> net, ffi::Buffer<ffi::F32> w,
                              ffi::Buffer<ffi::F32> gL, ffi::ResultBuffer<ffi::F32> n_bar,
                              int32_t vocab, int32_t param, int32_t gt_mode, int32_t n_timesteps,
                              int32_t flags, double gamma_min, double gamma_max) {
  auto dims = a.dimensions();
  mulan_desc d = MakeDesc(dims[0], dims[1], vocab, param, gt_mode, n_timesteps, gamma_min, gamma_max,
                          flags);
  return Check(mulan_bwd_post(&d, x.typed_data(), a.typed_data(), b.typed_data(), c.typed_data(),
                              t.typed_data(), eps.typed_data(), net.typed_data(),
                              mulan_kernel_param(param) == MULAN_PARAM_EPS ? w.typed_data() : nullptr,
                              gL.typed_data(), n_bar->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MulanBwdPost, BwdPostImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>()
        .Attr<int32_t>("vocab").Attr<int32_t>("param").Attr<int32_t>("gt_mode")
        .Attr<int32_t>("n_timesteps").Attr<int32_t>("flags").Attr<double>("gamma_min")
        .Attr<double>("gamma_max"));

// ---- bwd_pre: cotangents of (a, b, c) ----------------------------------------------------------
static ffi::Error BwdPreImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> x, ffi::Buffer<ffi::F32> a,
                             ffi::Buffer<ffi::F32> b, ffi::Buffer<ffi::F32> c,
                             ffi::Buffer<ffi::F32> t, ffi::Buffer<ffi::F32> eps,
                             ffi::Buffer<ffi::F32> net, ffi::Buffer<ffi::F32> z_bar,
                             ffi::Buffer<ffi::F32> g_bar, ffi::Buffer<ffi::F32> gL,
                             ffi::ResultBuffer<ffi::F32> a_bar, ffi::ResultBuffer<ffi::F32> b_bar,
                             ffi::ResultBuffer<ffi::F32> c_bar, int32_t vocab, int32_t param,
                             int32_t gt_mode, int32_t n_timesteps, int32_t flags,
                             double gamma_min, double gamma_max) {
  auto dims = a.dimensions();
  // n_timesteps > 0 selects the discrete-time branch of the backward (expm1 weight); with
  // MULAN_FLAG_C_RAW c_bar is the cotangent of the pre-activation
  mulan_desc d = MakeDesc(dims[0], dims[1], vocab, param, gt_mode, n_timesteps, gamma_min, gamma_max,
                          flags);
  return Check(mulan_bwd_pre(&d, x.typed_data(), a.typed_data(), b.typed_data(), c.typed_data(),
                             t.typed_data(), eps.typed_data(), net.typed_data(),
                             z_bar.typed_data(), g_bar.typed_data(), gL.typed_data(),
                             a_bar->typed_data(), b_bar->typed_data(), c_bar->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(
    MulanBwdPre, BwdPreImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
        .Arg<ffi::Buffer<ffi::F32>>()
        .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
        .Attr<int32_t>("vocab").Attr<int32_t>("param").Attr<int32_t>("gt_mode")
        .Attr<int32_t>("n_timesteps").Attr<int32_t>("flags").Attr<double>("gamma_min")
        .Attr<double>("gamma_max"));
