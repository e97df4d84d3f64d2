// mulan_elbo_host: the HOST-buffer entry of the C ABI (include/mulan_b200.h).  One call =
// H2D of the inputs, fwd_pre, denoiser callback (or the supplied net), fwd_post, bwd_post,
// bwd_pre, bpd_reduce, D2H of losses / scalars / gradients.
//
// The batch is cut into row chunks that flow through three streams -- copy-in, compute,
// copy-out -- so the H2D of chunk i+1, the kernels of chunk i and the D2H of chunk i-1
// overlap (PCIe is full duplex; the kernels are ~2 % of the transfer time).  Rows are
// independent, so chunking changes no result: per-row outputs are bitwise those of a single
// launch, and the six scalars are reduced once over all rows at the end.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "mulan_kernels.h"

namespace {

constexpr int kMaxChunks = 64;

struct HostWs {
  int device = -1;
  size_t cap_rows = 0;
  int dim = 0;
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[kMaxChunks] = {}, ev_cmp[kMaxChunks] = {}, ev_tail = nullptr;
  uint8_t* d_x = nullptr;
  float* d_f = nullptr;      // one slab: a,b,c,eps0,eps,net,z_t,w,nbar,abar,bbar,cbar [12][B,D]
  float* d_row = nullptr;    // t, g_net, recon, klz, diff, gL, var_sums[2] -> 8*B, + 8 scalars
  float* h_gl = nullptr;     // pinned [B]: the uniform loss cotangent
  void release() {
    if (d_x) cudaFree(d_x);
    if (d_f) cudaFree(d_f);
    if (d_row) cudaFree(d_row);
    if (h_gl) cudaFreeHost(h_gl);
    for (int i = 0; i < kMaxChunks; ++i) {
      if (ev_in[i]) cudaEventDestroy(ev_in[i]);
      if (ev_cmp[i]) cudaEventDestroy(ev_cmp[i]);
    }
    if (ev_tail) cudaEventDestroy(ev_tail);
    if (s_in) cudaStreamDestroy(s_in);
    if (s_cmp) cudaStreamDestroy(s_cmp);
    if (s_out) cudaStreamDestroy(s_out);
    *this = HostWs();
  }
};
thread_local HostWs g_ws;
thread_local char g_host_err[256] = "";   // scratch; published via mulan::set_last_error

#define CU(call)                                                        \
  do {                                                                  \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess) {                                            \
      snprintf(g_host_err, sizeof(g_host_err), "mulan_elbo_host: %s", cudaGetErrorString(e_)); \
      mulan::set_last_error(g_host_err);                                \
      return (int)MULAN_ERR_CUDA;                                       \
    }                                                                   \
  } while (0)

int alloc_ws(HostWs& ws, size_t B, int dim) {
  const size_t N = B * (size_t)dim;
  CU(cudaStreamCreateWithFlags(&ws.s_in, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&ws.s_cmp, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&ws.s_out, cudaStreamNonBlocking));
  for (int i = 0; i < kMaxChunks; ++i) {
    CU(cudaEventCreateWithFlags(&ws.ev_in[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ws.ev_cmp[i], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&ws.ev_tail, cudaEventDisableTiming));
  CU(cudaMalloc(&ws.d_x, N));
  CU(cudaMalloc(&ws.d_f, 12 * N * sizeof(float)));
  CU(cudaMalloc(&ws.d_row, (8 * B + 8) * sizeof(float)));
  CU(cudaMallocHost(&ws.h_gl, B * sizeof(float)));
  return 0;
}

// The workspace is marked valid (device / cap_rows / dim) only after EVERY allocation has
// succeeded; a failure midway (e.g. cudaMalloc out of memory) releases the partial state, so the
// next call starts from scratch instead of finding NULL pointers behind a matching shape.
int ensure_ws(HostWs& ws, size_t B, int dim) {
  int dev = 0;
  CU(cudaGetDevice(&dev));
  if (ws.device == dev && ws.cap_rows >= B && ws.dim == dim) return 0;
  ws.release();
  if (int r = alloc_ws(ws, B, dim)) {
    ws.release();
    cudaGetLastError();   // clear the sticky-free error state of the failed allocation
    return r;
  }
  ws.device = dev; ws.cap_rows = B; ws.dim = dim;
  return 0;
}

// Error exit of the pipeline: copies already enqueued still read / write the CALLER's host
// buffers, so drain the three streams before handing control (and those buffers) back.
int drain(HostWs& ws, int status) {
  if (ws.s_in) cudaStreamSynchronize(ws.s_in);
  if (ws.s_cmp) cudaStreamSynchronize(ws.s_cmp);
  if (ws.s_out) cudaStreamSynchronize(ws.s_out);
  return status;
}

}  // namespace

extern "C" {

void mulan_host_workspace_release(void) { g_ws.release(); }

}  // extern "C"

// The pipeline behind both host entries.  eps0 / eps are host arrays, or NULL with key_eps0 /
// key_eps set: then the two draws are generated on the device (threefry2x32 + erfinv, exactly
// jax.random.normal(key, [B, 32, 32, 3]) -- mulan_rng.cu) on the compute stream before the first
// chunk's kernels, while the first chunk's operands are still crossing PCIe.
static int elbo_host_impl(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                          const float* c, const float* t, const float* eps0, const float* eps,
                          const uint32_t* key_eps0, const uint32_t* key_eps,
                          const float* net, mulan_denoiser_fn denoiser, void* user,
                          int32_t want_grad, float* losses, float* scalars, float* a_bar,
                          float* b_bar, float* c_bar, float* n_bar) {
  const bool keyed = key_eps0 != nullptr && key_eps != nullptr;
  auto bad = [](const char* msg) {
    snprintf(g_host_err, sizeof(g_host_err), "mulan_elbo_host: %s", msg);
    mulan::set_last_error(g_host_err);
    return (int)MULAN_ERR_INVALID_ARG;
  };
  if (d == nullptr) return bad("desc is NULL");
  if (d->rows <= 0) return bad("rows must be positive");
  if (d->dim <= 0 || d->dim % 4 != 0) return bad("dim must be a positive multiple of 4");
  if (!x || !a || !b || !c || !t || !losses || !scalars)
    return bad("a required host pointer is NULL");
  if (!keyed && (!eps0 || !eps)) return bad("eps0 / eps are NULL and no keys were given");
  if (denoiser == nullptr && net == nullptr) return bad("net is NULL and no denoiser was given");
  if (d->gt_mode == MULAN_GT_PIXEL) {
    mulan::set_last_error("mulan_elbo_host: gt_mode=PIXEL is served by the device-pointer API");
    return (int)MULAN_ERR_UNSUPPORTED;
  }
  const size_t B = (size_t)d->rows, D = (size_t)d->dim, N = B * D;
  HostWs& ws = g_ws;
  if (int r = ensure_ws(ws, B, d->dim)) return r;
  // every chunk launches with the shape of the WHOLE batch: chunking changes no bit
  struct ShapeScope {
    explicit ShapeScope(int rows) { mulan::tl_shape_rows = rows; }
    ~ShapeScope() { mulan::tl_shape_rows = 0; }
  } shape_scope((int)B);
#undef CU
#define CU(call)                                                        \
  do {                                                                  \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess) {                                            \
      snprintf(g_host_err, sizeof(g_host_err), "mulan_elbo_host: %s", cudaGetErrorString(e_)); \
      mulan::set_last_error(g_host_err);                                \
      return drain(ws, (int)MULAN_ERR_CUDA);                            \
    }                                                                   \
  } while (0)

  float* dA = ws.d_f + 0 * N; float* dB = ws.d_f + 1 * N; float* dC = ws.d_f + 2 * N;
  float* dE0 = ws.d_f + 3 * N; float* dE = ws.d_f + 4 * N; float* dN = ws.d_f + 5 * N;
  float* dZ = ws.d_f + 6 * N; float* dW = ws.d_f + 7 * N; float* dNB = ws.d_f + 8 * N;
  float* dAB = ws.d_f + 9 * N; float* dBB = ws.d_f + 10 * N; float* dCB = ws.d_f + 11 * N;
  float* dT = ws.d_row; float* dG = ws.d_row + B; float* dRec = ws.d_row + 2 * B;
  float* dKlz = ws.d_row + 3 * B; float* dDiff = ws.d_row + 4 * B; float* dGL = ws.d_row + 5 * B;
  float* dVar = ws.d_row + 6 * B; float* dSc = ws.d_row + 8 * B;
  const bool save_w = mulan::kernel_param(d->param) == MULAN_PARAM_EPS;

  // chunking: ~1024 rows (75 MB in, 50 MB out) per chunk, at most kMaxChunks chunks
  size_t chunk = 1024;
  if ((B + chunk - 1) / chunk > (size_t)kMaxChunks) chunk = (B + kMaxChunks - 1) / kMaxChunks;
  const int nchunk = (int)((B + chunk - 1) / chunk);

  // per-example inputs first (tiny): t, and the uniform loss cotangent
  CU(cudaMemcpyAsync(dT, t, B * sizeof(float), cudaMemcpyHostToDevice, ws.s_in));
  if (want_grad) {
    // d bpd / d loss_diff_b = 1 / (B * D * ln 2)   (ldm/experiment_vdm.py:62-66)
    const float g = (float)(1.0 / ((double)B * (double)D * 0.6931471805599453));
    for (size_t i = 0; i < B; ++i) ws.h_gl[i] = g;
    CU(cudaMemcpyAsync(dGL, ws.h_gl, B * sizeof(float), cudaMemcpyHostToDevice, ws.s_in));
  }

  if (keyed) {
    // whole-batch draws (JAX's counter layout pairs element i with element i + N/2, so a draw
    // cannot be generated chunk by chunk without doing every block twice)
    if (N >= 0xffffffffULL) return drain(ws, bad("keyed draws need rows * dim < 2^32 - 1"));
    // jax.random.normal: sqrt(2) erf_inv(uniform on [nextafter(-1, 0), 1)), as mulan_rng_normal
    const float lo = nextafterf(-1.0f, 0.0f);
    cudaError_t e = mulan::launch_rng_draw(2, key_eps0[0], key_eps0[1], (long long)N, lo, 1.0f,
                                           dE0, ws.s_cmp);
    if (e == cudaSuccess)
      e = mulan::launch_rng_draw(2, key_eps[0], key_eps[1], (long long)N, lo, 1.0f, dE, ws.s_cmp);
    CU(e);
  }

  for (int ci = 0; ci < nchunk; ++ci) {
    const size_t r0 = (size_t)ci * chunk, nr = (r0 + chunk <= B) ? chunk : B - r0;
    const size_t o = r0 * D, n = nr * D;
    // ---- copy-in stream: straight from the caller's buffers (true DMA when page-locked)
    CU(cudaMemcpyAsync(ws.d_x + o, x + o, n, cudaMemcpyHostToDevice, ws.s_in));
    const float* srcs[6] = {a, b, c, keyed ? nullptr : eps0, keyed ? nullptr : eps,
                            denoiser == nullptr ? net : nullptr};
    float* dsts[6] = {dA, dB, dC, dE0, dE, dN};
    for (int k = 0; k < 6; ++k)
      if (srcs[k] != nullptr)
        CU(cudaMemcpyAsync(dsts[k] + o, srcs[k] + o, n * sizeof(float), cudaMemcpyHostToDevice,
                           ws.s_in));
    CU(cudaEventRecord(ws.ev_in[ci], ws.s_in));
    // ---- compute stream
    CU(cudaStreamWaitEvent(ws.s_cmp, ws.ev_in[ci], 0));
    mulan_desc dc = *d;
    dc.rows = (int32_t)nr;
    void* sc = (void*)ws.s_cmp;
    float* w_save = save_w ? dW + o : nullptr;
    int r = mulan_fwd_pre(&dc, ws.d_x + o, dA + o, dB + o, dC + o, dT + r0, dE0 + o, dE + o, dZ + o,
                          dG + r0, w_save, dRec + r0, dKlz + r0, dVar + 2 * r0, sc);
    if (r) return drain(ws, r);   // mulan_last_error() already holds the entry point's message
    if (denoiser != nullptr) {
      if (int rc = denoiser(user, (int32_t)nr, dZ + o, dG + r0, dN + o, sc)) {
        snprintf(g_host_err, sizeof(g_host_err), "mulan_elbo_host: denoiser callback returned %d", rc);
        mulan::set_last_error(g_host_err);
        return drain(ws, (int)MULAN_ERR_INVALID_ARG);
      }
    }
    if (!want_grad) {
      r = mulan_fwd_post(&dc, ws.d_x + o, dA + o, dB + o, dC + o, dT + r0, dE + o, dN + o, w_save,
                         dDiff + r0, sc);
    } else {
      // value-and-grad in one pass: the cotangent of a mean is known up front
      r = mulan_fwd_bwd_post(&dc, ws.d_x + o, dA + o, dB + o, dC + o, dT + r0, dE + o, dN + o,
                             w_save, dGL + r0, dDiff + r0, dNB + o, sc);
      if (!r)
        r = mulan_bwd_pre(&dc, ws.d_x + o, dA + o, dB + o, dC + o, dT + r0, dE + o, dN + o, nullptr,
                          nullptr, dGL + r0, dAB + o, dBB + o, dCB + o, sc);
    }
    if (r) return drain(ws, r);   // mulan_last_error() already holds the entry point's message
    CU(cudaEventRecord(ws.ev_cmp[ci], ws.s_cmp));
    // ---- copy-out stream
    if (want_grad) {
      CU(cudaStreamWaitEvent(ws.s_out, ws.ev_cmp[ci], 0));
      float* gdst[4] = {a_bar, b_bar, c_bar, n_bar};
      float* gsrc[4] = {dAB, dBB, dCB, dNB};
      for (int k = 0; k < 4; ++k)
        if (gdst[k] != nullptr)
          CU(cudaMemcpyAsync(gdst[k] + o, gsrc[k] + o, n * sizeof(float), cudaMemcpyDeviceToHost,
                             ws.s_out));
    }
  }
  // ---- tail: the six scalars over ALL rows, then losses + scalars back
  int r = mulan_bpd_reduce(d, dRec, dKlz, nullptr, dDiff, dVar, dSc, nullptr, nullptr,
                           (void*)ws.s_cmp);
  if (r) return drain(ws, r);
  CU(cudaMemcpyAsync(losses, dRec, 3 * B * sizeof(float), cudaMemcpyDeviceToHost, ws.s_cmp));
  CU(cudaMemcpyAsync(scalars, dSc, 6 * sizeof(float), cudaMemcpyDeviceToHost, ws.s_cmp));
  CU(cudaStreamSynchronize(ws.s_cmp));
  CU(cudaStreamSynchronize(ws.s_out));
  CU(cudaStreamSynchronize(ws.s_in));
  return 0;
}

extern "C" {

int mulan_elbo_host(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                    const float* c, const float* t, const float* eps0, const float* eps,
                    const float* net, mulan_denoiser_fn denoiser, void* user, int32_t want_grad,
                    float* losses, float* scalars, float* a_bar, float* b_bar, float* c_bar,
                    float* n_bar) {
  if (eps0 == nullptr || eps == nullptr) {
    mulan::set_last_error("mulan_elbo_host: eps0 / eps is NULL");
    return (int)MULAN_ERR_INVALID_ARG;
  }
  return elbo_host_impl(d, x, a, b, c, t, eps0, eps, nullptr, nullptr, net, denoiser, user,
                        want_grad, losses, scalars, a_bar, b_bar, c_bar, n_bar);
}

int mulan_elbo_host_keyed(const mulan_desc* d, const uint8_t* x, const float* a, const float* b,
                          const float* c, const float* t, const uint32_t* key_eps0,
                          const uint32_t* key_eps, const float* net, mulan_denoiser_fn denoiser,
                          void* user, int32_t want_grad, float* losses, float* scalars,
                          float* a_bar, float* b_bar, float* c_bar, float* n_bar) {
  if (key_eps0 == nullptr || key_eps == nullptr) {
    mulan::set_last_error("mulan_elbo_host_keyed: a key is NULL");
    return (int)MULAN_ERR_INVALID_ARG;
  }
  return elbo_host_impl(d, x, a, b, c, t, nullptr, nullptr, key_eps0, key_eps, net, denoiser,
                        user, want_grad, losses, scalars, a_bar, b_bar, c_bar, n_bar);
}

}  // extern "C"
