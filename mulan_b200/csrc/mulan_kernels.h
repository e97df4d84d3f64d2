// Internal: kernel parameter blocks and launchers shared by the .cu files and the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/mulan_b200.h"
#include "mulan_common.cuh"

namespace mulan {

// Bin geometry of EncDec (ldm/model_vdm.py:274-294): centres encode(k), k = 0..vocab-1.
struct VocabInfo {
  int vocab;
  int pow2;           // vocab is a power of two: (2k+1)/vocab - 1 is exact in one fma
  float vocab_f;      // (float)vocab
  float inv_vocab;    // 1/vocab (exact when pow2)
  float half_vocab;   // vocab/2
  float vocab_m1;     // vocab-1
  __device__ __forceinline__ float xval(int k) const {
    return pow2 ? fmaf((float)(2 * k + 1), inv_vocab, -1.0f) : encode_ref(k, vocab_f);
  }
};

inline VocabInfo make_vocab(int vocab) {
  VocabInfo v;
  v.vocab = vocab;
  v.pow2 = (vocab & (vocab - 1)) == 0;
  v.vocab_f = (float)vocab;
  v.inv_vocab = 1.0f / (float)vocab;
  v.half_vocab = 0.5f * (float)vocab;
  v.vocab_m1 = (float)(vocab - 1);
  return v;
}

// Constants of the fixed-end schedule: gamma(0) == f32(gamma_min) exactly, gamma(1) ==
// gmin + (delta*S)/S in {delta-ulp, delta, delta+ulp}.  Evaluated on the host in the
// reference's float32 op order with correctly rounded exp/log (double, then one rounding).
struct EndConsts {
  float s0, inv0, v0;       // exp(.5 g0), exp(-.5 g0), sigmoid(g0)
  float v1, om1, lv1;       // sigmoid(g1), 1 - v1, log(v1)
  int v1_uniform;           // sigmoid(gmin + r) identical for the three candidate r
};

inline float host_exp_f32(float x) { return (float)exp((double)x); }
inline float host_sigmoid_f32(float x) {
  const float e = host_exp_f32(-x);
  const float d = 1.0f + e;
  return 1.0f / d;
}
inline EndConsts make_end_consts(float gmin, float delta) {
  EndConsts k;
  k.s0 = host_exp_f32(0.5f * gmin);
  k.inv0 = host_exp_f32(-0.5f * gmin);
  k.v0 = host_sigmoid_f32(gmin);
  const float lo = nextafterf(delta, -INFINITY), hi = nextafterf(delta, INFINITY);
  const float g1a = gmin + lo, g1b = gmin + delta, g1c = gmin + hi;
  const float va = host_sigmoid_f32(g1a), vb = host_sigmoid_f32(g1b), vc = host_sigmoid_f32(g1c);
  k.v1_uniform = (va == vb) && (vb == vc);
  k.v1 = vb;
  k.om1 = 1.0f - vb;
  k.lv1 = (float)log((double)vb);
  return k;
}

// Constants of the closed-form 3-bin reconstruction term (mulan_fwd_pre.cu).
struct ReconFast {
  float inv0, s, s2, c0;   // e^{-g0/2}; s = (2/vocab) e^{-g0/2}; s log2(e); -s^2/2 log2(e)
  float two_iv, off;       // 2/vocab; 1/vocab - 1
  float half_vocab, vocab_m1;
};
inline ReconFast make_recon_fast(const EndConsts& k, const VocabInfo& vi) {
  ReconFast rc;
  const double s = (2.0 / (double)vi.vocab) * (double)k.inv0;
  rc.inv0 = k.inv0;
  rc.s = (float)s;
  rc.s2 = (float)(s * 1.4426950408889634);
  rc.c0 = (float)(-0.5 * s * s * 1.4426950408889634);
  rc.two_iv = 2.0f * vi.inv_vocab;
  rc.off = vi.inv_vocab - 1.0f;
  rc.half_vocab = vi.half_vocab;
  rc.vocab_m1 = vi.vocab_m1;
  return rc;
}

struct FwdPreParams {
  const uint8_t* x;
  const float *a, *b, *c, *t, *eps0, *eps;
  float *z_t, *g_net, *w_save, *loss_recon, *loss_klz, *var_sums;
  int rows, dim4, gt_mode;
  int c_raw;          // MULAN_FLAG_C_RAW: c holds the pre-activation of dense_out_c
  int pdl;            // MULAN_FLAG_PDL: launch with programmatic stream serialization
  int noise_rows;     // eps0 / eps are [noise_rows, D], row b reads b % noise_rows (0: [rows, D])
  int W;              // reconstruction window half-width for gamma_0 = gamma_min
  float gmin, delta;  // f32(gamma_min), f32(gamma_max - gamma_min)
  EndConsts k;
  ReconFast rc;
  VocabInfo vi;
};

// VDMOutput assembly + loss_fn scalars (ldm/model_mulan_epsilon.py:357-363,
// ldm/experiment_vdm.py:62-74), stand-alone or fused into the post kernel's epilogue.
// pmean of the six scalars without a collective (mulan_scalar_board, include/mulan_b200.h): the
// thread that writes a rank's scalars also stores them, tagged with a step counter, into slot
// [step % slots][rank] of EVERY rank's board over NVLink peer memory.
constexpr int kBoardSlots = 64, kBoardRow = 8;   // 6 scalars, 1 pad, 1 tag per (slot, rank)
struct ScalarBoard {
  float* boards[8];   // this process's mapping of rank r's board; boards[rank] is its own
  int world, rank;    // world == 0: no board
};

struct BpdReduceParams {
  const float *loss_recon, *loss_klz_prior, *kl_z, *loss_diff, *var_sums;
  float *scalars, *loss_klz_total;
  unsigned* ws;       // mulan_reduce_ws_bytes(rows) bytes, zero before first use (self-resetting)
  int rows, dim;
  ScalarBoard board;
};
cudaError_t launch_board_read(const ScalarBoard& b, float* mean_out, unsigned* epoch_out,
                              cudaStream_t s);

// mulan_fwd_pre_keyed: eps_0 / eps drawn inside the kernel from their threefry keys
struct FwdPreKeyedParams {
  FwdPreParams p;              // p.eps0 / p.eps unused
  uint32_t k_eps0[2], k_eps[2];
  float *eps0_out, *eps_out;   // optional [B, D] copies of the draws (eps feeds the post kernels)
};
cudaError_t launch_fwd_pre_keyed(const FwdPreKeyedParams& q, cudaStream_t s);

struct PostParams {
  const uint8_t* x;
  const float *a, *b, *c, *t, *eps, *net, *w_save, *gL;
  float* loss_diff;   // fwd
  float* n_bar;       // bwd
  int rows, dim4, param;
  int c_raw, pdl, noise_rows;   // as FwdPreParams (eps is the broadcast operand here)
  float gmin, delta;
  float scale;        // 0.5 (continuous) or 0.5*T (discrete)
  VocabInfo vi;
  BpdReduceParams red;   // fused loss-scalar reduction (red.ws != nullptr), see mulan_reduce.cuh
};

struct BwdPreParams {
  const uint8_t* x;
  const float *a, *b, *c, *t, *eps, *net, *z_bar, *g_bar, *gL;
  float *a_bar, *b_bar, *c_bar;
  int rows, dim4, param, gt_mode;
  int c_raw, pdl, noise_rows;   // c_raw: c_bar is the cotangent of the pre-activation
  float gmin, delta;
  int T;              // sm_n_timesteps (0 = continuous)
  float inv_T;        // f32(1/T): s = t - 1/T  (ldm/model_mulan_epsilon.py:350)
  VocabInfo vi;
};

// Discrete-time loss weight expm1(gamma(t) - gamma(t - 1/T)) (ldm/model_mulan_epsilon.py:350-354)
struct DiscreteWParams {
  const float *a, *b, *c, *t;
  float* w;
  int rows, dim4, c_raw;
  float gmin, delta, inv_T;
};

// Ancestral sampler step / decode (mulan_sampler.cu)
struct SamplerParams {
  const float *a, *b, *c, *t, *s;
  const float *z_t, *net, *eps;
  float* z_s;       // sample_step
  float* g_net;     // sample_gamma: [B] or [B,D]
  uint8_t* x;       // generate_x
  // ode_drift reuses: z_t = x_t, net = eps_hat, eps = Hutchinson noise v (or NULL),
  //                   z_s = drift out, g_net = net_bar out, div_direct = [B] out
  float* div_direct;
  int rows, dim4, abc_rows, param, gt_mode;
  float gmin, delta;
  float den0, inv0; // generate_x: sqrt(1 - sigmoid(g0)), exp(-g0/2)
  VocabInfo vi;
};
cudaError_t launch_sample_gamma(const SamplerParams& p, cudaStream_t s);
cudaError_t launch_sample_step(const SamplerParams& p, cudaStream_t s);
cudaError_t launch_generate_x(const SamplerParams& p, cudaStream_t s);
cudaError_t launch_ode_drift(const SamplerParams& p, bool high_precision, cudaStream_t s);
cudaError_t launch_row_dot(const float* u, const float* v, const float* add, float* out, int rows,
                           int dim4, cudaStream_t s);

// Adaptive RK45 integrator state (mulan_rk45.cu)
struct Rk45Params {
  int64_t n, k_stride;
  int n_k, of_y;
  double coef[7];
  double h, rtol, atol;
  const double *y, *y_new;
  const float* K;
  float* y_stage;
  double *y_out, *scratch;
};
cudaError_t launch_rk45_stage(const Rk45Params& p, cudaStream_t s);
cudaError_t launch_rk45_norm(const Rk45Params& p, double* out, cudaStream_t s);

// JAX-compatible threefry draws (mulan_rng.cu): kind 0 = bits, 1 = uniform, 2 = normal
cudaError_t launch_rng_draw(int kind, uint32_t k0, uint32_t k1, long long n, float minval,
                            float maxval, void* out, cudaStream_t s);

// <<<grid, block, 0, s>>> with, optionally, the programmatic-stream-serialization attribute
// (MULAN_FLAG_PDL): the kernel may start while its predecessor on the stream drains; every
// kernel of the path executes griddepcontrol.wait before its first global access.
template <typename P>
inline cudaError_t launch_kernel(void (*kernel)(const P), int grid, int block, cudaStream_t s,
                                 bool pdl, const P& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

// Number of CTAs of `kernel` (kThreads threads, static shared memory only) that are resident
// on the current device at once: the grid size of the persistent kernels.
// Queried per call (two driver look-ups, no caching): the answer belongs to the CURRENT device,
// and an XLA-style host calls in from one thread per device.
inline int resident_ctas(const void* kernel, int threads = kThreads) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess ||
      per_sm < 1)
    per_sm = 1;
  return sms * per_sm;
}

// Small launches (the literal per-GPU batch of 128 rows, ldm/configs/cifar10-conditioned.py:88)
// are latency bound: with one CTA per row and at most one row per SM, a row is spread over
// kLatencyThreads threads (one float4 column each for D = 3072) instead of the throughput
// shapes' 3-6 columns per thread.  Per-pixel outputs do not depend on the shape; per-row sums
// differ by float32 summation order (a launch's shape is a function of its row count only).
constexpr int kLatencyThreads = 768;
// The row count that decides the launch shape: the launch's own, unless the host-buffer pipeline
// (mulan_elbo_host) is cutting ONE batch into row chunks -- then the whole batch's, so that a
// chunked batch and a single launch of it produce the same bits (defined in mulan_abi.cu).
extern thread_local int tl_shape_rows;
inline int shape_rows(int rows) { return tl_shape_rows > 0 ? tl_shape_rows : rows; }
inline int latency_rows() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// Which loss formula the post / bwd_pre kernels run for desc->param.
//
// velocity_from_epsilon (ldm/model_mulan_velocity.py:246-249, 256-260) is the epsilon loss in
// disguise: with z_t = alpha f + sigma eps, e^gamma = v/(1-v) and alpha^2 + sigma^2 = 1,
//   v_hat = -e^{g/2} z_t + sqrt(1+e^g) net = (net - sigma z_t)/alpha
//   v_target - v_hat = (alpha^2 + sigma^2) eps/alpha - net/alpha = (eps - net)/alpha
//   (1-v) w (v_target - v_hat)^2 = w (eps - net)^2,
// so loss_diff, n_bar and (a,b,c)_bar are exactly those of MULAN_PARAM_EPS (every direct
// gamma / z_t dependence of the literal formula cancels analytically).  Against the golden
// vectors produced by the reference's own source the epsilon form differs by 1.3e-7 relative in
// loss_diff and 3e-6 in the gradients in float32, and by 1e-15 in float64
// (tests/test_oracle_golden.py::test_vfe_is_eps): far inside the 1e-5 / 1e-4 tolerances, and it
// saves 5 B/sub-pixel and ~40 instructions per sub-pixel.  MULAN_VFE_LITERAL=1 keeps the
// literal formula (A/B measurements, tests).
int kernel_param(int param);

// Publishes `msg` as this thread's mulan_last_error() (defined in mulan_abi.cu).
void set_last_error(const char* msg);

cudaError_t launch_fwd_pre(const FwdPreParams& p, cudaStream_t s);
int fwd_pre_variant(const FwdPreParams& p);
cudaError_t launch_fwd_post(const PostParams& p, cudaStream_t s);
cudaError_t launch_bwd_post(const PostParams& p, cudaStream_t s);
cudaError_t launch_fwd_bwd_post(const PostParams& p, cudaStream_t s);
cudaError_t launch_scale_rows(float* v, const float* num, const float* den, int rows, int dim4,
                              cudaStream_t s);
cudaError_t launch_bwd_pre(const BwdPreParams& p, cudaStream_t s);
cudaError_t launch_discrete_w(const DiscreteWParams& p, cudaStream_t s);
cudaError_t launch_adamw_ema(const mulan_adamw_desc& d, float* p, const float* g, float* mu,
                             float* nu, float* ema, cudaStream_t s);
cudaError_t launch_grad_sumsq(const float* g, long long n, double* scratch, float* out,
                              cudaStream_t s);
cudaError_t launch_aux_topk_fwd(int rows, int latent, int k, int noise_kind, const float* logits,
                                const float* noise, float* embedding, float* kl_z,
                                cudaStream_t s);
cudaError_t launch_aux_topk_bwd(int rows, int latent, int k, int noise_kind, const float* logits,
                                const float* noise, const float* emb_bar,
                                const float* klz_bar, float* logits_bar, cudaStream_t s);
cudaError_t launch_aux_gumbel(bool bwd, int rows, int latent, float tau, const float* logits,
                              const float* noise, const float* emb_bar, const float* klz_bar,
                              float* out, float* kl_z, cudaStream_t s);
cudaError_t launch_aux_gaussian(bool bwd, int rows, int latent, const float* mu, const float* var,
                                const float* eps, const float* emb_bar, const float* klz_bar,
                                float* out0, float* out1, cudaStream_t s);
cudaError_t launch_bpd_reduce(const BpdReduceParams& p, bool pdl, cudaStream_t s);
size_t reduce_ws_bytes(int rows);

}  // namespace mulan
