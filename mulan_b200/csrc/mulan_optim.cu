// mulan_adamw_ema: the optimizer update that follows the gradient all-reduce of every train
// step, fused over ONE flat float32 parameter buffer.
//
// Reference statements (what TrainState.apply_gradients runs, ldm/train_state.py:70-102, with
// the optax chain of ldm/experiment.py:132-182 and the schedule of :106-129):
//   optax.adamw = scale_by_adam(b1, b2, eps) -> add_decayed_weights(wd, mask) -> scale(-lr)
//     mu  = (1-b1) g + b1 mu ;  nu = (1-b2) g^2 + b2 nu          (update_moment)
//     u   = (mu / (1-b1^t)) / (sqrt(nu / (1-b2^t)) + eps)        (bias_correction, eps_root=0)
//     u  += wd * p   where the decay mask holds (everything but biases, :135-141)
//     p  += -lr * u                                               (optax.apply_updates)
//   ema  = ema + (1 - ema_rate) (p_new - ema)                    (train_state.py:91-95)
//   (+ the pmean of the gradients, ldm/experiment.py:341: NCCL sums, `grad_scale` = 1/world
//    finishes the mean here instead of in a separate pass over the bucket)
//   optional optax.clip_by_global_norm(config.gradient_clip_norm) in front of the chain
//   (ldm/experiment.py:176-178):  n = sqrt(sum g^2);  g = n < max ? g : (g / n) * max.
//   mulan_grad_sumsq makes one extra 4 B/param read pass over the bucket (deterministic
//   two-launch reduction, float64 accumulation of float32 squares); the update kernel reads the
//   scalar from device memory, so there is no host synchronisation.
//
// Layout: parameters are laid out decayed-first, so the mask is one boundary index instead of
// a per-element byte.  Purely HBM-bound: p, g, mu, nu, ema read (20 B) and p, mu, nu, ema
// written (16 B) = 36 B per parameter, float4 accesses, one float4 column per thread.
#include "mulan_kernels.h"

namespace mulan {

struct AdamwParams {
  float *p, *mu, *nu, *ema;
  const float* g;
  long long n4;        // float4 count (n padded to 4 by the host side)
  long long decay4;    // float4 index below which weight decay applies
  float lr, b1, b2, om_b1, om_b2, eps, wd, one_minus_ema;
  float bc1, bc2;      // 1 - b1^t, 1 - b2^t
  float grad_scale;
  const float* sumsq;  // device scalar: sum of squares of the RAW bucket (before grad_scale)
  float clip;          // max global norm
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& mu, float& nu, float& ema,
                                          const AdamwParams& k, bool decay, bool clip,
                                          float g_norm) {
  g = g * k.grad_scale;
  if (clip) g = __fdiv_rn(g, g_norm) * k.clip;
  mu = k.om_b1 * g + k.b1 * mu;
  nu = k.om_b2 * (g * g) + k.b2 * nu;
  const float mu_hat = __fdiv_rn(mu, k.bc1);
  const float nu_hat = __fdiv_rn(nu, k.bc2);
  float u = __fdiv_rn(mu_hat, sqrtf(nu_hat) + k.eps);
  if (decay) u = u + k.wd * p;
  p = p + (-k.lr) * u;
  ema = ema + k.one_minus_ema * (p - ema);
}

// One float4 column per thread, one CTA per 256 columns.  (A resident grid-stride loop measured
// 89 % of the HBM roofline -- one iteration of loads in flight per thread; this form lets the
// block scheduler keep every SM's load queue full: 105 %.  Streaming cache hints made no
// difference either way.)
template <bool CLIP>
__global__ void __launch_bounds__(kThreads)
adamw_ema_kernel(const AdamwParams k) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= k.n4) return;
  float g_norm = 0.f;
  bool clip = false;
  if (CLIP) {
    g_norm = k.grad_scale * sqrtf(__ldg(k.sumsq));
    clip = !(g_norm < k.clip);
  }
  float4 P = reinterpret_cast<float4*>(k.p)[i];
  const float4 G = __ldg(reinterpret_cast<const float4*>(k.g) + i);
  float4 M = reinterpret_cast<float4*>(k.mu)[i];
  float4 V = reinterpret_cast<float4*>(k.nu)[i];
  float4 E = reinterpret_cast<float4*>(k.ema)[i];
  const bool decay = i < k.decay4;
  adamw_one(P.x, G.x, M.x, V.x, E.x, k, decay, clip, g_norm);
  adamw_one(P.y, G.y, M.y, V.y, E.y, k, decay, clip, g_norm);
  adamw_one(P.z, G.z, M.z, V.z, E.z, k, decay, clip, g_norm);
  adamw_one(P.w, G.w, M.w, V.w, E.w, k, decay, clip, g_norm);
  reinterpret_cast<float4*>(k.p)[i] = P;
  reinterpret_cast<float4*>(k.mu)[i] = M;
  reinterpret_cast<float4*>(k.nu)[i] = V;
  reinterpret_cast<float4*>(k.ema)[i] = E;
}

cudaError_t launch_adamw_ema(const mulan_adamw_desc& d, float* p, const float* g, float* mu,
                             float* nu, float* ema, cudaStream_t s) {
  if (d.n == 0) return cudaSuccess;
  AdamwParams k;
  k.p = p; k.g = g; k.mu = mu; k.nu = nu; k.ema = ema;
  k.n4 = d.n / 4; k.decay4 = d.n_decay / 4;
  // Python-double hyper-parameters become float32 where optax's weak-typed scalars would
  k.lr = (float)d.lr; k.b1 = (float)d.b1; k.b2 = (float)d.b2; k.eps = (float)d.eps;
  k.wd = (float)d.weight_decay;
  k.om_b1 = (float)(1.0 - d.b1); k.om_b2 = (float)(1.0 - d.b2);
  k.one_minus_ema = (float)(1.0 - d.ema_rate);
  // 1 - b^t: optax raises the Python float to an int32 array -> a float32 power
  k.bc1 = 1.0f - powf((float)d.b1, (float)d.step);
  k.bc2 = 1.0f - powf((float)d.b2, (float)d.step);
  k.grad_scale = (float)d.grad_scale;
  k.sumsq = d.grad_sumsq;
  k.clip = (float)d.clip_norm;
  const long long want = (k.n4 + kThreads - 1) / kThreads;
  if (want > 0x7fffffffLL) {
    set_last_error("mulan_adamw_ema: n exceeds 2^31 * 1024 parameters");
    return cudaErrorInvalidValue;
  }
  const int grid = (int)want;
  if (d.clip_norm > 0.0) adamw_ema_kernel<true><<<grid, kThreads, 0, s>>>(k);
  else                   adamw_ema_kernel<false><<<grid, kThreads, 0, s>>>(k);
  return cudaGetLastError();
}

// ---- global sum of squares of the gradient bucket (optax.global_norm ** 2) ------------------
namespace {

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double cta_sum_f64(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum_f64(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < kWarps ? red[lane] : 0.0;
    t = warp_sum_f64(t);
  }
  return t;
}

__global__ void __launch_bounds__(kThreads)
grad_sumsq_partial_kernel(const float* __restrict__ g, long long n4, double* __restrict__ part) {
  __shared__ double red[kWarps];
  const long long stride = (long long)gridDim.x * kThreads;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
    const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += (double)(G.x * G.x) + (double)(G.y * G.y) + (double)(G.z * G.z) + (double)(G.w * G.w);
  }
  const double t = cta_sum_f64(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kThreads)
grad_sumsq_final_kernel(const double* __restrict__ part, int n_part, float* __restrict__ out) {
  __shared__ double red[kWarps];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_part; i += kThreads) acc += part[i];
  const double t = cta_sum_f64(acc, red);
  if (threadIdx.x == 0) out[0] = (float)t;
}

}  // namespace

cudaError_t launch_grad_sumsq(const float* g, long long n, double* scratch, float* out,
                              cudaStream_t s) {
  const long long n4 = n / 4;
  long long want = (n4 + kThreads - 1) / kThreads;
  if (want < 1) want = 1;
  // as many CTAs as the scratch holds partials for (a resident-only grid measured 88 % of the
  // HBM roofline: too few loads in flight); the fold order stays fixed for a given n
  int grid = (int)(want < MULAN_SUMSQ_SCRATCH ? want : MULAN_SUMSQ_SCRATCH);
  grad_sumsq_partial_kernel<<<grid, kThreads, 0, s>>>(g, n4, scratch);
  grad_sumsq_final_kernel<<<1, kThreads, 0, s>>>(scratch, grid, out);
  return cudaGetLastError();
}

}  // namespace mulan
