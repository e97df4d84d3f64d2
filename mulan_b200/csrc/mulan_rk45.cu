// Device-resident state of the adaptive Dormand-Prince 5(4) integrator behind the exact-
// likelihood evaluation and the probability-flow ODE sampler (SURVEY.md 8f "next" row 4).
//
// Reference statements:
//   get_ode_likelihood_fn / likelihood_fn   ldm/notebook_utils.py:264-373
//       init = concat(flatten(data), zeros(B));  solve_ivp(ode_func, (0, 1), init, rtol, atol,
//       method='RK45');  ode_func(t, x) = concat(flatten(drift), flatten(logp_grad))
//   get_sample_fn / sample_fn               ldm/notebook_utils.py:376-433   (t: 1 -> 0)
//   _to_flattened_numpy / _from_flattened_numpy  :193-200
//       the integrator state is float64 on the HOST; every function evaluation casts it to
//       float32, ships it to the devices, and ships the float32 derivative back.
//   scipy.integrate.solve_ivp(method='RK45') is the un-vendored third-party dependency
//   (scipy, unpinned in the reference's requirements; 1.18.1 in this image): explicit
//   Runge-Kutta of order 5(4), Dormand & Prince 1980, with the step control of Hairer,
//   Norsett & Wanner, "Solving ODEs I", Sec. II.4 - restated in oracle/rk45_oracle.py.
//
// Here the float64 state y[n] and the seven float32 stage derivatives K[7][n] stay in HBM.
// The only value that crosses to the host per step attempt is one double (the squared error
// norm); the accept/reject arithmetic on it is host code (mulan_b200/ode.py), exactly the
// scalar arithmetic scipy performs.
//
//   stage:  y_s = y + (sum_j a_sj K_j) h      -> float32 (denoiser input) and/or float64
//   norm :  sum_i (v_i / (atol + rtol max(|y_i|, |y_new_i|)))^2,   v = y  or  (sum_j e_j K_j) h
//           two launches: per-block partials in a fixed order, then one block folds them, so
//           the result does not depend on scheduling (an accept/reject decision is a
//           threshold on it).
//
// Traffic per element: stage s reads 8 + 4 s B and writes 4 (+8) B; the error norm reads
// 8 + 8 + 28 B.  n = B (3072 + 1) elements: HBM-bound, no reuse, coalesced grid-stride.
#include "mulan_kernels.h"

namespace mulan {

namespace {

constexpr int kRkThreads = 256;

__device__ __forceinline__ double stage_sum(const Rk45Params& p, int64_t i) {
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < 7; ++j)
    if (j < p.n_k) acc += p.coef[j] * (double)__ldg(p.K + (size_t)j * p.k_stride + i);
  return acc * p.h;
}

__global__ void __launch_bounds__(kRkThreads) rk45_stage_kernel(const Rk45Params p) {
  const int64_t step = (int64_t)gridDim.x * kRkThreads;
  for (int64_t i = (int64_t)blockIdx.x * kRkThreads + threadIdx.x; i < p.n; i += step) {
    const double v = p.y[i] + stage_sum(p, i);
    if (p.y_stage != nullptr) p.y_stage[i] = (float)v;
    if (p.y_out != nullptr) p.y_out[i] = v;
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum_d(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < kRkThreads / 32 ? red[lane] : 0.0;
    t = warp_sum_d(t);
  }
  return t;   // valid in thread 0
}

__global__ void __launch_bounds__(kRkThreads) rk45_norm_partial_kernel(const Rk45Params p) {
  __shared__ double red[kRkThreads / 32];
  const int64_t step = (int64_t)gridDim.x * kRkThreads;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kRkThreads + threadIdx.x; i < p.n; i += step) {
    const double yi = p.y[i];
    double mag = fabs(yi);
    if (p.y_new != nullptr) mag = fmax(mag, fabs(p.y_new[i]));
    const double scale = p.atol + mag * p.rtol;
    const double v = p.of_y ? yi : stage_sum(p, i);
    const double q = v / scale;
    acc += q * q;
  }
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) p.scratch[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kRkThreads) rk45_norm_final_kernel(const double* partial,
                                                                      int n_partial, double* out) {
  __shared__ double red[kRkThreads / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_partial; i += kRkThreads) acc += partial[i];
  const double t = block_sum_d(acc, red);
  if (threadIdx.x == 0) out[0] = t;
}

int rk_blocks(int64_t n) {
  const int64_t want = (n + kRkThreads - 1) / kRkThreads;
  static const int64_t cap = resident_ctas((const void*)rk45_norm_partial_kernel);
  int64_t b = want < cap ? want : cap;
  if (b > MULAN_RK45_SCRATCH) b = MULAN_RK45_SCRATCH;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

cudaError_t launch_rk45_stage(const Rk45Params& p, cudaStream_t stream) {
  rk45_stage_kernel<<<rk_blocks(p.n), kRkThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_rk45_norm(const Rk45Params& p, double* out, cudaStream_t stream) {
  const int blocks = rk_blocks(p.n);
  rk45_norm_partial_kernel<<<blocks, kRkThreads, 0, stream>>>(p);
  rk45_norm_final_kernel<<<1, kRkThreads, 0, stream>>>(p.scratch, blocks, out);
  return cudaGetLastError();
}

}  // namespace mulan
