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

// Four consecutive elements per thread: one LDG.128 per stage row, two per float64 vector
// (the scalar form issues 4-byte loads and runs at 64 % of the HBM roofline).  Per element the
// arithmetic and its order are those of stage_sum, so results are bit-identical to it.
struct Quad { double v[4]; };
__device__ __forceinline__ Quad stage_sum4(const Rk45Params& p, int64_t i4) {
  Quad a = {{0.0, 0.0, 0.0, 0.0}};
#pragma unroll
  for (int j = 0; j < 7; ++j)
    if (j < p.n_k) {
      const float4 k = __ldg(reinterpret_cast<const float4*>(p.K + (size_t)j * p.k_stride) + i4);
      const double c = p.coef[j];
      a.v[0] += c * (double)k.x; a.v[1] += c * (double)k.y;
      a.v[2] += c * (double)k.z; a.v[3] += c * (double)k.w;
    }
#pragma unroll
  for (int e = 0; e < 4; ++e) a.v[e] *= p.h;
  return a;
}
__device__ __forceinline__ Quad load4d(const double* q, int64_t i4) {
  const double2 lo = reinterpret_cast<const double2*>(q)[2 * i4];
  const double2 hi = reinterpret_cast<const double2*>(q)[2 * i4 + 1];
  return Quad{{lo.x, lo.y, hi.x, hi.y}};
}

template <bool VEC>
__global__ void __launch_bounds__(kRkThreads) rk45_stage_kernel(const Rk45Params p) {
  const int64_t step = (int64_t)gridDim.x * kRkThreads;
  const int64_t first = (int64_t)blockIdx.x * kRkThreads + threadIdx.x;
  int64_t tail0 = 0;
  if (VEC) {
    const int64_t n4 = p.n >> 2;
    tail0 = n4 << 2;
    for (int64_t i4 = first; i4 < n4; i4 += step) {
      const Quad y = load4d(p.y, i4), s = stage_sum4(p, i4);
      double v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = y.v[e] + s.v[e];
      if (p.y_stage != nullptr)
        reinterpret_cast<float4*>(p.y_stage)[i4] =
            make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
      if (p.y_out != nullptr) {
        reinterpret_cast<double2*>(p.y_out)[2 * i4] = make_double2(v[0], v[1]);
        reinterpret_cast<double2*>(p.y_out)[2 * i4 + 1] = make_double2(v[2], v[3]);
      }
    }
  }
  for (int64_t i = tail0 + first; i < p.n; i += step) {      // everything, or the <= 3 leftovers
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

template <bool VEC>
__global__ void __launch_bounds__(kRkThreads) rk45_norm_partial_kernel(const Rk45Params p) {
  __shared__ double red[kRkThreads / 32];
  const int64_t step = (int64_t)gridDim.x * kRkThreads;
  const int64_t first = (int64_t)blockIdx.x * kRkThreads + threadIdx.x;
  double acc = 0.0;
  int64_t tail0 = 0;
  if (VEC) {
    const int64_t n4 = p.n >> 2;
    tail0 = n4 << 2;
    for (int64_t i4 = first; i4 < n4; i4 += step) {
      const Quad y = load4d(p.y, i4);
      Quad yn = y;
      if (p.y_new != nullptr) yn = load4d(p.y_new, i4);
      Quad v = y;
      if (!p.of_y) v = stage_sum4(p, i4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double mag = fmax(fabs(y.v[e]), fabs(yn.v[e]));
        const double q = v.v[e] / (p.atol + mag * p.rtol);
        acc += q * q;
      }
    }
  }
  for (int64_t i = tail0 + first; i < p.n; i += step) {
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

// 16-byte alignment of everything a launch touches: the float64 vectors, the float32 stage
// vector, and every row of K (base and row stride).
bool vec_ok(const Rk45Params& p) {
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return a16(p.y) && a16(p.y_new) && a16(p.y_out) && a16(p.y_stage) && a16(p.K) &&
         (p.k_stride & 3) == 0 && p.n >= 4;
}

// Grid-stride kernels.  Measured at 50 M elements: the stage kernel is fastest with more CTAs
// than are resident at once (105 % of the measured HBM peak with 2048, 102 % with a resident-only
// grid); the norm kernel, whose float64 divisions keep each CTA busy longer, with a resident-only
// grid (97 % vs 91 %).  Both bounds keep the number of partials within the scratch, and the fold
// order is fixed for a given n.
int stage_blocks(int64_t n) {
  const int64_t want = ((n + 3) / 4 + kRkThreads - 1) / kRkThreads;
  const int64_t b = want < MULAN_RK45_SCRATCH ? want : MULAN_RK45_SCRATCH;
  return (int)(b < 1 ? 1 : b);
}
int norm_blocks(int64_t n) {
  const int64_t want = ((n + 3) / 4 + kRkThreads - 1) / kRkThreads;
  // per call, for the CURRENT device (no process-wide cache: one host thread per device)
  const int64_t cap = resident_ctas((const void*)rk45_norm_partial_kernel<true>, kRkThreads);
  int64_t b = want < cap ? want : cap;
  if (b > MULAN_RK45_SCRATCH) b = MULAN_RK45_SCRATCH;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

cudaError_t launch_rk45_stage(const Rk45Params& p, cudaStream_t stream) {
  if (vec_ok(p)) rk45_stage_kernel<true><<<stage_blocks(p.n), kRkThreads, 0, stream>>>(p);
  else           rk45_stage_kernel<false><<<stage_blocks(p.n), kRkThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_rk45_norm(const Rk45Params& p, double* out, cudaStream_t stream) {
  const int blocks = norm_blocks(p.n);
  if (vec_ok(p)) rk45_norm_partial_kernel<true><<<blocks, kRkThreads, 0, stream>>>(p);
  else           rk45_norm_partial_kernel<false><<<blocks, kRkThreads, 0, stream>>>(p);
  rk45_norm_final_kernel<<<1, kRkThreads, 0, stream>>>(p.scratch, blocks, out);
  return cudaGetLastError();
}

}  // namespace mulan
