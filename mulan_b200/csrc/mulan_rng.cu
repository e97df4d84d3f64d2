// Counter-based random draws of VDM.__call__ on the device (SURVEY.md 8f "next" row 2):
// jax.random.uniform / normal / bits for float32 under JAX's default threefry2x32 generator.
//
// Reference call sites: ldm/model_mulan_epsilon.py:287-292 (t0 = uniform(rng, ())), :315 and
// :327 (eps_0, eps = normal(rng, shape)); ldm/model_mulan_velocity.py:198-203, :223, :235.
// The arithmetic is JAX's (un-vendored; jax <= 0.4.23 per the reference's README.md:26):
//   bits    jax._src.prng.threefry_random_bits, non-partitionable: counters iota(n) (+ one 0 if
//           n is odd) split in halves x0 | x1; (y0, y1) = threefry2x32(key, x0, x1);
//           bits = concat(y0, y1)[:n]      -> element i pairs with element i + ceil(n/2)
//   uniform (bits >> 9 | 0x3F800000) as float - 1;  max(minval, f (maxval - minval) + minval)
//   normal  sqrt(2) erf_inv(uniform(nextafter(-1, 0), 1)),  erf_inv = XLA's float32 ErfInv
//           (Giles 2010): w = -log1p(-x x); degree-8 polynomial in w - 2.5 or sqrt(w) - 3
// Oracle: oracle/jax_rng_oracle.py, pinned to the Random123 vectors and to values printed in
// the JAX documentation (tests/test_rng.py).  The Flax make_rng key derivation stays on the
// framework side: the kernels take the raw 2 x uint32 key of each draw.
//
// One thread evaluates ONE threefry block and produces BOTH of its outputs (elements i and
// i + half), so no round is wasted; both stores are coalesced.  ~75 integer instructions per
// block + ~70 float instructions per normal: instruction-issue bound (4 B written per element).
#include "mulan_kernels.h"
#include "mulan_rng.cuh"

namespace mulan {

namespace {

enum { kBits = 0, kUniform = 1, kNormal = 2 };

struct RngParams {
  uint32_t k0, k1;
  long long n, half;
  float minval, span;
  void* out;
};

template <int KIND>
__device__ __forceinline__ void rng_store(const RngParams& p, long long i, uint32_t bits) {
  if (KIND == kBits) {
    reinterpret_cast<uint32_t*>(p.out)[i] = bits;
  } else if (KIND == kUniform) {
    reinterpret_cast<float*>(p.out)[i] = bits_to_uniform(bits, p.minval, p.span);
  } else {
    // sqrt(2) rounded to float32, as np.array(np.sqrt(2), float32)
    reinterpret_cast<float*>(p.out)[i] =
        1.41421354f * erf_inv_xla(bits_to_uniform(bits, p.minval, p.span));
  }
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) rng_kernel(const RngParams p) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < p.half; i += stride) {
    const long long j = i + p.half;
    uint32_t x0 = (uint32_t)i;
    uint32_t x1 = j < p.n ? (uint32_t)j : 0u;       // the zero appended for odd n
    threefry2x32(p.k0, p.k1, x0, x1);
    rng_store<KIND>(p, i, x0);
    if (j < p.n) rng_store<KIND>(p, j, x1);
  }
}

template <int KIND>
cudaError_t launch_rng(RngParams p, cudaStream_t s) {
  // per call, for the CURRENT device (no process-wide cache: one host thread per device)
  const int max_ctas = resident_ctas((const void*)rng_kernel<KIND>);
  const long long want = (p.half + kThreads - 1) / kThreads;
  const int grid = (int)(want < max_ctas ? want : max_ctas);
  rng_kernel<KIND><<<grid < 1 ? 1 : grid, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_rng_draw(int kind, uint32_t k0, uint32_t k1, long long n, float minval,
                            float maxval, void* out, cudaStream_t s) {
  RngParams p;
  p.k0 = k0; p.k1 = k1; p.n = n; p.half = (n + 1) / 2;
  p.minval = minval; p.span = maxval - minval; p.out = out;
  switch (kind) {
    case kBits: return launch_rng<kBits>(p, s);
    case kUniform: return launch_rng<kUniform>(p, s);
    default: return launch_rng<kNormal>(p, s);
  }
}

}  // namespace mulan
