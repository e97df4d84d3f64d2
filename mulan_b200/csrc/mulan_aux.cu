// Small per-example ops of the ELBO: the auxiliary-latent KL + relaxed top-k embedding
// ([B, 50]) and the final VDMOutput / bits-per-dim assembly ([B] -> 6 scalars).
//
// Reference statements:
//   _gumbel_kl_loss            ldm/model_mulan_epsilon.py:205-210
//   _gamma_noise               ldm/model_mulan_epsilon.py:221-231
//   _topk_embedding_and_loss   ldm/model_mulan_epsilon.py:233-252 (velocity.py:106-120)
//   VDMOutput assembly         ldm/model_mulan_epsilon.py:357-363
//   Experiment_VDM.loss_fn     ldm/experiment_vdm.py:62-74
//
// aux: one warp per row, two slots per lane (L <= 64); all reductions are warp shuffles.
#include "mulan_kernels.h"
#include "mulan_reduce.cuh"

namespace mulan {

constexpr int kAuxRowsPerCta = 4;  // 4 warps per CTA

struct AuxRow {
  float l[2];       // raw logits
  float q[2];       // softmax(logits)
  float lq[2];      // log_softmax(logits)
  float kl;         // KL(q || uniform)
  float l3[2];      // centred noisy logits
  float norm;       // ||l3||_2
  float soft[2];    // l3 / norm
  float hard[2];    // top-k indicator
  bool valid[2];
};

// noise_kind 0: G is the jax.random.gamma draw [10, B, L] of _gamma_noise; 1: G is an additive
// noise [B, L] (topk_noise_type == 'gumbel', ldm/model_mulan_epsilon.py:238-239).
__device__ __forceinline__ AuxRow aux_forward(int row, int rows, int L, int k, int noise_kind,
                                              const float* __restrict__ logits,
                                              const float* __restrict__ G) {
  const int lane = threadIdx.x & 31;
  AuxRow r;
  float mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    r.valid[s] = i < L;
    r.l[s] = r.valid[s] ? __ldg(logits + (size_t)row * L + i) : -INFINITY;
    mx = fmaxf(mx, r.l[s]);
  }
  mx = warp_max(mx);
  float un[2], se = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    un[s] = r.valid[s] ? expf(r.l[s] - mx) : 0.f;
    se += un[s];
  }
  se = warp_sum(se);
  const float lse = logf(se);
  const float log_unif = logf((float)(1.0 / (double)L));
  float kl = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    r.q[s] = __fdiv_rn(un[s], se);
    r.lq[s] = (r.l[s] - mx) - lse;
    if (r.valid[s]) kl += r.q[s] * (r.lq[s] - log_unif);
  }
  r.kl = warp_sum(kl);

  // gamma noise: s = 10 * ((sum_i G_i / (k/i)) - log 10) / k
  float l2[2], sm = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    float noise = 0.f;
    if (G != nullptr && r.valid[s] && noise_kind == 1) {
      noise = __ldg(G + (size_t)row * L + i);
    } else if (G != nullptr && r.valid[s]) {
      float acc = 0.f;
      for (int m = 1; m <= 10; ++m) {
        const float beta = __fdiv_rn((float)k, (float)m);
        acc += __fdiv_rn(__ldg(G + ((size_t)(m - 1) * rows + row) * L + i), beta);
      }
      acc = acc - logf(10.0f);
      noise = 10.0f * __fdiv_rn(acc, (float)k);
    }
    l2[s] = r.valid[s] ? r.l[s] + noise : 0.f;
    sm += l2[s];
  }
  const float mean = __fdiv_rn(warp_sum(sm), (float)L);
  float ss = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    r.l3[s] = r.valid[s] ? l2[s] - mean : 0.f;
    ss += r.l3[s] * r.l3[s];
  }
  r.norm = sqrtf(warp_sum(ss));
  // rank by counting strictly greater entries: hard = (l3 >= k-th largest)
  int cnt[2] = {0, 0};
  for (int j = 0; j < L; ++j) {
    const float vj = __shfl_sync(0xffffffffu, j < 32 ? r.l3[0] : r.l3[1], j & 31);
    cnt[0] += vj > r.l3[0];
    cnt[1] += vj > r.l3[1];
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    r.soft[s] = __fdiv_rn(r.l3[s], r.norm);
    r.hard[s] = (r.valid[s] && cnt[s] < k) ? 1.0f : 0.0f;
  }
  return r;
}

__global__ void __launch_bounds__(32 * kAuxRowsPerCta)
aux_topk_fwd_kernel(int rows, int L, int k, int noise_kind, const float* __restrict__ logits,
                    const float* __restrict__ G, float* __restrict__ emb,
                    float* __restrict__ kl_z) {
  const int row = blockIdx.x * kAuxRowsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const AuxRow r = aux_forward(row, rows, L, k, noise_kind, logits, G);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    // stop_gradient(hard - soft) + soft
    if (r.valid[s]) emb[(size_t)row * L + i] = (r.hard[s] - r.soft[s]) + r.soft[s];
  }
  if (lane == 0) kl_z[row] = r.kl;
}

__global__ void __launch_bounds__(32 * kAuxRowsPerCta)
aux_topk_bwd_kernel(int rows, int L, int k, int noise_kind, const float* __restrict__ logits,
                    const float* __restrict__ G, const float* __restrict__ emb_bar,
                    const float* __restrict__ klz_bar, float* __restrict__ logits_bar) {
  const int row = blockIdx.x * kAuxRowsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const AuxRow r = aux_forward(row, rows, L, k, noise_kind, logits, G);
  const float log_unif = logf((float)(1.0 / (double)L));
  float sb[2], dot = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    sb[s] = (emb_bar != nullptr && r.valid[s]) ? __ldg(emb_bar + (size_t)row * L + i) : 0.f;
    dot += sb[s] * r.soft[s];
  }
  dot = warp_sum(dot);
  // soft = l3/||l3||: l3_bar = (soft_bar - soft <soft, soft_bar>)/||l3||
  float l3b[2], sm = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    l3b[s] = r.valid[s] ? __fdiv_rn(sb[s] - r.soft[s] * dot, r.norm) : 0.f;
    sm += l3b[s];
  }
  const float mean = __fdiv_rn(warp_sum(sm), (float)L);
  const float kb = klz_bar != nullptr ? __ldg(klz_bar + row) : 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    if (r.valid[s]) {
      const float dkl = r.q[s] * ((r.lq[s] - log_unif) - r.kl);
      logits_bar[(size_t)row * L + i] = (l3b[s] - mean) + kb * dkl;
    }
  }
}

cudaError_t launch_aux_topk_fwd(int rows, int latent, int k, int noise_kind, const float* logits,
                                const float* noise, float* embedding, float* kl_z,
                                cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const int grid = (rows + kAuxRowsPerCta - 1) / kAuxRowsPerCta;
  aux_topk_fwd_kernel<<<grid, 32 * kAuxRowsPerCta, 0, s>>>(rows, latent, k, noise_kind, logits,
                                                          noise, embedding, kl_z);
  return cudaGetLastError();
}

cudaError_t launch_aux_topk_bwd(int rows, int latent, int k, int noise_kind, const float* logits,
                                const float* noise, const float* emb_bar,
                                const float* klz_bar, float* logits_bar, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const int grid = (rows + kAuxRowsPerCta - 1) / kAuxRowsPerCta;
  aux_topk_bwd_kernel<<<grid, 32 * kAuxRowsPerCta, 0, s>>>(rows, latent, k, noise_kind, logits,
                                                          noise, emb_bar, klz_bar, logits_bar);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// latent_type == 'gumbel' (ldm/model_mulan_epsilon.py:195-219): straight-through Gumbel-softmax.
//   kl    = _gumbel_kl_loss(logits)                      (raw logits)
//   l     = (logits + gumbel_noise) / tau                tau = max(.5, exp(-1e-5 step)), host
//   soft  = softmax(l); hard = one_hot(argmax(l)); emb = stop_grad(hard - soft) + soft
// latent_type == 'gaussian' (:264-270):
//   emb = mu + sqrt(var) eps_z;  kl = .5 sum(mu^2 + var - log var - 1)
// One warp per row, two slots per lane, like the top-k op.
// ---------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(32 * kAuxRowsPerCta)
aux_gumbel_kernel(int rows, int L, float tau, const float* __restrict__ logits,
                  const float* __restrict__ noise, const float* __restrict__ emb_bar,
                  const float* __restrict__ klz_bar, float* __restrict__ out,
                  float* __restrict__ kl_z) {
  const int row = blockIdx.x * kAuxRowsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // KL and softmax of the raw logits: reuse the top-k forward with no noise and k = L
  const AuxRow r = aux_forward(row, rows, L, L, 0, logits, nullptr);
  float l[2], mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    const float g = (noise != nullptr && r.valid[s]) ? __ldg(noise + (size_t)row * L + i) : 0.f;
    l[s] = r.valid[s] ? __fdiv_rn(r.l[s] + g, tau) : -INFINITY;
    mx = fmaxf(mx, l[s]);
  }
  mx = warp_max(mx);
  float un[2], se = 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    un[s] = r.valid[s] ? expf(l[s] - mx) : 0.f;
    se += un[s];
  }
  se = warp_sum(se);
  float soft[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) soft[s] = __fdiv_rn(un[s], se);
  if (!BWD) {
    // argmax: FIRST index attaining the maximum (jnp.argmax)
    int best = 1 << 30;
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (r.valid[s] && l[s] == mx) best = min(best, lane + 32 * s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int i = lane + 32 * s;
      const float hard = (i == best) ? 1.0f : 0.0f;
      if (r.valid[s]) out[(size_t)row * L + i] = (hard - soft[s]) + soft[s];
    }
    if (lane == 0) kl_z[row] = r.kl;
  } else {
    const float log_unif = logf((float)(1.0 / (double)L));
    float sb[2], dot = 0.f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int i = lane + 32 * s;
      sb[s] = (emb_bar != nullptr && r.valid[s]) ? __ldg(emb_bar + (size_t)row * L + i) : 0.f;
      dot += sb[s] * soft[s];
    }
    dot = warp_sum(dot);
    const float kb = klz_bar != nullptr ? __ldg(klz_bar + row) : 0.f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int i = lane + 32 * s;
      if (r.valid[s]) {
        const float lbar = soft[s] * (sb[s] - dot);                  // softmax backward
        const float dkl = r.q[s] * ((r.lq[s] - log_unif) - r.kl);
        out[(size_t)row * L + i] = __fdiv_rn(lbar, tau) + kb * dkl;
      }
    }
  }
}

cudaError_t launch_aux_gumbel(bool bwd, int rows, int latent, float tau, const float* logits,
                              const float* noise, const float* emb_bar, const float* klz_bar,
                              float* out, float* kl_z, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const int grid = (rows + kAuxRowsPerCta - 1) / kAuxRowsPerCta;
  if (bwd) aux_gumbel_kernel<true><<<grid, 32 * kAuxRowsPerCta, 0, s>>>(
      rows, latent, tau, logits, noise, emb_bar, klz_bar, out, kl_z);
  else aux_gumbel_kernel<false><<<grid, 32 * kAuxRowsPerCta, 0, s>>>(
      rows, latent, tau, logits, noise, emb_bar, klz_bar, out, kl_z);
  return cudaGetLastError();
}

template <bool BWD>
__global__ void __launch_bounds__(32 * kAuxRowsPerCta)
aux_gaussian_kernel(int rows, int L, const float* __restrict__ mu, const float* __restrict__ var,
                    const float* __restrict__ eps, const float* __restrict__ emb_bar,
                    const float* __restrict__ klz_bar, float* __restrict__ out0,
                    float* __restrict__ out1) {
  const int row = blockIdx.x * kAuxRowsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float kl = 0.f;
  const float kb = (BWD && klz_bar != nullptr) ? __ldg(klz_bar + row) : 0.f;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int i = lane + 32 * s;
    if (i >= L) continue;
    const size_t o = (size_t)row * L + i;
    const float m = __ldg(mu + o), v = __ldg(var + o), e = __ldg(eps + o);
    const float sd = sqrtf(v);
    if (!BWD) {
      out0[o] = m + sd * e;                                          // embedding
      kl += m * m + v - logf(v) - 1.0f;
    } else {
      const float eb = emb_bar != nullptr ? __ldg(emb_bar + o) : 0.f;
      out0[o] = eb + kb * m;                                         // mu_bar
      out1[o] = eb * __fdiv_rn(e, 2.0f * sd) + kb * 0.5f * (1.0f - __fdiv_rn(1.0f, v));  // var_bar
    }
  }
  if (!BWD) {
    kl = warp_sum(kl);
    if (lane == 0) out1[row] = 0.5f * kl;
  }
}

cudaError_t launch_aux_gaussian(bool bwd, int rows, int latent, const float* mu, const float* var,
                                const float* eps, const float* emb_bar, const float* klz_bar,
                                float* out0, float* out1, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const int grid = (rows + kAuxRowsPerCta - 1) / kAuxRowsPerCta;
  if (bwd) aux_gaussian_kernel<true><<<grid, 32 * kAuxRowsPerCta, 0, s>>>(
      rows, latent, mu, var, eps, emb_bar, klz_bar, out0, out1);
  else aux_gaussian_kernel<false><<<grid, 32 * kAuxRowsPerCta, 0, s>>>(
      rows, latent, mu, var, eps, emb_bar, klz_bar, out0, out1);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// VDMOutput + loss_fn scalars (mulan_bpd_reduce): the fixed-order reduction of mulan_reduce.cuh.
// With a workspace: one CTA per group of 128 rows, "last CTA done" finalisation (one launch, a
// few microseconds at 16384 rows; the single-CTA loop of round 1 took 16-22).  Without: one CTA
// walks the groups and reproduces the same summation order bit for bit.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
bpd_reduce_parallel_kernel(const BpdReduceParams p) {
  __shared__ float red[kWarps][5];
  __shared__ int s_flag;
  pdl_release_dependents();
  pdl_wait_for_primary();
  const int g = blockIdx.x;
  red_rows_done(p, g, min(kRedGroup, p.rows - g * kRedGroup), red, &s_flag);
}

__global__ void __launch_bounds__(kThreads)
bpd_reduce_single_kernel(const BpdReduceParams p) {
  __shared__ float red[kWarps][5];
  __shared__ float vacc[kThreads][5];   // the final step's per-thread accumulators
  pdl_release_dependents();
  pdl_wait_for_primary();
#pragma unroll
  for (int k = 0; k < 5; ++k) vacc[threadIdx.x][k] = 0.f;
  const int G = red_groups(p.rows);
  float acc[5];
  for (int g = 0; g < G; ++g) {
    red_group_sum(p, g, acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) vacc[g % kThreads][k] += acc[k];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 5; ++k) acc[k] = vacc[threadIdx.x][k];
  block_sum<5>(acc, red);
  if (threadIdx.x == 0) red_write_scalars(p, acc);
}

// mulan_scalar_board_read: the mean over the ranks of the latest step THIS rank has published.
// Waits (bounded) until every rank's row of that slot carries the step's tag, then sums in rank
// order (identical on every rank) and divides by the world size.
__global__ void __launch_bounds__(32)
board_read_kernel(const ScalarBoard b, float* __restrict__ mean_out,
                  unsigned* __restrict__ epoch_out) {
  const int lane = threadIdx.x;
  const float* mine = b.boards[b.rank];
  const unsigned step = *reinterpret_cast<const unsigned*>(mine + kBoardSlots * 8 * kBoardRow);
  const int slot = (int)(step % (unsigned)kBoardSlots);
  bool ok = true;
  float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
  if (lane < b.world) {
    const float4* row = reinterpret_cast<const float4*>(mine + ((size_t)slot * 8 + lane) * kBoardRow);
    const long long t0 = clock64();
    for (;;) {
      // each half of a row is one 128-bit store carrying its own tag (board_publish)
      asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w) : "l"(row) : "memory");
      asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "l"(row + 1) : "memory");
      const unsigned t_lo = __float_as_uint(lo.w), t_hi = __float_as_uint(hi.w);
      if (t_lo == step && t_hi == step) break;
      // a peer may already be LATER in the ring (same slot, step + k * slots): its row for this
      // step is gone; report the step as incomplete rather than mixing steps
      if ((int)(t_lo - step) > 0 || (int)(t_hi - step) > 0 || clock64() - t0 > 8000000000LL) {
        ok = false;
        break;
      }
      __nanosleep(200);
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  // sum over the ranks in rank order (lane r holds rank r's row): sequential shuffles
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < b.world; ++r) {
    s[0] += __shfl_sync(0xffffffffu, lo.x, r); s[1] += __shfl_sync(0xffffffffu, lo.y, r);
    s[2] += __shfl_sync(0xffffffffu, lo.z, r); s[3] += __shfl_sync(0xffffffffu, hi.x, r);
    s[4] += __shfl_sync(0xffffffffu, hi.y, r); s[5] += __shfl_sync(0xffffffffu, hi.z, r);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k)
      mean_out[k] = ok ? __fdiv_rn(s[k], (float)b.world) : __int_as_float(0x7fc00000);
  }
  if (lane == 0 && epoch_out != nullptr) *epoch_out = ok ? step : 0u;
}

cudaError_t launch_board_read(const ScalarBoard& b, float* mean_out, unsigned* epoch_out,
                              cudaStream_t s) {
  board_read_kernel<<<1, 32, 0, s>>>(b, mean_out, epoch_out);
  return cudaGetLastError();
}

size_t reduce_ws_bytes(int rows) {
  const int G = red_groups(rows < 1 ? 1 : rows);
  return sizeof(unsigned) * (size_t)(red_partials_offset(G) + 8 * G);
}

cudaError_t launch_bpd_reduce(const BpdReduceParams& p, bool pdl, cudaStream_t s) {
  if (p.ws != nullptr)
    return launch_kernel(bpd_reduce_parallel_kernel, red_groups(p.rows), kThreads, s, pdl, p);
  return launch_kernel(bpd_reduce_single_kernel, 1, kThreads, s, pdl, p);
}

}  // namespace mulan
