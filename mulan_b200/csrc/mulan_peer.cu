// mulan_adamw_ema_peer: the gradient pmean FUSED with the optimizer update over NVLink peer
// memory (SURVEY.md 8f row 1 as written: "fused AdamW(+EMA) update fused with the gradient
// all-reduce bucket").  Reference statements: grads = jax.lax.pmean(grads, 'batch')
// (ldm/experiment.py:341) followed by state.apply_gradients (ldm/experiment.py:344 ->
// ldm/train_state.py:70-102) on every device.
//
// One process per GPU; every rank's gradient bucket, parameter buffer and a small flag block
// live in cudaMalloc'ed memory that the other ranks of the node map through CUDA IPC
// (mulan_peer_alloc / mulan_peer_open), so a kernel can load from and store to any peer over
// NVLink 5 / NVSwitch.  ONE kernel per step (or per gradient bucket) on every rank:
//
//   A  flag barrier  "my gradients for this range are final" -> every peer; wait for all peers
//   1  reduce-scatter by peer LOADS: rank r owns 1/world of the range and sums the `world`
//      gradient buckets element-wise in rank order 0..world-1 (fixed order: deterministic)
//   2  AdamW + EMA on the owned shard only -- mu, nu, ema are touched for 1/world of the
//      parameters (36 B/param of optimizer traffic becomes 36/world + 4 B)
//   3  all-gather by peer STORES: the new parameters of the shard go to every rank's buffer
//   B  flag barrier  "my shard has landed everywhere"; the kernel retires only when every
//      peer's shard has landed HERE, so stream order protects the next forward pass
//
// versus ncclAllReduce (reduce-scatter + all-gather of GRADIENTS over the same links) followed
// by a full-size update on every rank: the same NVLink bytes, one launch, no intermediate
// reduced-gradient buffer, 1/world of the optimizer's HBM traffic.
//
// Ordering.  A rank signals A from a kernel that is stream-ordered after its backward pass, so
// when rank r has seen all A flags every bucket is final and no peer still needs the OLD
// parameters; it signals B after a system-scope fence behind its last peer store, and nobody
// re-zeroes or re-accumulates a gradient bucket before its own kernel (which waits for every B)
// has retired.  Flags carry the call's epoch (strictly increasing), so no reset is needed.
// A spin that exceeds ~4 s sets the error word instead of hanging the GPU.
#include <stdio.h>
#include <stdlib.h>

#include "mulan_kernels.h"

namespace mulan {
namespace {

constexpr int kMaxPeers = 8;
constexpr int kFlagA = 0, kFlagB = kMaxPeers, kFlagCount = 2 * kMaxPeers, kFlagErr = kFlagCount + 1;

struct PeerParams {
  float* grads[kMaxPeers];
  float* params[kMaxPeers];
  unsigned* flags[kMaxPeers];
  float *mu, *nu, *ema;
  float *mc_grads, *mc_params;   // NVSwitch multicast addresses of the two buffers (MC kernels)
  int world, rank;
  unsigned epoch;
  long long lo4, n4;       // this rank's shard: float4 columns [lo4, lo4 + n4)
  long long decay4;
  float lr, b1, b2, om_b1, om_b2, eps, wd, one_minus_ema, bc1, bc2, grad_scale;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Peer gradient loads: plain L2-only loads (ld.global.cg).  They are ordered after the flag
// acquire (thread 0's ld.acquire.sys + the CTA barrier behind it), every address is read once per
// kernel and L1 starts clean at a kernel boundary, so there is nothing stale to hit; unlike
// ld.volatile (system-scope relaxed, one request per access, a compiler barrier each) the
// hardware coalesces them and the compiler keeps all of a thread's loads in flight.
__device__ __forceinline__ float4 ld_peer4(const float* p, long long i4) {
  return __ldcg(reinterpret_cast<const float4*>(p) + i4);
}

// NVLink SHARP through a multicast mapping (sm_90+): ONE load returns the element-wise float32 sum
// of the addressed location on every GPU of the multicast group, reduced inside the switch; ONE
// store writes every GPU's copy.  Per rank the NVLink volume of the exchange falls from
// 2 (N-1)/N of the bucket per direction to 1/N in (the reduced shard) + 1/N out (the new
// parameters): the kernel becomes HBM bound.
__device__ __forceinline__ float4 multimem_ld_reduce4(const float* mc, long long i4) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(reinterpret_cast<const float4*>(mc) + i4) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st4(float* mc, long long i4, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               ::"l"(reinterpret_cast<float4*>(mc) + i4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Wait until every peer's flag in `slot` has reached this call's epoch (thread 0 of a CTA).
__device__ __forceinline__ void wait_peers(const PeerParams& k, int slot) {
  const unsigned* mine = k.flags[k.rank];
  const long long t0 = clock64();
  for (int r = 0; r < k.world; ++r) {
    while ((int)(ld_acquire_sys(mine + slot + r) - k.epoch) < 0) {
      if (clock64() - t0 > 8000000000LL) {       // ~4 s at 2 GHz: report, do not hang
        k.flags[k.rank][kFlagErr] = 1u;
        return;
      }
      __nanosleep(200);
    }
  }
}

__device__ __forceinline__ void adamw_elem(float& p, float g, float& mu, float& nu, float& ema,
                                           const PeerParams& k, bool decay) {
  g = g * k.grad_scale;                        // the 1/world of the pmean
  mu = k.om_b1 * g + k.b1 * mu;
  nu = k.om_b2 * (g * g) + k.b2 * nu;
  const float mu_hat = __fdiv_rn(mu, k.bc1);
  const float nu_hat = __fdiv_rn(nu, k.bc2);
  float u = __fdiv_rn(mu_hat, sqrtf(nu_hat) + k.eps);
  if (decay) u = u + k.wd * p;
  p = p + (-k.lr) * u;
  ema = ema + k.one_minus_ema * (p - ema);
}

// COLS float4 columns per thread (COLS * (WORLD - 1) peer loads in flight per thread before the
// first dependent use): with two ranks a single remote load per thread leaves NVLink latency
// exposed (measured 176 GB/s per direction), so small worlds take more columns (8 / 2 / 1 columns
// for 2 / 4 / 8 ranks, measured: two columns at 8 ranks cost 1.09 instead of 0.92 ms for 285 MB).
// PF (multicast form only): the switch-reduced gradients of a CTA's NEXT chunk are requested
// before the current chunk is updated and stored, so the reduce-scatter stream (the switch pulls
// every GPU's bucket: outbound-heavy) and the all-gather stream (multimem.st: inbound-heavy) of
// the fabric stay busy at the same time instead of alternating.
template <int WORLD, int COLS, bool MC = false, bool PF = false>
__global__ void __launch_bounds__(kThreads)
adamw_ema_peer_kernel(const PeerParams k) {
  __shared__ int s_last;
  const int tid = threadIdx.x;
  // ---- A: my gradients are final (this kernel is stream-ordered after backward)
  if (blockIdx.x == 0 && tid < WORLD) {
    __threadfence_system();
    st_release_sys(k.flags[tid] + kFlagA + k.rank, k.epoch);
  }
  if (tid == 0) wait_peers(k, kFlagA);
  __syncthreads();
  // ---- 1-3: chunks of kThreads * COLS float4 columns of the owned shard, COLS per thread
  // (CTA-strided: coalesced).  The grid is a few CTAs per SM that WALK the chunks: the flag
  // barrier at the head and the system fence + signal at the tail are paid once per CTA, not
  // once per 32 KB (one CTA per chunk measured 0.61 ms for 285 MB on two GPUs, this form 0.53 ms).
  const long long n_chunks = (k.n4 + (long long)kThreads * COLS - 1) / ((long long)kThreads * COLS);
  float4 Gnext[COLS];
  if (MC && PF && blockIdx.x < n_chunks) {
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const long long j = (long long)blockIdx.x * (kThreads * COLS) + tid + (long long)c * kThreads;
      if (j < k.n4) Gnext[c] = multimem_ld_reduce4(k.mc_grads, k.lo4 + j);
    }
  }
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
  const long long j0 = chunk * (kThreads * COLS) + tid;
  float4 G[COLS];
  if (MC && PF) {
#pragma unroll
    for (int c = 0; c < COLS; ++c) G[c] = Gnext[c];
    const long long nxt = chunk + gridDim.x;
    if (nxt < n_chunks) {
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const long long j = nxt * (kThreads * COLS) + tid + (long long)c * kThreads;
        if (j < k.n4) Gnext[c] = multimem_ld_reduce4(k.mc_grads, k.lo4 + j);
      }
    }
  } else if (MC) {
    // the switch sums the `world` gradient buckets (order fixed by the fabric, not by rank)
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const long long j = j0 + (long long)c * kThreads;
      if (j < k.n4) G[c] = multimem_ld_reduce4(k.mc_grads, k.lo4 + j);
    }
  } else {
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const long long j = j0 + (long long)c * kThreads;
      if (j < k.n4) G[c] = ld_peer4(k.grads[0], k.lo4 + j);
    }
#pragma unroll
    for (int r = 1; r < WORLD; ++r) {
      float4 H[COLS];
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const long long j = j0 + (long long)c * kThreads;
        if (j < k.n4) H[c] = ld_peer4(k.grads[r], k.lo4 + j);
      }
#pragma unroll
      for (int c = 0; c < COLS; ++c) {                    // rank order: deterministic
        G[c].x += H[c].x; G[c].y += H[c].y; G[c].z += H[c].z; G[c].w += H[c].w;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    const long long j = j0 + (long long)c * kThreads;
    if (j >= k.n4) continue;
    const long long i = k.lo4 + j;
    float4 P = reinterpret_cast<float4*>(k.params[k.rank])[i];
    float4 M = reinterpret_cast<float4*>(k.mu)[i];
    float4 V = reinterpret_cast<float4*>(k.nu)[i];
    float4 E = reinterpret_cast<float4*>(k.ema)[i];
    const bool decay = i < k.decay4;
    adamw_elem(P.x, G[c].x, M.x, V.x, E.x, k, decay);
    adamw_elem(P.y, G[c].y, M.y, V.y, E.y, k, decay);
    adamw_elem(P.z, G[c].z, M.z, V.z, E.z, k, decay);
    adamw_elem(P.w, G[c].w, M.w, V.w, E.w, k, decay);
    reinterpret_cast<float4*>(k.mu)[i] = M;
    reinterpret_cast<float4*>(k.nu)[i] = V;
    reinterpret_cast<float4*>(k.ema)[i] = E;
    if (MC) {
      multimem_st4(k.mc_params, i, P);                    // every rank's copy, one store
    } else {
#pragma unroll
      for (int r = 0; r < WORLD; ++r) reinterpret_cast<float4*>(k.params[r])[i] = P;
    }
  }
  }
  // ---- B: my shard has landed everywhere; retire only when every peer's has landed here.
  // ONE system-scope fence per CTA, by thread 0 behind the CTA barrier (the barrier makes every
  // thread's peer stores happen-before it, and the fence is cumulative): a fence in every thread
  // made each of the SM's 2048 threads wait for its own NVLink acknowledgements.
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    const unsigned done = atomicAdd(k.flags[k.rank] + kFlagCount, 1u);
    s_last = (done == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last == 0) return;
  if (tid < WORLD) {
    __threadfence_system();
    st_release_sys(k.flags[tid] + kFlagB + k.rank, k.epoch);
  }
  if (tid == 0) {
    k.flags[k.rank][kFlagCount] = 0;
    wait_peers(k, kFlagB);
  }
}

}  // namespace
}  // namespace mulan

extern "C" {

#define PEER_FAIL(code, ...)                                  \
  do {                                                        \
    char m_[256];                                             \
    snprintf(m_, sizeof(m_), __VA_ARGS__);                    \
    mulan::set_last_error(m_);                                \
    return (int)(code);                                       \
  } while (0)
#define PEER_CU(call, fn)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) PEER_FAIL(MULAN_ERR_CUDA, "%s: %s", fn, cudaGetErrorString(e_)); \
  } while (0)

int mulan_peer_alloc(size_t bytes, void** dev_ptr, void* handle_out) {
  const char* fn = "mulan_peer_alloc";
  if (dev_ptr == nullptr || handle_out == nullptr || bytes == 0)
    PEER_FAIL(MULAN_ERR_INVALID_ARG, "%s: bad argument", fn);
  static_assert(sizeof(cudaIpcMemHandle_t) == MULAN_PEER_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  PEER_CU(cudaMalloc(&p, bytes), fn);
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    PEER_FAIL(MULAN_ERR_CUDA, "%s: %s", fn, cudaGetErrorString(e));
  }
  memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return 0;
}

int mulan_peer_open(const void* handle, void** dev_ptr) {
  const char* fn = "mulan_peer_open";
  if (handle == nullptr || dev_ptr == nullptr) PEER_FAIL(MULAN_ERR_INVALID_ARG, "%s: NULL", fn);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  PEER_CU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess), fn);
  return 0;
}

int mulan_peer_close(void* dev_ptr) {
  if (dev_ptr != nullptr) PEER_CU(cudaIpcCloseMemHandle(dev_ptr), "mulan_peer_close");
  return 0;
}

int mulan_peer_free(void* dev_ptr) {
  if (dev_ptr != nullptr) PEER_CU(cudaFree(dev_ptr), "mulan_peer_free");
  return 0;
}

int mulan_adamw_ema_peer(const mulan_adamw_desc* d, const mulan_peer_desc* peers, int64_t lo,
                         int64_t hi, float* mu, float* nu, float* ema_params, void* stream) {
  const char* fn = "mulan_adamw_ema_peer";
  if (d == nullptr || peers == nullptr) PEER_FAIL(MULAN_ERR_INVALID_ARG, "%s: NULL desc", fn);
  const int W = peers->world;
  if (W < 1 || W > mulan::kMaxPeers || (W & (W - 1)) != 0 || peers->rank < 0 || peers->rank >= W)
    PEER_FAIL(MULAN_ERR_INVALID_ARG, "%s: world=%d (1, 2, 4 or 8), rank=%d", fn, W, peers->rank);
  if (lo < 0 || hi < lo || hi > d->n || lo % 4 != 0 || hi % 4 != 0 || d->n_decay % 4 != 0)
    PEER_FAIL(MULAN_ERR_INVALID_ARG, "%s: need 0 <= lo <= hi <= n, multiples of 4", fn);
  if (d->step < 1) PEER_FAIL(MULAN_ERR_INVALID_ARG, "%s: step must be >= 1", fn);
  if (d->clip_norm > 0.0)
    PEER_FAIL(MULAN_ERR_UNSUPPORTED, "%s: clip_by_global_norm needs the norm of the REDUCED "
              "gradient before any update; use the all-reduce path", fn);
  if (mu == nullptr || nu == nullptr || ema_params == nullptr)
    PEER_FAIL(MULAN_ERR_INVALID_ARG, "%s: NULL state pointer", fn);
  mulan::PeerParams k;
  for (int r = 0; r < W; ++r) {
    if (!peers->grads[r] || !peers->params[r] || !peers->flags[r])
      PEER_FAIL(MULAN_ERR_INVALID_ARG, "%s: peer %d has a NULL mapping", fn, r);
    k.grads[r] = peers->grads[r]; k.params[r] = peers->params[r]; k.flags[r] = peers->flags[r];
  }
  k.mu = mu; k.nu = nu; k.ema = ema_params;
  k.mc_grads = peers->mc_grads; k.mc_params = peers->mc_params;
  const bool mc = k.mc_grads != nullptr && k.mc_params != nullptr && W > 1;
  k.world = W; k.rank = peers->rank; k.epoch = peers->epoch;
  // the range is cut into `world` shards of whole float4 columns; the last takes the remainder
  const long long cols = (hi - lo) / 4, per = (cols + W - 1) / W;
  const long long first = per * k.rank < cols ? per * k.rank : cols;
  const long long last = first + per < cols ? first + per : cols;
  k.lo4 = lo / 4 + first; k.n4 = last - first;
  k.decay4 = d->n_decay / 4;
  k.lr = (float)d->lr; k.b1 = (float)d->b1; k.b2 = (float)d->b2; k.eps = (float)d->eps;
  k.wd = (float)d->weight_decay;
  k.om_b1 = (float)(1.0 - d->b1); k.om_b2 = (float)(1.0 - d->b2);
  k.one_minus_ema = (float)(1.0 - d->ema_rate);
  k.bc1 = 1.0f - powf((float)d->b1, (float)d->step);
  k.bc2 = 1.0f - powf((float)d->b2, (float)d->step);
  k.grad_scale = (float)d->grad_scale;
  const int per_thread = mc ? 4 : (W <= 2 ? 8 : (W == 4 ? 2 : 1));
  long long want = (k.n4 + (long long)mulan::kThreads * per_thread - 1) /
                   ((long long)mulan::kThreads * per_thread);
  if (want < 1) want = 1;                       // an empty shard still takes part in the barriers
  {
    // a few CTAs per SM walk the chunks (see the kernel)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 4;
    if (const char* e = getenv("MULAN_PEER_CTAS_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : 4;   // A/B
    if (want > (long long)per_sm * sms) want = (long long)per_sm * sms;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (mc) {
    bool pf = false;
    if (const char* e = getenv("MULAN_PEER_MC_PREFETCH")) pf = atoi(e) != 0;                  // A/B
    if (pf) {
      switch (W) {
        case 2: mulan::adamw_ema_peer_kernel<2, 4, true, true><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
        case 4: mulan::adamw_ema_peer_kernel<4, 4, true, true><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
        default: mulan::adamw_ema_peer_kernel<8, 4, true, true><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
      }
    } else {
      switch (W) {
        case 2: mulan::adamw_ema_peer_kernel<2, 4, true><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
        case 4: mulan::adamw_ema_peer_kernel<4, 4, true><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
        default: mulan::adamw_ema_peer_kernel<8, 4, true><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
      }
    }
    PEER_CU(cudaGetLastError(), fn);
    return 0;
  }
  switch (W) {
    case 1: mulan::adamw_ema_peer_kernel<1, 8><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
    case 2: mulan::adamw_ema_peer_kernel<2, 8><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
    case 4: mulan::adamw_ema_peer_kernel<4, 2><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
    default: mulan::adamw_ema_peer_kernel<8, 1><<<(int)want, mulan::kThreads, 0, s>>>(k); break;
  }
  PEER_CU(cudaGetLastError(), fn);
  return 0;
}

}  // extern "C"
