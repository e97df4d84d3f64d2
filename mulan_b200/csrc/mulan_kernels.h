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

struct FwdPreParams {
  const uint8_t* x;
  const float *a, *b, *c, *t, *eps0, *eps;
  float *z_t, *g_net, *w_save, *loss_recon, *loss_klz, *var_sums;
  int rows, dim4, gt_mode;
  int W;              // reconstruction window half-width for gamma_0 = gamma_min
  float gmin, delta;  // f32(gamma_min), f32(gamma_max - gamma_min)
  VocabInfo vi;
};

struct PostParams {
  const uint8_t* x;
  const float *a, *b, *c, *t, *eps, *net, *w_save, *gL;
  float* loss_diff;   // fwd
  float* n_bar;       // bwd
  int rows, dim4, param;
  float gmin, delta;
  float scale;        // 0.5 (continuous) or 0.5*T (discrete)
  VocabInfo vi;
};

struct BwdPreParams {
  const uint8_t* x;
  const float *a, *b, *c, *t, *eps, *net, *z_bar, *g_bar, *gL;
  float *a_bar, *b_bar, *c_bar;
  int rows, dim4, param, gt_mode;
  float gmin, delta;
  VocabInfo vi;
};

cudaError_t launch_fwd_pre(const FwdPreParams& p, cudaStream_t s);
cudaError_t launch_fwd_post(const PostParams& p, cudaStream_t s);
cudaError_t launch_bwd_post(const PostParams& p, cudaStream_t s);
cudaError_t launch_bwd_pre(const BwdPreParams& p, cudaStream_t s);
cudaError_t launch_aux_topk_fwd(int rows, int latent, int k, const float* logits,
                                const float* gamma_draw, float* embedding, float* kl_z,
                                cudaStream_t s);
cudaError_t launch_aux_topk_bwd(int rows, int latent, int k, const float* logits,
                                const float* gamma_draw, const float* emb_bar,
                                const float* klz_bar, float* logits_bar, cudaStream_t s);
cudaError_t launch_bpd_reduce(int rows, int dim, const float* loss_recon,
                              const float* loss_klz_prior, const float* kl_z,
                              const float* loss_diff, const float* var_sums, float* scalars,
                              float* loss_klz_total, cudaStream_t s);

}  // namespace mulan
