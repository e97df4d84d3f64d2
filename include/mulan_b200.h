/*
 * mulan_b200.h -- C ABI of libmulan_b200.so: the MuLAN per-pixel noise schedule +
 * variational-diffusion ELBO hot path as hand-written sm_100a CUDA kernels.
 *
 * The reference (s-sahoo/MuLAN, pure JAX/Flax) has no plugin/FFI registry; the seams this
 * library replaces are the Flax method bodies listed per entry point below (paths relative
 * to the reference root).  A JAX binding calls these from XLA-FFI handlers, a PyTorch
 * binding from torch.autograd.Function.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no framework types.
 *   - Unless an entry point says "host", every pointer is a DEVICE pointer owned by the
 *     caller; the call only ENQUEUES work on `stream` (a cudaStream_t passed as void*),
 *     never allocates, never synchronises, keeps no global mutable state, and is
 *     re-entrant from one host thread per device.
 *   - Arrays are float32, row-major [rows, dim] ("[B,D]", D = 32*32*3 = 3072 sub-pixels,
 *     NHWC-flattened) unless noted; x is uint8 [B,D]; per-example vectors are [B].
 *   - [B,D] float pointers must be 16-byte aligned and dim % 4 == 0 (float4 access);
 *     x must be 4-byte aligned.
 *   - Return value: 0 on success, negative mulan_status otherwise; a message for the last
 *     failure on this thread is available from mulan_last_error().
 *   - NaN/Inf propagate as in the reference (no trapping).
 */
#ifndef MULAN_B200_H_
#define MULAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MULAN_ABI_VERSION 2   /* 2: mulan_desc gained flags + noise_rows; mulan_bpd_reduce a workspace */

typedef enum mulan_status {
  MULAN_OK = 0,
  MULAN_ERR_INVALID_ARG = -1,   /* null pointer, bad enum, bad shape */
  MULAN_ERR_ALIGNMENT = -2,     /* pointer / dim not aligned for vector access */
  MULAN_ERR_UNSUPPORTED = -3,   /* combination not implemented (message says which) */
  MULAN_ERR_CUDA = -4           /* launch or runtime error; message has cudaGetErrorString */
} mulan_status;

/* Which network parameterisation the diffusion loss uses. */
typedef enum mulan_param {
  MULAN_PARAM_EPS = 0,          /* ldm/model_mulan_epsilon.py:335-347                     */
  MULAN_PARAM_VEL = 1,          /* ldm/model_mulan_velocity.py:243-260, from_epsilon=False */
  MULAN_PARAM_VEL_FROM_EPS = 2  /* ldm/model_mulan_velocity.py:246-249, from_epsilon=True  */
} mulan_param;

/* What the denoiser receives as its noise-level input (_get_score_model_gt,
 * ldm/model_mulan_epsilon.py:273-278). */
typedef enum mulan_gt_mode {
  MULAN_GT_MEAN = 0,            /* unet_type='vdm': per-example mean of gamma_t, [B]   */
  MULAN_GT_PIXEL = 1            /* unet_type='ldm': per-pixel gamma_t, [B,D]           */
} mulan_gt_mode;

/* mulan_desc.flags */
typedef enum mulan_flags {
  /* The `c` argument of mulan_fwd_pre / mulan_fwd_post / mulan_bwd_post / mulan_fwd_bwd_post /
   * mulan_post_bpd / mulan_bwd_pre is the PRE-ACTIVATION r of dense_out_c: the kernels apply the
   * epilogue of _compute_coefficients, c = 1e-3 + softplus(r) (ldm/model_mulan_epsilon.py:537),
   * themselves, and mulan_bwd_pre returns c_bar as the cotangent of r (c_bar * sigmoid(r)) --
   * the framework's softplus forward (8 B/sub-pixel) and backward (12 B/sub-pixel) disappear.
   * Power-of-two vocabularies only. */
  MULAN_FLAG_C_RAW = 1,
  /* Launch with programmatic stream serialization (cudaLaunchAttributeProgrammaticStream-
   * Serialization): the kernel's CTAs may become resident while the previous kernel on the
   * stream drains; every kernel of the path waits (griddepcontrol.wait) for that kernel to
   * complete before its first global-memory access, so results are unchanged.  Worth ~1-2 us per
   * launch boundary: the literal per-GPU batch of 128 rows is launch-latency bound. */
  MULAN_FLAG_PDL = 2
} mulan_flags;

/* POD descriptor shared by all entry points (the subset of VDMConfig,
 * ldm/model_vdm.py:33-82, the path reads). gamma_min/max are doubles because the
 * reference forms gamma_max-gamma_min in Python double before the float32 cast
 * (ldm/model_mulan_epsilon.py:489-490). */
typedef struct mulan_desc {
  int32_t rows;         /* B: examples in this launch (this device's shard)      */
  int32_t dim;          /* D: sub-pixels per example (3072); multiple of 4       */
  int32_t vocab;        /* vocab_size (256); 2..65536                            */
  int32_t param;        /* mulan_param                                           */
  int32_t gt_mode;      /* mulan_gt_mode                                         */
  int32_t n_timesteps;  /* sm_n_timesteps T; 0 = continuous time                 */
  double gamma_min;     /* -13.3                                                 */
  double gamma_max;     /*  5.0                                                  */
  uint32_t flags;       /* mulan_flags, OR-ed; 0 = the plain v1 behaviour        */
  /* 0: eps0 / eps are [rows, dim].  N > 0: they are [N, dim] and row b reads row b % N -- the
   * dense-VLB evaluation tiles ONE key's draws over every image of a launch
   * (ldm/notebook_utils.py:178, :185: the same PRNGKey(0) for each image), so 16 images x 128
   * timesteps read 128 noise rows, not 2048.  Applies to mulan_fwd_pre (eps0, eps) and to eps
   * in the post / bwd_pre entry points. */
  int32_t noise_rows;
} mulan_desc;

/* Thread-local message describing the last non-zero status returned on this thread. */
const char* mulan_last_error(void);
int mulan_abi_version(void);

/*
 * mulan_kernel_param -- the mulan_param whose loss formula mulan_fwd_post / mulan_bwd_post /
 * mulan_fwd_bwd_post / mulan_bwd_pre run for a descriptor with this `param`.  Identity except
 * MULAN_PARAM_VEL_FROM_EPS -> MULAN_PARAM_EPS: the velocity_from_epsilon loss
 * (ldm/model_mulan_velocity.py:246-249, 256-260) is algebraically the epsilon loss
 * (ldm/model_mulan_epsilon.py:345-347) -- (1-v) w (v_target - v_hat)^2 == w (eps - net)^2 --
 * with identical cotangents, so those entry points evaluate the epsilon form (1.3e-7 / 3e-6
 * relative from the reference's float32 loss / gradients on its own golden vectors, 1e-15 in
 * float64).  A caller uses it to decide whether a w_save buffer pays (it does when the result
 * is MULAN_PARAM_EPS) and whether x is read by the post kernels (it is not, then).
 * MULAN_VFE_LITERAL=1 in the environment keeps the literal formula (identity mapping).
 */
int mulan_kernel_param(int32_t param);

/*
 * mulan_fwd_pre -- everything in VDM.__call__ that precedes the denoiser call.
 * Replaces: EncDec.encode (ldm/model_vdm.py:274-280); NoiseSchedule_polynomial_fixedend.
 * _eval_polynomial at t in {0,1,t} (ldm/model_mulan_epsilon.py:514-529, call sites :307-309);
 * the jvp d-gamma/dt (:339-343); noising (:311,:327-328); _get_score_model_gt (:273-278);
 * reconstruction term (:315-318 -> EncDec.logprob/decode, ldm/model_vdm.py:282-303);
 * prior KL (:322-325); var_0/var_1 partial sums (:361-362).
 * The velocity model runs the same statements (ldm/model_mulan_velocity.py:208-236).
 *
 *   in : x[B,D] u8; a,b,c[B,D] polynomial coefficients (c already 1e-3+softplus);
 *        t[B]; eps0[B,D]; eps[B,D]
 *   out: z_t[B,D]; g_net ([B] for GT_MEAN, [B,D] for GT_PIXEL);
 *        w_save[B,D] or NULL -- the loss weight d-gamma/dt (T==0) or expm1(g_t-g_s) with
 *        s = t - 1/T (T>0: epsilon model only, w_save REQUIRED, t already discretised by the
 *        caller as ceil(t*T)/T), consumed by mulan_fwd_post / mulan_bwd_post;
 *        loss_recon[B]; loss_klz_prior[B]; var_sums[B,2] = per-row sum of sigmoid(g_0),
 *        sigmoid(g_1).
 */
int mulan_fwd_pre(const mulan_desc* desc,
                  const uint8_t* x, const float* a, const float* b, const float* c,
                  const float* t, const float* eps0, const float* eps,
                  float* z_t, float* g_net, float* w_save,
                  float* loss_recon, float* loss_klz_prior, float* var_sums,
                  void* stream);
/*
 * mulan_fwd_pre_consts -- mulan_fwd_pre with the five transcendental constants of the fixed
 * schedule ends SUPPLIED by the caller, as the caller's framework evaluates them in float32,
 * instead of computed on the host (double, rounded once = correctly rounded).  The reference's
 * own loss_recon depends on how its platform rounds them: one ulp of exp(g_0/2) moves a row's
 * loss_recon by up to 1.8e-5 relative, and for exp(-g_0/2) the exact value lies 0.40 ulp from the
 * correctly rounded float, where float32 exp implementations already disagree (DESIGN.md
 * section 2).  A binding that wants to match ITS platform bit for rounding passes
 *   exp_half_g0 = exp(.5 g_0), exp_neg_half_g0 = exp(-.5 g_0), sigmoid_g0, sigmoid_g1 =
 *   sigmoid(g_min + (g_max - g_min)), log_sigmoid_g1 = log(sigmoid_g1)     with g_0 = f32(g_min)
 * evaluated once by its own exp / log (ldm/model_mulan_epsilon.py:311-325; model_vdm.py:286).
 * consts == NULL is exactly mulan_fwd_pre.  The kernels read these values from their parameter
 * block either way (the immediates-specialised kernel is selected only when they equal the
 * shipped configuration's host-computed values bit for bit).
 * STATUS: added at the end of round 1 after the GPU budget was spent -- the plumbing is covered
 * by CPU tests (mulan_fwd_pre_variant_consts), the numerics run through the same, GPU-tested,
 * parameter-bank kernels, but the override itself has not yet been exercised on a GPU.
 */
typedef struct mulan_end_consts {
  float exp_half_g0, exp_neg_half_g0, sigmoid_g0, sigmoid_g1, log_sigmoid_g1;
} mulan_end_consts;
int mulan_fwd_pre_consts(const mulan_desc* desc, const mulan_end_consts* consts,
                         const uint8_t* x, const float* a, const float* b, const float* c,
                         const float* t, const float* eps0, const float* eps,
                         float* z_t, float* g_net, float* w_save,
                         float* loss_recon, float* loss_klz_prior, float* var_sums,
                         void* stream);
/* Host-only: what mulan_end_consts the library itself would use for this descriptor (out), and
 * which kernel variant (see mulan_fwd_pre_variant) a call with `consts` (NULL = own) selects. */
int mulan_host_end_consts(const mulan_desc* desc, mulan_end_consts* out);
int mulan_fwd_pre_variant_consts(const mulan_desc* desc, const mulan_end_consts* consts);

/*
 * mulan_fwd_pre_keyed -- mulan_fwd_pre with eps_0 and eps DRAWN INSIDE THE KERNEL from the raw
 * 2 x uint32 threefry keys that make_rng('sample') hands to jax.random.normal(rng, f.shape)
 * (ldm/model_mulan_epsilon.py:315, :327; "next" row 2 of the scope table as written).  The draws
 * equal mulan_rng_normal(key, rows * dim) reshaped to [rows, dim], bit for bit, and every output
 * equals mulan_fwd_pre's on those arrays.  eps_out [B,D] (eps is read again by the post /
 * bwd_pre entry points) and eps0_out [B,D] are optional (NULL: not written).
 * rows must be even (JAX pairs element e with e + N/2, so a CTA serves the row pair
 * (r, r + rows/2)); continuous time; the shipped configurations' closed-form reconstruction term.
 * 17 (+4 eps_out, +4 w_save) B/sub-pixel of HBM traffic instead of 25, but ~300 instead of ~105
 * instructions per sub-pixel: issue bound.  It replaces TWO stand-alone draws (8 B written, 8 B
 * read back) plus mulan_fwd_pre at about the same total time (profiles/r2_variants.md); against
 * draws that already sit in HBM, mulan_fwd_pre is faster.
 */
int mulan_fwd_pre_keyed(const mulan_desc* desc, const uint32_t* key_eps0, const uint32_t* key_eps,
                        const uint8_t* x, const float* a, const float* b, const float* c,
                        const float* t, float* z_t, float* g_net, float* w_save, float* eps0_out,
                        float* eps_out, float* loss_recon, float* loss_klz_prior, float* var_sums,
                        void* stream);

/* Host-only query: which mulan_fwd_pre kernel this descriptor selects on this host --
 * 0 generic (windowed log-softmax over the vocab bins), 1 closed-form 3-bin reconstruction term
 * with the launch constants in the parameter bank, 2 the same with the constants of the shipped
 * configuration (gamma in [-13.3, 5], vocab 256: both files under ldm/configs) as instruction
 * immediates (selected only when the host-computed constants match them bit for bit).  All three
 * produce the same outputs; negative = invalid descriptor. */
int mulan_fwd_pre_variant(const mulan_desc* desc);

/*
 * mulan_fwd_post -- the diffusion loss after the denoiser returned `net`.
 * Replaces ldm/model_mulan_epsilon.py:338-355 (EPS; T>0 variant needs w_save) and
 * ldm/model_mulan_velocity.py:246-260 (VEL / VEL_FROM_EPS).
 *   in : x, a, b, c, t, eps as for fwd_pre; net[B,D]; w_save[B,D] or NULL (recomputed
 *        from a,b,c,t when NULL; EPS with w_save never touches x,a,b,c)
 *   out: loss_diff[B]
 */
int mulan_fwd_post(const mulan_desc* desc,
                   const uint8_t* x, const float* a, const float* b, const float* c,
                   const float* t, const float* eps, const float* net, const float* w_save,
                   float* loss_diff, void* stream);

/*
 * mulan_bwd_post -- cotangent of the denoiser output, what jax.value_and_grad
 * (ldm/experiment.py:339) pushes into the U-Net: n_bar = d(sum_b gL_b loss_diff_b)/d net.
 *   in : as fwd_post, gL[B] = upstream cotangent of loss_diff
 *   out: n_bar[B,D]
 */
int mulan_bwd_post(const mulan_desc* desc,
                   const uint8_t* x, const float* a, const float* b, const float* c,
                   const float* t, const float* eps, const float* net, const float* w_save,
                   const float* gL, float* n_bar, void* stream);

/*
 * mulan_fwd_bwd_post -- value-and-grad of the diffusion loss in ONE pass: loss_diff[B] and
 * n_bar[B,D] = gL_b * d loss_diff_b / d net, for callers that know the loss cotangent gL up
 * front (jax.value_and_grad of a mean: gL_b = 1/(B*D*ln 2), ldm/experiment_vdm.py:62-66).
 * Reads what mulan_fwd_post reads once instead of twice (EPS: 12 B + 4 B written per
 * sub-pixel instead of 12 + 16).  If the true cotangent turns out different,
 * mulan_scale_rows(n_bar, true, assumed) corrects n_bar in place, touching only the rows
 * that differ.
 */
int mulan_fwd_bwd_post(const mulan_desc* desc,
                       const uint8_t* x, const float* a, const float* b, const float* c,
                       const float* t, const float* eps, const float* net, const float* w_save,
                       const float* gL, float* loss_diff, float* n_bar, void* stream);

/* v[b, :] *= num[b] / den[b] for the rows where num[b] != den[b]; v is [rows, dim]. */
int mulan_scale_rows(int32_t rows, int32_t dim, float* v, const float* num, const float* den,
                     void* stream);

/*
 * mulan_bwd_pre -- cotangents of the polynomial coefficients: every path from the loss to
 * (a,b,c): through z_t (z_bar, returned by the denoiser's backward), through the
 * denoiser's noise-level input (g_bar: [B] for GT_MEAN, [B,D] for GT_PIXEL), and through
 * loss_diff directly (weight d-gamma/dt, and for the velocity models (1-var_t), v_target,
 * and v_hat's own dependence on gamma_t and z_t).  gamma(0), gamma(1) are fixed ends and
 * contribute nothing.  Any of z_bar, g_bar, gL may be NULL (treated as zero); net may be
 * NULL only if gL is NULL.
 *   out: a_bar, b_bar, c_bar [B,D]
 */
int mulan_bwd_pre(const mulan_desc* desc,
                  const uint8_t* x, const float* a, const float* b, const float* c,
                  const float* t, const float* eps, const float* net,
                  const float* z_bar, const float* g_bar, const float* gL,
                  float* a_bar, float* b_bar, float* c_bar, void* stream);

/*
 * mulan_aux_topk_fwd / _bwd -- auxiliary-latent KL and relaxed top-k embedding.
 * Replaces _gumbel_kl_loss, _gamma_noise, _topk_embedding_and_loss
 * (ldm/model_mulan_epsilon.py:205-210, 221-252; ldm/model_mulan_velocity.py:78-120).
 *   in : logits[B,L]; gamma_draw[10,B,L] ~ Gamma(1/k) (the jax.random.gamma draw) or NULL
 *        for no noise; L <= 64; k = latent_k
 *   out: embedding[B,L]; kl_z[B]
 *   bwd: logits_bar[B,L] = d(sum emb_bar*embedding + sum klz_bar*kl_z)/d logits
 */
int mulan_aux_topk_fwd(int32_t rows, int32_t latent, int32_t k,
                       const float* logits, const float* gamma_draw,
                       float* embedding, float* kl_z, void* stream);
int mulan_aux_topk_bwd(int32_t rows, int32_t latent, int32_t k,
                       const float* logits, const float* gamma_draw,
                       const float* emb_bar, const float* klz_bar,
                       float* logits_bar, void* stream);

/*
 * The other auxiliary-latent variants of _get_embedding_and_kl_z
 * (ldm/model_mulan_epsilon.py:257-271; neither shipped config selects them):
 *   mulan_aux_topk_add_*  : top-k with an ADDITIVE noise [B,L] instead of the gamma draw
 *                           (topk_noise_type == 'gumbel', :238-239)
 *   mulan_aux_gumbel_*    : latent_type == 'gumbel' (:195-219): l = (logits + gumbel)/tau,
 *                           emb = stop_grad(one_hot(argmax l) - softmax l) + softmax l,
 *                           kl_z = KL(softmax(logits) || uniform);
 *                           tau = max(.5, exp(-1e-5 step)) is computed by the caller
 *   mulan_aux_gaussian_*  : latent_type == 'gaussian' (:264-270): emb = mu + sqrt(var) eps_z,
 *                           kl_z = .5 sum(mu^2 + var - log var - 1); bwd -> mu_bar, var_bar
 * All arrays are [B,L] (kl_z, klz_bar: [B]); L <= 64; noise / bar pointers may be NULL (zero).
 */
int mulan_aux_topk_add_fwd(int32_t rows, int32_t latent, int32_t k, const float* logits,
                           const float* noise, float* embedding, float* kl_z, void* stream);
int mulan_aux_topk_add_bwd(int32_t rows, int32_t latent, int32_t k, const float* logits,
                           const float* noise, const float* emb_bar, const float* klz_bar,
                           float* logits_bar, void* stream);
int mulan_aux_gumbel_fwd(int32_t rows, int32_t latent, double tau, const float* logits,
                         const float* gumbel_noise, float* embedding, float* kl_z, void* stream);
int mulan_aux_gumbel_bwd(int32_t rows, int32_t latent, double tau, const float* logits,
                         const float* gumbel_noise, const float* emb_bar, const float* klz_bar,
                         float* logits_bar, void* stream);
int mulan_aux_gaussian_fwd(int32_t rows, int32_t latent, const float* mu, const float* var,
                           const float* eps_z, float* embedding, float* kl_z, void* stream);
int mulan_aux_gaussian_bwd(int32_t rows, int32_t latent, const float* mu, const float* var,
                           const float* eps_z, const float* emb_bar, const float* klz_bar,
                           float* mu_bar, float* var_bar, void* stream);

/*
 * mulan_bpd_reduce -- VDMOutput assembly + Experiment_VDM.loss_fn scalars
 * (ldm/model_mulan_epsilon.py:357-363, ldm/experiment_vdm.py:62-74).
 *   in : loss_recon[B], loss_klz_prior[B], kl_z[B] or NULL, loss_diff[B], var_sums[B,2]
 *   out: scalars[6] = {bpd, bpd_latent, bpd_recon, bpd_diff, var0, var1};
 *        loss_klz_total[B] or NULL = kl_z + loss_klz_prior
 * The sums run in a FIXED order (groups of 128 rows, then the group partials; csrc/
 * mulan_reduce.cuh), so the result is deterministic and identical for the three ways of
 * obtaining it: this entry with reduce_ws (one CTA per group, last CTA finalises), this entry
 * with reduce_ws == NULL (one CTA walks the groups: slow, for callers that cannot supply
 * zeroed scratch) and mulan_post_bpd (fused into the post kernel).
 *   reduce_ws: mulan_reduce_ws_bytes(rows) bytes of device memory, 16-byte aligned, ZERO before
 *   its first use; every call leaves it zero again, so consecutive calls on one stream can share
 *   it.  Not shareable between calls that may run concurrently.
 */
size_t mulan_reduce_ws_bytes(int32_t rows);
int mulan_bpd_reduce(const mulan_desc* desc,
                     const float* loss_recon, const float* loss_klz_prior, const float* kl_z,
                     const float* loss_diff, const float* var_sums,
                     float* scalars, float* loss_klz_total, void* reduce_ws, void* stream);

/*
 * mulan_post_bpd -- mulan_fwd_post (gL == NULL: loss_diff only) or mulan_fwd_bwd_post (gL given:
 * loss_diff and n_bar) AND mulan_bpd_reduce in ONE launch: the CTA that completes a group of 128
 * rows folds the group's loss terms, the CTA that completes the last group writes the six
 * scalars.  Same bits as the separate calls; one launch and a single-CTA tail less per step
 * (ldm/experiment_vdm.py:62-74 follows ldm/model_mulan_epsilon.py:345-363 directly).
 * loss_recon, loss_klz_prior, var_sums are the outputs of mulan_fwd_pre for the same rows.
 */
int mulan_post_bpd(const mulan_desc* desc,
                   const uint8_t* x, const float* a, const float* b, const float* c,
                   const float* t, const float* eps, const float* net, const float* w_save,
                   const float* gL, const float* loss_recon, const float* loss_klz_prior,
                   const float* kl_z, const float* var_sums,
                   float* loss_diff, float* n_bar, float* scalars, float* loss_klz_total,
                   void* reduce_ws, void* stream);

/*
 * mulan_elbo_host -- HOST-buffer convenience entry: one call = H2D of the inputs, fwd_pre,
 * the denoiser callback (or the supplied `net`), fwd_post, bwd_post, bwd_pre, bpd_reduce and
 * D2H of the results, on the current device.  This is the call a ctypes/cgo-style binding
 * with host arrays would make; it allocates a workspace internally (cached per thread) and
 * synchronises before returning.  The batch flows through copy-in / compute / copy-out
 * streams in row chunks, so transfers in both directions overlap each other and the
 * kernels; page-locked host buffers make the copies true DMA.
 *   denoiser(user, rows, z_t_dev, g_net_dev, net_dev, stream): fills net_dev[rows,D] for
 *   one chunk of `rows` examples (called once per chunk, in row order); may be NULL, in
 *   which case `net` (host, [B,D]) is used.  z_bar is taken as zero (the path through the
 *   denoiser's own backward is the caller's).
 *   out (host): losses[3*B] = recon | klz_prior | diff; scalars[6];
 *               grads (want_grad): a_bar,b_bar,c_bar,n_bar [B,D] each (NULL to skip copy-out)
 */
typedef int (*mulan_denoiser_fn)(void* user, int32_t rows, const float* z_t, const float* g_net,
                                 float* net, void* stream);
int mulan_elbo_host(const mulan_desc* desc,
                    const uint8_t* x, const float* a, const float* b, const float* c,
                    const float* t, const float* eps0, const float* eps, const float* net,
                    mulan_denoiser_fn denoiser, void* user, int32_t want_grad,
                    float* losses, float* scalars,
                    float* a_bar, float* b_bar, float* c_bar, float* n_bar);

/*
 * mulan_elbo_host_keyed -- mulan_elbo_host with eps_0 and eps DRAWN ON THE DEVICE from the raw
 * 2 x uint32 threefry keys that the reference's make_rng('sample') hands to
 * jax.random.normal(rng, shape=f.shape) (ldm/model_mulan_epsilon.py:315, :327): which is what
 * VDM.__call__ itself does -- it receives images and an rng, not noise arrays.  The draws equal
 * mulan_rng_normal(key, B*D) reshaped to [B,D]; 8 of the 25 host-to-device bytes per sub-pixel
 * of mulan_elbo_host never cross PCIe.  rows * dim < 2^32 - 1.
 */
int mulan_elbo_host_keyed(const mulan_desc* desc,
                          const uint8_t* x, const float* a, const float* b, const float* c,
                          const float* t, const uint32_t* key_eps0, const uint32_t* key_eps,
                          const float* net, mulan_denoiser_fn denoiser, void* user,
                          int32_t want_grad, float* losses, float* scalars,
                          float* a_bar, float* b_bar, float* c_bar, float* n_bar);

/*
 * Ancestral sampler ("next" row 3 of the scope table): the schedule math of VDM.sample /
 * conditional_sample (ldm/model_mulan_epsilon.py:377-438, ldm/model_mulan_velocity.py:281-347)
 * and VDM.generate_x (ldm/model_mulan_epsilon.py:440-457), run T = 1000 times per generated
 * batch by Experiment_VDM.sample_fn (ldm/experiment_vdm.py:80-110).
 * a, b, c are [abc_rows, D] with abc_rows == rows, or == 1 to broadcast one coefficient row
 * over the batch (the unconditional sampler's deterministic embedding).
 *   mulan_sample_gamma : g_net = per-row mean (GT_MEAN, [B]) or per-pixel (GT_PIXEL, [B,D])
 *                        gamma(t) -- the denoiser's noise-level input (:397-411, :273-278)
 *   mulan_sample_step  : z_s = sqrt(a/b)(z_t - sigma_t c eps_hat) + sqrt((1-a) c) eps with
 *                        a = sigmoid(-g_s), b = sigmoid(-g_t), c = -expm1(g_s - g_t);
 *                        desc->param != EPS: eps_hat = net*sqrt(b) + sigma_t z_t (velocity model)
 *   mulan_generate_x   : x[B,D] u8 = argmax_k decode(z_0 / sqrt(1 - sigmoid(g_0)), g_0)
 *                        (sample_softmax = False; g_0 = gamma_min, the fixed end)
 */
int mulan_sample_gamma(const mulan_desc* desc, int32_t abc_rows, const float* a, const float* b,
                       const float* c, const float* t, float* g_net, void* stream);
int mulan_sample_step(const mulan_desc* desc, int32_t abc_rows, const float* a, const float* b,
                      const float* c, const float* t, const float* s, const float* z_t,
                      const float* net, const float* eps, float* z_s, void* stream);
int mulan_generate_x(const mulan_desc* desc, const float* z_0, uint8_t* x, void* stream);

/*
 * Probability-flow ODE ("next" row 4): drift of VDM.reverse_ode
 * (ldm/model_mulan_epsilon.py:459-478; the velocity model's version returns nothing) and the
 * local pieces of its Hutchinson divergence (_get_value_div_fn, ldm/notebook_utils.py:204-216).
 *   drift = 0.5 (-sigma x_t + eps_hat) sigma dgamma/dt,  sigma = sqrt(sigmoid(g_t))
 *           (high_precision != 0: sigma = exp(g_t/2) where sigmoid(g_t) <= 1e-3)
 *   with Hutchinson noise v (NULL: drift only), k = 0.5 sigma dgamma/dt:
 *     net_bar[B,D]  = k v                 cotangent for the denoiser's backward
 *     div_direct[B] = sum_d -sigma k v^2  the part of v^T J v that bypasses the denoiser
 *   mulan_row_dot(u = J_net^T net_bar, v, add = div_direct) then completes the divergence.
 */
int mulan_ode_drift(const mulan_desc* desc, int32_t abc_rows, const float* a, const float* b,
                    const float* c, const float* t, const float* x_t, const float* eps_hat,
                    const float* v, int32_t high_precision, float* drift, float* net_bar,
                    float* div_direct, void* stream);
/* out[b] = sum_d u[b,d] v[b,d] (+ add[b] when add != NULL); u, v are [rows, dim]. */
int mulan_row_dot(int32_t rows, int32_t dim, const float* u, const float* v, const float* add,
                  float* out, void* stream);

/*
 * Device-resident state of the adaptive RK45 (Dormand-Prince 5(4)) integrator that drives the
 * probability-flow ODE: replaces the host float64 numpy state of
 * scipy.integrate.solve_ivp(method='RK45') in likelihood_fn / sample_fn
 * (ldm/notebook_utils.py:343-353, :417-429), whose every function evaluation ships the whole
 * state host->device and the derivative back (_from/_to_flattened_numpy, :193-200).
 *   y[n] float64 state, K[7][n] float32 stage derivatives (row j at K + j*k_stride), all in HBM.
 *   mulan_rk45_stage: v = y + (sum_{j<n_k} coef[j] K_j) h;  y_stage = (float)v and/or y_out = v
 *                     (either may be NULL; n_k = 0 is a plain float64 -> float32 cast).
 *   mulan_rk45_norm : out[0] = sum_i (v_i / (atol + rtol max(|y_i|, |y_new_i|)))^2 with
 *                     v = y (of_y != 0) or (sum_j coef[j] K_j) h; y_new may be NULL.
 *                     Deterministic (fixed-order) reduction; scratch holds
 *                     MULAN_RK45_SCRATCH doubles; out is a DEVICE pointer.
 * coef is a HOST array of n_k <= 7 doubles, read before the call returns.
 * Any alignment works; with every vector 16-byte aligned and k_stride % 4 == 0 the kernels move
 * four elements per thread (128-bit loads), bit-identical per element to the scalar form.
 */
#define MULAN_RK45_SCRATCH 2048
int mulan_rk45_stage(int64_t n, int32_t n_k, const double* coef, double h, const double* y,
                     const float* K, int64_t k_stride, float* y_stage, double* y_out,
                     void* stream);
int mulan_rk45_norm(int64_t n, int32_t n_k, const double* coef, double h, double rtol,
                    double atol, const double* y, const double* y_new, const float* K,
                    int64_t k_stride, int32_t of_y, double* scratch, double* out, void* stream);

/*
 * mulan_adamw_ema -- "next" row 1 of the scope table: the AdamW + EMA update that follows the
 * gradient all-reduce of every train step, fused over one flat float32 buffer.
 * Replaces TrainState.apply_gradients (ldm/train_state.py:70-102) with the optax.adamw chain
 * of ldm/experiment.py:132-182 (scale_by_adam -> add_decayed_weights(mask) -> scale(-lr)),
 * the EMA of ldm/train_state.py:91-95, and the 1/world of pmean(grads) (ldm/experiment.py:341)
 * via grad_scale.  Parameters are laid out decayed-first: elements [0, n_decay) receive weight
 * decay (the reference's mask: everything but biases).  n and n_decay must be multiples of 4,
 * pointers 16-byte aligned.  36 B per parameter.
 */
typedef struct mulan_adamw_desc {
  int64_t n;            /* parameters in the flat buffer                    */
  int64_t n_decay;      /* [0, n_decay) get weight decay                    */
  int32_t step;         /* 1-based update count (bias correction)           */
  int32_t reserved;
  /* Hyper-parameters are doubles, as in the reference's Python config: each is rounded to
   * float32 where optax would use it (e.g. (1 - b1) is formed in double first).            */
  double lr;            /* learning rate for THIS step (schedule on the host) */
  double b1, b2, eps, weight_decay;
  double ema_rate;      /* 0.9999: ema += (1 - ema_rate) (p_new - ema)       */
  double grad_scale;    /* g is multiplied by this first (1/world of the pmean) */
  /* optax.clip_by_global_norm(config.gradient_clip_norm), ldm/experiment.py:176-178, applied to
   * grad_scale * g:  0 = off; else grad_sumsq (DEVICE scalar from mulan_grad_sumsq over the raw
   * bucket) must be set and g <- (g / norm) * clip_norm whenever norm >= clip_norm.            */
  double clip_norm;
  const float* grad_sumsq;
} mulan_adamw_desc;

int mulan_adamw_ema(const mulan_adamw_desc* desc, float* params, const float* grads,
                    float* mu, float* nu, float* ema_params, void* stream);

/*
 * Gradient pmean FUSED with the update over NVLink peer memory -- SURVEY.md 8f row 1 as written.
 * Replaces grads = jax.lax.pmean(grads, 'batch') (ldm/experiment.py:341) + state.apply_gradients
 * (ldm/experiment.py:344 -> ldm/train_state.py:70-102) for one process per GPU on one node.
 *
 *   mulan_peer_alloc : cudaMalloc + zero + CUDA-IPC handle (MULAN_PEER_HANDLE_BYTES bytes) of a
 *                      buffer the other ranks will map; mulan_peer_open maps a peer's handle
 *                      (peer access over NVLink is enabled on first use); _close / _free undo them.
 *   mulan_adamw_ema_peer(desc, peers, lo, hi, ...): for the parameter range [lo, hi) of the flat
 *                      buffers, in ONE kernel per rank: flag barrier -> rank r sums the `world`
 *                      gradient buckets over its 1/world shard by peer loads in rank order
 *                      (deterministic) -> AdamW + EMA on that shard (mu, nu, ema local, touched
 *                      for 1/world of the range) -> new parameters stored into every rank's
 *                      buffer -> flag barrier; the kernel retires when every peer's shard has
 *                      landed here.  desc->grad_scale = 1/world makes the sum the pmean;
 *                      desc->n, n_decay describe the WHOLE flat buffer; clip_norm must be 0.
 *                      Every rank must issue the same sequence of calls with the same (lo, hi)
 *                      and a strictly increasing peers->epoch (>= 1).  world in {1, 2, 4, 8}.
 *   flag block       : MULAN_PEER_FLAG_WORDS uint32 per rank, zero-initialised (mulan_peer_alloc
 *                      zeroes); word MULAN_PEER_FLAG_ERR becomes non-zero if a barrier timed out
 *                      (~4 s) instead of hanging the device.
 */
#define MULAN_PEER_HANDLE_BYTES 64
#define MULAN_PEER_MAX 8
#define MULAN_PEER_FLAG_WORDS 32
#define MULAN_PEER_FLAG_ERR 17
typedef struct mulan_peer_desc {
  int32_t world, rank;
  float* grads[MULAN_PEER_MAX];     /* this process's mapping of rank r's gradient bucket  */
  float* params[MULAN_PEER_MAX];    /* ... of rank r's parameter buffer                     */
  uint32_t* flags[MULAN_PEER_MAX];  /* ... of rank r's flag block                           */
  uint32_t epoch;                   /* 1, 2, 3, ...: one per call, identical on every rank  */
  uint32_t reserved;
  /* Optional NVSwitch MULTICAST addresses of the gradient and parameter buffers (both or neither;
   * NULL: plain peer loads / stores).  With them the reduce-scatter is one multimem.ld_reduce per
   * element -- the float32 sum over all ranks formed inside the switch, in the fabric's order
   * rather than rank order -- and the all-gather one multimem.st: per rank the NVLink volume falls
   * from 2 (N-1)/N of the bucket per direction to 1/N in + 1/N out, and the kernel is HBM bound.
   * The buffers must then be symmetric allocations bound to a multicast object (e.g.
   * torch.distributed._symmetric_memory, which also yields the per-rank mappings above). */
  float* mc_grads;
  float* mc_params;
} mulan_peer_desc;
int mulan_peer_alloc(size_t bytes, void** dev_ptr, void* handle_out);
int mulan_peer_open(const void* handle, void** dev_ptr);
int mulan_peer_close(void* dev_ptr);
int mulan_peer_free(void* dev_ptr);
int mulan_adamw_ema_peer(const mulan_adamw_desc* desc, const mulan_peer_desc* peers, int64_t lo,
                         int64_t hi, float* mu, float* nu, float* ema_params, void* stream);

/*
 * pmean of the six loss scalars WITHOUT a collective call (ldm/experiment.py:347-348, 365-366:
 * jax.lax.pmean of each metric right after the ELBO): mulan_post_bpd_peer is mulan_post_bpd whose
 * finalising thread also stores this rank's scalars, tagged with a step counter, into slot
 * [step % 64][rank] of EVERY rank's board over NVLink peer memory (boards allocated with
 * mulan_peer_alloc(mulan_scalar_board_bytes()), mapped with mulan_peer_open) -- an all-gather by
 * peer stores in the kernel's epilogue: no NCCL launch per step, nothing for the next kernel to
 * wait for.  The step counter lives in the rank's own board, so CUDA-graph replays advance it.
 * mulan_scalar_board_read(board, mean_out[6], epoch_out, stream): the mean over the ranks of the
 * latest step this rank has published -- waits (bounded, ~4 s) for every rank's row of that
 * slot, sums in rank order (the same bits on every rank), divides by world; epoch_out = the step
 * (0 and NaNs if a peer had already lapped the ring or the wait timed out).  Called when the host
 * wants the metrics (the reference logs every 1000 steps), not once per step.
 */
typedef struct mulan_scalar_board {
  int32_t world, rank;
  float* boards[MULAN_PEER_MAX];   /* this process's mapping of rank r's board */
} mulan_scalar_board;
size_t mulan_scalar_board_bytes(void);
int mulan_post_bpd_peer(const mulan_desc* desc,
                        const uint8_t* x, const float* a, const float* b, const float* c,
                        const float* t, const float* eps, const float* net, const float* w_save,
                        const float* gL, const float* loss_recon, const float* loss_klz_prior,
                        const float* kl_z, const float* var_sums,
                        float* loss_diff, float* n_bar, float* scalars, float* loss_klz_total,
                        void* reduce_ws, const mulan_scalar_board* board, void* stream);
int mulan_scalar_board_read(const mulan_scalar_board* board, float* mean_out, uint32_t* epoch_out,
                            void* stream);

/* out[0] (device float) = sum_i g[i]^2 over the flat bucket: optax.global_norm(grads)^2 for
 * clip_by_global_norm.  n % 4 == 0, g 16-byte aligned; scratch holds MULAN_SUMSQ_SCRATCH doubles.
 * Fixed-order two-launch reduction (float64 accumulation of float32 squares): run to run
 * identical, 4 B per parameter. */
#define MULAN_SUMSQ_SCRATCH 2048
int mulan_grad_sumsq(int64_t n, const float* g, double* scratch, float* out, void* stream);

/*
 * "Next" row 2: the random draws of VDM.__call__ generated on the device with JAX's default
 * counter-based generator, so that a JAX binding can hand over the 2 x uint32 key of a draw
 * instead of a materialised array: t0 = jax.random.uniform(rng, ()) and
 * eps_0 / eps = jax.random.normal(rng, shape) (ldm/model_mulan_epsilon.py:287-292, :315, :327).
 * out[i], i < n, equals jax.random.{bits,uniform,normal}(key, (n,)) (float32 / uint32,
 * threefry2x32, jax_threefry_partitionable=False, jax <= 0.4.2x) - the reshape to
 * [B,32,32,3] is free.  n < 2^32 - 1.  Flax's make_rng key folding stays with the caller.
 */
int mulan_rng_bits(uint32_t key0, uint32_t key1, int64_t n, uint32_t* out, void* stream);
int mulan_rng_uniform(uint32_t key0, uint32_t key1, int64_t n, float minval, float maxval,
                      float* out, void* stream);
int mulan_rng_normal(uint32_t key0, uint32_t key1, int64_t n, float* out, void* stream);

/* Frees the calling thread's cached mulan_elbo_host workspace (device + pinned host). */
void mulan_host_workspace_release(void);

#ifdef __cplusplus
}
#endif
#endif  /* MULAN_B200_H_ */
