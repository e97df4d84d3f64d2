"""Seeded inputs shared by the golden generator (make_golden.py, runs the reference source)
and the tests that replay them through the oracle / the CUDA path.  Only numpy's PCG64
streams are used, so the inputs are reproducible anywhere and the .npz fixtures only need
to hold OUTPUTS."""
from __future__ import annotations

import numpy as np

D = 3072
LATENT = 50
LATENT_K = 15

CONFIG = dict(
    vocab_size=256, sample_softmax=False, antithetic_time_sampling=True,
    with_fourier_features=True, with_attention=False, gamma_type='poly_fixedend',
    gamma_min=-13.3, gamma_max=5., sm_n_timesteps=0, sm_n_embd=128, sm_n_layer=32,
    sm_pdrop=0.1, latent_size=LATENT, encoder='unet', latent_type='topk', z_conditioning=True,
    reparam_type='true', unet_type='vdm', topk_noise_type='gamma', latent_k=LATENT_K,
    velocity_from_epsilon=False)   # values of ldm/configs/cifar10-conditioned.py:36-79


def softplus(v):
  return np.logaddexp(v, 0.0)


def glue_inputs(seed: int, B: int):
  """Inputs of the ELBO glue with (a, b, c) and the encoder logits supplied directly."""
  r = np.random.default_rng(seed)
  f32 = np.float32
  return dict(
      images=r.integers(0, 256, size=(B, 32, 32, 3), dtype=np.uint8),
      a=r.standard_normal((B, D)).astype(f32),
      b=r.standard_normal((B, D)).astype(f32),
      c=(1e-3 + softplus(r.standard_normal((B, D)))).astype(f32),
      logits=(2.0 * r.standard_normal((B, LATENT))).astype(f32),
      t0=f32(r.uniform()),
      G=r.gamma(1.0 / LATENT_K, size=(10, B, LATENT)).astype(f32),
      eps_0=r.standard_normal((B, 32, 32, 3)).astype(f32),
      eps=r.standard_normal((B, 32, 32, 3)).astype(f32),
      noise=(0.3 * r.standard_normal((B, 32, 32, 3))).astype(f32),
      w1=f32(0.7), w2=f32(0.05), w3=f32(0.02))


def mlp_weights(seed: int):
  """Flax-layout (kernel [in, out]) weights of NoiseSchedule_polynomial_fixedend's five Dense
  layers (ldm/model_mulan_epsilon.py:493-512): lecun-normal, with the zero-initialised
  `dense_out_a` perturbed so that a != 0."""
  r = np.random.default_rng(seed)
  f32 = np.float32
  def lecun(i, o, scale=1.0):
    return (scale * r.standard_normal((i, o)) / np.sqrt(i)).astype(f32)
  def bias(o):
    return (0.1 * r.standard_normal((o,))).astype(f32)
  return {
      'dense_1/kernel': lecun(LATENT, D), 'dense_1/bias': bias(D),
      'dense_2/kernel': lecun(D, D), 'dense_2/bias': bias(D),
      'dense_out_a/kernel': lecun(D, D, 3.0), 'dense_out_a/bias': bias(D),
      'dense_out_b/kernel': lecun(D, D, 3.0), 'dense_out_b/bias': bias(D),
      'dense_out_c/kernel': lecun(D, D, 3.0), 'dense_out_c/bias': bias(D),
  }


def encoder_weights(seed: int):
  """Stand-in encoder: logits = orig_f.reshape(B,-1)[:, :256] @ We  (the real UnetEncoder is
  outside the hot path)."""
  r = np.random.default_rng(seed)
  return (0.4 * r.standard_normal((256, LATENT))).astype(np.float32)


# ----------------------------------------------------------------------------------------
# Fixtures at BASELINE.json's batch sizes, the edge cases and the dense-VLB tile
# (make_golden_configs.py).  Large batches are stored COMPACTLY: per-example losses in full,
# [B, D] tensors as per-row statistics (sum, L2 norm, K seeded random projections).
# ----------------------------------------------------------------------------------------
PROJ_K = 8


def proj_vectors(seed: int = 999, k: int = PROJ_K, d: int = D):
  """K fixed N(0,1) directions: <grad_row, r_k> detects a relative L2 error eta of the row as
  a deviation ~ eta * |grad_row| in each projection."""
  return np.random.default_rng(seed).standard_normal((k, d))


def compact(v):
  """[B, ...] -> dict(sum[B], norm[B], proj[B,K]) in float64 (v flattened per row)."""
  v = np.asarray(v, np.float64).reshape(v.shape[0], -1)
  return dict(sum=v.sum(axis=1), norm=np.linalg.norm(v, axis=1), proj=v @ proj_vectors().T)


# name -> (kind, seed, B): configs[0] (B=8 eps), configs[1] (B=128 eps), configs[2] (velocity,
# 128 per GPU at 8 GPUs), configs[3] (v-from-eps, 256 per GPU)
SIZE_CASES = {
    'cfg1_eps_B8': ('eps', 301, 8),
    'cfg2_eps_B128': ('eps', 302, 128),
    'cfg3_vel_B128': ('vel', 303, 128),
    'cfg4_vfe_B256': ('vfe', 304, 256),
}

EDGE_CASES = {'edge_eps': ('eps', 401), 'edge_vel': ('vel', 402), 'edge_vfe': ('vfe', 403)}
EDGE_T = np.array([0., 1., 1e-6, .5, .999999, .25, .7, .123], np.float32)


def edge_inputs(seed: int):
  """B = 8, one edge per row (t is supplied per row: antithetic_time_sampling=False):
    0  t = 0,        a == 0 (the reference's zero-initialised dense_out_a, epsilon.py:495-500)
    1  t = 1,        a, b x 30
    2  t = 1e-6,     c -> 1e-3 (softplus of ~-20)
    3  t = .5,       x in {0, 255} with eps_0 x 4 (|eps_0| > 3: z_0 leaves its bin / the range)
    4  t = 1 - 1e-6, a == b == 0
    5  t = .25,      a x 30, b x -30, x in {0, 255}, eps_0 x 4
    6  t = .7,       a == b == 0, c == 1 (gamma exactly linear in t)
    7  t = .123,     an ordinary row
  """
  B = 8
  inp = glue_inputs(seed, B)
  r = np.random.default_rng(seed + 7)
  f32 = np.float32
  a, b, c = inp['a'].copy(), inp['b'].copy(), inp['c'].copy()
  x, e0 = inp['images'].copy(), inp['eps_0'].copy()
  a[0] = 0
  a[1] *= 30; b[1] *= 30
  c[2] = (1e-3 + softplus(-20. + r.standard_normal(D))).astype(f32)
  x[3] = r.choice(np.array([0, 255], np.uint8), size=(32, 32, 3)); e0[3] *= 4
  a[4] = 0; b[4] = 0
  a[5] *= 30; b[5] *= -30
  x[5] = r.choice(np.array([0, 255], np.uint8), size=(32, 32, 3)); e0[5] *= 4
  a[6] = 0; b[6] = 0; c[6] = 1
  inp.update(a=a, b=b, c=c, images=x, eps_0=e0, t=EDGE_T.copy())
  return inp


DENSE_CASES = {'dense_vel': ('vel', 501), 'dense_vfe': ('vfe', 502)}
DENSE_T, DENSE_IMAGES = 128, 2


def dense_inputs(seed: int):
  """eval_bpd_dense_sampling (ldm/notebook_utils.py:176-191): every image is tiled DENSE_T times
  and evaluated with THE SAME key, so the draws (t0, G, eps_0, eps) are shared by all images;
  the image, and with it the encoder logits and the schedule coefficients, differ per image."""
  base = glue_inputs(seed, DENSE_T)
  per_image = []
  for j in range(DENSE_IMAGES):
    g = glue_inputs(seed + 10 + j, DENSE_T)
    per_image.append(dict(image=g['images'][:1], a=g['a'], b=g['b'], c=g['c'],
                          logits=g['logits']))
  return base, per_image
