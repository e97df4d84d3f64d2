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
